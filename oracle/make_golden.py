#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference itself. TEST INFRASTRUCTURE.

Runs here (the build container), where /root/reference exists: `make -C oracle ref` compiles the
untouched reference sources into oracle/_ref/ref_dump; this script feeds it the reference's shipped
example (config 1) plus small seeded synthetic cases, with --threads 1 (the bit-reproducible mode,
SURVEY.md §4), and packs every dumped intermediate into one compressed .npz per case. The GPU box
has no /root/reference; tests only read the committed fixtures.

    python oracle/make_golden.py            # regenerates all fixtures
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("BAMM_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
ALPHABETS = {"STANDARD": "ACGT", "METHYLC": "ACGTM", "HYDROXYMETHYLC": "ACGTH", "EXTENDED": "ACGTMH"}


def synth(seed, nseq, L0, W, alphabet="STANDARD", n_frac=0.0, lower=False, plant=0.5, ragged=0):
    """Planted-motif sequences (SURVEY.md §8d generator, small): iid background, Dirichlet(0.3) PWM."""
    rng = np.random.default_rng(seed)
    letters = ALPHABETS[alphabet]
    A = len(letters)
    pwm = rng.dirichlet(np.full(A, 0.3), size=W)
    seqs = []
    for n in range(nseq):
        L = L0 + (int(rng.integers(0, ragged + 1)) if ragged else 0)
        s = rng.integers(0, A, size=L)
        if rng.random() < plant:
            site = np.array([rng.choice(A, p=pwm[j]) for j in range(W)])
            if rng.random() < 0.5 and alphabet == "STANDARD":
                site = (3 - site)[::-1]
            p = int(rng.integers(0, L - W + 1))
            s[p:p + W] = site
        txt = np.array(list(letters))[s]
        if n_frac > 0:
            mask = rng.random(L) < n_frac
            txt[mask] = "N"
        t = "".join(txt)
        if lower and n % 3 == 0:
            t = t.lower()
        seqs.append(t)
    sites = ["".join(letters[rng.choice(A, p=pwm[j])] for j in range(W)) for _ in range(200)]
    return seqs, sites


def write_inputs(tmp, seqs, sites):
    fa = os.path.join(tmp, "in.fasta")
    with open(fa, "w") as f:
        for n, s in enumerate(seqs):
            f.write(">seq%d\tdescr\n" % n)
            for i in range(0, len(s), 70):
                f.write(s[i:i + 70] + "\n")
    bs = os.path.join(tmp, "sites.block")
    with open(bs, "w") as f:
        for s in sites:
            f.write(s + "\n")
    return fa, bs


def run_case(name, fasta, sitefile, args, r_iters="1", max_iter=None, dump_neg=False, keep=None, extra_inputs=None, init=("--bindingSiteFile",),
             dump_mask=False, dump_pvalues=False):
    tmp = tempfile.mkdtemp(prefix="golden_")
    env = dict(os.environ, BAMM_DUMP_R_ITERS=r_iters, OMP_NUM_THREADS="1")
    if dump_pvalues:
        env["BAMM_DUMP_PVALUES"] = "1"
    if dump_mask:
        env["BAMM_DUMP_MASK"] = "1"
    if max_iter:
        env["BAMM_DUMP_MAXITER"] = str(max_iter)
    if dump_neg:
        env["BAMM_DUMP_NEG"] = "1"
    cmd = [os.path.join(HERE, "_ref", "ref_dump"), tmp, fasta, init[0], sitefile] + args + ["--threads", "1"]
    subprocess.check_call(cmd, env=env, stdout=subprocess.DEVNULL)
    d = os.path.join(tmp, "dump")
    arrays = {}
    meta = open(os.path.join(d, "meta.txt")).read()
    iters = None
    for line in meta.splitlines():
        t = line.split()
        if len(t) == 4 and t[0] == "motif" and t[2] == "iterations":
            iters = int(t[3])
    for fn in sorted(os.listdir(d)):
        if not fn.endswith(".npy"):
            continue
        key = fn[:-4]
        # keep per-iteration n/v only for the first two and the last iteration (fixture size)
        if "_it" in key and (key.startswith("m1_n_it") or key.startswith("m1_v_it")):
            it = int(key.split("_it")[1])
            if it not in (1, 2, iters):
                continue
        if keep is not None and not keep(key):
            continue
        arrays[key] = np.load(os.path.join(d, fn))
    arrays["meta"] = np.frombuffer(meta.encode(), np.uint8)
    arrays["args"] = np.frombuffer(" ".join(args).encode(), np.uint8)
    arrays["fasta_text"] = np.frombuffer(open(fasta, "rb").read(), np.uint8)
    arrays["sites_text"] = np.frombuffer(open(sitefile, "rb").read(), np.uint8)
    # the motif file the reference wrote (3 significant digits; file-format fixture)
    for fn in os.listdir(tmp):
        if fn.endswith(".ihbcp") or fn.endswith(".ihbp") or fn.endswith(".hbcp") or fn.endswith(".hbp") or fn.endswith(".zoops.stats"):
            arrays["file_" + fn.replace(".", "_")] = np.frombuffer(open(os.path.join(tmp, fn), "rb").read(), np.uint8)
    if extra_inputs:
        arrays.update(extra_inputs)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    shutil.rmtree(tmp)
    print("%-14s iterations=%s  %d arrays  %.1f KB" % (name, iters, len(arrays), os.path.getsize(path) / 1024))


def pwm_case():
    """Multi-motif initialisation from a MEME file (Motif::initFromPWM samples one site per sequence with a default-seeded
    std::mt19937; 1 thread makes it deterministic): the reference's shipped PWMs on its shipped sequences."""
    keep = lambda k: (k.startswith("m") and any(t in k for t in ("_v_init", "_alpha", "_llh", "_vdiff", "_v_final", "_q"))
                      and "_it" not in k) or k.startswith("bg_")
    run_case("jund_pwm_k1", os.path.join(REF, "example", "JunD.fasta"), os.path.join(REF, "example", "PWM_peng10.meme"),
             ["--EM", "-k", "1", "-K", "1", "--maxPWM", "2"], r_iters="", keep=keep, init=("--PWMFile",))


def neg_cases():
    """Negative sets only (SeqGenerator::sample_bgseqset_by_fold after srand(42), as mainBaMM.cpp:100-116 calls it): ragged
    both-strand templates with real N's, the 6-letter alphabet, single-stranded templates. One EM iteration keeps it short."""
    keep = lambda k: k in ("neg_codes", "neg_offsets", "pos_codes", "pos_offsets", "pos_kmer")
    tmp = tempfile.mkdtemp(prefix="golden_in_")
    cases = [
        ("neg_ragged_N", dict(seed=21, nseq=25, L0=34, W=6, n_frac=0.03, lower=True, ragged=7), ["--EM", "-k", "1", "-K", "1"]),
        ("neg_ext", dict(seed=22, nseq=20, L0=30, W=6, alphabet="EXTENDED"), ["--EM", "-k", "1", "-K", "1", "--alphabet", "EXTENDED"]),
        ("neg_ss", dict(seed=23, nseq=28, L0=45, W=6), ["--EM", "-k", "1", "-K", "1", "--ss"]),
    ]
    for name, kw, args in cases:
        seqs, sites = synth(**kw)
        d = os.path.join(tmp, name)
        os.makedirs(d)
        fa, bs = write_inputs(d, seqs, sites)
        run_case(name, fa, bs, args, r_iters="", max_iter=1, dump_neg=True, keep=keep)
    shutil.rmtree(tmp)


def mask_cases():
    """EM::mask (--advanceEM) from the binding-site initial model: final model, r, counts and log likelihood of the reference."""
    keep = lambda k: k.startswith("pos_") or k.startswith("bg_") or k in ("m1_v_init", "m1_alpha") or "_mask_" in k
    tmp = tempfile.mkdtemp(prefix="golden_in_")
    cases = [
        ("mask_k2", dict(seed=31, nseq=60, L0=70, W=8), ["--EM", "-k", "2", "-K", "2"]),
        ("mask_k3_ss", dict(seed=32, nseq=50, L0=90, W=10, n_frac=0.02), ["--EM", "-k", "3", "-K", "2", "--ss", "-q", "0.4"]),
    ]
    for name, kw, args in cases:
        seqs, sites = synth(**kw)
        d = os.path.join(tmp, name)
        os.makedirs(d)
        fa, bs = write_inputs(d, seqs, sites)
        run_case(name, fa, bs, args, r_iters="", max_iter=1, keep=keep, dump_mask=True)
    shutil.rmtree(tmp)


def pvalue_case():
    """Score statistics: ScoreSeqSet::calcPvalues of every positive window against all negative window
    scores (what --scoreSeqset computes, mainBaMM.cpp:203-230) and FDR::calculatePvalues with --savePvalues, MOPS and ZOOPS."""
    keep = lambda k: "_pval_" in k or ("_fdr_" in k and "All" not in k and "mops" not in k) or k in ("m1_score_mops", "pos_offsets")
    tmp = tempfile.mkdtemp(prefix="golden_in_")
    seqs, sites = synth(seed=41, nseq=50, L0=24, W=7)
    d = os.path.join(tmp, "pval")
    os.makedirs(d)
    fa, bs = write_inputs(d, seqs, sites)
    run_case("syn_pval", fa, bs, ["--EM", "-k", "2", "-K", "2", "--FDR", "-n", "4", "--savePvalues"], r_iters="",
             keep=keep, dump_pvalues=True)
    shutil.rmtree(tmp)


def main():
    subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    if "--only-pval" in sys.argv:
        return pvalue_case()
    if "--only-mask" in sys.argv:
        return mask_cases()
    if "--only-pwm" in sys.argv:
        return pwm_case()
    if "--only-neg" in sys.argv:
        return neg_cases()
    # config 1: the reference's shipped example
    run_case("jund_k2", os.path.join(REF, "example", "JunD.fasta"), os.path.join(REF, "example", "bindingsites.block"),
             ["--EM", "-k", "2", "-K", "2", "--FDR"], r_iters="1,41",
             keep=lambda k: not k.startswith("neg_"))
    tmp = tempfile.mkdtemp(prefix="golden_in_")
    cases = [
        # name, synth kwargs, args, r_iters, dump_neg
        ("syn_k2_N", dict(seed=11, nseq=40, L0=60, W=8, n_frac=0.03, lower=True, ragged=9), ["--EM", "-k", "2", "-K", "2"], "1,2", False),
        ("syn_ss_k1_q", dict(seed=12, nseq=50, L0=80, W=9), ["--EM", "-k", "1", "-K", "1", "--ss", "--optimizeQ", "-q", "0.5"], "1,3", False),
        ("syn_k4", dict(seed=13, nseq=60, L0=120, W=12), ["--EM", "-k", "4", "-K", "2"], "1", False),
        ("syn_ext_k1", dict(seed=14, nseq=40, L0=50, W=6, alphabet="EXTENDED"), ["--EM", "-k", "1", "-K", "1", "--alphabet", "EXTENDED"], "1", False),
        ("syn_k3_fdr", dict(seed=15, nseq=50, L0=50, W=8), ["--EM", "-k", "3", "-K", "2", "--FDR", "-n", "5"], "1", True),
        ("syn_k0", dict(seed=16, nseq=30, L0=40, W=7), ["--EM", "-k", "0", "-K", "0"], "1", False),
    ]
    for name, kw, args, r_iters, dump_neg in cases:
        seqs, sites = synth(**kw)
        d = os.path.join(tmp, name)
        os.makedirs(d)
        fa, bs = write_inputs(d, seqs, sites)
        run_case(name, fa, bs, args, r_iters=r_iters, dump_neg=dump_neg)
    shutil.rmtree(tmp)
    pwm_case()
    neg_cases()
    mask_cases()
    pvalue_case()


if __name__ == "__main__":
    sys.exit(main())
