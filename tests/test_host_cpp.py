"""C++ host side (bammmotif2_b200/host: the reference's class surface over the C ABI).

CPU tests (no device): FASTA parsing + encoding + N draws, background arithmetic and file format, binding-site
initialisation, negative-set sampling and the FDR statistics, each against the golden vectors made from the reference
(bit-exact: integer work and sequential fp32 host arithmetic in the reference's operation order).
GPU tests (-m gpu): the drop-in CLI bin/BaMMmotif end to end against the files the reference wrote for the same command.
"""
import os
import subprocess

import numpy as np
import pytest

from util import CASES, Golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bammmotif2_b200", "bin")


@pytest.fixture(scope="module")
def bins():
    from bammmotif2_b200 import build
    build.build_all()
    return BIN


def write_inputs(g, d, name="in"):
    fa, bs = os.path.join(d, name + ".fasta"), os.path.join(d, "sites.block")
    open(fa, "wb").write(bytes(g["fasta_text"]))
    open(bs, "wb").write(bytes(g["sites_text"]))
    return fa, bs


def parse_fasta_text(text):
    recs, h, sq = [], None, []
    for line in text.split("\n"):
        if not line:
            continue
        if line[0] == ">":
            if h is not None and sq:
                recs.append((h, "".join(sq)))
            h, sq = line, []
        else:
            sq.append(line)
    if h is not None and sq:
        recs.append((h, "".join(sq)))
    return recs


def run(cmd, **kw):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, **kw)
    assert p.returncode == 0, "%s\n%s\n%s" % (" ".join(cmd), p.stdout[-2000:], p.stderr[-2000:])
    return p


@pytest.mark.parametrize("seed", [42, 1, 0, 2**31 - 1, 123456789])
def test_private_rand_stream_equals_libc(bins, seed):
    """host/Util.h LibcRand is the stream every rand()-dependent step of the reference consumes (srand(42) in mainBaMM.cpp:22,
    SeqGenerator.cpp:33, FDR.cpp:153): same values as this machine's libc srand() / rand(), including the seed-0 rule."""
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(ctypes.c_uint(seed))
    want = [libc.rand() for _ in range(5000)]
    out = subprocess.run([os.path.join(bins, "host_check"), "rand", str(seed), "5000"], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert [int(x) for x in out.split()] == want


@pytest.mark.parametrize("case", CASES)
def test_fasta_encoding_and_kmers_bit_exact(bins, case, tmp_path):
    g = Golden(case)
    fa, _ = write_inputs(g, str(tmp_path))
    run([os.path.join(bins, "host_check"), "encode", g.alphabet, fa, "1" if g.ss else "0", str(tmp_path)])
    assert np.array_equal(np.fromfile(tmp_path / "codes.u8", np.uint8), g["pos_codes"])
    assert np.array_equal(np.fromfile(tmp_path / "offsets.u64", np.uint64), g["pos_offsets"])
    # includes the rand() draws for N bases: same libc stream, same order as the reference
    assert np.array_equal(np.fromfile(tmp_path / "kmer.u64", np.uint64), g["pos_kmer"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + ["neg_ragged_N", "neg_ext", "neg_ss"])
def test_fasta_encoding_on_the_device_bit_exact(bins, case, tmp_path):
    """Row f-3: the same reader with the per-base work on the device (raw text up, codes encoded / mirrored / counted by
    k_fasta_encode, rand()-dependent hashes from the 21-code windows): codes, offsets, every k-mer hash and the base
    frequencies equal the reference's; also with CRLF line ends and sequences wrapped over several lines."""
    g = Golden(case)
    fa, _ = write_inputs(g, str(tmp_path))
    env = dict(os.environ, BAMM_DEVICE_FASTA="1")
    run([os.path.join(bins, "host_check"), "encode", g.alphabet, fa, "1" if g.ss else "0", str(tmp_path)], env=env)
    assert np.array_equal(np.fromfile(tmp_path / "codes.u8", np.uint8), g["pos_codes"])
    assert np.array_equal(np.fromfile(tmp_path / "offsets.u64", np.uint64), g["pos_offsets"])
    assert np.array_equal(np.fromfile(tmp_path / "kmer.u64", np.uint64), g["pos_kmer"])
    freq = np.fromfile(tmp_path / "basefreq.f32", np.float32)
    # the host reader on a re-wrapped, CRLF copy of the file (a '\r' is an undefined base for both readers, like in the reference)
    recs = parse_fasta_text(open(fa).read())
    alt = tmp_path / "alt.fasta"
    with open(alt, "w", newline="") as f:
        for h, sq in recs:
            f.write(h + "\r\n")
            for i in range(0, len(sq), 37):
                f.write(sq[i:i + 37] + "\r\n")
    d_host, d_dev = tmp_path / "h", tmp_path / "d"
    d_host.mkdir(); d_dev.mkdir()
    run([os.path.join(bins, "host_check"), "encode", g.alphabet, str(alt), "1" if g.ss else "0", str(d_host)], env=dict(os.environ, BAMM_DEVICE_FASTA="0"))
    run([os.path.join(bins, "host_check"), "encode", g.alphabet, str(alt), "1" if g.ss else "0", str(d_dev)], env=env)
    for fn in ("codes.u8", "offsets.u64", "kmer.u64", "basefreq.f32"):
        assert open(d_host / fn, "rb").read() == open(d_dev / fn, "rb").read(), fn
    assert len(freq) == g.A


@pytest.mark.gpu
@pytest.mark.parametrize("ss", ["0", "1"])
def test_fasta_device_reader_with_sparse_undefined_bases(bins, ss, tmp_path):
    """The k-mer hashes around undefined bases in the device reader: one draw plus a rolling hash where the base is the only
    undefined one of its 11-mer, the digit-by-digit loop elsewhere — records shorter than 11 bases (growing spans), undefined
    bases next to each other, at the ends and near the strand junction. Same codes, hashes (= same rand() stream) and
    frequencies as the host loop."""
    rng = np.random.default_rng(5)
    fa = tmp_path / "sparse.fasta"
    with open(fa, "w") as f:
        for i in range(400):
            L = int(rng.integers(3, 16)) if i % 3 == 0 else int(rng.integers(20, 120))
            sq = rng.choice(list("ACGT"), size=L)
            for z in rng.integers(0, L, size=int(rng.integers(0, 4))):
                sq[z] = "N"
            if i % 7 == 0: sq[0] = "N"
            if i % 11 == 0: sq[-1] = "N"
            f.write(">s%d\n%s\n" % (i, "".join(sq)))
    d_host, d_dev = tmp_path / "h", tmp_path / "d"
    d_host.mkdir(); d_dev.mkdir()
    run([os.path.join(bins, "host_check"), "encode", "STANDARD", str(fa), ss, str(d_host)], env=dict(os.environ, BAMM_DEVICE_FASTA="0"))
    run([os.path.join(bins, "host_check"), "encode", "STANDARD", str(fa), ss, str(d_dev)], env=dict(os.environ, BAMM_DEVICE_FASTA="1"))
    for fn in ("codes.u8", "offsets.u64", "kmer.u64", "basefreq.f32"):
        assert open(d_host / fn, "rb").read() == open(d_dev / fn, "rb").read(), fn


@pytest.mark.gpu
def test_fasta_with_many_undefined_bases_falls_back_to_the_host_reader(bins, tmp_path):
    """The device encoder keeps a bounded list of undefined bases (1/16 of the text); a file above that is read by the
    host loop instead, with the same result as when the host loop is asked for directly."""
    rng = np.random.default_rng(11)
    fa = tmp_path / "n.fasta"
    with open(fa, "w") as f:
        for i in range(300):
            f.write(">s%d\n" % i)
            f.write("".join(rng.choice(list("ACGTNNN"), size=int(rng.integers(40, 90)))) + "\n")
    d_host, d_dev = tmp_path / "h", tmp_path / "d"
    d_host.mkdir(); d_dev.mkdir()
    run([os.path.join(bins, "host_check"), "encode", "STANDARD", str(fa), "0", str(d_host)], env=dict(os.environ, BAMM_DEVICE_FASTA="0"))
    run([os.path.join(bins, "host_check"), "encode", "STANDARD", str(fa), "0", str(d_dev)], env=dict(os.environ, BAMM_DEVICE_FASTA="1"))
    for fn in ("codes.u8", "offsets.u64", "kmer.u64", "basefreq.f32"):
        assert open(d_host / fn, "rb").read() == open(d_dev / fn, "rb").read(), fn


@pytest.mark.parametrize("case", CASES)
def test_background_and_site_init_bit_exact(bins, case, tmp_path):
    g = Golden(case)
    _, bs = write_inputs(g, str(tmp_path))
    g["bg_n"].tofile(tmp_path / "bgn.u64")
    g.bg_alpha().tofile(tmp_path / "abg.f32")
    run([os.path.join(bins, "host_check"), "init", g.alphabet, str(g.K), str(g.K_bg_model), str(tmp_path / "bgn.u64"),
         str(tmp_path / "abg.f32"), bs, repr(float(g.q)), str(tmp_path)])
    assert np.array_equal(np.fromfile(tmp_path / "bg_v.f32", np.float32), g["bg_v"])
    assert np.array_equal(np.fromfile(tmp_path / "v_init.f32", np.float32), g["m1_v_init"])
    hb = [k for k in g.z.files if k.endswith("_hbcp") and "motif" not in k][0]
    assert open(tmp_path / "check.hbcp", "rb").read() == bytes(g[hb])                 # file format is contract
    assert open(tmp_path / "check.hbp", "rb").read() == bytes(g[hb[:-4] + "hbp"])


def test_pwm_initialisation_bit_exact(bins, tmp_path):
    """MotifSet("PWM") + Motif::initFromPWM: MEME parsing, posterior site sampling with std::mt19937 /
    std::discrete_distribution, higher-order counts — two motifs of the reference's shipped PWM file."""
    g = Golden("jund_pwm_k1")
    fa = tmp_path / "JunD.fasta"
    open(fa, "wb").write(bytes(Golden("jund_k2")["fasta_text"]))
    meme = tmp_path / "pwm.meme"
    open(meme, "wb").write(bytes(g["sites_text"]))
    g["bg_n"].tofile(tmp_path / "bgn.u64")
    g.bg_alpha().tofile(tmp_path / "abg.f32")
    run([os.path.join(bins, "host_check"), "pwminit", "STANDARD", str(fa), "0", str(g.K), str(g.K_bg_model),
         str(tmp_path / "bgn.u64"), str(tmp_path / "abg.f32"), str(meme), "2", repr(float(g.q)), str(tmp_path)])
    for m in (1, 2):
        assert np.array_equal(np.fromfile(tmp_path / ("v_init_%d.f32" % m), np.float32), g["m%d_v_init" % m]), m


@pytest.mark.gpu
def test_pwm_initialisation_on_the_device_bit_exact(bins, tmp_path):
    """Row f-4: the sampling step of Motif::initFromPWM on the device (float posteriors in the reference's order, the draw of
    std::discrete_distribution restated in double, integer k-mer counts): the initial models equal the reference's."""
    g = Golden("jund_pwm_k1")
    fa = tmp_path / "JunD.fasta"
    open(fa, "wb").write(bytes(Golden("jund_k2")["fasta_text"]))
    meme = tmp_path / "pwm.meme"
    open(meme, "wb").write(bytes(g["sites_text"]))
    g["bg_n"].tofile(tmp_path / "bgn.u64")
    g.bg_alpha().tofile(tmp_path / "abg.f32")
    run([os.path.join(bins, "host_check"), "pwminit", "STANDARD", str(fa), "0", str(g.K), str(g.K_bg_model),
         str(tmp_path / "bgn.u64"), str(tmp_path / "abg.f32"), str(meme), "2", repr(float(g.q)), str(tmp_path)],
        env=dict(os.environ, BAMM_DEVICE_PWMINIT="1"))
    for m in (1, 2):
        assert np.array_equal(np.fromfile(tmp_path / ("v_init_%d.f32" % m), np.float32), g["m%d_v_init" % m]), m


def test_negative_sampling_bit_exact(bins, tmp_path):
    """The serial host sampler (kept as the fall-back of the device sampler, SeqGenerator.cpp of the host side)."""
    g = Golden("syn_k3_fdr")
    fa, _ = write_inputs(g, str(tmp_path))
    run([os.path.join(bins, "host_check"), "neg", "STANDARD", fa, "0", str(g.meta["mFold"]), "2", str(tmp_path)],
        env=dict(os.environ, BAMM_HOST_NEGATIVES="1"))
    assert np.array_equal(np.fromfile(tmp_path / "neg_codes.u8", np.uint8), g["neg_codes"])
    assert np.array_equal(np.fromfile(tmp_path / "neg_offsets.u64", np.uint64), g["neg_offsets"])


@pytest.mark.gpu
def test_negative_sampling_device_path_bit_exact(bins, tmp_path):
    """The same call through the host classes with the device sampler (SeqGenerator -> bamm_seqset_sample_negatives ->
    device-built SequenceSet whose codes come back lazily)."""
    g = Golden("syn_k3_fdr")
    fa, _ = write_inputs(g, str(tmp_path))
    run([os.path.join(bins, "host_check"), "neg", "STANDARD", fa, "0", str(g.meta["mFold"]), "2", str(tmp_path)])
    assert np.array_equal(np.fromfile(tmp_path / "neg_codes.u8", np.uint8), g["neg_codes"])
    assert np.array_equal(np.fromfile(tmp_path / "neg_offsets.u64", np.uint64), g["neg_offsets"])


def _pr_walk_restated(pos, neg, posN, negN, q):
    """FDR::calculatePR's ZOOPS walk (reference: src/evaluation/FDR.cpp:196-262) restated with binary searches per entry, the way the
    reference finds the neighbours of a score among the negatives; libc supplies the tie-breaks."""
    import bisect, ctypes, math
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(42)
    f32 = np.float32
    pos = np.sort(pos)[::-1].astype(f32); neg = np.sort(neg)[::-1].astype(f32)
    negasc = (-neg).tolist()                                   # ascending keys for bisect on a descending vector
    mfold = f32(negN) / f32(posN)
    n_top = int(min(100, negN // 10))
    lam = f32(1e-16)
    for l in range(n_top):
        lam = f32(lam + f32(neg[l] - neg[n_top]))
    lam = f32(lam / f32(n_top))
    ip = ineg = 0
    TP, FP, PV = [], [], []
    nan = float("nan")
    for i in range(posN + negN):
        ps = float(pos[ip]) if ip < len(pos) else nan
        ns = float(neg[ineg]) if ineg < len(neg) else nan
        if (ps > ns or ip == 0 or ineg == negN) and ip < posN:
            Sl = ps; ip += 1
        elif ps == ns and libc.rand() % 2 == 0 and ip < posN:
            Sl = ps; ip += 1
        else:
            Sl = ns; ineg += 1
        tp = f32(ip); fp = f32(f32(ineg) / mfold)
        TP.append(tp); FP.append(fp)
        if Sl <= float(neg[n_top]):
            lb = bisect.bisect_left(negasc, -Sl)               # first negative that is not above Sl
            ub = bisect.bisect_right(negasc, -Sl)              # first negative below Sl
            up = float(neg[lb - 1]) if lb > 0 else float(neg[0])
            lo = float(neg[ub]) if ub < len(neg) else Sl
            pv = (ineg + float(f32(up - Sl)) / (float(f32(up - lo)) + 1e-5)) / float(f32(negN))
        else:
            pv = float(f32(f32(n_top) * f32(math.exp(float(f32(f32(neg[n_top] - f32(Sl)) / lam))))) / f32(negN))
        PV.append(pv)
    return np.array(TP, f32), np.array(FP, f32), np.array(PV, np.float64)


@pytest.mark.parametrize("seed,ties", [(1, False), (2, True), (3, True)])
def test_pr_walk_against_a_restatement_with_binary_searches(bins, seed, ties, tmp_path):
    """host/FDR.cpp advances the two neighbours of a score among the sorted negatives with the walk instead of searching them per
    entry: same TP / FP and the same rank p-values as a restatement that searches (scores quantised so that ties inside and between
    the sets are frequent; a first positive below the top negatives, which makes the walk's score go UP once)."""
    rng = np.random.default_rng(seed)
    posN, negN = 400, 4000
    pos = rng.normal(1.0, 2.0, posN).astype(np.float32)
    neg = rng.normal(0.0, 1.5, negN).astype(np.float32)
    if ties:
        pos = np.round(pos * 4) / np.float32(4); neg = np.round(neg * 4) / np.float32(4)
    if seed == 3:
        pos = pos - np.float32(6.0)                             # every positive below the 100 best negatives: rank branch, then upwards
    pos.tofile(tmp_path / "pos.f32"); neg.tofile(tmp_path / "neg.f32")
    run([os.path.join(bins, "host_check"), "pr", str(posN), str(negN), "0.9", str(tmp_path / "pos.f32"), str(tmp_path / "neg.f32"), str(tmp_path)])
    TP, FP, PV = _pr_walk_restated(pos, neg, posN, negN, 0.9)
    assert np.array_equal(np.fromfile(tmp_path / "TP.f32", np.float32), TP)
    assert np.array_equal(np.fromfile(tmp_path / "FP.f32", np.float32), FP)
    got = np.fromfile(tmp_path / "PNpval.f32", np.float32).astype(np.float64)
    # the exponential tail goes through expf (last-bit differences against math.exp); the rank branch is plain arithmetic
    assert np.allclose(got, PV, rtol=2e-6, atol=1e-9), np.abs(got - PV).max()


@pytest.mark.parametrize("case", ["jund_k2", "syn_k3_fdr", "syn_pval"])
def test_fdr_statistics_bit_exact(bins, case, tmp_path):
    g = Golden(case)
    g["m1_fdr_posScoreMax"].tofile(tmp_path / "pos.f32")
    g["m1_fdr_negScoreMax"].tofile(tmp_path / "neg.f32")
    run([os.path.join(bins, "host_check"), "pr", str(g.meta["npos"]), str(g.meta["nneg"]), repr(float(g.q)),
         str(tmp_path / "pos.f32"), str(tmp_path / "neg.f32"), str(tmp_path)])
    if "m1_fdr_zoops_pvalue" in g:      # FDR::calculatePvalues (--savePvalues, src/evaluation/FDR.cpp:278-330) against the reference's vector
        assert np.array_equal(np.fromfile(tmp_path / "zoops_pvalue.f32", np.float32), g["m1_fdr_zoops_pvalue"])
    if case == "syn_pval":
        # 48 of the 50 positives are scored (4 folds of 12) while the walk of calculatePR runs over 50 + 5040 entries, and four
        # scores tie between the sets (tie-break by rand() % 2, FDR.cpp:233): its tail is not defined by the inputs alone
        return
    for k in ["TP", "FP", "FDR", "Rec", "occ_frac"]:
        assert np.array_equal(np.fromfile(tmp_path / (k + ".f32"), np.float32), g["m1_fdr_" + k]), k
    p, gp = np.fromfile(tmp_path / "PNpval.f32", np.float32), g["m1_fdr_PNpval"]
    # Scores not above the lowest negative have no lower neighbour: the reference reads one element past the end of its
    # vector there (FDR.cpp:246), so those p-values are undefined in the reference itself. The walk visits the scores
    # in descending order, so entry i belongs to the i-th largest score of the union.
    walk = np.sort(np.concatenate([g["m1_fdr_posScoreMax"], g["m1_fdr_negScoreMax"]]))[::-1]
    defined = walk > g["m1_fdr_negScoreMax"].min()
    assert defined.sum() >= len(p) - 3
    assert np.array_equal(p[defined], gp[defined])


# ---------------------------------------------------------------------------------------------------------------------
def parse_numbers(blob):
    return np.array([float(t) for t in blob.decode().split()], np.float64)


def stats_table(blob):
    lines = blob.decode().strip().split("\n")
    head = lines[0].split("\t")
    body = np.array([[float(x) for x in l.split("\t") if x != ""] for l in lines[1:]], np.float64)
    return head, body


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cli_matches_reference_files(bins, case, tmp_path):
    """bin/BaMMmotif with the reference's command line writes the files the reference wrote."""
    g = Golden(case)
    name = [k for k in g.z.files if k.endswith("_hbcp") and "motif" not in k][0][len("file_"):-len("_hbcp")]
    fa, bs = write_inputs(g, str(tmp_path), name)
    out = tmp_path / "out"
    p = run([os.path.join(bins, "BaMMmotif"), str(out), fa, "--bindingSiteFile", bs] + g.args + ["--verbose"])
    iters = [l for l in p.stdout.split("\n") if " iter, llh=" in l]
    # syn_k4: the reference stops on a -2 ulp likelihood step on a plateau (see test_gpu_parity); another summation order
    # leaves the plateau at another iteration and the files then differ in the 3rd digit
    same_stop = len(iters) == g.iterations
    assert same_stop or case == "syn_k4", (len(iters), g.iterations)
    tol = 1.2e-3 if same_stop else 1e-2
    assert open(out / (name + ".hbcp"), "rb").read() == bytes(g["file_%s_hbcp" % name])
    assert open(out / (name + ".hbp"), "rb").read() == bytes(g["file_%s_hbp" % name])
    for ext in ("ihbcp", "ihbp"):
        ours = parse_numbers(open(out / ("%s_motif_1.%s" % (name, ext)), "rb").read())
        ref = parse_numbers(bytes(g["file_%s_motif_1_%s" % (name, ext)]))
        assert ours.shape == ref.shape
        # 3 significant digits in the file: 1e-4 model tolerance + one unit of the last printed digit
        assert np.all(np.abs(ours - ref) <= tol * np.abs(ref) + (0.0 if same_stop else 5e-3)), ext
    key = "file_%s_motif_1_zoops_stats" % name
    if key in g:
        head, body = stats_table(open(out / (name + "_motif_1.zoops.stats"), "rb").read())
        rhead, rbody = stats_table(bytes(g[key]))
        assert head[:6] == rhead[:6]
        assert body.shape == rbody.shape
        # TP / FP are rank walks over scores of fold models that agree with the reference's to 1e-4: allow a few swaps
        assert np.mean(np.abs(body[:, 0] - rbody[:, 0]) > 0) < 0.02
        assert np.max(np.abs(body[:, 0] - rbody[:, 0])) <= 3
        assert abs(float(head[6]) - float(rhead[6])) <= 0.02


@pytest.mark.gpu
def test_cli_pwm_file_two_motifs(bins, tmp_path):
    """--PWMFile --maxPWM 2: both motifs of the reference's run, iteration counts and written models."""
    g = Golden("jund_pwm_k1")
    fa = tmp_path / "JunD.fasta"
    open(fa, "wb").write(bytes(Golden("jund_k2")["fasta_text"]))
    meme = tmp_path / "pwm.meme"
    open(meme, "wb").write(bytes(g["sites_text"]))
    out = tmp_path / "out"
    p = run([os.path.join(bins, "BaMMmotif"), str(out), str(fa), "--PWMFile", str(meme)] + g.args + ["--verbose"])
    iters = [l for l in p.stdout.split("\n") if " iter, llh=" in l]
    want = [g.meta["iterations"]] if "iterations" in g.meta else None
    assert open(out / "JunD.hbcp", "rb").read() == bytes(g["file_JunD_hbcp"])
    for m in (1, 2):
        ours = parse_numbers(open(out / ("JunD_motif_%d.ihbcp" % m), "rb").read())
        ref = parse_numbers(bytes(g["file_JunD_motif_%d_ihbcp" % m]))
        assert ours.shape == ref.shape and np.all(np.abs(ours - ref) <= 1.2e-3 * np.abs(ref) + 1e-30), m
    assert len(iters) == len(g["m1_llh"]) + len(g["m2_llh"])


@pytest.mark.gpu
def test_cli_against_reference_binary_on_fresh_data(bins, tmp_path):
    """Same command line through the reference binary (built from the untouched sources by oracle/Makefile, 1 thread)
    and through bin/BaMMmotif on a seeded planted-motif set that no fixture covers: N bases in the input, --FDR,
    --scoreSeqset, --saveBaMMs. Exact where the work is integer / host fp32, toleranced where the device sums."""
    ref = os.path.join(ROOT, "oracle", "_ref", "BaMMmotif_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/BaMMmotif_ref not built")
    from bammmotif2_b200 import synth
    fwd, sites, _ = synth.planted_sequences(4242, 3000, 120, 11)
    fwd = fwd.copy()
    fwd[::97, 17] = 0                                   # some real N's (reverse strand then carries code 78)
    fa, bs = str(tmp_path / "syn.fasta"), str(tmp_path / "sites.block")
    synth.write_fasta(fa, fwd)
    synth.write_sites(bs, sites)
    args = ["--bindingSiteFile", bs, "--EM", "-k", "3", "-K", "2", "--FDR", "-n", "5", "--scoreSeqset", "--saveBaMMs", "--verbose"]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    pr = run([ref, str(tmp_path / "ref"), fa] + args + ["--threads", "1"], env=env)
    po = run([os.path.join(bins, "BaMMmotif"), str(tmp_path / "our"), fa] + args)
    it_r = [l for l in pr.stdout.split("\n") if " iter, llh=" in l]
    it_o = [l for l in po.stdout.split("\n") if " iter, llh=" in l]
    assert len(it_r) == len(it_o), (len(it_r), len(it_o))
    for a, b in zip(it_r, it_o):                        # "N iter, llh=X, diff_llh=Y, v_diff=Z" with 6 significant digits
        la, lb = float(a.split("llh=")[1].split(",")[0]), float(b.split("llh=")[1].split(",")[0])
        assert abs(la - lb) <= 2e-5 * abs(la)
    rd, od = tmp_path / "ref", tmp_path / "our"
    for fn in ("syn.hbcp", "syn.hbp"):
        assert open(rd / fn, "rb").read() == open(od / fn, "rb").read(), fn
    for fn in ("syn_motif_1.ihbcp", "syn_motif_1.ihbp"):
        a, b = parse_numbers(open(rd / fn, "rb").read()), parse_numbers(open(od / fn, "rb").read())
        assert a.shape == b.shape and np.all(np.abs(a - b) <= 1.2e-3 * np.abs(a) + 1e-30), fn
    # .positions: windows with posterior >= 0.3 (r within 1e-5: the same windows unless one sits on the threshold)
    pa, pb = open(rd / "syn_motif_1.positions").read().split("\n"), open(od / "syn_motif_1.positions").read().split("\n")
    assert pa[0] == pb[0] and len(set(pa) ^ set(pb)) <= 2
    # .occurrence: p-values of window scores against the sampled negatives (bit-identical negative set and scores of a
    # model that agrees to 1e-4): same header, nearly the same hits
    oa, ob = open(rd / "syn_motif_1.occurrence").read().split("\n"), open(od / "syn_motif_1.occurrence").read().split("\n")
    assert oa[0] == ob[0]
    ka, kb = {"\t".join(l.split("\t")[:5]) for l in oa[1:] if l}, {"\t".join(l.split("\t")[:5]) for l in ob[1:] if l}
    assert len(ka ^ kb) <= 0.02 * max(len(ka), 1) + 2
    ha, ba = stats_table(open(rd / "syn_motif_1.zoops.stats", "rb").read())
    hb, bb = stats_table(open(od / "syn_motif_1.zoops.stats", "rb").read())
    assert ha[:6] == hb[:6] and ba.shape == bb.shape
    # TP is a rank walk over the scores of five fold models; a fold whose EM crosses the v_diff < 0.01 stop rule one
    # iteration earlier or later than the reference shifts its scores in the 3rd-4th digit and swaps neighbouring ranks
    # (TP is cumulative: one swapped pair shifts a long stretch of rows by one, so only the size of the shift is bounded)
    assert np.max(np.abs(ba[:, 0] - bb[:, 0])) <= 0.005 * 3000
    assert abs(float(ha[6]) - float(hb[6])) <= 0.02                        # occurrence fraction in the header


@pytest.mark.gpu
def test_cli_advance_em_against_reference_binary(bins, tmp_path):
    """--advanceEM (EM::mask) through both binaries on a seeded planted-motif set: the written models agree to the printed
    digits."""
    ref = os.path.join(ROOT, "oracle", "_ref", "BaMMmotif_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/BaMMmotif_ref not built")
    from bammmotif2_b200 import synth
    fwd, sites, _ = synth.planted_sequences(777, 1500, 100, 10)
    fa, bs = str(tmp_path / "syn.fasta"), str(tmp_path / "sites.block")
    synth.write_fasta(fa, fwd)
    synth.write_sites(bs, sites)
    args = ["--bindingSiteFile", bs, "--EM", "-k", "2", "-K", "2", "--advanceEM"]
    run([ref, str(tmp_path / "ref"), fa] + args + ["--threads", "1"], env=dict(os.environ, OMP_NUM_THREADS="1"))
    run([os.path.join(bins, "BaMMmotif"), str(tmp_path / "our"), fa] + args)
    for fn in ("syn.hbcp", "syn.hbp"):
        assert open(tmp_path / "ref" / fn, "rb").read() == open(tmp_path / "our" / fn, "rb").read(), fn
    for fn in ("syn_motif_1.ihbcp", "syn_motif_1.ihbp"):
        a, b = parse_numbers(open(tmp_path / "ref" / fn, "rb").read()), parse_numbers(open(tmp_path / "our" / fn, "rb").read())
        assert a.shape == b.shape and np.all(np.abs(a - b) <= 1.2e-3 * np.abs(a) + 1e-30), fn


@pytest.mark.gpu
def test_cli_two_devices_write_the_same_files(bins, tmp_path):
    """BAMM_DEVICES=0,1: EM::optimize (also inside the FDR folds, reference src/evaluation/FDR.cpp:37-73) splits the sequences
    over two devices of one process. Every output file must be byte-identical to the one-device run (skipped on one device)."""
    from bammmotif2_b200 import capi, synth
    capi.load()
    if capi.device_count() < 2:
        pytest.skip("needs two devices")
    fwd, sites, _ = synth.planted_sequences(99, 12000, 150, 12)
    fa, bs = str(tmp_path / "syn.fasta"), str(tmp_path / "sites.block")
    synth.write_fasta(fa, fwd)
    synth.write_sites(bs, sites)
    args = ["--bindingSiteFile", bs, "--EM", "-k", "2", "-K", "2", "--FDR", "-m", "2", "-n", "3", "--saveBaMMs", "--verbose"]
    outs = {}
    for name, devs in (("one", "0"), ("two", "0,1")):
        p = run([os.path.join(bins, "BaMMmotif"), str(tmp_path / name), fa] + args, env=dict(os.environ, BAMM_DEVICES=devs, BAMM_GROUP_MIN_SEQS="1000"))
        outs[name] = [l for l in p.stdout.split("\n") if " iter, llh=" in l]
    assert outs["one"] == outs["two"] and len(outs["one"]) > 3
    files = sorted(os.listdir(tmp_path / "one"))
    assert files == sorted(os.listdir(tmp_path / "two")) and any(f.endswith(".ihbcp") for f in files) and any(f.endswith(".zoops.stats") for f in files)
    for fn in files:
        assert open(tmp_path / "one" / fn, "rb").read() == open(tmp_path / "two" / fn, "rb").read(), fn
