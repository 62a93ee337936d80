"""Synthetic planted-motif data of BASELINE.json's shapes: uniform iid ACGT background, a
Dirichlet(0.3) PWM of width W, one sampled site planted on a random strand in half of the sequences, and a
binding-site block sampled from the same PWM. Everything derives from one NumPy seed that bench.py records.

Stored layout is the reference's (Sequence.cpp:10-14, 91-99): forward | 0 | reverse complement, L = 2*L0+1.
The structural N in the middle makes the 11 following k-mer hashes depend on rand() draws in the reference
(Sequence.cpp:38); here they come from the same NumPy generator and travel as the seqset's patch list.
"""
import numpy as np

WORKLOADS = {
    # name: nseq, L0, W, K (motif order), K_bg (background order)   -- BASELINE.json configs[1], configs[2]
    "c2": dict(nseq=50_000, L0=200, W=12, K=2, K_bg=2, desc="synthetic 50k x 200 bp, planted W=12, order-2 motif / order-2 background"),
    "c3": dict(nseq=1_000_000, L0=500, W=20, K=4, K_bg=2, desc="synthetic 1M x 500 bp, order-4 motif W=20, both strands"),
    # BASELINE.json configs[3]: --FDR, 5 folds, mFold=10 sampled negatives, order 3; one "step" scores one fold's test set
    "c4": dict(nseq=1_000_000, L0=500, W=12, K=3, K_bg=2, mfold=10, cvfold=5,
               desc="synthetic 1M x 500 bp positives, 10x sampled negatives (device), order-3 W=12, one fold of 5-fold FDR scoring (ZOOPS)"),
    # BASELINE.json configs[4]: multi-motif EM from the reference's shipped PWMs (example/PWM_peng10.meme, six motifs), order 5,
    # EXTENDED (ACGTMH) alphabet; runs through the drop-in CLI (bench.py --workload c5)
    "c5": dict(nseq=100_000, L0=500, W=12, K=5, K_bg=2, desc="synthetic 100k x 500 bp, six motifs of PWM_peng10.meme, order 5, EXTENDED alphabet (CLI)"),
    "tiny": dict(nseq=2_000, L0=100, W=10, K=2, K_bg=2, desc="smoke-sized planted-motif set"),
}


def planted_sequences(seed, nseq, L0, W, A=4, plant_frac=0.5, nsites=500, motif_seed=None):
    """Returns (fwd codes [nseq, L0] uint8 in 1..A, sites [nsites, W] uint8 in 1..A, pwm [W, A]).
    motif_seed: seed of the PWM and of the binding-site block (default: `seed`). Shards of ONE data set (multi-GPU runs)
    share the motif_seed and differ in `seed`, which draws the background, the planted positions and strands."""
    rng = np.random.default_rng(seed)
    mrng = rng if motif_seed is None else np.random.default_rng(motif_seed)
    pwm = mrng.dirichlet(np.full(A, 0.3), size=W)
    cdf = np.cumsum(pwm, axis=1)
    fwd = rng.integers(1, A + 1, size=(nseq, L0), dtype=np.uint8)
    nplant = int(nseq * plant_frac)
    rows = rng.permutation(nseq)[:nplant]
    u = rng.random((nplant, W))
    site = (u[:, :, None] > cdf[None, :, :]).sum(axis=2).astype(np.uint8)          # 0..A-1
    np.minimum(site, A - 1, out=site)
    rc = rng.random(nplant) < 0.5
    if A == 4:
        site_rc = (3 - site)[:, ::-1]
        site = np.where(rc[:, None], site_rc, site)
    start = rng.integers(0, L0 - W + 1, size=nplant)
    cols = start[:, None] + np.arange(W)[None, :]
    fwd[rows[:, None], cols] = site + 1
    us = (rng if motif_seed is None else mrng).random((nsites, W))
    sites = np.minimum((us[:, :, None] > cdf[None, :, :]).sum(axis=2), A - 1).astype(np.uint8) + 1
    return fwd, sites, pwm


def stored_both_strands(fwd, A=4, out=None):
    """fwd | 0 | revcomp  (complement of code c in ACGT is 5-c; reference Alphabet.cpp:12-15)."""
    assert A == 4
    nseq, L0 = fwd.shape
    L = 2 * L0 + 1
    codes = out if out is not None else np.empty((nseq, L), np.uint8)
    codes[:, :L0] = fwd
    codes[:, L0] = 0
    codes[:, L0 + 1:] = (5 - fwd)[:, ::-1]
    return codes


def middle_n_patches(codes, seed, A=4):
    """Patch list for the structural N at column L0 of every stored sequence: positions L0..L0+10 and their
    11-mer hash with an independent draw per (position, digit) for the N, as in Sequence.cpp:35-41."""
    nseq, L = codes.shape
    L0 = (L - 1) // 2
    rng = np.random.default_rng(seed + 7919)
    npos = min(11, L - L0)
    pos_local = L0 + np.arange(npos)
    kmer = np.zeros((nseq, npos), np.uint64)
    for a, i in enumerate(pos_local):
        for t in range(min(i + 1, 11)):
            c = codes[:, i - t].astype(np.int64)
            if i - t == L0:
                d = rng.integers(0, A, size=nseq)
            else:
                d = c - 1
            kmer[:, a] += (d * (A ** t)).astype(np.uint64)
    gpos = (np.arange(nseq, dtype=np.uint64)[:, None] * np.uint64(L) + pos_local[None, :].astype(np.uint64))
    return gpos.ravel(), kmer.ravel()


def full_kmers(codes, patch_pos, patch_kmer, A=4):
    """The reference's kmer_ array (11-mer hash per stored position) for a [nseq, L] code matrix — used by the
    tests / CPU baseline to feed the oracle the same data. Positions in the patch list take the patched value."""
    nseq, L = codes.shape
    d = np.where(codes == 0, 0, codes.astype(np.int64) - 1)
    kmer = np.zeros((nseq, L), np.uint64)
    for t in range(11):
        kmer[:, t:] += (d[:, :L - t] * (A ** t)).astype(np.uint64)
    flat = kmer.ravel()
    flat[patch_pos.astype(np.int64)] = patch_kmer
    return flat


def write_fasta(path, fwd, letters="ACGT"):
    lut = np.frombuffer(("N" + letters).encode(), np.uint8)
    with open(path, "wb") as f:
        for n in range(fwd.shape[0]):
            f.write(b">s%d\n" % n)
            f.write(lut[fwd[n]].tobytes())
            f.write(b"\n")


def write_sites(path, sites, letters="ACGT"):
    lut = np.frombuffer(("N" + letters).encode(), np.uint8)
    with open(path, "wb") as f:
        for s in sites:
            f.write(lut[s].tobytes() + b"\n")
