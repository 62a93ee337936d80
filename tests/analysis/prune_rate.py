"""Development analysis (CPU, uses the oracle: test infrastructure, never shipped): how many windows of a c3-like set can reach
the M-step's threshold, and how many survive an upper bound of the window product built from G1 lookup tables over base
ranges (columns whose context sticks out of their group's bases take the max of s over the missing context bases).
Decides the formulation of the pruned E-step.   usage: prune_rate.py [nseq] [alpha_div] [K] [W]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bammmotif2_b200 import synth, hostmodel
from oracle import oracle as orc

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
alpha_div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
K = int(sys.argv[3]) if len(sys.argv) > 3 else 4
W = int(sys.argv[4]) if len(sys.argv) > 4 else 20
L0, Kbg, A, Q = 500, 2, 4, 0.3
iters = [0, 3, 8]
fwd, sites, _ = synth.planted_sequences(1234, nseq, L0, W)
codes = synth.stored_both_strands(fwd)
ppos, pkmer = synth.middle_n_patches(codes, 1234)
kmer = synth.full_kmers(codes, ppos, pkmer)
L = codes.shape[1]
offsets = np.arange(nseq + 1, dtype=np.uint64) * np.uint64(L)
nb, vbg = orc.bg_model(kmer, A, Kbg, hostmodel.default_bg_alpha(Kbg))
alpha = hostmodel.default_motif_alpha(K, W)
alpha[1:] /= alpha_div
v = hostmodel.motif_from_sites(sites, A, K, alpha, vbg)
r = np.zeros(len(kmer), np.float32)
n_all = np.zeros(orc.model_size(A, K, W), np.float32)
Yn = A ** (K + 1)
ns = min(nseq, 3000)
ks = (kmer[: ns * L] % np.uint64(Yn)).astype(np.int64).reshape(ns, L)
LW1 = L - W + 1
thr0 = 2.0 ** -41 * (1 - Q) * 0.999
pos = Q / LW1
full = L - 2 * W + 2            # windows p < full are untruncated

def plan_bound(S, groups):
    """groups: list of (lo, hi) base ranges relative to the window start; column j belongs to the group that holds base p+j."""
    b = np.ones((ns, full))
    for j in range(W):
        g = max([x for x in groups if x[0] <= j <= x[1]], key=lambda x: j - x[0])
        avail = min(j - g[0] + 1, K + 1)
        U = S[:, j].reshape(A ** (K + 1 - avail), A ** avail).max(axis=0)
        b *= U[(ks % (A ** avail))[:, j:j + full]]
    return b

PLANS = {
    "pairs a: [0..5|6..11|12..17|18..19]": [(0, 5), (6, 11), (12, 17), (18, 19)],
    "pairs b: [-2..3|4..9|10..15|15..19]": [(-2, 3), (4, 9), (10, 15), (15, 19)],
    "pairs c: [-2..3|3..8|9..14|15..19]": [(-2, 3), (3, 8), (9, 14), (15, 19)],
    "pairs d: [-1..4|5..10|11..16|15..19]": [(-1, 4), (5, 10), (11, 16), (15, 19)],
    "3x7 [-1..5|6..12|13..19]": [(-1, 5), (6, 12), (13, 19)],
    "3: [-2..4|5..12(8b,16bit)|13..19]": [(-2, 4), (5, 12), (13, 19)],
    "3: 8b,8b,7b 16-bit [-2..5|6..13|13..19]": [(-2, 5), (6, 13), (13, 19)],
    "4: [-2..4|3..9|8..14|14..19]": [(-2, 4), (3, 9), (8, 14), (14, 19)],
    "4: 7,7,7,6 [-2..4|4..10|9..15|14..19]": [(-2, 4), (4, 10), (9, 15), (14, 19)],
    "5x(6-7)": [(-2, 4), (3, 8), (7, 12), (11, 16), (14, 19)],
}
for it in range(max(iters) + 1):
    s = orc.linear_s(v, vbg, A, K, Kbg, W)
    if it in iters:
        S = s.reshape(Yn, W).astype(np.float64)
        ex = np.ones((ns, full))
        for j in range(W):
            ex *= S[ks[:, j:j + full], j]
        val = ex * pos
        print("it %2d  exact>=thr0 %.4f  r>=2^-41 %.4f  r>=2^-33 %.4f" % (it, (val >= thr0).mean(),
              ((val / (1 - Q + val.sum(1, keepdims=True))) >= 2.0 ** -41).mean(), ((val / (1 - Q + val.sum(1, keepdims=True))) >= 2.0 ** -33).mean()), flush=True)
        for name, groups in PLANS.items():
            b = plan_bound(S, groups)
            assert (b >= ex * (1 - 1e-9)).all()
            print("        %-45s cand %.4f" % (name, (b * pos >= thr0).mean()), flush=True)
    orc.em_iteration_omp(kmer, offsets, A, K, W, s, Q, r, n_all, 8)
    orc.update_v(n_all, alpha.ravel(), vbg, A, K, W, v)
