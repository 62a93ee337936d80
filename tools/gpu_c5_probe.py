"""One-off probe of BASELINE config 5's shape on the generic (index-array) path: 6-letter alphabet, order 5, W in {12, 9, 13}.
Prints ms per EM iteration and positions.iter/s; not a bench line (parity for this path: tests/test_gpu_parity.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bammmotif2_b200 import capi, hostmodel
A, K, Kbg = 6, int(os.environ.get("K", 5)), 2
nseq, L0 = int(os.environ.get("NSEQ", 100000)), 500
rng = np.random.default_rng(3)
fwd = rng.integers(1, A + 1, size=(nseq, L0), dtype=np.uint8)
comp = np.array([0, 4, 3, 2, 1, 3, 3], np.uint8)
L = 2 * L0 + 1
codes = np.zeros((nseq, L), np.uint8)
codes[:, :L0] = fwd
codes[:, L0 + 1:] = comp[fwd][:, ::-1]
# the structural N: patch list with the reference's 11-mer hashes (independent draws per digit)
kmer_cols = []
d = np.where(codes == 0, 0, codes.astype(np.int64) - 1)
ppos = (np.arange(nseq, dtype=np.uint64)[:, None] * np.uint64(L) + (L0 + np.arange(11))[None, :].astype(np.uint64)).ravel()
pk = np.zeros((nseq, 11), np.uint64)
for a, i in enumerate(range(L0, L0 + 11)):
    for t in range(11):
        dig = rng.integers(0, A, size=nseq) if i - t == L0 else d[:, i - t]
        pk[:, a] += (dig * (A ** t)).astype(np.uint64)
offsets = np.arange(nseq + 1, dtype=np.uint64) * np.uint64(L)
ss = capi.SeqSet(codes.ravel(), offsets, A, ppos, pk.ravel())
vbg = hostmodel.background_from_counts(ss.count_kmers(Kbg), A, Kbg, hostmodel.default_bg_alpha(Kbg))
for W in (12, 9, 13):
    alpha = hostmodel.default_motif_alpha(K, W)
    sites = rng.integers(1, A + 1, size=(500, W), dtype=np.uint8)
    em = capi.EM(ss, W, K, Kbg)
    em.set_model(hostmodel.motif_from_sites(sites, A, K, alpha, vbg), vbg, alpha, 0.3)
    em.iterate(2)
    em.iterate(5)
    it, e, m, u, tot = em.loop_timing()
    print("A=6 K=" + str(K) + " W=%d  %d x %d bp: %.2f ms/iteration (E %.2f, M %.2f, update %.2f)  %.3e positions.iter/s" %
          (W, nseq, L0, tot / it, e / it, m / it, u / it, nseq * L * it / (tot * 1e-3)))
    em.close()
