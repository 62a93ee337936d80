"""Timing probe of the row f-4 entry points at the c3 shape (300k x 500 bp, W=20, K=4): bamm_em_mask and
bamm_seqset_sample_pwm_sites. Wall clock around the C-ABI calls (synchronous)."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bammmotif2_b200 import capi, synth, hostmodel
import bench
wl = dict(synth.WORKLOADS["c3"], name="c3")
nseq = int(os.environ.get("NSEQ", 300000))
data = bench.make_data(wl, nseq, 1234)
ss = capi.SeqSet(data["codes"].reshape(-1), data["offsets"], 4, data["ppos"], data["pkmer"])
v0, vbg, alpha = bench.initial_model(capi, ss, wl, data["sites"], None)
em = capi.EM(ss, wl["W"], wl["K"], wl["K_bg"])
em.set_model(v0, vbg, alpha, 0.3)
t0 = time.perf_counter(); res = em.mask(f=0.05); t1 = time.perf_counter()
print("bamm_em_mask: %d sequences x %d positions, f=0.05: %d windows kept, %d iterations, %.3f s (%.1f ms per iteration incl. phase 1 + sort)" %
      (nseq, data["L"], res["nkept"], res["iterations"], t1 - t0, 1e3 * (t1 - t0) / res["iterations"]))
em.set_model(v0, vbg, alpha, 0.3)
t0 = time.perf_counter(); r2 = em.optimize(); t1 = time.perf_counter()
print("bamm_em_optimize on the same data: %d iterations, %.3f s" % (r2["iterations"], t1 - t0))
W, K = wl["W"], wl["K"]
score = (v0[:4 * W].reshape(4, W) / vbg[:4, None]).astype(np.float32)
u = np.random.default_rng(1).random(nseq)
n_all = np.zeros(capi.model_size(4, K, W), np.int32)
lib = capi.load()
t0 = time.perf_counter()
capi._check(lib.bamm_seqset_sample_pwm_sites(ss.h, None, 0, W, K, 4, score.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 0.3,
                                             u.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n_all.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), None))
t1 = time.perf_counter()
print("bamm_seqset_sample_pwm_sites: %d sequences, %d sites counted, %.3f s" % (nseq, n_all[:4 * W].sum() // W, t1 - t0))
