"""GPU diagnostic: fraction of windows whose posterior survives the M-step's 2^-41 fixed-point threshold, per iteration."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bammmotif2_b200 import capi, synth, hostmodel
import importlib
bench = importlib.import_module("bench")
wl = dict(synth.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"])
nseq = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
data = bench.make_data(wl, nseq, 1234)
ss = capi.SeqSet(data["codes"].reshape(-1), data["offsets"], 4, data["ppos"], data["pkmer"])
v0, vbg, alpha = bench.initial_model(capi, ss, wl, data["sites"])
em = capi.EM(ss, wl["W"], wl["K"], wl["K_bg"])
em.set_model(v0, vbg, alpha, 0.3)
for it in range(1, 13):
    llh = em.estep()
    r = em.r()
    print("iter %2d llh %.1f active %.4f  (r>=1e-6: %.4f)" % (it, llh, float((r >= 2.0 ** -41).mean()), float((r >= 1e-6).mean())), flush=True)
    em.mstep()
