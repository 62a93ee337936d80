#!/usr/bin/env python
"""One process, N devices (bamm_set_device_group): EM iterations on ONE c3-shaped set split over the devices — the strong-scaling
figure of the in-process route (the torchrun route is bench.py --gpus N). Usage: group_bench.py [nseq] [steps]"""
import hashlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bammmotif2_b200 import capi, synth, hostmodel

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
wl = synth.WORKLOADS["c3"]
capi.load()
ndev = capi.device_count()
fwd, sites, _ = synth.planted_sequences(1234, nseq, wl["L0"], wl["W"], motif_seed=1234)
codes = synth.stored_both_strands(fwd)
ppos, pkmer = synth.middle_n_patches(codes, 1234)
L = codes.shape[1]
offsets = np.arange(nseq + 1, dtype=np.uint64) * np.uint64(L)
capi._check(capi.load().bamm_set_device(0))
ss = capi.SeqSet(codes.ravel(), offsets, 4, ppos, pkmer)
vbg = hostmodel.background_from_counts(ss.count_kmers(wl["K_bg"]), 4, wl["K_bg"], hostmodel.default_bg_alpha(wl["K_bg"]))
alpha = hostmodel.default_motif_alpha(wl["K"], wl["W"])
v0 = hostmodel.motif_from_sites(sites, 4, wl["K"], alpha, vbg)
out = []
n = 1
while n <= ndev:
    capi.set_device_group(list(range(n)))
    t0 = time.perf_counter()
    em = capi.EM(ss, wl["W"], wl["K"], wl["K_bg"])
    t_create = time.perf_counter() - t0
    em.set_model(v0, vbg, alpha, 0.3)
    em.iterate(3)
    t0 = time.perf_counter()
    em.iterate(steps)                       # returns after every device has finished (scalars read back)
    dt = time.perf_counter() - t0
    it, e_ms, m_ms, u_ms, tot = em.loop_timing()
    row = dict(devices=n, ms_per_step_wall=dt / steps * 1e3, ms_per_step_device0=tot / max(it, 1), em_create_s=t_create,
               bp_iter_per_s=nseq * wl["L0"] * steps / dt, model_sha1=hashlib.sha1(em.model().tobytes()).hexdigest())
    out.append(row)
    print(json.dumps(row), flush=True)
    em.close()
    n *= 2
capi.set_device_group([])
base = out[0]["ms_per_step_wall"]
print(json.dumps(dict(summary="strong scaling, one process", nseq=nseq, efficiency={r["devices"]: base / (r["devices"] * r["ms_per_step_wall"]) for r in out},
                      models_identical=len({r["model_sha1"] for r in out}) == 1)))
