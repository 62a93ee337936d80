"""2-GPU probe: where does the per-iteration time go in the sharded loop? (torchrun --nproc-per-node 2 tools/mgpu_probe.py)"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bammmotif2_b200 import capi, synth, sharding
import importlib
bench = importlib.import_module("bench")
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); capi.load(); capi._check(capi.load().bamm_set_device(lr))
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", lr))
wl = dict(synth.WORKLOADS["c3"]); nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
data = bench.make_data(wl, nseq, 1234 + 1000 * rank, motif_seed=1234)
ss = capi.SeqSet(data["codes"].reshape(-1), data["offsets"], 4, data["ppos"], data["pkmer"])
v0, vbg, alpha = bench.initial_model(capi, ss, wl, data["sites"])
em = capi.EM(ss, wl["W"], wl["K"], wl["K_bg"]); em.set_model(v0, vbg, alpha, 0.3)
stream = torch.cuda.ExternalStream(em.stream(), device=torch.device("cuda", lr))
words = em.exchange_buffer()[1]
xt = torch.zeros(words, dtype=torch.int64, device="cuda"); em.set_exchange_buffer(xt.data_ptr(), words); em.set_global_nseq(nseq * world)
def loop(n, mode):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for _ in range(n):
            if mode != "ar_only":
                em.estep_local(); em.mstep_local()
            if mode != "no_ar":
                sharding.allreduce_exchange(xt)
            if mode != "ar_only":
                em.finish_iteration(sync=False)
    t1 = time.perf_counter()
    stream.synchronize(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3
import subprocess, threading
def clocks(tag):
    out = subprocess.run(["nvidia-smi", "--query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown", "--format=csv,noheader"], stdout=subprocess.PIPE, text=True).stdout
    if rank == 0: print("clocks", tag, out.replace("\n", " | "), flush=True)
class Sampler(threading.Thread):
    def __init__(self): super().__init__(daemon=True); self.stop = False; self.rows = []
    def run(self):
        while not self.stop:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"], stdout=subprocess.PIPE, text=True).stdout
            self.rows.append(out.strip().replace("\n", " ; "))
# peer-exchange object
em_nccl = em
em_peer = capi.EM(ss, wl["W"], wl["K"], wl["K_bg"]); em_peer.set_model(v0, vbg, alpha, 0.3); em_peer.set_global_nseq(nseq * world)
mine = torch.frombuffer(bytearray(em_peer.peer_alloc(rank, world)), dtype=torch.uint8).cuda()
allh = torch.empty(world * 64, dtype=torch.uint8, device="cuda"); dist.all_gather_into_tensor(allh, mine)
em_peer.peer_attach(allh.cpu().numpy().tobytes())
stream_peer = torch.cuda.ExternalStream(em_peer.stream(), device=torch.device("cuda", lr))
def loop_peer(n):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        em_peer.estep_local(); em_peer.mstep_local(); em_peer.finish_iteration(sync=False)
    stream_peer.synchronize(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
loop_peer(3)
smp = Sampler(); 
if rank == 0: smp.start()
tp = loop_peer(60)
smp.stop = True
e_ms, m_ms = em_peer.timing()
print("rank %d peer     total %.3f ms/iter   last E %.3f  last M+exchange+update %.3f" % (rank, tp, e_ms, m_ms), flush=True)
if rank == 0: print("clock samples during peer loop:", smp.rows[:12], flush=True)
smp2 = Sampler()
if rank == 0: smp2.start()
loop(3, "no_ar"); enq, tot = loop(60, "no_ar")
smp2.stop = True
if rank == 0: print("no_ar 60 iters total %.3f; clock samples:" % tot, smp2.rows[:12], flush=True)
for mode in ("full", "no_ar", "ar_only", "full"):
    loop(3, mode)
    enq, tot = loop(20, mode)
    e_ms, m_ms = em.timing() if mode != "ar_only" else (0, 0)
    print("rank %d %-8s enqueue %.3f ms/iter   total %.3f ms/iter   last E %.3f  last M+exchange+update %.3f" % (rank, mode, enq, tot, e_ms, m_ms), flush=True)
dist.destroy_process_group()
