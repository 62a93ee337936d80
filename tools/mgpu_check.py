"""N-GPU parity check (torchrun --nproc-per-node N tools/mgpu_check.py): every rank runs EM on its shard of a golden case
with the NVLink peer exchange (and again with the NCCL int64 all-reduce); all ranks must hold the SAME model bits as a
single-GPU run over the whole set, and the reference's model within 1e-4. Prints one line per mode; exit code 1 on mismatch."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bammmotif2_b200 import capi, sharding
from util import Golden
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); capi.load(); capi._check(capi.load().bamm_set_device(lr))
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", lr))
bad = 0
for case in ("jund_k2", "syn_k4", "syn_k3_fdr"):
    g = Golden(case)
    iters = min(g.iterations, 12)
    pp, pk = capi.kmer_patches(g["pos_codes"], g["pos_kmer"])
    ss = capi.SeqSet(g["pos_codes"], g["pos_offsets"], g.A, pp, pk)
    nseq = ss.nseq
    single = capi.EM(ss, g.W, g.K, g.K_bg_model)
    single.set_model(g["m1_v_init"], g["bg_v"], g["m1_alpha"], g.q)
    llh1, _ = single.iterate(iters)
    v1 = single.model()
    L = np.diff(g["pos_offsets"].astype(np.int64))
    lo, hi = sharding.shard_bounds(L, world)[rank]
    for mode in ("peer", "nccl"):
        em = capi.EM(ss, g.W, g.K, g.K_bg_model, subset=np.arange(lo, hi, dtype=np.uint64))
        em.set_model(g["m1_v_init"], g["bg_v"], g["m1_alpha"], g.q)
        em.set_global_nseq(nseq)
        stream = torch.cuda.ExternalStream(em.stream(), device=torch.device("cuda", lr))
        xt = None
        if mode == "peer":
            mine = torch.frombuffer(bytearray(em.peer_alloc(rank, world)), dtype=torch.uint8).cuda()
            allh = torch.empty(world * 64, dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(allh, mine)
            em.peer_attach(allh.cpu().numpy().tobytes())
        else:
            words = em.exchange_buffer()[1]
            xt = torch.zeros(words, dtype=torch.int64, device="cuda")
            em.set_exchange_buffer(xt.data_ptr(), words)
        for it in range(iters):
            with torch.cuda.stream(stream):
                em.estep_local(); em.mstep_local()
                if xt is not None:
                    sharding.allreduce_exchange(xt)
            llh, vd = em.finish_iteration(sync=True)
        v = em.model()
        same = np.array_equal(v, v1)
        ref = g["m1_v_it%d" % iters] if ("m1_v_it%d" % iters) in g else None
        close = True if ref is None else bool(np.all(np.abs(v - ref) <= 1e-4 * np.abs(ref)))
        t = torch.tensor([int(same and close)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("%-10s %-4s world %d iters %d: model bits equal to 1-GPU: %s, llh %.6f vs %.6f, within 1e-4 of the reference: %s" % (
                case, mode, world, iters, bool(t.item()), llh, llh1, close if ref is not None else "n/a"), flush=True)
        bad += 0 if t.item() else 1
        em.close()
        dist.barrier()
dist.destroy_process_group()
sys.exit(1 if bad else 0)
