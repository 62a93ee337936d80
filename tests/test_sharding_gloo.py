"""Multi-rank host logic on CPU (gloo, world_size 2): sequence sharding + the int64 exchange of the M-step.

Each rank takes its shard of a golden case, computes its posteriors with the CPU oracle, turns them into the
library's fixed-point partial counts, all-reduces the exchange buffer, and every rank must end up with the SAME bits
as a single process — and with the reference's first-iteration model within the 1e-5 tolerance.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from bammmotif2_b200 import sharding  # noqa: E402


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(0)
    L = rng.integers(50, 2000, size=1000)
    for world in (1, 2, 3, 4, 8):
        b = sharding.shard_bounds(L, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == len(L)
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))            # contiguous, whole sequences
        loads = np.array([L[s:e].sum() for s, e in b])
        assert loads.max() - loads.min() <= 2 * L.max()                           # balanced to within a sequence or two
    assert sharding.shard_bounds([], 2) == [(0, 0), (0, 0)]
    assert sharding.shard_bounds([10], 4)[0] == (0, 1)                            # fewer sequences than ranks: empty shards


def partial_exchange(orc, g, lo, hi):
    """Exchange buffer of the sequences [lo, hi) of golden case g for its initial model (E-step on the oracle)."""
    A, K, W = g.A, g.K, g.W
    Yn = A ** (K + 1)
    off = g["pos_offsets"].astype(np.int64)
    kmer = np.ascontiguousarray(g["pos_kmer"][off[lo]:off[hi]])
    soff = (off[lo:hi + 1] - off[lo]).astype(np.uint64)
    counts = np.zeros(W * Yn, np.int64)
    if hi == lo:
        return sharding.pack_exchange(counts, 0.0, 0.0), 0
    s = orc.linear_s(g["m1_v_init"], g["bg_v"], A, K, g.K_bg, W)
    r, llh = orc.estep(kmer, soff, A, K, W, s, g.q)
    y = (kmer % np.uint64(Yn)).astype(np.int64)
    for n in range(hi - lo):                                   # gather form of EM::MStep
        b, L = int(soff[n]), int(soff[n + 1] - soff[n])
        rn = r[b:b + L]
        for p in range(L - W + 1):
            X = np.int64(np.rint(np.float32(rn[L - W - p]) * np.float32(sharding.COUNT_SCALE)))
            if X == 0:
                continue
            jmax = min(W - 1, L - W - p)
            j = np.arange(jmax + 1)
            np.add.at(counts, j * Yn + y[b + p + j], X)
    return sharding.pack_exchange(counts, llh, float(r.sum(dtype=np.float64))), hi - lo


def _worker(rank, world, port, case, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from util import Golden
    from oracle import oracle as orc
    orc.build()
    g = Golden(case)
    L = np.diff(g["pos_offsets"].astype(np.int64))
    lo, hi = sharding.shard_bounds(L, world)[rank]
    buf, _ = partial_exchange(orc, g, lo, hi)
    t = torch.from_numpy(buf.copy())
    sharding.allreduce_exchange(t)
    np.save(os.path.join(out, "rank%d.npy" % rank), t.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["syn_k2_N", "syn_k3_fdr"])
def test_two_rank_exchange_equals_single_process(case, tmp_path, oracle):
    from util import Golden
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, r1)                                                   # every rank holds the same bits
    g = Golden(case)
    nseq = len(g["pos_offsets"]) - 1
    single, _ = partial_exchange(oracle, g, 0, nseq)
    assert np.array_equal(r0[:-2], single[:-2])                                     # integer sums: rank count does not matter
    # scalars: the oracle sums each shard's log likelihood sequentially in fp32 before the fixed-point conversion
    assert abs(int(r0[-2]) - int(single[-2])) <= 1e-6 * abs(int(single[-2])) and abs(int(r0[-1]) - int(single[-1])) <= 1e-6 * abs(int(single[-1])) + 64
    # and the model every rank derives from the reduced buffer is the reference's first-iteration model
    nK, llh, rsum = sharding.unpack_exchange(r0, g.A, g.K, g.W)
    assert abs(llh - g["m1_llh"][0]) <= 1e-5 * abs(g["m1_llh"][0])
    n_all = np.zeros(len(g["m1_v_init"]), np.float32)
    off = [0]
    for k in range(g.K + 1):
        off.append(off[-1] + g.A ** (k + 1) * g.W)
    n_all[off[g.K]:] = nK.ravel()
    for k in range(g.K, 0, -1):                                                     # fold to lower orders (EM.cpp:247-254)
        cur = n_all[off[k]:off[k + 1]].reshape(g.A ** (k + 1), g.W)
        low = np.zeros((g.A ** k, g.W), np.float32)
        for a in range(g.A):
            low += cur[a * g.A ** k:(a + 1) * g.A ** k]
        n_all[off[k - 1]:off[k]] = low.ravel()
    v = oracle.update_v(n_all, g["m1_alpha"], g["bg_v"], g.A, g.K, g.W, g["m1_v_init"].copy())
    ref = g["m1_v_it1"]
    assert np.all(np.abs(v - ref) <= 1e-5 * np.abs(ref))


def _neg_worker(rank, world, port, case, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from util import Golden
    g = Golden(case)
    off = g["pos_offsets"].astype(np.int64)
    L = np.diff(off)
    bounds = sharding.shard_bounds(L, world)
    lo, hi = bounds[rank]
    y2 = (g["pos_kmer"][off[lo]:off[hi]] % np.uint64(g.A ** 3)).astype(np.int64)
    cnt = torch.from_numpy(sharding.negative_kmer_counts(y2, off[lo:hi + 1] - off[lo], g.A))
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    out[rank] = (cnt.numpy().copy(), sharding.negative_draw_offset(L, bounds, rank, g.meta["mFold"]), hi - lo)
    dist.destroy_process_group()


def test_negative_sampling_shards_gloo():
    """Host logic of the sharded negative sampler on two gloo ranks: the all-reduced k-mer counters equal the whole set's,
    and the draw offsets tile the reference's rand() stream without gap or overlap."""
    from util import Golden
    case = "neg_ragged_N"
    g = Golden(case)
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_neg_worker, args=(world, 29533, case, out), nprocs=world, join=True)
    off = g["pos_offsets"].astype(np.int64)
    whole = sharding.negative_kmer_counts((g["pos_kmer"] % np.uint64(g.A ** 3)).astype(np.int64), off, g.A)
    assert np.array_equal(out[0][0], whole) and np.array_equal(out[1][0], whole)
    L = np.diff(off)
    fold = g.meta["mFold"]
    assert out[0][1] == 0
    assert out[1][1] == fold * int(L[:out[0][2]].sum())
    assert out[1][1] + fold * int(L[out[0][2]:].sum()) == len(g["neg_codes"])       # one draw per sampled base


def _score_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    L = rng.integers(30, 400, size=101)
    zo_all = rng.standard_normal(101).astype(np.float32)
    zo_all[5] = -np.inf                                         # a window score can be -inf (log of a zero probability)
    z_all = rng.integers(0, L - 20)
    bounds = sharding.shard_bounds(L, world)
    lo, hi = bounds[rank]
    zo, z = sharding.gather_scores(zo_all[lo:hi], z_all[lo:hi], [e - s for s, e in bounds])
    np.save(os.path.join(out, "zo%d.npy" % rank), zo)
    np.save(os.path.join(out, "z%d.npy" % rank), z)
    if rank == 0:
        np.save(os.path.join(out, "zo_all.npy"), zo_all)
        np.save(os.path.join(out, "z_all.npy"), z_all)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gathered_scores_equal_the_unsharded_arrays(world, tmp_path):
    """Row e, scoring: every rank's ZOOPS scores + argmax arrive in sequence order, bit for bit (ragged shards)."""
    port = 29620 + world + (os.getpid() % 200)
    mp.spawn(_score_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    zo_all, z_all = np.load(tmp_path / "zo_all.npy"), np.load(tmp_path / "z_all.npy")
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("zo%d.npy" % r)).view(np.uint32), zo_all.view(np.uint32))
        assert np.array_equal(np.load(tmp_path / ("z%d.npy" % r)), z_all)
