"""Host-side logic of the multi-GPU path: which sequences a rank owns and what the ranks exchange.

Sequences are independent in the E-step (per-sequence normalisation, EM.cpp:149-196) and the M-step is a sum over
sequences (EM.cpp:231-243), so ranks own contiguous blocks of whole sequences, balanced by stored length. The only
per-iteration exchange is a SUM over ranks of the library's exchange buffer (bamm_em_exchange_buffer): int64 words,
  [0, W*A^(K+1))   counts n[K][y][j] in 2^-40 fixed point, laid out [j][y]
  [nbin]           log likelihood in 2^-32 fixed point
  [nbin+1]         sum of all posteriors in 2^-32 fixed point (optimize_q, EM.cpp:505-519)
Integer sums are associative: the reduced buffer — and with it the model every rank computes from it — does not depend
on the rank count or the reduction order.
"""
import numpy as np

COUNT_SCALE = float(2 ** 40)
SCALAR_SCALE = float(2 ** 32)


def shard_bounds(lengths, world_size):
    """Contiguous [start, end) sequence ranges, one per rank, whole sequences only, balanced by sum of lengths:
    rank r ends at the first sequence boundary at or after r+1 equal shares of the total (FASTA order is preserved
    inside a shard, so r indexing is unchanged)."""
    lengths = np.asarray(lengths, np.int64)
    n = len(lengths)
    if world_size <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(world_size, 1) - 1)
    csum = np.cumsum(lengths)
    total = int(csum[-1])
    bounds, start = [], 0
    for r in range(world_size):
        if r == world_size - 1:
            end = n
        else:
            target = total * (r + 1) / world_size
            end = int(np.searchsorted(csum, target, side="left")) + 1
            end = min(max(end, start), n)
        bounds.append((start, end))
        start = end
    return bounds


def exchange_words(A, K, W):
    return W * A ** (K + 1) + 2


def pack_exchange(counts_jy_fx, llh, rsum):
    """counts_jy_fx: int64 [W*Yn] fixed-point counts in [j][y] order; returns the int64 exchange buffer."""
    buf = np.empty(len(counts_jy_fx) + 2, np.int64)
    buf[:-2] = counts_jy_fx
    buf[-2] = np.int64(np.rint(float(llh) * SCALAR_SCALE))
    buf[-1] = np.int64(np.rint(float(rsum) * SCALAR_SCALE))
    return buf


def unpack_exchange(buf, A, K, W):
    """Returns (n[K] as float32 in the reference's [y][j] order, llh, sum of posteriors)."""
    Yn = A ** (K + 1)
    buf = np.asarray(buf, np.int64)
    nK = (buf[:W * Yn].astype(np.float64) / COUNT_SCALE).astype(np.float32).reshape(W, Yn).T.copy()
    return nK, np.float32(buf[W * Yn] / SCALAR_SCALE), np.float32(buf[W * Yn + 1] / SCALAR_SCALE)


def allreduce_exchange(tensor, group=None):
    """In-place SUM of the exchange buffer over all ranks (torch.distributed: NCCL on the GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def optimize_q(nseq_global, rsum):
    """reference: EM::optimize_q, src/refinement/EM.cpp:515 (formula kept as is), with the GLOBAL sequence count"""
    f = np.float32
    return (f(nseq_global) - f(rsum) + f(1.0)) / (f(nseq_global) + f(2.0))


# ---- negative-set sampling in shards (bamm_seqset_sample_negatives_shard) ------------------------------------------------
def negative_draw_offset(lengths, bounds, rank, fold):
    """First rand() draw of a rank's shard of the negative set: every sampled base consumes one draw
    (SeqGenerator.cpp:285-348) and every template of stored length L yields `fold` records of length L, so the shard of
    rank r starts after fold * (stored bases of all templates on earlier ranks)."""
    lengths = np.asarray(lengths, np.int64)
    return int(fold) * int(lengths[:bounds[rank][0]].sum())


def negative_kmer_counts(codes_kmer_order2, offsets, A):
    """Host restatement of the counters bamm_seqset_negative_kmer_counts returns for one shard (orders 0, 1, 2
    concatenated; positions j >= k only, SeqGenerator.cpp:73-84) from the order-2 k-mer indices kmer % A^3."""
    y = np.asarray(codes_kmer_order2, np.int64)
    off = np.asarray(offsets, np.int64)
    out = np.zeros(A + A * A + A ** 3, np.int64)
    local = np.arange(len(y), dtype=np.int64) - np.repeat(off[:-1], np.diff(off))
    out[:A] = np.bincount(y % A, minlength=A)
    out[A:A + A * A] = np.bincount((y % (A * A))[local >= 1], minlength=A * A)
    out[A + A * A:] = np.bincount(y[local >= 2], minlength=A ** 3)
    return out


# ---- scores of a sharded scoring call -> rank 0 --------------------------------------------------------------
def gather_scores(zoops, z, shard_sizes, group=None, device=None):
    """ZOOPS score (float32) and argmax (int64) per sequence of every rank's shard, concatenated in rank order on EVERY
    rank (all_gather of shards padded to the largest one; NCCL on the GPUs, gloo in the CPU tests). Rank 0 then runs the
    reference's host-side sort / precision-recall walk (FDR::calculatePR, src/evaluation/FDR.cpp:147-332) on the whole set.
    shard_sizes: number of sequences each rank scored (known to every rank from shard_bounds)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    assert len(shard_sizes) == world
    mine = int(shard_sizes[dist.get_rank(group)])
    assert len(zoops) == mine and len(z) == mine
    cap = max(int(max(shard_sizes)), 1)
    # one 8-byte word per sequence: score bits in the low half, argmax in the high half (window starts are < 2^31)
    zi = np.asarray(z, np.int64)
    assert mine == 0 or (zi.min() >= 0 and zi.max() < (1 << 31))
    word = np.zeros(cap, np.int64)
    word[:mine] = np.ascontiguousarray(zoops, np.float32).view(np.uint32).astype(np.int64) | (zi << 32)
    t = torch.from_numpy(word)
    if device is not None:
        t = t.to(device)
    allw = torch.empty(world * cap, dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(allw, t, group=group)
    allw = allw.cpu().numpy().reshape(world, cap)
    parts = [allw[r, :int(shard_sizes[r])] for r in range(world)]
    words = np.concatenate(parts) if parts else np.zeros(0, np.int64)
    return (words & 0xffffffff).astype(np.uint32).view(np.float32), (words >> 32).astype(np.int64)
