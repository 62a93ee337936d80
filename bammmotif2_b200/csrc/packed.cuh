// packed.cuh — the fast path for the 4-letter alphabet: sequences as 2-bit packed base streams.
//
// Layout. A "regular" sequence (codes 1..4 everywhere, except optionally the structural N in the
// middle of a both-strand sequence, Sequence.cpp:10-14) is stored as 32-bit words of 16 bases, FIRST base in the
// most significant bits, preceded by two zero pad words (32 bases) and followed by six. The order-K k-mer index of the
// reference, y(i) = sum_t c(i-t) 4^t (Sequence.cpp:35-41), is then literally a bit field of the stream: for a
// window start p the 64-bit word  w = bases p-K .. p-K+31  gives  y(p+j) = (w >> (62-2K-2j)) & (4^(K+1)-1), and the
// zero pad supplies the implicit leading 'A's of k-mers that start before the sequence does. No index array is read.
//
// E-step: tuple tables. T consecutive motif columns are folded into one lookup over the (K+T)-mer z that ends
// at the tuple's last base:  tab[c][z] = prod_{t<T} s[cT+t][ (z >> 2(T-1-t)) & maskK ],  so a window costs
// ceil(W/T) shared-memory lookups instead of W. Windows that cannot use tuples (the last W-1 truncated windows,
// EM.cpp:167, and windows over the N whose k-mers hold rand() draws, Sequence.cpp:38) take an exact slow path
// over the plain table in global memory.
//
// M-step: sparse. r is converted to 2^40 fixed point; every window whose r rounds to 0 contributes exactly
// nothing, and in practice that is 60-95 % of all windows. Surviving (window, value) pairs are compacted into a
// per-warp queue so that the scatter loop runs with full lanes.
#pragma once
#include "kernels.cuh"

namespace bamm {

struct PackedSeq {           // per regular sequence
    uint64_t word_off;       // index of the sequence's FIRST DATA word (two zero pad words sit before it)
    uint32_t L;              // stored length
    uint32_t mid;            // position of the structural N, or 0xffffffff
};

struct PackedView {
    const uint32_t* words;             // packed stream, 16 bases per 32-bit word
    const PackedSeq* seqs;             // [nseq] (entries of irregular sequences are unused)
    const uint16_t* ypatch;            // [nseq][K+1] order-K k-mer index at positions mid..mid+K (rand() draws inside)
    const uint32_t* seq_ids;           // list -> seqset index
    const uint64_t* r_off;             // list -> offset of the sequence's r
    uint32_t nlist;
};

// ---- classification + packing (device side of bamm_seqset_create) --------------------------------------------------
// kind: 0 irregular, 1 regular without N, 2 regular with exactly one 0 code at (L-1)/2, L odd
__global__ void k_classify(const uint8_t* __restrict__ codes, const uint64_t* __restrict__ off, uint64_t nseq,
                           uint8_t* __restrict__ kind) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t n = warp; n < nseq; n += nwarps) {
        const uint64_t base = off[n], L = off[n + 1] - base;
        const uint64_t mid = (L & 1) ? (L - 1) / 2 : ~0ull;
        int bad = 0, zeros_at_mid = 0;
        for (uint64_t i = lane; i < L; i += 32) {
            const uint32_t c = codes[base + i];
            if (c == 0) { if (i == mid) zeros_at_mid = 1; else bad = 1; }
            else if (c > 4) bad = 1;
        }
        bad = __any_sync(FULL, bad);
        zeros_at_mid = __any_sync(FULL, zeros_at_mid);
        if (lane == 0) kind[n] = (bad || L == 0 || L >= 0xffffff00ull) ? 0 : (zeros_at_mid ? 2 : 1);
    }
}

// every patch must sit at mid..mid+10 of a kind-2 sequence; anything else makes its sequence irregular.
// cover[n] counts the patches that landed in the structural region.
__global__ void k_check_patches(const uint64_t* __restrict__ ppos, uint64_t np, const uint64_t* __restrict__ off,
                                uint64_t nseq, uint8_t* __restrict__ kind, uint32_t* __restrict__ cover) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const uint64_t pos = ppos[i];
    uint64_t lo = 0, hi = nseq;                      // last n with off[n] <= pos
    while (hi - lo > 1) { const uint64_t m = (lo + hi) >> 1; if (off[m] <= pos) lo = m; else hi = m; }
    const uint64_t n = lo, L = off[n + 1] - off[n], local = pos - off[n];
    const uint64_t mid = (L - 1) / 2;
    if (kind[n] == 2 && local >= mid && local <= mid + 10) atomicAdd(&cover[n], 1u);
    else kind[n] = 0;                                // benign race: every writer stores 0
}

__global__ void k_finish_kinds(const uint64_t* __restrict__ off, uint64_t nseq, const uint32_t* __restrict__ cover,
                               uint8_t* __restrict__ kind) {
    const uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nseq) return;
    if (kind[n] == 2) {
        const uint64_t L = off[n + 1] - off[n], mid = (L - 1) / 2;
        const uint64_t need = (L - mid < 11) ? L - mid : 11;      // positions mid..min(mid+10, L-1)
        if (cover[n] != need) kind[n] = 0;                         // the caller did not patch the N: generic path
    }
}

// one lane per output word (16 bases)
__global__ void k_pack(const uint8_t* __restrict__ codes, const uint64_t* __restrict__ off, uint64_t nseq,
                       const uint8_t* __restrict__ kind, const PackedSeq* __restrict__ seqs,
                       uint32_t* __restrict__ words) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t n = warp; n < nseq; n += nwarps) {
        if (kind[n] == 0) continue;
        const uint64_t base = off[n], L = off[n + 1] - base;
        const uint64_t nw = (L + 15) / 16;
        uint32_t* dst = words + seqs[n].word_off;
        if (lane < 2) dst[-1 - lane] = 0u;
        if (lane < 6) dst[nw + lane] = 0u;
        for (uint64_t wi = lane; wi < nw; wi += 32) {
            uint32_t w = 0;
            for (int b = 0; b < 16; b++) {
                const uint64_t i = wi * 16 + b;
                const uint32_t c = (i < L) ? codes[base + i] : 0u;
                w = (w << 2) | (c ? c - 1 : 0u);
            }
            dst[wi] = w;
        }
    }
}

__global__ void k_make_ypatch(const uint64_t* __restrict__ ppos, const uint64_t* __restrict__ pkmer, uint64_t np,
                              const uint64_t* __restrict__ off, uint64_t nseq, const uint8_t* __restrict__ kind,
                              int K, uint64_t Yn, uint16_t* __restrict__ ypatch) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const uint64_t pos = ppos[i];
    uint64_t lo = 0, hi = nseq;
    while (hi - lo > 1) { const uint64_t m = (lo + hi) >> 1; if (off[m] <= pos) lo = m; else hi = m; }
    const uint64_t n = lo;
    if (kind[n] != 2) return;
    const uint64_t L = off[n + 1] - off[n], d = pos - off[n] - (L - 1) / 2;
    if (d <= (uint64_t)K) ypatch[n * (uint64_t)(K + 1) + d] = (uint16_t)(pkmer[i] % Yn);
}

// ---- tables ------------------------------------------------------------------------------------------------------
// tab[c][z] for c < C = ceil(W/T), z < 4^(K+T); s is the plain table [j][y]. LOG != 0: sums instead of products.
__global__ void k_make_tuple_table(const float* __restrict__ s, int W, int K, int T, int C, uint32_t Yn,
                                   float* __restrict__ tab) {
    const uint32_t Zn = Yn << (2 * (T - 1));
    const uint32_t maskK = Yn - 1;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)C * Zn; i += gridDim.x * blockDim.x) {
        const uint32_t c = i / Zn, z = i % Zn;
        float p = 1.0f;
        for (int t = 0; t < T; t++) {
            const int j = (int)c * T + t;
            if (j < W) p *= s[(uint32_t)j * Yn + ((z >> (2 * (T - 1 - t))) & maskK)];
        }
        tab[i] = p;
    }
}

// ---- window extraction ---------------------------------------------------------------------------------------------
// 64 bits holding bases b0 .. b0+31 of a sequence (b0 >= -32: the pad words supply leading zeros), as (hi, lo).
__device__ __forceinline__ void window_bits(const uint32_t* __restrict__ wd, int b0, uint32_t& whi, uint32_t& wlo) {
    const int wi = b0 >> 4;                      // floor
    const int s = 2 * (b0 & 15);
    const uint32_t t0 = wd[wi], t1 = wd[wi + 1], t2 = wd[wi + 2];
    whi = __funnelshift_l(t1, t0, s);
    wlo = __funnelshift_l(t2, t1, s);
}
__device__ __forceinline__ unsigned long long window_word(const uint32_t* __restrict__ wd, int b0) {
    uint32_t hi, lo; window_bits(wd, b0, hi, lo);
    return ((unsigned long long)hi << 32) | lo;
}

__device__ __forceinline__ uint32_t field(unsigned long long w, int shift, uint32_t mask) {
    return (uint32_t)(w >> shift) & mask;
}

struct Plan {                 // launch-invariant parameters of the packed kernels
    int W, K, T, C;           // tuple size T, C = ceil(W/T) lookups per window
    uint32_t Yn, Zn;          // 4^(K+1), 4^(K+T)
    float q;
};

// ---- E-step --------------------------------------------------------------------------------------------------------
// reference: EM::EStep, src/refinement/EM.cpp:139-200 (gather form). One warp per sequence, lanes =
// window starts. The C tuple lookups of a window are fully unrolled (C is a template parameter); the 64-bit window
// word is consumed from the top, 2T bits per lookup. Unnormalised posteriors go to r, the normaliser is reduced
// over the warp, and a second sweep scales by 1/norm (the lines are still in L2).
//
// Chunks that contain windows tuples cannot fully serve — the last W-1 truncated windows (EM.cpp:167) and the
// windows over the k-mers that hold the N's rand() draws (positions mid..mid+K, Sequence.cpp:38) — run a masked
// variant: per lane a bit mask of the tuples that are whole and untouched (taken from the shared table as usual)
// and a bit mask of the single columns that must come from the plain table in global memory.
template <int C>
__global__ void __launch_bounds__(1024, 1)
k_estep_packed(PackedView pv, Plan pl, const float* __restrict__ tab_g /* [C][Zn] */, const float* __restrict__ s_g /* [W][Yn] */,
               float* __restrict__ r, unsigned long long* __restrict__ scal) {
    extern __shared__ float tab[];
    for (uint32_t i = threadIdx.x; i < (uint32_t)C * pl.Zn; i += blockDim.x) tab[i] = tab_g[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    const int W = pl.W, K = pl.K, T = pl.T;
    const int zb = 2 * (K + T);                            // bits of one tuple index
    const int roll = 2 * T;
    const uint32_t zn_bytes = pl.Zn * 4u;
    const uint32_t maskK = pl.Yn - 1;
    const char* tabc = reinterpret_cast<const char*>(tab);
    long long llh_fx = 0, rsum_fx = 0;
    const float one_minus_q = 1.0f - pl.q;
    for (uint32_t li = warp; li < pv.nlist; li += nwarps) {
        const uint32_t n = pv.seq_ids[li];
        const PackedSeq sq = pv.seqs[n];
        const int L = (int)sq.L, LW1 = L - W + 1;
        const int mid = (int)sq.mid;                       // -1 when there is no N
        // per-lane view of the stream: this lane's windows start at bases lane-K + 32*ch, i.e. always at the same
        // bit offset sft inside a 32-bit word; three words are kept and two new ones are fetched per chunk
        const uint32_t* __restrict__ wl = pv.words + sq.word_off + ((lane - K) >> 4);
        const int sft = 2 * ((lane - K) & 15);
        uint32_t t0 = wl[0], t1 = wl[1], t2 = wl[2];
        float* __restrict__ rn = r + pv.r_off[li];
        const float pos = pl.q / (float)LW1;
        const int tail0 = L - 2 * W + 2;                   // first truncated window (p > L-2W+1)
        float sum = 0.0f;
        for (int p0 = 0; p0 < LW1; p0 += 32) {
            const int p = p0 + lane;
            uint32_t whi = __funnelshift_l(t1, t0, sft), wlo = __funnelshift_l(t2, t1, sft);
            const unsigned long long w = ((unsigned long long)whi << 32) | wlo;
            wl += 2;
            t0 = t2; t1 = wl[1]; t2 = wl[2];
            float prod = 1.0f;
            const bool chunk_slow = (p0 + 31 >= tail0) || (mid >= 0 && p0 <= mid + K && p0 + 31 + W - 1 >= mid);   // warp-uniform
            if (!chunk_slow) {
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const uint32_t z4 = (whi >> (32 - zb)) << 2;
                    prod *= *reinterpret_cast<const float*>(tabc + (uint32_t)c * zn_bytes + z4);
                    whi = __funnelshift_l(wlo, whi, roll);
                    wlo <<= roll;
                }
            } else {
                const int jmax = (p < LW1) ? min(W - 1, L - W - p) : -1;
                const int cfull = (jmax + 1) / T;                               // whole tuples inside the truncation
                uint32_t good = (1u << cfull) - 1u;
                uint32_t cols = 0;
                if (cfull * T <= jmax) cols = ((2u << jmax) - 1u) & ~((1u << (cfull * T)) - 1u);   // truncated remainder
                const bool over_n = mid >= 0 && p <= mid + K && p + W - 1 >= mid;
                if (over_n) {
                    // tuples c with p+cT+T-1 >= mid and p+cT <= mid+K
                    const int a = mid - p - T + 1;
                    const int c_lo = a <= 0 ? 0 : (a + T - 1) / T;
                    const int c_hi = min(C - 1, (mid + K - p) / T);
                    if (c_hi >= c_lo) {
                        const uint32_t cm = ((2u << c_hi) - 1u) & ~((1u << c_lo) - 1u);
                        const int jl = c_lo * T, jh = min(jmax, c_hi * T + T - 1);
                        if (jh >= jl) cols |= ((2u << jh) - 1u) & ~((1u << jl) - 1u);
                        good &= ~cm;
                    }
                }
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const uint32_t z4 = (whi >> (32 - zb)) << 2;
                    const float v = *reinterpret_cast<const float*>(tabc + (uint32_t)c * zn_bytes + z4);
                    prod *= ((good >> c) & 1u) ? v : 1.0f;
                    whi = __funnelshift_l(wlo, whi, roll);
                    wlo <<= roll;
                }
                while (cols) {
                    const int j = __ffs(cols) - 1;
                    cols &= cols - 1u;
                    uint32_t y = field(w, 62 - 2 * K - 2 * j, maskK);
                    const int d = p + j - mid;
                    if (over_n && d >= 0 && d <= K) y = pv.ypatch[(uint64_t)n * (K + 1) + d];
                    prod *= __ldg(&s_g[(uint32_t)j * pl.Yn + y]);
                }
            }
            if (p < LW1) {
                const float val = prod * pos;
                rn[L - W - p] = val;
                sum += val;
            }
        }
        sum = warp_sum(sum);
        const float norm = one_minus_q + sum;
        const float rnorm = __frcp_rn(norm);
        __syncwarp();
        for (int k = lane; k < L; k += 32) rn[k] = (k < LW1) ? rn[k] * rnorm : 0.0f;
        if (lane == 0) {
            llh_fx += __double2ll_rn((double)logf(norm) * SC_SCALE_D);
            rsum_fx += __double2ll_rn((double)(sum * rnorm) * SC_SCALE_D);
        }
    }
    if (lane == 0) {
        if (llh_fx) atomicAdd(&scal[0], (unsigned long long)llh_fx);
        if (rsum_fx) atomicAdd(&scal[1], (unsigned long long)rsum_fx);
    }
}

// ---- M-step --------------------------------------------------------------------------------------------------------
// reference: EM::MStep accumulation, src/refinement/EM.cpp:230-243 (gather form).
// Pass 1 per warp and sequence: stream r in batches of 256 windows (8 per lane; the next batch's loads are issued
// before the current one is examined), keep the windows whose r is at least half a fixed-point unit (everything
// below rounds to exactly 0 and contributes nothing) and compact them with a warp scan into the warp's ring queue.
// Pass 2 whenever 32 entries are queued (and once more at the end of the sequence): each lane scatters one
// window's value into the W bins it touches with native 32-bit shared atomics. Low-word wrap-arounds are collected
// in a bit mask and carried into the CTA's 64-bit partial table after the loop, so the scatter loop is branch-free.
struct QEntry { uint32_t p; float rv; };
constexpr int QCAP = 256;                      // ring entries per warp (>= 31 + 128)
constexpr int M_UNROLL = 8;                    // chunks of 32 windows in flight per lane
constexpr int M_GROUP = 4;                     // chunks compacted per warp scan
constexpr float FX_HALF_UNIT = 4.547473508864641e-13f;   // 2^-41: smallest r that rounds to a non-zero count

struct SeqCtx { const uint32_t* wd; const uint16_t* yp; int L, mid; };

__device__ __forceinline__ void scatter_window(const SeqCtx& sc, const Plan& pl, uint32_t* __restrict__ lo_sh,
                                               unsigned long long* __restrict__ mypart, int p, float rv) {
    const int W = pl.W, K = pl.K;
    const unsigned long long X = __float2ull_rn(rv * FX_SCALE_F);
    if (X == 0) return;
    const int L = sc.L, mid = sc.mid;
    const unsigned long long w = window_word(sc.wd, p - K);      // bases p-K .. p-K+31
    const uint32_t maskK = pl.Yn - 1;
    const int jmax = min(W - 1, L - W - p);
    const bool over_n = mid >= 0 && p <= mid + K && p + W - 1 >= mid;
    const uint32_t xlo = (uint32_t)X, xhi = (uint32_t)(X >> 32);
    int sh = 62 - 2 * K;
    uint32_t jb = 0;
    uint32_t carry = 0;                        // bit j: the low word of bin (j, y_j) wrapped
    if (!over_n) {
        for (int j = 0; j <= jmax; j++) {
            const uint32_t old = atomicAdd(&lo_sh[jb + field(w, sh, maskK)], xlo);
            carry |= ((uint32_t)(old + xlo) < old ? 1u : 0u) << j;
            sh -= 2; jb += pl.Yn;
        }
    } else {
        for (int j = 0; j <= jmax; j++) {
            uint32_t y = field(w, sh, maskK);
            const int d = p + j - mid;
            if (d >= 0 && d <= K) y = sc.yp[d];
            const uint32_t old = atomicAdd(&lo_sh[jb + y], xlo);
            carry |= ((uint32_t)(old + xlo) < old ? 1u : 0u) << j;
            sh -= 2; jb += pl.Yn;
        }
    }
    // high words: large r (xhi != 0) touches every bin, otherwise only the bins whose low word wrapped
    uint32_t todo = xhi ? (jmax >= 31 ? 0xffffffffu : ((2u << jmax) - 1u)) : carry;
    while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1u;
        uint32_t y = field(w, 62 - 2 * K - 2 * j, maskK);
        const int d = p + j - mid;
        if (over_n && d >= 0 && d <= K) y = sc.yp[d];
        const uint32_t h = xhi + ((carry >> j) & 1u);
        atomicAdd(&mypart[(uint32_t)j * pl.Yn + y], (unsigned long long)h << 32);
    }
}

__global__ void __launch_bounds__(512)
k_mstep_packed(PackedView pv, Plan pl, const float* __restrict__ r, unsigned long long* __restrict__ part /* [gridDim.x][W*Yn] */) {
    extern __shared__ uint32_t smem_u32[];
    const uint32_t nbin = (uint32_t)pl.W * pl.Yn;
    uint32_t* lo_sh = smem_u32;
    QEntry* queues = reinterpret_cast<QEntry*>(smem_u32 + ((nbin + 3) & ~3u));
    unsigned long long* mypart = part + (uint64_t)blockIdx.x * nbin;
    for (uint32_t i = threadIdx.x; i < nbin; i += blockDim.x) lo_sh[i] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    QEntry* q = queues + wib * QCAP;
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + wib;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    const int W = pl.W, K = pl.K;
    for (uint32_t li = warp; li < pv.nlist; li += nwarps) {
        const uint32_t n = pv.seq_ids[li];
        const PackedSeq sq = pv.seqs[n];
        SeqCtx sc; sc.wd = pv.words + sq.word_off; sc.yp = pv.ypatch + (uint64_t)n * (K + 1); sc.L = (int)sq.L; sc.mid = (int)sq.mid;
        const int L = sc.L, LW1 = L - W + 1;
        const float* __restrict__ rn = r + pv.r_off[li];
        uint32_t qhead = 0, qcount = 0;        // warp-uniform
        // r index i = L-W-p runs over [0, LW1); lane takes i = i0 + u*32 + lane
        float cur[M_UNROLL], nxt[M_UNROLL];
#pragma unroll
        for (int u = 0; u < M_UNROLL; u++) { const int i = u * 32 + lane; cur[u] = (i < LW1) ? __ldcs(&rn[i]) : 0.0f; }
        for (int i0 = 0; i0 < LW1; i0 += 32 * M_UNROLL) {
#pragma unroll
            for (int u = 0; u < M_UNROLL; u++) {
                const int i = i0 + 32 * M_UNROLL + u * 32 + lane;
                nxt[u] = (i < LW1) ? __ldcs(&rn[i]) : 0.0f;
            }
#pragma unroll
            for (int g = 0; g < M_UNROLL; g += M_GROUP) {
                uint32_t act = 0;
#pragma unroll
                for (int u = 0; u < M_GROUP; u++) act |= (cur[g + u] >= FX_HALF_UNIT ? 1u : 0u) << u;
                if (__any_sync(FULL, act != 0)) {
                    const uint32_t cnt = __popc(act);
                    uint32_t incl = cnt;                                   // inclusive prefix of the per-lane counts
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
                    const uint32_t total = __shfl_sync(FULL, incl, 31);
                    uint32_t slot = qhead + qcount + incl - cnt;
#pragma unroll
                    for (int u = 0; u < M_GROUP; u++) {
                        if (act & (1u << u)) {
                            QEntry& e = q[slot & (QCAP - 1)];
                            e.p = (uint32_t)(L - W - (i0 + (g + u) * 32 + lane)); e.rv = cur[g + u];
                            slot++;
                        }
                    }
                    qcount += total;
                    __syncwarp();
                    while (qcount >= 32) {
                        const QEntry e = q[(qhead + lane) & (QCAP - 1)];
                        scatter_window(sc, pl, lo_sh, mypart, (int)e.p, e.rv);
                        qhead = (qhead + 32) & (QCAP - 1); qcount -= 32;
                    }
                    __syncwarp();
                }
            }
#pragma unroll
            for (int u = 0; u < M_UNROLL; u++) cur[u] = nxt[u];
        }
        if (qcount) {                          // flush the sequence's remainder (< 32 entries)
            if (lane < (int)qcount) {
                const QEntry e = q[(qhead + lane) & (QCAP - 1)];
                scatter_window(sc, pl, lo_sh, mypart, (int)e.p, e.rv);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nbin; i += blockDim.x) {
        const uint32_t v = lo_sh[i];
        if (v) atomicAdd(&mypart[i], (unsigned long long)v);
    }
}

// ---- scoring -------------------------------------------------------------------------------------------------------
// reference: ScoreSeqSet::calcLogOdds, src/seq_scoring/ScoreSeqSet.cpp:25-67. Plain table, sum in ascending j from
// 0.0f: bit-identical to the reference for the same table. Only the k-mer fetch differs from k_score.
__global__ void __launch_bounds__(512)
k_score_packed(PackedView pv, Plan pl, const uint64_t* __restrict__ mops_off, const float* __restrict__ s_g,
               float* __restrict__ zoops, unsigned long long* __restrict__ z, float* __restrict__ mops, const uint32_t* __restrict__ out_idx) {
    extern __shared__ float s_sh[];
    for (uint32_t i = threadIdx.x; i < (uint32_t)pl.W * pl.Yn; i += blockDim.x) s_sh[i] = s_g[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    const int W = pl.W, K = pl.K;
    const uint32_t maskK = pl.Yn - 1;
    for (uint32_t li = warp; li < pv.nlist; li += nwarps) {
        const uint32_t n = pv.seq_ids[li];
        const uint32_t oi = out_idx[li];                   // position of this sequence in the caller's subset
        const PackedSeq sq = pv.seqs[n];
        const int L = (int)sq.L, LW1 = L - W + 1, mid = (int)sq.mid;
        const uint32_t* __restrict__ wd = pv.words + sq.word_off;
        float best = -3.402823466e+38f;
        int bestp = 0;
        for (int p0 = 0; p0 < LW1; p0 += 32) {
            const int p = p0 + lane;
            const unsigned long long w = window_word(wd, p - K);
            const bool over_n = mid >= 0 && p <= mid + K && p + W - 1 >= mid;
            float sc = 0.0f;
            int sh = 62 - 2 * K;
            uint32_t jb = 0;
            if (__any_sync(FULL, over_n)) {
                for (int j = 0; j < W; j++) {
                    uint32_t y = field(w, sh, maskK);
                    const int d = p + j - mid;
                    if (over_n && d >= 0 && d <= K) y = pv.ypatch[(uint64_t)n * (K + 1) + d];
                    sc += s_sh[jb + y];
                    sh -= 2; jb += pl.Yn;
                }
            } else {
                for (int j = 0; j < W; j++) {
                    sc += s_sh[jb + field(w, sh, maskK)];
                    sh -= 2; jb += pl.Yn;
                }
            }
            if (p < LW1) {
                if (mops) mops[mops_off[oi] + p] = sc;
                if (sc > best) { best = sc; bestp = p; }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(FULL, best, o);
            const int op = __shfl_xor_sync(FULL, bestp, o);
            if (ob > best || (ob == best && op < bestp)) { best = ob; bestp = op; }
        }
        if (lane == 0) { zoops[oi] = best; z[oi] = (unsigned long long)bestp; }
    }
}

}  // namespace bamm
