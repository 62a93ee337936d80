// mstep.cuh — packed M-step kernels (active list / scan of r). Included by launch_mstep.cu only.
#pragma once
#include "common.cuh"

namespace bamm {

// ---- M-step --------------------------------------------------------------------------------------------------------
// reference: EM::MStep accumulation, src/refinement/EM.cpp:230-243 (gather form).
// Every posterior is converted to 2^-40 fixed point; a window whose posterior rounds to 0 contributes exactly nothing.
// Each lane scatters one window's value into the bins (j, y(p+j)) it touches with native 32-bit shared atomics on two
// CTA-private tables: the low 32 bits of the sums, and a second table that collects the high parts (r >= 2^-8) and the
// wrap-arounds of the low words — no global atomics while scattering, integer sums only, so the counts are
// bit-reproducible for any grid, schedule, kernel variant or GPU count.
//
// Column split: a CTA owns NC consecutive motif columns [j0, j0+NC), j0 = (blockIdx.x % nsplit) * NC, so that its two
// tables (NC * 4^(K+1) * 8 bytes) fit shared memory for every order the packed stream supports; the CTAs of one split
// together walk all windows. nsplit = 1, NC = W whenever the whole table fits (K <= 4 at W = 20).
//
// Two sources of windows: the E-step's active list (sparse posteriors: full lanes, no scan of r) and a scan of r
// (dense posteriors, list switched off or overflowed).
struct SeqCtx { const uint32_t* wd; const uint16_t* yp; int L, mid; };

// Shared atomics cannot be predicated on sm_100a (ptxas wraps every guarded ATOMS in a BSSY / BRA / BSYNC region), so
// the scatter of a full window is written without guards: the low-word atomic of every column is unconditional — lanes
// without a value add 0, which changes nothing — and only the rare add to the high table (r >= 2^-8, or a wrap-around of
// the low word detected from the returned old value) sits behind a branch.
__device__ __forceinline__ uint32_t atoms_add_ret(uint32_t addr, uint32_t x) {
    uint32_t o;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(addr), "r"(x) : "memory");
    return o;
}
__device__ __forceinline__ void reds_add(uint32_t addr, uint32_t h) {
    asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(h) : "memory");
}
// Where the high parts go: the CTA's second shared-memory table, or — when that table is given up so that twice the columns
// fit one CTA (orders >= 5: fewer passes over the list) — a 64-bit count table in global memory whose bins [j0 Yn, ...) mirror
// the shared low-word table (one copy: nrep = 1 there). Rare either way: posteriors >= 2^-8 and the wrap-arounds of a low
// word. In that mode all CTAs share ONE global table (the flush adds the low words with atomics as well): at 4^6 and more
// rows per column a partial table per CTA would cost more to clear and to sum than the M-step itself.
// HG is a template parameter of the kernels: the variant with the shared-memory table carries no pointer (it runs at the
// register limit of 1024 threads per CTA).
template <bool HG> struct HiDst;
template <> struct HiDst<false> { __device__ __forceinline__ void set(unsigned long long*, uint32_t) {} };
template <> struct HiDst<true> {
    unsigned long long* ghi; uint32_t lo_s;
    __device__ __forceinline__ void set(unsigned long long* g, uint32_t l) { ghi = g; lo_s = l; }
};
__device__ __forceinline__ void hi_add(uint32_t addr, uint32_t hioff, uint32_t h, const HiDst<false>&) { reds_add(addr + hioff, h); }
__device__ __forceinline__ void hi_add(uint32_t addr, uint32_t, uint32_t h, const HiDst<true>& hd) {
    atomicAdd(hd.ghi + ((addr - hd.lo_s) >> 2), (unsigned long long)h << 32);
}
// guarded single step for the slow paths (windows over the N, truncated windows)
template <bool HG>
__device__ __forceinline__ void atoms_add_carry(uint32_t addr, uint32_t hioff, uint32_t xlo, uint32_t xhi, const HiDst<HG>& hd) {
    const uint32_t old = atoms_add_ret(addr, xlo);
    const uint32_t h = xhi + ((uint32_t)(old + xlo) < old ? 1u : 0u);
    if (h) hi_add(addr, hioff, h, hd);
}

// the NC columns of one window, fully unrolled: `up` holds the window's bases right-aligned so that column j0+jj's k-mer
// is the bit field at 2(NC-1-jj). MASKED = false: every column exists (full window, full split) — no per-column test at all.
// MASKED = true: columns jj > jrel_max add 0 (truncated tail windows, EM.cpp:236; columns past W in the last split).
// Batches of M_BATCH columns: first the low-word atomics of the batch, then the carries / high parts from the returned
// values, so several atomics of a lane are in flight.
constexpr int M_BATCH = 8;
template <int NC, bool MASKED, bool HG>
__device__ __forceinline__ void scatter_cols(unsigned long long up, uint32_t maskK, uint32_t lo_s, uint32_t hi_off, uint32_t yn4,
                                             uint32_t xlo, uint32_t xhi, const HiDst<HG>& hd, int jrel_max, int jmax_u = 31 /* warp-uniform bound of jrel_max */) {
    const uint32_t ulo = (uint32_t)up, uhi = (uint32_t)(up >> 32);
    const bool hi_nz = xhi != 0u;
#pragma unroll
    for (int g0 = 0; g0 < NC; g0 += M_BATCH) {
        uint32_t adr[M_BATCH], old[M_BATCH];
#pragma unroll
        for (int k = 0; k < M_BATCH; k++) {
            const int jj = g0 + k;
            if (jj < NC && (!MASKED || jj <= jmax_u)) {                    // columns no lane of the warp has are skipped (uniform)
                const int sh = 2 * (NC - 1 - jj);
                const uint32_t y = (sh >= 32 ? (uhi >> (sh - 32)) : __funnelshift_r(ulo, uhi, sh)) & maskK;
                adr[k] = lo_s + (uint32_t)jj * yn4 + (y << 2);
                old[k] = atoms_add_ret(adr[k], (!MASKED || jj <= jrel_max) ? xlo : 0u);
            }
        }
#pragma unroll
        for (int k = 0; k < M_BATCH; k++) {
            const int jj = g0 + k;
            if (jj < NC && (!MASKED || jj <= jmax_u)) {
                const bool on = !MASKED || jj <= jrel_max;
                const bool carry = on && (uint32_t)~old[k] < xlo;          // old + xlo wrapped
                if (carry || (on && hi_nz)) hi_add(adr[k], hi_off, xhi + (carry ? 1u : 0u), hd);
            }
        }
    }
}

// slow path, one window per lane with a run-time column loop: windows over the N, whose k-mers at positions mid..mid+K
// hold rand() draws (Sequence.cpp:38) and come from the patch list, and the truncated last W-1 windows (EM.cpp:236)
template <bool HG>
__device__ __forceinline__ void scatter_window_slow(const SeqCtx& sc, const Plan& pl, uint32_t lo_s, uint32_t hi_off, int j0, int nc,
                                                    int p, unsigned long long X, const HiDst<HG>& hd) {
    if (X == 0) return;
    const int W = pl.W, K = pl.K;
    const unsigned long long w = window_word(sc.wd, p + j0 - K); // bases p+j0-K .. p+j0-K+31
    const uint32_t maskK = pl.Yn - 1;
    const int jmax = min(min(W - 1, sc.L - W - p), j0 + nc - 1);  // sc.L - W - p: the window's last column (EM.cpp:236)
    const uint32_t xlo = (uint32_t)X, xhi = (uint32_t)(X >> 32);
    for (int j = j0; j <= jmax; j++) {
        uint32_t y = field(w, 62 - 2 * K - 2 * (j - j0), maskK);
        const int d = p + j - sc.mid;
        if (sc.mid >= 0 && d >= 0 && d <= K) y = sc.yp[d];
        atoms_add_carry(lo_s + (((uint32_t)(j - j0) * pl.Yn + y) << 2), hi_off, xlo, xhi, hd);
    }
}

// CTA-private tables -> this CTA's slice of the partial-count array (only the CTA's own columns; the rest stays zero)
// Replicas: for tiny tables (orders 0 and 1: 4 or 16 bins per column) most lanes of a warp hit the same few addresses and
// the atomics serialise; the CTA then keeps nrep copies of both tables (lane l uses copy l % nrep, copies rstride words
// apart with rstride = 1 mod 32 so that equal bins of different copies fall into different banks) and sums them here.
template <bool HG>
__device__ __forceinline__ void flush_cols(const uint32_t* lo_sh, MTables mt, uint32_t nb, uint32_t Yn, int j0, int W,
                                           unsigned long long* __restrict__ mypart) {
    const uint32_t* hi_sh = lo_sh + mt.nrep * mt.rstride;
    const uint32_t lim = (uint32_t)max(0, min((int)(nb / Yn), W - j0)) * Yn;
    for (uint32_t i = threadIdx.x; i < lim; i += blockDim.x) {
        if (HG) {                                         // the high parts are in mypart already (global atomics of this CTA)
            const uint32_t lo = lo_sh[i];
            if (lo) atomicAdd(&mypart[(uint32_t)j0 * Yn + i], (unsigned long long)lo);
        } else {
            unsigned long long lo = 0, hi = 0;
            for (uint32_t rp = 0; rp < mt.nrep; rp++) { lo += lo_sh[rp * mt.rstride + i]; hi += hi_sh[rp * mt.rstride + i]; }
            const unsigned long long v = lo + (hi << 32);
            if (v) mypart[(uint32_t)j0 * Yn + i] = v;
        }
    }
}

// M-step from the E-step's active list: every lane scatters one listed window.
template <int NC, bool HG>
__global__ void __launch_bounds__(1024, 1)
k_mstep_list_w(PackedView pv, Plan pl, ActiveList al, uint32_t nregions, int nsplit, MTables mt, unsigned long long* __restrict__ part /* [gridDim.x][W*Yn]; HG: one [W*Yn] table */) {
    extern __shared__ uint32_t smem_u32[];
    if (*al.overflow != 0u) return;                                        // k_mstep_scan_w scans r instead
    const uint32_t nb = (uint32_t)NC * pl.Yn;
    uint32_t* lo_sh = smem_u32;
    for (uint32_t i = threadIdx.x; i < (HG ? 1u : 2u) * mt.nrep * mt.rstride; i += blockDim.x) smem_u32[i] = 0u;
    __syncthreads();
    const int split = (int)(blockIdx.x % (uint32_t)nsplit), j0 = split * NC;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x / (uint32_t)nsplit) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t nwarps = (gridDim.x / (uint32_t)nsplit) * (blockDim.x >> 5);
    const int W = pl.W, K = pl.K;
    const uint32_t maskK = pl.Yn - 1;
    const int ralign = 62 - 2 * (K + NC - 1);                              // word = bases p+j0-K ..: column j0+NC-1's last base lowest
    const uint32_t lo_s = (uint32_t)__cvta_generic_to_shared(lo_sh) + ((uint32_t)lane & (mt.nrep - 1u)) * mt.rstride * 4u;   // this lane's copy
    const uint32_t hi_off = mt.nrep * mt.rstride * 4u, yn4 = pl.Yn * 4u;
    HiDst<HG> hd; hd.set(part + (uint64_t)j0 * pl.Yn, lo_s);            // HG: one global table for all CTAs (atomics only)
    const int nc_valid = min(NC, W - j0);
    const bool split_full = nc_valid == NC;                                // CTA-uniform
    const uint4 none = make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t rg = warp; rg < nregions; rg += nwarps) {
        // An entry carries the word offset of its sequence and the window start (common.cuh, ActiveEntry): the only dependent
        // loads are the three stream words and the sequence's normaliser; the next batch of entries is requested before this
        // one is scattered.
        const ActiveEntry* __restrict__ ent = al.ent + al.reg_off[rg];
        // front of the region: full windows away from the N, written by the E-step's fast chunks — the unguarded path
        const uint32_t cnt = al.cnt[rg];
        uint4 nxt = lane < cnt ? __ldcs(reinterpret_cast<const uint4*>(ent + lane)) : none;
        for (uint32_t e0 = 0; e0 < cnt; e0 += 32) {                        // warp-uniform batches
            const uint4 raw = nxt;
            const bool on = e0 + lane < cnt;
            nxt = e0 + 32 + lane < cnt ? __ldcs(reinterpret_cast<const uint4*>(ent + e0 + 32 + lane)) : none;
            unsigned long long X = 0ull, up = 0ull;                        // lanes past the end add 0 to column bins
            if (on) {
                const float rv = __uint_as_float(raw.z) * al.scale[raw.w];
                X = __float2ull_rn(rv * FX_SCALE_F);
                up = window_word(pv.words + raw.x, (int)(raw.y & ACT_P_MASK) + j0 - K) >> ralign;
            }
            if (split_full) scatter_cols<NC, false, HG>(up, maskK, lo_s, hi_off, yn4, (uint32_t)X, (uint32_t)(X >> 32), hd, 0);
            else scatter_cols<NC, true, HG>(up, maskK, lo_s, hi_off, yn4, (uint32_t)X, (uint32_t)(X >> 32), hd, on ? nc_valid - 1 : -1);
        }
        // back of the region (filled downwards): windows of the E-step's masked evaluation — truncated tail windows, windows
        // over the N, and the full windows that share their chunks
        const uint32_t cntb = al.cnt_back[rg];
        const ActiveEntry* __restrict__ entb = al.ent + al.reg_off[rg + 1] - cntb;
        nxt = lane < cntb ? __ldcs(reinterpret_cast<const uint4*>(entb + lane)) : none;
        for (uint32_t e0 = 0; e0 < cntb; e0 += 32) {
            const uint4 raw = nxt;
            nxt = e0 + 32 + lane < cntb ? __ldcs(reinterpret_cast<const uint4*>(entb + e0 + 32 + lane)) : none;
            unsigned long long X = 0ull, up = 0ull;
            int jrel_max = -1;
            if (e0 + lane < cntb) {
                const int p = (int)(raw.y & ACT_P_MASK), jm = (int)((raw.y >> ACT_JMAX_SHIFT) & 31u);
                const float rv = __uint_as_float(raw.z) * al.scale[raw.w];
                X = __float2ull_rn(rv * FX_SCALE_F);
                if (raw.y >> 31) {                                         // patched k-mers: slow path
                    const uint32_t n = pv.seq_ids[raw.w];
                    const PackedSeq sq = pv.seqs[n];
                    SeqCtx sc; sc.wd = pv.words + sq.word_off; sc.yp = pv.ypatch + (uint64_t)n * (K + 1); sc.L = (int)sq.L; sc.mid = (int)sq.mid;
                    scatter_window_slow(sc, pl, lo_s, hi_off, j0, NC, p, X, hd);
                    X = 0ull;
                } else {
                    up = window_word(pv.words + raw.x, p + j0 - K) >> ralign;
                    jrel_max = min(jm - j0, nc_valid - 1);
                }
            }
            __syncwarp();
            // the truncated windows arrive sorted by their last column (k_emasked lists them window index by window index)
            scatter_cols<NC, true, HG>(up, maskK, lo_s, hi_off, yn4, (uint32_t)X, (uint32_t)(X >> 32), hd, jrel_max, __reduce_max_sync(FULL, jrel_max));
        }
    }
    __syncthreads();
    flush_cols<HG>(lo_sh, mt, nb, pl.Yn, j0, W, HG ? part : part + (uint64_t)blockIdx.x * ((uint64_t)W * pl.Yn));
}

// M-step from r itself: one warp per sequence, lanes = 32 consecutive window starts, the packed stream is followed with the
// E-step's rolling three-word fetch. Chunks without a surviving posterior are skipped after one vote.
// only_if: nullptr, or a device flag — the kernel runs only when it is non-zero (active list overflowed).
// scale: nullptr when r is already normalised, else 1/normaliser per list sequence.
template <int NC, bool HG>
__global__ void __launch_bounds__(1024, 1)
k_mstep_scan_w(PackedView pv, Plan pl, const float* __restrict__ r, const float* __restrict__ scale, const uint32_t* __restrict__ only_if,
               int nsplit, MTables mt, unsigned long long* __restrict__ part /* [gridDim.x][W*Yn]; HG: one [W*Yn] table */) {
    extern __shared__ uint32_t smem_u32[];
    if (only_if != nullptr && *only_if == 0u) return;                      // k_mstep_list_w did the work
    const uint32_t nb = (uint32_t)NC * pl.Yn;
    uint32_t* lo_sh = smem_u32;
    for (uint32_t i = threadIdx.x; i < (HG ? 1u : 2u) * mt.nrep * mt.rstride; i += blockDim.x) smem_u32[i] = 0u;
    __syncthreads();
    const int split = (int)(blockIdx.x % (uint32_t)nsplit), j0 = split * NC;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x / (uint32_t)nsplit) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t nwarps = (gridDim.x / (uint32_t)nsplit) * (blockDim.x >> 5);
    const int W = pl.W, K = pl.K;
    const uint32_t maskK = pl.Yn - 1;
    const int ralign = 62 - 2 * (K + NC - 1);
    const uint32_t lo_s = (uint32_t)__cvta_generic_to_shared(lo_sh) + ((uint32_t)lane & (mt.nrep - 1u)) * mt.rstride * 4u;   // this lane's copy
    const uint32_t hi_off = mt.nrep * mt.rstride * 4u, yn4 = pl.Yn * 4u;
    HiDst<HG> hd; hd.set(part + (uint64_t)j0 * pl.Yn, lo_s);            // HG: one global table for all CTAs (atomics only)
    const int nc_valid = min(NC, W - j0);
    const bool split_full = nc_valid == NC;                                // CTA-uniform
    const int lane_word = (lane + j0 - K) >> 4;                            // this lane's words start at bases lane+j0-K + 32*chunk
    const int sft = 2 * ((lane + j0 - K) & 15);
    for (uint32_t li = warp; li < pv.nlist; li += nwarps) {
        const uint32_t n = pv.seq_ids[li];
        const PackedSeq sq = pv.seqs[n];
        const int L = (int)sq.L, LW1 = L - W + 1, mid = (int)sq.mid;
        const uint32_t* __restrict__ wl = pv.words + sq.word_off + lane_word;
        uint32_t t0 = wl[0], t1 = wl[1], t2 = wl[2];
        const float* __restrict__ rp = r + pv.r_off[li] + (L - W - lane);  // r index of this lane's window; moves down 32 per chunk
        const float sc_f = scale ? scale[li] : 1.0f;                       // x * 1.0f is exact: one code path for both states of r
        const int nch = (LW1 + 31) >> 5;
        int cn0 = nch, cn1 = -1;                                           // chunks that hold windows over the N
        if (mid >= 0) { cn0 = max(mid - W + 1, 0) >> 5; cn1 = (mid + K) >> 5; }
        const int ctail = max(L - 2 * W + 2, 0) >> 5;                      // first chunk with a truncated window (p > L-2W+1)
        float rv_next = lane < LW1 ? __ldcs(rp) : 0.0f;
        for (int c = 0; c < nch; c++) {
            const int p = (c << 5) + lane;
            const uint32_t whi = __funnelshift_l(t1, t0, sft), wlo = __funnelshift_l(t2, t1, sft);
            wl += 2;
            t0 = t2; t1 = wl[1]; t2 = wl[2];
            const float rv = rv_next * sc_f;
            rp -= 32;
            rv_next = (p + 32 < LW1) ? __ldcs(rp) : 0.0f;
            unsigned long long X = __float2ull_rn(rv * FX_SCALE_F);
            if (!__any_sync(FULL, X != 0ull)) continue;
            const unsigned long long up = (((unsigned long long)whi << 32) | wlo) >> ralign;
            if ((c >= cn0 && c <= cn1) || c >= ctail || !split_full) {     // windows over the N / truncated windows / partial split
                if (mid >= 0 && p <= mid + K && p + W - 1 >= mid) {
                    SeqCtx sc; sc.wd = pv.words + sq.word_off; sc.yp = pv.ypatch + (uint64_t)n * (K + 1); sc.L = L; sc.mid = mid;
                    scatter_window_slow(sc, pl, lo_s, hi_off, j0, NC, p, X, hd);
                    X = 0ull;
                }
                __syncwarp();
                scatter_cols<NC, true, HG>(up, maskK, lo_s, hi_off, yn4, (uint32_t)X, (uint32_t)(X >> 32), hd, min(min(W - 1, L - W - p) - j0, nc_valid - 1));
            } else {
                scatter_cols<NC, false, HG>(up, maskK, lo_s, hi_off, yn4, (uint32_t)X, (uint32_t)(X >> 32), hd, 0);
            }
        }
    }
    __syncthreads();
    flush_cols<HG>(lo_sh, mt, nb, pl.Yn, j0, W, HG ? part : part + (uint64_t)blockIdx.x * ((uint64_t)W * pl.Yn));
}

}  // namespace bamm
