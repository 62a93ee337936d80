// packed.cuh — the fast path for the 4-letter alphabet: sequences as 2-bit packed base streams.
//
// Layout. A "regular" sequence (codes 1..4 everywhere, except optionally the structural N in the
// middle of a both-strand sequence, Sequence.cpp:10-14) is stored as 32-bit words of 16 bases, FIRST base in the
// most significant bits, preceded by two zero pad words (32 bases) and followed by six. The order-K k-mer index of the
// reference, y(i) = sum_t c(i-t) 4^t (Sequence.cpp:35-41), is then literally a bit field of the stream: for a
// window start p the 64-bit word  w = bases p-K .. p-K+31  gives  y(p+j) = (w >> (62-2K-2j)) & (4^(K+1)-1), and the
// zero pad supplies the implicit leading 'A's of k-mers that start before the sequence does. No index array is read.
//
// E-step: column groups. Consecutive motif columns are folded into one lookup over the bases they depend on,
//   tab[g][z] = prod_{j in group g} s[j][ y_j(z) ],  so a window costs G shared-memory lookups instead of W. Windows that
// cannot use whole groups (the last W-1 truncated windows, EM.cpp:167, and windows over the N whose k-mers hold rand()
// draws, Sequence.cpp:38) take a masked path that mixes whole groups with single columns of the plain table. When the
// tables of all W columns do not fit shared memory the kernel runs once per column range ("column passes").
//
// M-step: sparse or dense. r is converted to 2^40 fixed point; every window whose r rounds to 0 contributes exactly
// nothing. The E-step lists the surviving (window, value) pairs per warp; the list kernel scatters them with full
// lanes. When the posteriors are dense (list overflow) a scan kernel walks r instead. Both use unguarded 32-bit shared
// atomics on CTA-private low / high tables, optionally split by motif column over the CTAs.
#pragma once
#include <type_traits>
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
        auto check = [&](uint64_t i, uint32_t c) {
            if (c == 0) { if (i == mid) zeros_at_mid = 1; else bad = 1; }
            else if (c > 4) bad = 1;
        };
        // unaligned head and tail byte by byte, the aligned interior four codes per load: a word is suspicious when one of
        // its bytes is 0 or above 4 (bit tricks), only then its bytes are looked at one by one
        const uint64_t a0 = (base + 3) & ~3ull, a1 = (base + L) & ~3ull;     // aligned interior [a0, a1) in global offsets
        if (a0 >= a1) {
            for (uint64_t i = lane; i < L; i += 32) check(i, codes[base + i]);
        } else {
            for (uint64_t g = base + lane; g < a0; g += 32) check(g - base, codes[g]);
            for (uint64_t g = a1 + lane; g < base + L; g += 32) check(g - base, codes[g]);
            const uint32_t* __restrict__ w4 = reinterpret_cast<const uint32_t*>(codes + a0);
            const uint64_t nw = (a1 - a0) >> 2;
            for (uint64_t k = lane; k < nw; k += 32) {
                const uint32_t w = w4[k];
                const uint32_t zero = (w - 0x01010101u) & ~w & 0x80808080u;
                const uint32_t big = (((w & 0x7f7f7f7fu) + 0x7b7b7b7bu) | w) & 0x80808080u;
                if (zero | big) {
                    const uint64_t i = a0 - base + (k << 2);
                    check(i, w & 0xffu); check(i + 1, (w >> 8) & 0xffu); check(i + 2, (w >> 16) & 0xffu); check(i + 3, w >> 24);
                }
            }
        }
        bad = __any_sync(FULL, bad);
        zeros_at_mid = __any_sync(FULL, zeros_at_mid);
        if (lane == 0) kind[n] = (bad || L == 0 || L >= 0xffffff00ull) ? 0 : (zeros_at_mid ? 2 : 1);
    }
}


// patch list contract of bamm_seqset_create: positions inside the set (bit 0 of *bad otherwise) and strictly increasing (bit 1)
__global__ void k_validate_patches(const uint64_t* __restrict__ ppos, uint64_t np, uint64_t npos, uint32_t* __restrict__ bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const uint64_t p = ppos[i];
    if (p >= npos) atomicOr(bad, 1u);
    if (i && p <= ppos[i - 1]) atomicOr(bad, 2u);
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

// words of every regular sequence in the packed stream: data words + 2 pad words in front + 6 behind (0 for irregular ones)
__global__ void k_word_counts(const uint64_t* __restrict__ off, uint64_t nseq, const uint8_t* __restrict__ kind, unsigned long long* __restrict__ wc) {
    const uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nseq) return;
    const uint64_t L = off[n + 1] - off[n];
    wc[n] = kind[n] ? (L + 15) / 16 + 8 : 0ull;
}
// PackedSeq records from the exclusive scan of the word counts
__global__ void k_fill_pseq(const uint64_t* __restrict__ off, uint64_t nseq, const uint8_t* __restrict__ kind,
                            const unsigned long long* __restrict__ wscan, PackedSeq* __restrict__ seqs) {
    const uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nseq) return;
    PackedSeq q; q.word_off = 0; q.L = 0; q.mid = 0xffffffffu;
    if (kind[n]) {
        const uint64_t L = off[n + 1] - off[n];
        q.word_off = wscan[n] + 2; q.L = (uint32_t)L; q.mid = kind[n] == 2 ? (uint32_t)((L - 1) / 2) : 0xffffffffu;
    }
    seqs[n] = q;
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

// ---- column groups ------------------------------------------------------------------------------------------------
// The W motif columns are cut into G consecutive groups; group g folds its columns into ONE table lookup over the
// bases its columns depend on. Column j reads the (K+1)-mer ending at base p+j, but its value only depends on
// ctx(j)+1 of those bases, ctx(j) = max(min(j,K), K_bg): for j < K the model's v[K][y][j] is a copy of the order-j
// entry (Motif::updateV, Motif.h:126-128) and the background factor uses K_bg+1 bases (Motif.cpp:485-494). The first
// group therefore reaches only K_bg bases left of the window and can be wider for the same table size
// (C3: columns 0-4 in one 4^7 table). A host-side DP picks the cut with the fewest groups that fits shared memory.
constexpr int MAXG = 16;
struct GroupPlan {
    int W, K, G, kd;             // window word = the 32 bases starting at base p-kd (kd may be negative in later column passes)
    uint32_t Yn;                 // 4^(K+1)
    float q;
    uint32_t table_bytes;
    uint32_t shift[MAXG];        // (w >> shift) & mask4 = byte offset of the group's entry
    uint32_t shift2[MAXG];       // general mode: extra shift after the clamped funnel shift
    uint32_t mask4[MAXG];
    uint32_t base[MAXG];         // byte offset of the group's table
    uint32_t colmask[MAXG];      // bit j for every column of the group
    int col0[MAXG], ncol[MAXG], lo[MAXG];   // first column, column count, first base relative to the window start p
    // column passes: when the tables of all W columns do not fit shared memory (orders >= 5 with wide motifs) the E-step
    // runs once per pass over a column range; the partial product of the earlier passes travels through r
    uint32_t passmask;           // bit j for every column of this pass (all W columns in a single-pass plan)
    int pass_first, pass_last;
    float thr0;                  // 2^-41 (1-q) 0.999: smallest unnormalised value that can reach the M-step's threshold (norm >= 1-q)
};

// tab[g][z] = prod_{j in group g} s[j][ y_j(z) ], product in ascending j from 1.0f; z holds bases p+lo .. p+hi
// (newest base in the low digits), y_j = the part of the (K+1)-mer ending at p+j that lies inside the group's bases
// (digits outside do not influence s[j][.], see above).
template <bool LOGSUM>
__global__ void k_make_group_tables(const float* __restrict__ s, GroupPlan gp, float* __restrict__ tab) {
    const uint32_t total = gp.table_bytes >> 2;
    const uint32_t maskK = gp.Yn - 1;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int g = 0;
        while (g + 1 < gp.G && i >= (gp.base[g + 1] >> 2)) g++;
        const uint32_t z = i - (gp.base[g] >> 2);
        const int hi = gp.col0[g] + gp.ncol[g] - 1;
        float p = LOGSUM ? 0.0f : 1.0f;
        for (int t = 0; t < gp.ncol[g]; t++) {
            const int j = gp.col0[g] + t;
            const int avail = j - gp.lo[g] + 1;                          // bases of the group up to column j
            uint32_t y = (z >> (2 * (hi - j))) & maskK;
            if (avail < gp.K + 1) y &= (1u << (2 * avail)) - 1u;
            if (LOGSUM) p += s[(uint32_t)j * gp.Yn + y]; else p *= s[(uint32_t)j * gp.Yn + y];
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

struct Plan {                 // launch-invariant parameters of the packed M-step / scoring kernels
    int W, K, T, C;
    uint32_t Yn, Zn;          // 4^(K+1)
    float q;
};

// Windows whose posterior survives the M-step's fixed-point rounding, written by the E-step while it normalises:
// one region per E-step warp (no atomics, no ordering requirement: the M-step sums integers).
struct ActiveEntry { uint32_t li, p; float rv; uint32_t pad; };   // 16 bytes: one 128-bit store / load per entry
struct ActiveList {
    ActiveEntry* ent;         // entries (li = index into the packed list of the EM object, p = window start, rv = UNNORMALISED posterior)
    float* scale;             // [nlist] 1/normaliser of every list sequence: posterior = rv * scale[li]
    const uint64_t* reg_off;  // [nregions+1] first entry of every region
    uint32_t* cnt;            // [nregions] entries written at the front of the region (full windows of the fast chunks)
    uint32_t* cnt_back;       // [nregions] entries written at the back, downwards (windows of the masked chunks)
    uint32_t* overflow;       // set when a region was too small: the M-step then scans r instead
};
constexpr float FX_HALF_UNIT = 4.547473508864641e-13f;   // 2^-41: smallest r that rounds to a non-zero count

// shared-memory load at (per-lane byte offset) + (warp-uniform base): one LDS with a uniform-register base operand
__device__ __forceinline__ float lds_f32(uint32_t off, uint32_t ubase) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(off + ubase));
    return v;
}

// ---- E-step --------------------------------------------------------------------------------------------------------
// reference: EM::EStep, src/refinement/EM.cpp:139-200 (gather form). One warp per sequence, lanes =
// window starts. The G group lookups of a window are fully unrolled; the byte offset of group g's entry is a bit field
// of the 64-bit window word: FAST (every field below bit 32 after the alignment shift delta): one funnel shift + one
// mask per lookup; otherwise a clamped funnel shift plus a second shift. ONE pass: unnormalised posteriors go to r and,
// when they can reach the M-step's threshold, to the warp's region of the active list; the warp reduces the normaliser
// and stores its reciprocal per sequence. Normalisation is a multiplication by that factor wherever r is consumed.
//
// Chunks that contain windows the group tables cannot fully serve — the last W-1 truncated windows (EM.cpp:167) and the
// windows over the k-mers that hold the N's rand() draws (positions mid..mid+K, Sequence.cpp:38) — run a masked
// variant: per lane a bit mask of the groups that are whole and untouched (taken from the shared tables as usual)
// and a bit mask of the single columns that must come from the plain table in global memory.
#ifndef BAMM_E_THREADS
#define BAMM_E_THREADS 1024      // threads per CTA of the packed E-step (one CTA per SM: the tables fill shared memory)
#endif
#ifndef BAMM_E_UNROLL
#define BAMM_E_UNROLL 1          // unroll factor of the fast chunk loop
#endif
#ifndef BAMM_E_PIN
#define BAMM_E_PIN 0             // keep the per-group extraction constants in registers instead of re-reading the constant bank
#endif
template <int G, bool FAST, bool MULTI>
__global__ void __launch_bounds__(BAMM_E_THREADS, 1)
k_estep_packed(PackedView pv, const __grid_constant__ GroupPlan gp, const float* __restrict__ tab_g, const float* __restrict__ s_g /* [W][Yn] */,
               const float* __restrict__ s_rows /* [Yn][W] */, float* __restrict__ r, unsigned long long* __restrict__ scal, ActiveList al) {
    extern __shared__ float tab[];
    for (uint32_t i = threadIdx.x; i < (gp.table_bytes >> 2); i += blockDim.x) tab[i] = tab_g[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    const int W = gp.W, K = gp.K, KD = gp.kd;
    const uint32_t maskK = gp.Yn - 1;
    const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(tab);        // 32-bit shared-window address
    long long llh_fx = 0, rsum_fx = 0;
    const float one_minus_q = 1.0f - gp.q;
    uint32_t lt_mask;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
    constexpr int E_UNROLL = BAMM_E_UNROLL;
    // a window can only reach the M-step's threshold r >= 2^-41 if val >= 2^-41 (1-q): norm >= 1-q (margin for rounding)
    const float thr0 = gp.thr0;
    uint32_t c_sh[G], c_mk[G], c_ab[G], c_s2[G];
#pragma unroll
    for (int g = 0; g < G; g++) {
        c_sh[g] = gp.shift[g]; c_mk[g] = gp.mask4[g]; c_ab[g] = tab_s + gp.base[g]; c_s2[g] = gp.shift2[g];
#if BAMM_E_PIN
        if (G <= BAMM_E_PIN) asm volatile("" : "+r"(c_sh[g]), "+r"(c_mk[g]), "+r"(c_ab[g]));
#endif
    }
    const bool first = !MULTI || gp.pass_first != 0, last = !MULTI || gp.pass_last != 0;    // CTA-uniform
    const uint32_t passmask = MULTI ? gp.passmask : 0xffffffffu;
    bool emit = al.ent != nullptr && last;
    ActiveEntry* __restrict__ lreg = emit ? al.ent + al.reg_off[warp] : nullptr;      // this warp's region
    const uint32_t lcap = emit ? (uint32_t)(al.reg_off[warp + 1] - al.reg_off[warp]) : 0u;
    uint32_t lpos = 0, bpos = 0;                            // entries written so far at the front / at the back (downwards)
    for (uint32_t li = warp; li < pv.nlist; li += nwarps) {
        const uint32_t n = pv.seq_ids[li];
        const PackedSeq sq = pv.seqs[n];
        const int L = (int)sq.L, LW1 = L - W + 1;
        const int mid = (int)sq.mid;                       // -1 when there is no N
        const uint32_t* __restrict__ wseq = pv.words + sq.word_off;
        float* __restrict__ rn = r + pv.r_off[li];
        const float pos = gp.q / (float)LW1;
        // windows that need the masked path: [n0,n1) over the N's patched k-mers (clipped to the tail), [tl,LW1) the truncated
        // tail (p > L-2W+1). Segment schedule [0,b1) fast | [b1,b2) masked | [b2,b3) fast | [b3,LW1) masked; every window of
        // a fast segment is a full, untouched window. Two candidates: cuts at the exact ranges (fewest masked chunks, but a
        // partial chunk at the end of each fast segment) or cuts rounded outwards to multiples of 32 (every fast chunk full,
        // more windows in masked chunks); a masked chunk costs about SLOW_COST fast chunks.
        const int tl = min(max(L - 2 * W + 2, 0), LW1);
        int b1 = tl, b2 = tl, b3 = tl;
        if (mid >= 0) { b1 = min(max(mid - W + 1, 0), tl); b2 = min(mid + K + 1, tl); }
        {
            constexpr int SLOW_COST = 4;
            const int a1 = b1 & ~31, a3 = tl & ~31, a2 = min((b2 + 31) & ~31, a3);
            const int cost_exact = ((b1 + 31) >> 5) + ((b3 - b2 + 31) >> 5) + SLOW_COST * (((b2 - b1 + 31) >> 5) + ((LW1 - b3 + 31) >> 5));
            const int cost_align = (a1 >> 5) + ((a3 - a2) >> 5) + SLOW_COST * (((a2 - a1) >> 5) + ((LW1 - a3 + 31) >> 5));
            if (cost_align <= cost_exact) { b1 = a1; b2 = a2; b3 = a3; }
        }
        float sum = 0.0f;
#pragma unroll 1
        for (int seg = 0; seg < 4; seg++) {
            const int p0 = seg == 0 ? 0 : seg == 1 ? b1 : seg == 2 ? b2 : b3;
            const int pe = seg == 0 ? b1 : seg == 1 ? b2 : seg == 2 ? b3 : LW1;
            if (p0 >= pe) continue;
            // this lane's windows start at p0+lane + 32*chunk; its window word starts KD bases earlier. Three stream words are
            // kept and two new ones are fetched per chunk of 32 windows.
            const int bb = p0 + lane - KD;
            const uint32_t* __restrict__ wl = wseq + (bb >> 4);
            const int sft = 2 * (bb & 15);
            uint32_t t0 = wl[0], t1 = wl[1], t2 = wl[2];
            int p = p0 + lane;
            float* __restrict__ rp = rn + (L - W - p);     // r index of this lane's window; moves down 32 per chunk
            if (!(seg & 1)) {
                // fast chunk: G table lookups, no masks; `tail` = the partial last chunk of the segment (lanes p >= pe are off)
                auto fast_chunk = [&](auto tail_tag) {
                    constexpr bool TAILC = decltype(tail_tag)::value;
                    const uint32_t whi = __funnelshift_l(t1, t0, sft), wlo = __funnelshift_l(t2, t1, sft);
                    wl += 2;
                    t0 = t2; t1 = wl[1]; t2 = wl[2];
                    const bool on = !TAILC || p < pe;
                    float prod = 1.0f;
                    if (MULTI && !first && on) prod = *rp;      // product over the columns of the earlier passes
#pragma unroll
                    for (int g = 0; g < G; g++) {
                        uint32_t off;
                        if (FAST) off = __funnelshift_r(wlo, whi, c_sh[g]) & c_mk[g];
                        else      off = (__funnelshift_rc(wlo, whi, c_sh[g]) >> c_s2[g]) & c_mk[g];
                        prod *= lds_f32(off, c_ab[g]);
                    }
                    if (MULTI && !last) { if (on) *rp = prod; }
                    else {
                        const float val = on ? prod * pos : 0.0f;
                        if (on) *rp = val;
                        sum += val;
                        if (emit) {
                            const bool act = val >= thr0;
                            const uint32_t m = __ballot_sync(FULL, act);
                            if (m) {
                                const uint32_t cnt = __popc(m);
                                if (lpos + bpos + cnt > lcap) { emit = false; *al.overflow = 1u; }
                                else {
                                    if (act) *reinterpret_cast<uint4*>(lreg + lpos + __popc(m & lt_mask)) = make_uint4(li, (uint32_t)p, __float_as_uint(val), 0u);
                                    lpos += cnt;
                                }
                            }
                        }
                    }
                    rp -= 32; p += 32;
                };
#pragma unroll E_UNROLL
                for (int c = (pe - p0) >> 5; c > 0; c--) fast_chunk(std::false_type{});
                if ((pe - p0) & 31) fast_chunk(std::true_type{});
            } else {
#pragma unroll 1
                for (int c = (pe - p0 + 31) >> 5; c > 0; c--) {
                    const uint32_t whi = __funnelshift_l(t1, t0, sft), wlo = __funnelshift_l(t2, t1, sft);
                    wl += 2;
                    t0 = t2; t1 = wl[1]; t2 = wl[2];
                    const unsigned long long w = ((unsigned long long)whi << 32) | wlo;
                    const int jmax = (p < pe) ? min(W - 1, L - W - p) : -1;
                    const uint32_t valid = (jmax >= 0 ? (jmax >= 31 ? 0xffffffffu : ((2u << jmax) - 1u)) : 0u) & passmask;
                    uint32_t ncols = 0;                     // columns whose k-mer holds a rand() draw of the N
                    const bool over_n = mid >= 0 && p <= mid + K && p + W - 1 >= mid;
                    if (over_n) {
                        const int ja = max(mid - p, 0), jb = min(mid - p + K, W - 1);
                        if (jb >= ja) ncols = ((jb >= 31 ? 0xffffffffu : ((2u << jb) - 1u))) & ~((1u << ja) - 1u) & passmask;
                    }
                    uint32_t cols = valid;
                    float prod = 1.0f;
                    if (MULTI && !first && p < pe) prod = *rp;
#pragma unroll
                    for (int g = 0; g < G; g++) {
                        const uint32_t cm = gp.colmask[g];
                        const bool good = ((cm & ~valid) == 0u) && ((cm & ncols) == 0u);
                        uint32_t off;
                        if (FAST) off = __funnelshift_r(wlo, whi, c_sh[g]) & c_mk[g];
                        else      off = (__funnelshift_rc(wlo, whi, c_sh[g]) >> c_s2[g]) & c_mk[g];
                        const float v = lds_f32(off, c_ab[g]);
                        prod *= good ? v : 1.0f;
                        if (good) cols &= ~cm;
                    }
                    // the single columns are multiplied after the whole groups: a re-association of the reference's
                    // ascending-j product, far inside the 1e-5 tolerance
                    // (a) columns whose k-mer holds a draw of the N: the K+1 patched k-mers are the same for the whole
                    //     sequence, so for each of them the lanes read neighbouring elements of ONE row of the [y][j]
                    //     copy of the table — a coalesced load instead of a 32-way gather
                    if (__any_sync(FULL, over_n)) {
                        for (int d0 = 0; d0 <= K; d0 += 4) {
                            float f[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int d = d0 + u, j = mid + d - p;
                                f[u] = 1.0f;
                                if (d <= K && over_n && j >= 0 && j <= jmax && (!MULTI || ((passmask >> j) & 1u)))
                                    f[u] = __ldg(&s_rows[(uint64_t)pv.ypatch[(uint64_t)n * (K + 1) + d] * W + j]);
                            }
#pragma unroll
                            for (int u = 0; u < 4; u++) prod *= f[u];
                        }
                        cols &= ~ncols;
                    }
                    // (b) the other single columns (partial group of a truncated window, unpatched neighbours inside a
                    //     group the N broke), four per round so that their gathers are in flight together
                    while (cols) {
                        float f[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            f[u] = 1.0f;
                            if (cols) {
                                const int j = __ffs(cols) - 1;
                                cols &= cols - 1u;
                                f[u] = __ldg(&s_g[(uint32_t)j * gp.Yn + field(w, 62 - 2 * KD - 2 * j, maskK)]);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) prod *= f[u];
                    }
                    float val = 0.0f;
                    if (p < pe) {
                        if (MULTI && !last) *rp = prod;
                        else {
                            val = prod * pos;
                            *rp = val;
                            sum += val;
                        }
                    }
                    if (emit) {
                        const bool act = val >= thr0;
                        const uint32_t m = __ballot_sync(FULL, act);
                        if (m) {
                            const uint32_t cnt = __popc(m);
                            if (lpos + bpos + cnt > lcap) { emit = false; *al.overflow = 1u; }
                            else {                              // windows of the masked chunks go to the back of the region
                                bpos += cnt;
                                if (act) *reinterpret_cast<uint4*>(lreg + (lcap - bpos) + __popc(m & lt_mask)) = make_uint4(li, (uint32_t)p, __float_as_uint(val), 0u);
                            }
                        }
                    }
                    rp -= 32; p += 32;
                }
            }
        }
        if (MULTI && !last) continue;
        sum = warp_sum(sum);
        const float norm = one_minus_q + sum;
        const float rnorm = __frcp_rn(norm);
        // r keeps the unnormalised values; the factor is applied by whoever reads them (list / scan M-step,
        // k_normalise_r before r leaves the device). The tail beyond LW1 is zero since allocation.
        if (lane == 0) {
            al.scale[li] = rnorm;
            llh_fx += __double2ll_rn((double)logf(norm) * SC_SCALE_D);
            rsum_fx += __double2ll_rn((double)(sum * rnorm) * SC_SCALE_D);
        }
    }
    if (lane == 0) {
        if (llh_fx) atomicAdd(&scal[0], (unsigned long long)llh_fx);
        if (rsum_fx) atomicAdd(&scal[1], (unsigned long long)rsum_fx);
    }
    if (al.ent != nullptr && last) { al.cnt[warp] = lpos; al.cnt_back[warp] = bpos; }   // every lane holds the same counts
}

// r <- r * scale for every packed-list sequence: run before r leaves the device (bamm_em_get_r). One warp per sequence.
__global__ void k_normalise_r(PackedView pv, int W, const float* __restrict__ scale, float* __restrict__ r) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t li = warp; li < pv.nlist; li += nwarps) {
        const int LW1 = (int)pv.seqs[pv.seq_ids[li]].L - W + 1;
        const float f = scale[li];
        float* __restrict__ rn = r + pv.r_off[li];
        for (int k = lane; k < LW1; k += 32) rn[k] *= f;
    }
}

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
// guarded single step for the slow paths (windows over the N, truncated windows)
__device__ __forceinline__ void atoms_add_carry(uint32_t addr, uint32_t hioff, uint32_t xlo, uint32_t xhi) {
    const uint32_t old = atoms_add_ret(addr, xlo);
    const uint32_t h = xhi + ((uint32_t)(old + xlo) < old ? 1u : 0u);
    if (h) reds_add(addr + hioff, h);
}

// the NC columns of one window, fully unrolled: `up` holds the window's bases right-aligned so that column j0+jj's k-mer
// is the bit field at 2(NC-1-jj). MASKED = false: every column exists (full window, full split) — no per-column test at all.
// MASKED = true: columns jj > jrel_max add 0 (truncated tail windows, EM.cpp:236; columns past W in the last split).
// Batches of M_BATCH columns: first the low-word atomics of the batch, then the carries / high parts from the returned
// values, so several atomics of a lane are in flight.
constexpr int M_BATCH = 8;
template <int NC, bool MASKED>
__device__ __forceinline__ void scatter_cols(unsigned long long up, uint32_t maskK, uint32_t lo_s, uint32_t hi_off, uint32_t yn4,
                                             uint32_t xlo, uint32_t xhi, int jrel_max) {
    const uint32_t ulo = (uint32_t)up, uhi = (uint32_t)(up >> 32);
    const bool hi_nz = xhi != 0u;
#pragma unroll
    for (int g0 = 0; g0 < NC; g0 += M_BATCH) {
        uint32_t adr[M_BATCH], old[M_BATCH];
#pragma unroll
        for (int k = 0; k < M_BATCH; k++) {
            const int jj = g0 + k;
            if (jj < NC) {
                const int sh = 2 * (NC - 1 - jj);
                const uint32_t y = (sh >= 32 ? (uhi >> (sh - 32)) : __funnelshift_r(ulo, uhi, sh)) & maskK;
                adr[k] = lo_s + (uint32_t)jj * yn4 + (y << 2);
                old[k] = atoms_add_ret(adr[k], (!MASKED || jj <= jrel_max) ? xlo : 0u);
            }
        }
#pragma unroll
        for (int k = 0; k < M_BATCH; k++) {
            const int jj = g0 + k;
            if (jj < NC) {
                const bool on = !MASKED || jj <= jrel_max;
                const bool carry = on && (uint32_t)~old[k] < xlo;          // old + xlo wrapped
                if (carry || (on && hi_nz)) reds_add(adr[k] + hi_off, xhi + (carry ? 1u : 0u));
            }
        }
    }
}

// slow path, one window per lane with a run-time column loop: windows over the N, whose k-mers at positions mid..mid+K
// hold rand() draws (Sequence.cpp:38) and come from the patch list, and the truncated last W-1 windows (EM.cpp:236)
__device__ __forceinline__ void scatter_window_slow(const SeqCtx& sc, const Plan& pl, uint32_t lo_s, uint32_t hi_off, int j0, int nc,
                                                    int p, unsigned long long X) {
    if (X == 0) return;
    const int W = pl.W, K = pl.K;
    const unsigned long long w = window_word(sc.wd, p + j0 - K); // bases p+j0-K .. p+j0-K+31
    const uint32_t maskK = pl.Yn - 1;
    const int jmax = min(min(W - 1, sc.L - W - p), j0 + nc - 1);
    const uint32_t xlo = (uint32_t)X, xhi = (uint32_t)(X >> 32);
    for (int j = j0; j <= jmax; j++) {
        uint32_t y = field(w, 62 - 2 * K - 2 * (j - j0), maskK);
        const int d = p + j - sc.mid;
        if (sc.mid >= 0 && d >= 0 && d <= K) y = sc.yp[d];
        atoms_add_carry(lo_s + (((uint32_t)(j - j0) * pl.Yn + y) << 2), hi_off, xlo, xhi);
    }
}

// CTA-private tables -> this CTA's slice of the partial-count array (only the CTA's own columns; the rest stays zero)
// Replicas: for tiny tables (orders 0 and 1: 4 or 16 bins per column) most lanes of a warp hit the same few addresses and
// the atomics serialise; the CTA then keeps nrep copies of both tables (lane l uses copy l % nrep, copies rstride words
// apart with rstride = 1 mod 32 so that equal bins of different copies fall into different banks) and sums them here.
struct MTables { uint32_t nrep, rstride; };      // rstride >= NC * Yn
__device__ __forceinline__ void flush_cols(const uint32_t* lo_sh, MTables mt, uint32_t nb, uint32_t Yn, int j0, int W,
                                           unsigned long long* __restrict__ mypart) {
    const uint32_t* hi_sh = lo_sh + mt.nrep * mt.rstride;
    const uint32_t lim = (uint32_t)max(0, min((int)(nb / Yn), W - j0)) * Yn;
    for (uint32_t i = threadIdx.x; i < lim; i += blockDim.x) {
        unsigned long long lo = 0, hi = 0;
        for (uint32_t rp = 0; rp < mt.nrep; rp++) { lo += lo_sh[rp * mt.rstride + i]; hi += hi_sh[rp * mt.rstride + i]; }
        const unsigned long long v = lo + (hi << 32);
        if (v) mypart[(uint32_t)j0 * Yn + i] = v;
    }
}

// M-step from the E-step's active list: every lane scatters one listed window.
template <int NC>
__global__ void __launch_bounds__(1024, 1)
k_mstep_list_w(PackedView pv, Plan pl, ActiveList al, uint32_t nregions, int nsplit, MTables mt, unsigned long long* __restrict__ part /* [gridDim.x][W*Yn] */) {
    extern __shared__ uint32_t smem_u32[];
    if (*al.overflow != 0u) return;                                        // k_mstep_scan_w scans r instead
    const uint32_t nb = (uint32_t)NC * pl.Yn;
    uint32_t* lo_sh = smem_u32;
    for (uint32_t i = threadIdx.x; i < 2 * mt.nrep * mt.rstride; i += blockDim.x) smem_u32[i] = 0u;
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
    const int nc_valid = min(NC, W - j0);
    const bool split_full = nc_valid == NC;                                // CTA-uniform
    for (uint32_t rg = warp; rg < nregions; rg += nwarps) {
        const ActiveEntry* __restrict__ ent = al.ent + al.reg_off[rg];
        // front of the region: full windows away from the N, written by the E-step's fast chunks — the unguarded path
        const uint32_t cnt = al.cnt[rg];
        for (uint32_t e0 = 0; e0 < cnt; e0 += 32) {                        // warp-uniform batches
            const uint32_t e = e0 + lane;
            unsigned long long X = 0ull, up = 0ull;                        // lanes past the end add 0 to column bins
            int jrel_max = -1;
            if (e < cnt) {
                const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(ent + e));
                const uint32_t li = raw.x;
                const float rv = __uint_as_float(raw.z) * al.scale[li];
                X = __float2ull_rn(rv * FX_SCALE_F);
                up = window_word(pv.words + pv.seqs[pv.seq_ids[li]].word_off, (int)raw.y + j0 - K) >> ralign;
                jrel_max = NC;
            }
            if (split_full) scatter_cols<NC, false>(up, maskK, lo_s, hi_off, yn4, (uint32_t)X, (uint32_t)(X >> 32), 0);
            else scatter_cols<NC, true>(up, maskK, lo_s, hi_off, yn4, (uint32_t)X, (uint32_t)(X >> 32), min(jrel_max, nc_valid - 1));
        }
        // back of the region (filled downwards): windows of the E-step's masked chunks — truncated tail windows, windows
        // over the N, and the full windows that share their chunks
        const uint32_t cntb = al.cnt_back[rg];
        const ActiveEntry* __restrict__ entb = al.ent + al.reg_off[rg + 1] - cntb;
        for (uint32_t e0 = 0; e0 < cntb; e0 += 32) {
            const uint32_t e = e0 + lane;
            unsigned long long X = 0ull, up = 0ull;
            int jrel_max = -1;
            if (e < cntb) {
                const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(entb + e));
                const uint32_t li = raw.x;
                const int p = (int)raw.y;
                const float rv = __uint_as_float(raw.z) * al.scale[li];
                const uint32_t n = pv.seq_ids[li];
                const PackedSeq sq = pv.seqs[n];
                const int L = (int)sq.L, mid = (int)sq.mid;
                X = __float2ull_rn(rv * FX_SCALE_F);
                if (mid >= 0 && p <= mid + K && p + W - 1 >= mid) {        // patched k-mers: slow path
                    SeqCtx sc; sc.wd = pv.words + sq.word_off; sc.yp = pv.ypatch + (uint64_t)n * (K + 1); sc.L = L; sc.mid = mid;
                    scatter_window_slow(sc, pl, lo_s, hi_off, j0, NC, p, X);
                    X = 0ull;
                } else {
                    up = window_word(pv.words + sq.word_off, p + j0 - K) >> ralign;
                    jrel_max = min(min(W - 1, L - W - p) - j0, nc_valid - 1);
                }
            }
            __syncwarp();
            scatter_cols<NC, true>(up, maskK, lo_s, hi_off, yn4, (uint32_t)X, (uint32_t)(X >> 32), jrel_max);
        }
    }
    __syncthreads();
    flush_cols(lo_sh, mt, nb, pl.Yn, j0, W, part + (uint64_t)blockIdx.x * ((uint64_t)W * pl.Yn));
}

// M-step from r itself: one warp per sequence, lanes = 32 consecutive window starts, the packed stream is followed with the
// E-step's rolling three-word fetch. Chunks without a surviving posterior are skipped after one vote.
// only_if: nullptr, or a device flag — the kernel runs only when it is non-zero (active list overflowed).
// scale: nullptr when r is already normalised, else 1/normaliser per list sequence.
template <int NC>
__global__ void __launch_bounds__(1024, 1)
k_mstep_scan_w(PackedView pv, Plan pl, const float* __restrict__ r, const float* __restrict__ scale, const uint32_t* __restrict__ only_if,
               int nsplit, MTables mt, unsigned long long* __restrict__ part /* [gridDim.x][W*Yn] */) {
    extern __shared__ uint32_t smem_u32[];
    if (only_if != nullptr && *only_if == 0u) return;                      // k_mstep_list_w did the work
    const uint32_t nb = (uint32_t)NC * pl.Yn;
    uint32_t* lo_sh = smem_u32;
    for (uint32_t i = threadIdx.x; i < 2 * mt.nrep * mt.rstride; i += blockDim.x) smem_u32[i] = 0u;
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
                    scatter_window_slow(sc, pl, lo_s, hi_off, j0, NC, p, X);
                    X = 0ull;
                }
                __syncwarp();
                scatter_cols<NC, true>(up, maskK, lo_s, hi_off, yn4, (uint32_t)X, (uint32_t)(X >> 32), min(min(W - 1, L - W - p) - j0, nc_valid - 1));
            } else {
                scatter_cols<NC, false>(up, maskK, lo_s, hi_off, yn4, (uint32_t)X, (uint32_t)(X >> 32), 0);
            }
        }
    }
    __syncthreads();
    flush_cols(lo_sh, mt, nb, pl.Yn, j0, W, part + (uint64_t)blockIdx.x * ((uint64_t)W * pl.Yn));
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
        const uint32_t oi = out_idx ? out_idx[li] : li;    // position of this sequence in the caller's subset
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

// ZOOPS-only scoring (FDR's default consumer, Global.cpp:31): only the maximum window score and its first position are
// wanted, but bit-identical to the reference's ascending-j fp32 sum. Every window gets a CHEAP score from column-group
// tables (G lookups; sums of the same table entries in another association, so |cheap - exact| <= eps, see the host
// side). A window can only hold the true maximum if its cheap score is within 2*eps of the maximum cheap score of the
// sequence (cheap(p*) >= M - eps >= max_cheap - 2 eps), so each lane defers
// its latest qualifying window and re-scores it exactly (plain table, ascending j) only if it still qualifies at the end
// of the sequence — or when a second qualifying window of the same lane displaces it. Windows over the N's patched
// k-mers are re-scored at once. About one exact evaluation per sequence instead of one per running-maximum record.
template <int G, bool FAST>
__global__ void __launch_bounds__(1024, 1)
k_score_zoops_packed(PackedView pv, const __grid_constant__ GroupPlan gp, const float* __restrict__ tab_g, const float* __restrict__ s_g,
                     float two_eps, float* __restrict__ zoops, unsigned long long* __restrict__ z, const uint32_t* __restrict__ out_idx) {
    extern __shared__ float smem_f[];
    float* tab = smem_f;                                                   // group tables
    float* s_sh = smem_f + (gp.table_bytes >> 2);                          // plain [j][y] log-odds table
    const uint32_t nplain = (uint32_t)gp.W * gp.Yn;
    for (uint32_t i = threadIdx.x; i < (gp.table_bytes >> 2); i += blockDim.x) tab[i] = tab_g[i];
    for (uint32_t i = threadIdx.x; i < nplain; i += blockDim.x) s_sh[i] = s_g[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    const int W = gp.W, K = gp.K, KD = gp.kd;
    const uint32_t maskK = gp.Yn - 1;
    const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(tab);
    const int lane_word = (lane - KD) >> 4;
    const int sft = 2 * ((lane - KD) & 15);
    constexpr float NEG_INF = -3.402823466e+38f;
    uint32_t c_sh[G], c_mk[G], c_ab[G], c_s2[G];
#pragma unroll
    for (int g = 0; g < G; g++) { c_sh[g] = gp.shift[g]; c_mk[g] = gp.mask4[g]; c_ab[g] = tab_s + gp.base[g]; c_s2[g] = gp.shift2[g]; }
    for (uint32_t li = warp; li < pv.nlist; li += nwarps) {
        const uint32_t n = pv.seq_ids[li];
        const uint32_t oi = out_idx ? out_idx[li] : li;
        const PackedSeq sq = pv.seqs[n];
        const int L = (int)sq.L, LW1 = L - W + 1, mid = (int)sq.mid;
        const uint32_t* __restrict__ wseq = pv.words + sq.word_off;
        const uint32_t* __restrict__ wl = wseq + lane_word;
        uint32_t t0 = wl[0], t1 = wl[1], t2 = wl[2];
        float best = NEG_INF, run_max = NEG_INF;
        int bestp = 0;
        float held = NEG_INF;                                              // cheap score of this lane's deferred window
        int heldp = -1;
        // exact score of window p (ascending j from 0.0f, ScoreSeqSet.cpp:49-54), folded into the lane's best
        auto rescore = [&](int p) {
            const unsigned long long w = window_word(wseq, p - KD);
            const bool over_n = mid >= 0 && p <= mid + K && p + W - 1 >= mid;
            float sc = 0.0f;
            int sh = 62 - 2 * KD;
            uint32_t jb = 0;
            for (int j = 0; j < W; j++) {
                uint32_t y = field(w, sh, maskK);
                const int d = p + j - mid;
                if (over_n && d >= 0 && d <= K) y = pv.ypatch[(uint64_t)n * (K + 1) + d];
                sc += s_sh[jb + y];
                sh -= 2; jb += gp.Yn;
            }
            if (sc > best || (sc == best && p < bestp)) { best = sc; bestp = p; }
        };
        const int nch = (LW1 + 31) >> 5;
        for (int c = 0; c < nch; c++) {
            const int p = (c << 5) + lane;
            const uint32_t whi = __funnelshift_l(t1, t0, sft), wlo = __funnelshift_l(t2, t1, sft);
            wl += 2;
            t0 = t2; t1 = wl[1]; t2 = wl[2];
            float cheap = 0.0f;
#pragma unroll
            for (int g = 0; g < G; g++) {
                uint32_t off;
                if (FAST) off = __funnelshift_r(wlo, whi, c_sh[g]) & c_mk[g];
                else      off = (__funnelshift_rc(wlo, whi, c_sh[g]) >> c_s2[g]) & c_mk[g];
                cheap += lds_f32(off, c_ab[g]);
            }
            const bool on = p < LW1;
            const bool over_n = on && mid >= 0 && p <= mid + K && p + W - 1 >= mid;
            // the lane's OWN running maximum is a valid (lower) stand-in for the warp's while scanning: it only lets a few
            // more windows qualify; the warp-wide maximum is taken once per sequence
            if (over_n) rescore(p);                                        // group tables do not know the patched k-mers
            else if (on) {
                run_max = fmaxf(run_max, cheap);
                const float thr = run_max - two_eps;
                if (cheap >= thr) {
                    if (heldp >= 0 && held >= thr) rescore(heldp);         // displaced while still qualifying (rare)
                    held = cheap; heldp = p;
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) run_max = fmaxf(run_max, __shfl_xor_sync(FULL, run_max, o));
        if (heldp >= 0 && held >= run_max - two_eps) rescore(heldp);
        __syncwarp();
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(FULL, best, o);
            const int op = __shfl_xor_sync(FULL, bestp, o);
            if (ob > best || (ob == best && op < bestp)) { best = ob; bestp = op; }
        }
        if (lane == 0) { zoops[oi] = best; z[oi] = (unsigned long long)bestp; }
    }
}

}  // namespace bamm
