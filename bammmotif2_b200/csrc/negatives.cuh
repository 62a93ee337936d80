// negatives.cuh — negative (background) sequence sampling on the device.
//
// reference: SeqGenerator::sample_bgseqset_by_fold / calculate_kmer_frequency / rescale_kmer_frequency /
// bgseq_on_rescaled_v, src/seq_generator/SeqGenerator.cpp:63-206, 285-348. For every positive sequence ("template")
// the reference rescales the set-wide order-2 k-mer model with the template's own k-mer counts and samples `fold`
// sequences of the template's stored length from it, one libc rand() draw per base, all on one host thread (10^10
// draws for BASELINE config 4). The sampled bases depend on the rand() stream after srand(42), so bit parity needs
// the identical stream:
//
//   glibc rand() (TYPE_3) is the additive lagged-Fibonacci generator  r[i] = r[i-31] + r[i-3]  (mod 2^32), output
//   r[i+344] >> 1 for the i-th call. With u_m = r[3+m] the recurrence holds for all m >= 0, i.e. the stream is linear
//   with characteristic polynomial x^31 = x^28 + 1 over Z/2^32, and u_E = sum_k c_k u_k with c = x^E mod that polynomial.
//   Every negative sequence consumes exactly L draws (one per base; sampled codes are never 0, so the Sequence
//   constructor draws nothing), hence sequence g starts at draw D_g = sum of the lengths before it. A thread jumps to
//   the start of ITS block of consecutive negatives with ~log2(D) polynomial products and then simply runs the
//   generator — the device reproduces the host stream exactly, in parallel.
//
// The float arithmetic of the model rescaling repeats the reference's operation order in fp32 (the library is
// compiled with -fmad=false), so the cumulative bars and therefore every sampled base are identical.
#pragma once
#include "kernels.cuh"

namespace bamm {

struct NegDims {
    int A;                      // alphabet size
    uint32_t Y1, Y2, Y3;        // A, A^2, A^3
    uint32_t total;             // Y1 + Y2 + Y3: tables of order 0, 1, 2 concatenated
};
constexpr int NEG_MAXTOT = 6 + 36 + 216;
constexpr int LFG_N = 31;       // state words of glibc's TYPE_3 generator
constexpr int LFG_NPOW = 48;    // x^(2^b) for b < 48

// ---- (1) set-wide k-mer counts: n[k][kmer[j] % A^(k+1)] for j >= k   (SeqGenerator.cpp:73-84) ----------------------
// Y2: order-2 index array of the positive set (kmer % A^3, N draws patched in). One warp per sequence.
__global__ void __launch_bounds__(256)
k_neg_count_set(const uint16_t* __restrict__ Y2, const uint64_t* __restrict__ off, const uint32_t* __restrict__ ids /* nullable: template list */,
                uint64_t nseq, NegDims d, unsigned long long* __restrict__ cnt /* [total] */) {
    __shared__ uint32_t hist[NEG_MAXTOT];
    for (uint32_t b = threadIdx.x; b < d.total; b += blockDim.x) hist[b] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t n = warp; n < nseq; n += nwarps) {
        const uint64_t sid = ids ? ids[n] : n;
        const uint64_t base = off[sid], L = off[sid + 1] - base;
        for (uint64_t j = lane; j < L; j += 32) {
            const uint32_t y = Y2[base + j];
            atomicAdd(&hist[y % d.Y1], 1u);
            if (j >= 1) atomicAdd(&hist[d.Y1 + y % d.Y2], 1u);
            if (j >= 2) atomicAdd(&hist[d.Y1 + d.Y2 + y], 1u);
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < d.total; b += blockDim.x) if (hist[b]) atomicAdd(&cnt[b], (unsigned long long)hist[b]);
}

// ---- (2) per-template rescaled model -> cumulative bars of order 1 and 2   (SeqGenerator.cpp:114-185) --------------
// vset: set-wide probabilities v_[k][y], tables concatenated. pc: pseudo-count weight A_[k] (20 for every k).
// rb: [nseq][Y2 + Y3]. One warp per template: the lanes count, lane 0 does the (strictly ordered) float arithmetic.
__global__ void __launch_bounds__(128)
k_neg_models(const uint16_t* __restrict__ Y2, const uint64_t* __restrict__ off, const uint32_t* __restrict__ ids, uint64_t nseq, NegDims d,
             const float* __restrict__ vset, float pc, float* __restrict__ rb) {
    __shared__ uint32_t hist_all[4][NEG_MAXTOT];
    __shared__ float vs1_all[4][36];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* hist = hist_all[wib];
    float* vs1 = vs1_all[wib];
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const float* v0 = vset; const float* v1 = vset + d.Y1;
    for (uint64_t n = warp; n < nseq; n += nwarps) {
        for (uint32_t b = lane; b < d.total; b += 32) hist[b] = 0u;
        __syncwarp();
        const uint64_t sid = ids ? ids[n] : n;
        const uint64_t base = off[sid], L = off[sid + 1] - base;
        for (uint64_t j = lane; j < L; j += 32) {
            const uint32_t y = Y2[base + j];
            atomicAdd(&hist[y % d.Y1], 1u);
            if (j >= 1) atomicAdd(&hist[d.Y1 + y % d.Y2], 1u);
            if (j >= 2) atomicAdd(&hist[d.Y1 + d.Y2 + y], 1u);
        }
        __syncwarp();
        if (lane == 0) {
            const uint32_t* n0 = hist; const uint32_t* n1 = hist + d.Y1; const uint32_t* n2 = hist + d.Y1 + d.Y2;
            float* rb1 = rb + n * (uint64_t)(d.Y2 + d.Y3);
            float* rb2 = rb1 + d.Y2;
            const float Lf = (float)L;
            // k = 1 (SeqGenerator.cpp:144-166)
            for (uint32_t y = 0; y < d.Y2; y++) {
                const uint32_t y2 = y % d.Y1;
                vs1[y] = v1[y] * ((float)n1[y] + pc * v0[y2]) / v0[y2] / (Lf + pc);
            }
            float nf[6];
            for (uint32_t a = 0; a < d.Y1; a++) nf[a] = 0.0f;
            for (uint32_t y = 0; y < d.Y2; y++) {
                const uint32_t yk = y / d.Y1;
                vs1[y] = ((float)n1[y] + pc * vs1[y]) / ((float)n0[yk] + pc);
                nf[yk] += vs1[y];
            }
            float sum = 0.0f;
            for (uint32_t y = 0; y < d.Y2; y++) {
                vs1[y] /= nf[y / d.Y1];
                if (y % d.Y1 == 0) sum = 0.0f;
                sum += vs1[y];
                rb1[y] = sum;
            }
            // k = 2 (SeqGenerator.cpp:168-180)
            for (uint32_t y = 0; y < d.Y3; y++) {
                const uint32_t y2 = y % d.Y2, yk = y / d.Y1;
                const float v = ((float)n2[y] + pc * vs1[y2]) / ((float)n1[yk] + pc);
                if (y % d.Y1 == 0) sum = 0.0f;
                sum += v;
                rb2[y] = sum;
            }
        }
        __syncwarp();
    }
}

// ---- glibc rand() as a linear recurrence -----------------------------------------------------------------------------
// c <- c * p  mod (x^31 - x^28 - 1), coefficients mod 2^32
__device__ __host__ inline void lfg_poly_mul(uint32_t* c, const uint32_t* p) {
    uint32_t t[2 * LFG_N - 1];
    for (int i = 0; i < 2 * LFG_N - 1; i++) t[i] = 0u;
    for (int i = 0; i < LFG_N; i++) {
        const uint32_t ci = c[i];
        if (ci == 0u) continue;
        for (int j = 0; j < LFG_N; j++) t[i + j] += ci * p[j];
    }
    for (int dgr = 2 * LFG_N - 2; dgr >= LFG_N; dgr--) { t[dgr - 3] += t[dgr]; t[dgr - LFG_N] += t[dgr]; }
    for (int i = 0; i < LFG_N; i++) c[i] = t[i];
}
// state[t] = u_{E+t}, t < 31, from the base window u0[k] = u_k and the table pw[b] = x^(2^b)
__device__ __host__ inline void lfg_jump(unsigned long long E, const uint32_t* u0, const uint32_t* pw, uint32_t* state) {
    uint32_t c[LFG_N];
    for (int i = 0; i < LFG_N; i++) c[i] = 0u;
    c[0] = 1u;
    for (int b = 0; b < LFG_NPOW && (E >> b) != 0ull; b++)
        if ((E >> b) & 1ull) lfg_poly_mul(c, pw + b * LFG_N);
    for (int t = 0; t < LFG_N; t++) {
        uint32_t acc = 0u;
        for (int k = 0; k < LFG_N; k++) acc += c[k] * u0[k];
        state[t] = acc;
        const uint32_t top = c[LFG_N - 1];                       // c <- x * c
        for (int k = LFG_N - 1; k > 0; k--) c[k] = c[k - 1];
        c[0] = top; c[28] += top;
    }
}

// draws [first, first+count) of the stream (test hook): one thread per block of `per_thread` draws
__global__ void k_rand_stream(const uint32_t* __restrict__ u0, const uint32_t* __restrict__ pw, unsigned long long first,
                              unsigned long long count, unsigned long long per_thread, int* __restrict__ out) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long a = t * per_thread;
    if (a >= count) return;
    const unsigned long long b = a + per_thread < count ? a + per_thread : count;
    uint32_t st[LFG_N];
    lfg_jump(first + a + 310ull, u0, pw, st);                     // draw D is u_{D+341}: the 31 words before it
    int idx = 0;
    for (unsigned long long i = a; i < b; i++) {
        const int i28 = idx + 28 >= LFG_N ? idx + 28 - LFG_N : idx + 28;
        const uint32_t v = st[idx] + st[i28];
        st[idx] = v;
        idx = idx + 1 == LFG_N ? 0 : idx + 1;
        out[i] = (int)(v >> 1);
    }
}

// ---- (3) sampling   (SeqGenerator.cpp:285-348 with sOrder = 2) ------------------------------------------------------
// Thread t generates the consecutive negatives [t*per_thread, (t+1)*per_thread): negative g = template g / fold, length =
// the template's stored length, written at byte offset = draw offset = fold * off[template] + (g % fold) * L. The
// generator state lives in shared memory ([word][thread], conflict-free).
// flags[0] is set when a first base stayed 0 (the reference leaves code 0 when the draw exceeds the last bar — about one
// draw in 10^7; the Sequence constructor would then consume extra rand() calls): the caller falls back to the host path.
constexpr int NEG_THREADS = 128;
__global__ void __launch_bounds__(NEG_THREADS)
k_neg_sample(const uint64_t* __restrict__ off /* [nseq+1] prefix sums of the TEMPLATE LIST's lengths */, uint64_t nseq, uint64_t fold, NegDims d, const float* __restrict__ rb0,
             const float* __restrict__ rb, const uint32_t* __restrict__ u0, const uint32_t* __restrict__ pw,
             uint64_t per_thread, uint64_t draw0 /* draws consumed by the templates of earlier shards */, uint8_t* __restrict__ codes,
             uint32_t* __restrict__ flags) {
    __shared__ uint32_t st_sh[LFG_N][NEG_THREADS];
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nneg = nseq * fold;
    uint64_t g = t * per_thread;
    if (g >= nneg) return;
    const uint64_t gend = g + per_thread < nneg ? g + per_thread : nneg;
    uint64_t tmpl = g / fold, m = g % fold;
    uint64_t tbase = off[tmpl], L = off[tmpl + 1] - tbase;
    uint64_t o = fold * tbase + m * L;                           // byte offset of negative g = index of its first draw
    {
        uint32_t st[LFG_N];
        lfg_jump(draw0 + o + 310ull, u0, pw, st);
#pragma unroll 1
        for (int k = 0; k < LFG_N; k++) st_sh[k][threadIdx.x] = st[k];
    }
    int idx = 0;
    const uint32_t A = d.Y1;
    bool bad = false;
    // the thread's negatives are contiguous in the output: whole 32-bit words are stored at once, single bytes only at the
    // unaligned head and tail of its range
    const uint64_t head_end = (o + 3) & ~3ull;                  // first 4-byte aligned output offset of this thread
    uint32_t acc = 0;
    for (; g < gend; g++) {
        const float* __restrict__ rb1 = rb + tmpl * (uint64_t)(d.Y2 + d.Y3);
        const float* __restrict__ rb2 = rb1 + d.Y2;
        uint32_t c1 = 0, c2 = 0;                                 // codes-1 of the two previous bases
        for (uint64_t i = 0; i < L; i++) {
            const int i28 = idx + 28 >= LFG_N ? idx + 28 - LFG_N : idx + 28;
            const uint32_t v = st_sh[idx][threadIdx.x] + st_sh[i28][threadIdx.x];
            st_sh[idx][threadIdx.x] = v;
            idx = idx + 1 == LFG_N ? 0 : idx + 1;
            const float rnd = (float)(int)(v >> 1) / 2147483648.0f;      // (float)rand() / (float)RAND_MAX
            uint32_t code;
            if (i == 0) {
                code = 0;
                for (uint32_t y = 0; y < A; y++) if (rnd <= rb0[y]) { code = y + 1; break; }
                if (code == 0) { bad = true; code = 1; }
            } else {
                const float* __restrict__ bar = (i == 1) ? rb1 + c1 * A : rb2 + (c2 * A + c1) * A;
                code = A;
                for (uint32_t a = 0; a + 1 < A; a++) if (rnd <= bar[a]) { code = a + 1; break; }
            }
            const uint64_t a = o + i;
            if (a < head_end) codes[a] = (uint8_t)code;
            else {
                acc |= code << (8 * (uint32_t)(a & 3));
                if ((a & 3) == 3) { *reinterpret_cast<uint32_t*>(codes + (a - 3)) = acc; acc = 0; }
            }
            c2 = c1; c1 = code - 1;
        }
        o += L;
        if (++m == fold) {
            m = 0; tmpl++;
            if (tmpl < nseq) { tbase = off[tmpl]; L = off[tmpl + 1] - tbase; }
        }
    }
    if (o > head_end) for (uint64_t a = o & ~3ull; a < o; a++) codes[a] = (uint8_t)(acc >> (8 * (uint32_t)(a & 3)));   // tail bytes of a partial word
    if (bad) flags[0] = 1u;
}

}  // namespace bamm
