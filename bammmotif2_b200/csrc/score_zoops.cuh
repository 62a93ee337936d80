// score_zoops.cuh — ZOOPS-only scoring kernel with column-group pruning. Included by launch_score.cu only.
#pragma once
#include "common.cuh"

namespace bamm {

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
