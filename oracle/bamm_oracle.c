/*
 * bamm_oracle.c — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's EM-refinement / sequence-scoring hot path
 * (soedinglab/BaMMmotif2, C++11). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product (libbamm_b200.so)
 * never links, loads or calls it.
 *
 * Parity status: PINNED. The reference ships no golden vectors, so this
 * restatement is pinned against the reference ITSELF: oracle/Makefile compiles the untouched
 * reference sources into oracle/_ref/ref_dump, oracle/make_golden.py runs it (1 thread) and
 * commits the per-iteration intermediates under tests/golden/; tests/test_oracle_golden.py
 * checks every function below against those vectors (bit-exact for integer work and for the
 * sequential float arithmetic, which this file reproduces operation by operation).
 *
 * Every function cites the reference file:line it follows. Array layouts are flat:
 *   kmer     : uint64 per stored position, all sequences concatenated (reference: size_t* kmer_)
 *   offsets  : nseq+1 prefix sums of stored lengths L_n
 *   v_all    : for k=0..K, for y<A^(k+1), for j<W  (reference float*** v_[k][y][j])
 *   vbg_all  : for k=0..Kbg_model, for y<A^(k+1)   (reference float** v_[k][y])
 *   alpha    : [K+1][W]                             (reference float** A_[k][j])
 *   s        : [A^(K+1)][W]
 *   r        : one float per stored position, per sequence indexed i = L-W-p (reversed, EM.cpp:156-159)
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* reference: src/refinement/utils.h:167-179 (ipow) */
static uint64_t orc_ipow(uint64_t base, uint64_t e) {
    uint64_t r = 1;
    while (e) { if (e & 1) r *= base; e >>= 1; base *= base; }
    return r;
}

static size_t v_offset(int A, int k, int W) { /* start of order k inside v_all */
    size_t off = 0;
    for (int kk = 0; kk < k; kk++) off += (size_t)orc_ipow(A, kk + 1) * (size_t)W;
    return off;
}
static size_t bg_offset(int A, int k) {
    size_t off = 0;
    for (int kk = 0; kk < k; kk++) off += (size_t)orc_ipow(A, kk + 1);
    return off;
}

ORC_API uint64_t orc_model_size(int A, int K, int W) { return v_offset(A, K + 1, W); }
ORC_API uint64_t orc_bg_size(int A, int K) { return bg_offset(A, K + 1); }

/* ---------------------------------------------------------------- Alphabet ------------- */
/* reference: src/init/Alphabet.cpp:10-55. type: 0 STANDARD, 1 METHYLC, 2 HYDROXYMETHYLC, 3 EXTENDED.
 * Fills base_to_code[128] and code_to_comp[128]; returns alphabet size (0 on bad type). */
ORC_API int orc_alphabet_tables(int type, uint8_t* base_to_code, uint8_t* code_to_comp) {
    const char* alpha; const char* comp; int size;
    switch (type) {
        case 0: size = 4; alpha = "ACGT";   comp = "TGCA";   break;
        case 1: size = 5; alpha = "ACGTM";  comp = "TGCAG";  break;
        case 2: size = 5; alpha = "ACGTH";  comp = "TGCAG";  break;
        case 3: size = 6; alpha = "ACGTMH"; comp = "TGCAGG"; break;
        default: return 0;
    }
    memset(base_to_code, 0, 128); memset(code_to_comp, 0, 128);
    for (int i = 0; i < size; i++) {
        base_to_code[(int)alpha[i]] = (uint8_t)(i + 1);
        base_to_code[(int)(alpha[i] | 0x20)] = (uint8_t)(i + 1);      /* tolower, Alphabet.cpp:38 */
    }
    for (int i = 0; i < size; i++) code_to_comp[i + 1] = base_to_code[(int)comp[i]];
    code_to_comp[0] = 'N';                                           /* Alphabet.cpp:51 (78, not 0) */
    for (int i = size; i < 127; i++) code_to_comp[i + 1] = 'N';
    return size;
}

ORC_API void orc_srand(unsigned seed) { srand(seed); }
ORC_API int orc_rand(void) { return rand(); }

/* reference: src/init/Sequence.cpp:4-43 + appendRevComp :91-99.
 * enc: L0 codes of the input strand. Writes stored codes (L = 2*L0+1 or L0) and the 11-mer hash
 * kmer[i]; a code-0 base draws rand() % A independently for every (i,k) pair, in the reference's
 * order (i ascending, k descending). Returns L. */
ORC_API uint64_t orc_sequence_build(const uint8_t* enc, uint64_t L0, int A, const uint8_t* code_to_comp,
                                    int single_strand, uint8_t* codes, uint64_t* kmer) {
    uint64_t L;
    if (!single_strand) {
        L = 2 * L0 + 1;
        memset(codes, 0, L);
        for (uint64_t i = 0; i < L0; i++) { codes[i] = enc[i]; codes[2 * L0 - i] = code_to_comp[enc[i]]; }
    } else {
        L = L0;
        memcpy(codes, enc, L);
    }
    uint64_t Y[12];
    for (int i = 0; i < 12; i++) Y[i] = orc_ipow((uint64_t)A, (uint64_t)i);
    for (uint64_t i = 0; i < L; i++) {
        kmer[i] = 0;
        for (uint64_t k = (i < 10 ? i + 1 : 11); k > 0; k--) {
            uint8_t c = codes[i - k + 1];
            kmer[i] += ((c == 0) ? ((uint64_t)rand() % Y[1]) : (uint64_t)(c - 1)) * Y[k - 1];
        }
    }
    return L;
}

/* ---------------------------------------------------------------- BackgroundModel ------ */
/* reference: src/init/BackgroundModel.cpp:26-42 (counts) and :441-472 (calculateV). */
ORC_API void orc_bg_model(const uint64_t* kmer, uint64_t npos, int A, int K, const float* alpha,
                          int interpolate, uint64_t* n_all, float* v_all) {
    uint64_t Y[16];
    for (int i = 0; i < 16; i++) Y[i] = orc_ipow((uint64_t)A, (uint64_t)i);
    memset(n_all, 0, sizeof(uint64_t) * bg_offset(A, K + 1));
    for (uint64_t i = 0; i < npos; i++)
        for (int k = 0; k <= K; k++) n_all[bg_offset(A, k) + kmer[i] % Y[k + 1]]++;
    uint64_t base = 0;
    for (uint64_t y = 0; y < Y[1]; y++) base += n_all[y];
    for (uint64_t y = 0; y < Y[1]; y++)
        v_all[y] = ((float)n_all[y] + alpha[0] * 0.25f) / ((float)base + alpha[0]);
    for (int k = 1; k <= K; k++) {
        const uint64_t* nk = n_all + bg_offset(A, k); const uint64_t* nk1 = n_all + bg_offset(A, k - 1);
        float* vk = v_all + bg_offset(A, k); const float* vk1 = v_all + bg_offset(A, k - 1);
        for (uint64_t y = 0; y < Y[k + 1]; y++) {
            uint64_t y2 = y % Y[k], yk = y / Y[1];
            if (interpolate) vk[y] = ((float)nk[y] + alpha[k] * vk1[y2]) / ((float)nk1[yk] + alpha[k]);
            else             vk[y] = ((float)nk[y] + alpha[k] * 0.25f)   / ((float)nk1[yk] + alpha[k]);
        }
    }
}

/* ---------------------------------------------------------------- Motif ---------------- */
/* reference: src/init/Motif.cpp:134-189 (initFromBindingSites, no flanks) + calculateV :403-428.
 * sites: C rows of W codes (1..A). alpha: [K+1][W]. */
ORC_API void orc_motif_from_sites(const uint8_t* sites, uint64_t C, int W, int A, int K,
                                  const float* alpha, const float* vbg_all, float* v_all) {
    uint64_t Y[16];
    for (int i = 0; i < 16; i++) Y[i] = orc_ipow((uint64_t)A, (uint64_t)i);
    size_t total = v_offset(A, K + 1, W);
    int* n = (int*)calloc(total, sizeof(int));
    for (uint64_t c = 0; c < C; c++) {
        const uint8_t* bs = sites + c * (uint64_t)W;
        for (int k = 0; k <= K; k++)
            for (int j = k; j < W; j++) {
                uint64_t y = 0;
                for (int a = 0; a <= k; a++) y += Y[a] * (uint64_t)(bs[j - a] - 1);
                n[v_offset(A, k, W) + y * W + j]++;
            }
    }
    for (uint64_t y = 0; y < Y[1]; y++)
        for (int j = 0; j < W; j++)
            v_all[y * W + j] = ((float)n[y * W + j] + alpha[j] * vbg_all[y]) / ((float)C + alpha[j]);
    for (int k = 1; k <= K; k++) {
        float* vk = v_all + v_offset(A, k, W); const float* vk1 = v_all + v_offset(A, k - 1, W);
        const int* nk = n + v_offset(A, k, W); const int* nk1 = n + v_offset(A, k - 1, W);
        for (uint64_t y = 0; y < Y[k + 1]; y++) {
            uint64_t y2 = y % Y[k], yk = y / Y[1];
            for (int j = 0; j < k; j++) vk[y * W + j] = vk1[y2 * W + j];
            for (int j = k; j < W; j++)
                vk[y * W + j] = ((float)nk[y * W + j] + alpha[k * W + j] * vk1[y2 * W + j])
                              / ((float)nk1[yk * W + j - 1] + alpha[k * W + j]);
        }
    }
    free(n);
}

/* reference: src/init/Motif.h:95-136 (updateV). n_all: float counts, all orders. */
ORC_API void orc_update_v(const float* n_all, const float* alpha, const float* vbg_all,
                          int A, int K, int W, float* v_all) {
    uint64_t Y[16];
    for (int i = 0; i < 16; i++) Y[i] = orc_ipow((uint64_t)A, (uint64_t)i);
    float* sumN = (float*)calloc((size_t)W, sizeof(float));
    for (uint64_t y = 0; y < Y[1]; y++) for (int j = 0; j < W; j++) sumN[j] += n_all[y * W + j];
    for (uint64_t y = 0; y < Y[1]; y++)
        for (int j = 0; j < W; j++)
            v_all[y * W + j] = (n_all[y * W + j] + alpha[j] * vbg_all[y]) / (sumN[j] + alpha[j]);
    for (int k = 1; k <= K; k++) {
        float* vk = v_all + v_offset(A, k, W); const float* vk1 = v_all + v_offset(A, k - 1, W);
        const float* nk = n_all + v_offset(A, k, W); const float* nk1 = n_all + v_offset(A, k - 1, W);
        for (uint64_t y = 0; y < Y[k + 1]; y++) {
            uint64_t y2 = y % Y[k], yk = y / Y[1];
            for (int j = 0; j < k; j++) vk[y * W + j] = vk1[y2 * W + j];
            for (int j = k; j < W; j++)
                vk[y * W + j] = (nk[y * W + j] + alpha[k * W + j] * vk1[y2 * W + j])
                              / (nk1[yk * W + j - 1] + alpha[k * W + j]);
        }
    }
    free(sumN);
}

/* reference: src/init/Motif.cpp:430-469 (calculateP). k_bg = order of the bg model held by the motif. */
ORC_API void orc_calculate_p(const float* v_all, const float* vbg_all, int k_bg, int A, int K, int W, float* p_all) {
    uint64_t Y[16];
    for (int i = 0; i < 16; i++) Y[i] = orc_ipow((uint64_t)A, (uint64_t)i);
    for (int j = 0; j < W; j++) for (uint64_t y = 0; y < Y[1]; y++) p_all[y * W + j] = v_all[y * W + j];
    for (int k = 1; k <= K; k++) {
        float* pk = p_all + v_offset(A, k, W); const float* pk1 = p_all + v_offset(A, k - 1, W);
        const float* vk = v_all + v_offset(A, k, W);
        for (uint64_t y = 0; y < Y[k + 1]; y++) {
            uint64_t yk = y / Y[1];
            for (int j = 0; j < k; j++) {
                float p = 1;
                for (int i = 0; i <= j; i++) {
                    uint64_t yi = y / Y[i];
                    p *= v_all[v_offset(A, k - i, W) + yi * W + (j - i)];
                }
                for (int i = j + 1; i <= k; i++) {
                    if ((k - i) <= k_bg || k <= k_bg) {
                        uint64_t yi = y / Y[i];
                        p *= vbg_all[bg_offset(A, k - i) + yi];
                    } else {
                        uint64_t yi = y / Y[1] % Y[k_bg + 1];
                        p *= vbg_all[bg_offset(A, k_bg) + yi];
                    }
                }
                pk[y * W + j] = p;
            }
            for (int j = k; j < W; j++) pk[y * W + j] = vk[y * W + j] * pk1[yk * W + j - 1];
        }
    }
}

/* reference: src/init/Motif.cpp:485-494 (calculateLinearS) */
ORC_API void orc_linear_s(const float* v_all, const float* vbg_all, int A, int K, int K_bg, int W, float* s) {
    uint64_t YK1 = orc_ipow(A, K + 1), YB = orc_ipow(A, K_bg + 1);
    const float* vK = v_all + v_offset(A, K, W); const float* vb = vbg_all + bg_offset(A, K_bg);
    for (uint64_t y = 0; y < YK1; y++) { uint64_t yb = y % YB; for (int j = 0; j < W; j++) s[y * W + j] = vK[y * W + j] / vb[yb]; }
}
/* reference: src/init/Motif.cpp:471-483 (calculateLogS) */
ORC_API void orc_log_s(const float* v_all, const float* vbg_all, int A, int K, int K_bg, int W, float* s) {
    uint64_t YK1 = orc_ipow(A, K + 1), YB = orc_ipow(A, K_bg + 1);
    const float* vK = v_all + v_offset(A, K, W); const float* vb = vbg_all + bg_offset(A, K_bg);
    for (uint64_t y = 0; y < YK1; y++) {
        uint64_t yb = y % YB;
        for (int j = 0; j < W; j++) { float rnd = 1e-5f; s[y * W + j] = logf(vK[y * W + j] + rnd) - logf(vb[yb]); }
    }
}

/* ---------------------------------------------------------------- EM ------------------- */
/* reference: src/refinement/EM.cpp:139-200 (EStep), scatter form, operation order preserved.
 * r must hold sum(L_n) floats. Returns the log likelihood (float accumulation over n, as with 1 thread).
 * If llh_double != NULL it also receives the same sum accumulated in double (scale check). */
ORC_API float orc_estep(const uint64_t* kmer, const uint64_t* offsets, uint64_t nseq, int A, int K, int W,
                        const float* s, float q, float* r, double* llh_double) {
    uint64_t YK1 = orc_ipow(A, K + 1);
    float llh = 0.0f; double llhd = 0.0;
    for (uint64_t n = 0; n < nseq; n++) {
        uint64_t L = offsets[n + 1] - offsets[n];
        uint64_t LW1 = L - W + 1;
        const uint64_t* km = kmer + offsets[n];
        float* rn = r + offsets[n];
        float norm = 1.0f - q;
        float pos_i = q / (float)LW1;
        for (uint64_t i = 0; i < L; i++) rn[i] = (i < LW1) ? 1.0f : 0.0f;
        for (uint64_t ij = 0; ij < LW1; ij++) {
            uint64_t y = km[ij] % YK1;
            for (int j = 0; j < W; j++) rn[L - W - ij + j] *= s[y * W + j];
        }
        for (uint64_t i = 0; i < LW1; i++) { rn[i] *= pos_i; norm += rn[i]; }
        for (uint64_t i = 0; i < LW1; i++) rn[i] /= norm;
        for (uint64_t i = LW1; i < L; i++) rn[i] = 0.0f;
        llh += logf(norm); llhd += (double)logf(norm);
    }
    if (llh_double) *llh_double = llhd;
    return llh;
}

/* reference: src/refinement/EM.cpp:217-259 (MStep without the final updateV), 1-thread order.
 * n_all receives all orders (top order accumulated, lower orders folded, EM.cpp:247-254).
 * accumulate_double != 0: top-order sums are taken in double and rounded once (scale check). */
ORC_API void orc_mstep(const uint64_t* kmer, const uint64_t* offsets, uint64_t nseq, int A, int K, int W,
                       const float* r, float* n_all, int accumulate_double) {
    uint64_t Y[16];
    for (int i = 0; i < 16; i++) Y[i] = orc_ipow((uint64_t)A, (uint64_t)i);
    size_t total = v_offset(A, K + 1, W);
    memset(n_all, 0, total * sizeof(float));
    float* nK = n_all + v_offset(A, K, W);
    double* nd = accumulate_double ? (double*)calloc(Y[K + 1] * (size_t)W, sizeof(double)) : NULL;
    for (uint64_t n = 0; n < nseq; n++) {
        uint64_t L = offsets[n + 1] - offsets[n];
        const uint64_t* km = kmer + offsets[n];
        const float* rn = r + offsets[n];
        for (uint64_t ij = 0; ij < L - W + 1; ij++) {
            uint64_t y = km[ij] % Y[K + 1];
            if (nd) for (int j = 0; j < W; j++) nd[y * W + j] += (double)rn[L - W - ij + j];
            else    for (int j = 0; j < W; j++) nK[y * W + j] += rn[L - W - ij + j];
        }
    }
    if (nd) { for (size_t i = 0; i < Y[K + 1] * (size_t)W; i++) nK[i] = (float)nd[i]; free(nd); }
    for (int k = K; k > 0; k--) {
        float* nk = n_all + v_offset(A, k, W); float* nk1 = n_all + v_offset(A, k - 1, W);
        for (uint64_t y = 0; y < Y[k + 1]; y++) {
            uint64_t y2 = y % Y[k];
            for (int j = 0; j < W; j++) nk1[y2 * W + j] += nk[y * W + j];
        }
    }
}

/* reference: src/refinement/EM.cpp:505-519 (optimize_q) */
ORC_API float orc_optimize_q(const uint64_t* offsets, uint64_t nseq, int W, const float* r) {
    float N1 = 0.f;
    for (uint64_t n = 0; n < nseq; n++) {
        uint64_t L = offsets[n + 1] - offsets[n];
        const float* rn = r + offsets[n];
        for (uint64_t i = 0; i < L - W + 1; i++) N1 += rn[i];
    }
    return ((float)nseq - N1 + 1.f) / ((float)nseq + 2.f);
}

/* reference: src/refinement/EM.cpp:62-137 (optimize). Runs the loop with the reference's stop rule
 * (epsilon 0.01, llh drop after iteration 10, at most max_iter=1000; EM.h:61-63). Traces (length
 * max_iter) may be NULL. r: workspace of sum(L_n) floats. Returns the iteration count. */
ORC_API int orc_em_optimize(const uint64_t* kmer, const uint64_t* offsets, uint64_t nseq, int A, int K, int W,
                            int K_bg_model, const float* vbg_all, const float* alpha, float* v_all, float* q_io,
                            int optimize_q, float epsilon, int max_iter, float* r,
                            float* llh_trace, float* vdiff_trace, float* q_trace, float* n_all_out) {
    int K_bg = K_bg_model < K ? K_bg_model : K;
    uint64_t YK1 = orc_ipow(A, K + 1);
    size_t total = v_offset(A, K + 1, W);
    float* s = (float*)malloc(YK1 * (size_t)W * sizeof(float));
    float* n_all = (float*)malloc(total * sizeof(float));
    float* v_before = (float*)malloc(YK1 * (size_t)W * sizeof(float));
    float* vK = v_all + v_offset(A, K, W);
    float q = *q_io, llh = 0.0f, llh_prev;
    int iterate = 1, it = 0;
    while (iterate && it < max_iter) {
        it++;
        llh_prev = llh;
        memcpy(v_before, vK, YK1 * (size_t)W * sizeof(float));
        orc_linear_s(v_all, vbg_all, A, K, K_bg, W, s);
        llh = orc_estep(kmer, offsets, nseq, A, K, W, s, q, r, NULL);
        orc_mstep(kmer, offsets, nseq, A, K, W, r, n_all, 0);
        orc_update_v(n_all, alpha, vbg_all, A, K, W, v_all);
        if (optimize_q && it <= 5) q = orc_optimize_q(offsets, nseq, W, r);
        float v_diff = 0.0f;
        for (size_t i = 0; i < YK1 * (size_t)W; i++) v_diff += fabsf(vK[i] - v_before[i]);
        float llh_diff = llh - llh_prev;
        if (llh_trace) llh_trace[it - 1] = llh;
        if (vdiff_trace) vdiff_trace[it - 1] = v_diff;
        if (q_trace) q_trace[it - 1] = q;
        if (v_diff < epsilon) iterate = 0;
        if (llh_diff < 0 && it > 10) iterate = 0;
    }
    if (n_all_out) memcpy(n_all_out, n_all, total * sizeof(float));
    *q_io = q;
    free(s); free(n_all); free(v_before);
    return it;
}

/* reference: src/refinement/EM.cpp:261-503 (mask, the "advanced EM" of --advanceEM) without optimizeQ:
 * (1) one E-step with the order-0 model over all windows (full products; the window at p = 0 keeps the bare prior because
 *     the loop bound j < min(W, ij) skips it, :300-303), ascending-i normaliser;
 * (2) the cutoff that keeps the fraction f of all windows with the largest r (descending sort, :318-345);
 * (3) EM over the kept windows only, full W-column products and counts (no truncated windows here), the position prior read
 *     at pos_[LW1 - i] (so the window i = 0, when kept, gets prior 0: :417), r[0] divided by the normaliser once more (:422),
 *     the stop rule of optimize().
 * r: sum(L_n) floats, zero-initialised by the caller like the reference's calloc. Returns the iteration count. */
static int cmp_float_desc(const void* a, const void* b) {
    const float x = *(const float*)a, y = *(const float*)b;
    return (x < y) - (x > y);
}
ORC_API int orc_em_mask(const uint64_t* kmer, const uint64_t* offsets, uint64_t nseq, int A, int K, int W, int K_bg_model,
                        const float* vbg_all, const float* alpha, float* v_all, float q, float f, float epsilon, int max_iter,
                        float* r, float* llh_out, float* cutoff_out, uint64_t* nkept_out, float* n_all_out) {
    int K_bg = K_bg_model < K ? K_bg_model : K;
    uint64_t YK1 = orc_ipow(A, K + 1), Y1 = (uint64_t)A;
    size_t total = v_offset(A, K + 1, W);
    float* s = (float*)calloc(YK1 * (size_t)W, sizeof(float));
    float* n_all = (float*)calloc(total, sizeof(float));
    float* v_before = (float*)malloc(YK1 * (size_t)W * sizeof(float));
    float* vK = v_all + v_offset(A, K, W);
    uint64_t npos = offsets[nseq];
    float* pos = (float*)calloc(npos ? npos : 1, sizeof(float));
    /* (1) */
    for (uint64_t y = 0; y < Y1; y++) for (int j = 0; j < W; j++) s[y * W + j] = v_all[y * W + j] / vbg_all[y];
    uint64_t pos_count = 0;
    for (uint64_t n = 0; n < nseq; n++) {
        uint64_t L = offsets[n + 1] - offsets[n], LW1 = L - W + 1;
        const uint64_t* km = kmer + offsets[n];
        float* rn = r + offsets[n]; float* pn = pos + offsets[n];
        float normFactor = 1.0f - q;
        float pos_i = q / (float)LW1;
        for (uint64_t i = 0; i < LW1; i++) { rn[i] = 1.0f; pn[i] = pos_i; }
        for (uint64_t ij = 0; ij < L; ij++) {
            uint64_t y = km[ij] % Y1;
            uint64_t padding = ((int)(ij - L + W) > 0) * (ij - L + W);
            for (uint64_t j = padding; j < ((uint64_t)W < ij ? (uint64_t)W : ij); j++) rn[L - W - ij + j] *= s[y * W + j];
        }
        for (uint64_t i = 0; i < LW1; i++) { rn[i] *= pn[L - W - i]; normFactor += rn[i]; }
        for (uint64_t i = 0; i < LW1; i++) rn[i] /= normFactor;
        pos_count += LW1;
    }
    /* (2) */
    float* r_all = (float*)malloc((pos_count ? pos_count : 1) * sizeof(float));
    { uint64_t c = 0; for (uint64_t n = 0; n < nseq; n++) { uint64_t L = offsets[n + 1] - offsets[n]; for (uint64_t i = 0; i < L - W + 1; i++) r_all[c++] = r[offsets[n] + i]; } }
    qsort(r_all, pos_count, sizeof(float), cmp_float_desc);
    float r_cutoff = r_all[(size_t)((float)pos_count * f)];
    free(r_all);
    uint64_t* ri_off = (uint64_t*)calloc(nseq + 1, sizeof(uint64_t));
    for (uint64_t n = 0; n < nseq; n++) {
        uint64_t L = offsets[n + 1] - offsets[n], c = 0;
        for (uint64_t i = 0; i < L - W + 1; i++) if (r[offsets[n] + i] >= r_cutoff) c++;
        ri_off[n + 1] = ri_off[n] + c;
    }
    uint64_t* ri = (uint64_t*)malloc((ri_off[nseq] ? ri_off[nseq] : 1) * sizeof(uint64_t));
    for (uint64_t n = 0; n < nseq; n++) {
        uint64_t L = offsets[n + 1] - offsets[n], c = ri_off[n];
        for (uint64_t i = 0; i < L - W + 1; i++) if (r[offsets[n] + i] >= r_cutoff) ri[c++] = i;
    }
    if (cutoff_out) *cutoff_out = r_cutoff;
    if (nkept_out) *nkept_out = ri_off[nseq];
    /* (3) */
    float llh = 0.0f, llh_prev;
    int iterate = 1, it = 0;
    while (iterate && it < max_iter) {
        it++;
        llh_prev = llh;
        memcpy(v_before, vK, YK1 * (size_t)W * sizeof(float));
        float llikelihood = 0.0f;
        orc_linear_s(v_all, vbg_all, A, K, K_bg, W, s);
        for (uint64_t n = 0; n < nseq; n++) {
            uint64_t L = offsets[n + 1] - offsets[n], LW1 = L - W + 1;
            const uint64_t* km = kmer + offsets[n];
            float* rn = r + offsets[n]; float* pn = pos + offsets[n];
            float normFactor = 1.0f - q;
            float pos_i = q / (float)LW1;
            for (uint64_t x = ri_off[n]; x < ri_off[n + 1]; x++) { rn[ri[x]] = 1.0f; pn[ri[x]] = pos_i; }
            for (uint64_t x = ri_off[n]; x < ri_off[n + 1]; x++) {
                uint64_t i = ri[x];
                for (int j = 0; j < W; j++) { uint64_t y = km[L - W - i + j] % YK1; rn[i] *= s[y * W + j]; }
                rn[i] *= pn[LW1 - i];
                normFactor += rn[i];
            }
            rn[0] /= normFactor;
            for (uint64_t x = ri_off[n]; x < ri_off[n + 1]; x++) rn[ri[x]] /= normFactor;
            for (uint64_t i = LW1; i < L; i++) rn[i] = 0.0f;
            llikelihood += logf(normFactor);
        }
        llh = llikelihood;
        memset(n_all, 0, total * sizeof(float));
        float* nK = n_all + v_offset(A, K, W);
        for (uint64_t n = 0; n < nseq; n++) {
            uint64_t L = offsets[n + 1] - offsets[n];
            const uint64_t* km = kmer + offsets[n];
            const float* rn = r + offsets[n];
            for (uint64_t x = ri_off[n]; x < ri_off[n + 1]; x++) {
                uint64_t i = ri[x];
                for (int j = 0; j < W; j++) { uint64_t y = km[L - W - i + j] % YK1; nK[y * W + j] += rn[i]; }
            }
        }
        uint64_t Y[16]; Y[0] = 1; for (int k = 1; k < 16; k++) Y[k] = Y[k - 1] * (uint64_t)A;
        for (int k = K; k > 0; k--) {
            float* nk = n_all + v_offset(A, k, W); float* nk1 = n_all + v_offset(A, k - 1, W);
            for (uint64_t y = 0; y < Y[k + 1]; y++) {
                uint64_t y2 = y % Y[k];
                for (int j = 0; j < W; j++) nk1[y2 * W + j] += nk[y * W + j];
            }
        }
        orc_update_v(n_all, alpha, vbg_all, A, K, W, v_all);
        float v_diff = 0.0f;
        for (size_t i = 0; i < YK1 * (size_t)W; i++) v_diff += fabsf(vK[i] - v_before[i]);
        float llh_diff = llh - llh_prev;
        if (v_diff < epsilon) iterate = 0;
        if (llh_diff < 0 && it > 10) iterate = 0;
    }
    if (llh_out) *llh_out = llh;
    if (n_all_out) memcpy(n_all_out, n_all, total * sizeof(float));
    free(s); free(n_all); free(v_before); free(pos); free(ri_off); free(ri);
    return it;
}

/* ---------------------------------------------------------------- ScoreSeqSet ---------- */
/* reference: src/seq_scoring/ScoreSeqSet.cpp:25-67 (calcLogOdds) given the log table s (calculateLogS).
 * mops: sum(L_n - W + 1) floats (may be NULL), zoops/z: nseq entries. */
ORC_API void orc_logodds(const uint64_t* kmer, const uint64_t* offsets, uint64_t nseq, int A, int K, int W,
                         const float* s, float* mops, float* zoops, uint64_t* z) {
    uint64_t YK1 = orc_ipow(A, K + 1);
    uint64_t mo = 0;
    for (uint64_t n = 0; n < nseq; n++) {
        uint64_t L = offsets[n + 1] - offsets[n];
        uint64_t LW1 = L - W + 1;
        const uint64_t* km = kmer + offsets[n];
        float maxScore = -FLT_MAX; uint64_t zi = 0;
        for (uint64_t i = 0; i < LW1; i++) {
            float lo = 0.0f;
            for (int j = 0; j < W; j++) lo += s[(km[i + j] % YK1) * W + j];
            if (mops) mops[mo + i] = lo;
            if (lo > maxScore) { maxScore = lo; zi = i; }
        }
        mo += LW1;
        zoops[n] = maxScore; z[n] = zi;
    }
}

/* ---------------------------------------------------------------- timed multi-thread EM (cpu_baseline "port") --- */
/* The same E-step / M-step as above with the reference's OpenMP structure (EM.cpp:148-149, :230-243:
 * parallel for over sequences, llh reduction, CAS float atomics in the M-step). Used only by bench.py
 * as the "port" CPU baseline when oracle/_ref is unavailable. */
static inline void atomic_float_add(float* src, float v) {   /* reference: EM.cpp:203-215 */
    union { unsigned int i; float f; } nv, pv;
    do { pv.f = *src; nv.f = pv.f + v; }
    while (!__atomic_compare_exchange_n((volatile unsigned int*)src, &pv.i, nv.i, 0, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
}
ORC_API float orc_em_iteration_omp(const uint64_t* kmer, const uint64_t* offsets, uint64_t nseq, int A, int K, int W,
                                   const float* s, float q, float* r, float* n_all, int threads) {
    uint64_t Y[16];
    for (int i = 0; i < 16; i++) Y[i] = orc_ipow((uint64_t)A, (uint64_t)i);
    float llh = 0.0f;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    #pragma omp parallel for reduction(+:llh)
    for (uint64_t n = 0; n < nseq; n++) {
        uint64_t L = offsets[n + 1] - offsets[n], LW1 = L - W + 1;
        const uint64_t* km = kmer + offsets[n]; float* rn = r + offsets[n];
        float norm = 1.0f - q, pos_i = q / (float)LW1;
        for (uint64_t i = 0; i < L; i++) rn[i] = (i < LW1) ? 1.0f : 0.0f;
        for (uint64_t ij = 0; ij < LW1; ij++) { uint64_t y = km[ij] % Y[K + 1]; for (int j = 0; j < W; j++) rn[L - W - ij + j] *= s[y * W + j]; }
        for (uint64_t i = 0; i < LW1; i++) { rn[i] *= pos_i; norm += rn[i]; }
        for (uint64_t i = 0; i < LW1; i++) rn[i] /= norm;
        for (uint64_t i = LW1; i < L; i++) rn[i] = 0.0f;
        llh += logf(norm);
    }
    size_t total = v_offset(A, K + 1, W);
    memset(n_all, 0, total * sizeof(float));
    float* nK = n_all + v_offset(A, K, W);
    #pragma omp parallel for
    for (uint64_t n = 0; n < nseq; n++) {
        uint64_t L = offsets[n + 1] - offsets[n];
        const uint64_t* km = kmer + offsets[n]; const float* rn = r + offsets[n];
        for (uint64_t ij = 0; ij < L - W + 1; ij++) { uint64_t y = km[ij] % Y[K + 1]; for (int j = 0; j < W; j++) atomic_float_add(&nK[y * W + j], rn[L - W - ij + j]); }
    }
    for (int k = K; k > 0; k--) {
        float* nk = n_all + v_offset(A, k, W); float* nk1 = n_all + v_offset(A, k - 1, W);
        for (uint64_t y = 0; y < Y[k + 1]; y++) { uint64_t y2 = y % Y[k]; for (int j = 0; j < W; j++) nk1[y2 * W + j] += nk[y * W + j]; }
    }
    return llh;
}
