// capi_score.inl — part of capi.cu (one translation unit: included there, in this order).
// log-odds scoring (rows a-9 / a-10) and score statistics (row f-1)
// ------------------------------------------------------------------------------------------- scoring
extern "C" int bamm_score_logodds(bamm_seqset* s, const uint64_t* subset, uint64_t nsub, int W, int K, int K_bg_model,
                                  const float* v_all, const float* vbg_all, float* zoops, uint64_t* z, float* mops) {
    REQUIRE(s && v_all && vbg_all && zoops && z, "NULL argument");
    REQUIRE(W >= 1 && W <= 32, "motif width W=%d not in [1,32]", W);
    Trace tr("score_logodds");
    if (!subset) nsub = s->nseq;
    REQUIRE(K >= 0 && K <= 10, "order K=%d not in [0,10]", K);
    const int K_bg = K_bg_model < K ? K_bg_model : K;
    ModelDims d; fill_dims(d, s->A, K, W, K_bg);
    const uint64_t ia_Yn = ipow_u64((uint64_t)s->A, K + 1);
    REQUIRE(ia_Yn * (uint64_t)W < (1ull << 31), "table too large");
    const uint32_t Yn = (uint32_t)ia_Yn, nbin = Yn * (uint32_t)W;
    // Motif::calculateLogS (Motif.cpp:471-483) on the host: same libm logf as the reference; [j][y] layout
    std::vector<float> slog(nbin);
    {
        const float* vK = v_all + d.voff[K]; const float* vb = vbg_all + d.bgoff[K_bg];
        const uint32_t YB = d.Y[K_bg + 1];
        for (uint32_t y = 0; y < Yn; y++) {
            const float lb = logf(vb[y % YB]);
            for (int j = 0; j < W; j++) slog[(uint64_t)j * Yn + y] = logf(vK[(uint64_t)y * W + j] + 1e-5f) - lb;
        }
    }
    std::vector<uint32_t> gen_ids, gen_out, pk_ids, pk_out;
    pk_ids.reserve(nsub); pk_out.reserve(nsub);
    std::vector<uint64_t> moff(mops ? nsub + 1 : 1, 0);      // window offsets of the subset: only the MOPS output needs them
    int max_optin = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device);
    const size_t tb = (size_t)nbin * 4;
    const bool smem = tb <= (size_t)max_optin;
    const bool packed_ok = s->A == 4 && s->nregular > 0 && W + K <= 32 && ia_Yn <= 65536 && smem && !getenv("BAMM_NO_PACKED");
    // whole set regular and long enough (e.g. a sampled negative set): no per-sequence look-ups — the caller's subset IS the
    // list; it is uploaded as it is and narrowed / range-checked on the device (k_ids_from_u64), no host pass
    const bool dev_ids = !mops && packed_ok && s->nregular == s->nseq && s->minL >= (uint64_t)W && nsub > 0;
    if (!dev_ids)
    for (uint64_t i = 0; i < nsub; i++) {
        const uint64_t n = subset ? subset[i] : i;
        REQUIRE(n < s->nseq, "subset index out of range");
        const uint64_t L = s->h_off[n + 1] - s->h_off[n];
        REQUIRE(L >= (uint64_t)W, "sequence %llu is shorter than the motif", (unsigned long long)n);
        if (mops) moff[i + 1] = moff[i] + (L - W + 1);
        if (packed_ok && s->h_kind[n]) { pk_ids.push_back((uint32_t)n); pk_out.push_back((uint32_t)i); }
        else { gen_ids.push_back((uint32_t)n); gen_out.push_back((uint32_t)i); }
    }
    const bool identity_out = gen_ids.empty();               // every sequence on the packed path: list index == output index
    const uint64_t npk = dev_ids ? nsub : pk_ids.size();
    tr.mark("log table + subset split (host)");
    // ZOOPS-only calls on the packed path: prune with column-group tables, re-score exactly near the running maximum
    // (k_score_zoops_packed). eps bounds |cheap - exact|: both are fp32 sums of the same W table entries (|entry| <= S) in
    // different associations, each within (W-1) * 2^-24 * W * S of the real sum; factor 1.5 for slack.
    GroupPlan zplan; bool zfast = false, zoops_fast = false; float two_eps = 0.0f;
    if (!mops && npk && !getenv("BAMM_NO_ZOOPS_FAST")) {
        float S = 0.0f;
        for (uint32_t i = 0; i < nbin; i++) { const float a = fabsf(slog[i]); if (!(a <= 3.0e38f)) { S = -1.0f; break; } if (a > S) S = a; }
        const bool reduced = leading_columns_are_copies(d, K, W, Yn, v_all);
        if (S >= 0.0f && tb + 4096 < (size_t)max_optin &&
            make_group_plan(W, K, K_bg, reduced, (size_t)max_optin - tb, 0, W, zplan, zfast)) {
            zoops_fast = true;
            two_eps = 2.0f * 1.5f * 2.0f * (float)W * (float)W * S * 5.9604645e-8f;
        }
    }
    IndexArray* ia = nullptr;
    uint16_t* d_yp = nullptr;
    if (!gen_ids.empty()) { std::lock_guard<std::mutex> g(s->mu); int rc = seqset_index_locked(s, K, &ia); if (rc) return rc; }
    if (npk)  { std::lock_guard<std::mutex> g(s->mu); int rc = seqset_ypatch_locked(s, K, &d_yp); if (rc) return rc; }
    tr.mark("plan + index");
    CU(cudaSetDevice(s->device));
    cudaStream_t st; CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float *d_s = nullptr, *d_zoops = nullptr, *d_mops = nullptr, *d_ztab = nullptr; unsigned long long* d_z = nullptr;
    uint32_t *d_gids = nullptr, *d_gout = nullptr, *d_pids = nullptr, *d_pout = nullptr; uint64_t* d_moff = nullptr;
    uint64_t* d_sub = nullptr; uint32_t* d_bad = nullptr; uint32_t bad = 0;
    int rc = BAMM_OK;
#define CUX(call) do { cudaError_t e2_ = (call); if (e2_ != cudaSuccess) { rc = fail(e2_ == cudaErrorMemoryAllocation ? BAMM_E_NOMEM : BAMM_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e2_)); goto done; } } while (0)
    {
        CUX(dev_malloc(&d_s, (uint64_t)nbin * 4));
        CUX(dev_malloc(&d_zoops, (nsub ? nsub : 1) * 4));
        CUX(dev_malloc(&d_z, (nsub ? nsub : 1) * 8));
        CUX(dev_malloc(&d_gids, (gen_ids.size() ? gen_ids.size() : 1) * 4));
        CUX(dev_malloc(&d_gout, (gen_ids.size() ? gen_ids.size() : 1) * 4));
        CUX(dev_malloc(&d_pids, (npk ? npk : 1) * 4));
        CUX(dev_malloc(&d_pout, (npk && !identity_out ? npk : 1) * 4));
        if (dev_ids) {
            CUX(dev_malloc(&d_bad, 4));
            CUX(cudaMemsetAsync(d_bad, 0, 4, st));
            if (subset) {
                CUX(dev_malloc(&d_sub, nsub * 8));
                CUX(cudaMemcpyAsync(d_sub, subset, nsub * 8, cudaMemcpyHostToDevice, st));
            }
            k_ids_from_u64<<<(unsigned)((nsub + 255) / 256), 256, 0, st>>>(d_sub, nsub, s->nseq, d_pids, d_bad);
            CUX(cudaGetLastError());
        }
        CUX(dev_malloc(&d_moff, moff.size() * 8));
        if (mops) CUX(dev_malloc(&d_mops, (moff[nsub] ? moff[nsub] : 1) * 4));
        CUX(cudaMemcpyAsync(d_s, slog.data(), (uint64_t)nbin * 4, cudaMemcpyHostToDevice, st));
        CUX(cudaMemcpyAsync(d_gids, gen_ids.data(), gen_ids.size() * 4, cudaMemcpyHostToDevice, st));
        CUX(cudaMemcpyAsync(d_gout, gen_out.data(), gen_out.size() * 4, cudaMemcpyHostToDevice, st));
        if (!dev_ids) CUX(cudaMemcpyAsync(d_pids, pk_ids.data(), pk_ids.size() * 4, cudaMemcpyHostToDevice, st));
        if (!identity_out) CUX(cudaMemcpyAsync(d_pout, pk_out.data(), pk_out.size() * 4, cudaMemcpyHostToDevice, st));
        CUX(cudaMemcpyAsync(d_moff, moff.data(), moff.size() * 8, cudaMemcpyHostToDevice, st));
        int per_sm = smem ? (int)((size_t)(max_optin + 1024) / (tb + 1024)) : 4;
        if (per_sm < 1) per_sm = 1; if (per_sm > 4) per_sm = 4;
        const int grid = s->sm_count * per_sm;
        CUX(cudaEventCreate(&ev0)); CUX(cudaEventCreate(&ev1));
        tr.mark("alloc + H2D");
        CUX(cudaEventRecord(ev0, st));
        if (npk) {
            PackedView pv; pv.words = s->d_words; pv.seqs = s->d_pseq; pv.ypatch = d_yp; pv.seq_ids = d_pids; pv.r_off = nullptr; pv.nlist = (uint32_t)npk;
            Plan pl; pl.W = W; pl.K = K; pl.T = 1; pl.C = W; pl.Yn = Yn; pl.Zn = Yn; pl.q = 0.f;
            if (zoops_fast) {
                CUX(dev_malloc(&d_ztab, zplan.table_bytes));
                const uint32_t total = zplan.table_bytes >> 2, blocks = (total + 255) / 256;
                k_make_group_tables<true><<<blocks < 1184 ? blocks : 1184, 256, 0, st>>>(d_s, zplan, d_ztab);
                CUX(cudaGetLastError());
                if (launch_score_zoops(zplan, zfast, s->sm_count, st, pv, d_ztab, d_s, two_eps, d_zoops, d_z, identity_out ? nullptr : d_pout, tb)) {
                    rc = fail(BAMM_E_CUDA, "ZOOPS scoring launch failed: %s", cudaGetErrorString(cudaGetLastError())); goto done;
                }
            } else {
                CUX(cudaFuncSetAttribute(k_score_packed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb));
                k_score_packed<<<grid, 512, tb, st>>>(pv, pl, d_moff, d_s, d_zoops, d_z, d_mops, identity_out ? nullptr : d_pout);
                CUX(cudaGetLastError());
            }
        }
        if (!gen_ids.empty()) {
            SubsetView sv; sv.seq_off = s->d_off; sv.seq_ids = d_gids; sv.r_off = nullptr; sv.nsub = (uint32_t)gen_ids.size();
            if (ia->bytes == 2) {
                if (smem) { CUX(cudaFuncSetAttribute(k_score<uint16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb));
                            k_score<uint16_t, true><<<grid, 512, tb, st>>>((const uint16_t*)ia->d, sv, d_moff, W, Yn, d_s, d_zoops, d_z, d_mops, d_gout); }
                else        k_score<uint16_t, false><<<grid, 512, 0, st>>>((const uint16_t*)ia->d, sv, d_moff, W, Yn, d_s, d_zoops, d_z, d_mops, d_gout);
            } else {
                if (smem) { CUX(cudaFuncSetAttribute(k_score<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb));
                            k_score<uint32_t, true><<<grid, 512, tb, st>>>((const uint32_t*)ia->d, sv, d_moff, W, Yn, d_s, d_zoops, d_z, d_mops, d_gout); }
                else        k_score<uint32_t, false><<<grid, 512, 0, st>>>((const uint32_t*)ia->d, sv, d_moff, W, Yn, d_s, d_zoops, d_z, d_mops, d_gout);
            }
            CUX(cudaGetLastError());
        }
        CUX(cudaEventRecord(ev1, st));
        tr.mark("kernels");
        CUX(cudaMemcpyAsync(zoops, d_zoops, nsub * 4, cudaMemcpyDeviceToHost, st));
        CUX(cudaMemcpyAsync(z, d_z, nsub * 8, cudaMemcpyDeviceToHost, st));
        if (mops) CUX(cudaMemcpyAsync(mops, d_mops, moff[nsub] * 4, cudaMemcpyDeviceToHost, st));
        if (dev_ids) CUX(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
        CUX(cudaStreamSynchronize(st));
        if (bad) { rc = fail(BAMM_E_INVALID, "subset index out of range"); goto done; }
        CUX(cudaEventElapsedTime(&g_score_ms, ev0, ev1));
        tr.mark("D2H");
    }
done:
#undef CUX
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    cudaFree(d_ztab); cudaFree(d_sub); cudaFree(d_bad);
    cudaFree(d_s); cudaFree(d_zoops); cudaFree(d_z); cudaFree(d_gids); cudaFree(d_gout); cudaFree(d_pids); cudaFree(d_pout); cudaFree(d_moff); cudaFree(d_mops);
    cudaStreamDestroy(st);
    return rc;
}

extern "C" int bamm_score_last_timing(float* kernel_ms) {
    REQUIRE(kernel_ms, "NULL argument");
    *kernel_ms = g_score_ms;
    return BAMM_OK;
}

// ------------------------------------------------------------------------------------------- score statistics (row f-1)
static int device_sort_f32(float* d_keys, uint64_t n, bool descending, cudaStream_t st) {
    if (n < 2) return BAMM_OK;
    float* d_alt = nullptr; void* d_tmp = nullptr; size_t tmp_bytes = 0;
    cudaError_t e = dev_malloc(&d_alt, n * sizeof(float));
    if (e != cudaSuccess) { cudaGetLastError(); return fail(BAMM_E_NOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e)); }
    cub::DoubleBuffer<float> buf(d_keys, d_alt);
    const unsigned long long items = n;                                // 64-bit item count: no 2^31 limit
    if (descending) cub::DeviceRadixSort::SortKeysDescending(nullptr, tmp_bytes, buf, items, 0, 32, st);
    else            cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, buf, items, 0, 32, st);
    e = dev_malloc(&d_tmp, tmp_bytes ? tmp_bytes : 16);
    const bool nomem = e == cudaErrorMemoryAllocation;
    if (e == cudaSuccess) {
        if (descending) e = cub::DeviceRadixSort::SortKeysDescending(d_tmp, tmp_bytes, buf, items, 0, 32, st);
        else            e = cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, buf, items, 0, 32, st);
    }
    if (e == cudaSuccess && buf.Current() != d_keys) e = cudaMemcpyAsync(d_keys, buf.Current(), n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_alt); cudaFree(d_tmp);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nomem ? BAMM_E_NOMEM : BAMM_E_CUDA, "device sort failed: %s", cudaGetErrorString(e)); }
    return BAMM_OK;
}

// one run of at most `n` scores through the device; BAMM_E_NOMEM when the device cannot hold it (2 n floats + workspace)
static int sort_run_on_device(float* scores, uint64_t n, bool descending) {
    float* d = nullptr;
    cudaError_t e = dev_malloc(&d, n * sizeof(float));
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? BAMM_E_NOMEM : BAMM_E_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e)); }
    e = cudaMemcpy(d, scores, n * sizeof(float), cudaMemcpyHostToDevice);
    int rc = e == cudaSuccess ? device_sort_f32(d, n, descending, 0) : fail(BAMM_E_CUDA, "H2D failed: %s", cudaGetErrorString(e));
    if (!rc) { e = cudaMemcpy(scores, d, n * sizeof(float), cudaMemcpyDeviceToHost); if (e != cudaSuccess) rc = fail(BAMM_E_CUDA, "D2H failed: %s", cudaGetErrorString(e)); }
    cudaFree(d);
    return rc;
}

// Sorts like std::sort (the reference's sorts in FDR::calculatePR / calculatePvalues, src/evaluation/FDR.cpp:161-162, 207-208) for
// ANY n: one device sort when the device can hold the vector; otherwise runs of halving size sorted on the device and merged on
// the host; a run the device cannot take at all is sorted on the host. BAMM_SORT_RUN caps the run length (tests).
extern "C" int bamm_sort_scores(float* scores, uint64_t n, int descending) {
    REQUIRE(scores || n == 0, "scores is NULL");
    if (n < 2) return BAMM_OK;
    const bool desc = descending != 0;
    uint64_t run = n;
    if (const char* cap = getenv("BAMM_SORT_RUN")) { const uint64_t c = (uint64_t)atoll(cap); if (c >= 2 && c < run) run = c; }
    auto host_sort = [&](float* a, float* b) { if (desc) std::sort(a, b, std::greater<float>()); else std::sort(a, b); };
    for (;;) {
        int rc = BAMM_OK;
        uint64_t done = 0;
        for (; done < n && rc == BAMM_OK; done += run) rc = sort_run_on_device(scores + done, std::min(run, n - done), desc);
        if (rc == BAMM_OK) break;
        if (rc != BAMM_E_NOMEM) return rc;
        if (run <= (1ull << 20)) { host_sort(scores, scores + n); return BAMM_OK; }   // no room for even a small run: host
        run = (run + 1) / 2;                                                           // sorted prefixes stay sorted: harmless
    }
    for (uint64_t width = run; width < n; width *= 2)                                  // merge neighbouring runs
        for (uint64_t lo = 0; lo + width < n; lo += 2 * width) {
            float* a = scores + lo; float* m = a + width; float* b = scores + std::min(n, lo + 2 * width);
            if (desc) std::inplace_merge(a, m, b, std::greater<float>()); else std::inplace_merge(a, m, b);
        }
    return BAMM_OK;
}

extern "C" int bamm_mops_pvalues(const float* neg_scores, uint64_t nneg, const float* pos_scores, uint64_t npos, uint64_t n_pos_sequences,
                                 float* p_values, float* e_values) {
    REQUIRE(neg_scores && nneg >= 1, "no negative scores");
    REQUIRE((pos_scores && p_values && e_values) || npos == 0, "NULL argument");
    float *d_neg = nullptr, *d_pos = nullptr, *d_p = nullptr, *d_e = nullptr;
    int rc = BAMM_OK;
    const uint64_t CH = 1ull << 26;                                 // positive scores go through in chunks of 64M
    const uint64_t chn = npos < CH ? (npos ? npos : 1) : CH;
#define CUX(call) do { cudaError_t e2_ = (call); if (e2_ != cudaSuccess) { rc = fail(e2_ == cudaErrorMemoryAllocation ? BAMM_E_NOMEM : BAMM_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e2_)); goto done; } } while (0)
    {
        CUX(dev_malloc(&d_neg, nneg * sizeof(float)));
        CUX(cudaMemcpy(d_neg, neg_scores, nneg * sizeof(float), cudaMemcpyHostToDevice));
        rc = device_sort_f32(d_neg, nneg, false, 0);
        if (rc) goto done;
        // rate parameter of the exponential tail from the first nTop sorted values, in the reference's order (ScoreSeqSet.cpp:88-96)
        const size_t nTop = (size_t)std::min<uint64_t>(100, nneg / 10);
        std::vector<float> head(nTop + 1);
        CUX(cudaMemcpy(head.data(), d_neg, (nTop + 1) * sizeof(float), cudaMemcpyDeviceToHost));
        const float S_ntop = head[nTop];
        float lambda = 0.f;
        for (size_t n = 0; n < nTop; n++) lambda += (head[n] - S_ntop);
        lambda = lambda / (float)nTop;
        CUX(dev_malloc(&d_pos, chn * sizeof(float)));
        CUX(dev_malloc(&d_p, chn * sizeof(float)));
        CUX(dev_malloc(&d_e, chn * sizeof(float)));
        int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        for (uint64_t b = 0; b < npos; b += CH) {
            const uint64_t m = std::min(CH, npos - b);
            CUX(cudaMemcpy(d_pos, pos_scores + b, m * sizeof(float), cudaMemcpyHostToDevice));
            k_mops_pvalues<<<sms * 8, 256>>>(d_neg, nneg, d_pos, m, S_ntop, lambda, (float)nTop, (float)n_pos_sequences, d_p, d_e);
            CUX(cudaGetLastError());
            CUX(cudaMemcpy(p_values + b, d_p, m * sizeof(float), cudaMemcpyDeviceToHost));
            CUX(cudaMemcpy(e_values + b, d_e, m * sizeof(float), cudaMemcpyDeviceToHost));
        }
    }
done:
#undef CUX
    cudaFree(d_neg); cudaFree(d_pos); cudaFree(d_p); cudaFree(d_e);
    return rc;
}
