// capi_negatives.inl — part of capi.cu (one translation unit: included there, in this order).
// negative-set sampling (row f-2) and the device-built sets it returns
// ------------------------------------------------------------------------------------------- negative sampling
// glibc srandom_r(seed) for the TYPE_3 generator: the 31 words u_m = r[3+m] the recurrence starts from, and x^(2^b)
// modulo its characteristic polynomial (negatives.cuh)
struct LfgTables { uint32_t u0[LFG_N]; uint32_t pw[LFG_NPOW * LFG_N]; };
static void lfg_tables(uint32_t seed, LfgTables& t) {
    int32_t r[34];
    r[0] = seed ? (int32_t)seed : 1;
    for (int i = 1; i < 31; i++) {
        const long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (int32_t)w;
    }
    for (int i = 31; i < 34; i++) r[i] = r[i - 31];
    for (int m = 0; m < LFG_N; m++) t.u0[m] = (uint32_t)r[3 + m];
    for (int k = 0; k < LFG_N; k++) t.pw[k] = k == 1 ? 1u : 0u;                     // x
    for (int b = 1; b < LFG_NPOW; b++) {
        uint32_t* cur = t.pw + b * LFG_N;
        memcpy(cur, t.pw + (b - 1) * LFG_N, LFG_N * sizeof(uint32_t));
        lfg_poly_mul(cur, t.pw + (b - 1) * LFG_N);
    }
}

extern "C" int bamm_rand_stream(uint32_t seed, uint64_t first, uint64_t count, int32_t* out) {
    REQUIRE(out || count == 0, "out is NULL");
    REQUIRE(first + count + 400 < (1ull << (LFG_NPOW - 1)), "draw index out of range");
    if (!count) return BAMM_OK;
    LfgTables t; lfg_tables(seed, t);
    uint32_t* d_t = nullptr; int* d_out = nullptr;
    CU(dev_malloc(&d_t, sizeof(t)));
    cudaError_t e = dev_malloc(&d_out, count * sizeof(int));
    if (e != cudaSuccess) { cudaFree(d_t); return fail(BAMM_E_NOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e)); }
    cudaMemcpy(d_t, &t, sizeof(t), cudaMemcpyHostToDevice);
    const uint64_t threads = std::min<uint64_t>(count, 148ull * 1024ull), per = (count + threads - 1) / threads;
    k_rand_stream<<<(unsigned)((threads + 127) / 128), 128>>>(d_t, d_t + LFG_N, first, count, per, d_out);
    e = cudaMemcpy(out, d_out, count * sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d_t); cudaFree(d_out);
    if (e != cudaSuccess) return fail(BAMM_E_CUDA, "rand stream kernel failed: %s", cudaGetErrorString(e));
    return BAMM_OK;
}

static int sample_negatives_impl(bamm_seqset* pos, const uint64_t* subset, uint64_t nsub, uint64_t fold, uint32_t seed, uint64_t draw_offset,
                                 const uint64_t* global_counts, uint64_t* local_counts_out, bamm_seqset** out);

extern "C" int bamm_seqset_sample_negatives(bamm_seqset* pos, const uint64_t* subset, uint64_t nsub, uint64_t fold, uint32_t seed, bamm_seqset** out) {
    return sample_negatives_impl(pos, subset, nsub, fold, seed, 0, nullptr, nullptr, out);
}

extern "C" int bamm_seqset_negative_kmer_counts(bamm_seqset* pos, const uint64_t* subset, uint64_t nsub, uint64_t* counts) {
    REQUIRE(counts, "counts is NULL");
    return sample_negatives_impl(pos, subset, nsub, 1, 42, 0, nullptr, counts, nullptr);
}

extern "C" int bamm_seqset_sample_negatives_shard(bamm_seqset* pos, const uint64_t* subset, uint64_t nsub, uint64_t fold, uint32_t seed,
                                                  uint64_t draw_offset, const uint64_t* global_counts, bamm_seqset** out) {
    REQUIRE(global_counts, "global_counts is NULL");
    return sample_negatives_impl(pos, subset, nsub, fold, seed, draw_offset, global_counts, nullptr, out);
}

// out == nullptr: only the template list's k-mer counts are wanted (local_counts_out)
static int sample_negatives_impl(bamm_seqset* pos, const uint64_t* subset, uint64_t nsub, uint64_t fold, uint32_t seed, uint64_t draw_offset,
                                 const uint64_t* global_counts, uint64_t* local_counts_out, bamm_seqset** out) {
    REQUIRE(out || local_counts_out, "out is NULL");
    bamm_seqset* dummy_out = nullptr;
    if (!out) out = &dummy_out;
    *out = nullptr;
    REQUIRE(pos, "seqset is NULL");
    REQUIRE(fold >= 1, "fold must be at least 1");
    if (!subset) nsub = pos->nseq;
    REQUIRE(nsub >= 1 && nsub * fold < (1ull << 32), "number of negative sequences out of range");
    // the template list: prefix sums of its lengths (= draw offsets / fold) and, for a true subset, its sequence ids
    std::vector<uint64_t> toff(nsub + 1, 0);
    std::vector<uint32_t> tids(subset ? nsub : 0);
    for (uint64_t i = 0; i < nsub; i++) {
        const uint64_t n = subset ? subset[i] : i;
        REQUIRE(n < pos->nseq, "subset index out of range");
        const uint64_t L = pos->h_off[n + 1] - pos->h_off[n];
        REQUIRE(L >= 1, "empty template sequence");
        toff[i + 1] = toff[i] + L;
        if (subset) tids[i] = (uint32_t)n;
    }
    REQUIRE(draw_offset + toff[nsub] * fold + 400 < (1ull << (LFG_NPOW - 1)), "too many draws");
    Trace tr("sample_negatives");
    CU(cudaSetDevice(pos->device));
    NegDims d; d.A = pos->A; d.Y1 = (uint32_t)pos->A; d.Y2 = d.Y1 * d.Y1; d.Y3 = d.Y2 * d.Y1; d.total = d.Y1 + d.Y2 + d.Y3;
    IndexArray* ia = nullptr;
    { std::lock_guard<std::mutex> g(pos->mu); int rc = seqset_index_locked(pos, 2, &ia); if (rc) return rc; }
    const uint16_t* Y2 = (const uint16_t*)ia->d;
    const float pc = 20.0f;                                   // SeqGenerator.cpp:30-32: A_[k] = 20 for every order
    unsigned long long* d_cnt = nullptr; float *d_v = nullptr, *d_rb = nullptr; uint32_t *d_lfg = nullptr, *d_flags = nullptr, *d_tids = nullptr;
    uint64_t* d_toff = nullptr;
    bamm_seqset* neg = nullptr;
    int rc = BAMM_OK;
#define CUX(call) do { cudaError_t e2_ = (call); if (e2_ != cudaSuccess) { rc = fail(e2_ == cudaErrorMemoryAllocation ? BAMM_E_NOMEM : BAMM_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e2_)); goto done; } } while (0)
    {
        // set-wide frequencies (SeqGenerator::calculate_kmer_frequency, SeqGenerator.cpp:63-112): counts on the device, the
        // 84 probabilities on the host in the reference's operation order
        if (subset) {
            CUX(dev_malloc(&d_tids, nsub * sizeof(uint32_t)));
            CUX(cudaMemcpy(d_tids, tids.data(), nsub * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
        CUX(dev_malloc(&d_toff, (nsub + 1) * sizeof(uint64_t)));
        CUX(cudaMemcpy(d_toff, toff.data(), (nsub + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
        CUX(dev_malloc(&d_cnt, d.total * sizeof(unsigned long long)));
        CUX(cudaMemset(d_cnt, 0, d.total * sizeof(unsigned long long)));
        tr.mark("order-2 index + template list");
        k_neg_count_set<<<pos->sm_count * 8, 256>>>(Y2, pos->d_off, d_tids, nsub, d, d_cnt);
        CUX(cudaGetLastError());
        std::vector<unsigned long long> cnt(d.total);
        CUX(cudaMemcpy(cnt.data(), d_cnt, d.total * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        if (local_counts_out) {                                // the caller sums these over the shards of one set
            for (uint32_t b = 0; b < d.total; b++) local_counts_out[b] = cnt[b];
            goto done;
        }
        if (global_counts) for (uint32_t b = 0; b < d.total; b++) cnt[b] = global_counts[b];
        std::vector<float> v(d.total), rb0(d.Y1);
        const unsigned long long *n0 = cnt.data(), *n1 = n0 + d.Y1, *n2 = n1 + d.Y2;
        float *v0 = v.data(), *v1 = v0 + d.Y1, *v2 = v1 + d.Y2;
        size_t normFactor = 0;
        for (uint32_t y = 0; y < d.Y1; y++) normFactor += n0[y];
        float sum = 0.0f;
        for (uint32_t y = 0; y < d.Y1; y++) {
            v0[y] = ((float)n0[y] + pc * 0.25f) / ((float)normFactor + pc);
            sum += v0[y];
            rb0[y] = sum;
        }
        for (uint32_t y = 0; y < d.Y2; y++) v1[y] = ((float)n1[y] + pc * v0[y % d.Y1]) / ((float)n0[y / d.Y1] + pc);
        for (uint32_t y = 0; y < d.Y3; y++) v2[y] = ((float)n2[y] + pc * v1[y % d.Y2]) / ((float)n1[y / d.Y1] + pc);
        CUX(dev_malloc(&d_v, (d.total + d.Y1) * sizeof(float)));
        CUX(cudaMemcpy(d_v, v.data(), d.total * sizeof(float), cudaMemcpyHostToDevice));
        CUX(cudaMemcpy(d_v + d.total, rb0.data(), d.Y1 * sizeof(float), cudaMemcpyHostToDevice));
        // per-template bars
        CUX(dev_malloc(&d_rb, nsub * (uint64_t)(d.Y2 + d.Y3) * sizeof(float)));
        k_neg_models<<<pos->sm_count * 16, 128>>>(Y2, pos->d_off, d_tids, nsub, d, d_v, pc, d_rb);
        CUX(cudaGetLastError());
        // the negative set: `fold` records per template, each of the template's stored length
        const uint64_t nneg = nsub * fold;
        std::vector<uint64_t> noff(nneg + 1);
        noff[0] = 0;
        for (uint64_t i = 0, g = 0; i < nsub; i++) {
            const uint64_t L = toff[i + 1] - toff[i];
            for (uint64_t m = 0; m < fold; m++, g++) noff[g + 1] = noff[g] + L;
        }
        tr.mark("set-wide model + per-template bars + offsets (host)");
        rc = seqset_new(noff.data(), nneg, pos->A, &neg, nullptr, &noff);
        if (rc) goto done;
        tr.mark("seqset_new (alloc + offsets H2D)");
        LfgTables t; lfg_tables(seed, t);
        CUX(dev_malloc(&d_lfg, sizeof(t)));
        CUX(cudaMemcpy(d_lfg, &t, sizeof(t), cudaMemcpyHostToDevice));
        CUX(dev_malloc(&d_flags, sizeof(uint32_t)));
        CUX(cudaMemset(d_flags, 0, sizeof(uint32_t)));
        const uint64_t want = (uint64_t)pos->sm_count * 2048ull;
        const uint64_t per = (nneg + want - 1) / want;
        const uint64_t threads = (nneg + per - 1) / per;
        k_neg_sample<<<(unsigned)((threads + NEG_THREADS - 1) / NEG_THREADS), NEG_THREADS>>>(d_toff, nsub, fold, d, d_v + d.total, d_rb,
                                                                                          d_lfg, d_lfg + LFG_N, per, draw_offset, neg->d_codes, d_flags);
        CUX(cudaGetLastError());
        uint32_t flags = 0;
        CUX(cudaMemcpy(&flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost));
        if (flags) {
            rc = fail(BAMM_E_STATE, "a sampled sequence starts with an undetermined base (draw above the last cumulative bar): "
                                    "the reference's rand() stream diverges here, use the host sampler for this set");
            goto done;
        }
        tr.mark("sampling kernel");
        rc = seqset_finish(neg, true);                         // sampled codes are 1..A by construction (flags checked above)
        tr.mark("classify + pack");
        if (rc) { neg = nullptr; goto done; }                  // seqset_finish destroys the set on failure
    }
done:
#undef CUX
    cudaFree(d_cnt); cudaFree(d_v); cudaFree(d_rb); cudaFree(d_lfg); cudaFree(d_flags); cudaFree(d_tids); cudaFree(d_toff);
    if (rc) { if (neg) bamm_seqset_destroy(neg); return rc; }
    *out = neg;
    return BAMM_OK;
}

extern "C" int bamm_seqset_get_codes(bamm_seqset* s, uint8_t* out) {
    REQUIRE(s && out, "NULL argument");
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpy(out, s->d_codes, s->npos, cudaMemcpyDeviceToHost));
    return BAMM_OK;
}

extern "C" int bamm_seqset_get_offsets(const bamm_seqset* s, uint64_t* out) {
    REQUIRE(s && out, "NULL argument");
    memcpy(out, s->h_off.data(), (s->nseq + 1) * sizeof(uint64_t));
    return BAMM_OK;
}
