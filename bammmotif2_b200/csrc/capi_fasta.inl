// capi_fasta.inl — part of capi.cu (one translation unit: included there, in this order).
// FASTA text -> device set (row f-3) and PWM site sampling (row f-4)
// ------------------------------------------------------------------------------------------- FASTA text -> device set (row f-3)
static_assert(sizeof(FastaSeg) == 24, "FastaSeg layout is part of the C ABI (bamm_fasta_seg)");

extern "C" int bamm_seqset_encode_text(const char* text, uint64_t nbytes, const bamm_fasta_seg* segs, uint64_t nseg,
                                       const uint64_t* offsets, const uint32_t* rec_L0, uint64_t nrec, int single_strand, int A,
                                       const uint8_t* base2code, const uint8_t* code2comp, uint64_t* base_counts, uint64_t* n_forward_zeros,
                                       bamm_seqset** out) {
    REQUIRE(out, "out is NULL");
    *out = nullptr;
    REQUIRE(text && segs && offsets && rec_L0 && base2code && code2comp && base_counts && n_forward_zeros, "NULL argument");
    REQUIRE(A >= 2 && A <= 6, "alphabet size %d not in [2,6]", A);
    Trace tr("encode_text");
    bamm_seqset* s = nullptr;
    { int rc = seqset_new(offsets, nrec, A, &s); if (rc) return rc; }
    uint8_t *d_text = nullptr, *d_lut = nullptr; FastaSeg* d_segs = nullptr; uint32_t* d_L0 = nullptr;
    unsigned long long *d_cnt = nullptr;
    int rc = BAMM_OK;
    const uint64_t zero_cap = std::max<uint64_t>(1024, s->npos / 16);          // forward undefined bases kept (more => error below)
#define CUX(call) do { cudaError_t e2_ = (call); if (e2_ != cudaSuccess) { rc = fail(e2_ == cudaErrorMemoryAllocation ? BAMM_E_NOMEM : BAMM_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e2_)); goto done; } } while (0)
    {
        CUX(dev_malloc(&d_text, nbytes ? nbytes : 1));
        CUX(cudaMemcpy(d_text, text, nbytes, cudaMemcpyHostToDevice));
        CUX(dev_malloc(&d_segs, (nseg ? nseg : 1) * sizeof(FastaSeg)));
        CUX(cudaMemcpy(d_segs, segs, nseg * sizeof(FastaSeg), cudaMemcpyHostToDevice));
        CUX(dev_malloc(&d_L0, (nrec ? nrec : 1) * sizeof(uint32_t)));
        CUX(cudaMemcpy(d_L0, rec_L0, nrec * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CUX(dev_malloc(&d_lut, 512));
        CUX(cudaMemcpy(d_lut, base2code, 256, cudaMemcpyHostToDevice));
        CUX(cudaMemcpy(d_lut + 256, code2comp, 256, cudaMemcpyHostToDevice));
        CUX(dev_malloc(&d_cnt, 16 * sizeof(unsigned long long)));
        CUX(cudaMemset(d_cnt, 0, 16 * sizeof(unsigned long long)));
        CUX(dev_malloc(&s->d_zero_pos, zero_cap * sizeof(unsigned long long)));
        tr.mark("alloc + text H2D");
        if (nseg) {
            k_fasta_encode<<<s->sm_count * 8, 256>>>(d_text, d_segs, nseg, s->d_off, d_L0, single_strand, d_lut, d_lut + 256, A, s->d_codes,
                                                     d_cnt, (unsigned long long*)s->d_zero_pos, zero_cap, d_cnt + 8);
            CUX(cudaGetLastError());
        }
        unsigned long long h[16];
        CUX(cudaMemcpy(h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost));
        tr.mark("encode kernel");
        for (int a = 0; a < A; a++) base_counts[a] = h[a];
        if (h[8] > zero_cap) { rc = fail(BAMM_E_INVALID, "more than 1/16 of the bases are undefined: use the host encoder"); goto done; }
        s->n_zero_fwd = h[8];
        *n_forward_zeros = h[8];
    }
done:
#undef CUX
    cudaFree(d_text); cudaFree(d_segs); cudaFree(d_L0); cudaFree(d_lut); cudaFree(d_cnt);
    if (rc) { bamm_seqset_destroy(s); return rc; }
    *out = s;
    return BAMM_OK;
}

extern "C" int bamm_seqset_forward_zeros(bamm_seqset* s, uint64_t* positions) {
    REQUIRE(s && (positions || s->n_zero_fwd == 0), "NULL argument");
    if (s->n_zero_fwd) CU(cudaMemcpy(positions, s->d_zero_pos, s->n_zero_fwd * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return BAMM_OK;
}

extern "C" int bamm_seqset_code_windows(bamm_seqset* s, const uint64_t* zpos, const uint64_t* zbeg, const uint64_t* zend, uint64_t nz, uint8_t* windows) {
    REQUIRE(s && ((zpos && zbeg && zend && windows) || nz == 0), "NULL argument");
    if (!nz) return BAMM_OK;
    uint64_t* d = nullptr; uint8_t* d_w = nullptr;
    CU(dev_malloc(&d, 3 * nz * sizeof(uint64_t)));
    cudaError_t e = dev_malloc(&d_w, nz * 21);
    if (e == cudaSuccess) e = cudaMemcpy(d, zpos, nz * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + nz, zbeg, nz * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + 2 * nz, zend, nz * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_zero_windows<<<(unsigned)((nz * 21 + 255) / 256), 256>>>(s->d_codes, d, d + nz, d + 2 * nz, nz, d_w);
        e = cudaMemcpy(windows, d_w, nz * 21, cudaMemcpyDeviceToHost);
    }
    cudaFree(d); cudaFree(d_w);
    if (e != cudaSuccess) return fail(BAMM_E_CUDA, "code windows failed: %s", cudaGetErrorString(e));
    return BAMM_OK;
}

extern "C" int bamm_seqset_finish_patches(bamm_seqset* s, const uint64_t* patch_pos, const uint64_t* patch_kmer, uint64_t npatch) {
    REQUIRE(s && ((patch_pos && patch_kmer) || npatch == 0), "NULL argument");
    REQUIRE(!s->d_pseq && !s->d_kind, "the set is already finished");
    cudaFree(s->d_zero_pos); s->d_zero_pos = nullptr;
    s->npatch = npatch;
    if (npatch) {
        CU(dev_malloc(&s->d_ppos, npatch * sizeof(uint64_t)));
        CU(dev_malloc(&s->d_pkmer, npatch * sizeof(uint64_t)));
        CU(cudaMemcpy(s->d_ppos, patch_pos, npatch * sizeof(uint64_t), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(s->d_pkmer, patch_kmer, npatch * sizeof(uint64_t), cudaMemcpyHostToDevice));
        uint32_t* d_bad = nullptr; uint32_t bad = 0;
        CU(dev_malloc(&d_bad, sizeof(uint32_t)));
        cudaMemset(d_bad, 0, sizeof(uint32_t));
        k_validate_patches<<<(unsigned)((npatch + 255) / 256), 256>>>(s->d_ppos, npatch, s->npos, d_bad);
        cudaError_t ev = cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
        cudaFree(d_bad);
        CU(ev);
        if (bad) return fail(BAMM_E_INVALID, bad & 1u ? "patch position out of range" : "patch positions must be strictly increasing");
    }
    return seqset_finish(s);                                   // destroys the set on failure
}

// ------------------------------------------------------------------------------------------- Motif::initFromPWM sampling (row f-4)
extern "C" int bamm_seqset_sample_pwm_sites(bamm_seqset* s, const uint64_t* subset, uint64_t nsub, int W, int K, int asize,
                                            const float* score, float q, const double* uniforms, int32_t* n_all, uint64_t* z_out) {
    REQUIRE(s && score && uniforms && n_all, "NULL argument");
    REQUIRE(W >= 1 && W <= 32, "motif width W=%d not in [1,32]", W);
    REQUIRE(K >= 0 && K <= 10, "order K=%d not in [0,10]", K);
    REQUIRE(asize >= 1 && asize <= 6, "PWM alphabet size %d not in [1,6]", asize);
    if (!subset) nsub = s->nseq;
    REQUIRE(nsub < (1ull << 32), "subset too large");
    ModelDims d; fill_dims(d, s->A, K, W, 0);
    const size_t msize = d.voff[K + 1];
    std::vector<uint32_t> ids(nsub);
    uint64_t maxL = 0;
    for (uint64_t i = 0; i < nsub; i++) {
        const uint64_t n = subset ? subset[i] : i;
        REQUIRE(n < s->nseq, "subset index out of range");
        const uint64_t L = s->h_off[n + 1] - s->h_off[n];
        REQUIRE(L >= (uint64_t)W, "sequence %llu is shorter than the motif", (unsigned long long)n);
        ids[i] = (uint32_t)n;
        maxL = std::max(maxL, L);
    }
    const int Kidx = K > 1 ? K : 1;                              // kmer % asize needs an index whose modulus asize divides (6^2 = 36 for the 4-letter PWM on ACGTMH)
    IndexArray* ia = nullptr;
    { std::lock_guard<std::mutex> g(s->mu); int rc = seqset_index_locked(s, Kidx, &ia); if (rc) return rc; }
    CU(cudaSetDevice(s->device));
    const int grid = s->sm_count * 8, warps = grid * 8;
    const uint64_t stride = ((maxL + 1 + 31) / 32) * 32;
    uint32_t *d_ids = nullptr, *d_voff = nullptr; float *d_score = nullptr, *d_scratch = nullptr; double* d_u = nullptr; int* d_n = nullptr;
    unsigned long long* d_z = nullptr;
    int rc = BAMM_OK;
#define CUX(call) do { cudaError_t e2_ = (call); if (e2_ != cudaSuccess) { rc = fail(e2_ == cudaErrorMemoryAllocation ? BAMM_E_NOMEM : BAMM_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e2_)); goto done; } } while (0)
    {
        CUX(dev_malloc(&d_ids, (nsub ? nsub : 1) * 4));
        CUX(cudaMemcpy(d_ids, ids.data(), nsub * 4, cudaMemcpyHostToDevice));
        CUX(dev_malloc(&d_voff, 16 * 4));
        CUX(cudaMemcpy(d_voff, d.voff, 16 * 4, cudaMemcpyHostToDevice));
        CUX(dev_malloc(&d_score, (size_t)asize * W * 4));
        CUX(cudaMemcpy(d_score, score, (size_t)asize * W * 4, cudaMemcpyHostToDevice));
        CUX(dev_malloc(&d_u, (nsub ? nsub : 1) * 8));
        CUX(cudaMemcpy(d_u, uniforms, nsub * 8, cudaMemcpyHostToDevice));
        CUX(dev_malloc(&d_scratch, (uint64_t)warps * stride * 4));
        CUX(dev_malloc(&d_n, msize * 4));
        CUX(cudaMemset(d_n, 0, msize * 4));
        if (z_out) CUX(dev_malloc(&d_z, (nsub ? nsub : 1) * 8));
        if (ia->bytes == 2) k_pwm_sample_sites<uint16_t><<<grid, 256>>>((const uint16_t*)ia->d, s->d_off, d_ids, (uint32_t)nsub, W, K, (uint32_t)s->A, (uint32_t)asize,
                                                                        d_score, q, d_u, d_scratch, stride, d_n, d_voff, d_z);
        else                k_pwm_sample_sites<uint32_t><<<grid, 256>>>((const uint32_t*)ia->d, s->d_off, d_ids, (uint32_t)nsub, W, K, (uint32_t)s->A, (uint32_t)asize,
                                                                        d_score, q, d_u, d_scratch, stride, d_n, d_voff, d_z);
        CUX(cudaGetLastError());
        CUX(cudaMemcpy(n_all, d_n, msize * 4, cudaMemcpyDeviceToHost));
        if (z_out) CUX(cudaMemcpy(z_out, d_z, nsub * 8, cudaMemcpyDeviceToHost));
    }
done:
#undef CUX
    cudaFree(d_ids); cudaFree(d_voff); cudaFree(d_score); cudaFree(d_u); cudaFree(d_scratch); cudaFree(d_n); cudaFree(d_z);
    return rc;
}
