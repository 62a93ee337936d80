// capi_em.inl — part of capi.cu (one translation unit: included there, in this order).
// EM objects: column-group planner, create / set_model, E-step / M-step / update launches, optimize, mask, NVLink peer exchange
// ------------------------------------------------------------------------------------------- EM
// device groups (capi_group.inl): an EM object over several devices is a facade over one shard object per device
static bool group_wanted(const bamm_seqset* s, uint64_t nsub);
static int group_create(bamm_seqset* s, const uint64_t* subset, uint64_t nsub, int W, int K, int K_bg_model, bamm_em** out);
static void group_destroy(bamm_em* em);
static int group_set_model(bamm_em* em, const float* v_all, const float* vbg_all, const float* alpha, float q);
static int group_optimize(bamm_em* em, int optimize_q, float epsilon, int max_iter, int* iterations, float* llh_trace, float* vdiff_trace, float* q_trace);
static int group_iterate(bamm_em* em, int n_iter, float* llh_last, float* vdiff_last);
static int group_estep(bamm_em* em, float* llh);
static int group_mstep(bamm_em* em);
static int group_get_r(bamm_em* em, uint64_t first, uint64_t count, float* out);
#define IS_GROUP(em) ((em) && !(em)->shards.empty())
#define NOT_FOR_GROUPS(em, what) REQUIRE(!IS_GROUP(em), what " is not available on an EM object that spans a device group")
static void fill_dims(ModelDims& d, int A, int K, int W, int K_bg) {
    memset(&d, 0, sizeof(d));
    d.A = A; d.K = K; d.W = W; d.K_bg = K_bg;
    uint64_t p = 1;
    for (int i = 0; i < 16; i++) { d.Y[i] = (uint32_t)(p > 0xffffffffull ? 0xffffffffull : p); p *= A; }
    uint32_t vo = 0, bo = 0;
    for (int k = 0; k < 16; k++) {
        d.voff[k] = vo; d.bgoff[k] = bo;
        if (k <= K + 1) { vo += d.Y[k + 1 < 16 ? k + 1 : 15] * (uint32_t)W; }
        if (k <= 12) bo += d.Y[k + 1 < 16 ? k + 1 : 15];
    }
}

extern "C" void bamm_em_destroy(bamm_em* em) {
    if (!em) return;
    if (IS_GROUP(em)) { group_destroy(em); return; }
    cudaSetDevice(em->device);
    if (em->stream) cudaStreamSynchronize(em->stream);
    for (int p = 0; p < MAX_PEERS; p++) if (em->peer_mapped[p]) cudaIpcCloseMemHandle(em->peer_mapped[p]);
    cudaFree(em->d_peer_local); cudaFree(em->d_peer_done); cudaFree(em->d_peer_wait);
    cudaFree(em->d_act); cudaFree(em->d_scale); cudaFree(em->d_act_cnt); cudaFree(em->d_overflow); cudaFree(em->d_reg_off);
    cudaFree(em->d_gen_ids); cudaFree(em->d_gen_roff); cudaFree(em->d_pk_ids); cudaFree(em->d_pk_roff); cudaFree(em->d_tab); cudaFree(em->d_tab_alt);
    cudaFree(em->d_btab); cudaFree(em->d_U); cudaFree(em->d_cand); cudaFree(em->d_cand_seq); cudaFree(em->d_seqacc); cudaFree(em->d_cand_part); cudaFree(em->d_mask_part); cudaFree(em->d_creg_off); cudaFree(em->d_eflags);
    cudaFree(em->d_s_alt); cudaFree(em->d_sT_alt); cudaFree(em->d_rows);
    cudaFree(em->d_seq_ids); cudaFree(em->d_r_off); cudaFree(em->d_r); cudaFree(em->d_s); cudaFree(em->d_sT); cudaFree(em->d_v);
    cudaFree(em->d_vK_prev); cudaFree(em->d_n); cudaFree(em->d_vbg); cudaFree(em->d_alpha); cudaFree(em->d_part);
    if (em->own_xbuf) cudaFree(em->d_xbuf);
    cudaFree(em->d_vdiff); cudaFree(em->d_vdiff_part);
    cudaFree(em->d_m_ids); cudaFree(em->d_m_roff); cudaFree(em->d_m_woff); cudaFree(em->d_m_seloff); cudaFree(em->d_m_sel);
    for (cudaEvent_t e : em->loop_ev) cudaEventDestroy(e);
    pinned_scalars_put(em->h_scal);                    // h_vdiff lives in the same slot
    for (int i = 0; i < 4; i++) if (em->ev[i]) cudaEventDestroy(em->ev[i]);
    if (em->stream) cudaStreamDestroy(em->stream);
    delete em;
}

static int mstep_w_dispatch(bamm_em* em, const PackedView* pv, const Plan* pl, int mode);

template <typename K> static int max_smem_optin(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess ? 0 : -1;
}


// ---- column-group planner of the packed E-step (packed.cuh, "column groups") -----------------------------------------
// Cuts columns 0..W-1 into the fewest consecutive groups whose lookup tables (4^bases floats each) fit `budget` bytes.
// reduced: columns j < K only depend on max(j, K_bg)+1 bases (true for every model produced by updateV; checked for
// models passed to bamm_em_set_model). Returns false when even one column per group does not fit.
static bool make_group_plan(int W, int K, int K_bg, bool reduced, size_t budget, int ca, int cb, GroupPlan& gp, bool& fast) {
    auto ctx = [&](int j) { int c = j < K ? j : K; if (!reduced) c = K; return c > K_bg ? c : K_bg; };
    auto first_base = [&](int a, int b) { int lo = 1 << 30; for (int j = a; j <= b; j++) lo = std::min(lo, j - ctx(j)); return lo; };
    const double INF = 1e300;
    auto bytes_of = [&](int a, int b) {               // table of the group covering columns a..b
        const int nb = b - first_base(a, b) + 1;
        return nb > 12 ? INF : 4.0 * (double)(1ull << (2 * nb));
    };
    // sfx[j][n]: least bytes covering columns [j,W) with n groups
    std::vector<std::vector<double>> sfx(W + 1, std::vector<double>(MAXG + 1, INF));
    std::vector<std::vector<int>> nxt(W + 1, std::vector<int>(MAXG + 1, -1));
    sfx[cb][0] = 0;
    for (int j = cb - 1; j >= ca; j--)
        for (int n = 1; n <= MAXG; n++)
            for (int e = j + 1; e <= cb; e++) {
                if (sfx[e][n - 1] >= INF) continue;
                const double c = bytes_of(j, e - 1) + sfx[e][n - 1];
                if (c < sfx[j][n]) { sfx[j][n] = c; nxt[j][n] = e; }
            }
    // fewest groups first; then a first group wide enough for the one-shift extraction (see `fast` below); then bytes
    int G = -1, best_a1 = -1; bool best_fast = false; double best_bytes = INF;
    for (int n = 1; n <= MAXG && G < 0; n++) {
        for (int a1 = ca + 1; a1 <= cb; a1++) {
            const double c = bytes_of(ca, a1 - 1) + sfx[a1][n - 1];
            if (c > (double)budget) continue;
            const bool f = std::max(K - ca, 15 - (a1 - 1)) <= 31 - cb;
            if (G < 0 || (f && !best_fast) || (f == best_fast && c < best_bytes)) { G = n; best_a1 = a1; best_fast = f; best_bytes = c; }
        }
    }
    if (G < 0) return false;
    memset(&gp, 0, sizeof(gp));
    gp.W = W; gp.K = K; gp.G = G; gp.Yn = 1u << (2 * (K + 1));
    std::vector<int> cuts(G + 1);
    cuts[0] = ca; cuts[1] = best_a1;
    for (int g = 1, j = best_a1; g < G; g++) { j = nxt[j][G - g]; cuts[g + 1] = j; }
    uint32_t base = 0;
    for (int g = 0; g < G; g++) {
        const int a = cuts[g], b = cuts[g + 1] - 1;
        gp.col0[g] = a; gp.ncol[g] = b - a + 1; gp.lo[g] = first_base(a, b);
        const int nb = b - gp.lo[g] + 1;
        gp.base[g] = base;
        gp.mask4[g] = (uint32_t)(((1ull << (2 * nb)) - 1ull) << 2);
        gp.colmask[g] = (uint32_t)(((b >= 31 ? 0xffffffffull : ((2ull << b) - 1ull))) & ~((1ull << a) - 1ull));
        base += 4u << (2 * nb);
    }
    gp.table_bytes = base;
    gp.passmask = (uint32_t)(((cb >= 32 ? 0x100000000ull : (1ull << cb)) - 1ull) & ~((1ull << ca) - 1ull));
    gp.pass_first = ca == 0; gp.pass_last = cb == W;
    // alignment of the window word (32 bases from p-kd): base p+hi sits at bit 62-2(hi+kd); the byte offset of a group's
    // entry needs shift = 60-2(hi+kd) >= 0, i.e. kd <= 31-cb; the oldest base any column of the pass reads is p+ca-K, i.e.
    // kd >= K-ca; the one-shift extraction needs every shift <= 31, i.e. kd >= 15-hi0
    const int hi0 = cuts[1] - 1;
    const int kd_min = K - ca, kd_max = 31 - cb;
    if (kd_min > kd_max) return false;
    const int kd_fast = std::max(kd_min, 15 - hi0);
    fast = kd_fast <= kd_max;
    const int kd = fast ? kd_fast : kd_min;
    gp.kd = kd;
    for (int g = 0; g < G; g++) {
        const int hi = cuts[g + 1] - 1;
        const int sh = 60 - 2 * (hi + kd);
        gp.shift[g] = (uint32_t)sh;
        gp.shift2[g] = sh > 32 ? (uint32_t)(sh - 32) : 0u;
    }
    return true;
}

// Column passes: the fewest lookups per window over all cuts of [0,W) into consecutive column ranges whose group tables
// fit `budget` each; every extra pass costs one read and one write of r, weighted like PASS_COST lookups.
static bool plan_passes(int W, int K, int K_bg, bool reduced, size_t budget, std::vector<GroupPlan>& plans, std::vector<char>& fast) {
    const int PASS_COST = 3, INF = 1 << 28;
    std::vector<int> best(W + 1, INF), from(W + 1, -1);
    best[0] = 0;
    GroupPlan gp; bool f;
    if (make_group_plan(W, K, K_bg, reduced, budget, 0, W, gp, f)) { plans.assign(1, gp); fast.assign(1, (char)f); return true; }
    for (int b = 1; b <= W; b++)
        for (int a = 0; a < b; a++) {
            if (best[a] >= INF || !make_group_plan(W, K, K_bg, reduced, budget, a, b, gp, f)) continue;
            const int c = best[a] + gp.G + PASS_COST;
            if (c < best[b]) { best[b] = c; from[b] = a; }
        }
    if (best[W] >= INF) return false;
    std::vector<int> cuts;
    for (int b = W; b > 0; b = from[b]) cuts.push_back(b);
    cuts.push_back(0);
    std::reverse(cuts.begin(), cuts.end());
    plans.clear(); fast.clear();
    for (size_t i = 0; i + 1 < cuts.size(); i++) {
        make_group_plan(W, K, K_bg, reduced, budget, cuts[i], cuts[i + 1], gp, f);
        plans.push_back(gp); fast.push_back((char)f);
    }
    return true;
}

// ---- bound plan of the pruned E-step (estep.cuh) -------------------------------------------------------------------
// Groups of T <= 6 bases whose tables are indexed by T+1 bases (two neighbouring windows per entry, k_make_bound_tables): the
// fewest groups (6-base groups first: 64 KB each, then 5 bases: 16 KB, ...) whose base ranges cover columns 0..W-1 within
// `budget`. Spare bases: one to the left of the first group (background context of column 0), the rest as overlaps of the
// LAST groups with their predecessors, so that fewer leading columns of those groups lose context. Group g owns the columns
// right of group g-1's last base. Returns false when W columns cannot be covered (or W+1 bases do not fit one window word).
static bool make_bound_plan(int W, int K, int K_bg, size_t budget, GroupPlan& gp, bool& fast) {
    if (W + 1 > 31) return false;
    std::vector<int> T;
    {
        size_t bytes = 0; int bases = 0;
        while (bases < W && (int)T.size() < 8) {
            int t = 6;
            while (t >= 1 && bytes + ((size_t)4 << (2 * (t + 1))) > budget) t--;
            if (t < 1) return false;
            // the last group only takes what it needs (plus context it can use)
            if (bases + t > W + K) t = std::max(1, W + K - bases);
            T.push_back(t); bytes += (size_t)4 << (2 * (t + 1)); bases += t;
        }
        if (bases < W) return false;
    }
    const int G = (int)T.size();
    int spare = 0; for (int t : T) spare += t; spare -= W;
    memset(&gp, 0, sizeof(gp));
    gp.W = W; gp.K = K; gp.G = G; gp.Yn = 1u << (2 * (K + 1));
    const int lead = std::min(std::min(spare, 31 - (W + 1)), std::min(1, std::min(K_bg, K)));    // the window word must still hold base p+W
    spare -= lead;
    std::vector<int> ov(G, 0);
    for (int g = G - 1; g >= 1 && spare > 0; g--) { ov[g] = std::min(spare, std::min(K, T[g] - 1)); spare -= ov[g]; }
    int lo = -lead - spare;                                  // anything still left goes to the front as well
    uint32_t base = 0;
    int prev_hi = -1;
    for (int g = 0; g < G; g++) {
        if (g > 0) lo = prev_hi + 1 - ov[g];
        int hi = lo + T[g] - 1;
        if (hi > W - 1) hi = W - 1;
        if (hi <= prev_hi) return false;
        const int Tg = hi - lo + 1;
        gp.col0[g] = prev_hi + 1; gp.ncol[g] = hi - prev_hi; gp.lo[g] = lo;
        gp.base[g] = base;
        gp.mask4[g] = (uint32_t)(((1ull << (2 * (Tg + 1))) - 1ull) << 2);
        gp.colmask[g] = (uint32_t)(((hi >= 31 ? 0xffffffffull : ((2ull << hi) - 1ull))) & ~((1ull << gp.col0[g]) - 1ull));
        base += 4u << (2 * (Tg + 1));
        prev_hi = hi;
    }
    if (prev_hi != W - 1 || (size_t)base > budget) return false;
    gp.table_bytes = base;
    gp.passmask = (uint32_t)((1ull << W) - 1ull);
    gp.pass_first = 1; gp.pass_last = 1;
    // alignment of the window word (32 bases from p-kd, p = the FIRST window of the pair): kd >= -lo[0] (oldest base read),
    // base p+W (last base of the second window) inside the word: kd <= 31-(W+1); one-shift extraction needs kd >= 15-(hi[0]+1)
    const int hi0 = gp.col0[0] + gp.ncol[0] - 1;
    const int kd_min = -gp.lo[0], kd_max = 31 - (W + 1);
    if (kd_min > kd_max) return false;
    const int kd_fast = std::max(kd_min, 15 - (hi0 + 1));
    fast = kd_fast <= kd_max;
    gp.kd = fast ? kd_fast : kd_min;
    for (int g = 0; g < G; g++) {
        const int hi = gp.col0[g] + gp.ncol[g] - 1;
        const int sh = 60 - 2 * (hi + 1 + gp.kd);
        gp.shift[g] = (uint32_t)sh;
        gp.shift2[g] = sh > 32 ? (uint32_t)(sh - 32) : 0u;
    }
    return true;
}

// The plan as plain numbers (no device work): what bamm_em_create / bamm_em_set_model would choose for these parameters.
extern "C" int bamm_plan_describe(int W, int K, int K_bg_model, int reduced, uint64_t table_budget_bytes, int32_t* out, uint64_t cap,
                                  uint64_t* n_used) {
    REQUIRE(out && n_used, "NULL argument");
    REQUIRE(W >= 1 && W <= 32, "motif width W=%d not in [1,32]", W);
    REQUIRE(K >= 0 && K <= 10 && K_bg_model >= 0 && K_bg_model <= 10, "order out of range");
    std::vector<GroupPlan> plans; std::vector<char> fast;
    *n_used = 0;
    const int K_bg = K_bg_model < K ? K_bg_model : K;          // as in bamm_em_create (reference EM.cpp:23)
    if (!plan_passes(W, K, K_bg, reduced != 0, (size_t)table_budget_bytes, plans, fast)) { REQUIRE(cap >= 1, "buffer too small"); out[0] = 0; *n_used = 1; return BAMM_OK; }
    uint64_t n = 1;
    for (const GroupPlan& gp : plans) n += 8 + 8 * (uint64_t)gp.G;
    REQUIRE(cap >= n, "buffer too small: %llu words needed", (unsigned long long)n);
    int32_t* o = out;
    *o++ = (int32_t)plans.size();
    for (size_t i = 0; i < plans.size(); i++) {
        const GroupPlan& gp = plans[i];
        int ca = 0; while (ca < 32 && !((gp.passmask >> ca) & 1u)) ca++;
        int cb = 32; while (cb > 0 && !((gp.passmask >> (cb - 1)) & 1u)) cb--;
        *o++ = gp.G; *o++ = gp.kd; *o++ = fast[i] ? 1 : 0; *o++ = (int32_t)gp.table_bytes; *o++ = ca; *o++ = cb;
        *o++ = (int32_t)gp.pass_first; *o++ = (int32_t)gp.pass_last;
        for (int g = 0; g < gp.G; g++) {
            *o++ = (int32_t)gp.col0[g]; *o++ = (int32_t)gp.ncol[g]; *o++ = (int32_t)gp.lo[g]; *o++ = (int32_t)gp.shift[g];
            *o++ = (int32_t)gp.shift2[g]; *o++ = (int32_t)gp.mask4[g]; *o++ = (int32_t)gp.base[g]; *o++ = (int32_t)gp.colmask[g];
        }
    }
    *n_used = n;
    return BAMM_OK;
}

// The bound plan of the pruned E-step as plain numbers (no device work).
extern "C" int bamm_bound_plan_describe(int W, int K, int K_bg_model, uint64_t table_budget_bytes, int32_t* out, uint64_t cap, uint64_t* n_used) {
    REQUIRE(out && n_used, "NULL argument");
    REQUIRE(W >= 1 && W <= 32, "motif width W=%d not in [1,32]", W);
    REQUIRE(K >= 0 && K <= 10 && K_bg_model >= 0 && K_bg_model <= 10, "order out of range");
    GroupPlan gp; bool fast = false;
    *n_used = 0;
    REQUIRE(cap >= 1, "buffer too small");
    if (!make_bound_plan(W, K, K_bg_model < K ? K_bg_model : K, (size_t)table_budget_bytes, gp, fast)) { out[0] = 0; *n_used = 1; return BAMM_OK; }
    const uint64_t n = 4 + 7 * (uint64_t)gp.G;
    REQUIRE(cap >= n, "buffer too small: %llu words needed", (unsigned long long)n);
    int32_t* o = out;
    *o++ = gp.G; *o++ = gp.kd; *o++ = fast ? 1 : 0; *o++ = (int32_t)gp.table_bytes;
    for (int g = 0; g < gp.G; g++) {
        *o++ = gp.col0[g]; *o++ = gp.ncol[g]; *o++ = gp.lo[g]; *o++ = (int32_t)gp.shift[g]; *o++ = (int32_t)gp.shift2[g];
        *o++ = (int32_t)gp.mask4[g]; *o++ = (int32_t)gp.base[g];
    }
    *n_used = n;
    return BAMM_OK;
}

extern "C" int bamm_em_create(bamm_seqset* s, const uint64_t* subset, uint64_t nsub, int W, int K, int K_bg_model,
                              bamm_em** out) {
    REQUIRE(out, "out is NULL");
    *out = nullptr;
    REQUIRE(s, "seqset is NULL");
    REQUIRE(W >= 1 && W <= 32, "motif width W=%d not in [1,32]", W);
    REQUIRE(K >= 0 && K <= 10 && K_bg_model >= 0 && K_bg_model <= 10, "order out of range");
    Trace tr("em_create");
    if (!subset) nsub = s->nseq;
    if (group_wanted(s, nsub)) return group_create(s, subset, nsub, W, K, K_bg_model, out);
    REQUIRE(nsub < (1ull << 32), "subset too large");
    const uint64_t Yn64 = ipow_u64((uint64_t)s->A, K + 1);
    REQUIRE(Yn64 * (uint64_t)W < (1ull << 31), "table too large");
    bamm_em* em = new (std::nothrow) bamm_em();
    if (!em) return fail(BAMM_E_NOMEM, "host allocation failed");
    em->ss = s; em->device = s->device; em->W = W; em->K = K; em->K_bg_model = K_bg_model;
    em->K_bg = K_bg_model < K ? K_bg_model : K; em->A = s->A;
    em->Yn = (uint32_t)Yn64; em->nbin = em->Yn * (uint32_t)W; em->nsub = nsub; em->nseq_global = nsub;
    fill_dims(em->dims, s->A, K, W, em->K_bg);
    em->model_size = em->dims.voff[K + 1];
    em->bg_size = em->dims.bgoff[K_bg_model + 1];
    int max_optin = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, em->device);
    const int sms = s->sm_count;
    const size_t table_bytes = (size_t)em->nbin * sizeof(float);
    // ---- packed path: possible when the column-group tables and the M-step's count table fit shared memory
    // (the M-step needs two 32-bit count tables of at least one column: 4^(K+1) * 8 bytes)
    bool packed_ok = s->A == 4 && s->nregular > 0 && Yn64 <= 65536 &&
                     Yn64 * 8 <= (uint64_t)max_optin && !getenv("BAMM_NO_PACKED");
    em->tab_capacity = (size_t)max_optin;
    if (getenv("BAMM_TABLE_BYTES")) em->tab_capacity = std::min(em->tab_capacity, (size_t)atol(getenv("BAMM_TABLE_BYTES")) & ~(size_t)255);   // the tables of a pass start 16-byte aligned (bulk copies)
    if (packed_ok) {
        // both variants (with / without the reduced context of the leading columns) must be plannable
        std::vector<GroupPlan> tmp; std::vector<char> f;
        packed_ok = plan_passes(W, K, em->K_bg, false, em->tab_capacity, tmp, f);
        em->plan.W = W; em->plan.K = K; em->plan.T = 1; em->plan.C = W;
        em->plan.Yn = em->Yn; em->plan.Zn = em->Yn; em->plan.q = 0.3f;
    }
    // ---- split the subset
    std::vector<uint32_t> ids, gen_ids, pk_ids;
    std::vector<uint64_t> gen_roff, pk_roff;
    uint64_t max_lw1_pk = 0;
    if (!subset && packed_ok && s->nregular == s->nseq && s->minL >= (uint64_t)W && nsub > 0) {
        // the whole set, every sequence regular and long enough: the lists are the identity and the set's own offsets
        // (prefix sums of L) — made on the device below; no host pass over the sequences
        em->whole_set = true;
        max_lw1_pk = s->maxL - (uint64_t)W + 1;
    } else {
    em->h_r_off.resize(nsub + 1);
    ids.resize(nsub);
    pk_ids.reserve(nsub); pk_roff.reserve(nsub);
    em->h_r_off[0] = 0;
    for (uint64_t i = 0; i < nsub; i++) {
        const uint64_t n = subset ? subset[i] : i;
        if (n >= s->nseq) { delete em; return fail(BAMM_E_INVALID, "subset[%llu]=%llu out of range", (unsigned long long)i, (unsigned long long)n); }
        const uint64_t L = s->h_off[n + 1] - s->h_off[n];
        if (L < (uint64_t)W) { delete em; return fail(BAMM_E_INVALID, "sequence %llu is shorter (L=%llu) than the motif (W=%d)", (unsigned long long)n, (unsigned long long)L, W); }
        ids[i] = (uint32_t)n;
        em->h_r_off[i + 1] = em->h_r_off[i] + L;
        if (packed_ok && s->h_kind[n]) {
            pk_ids.push_back((uint32_t)n); pk_roff.push_back(em->h_r_off[i]);
            if (L - W + 1 > max_lw1_pk) max_lw1_pk = L - W + 1;
        } else {
            gen_ids.push_back((uint32_t)n); gen_roff.push_back(em->h_r_off[i]);
        }
    }
    }
    em->rsize = em->whole_set ? s->npos : em->h_r_off[nsub];
    em->h_ids.swap(ids);
    em->ngen = (uint32_t)gen_ids.size(); em->npk = em->whole_set ? (uint32_t)nsub : (uint32_t)pk_ids.size();
    tr.mark("subset split (host)");
    IndexArray* ia = nullptr;
    if (em->ngen) { std::lock_guard<std::mutex> g(s->mu); int rc = seqset_index_locked(s, K, &ia); if (rc) { delete em; return rc; } }
    if (em->npk)  { std::lock_guard<std::mutex> g(s->mu); int rc = seqset_ypatch_locked(s, K, &em->d_ypatch); if (rc) { delete em; return rc; } }
#define CUE(call) do { cudaError_t e2_ = (call); if (e2_ != cudaSuccess) { int code_ = e2_ == cudaErrorMemoryAllocation ? BAMM_E_NOMEM : BAMM_E_CUDA; \
    fail(code_, "%s failed: %s", #call, cudaGetErrorString(e2_)); bamm_em_destroy(em); return code_; } } while (0)
    CUE(cudaSetDevice(em->device));
    CUE(cudaStreamCreateWithFlags(&em->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; i++) CUE(cudaEventCreate(&em->ev[i]));
    auto upload = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
        cudaError_t e = dev_malloc(dst, bytes ? bytes : 16);
        if (e == cudaSuccess && bytes) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
        return e;
    };
    CUE(upload(gen_ids.data(), gen_ids.size() * 4, (void**)&em->d_gen_ids));
    CUE(upload(gen_roff.data(), gen_roff.size() * 8, (void**)&em->d_gen_roff));
    if (em->whole_set) {
        CUE(dev_malloc(&em->d_pk_ids, nsub * 4));
        CUE(dev_malloc(&em->d_pk_roff, nsub * 8));
        k_iota_u32<<<(unsigned)((nsub + 255) / 256), 256>>>(em->d_pk_ids, nsub);
        CUE(cudaGetLastError());
        CUE(cudaMemcpy(em->d_pk_roff, s->d_off, nsub * 8, cudaMemcpyDeviceToDevice));
    } else {
        CUE(upload(pk_ids.data(), pk_ids.size() * 4, (void**)&em->d_pk_ids));
        CUE(upload(pk_roff.data(), pk_roff.size() * 8, (void**)&em->d_pk_roff));
    }
    tr.mark("index / ypatch + id uploads");
    CUE(dev_malloc(&em->d_r, (em->rsize ? em->rsize : 1) * sizeof(float)));
    CUE(cudaMemset(em->d_r, 0, (em->rsize ? em->rsize : 1) * sizeof(float)));   // the packed E-step never touches the tail i >= LW1
    tr.mark("r alloc + memset");
    CUE(dev_malloc(&em->d_s, (uint64_t)em->nbin * sizeof(float)));
    CUE(dev_malloc(&em->d_sT, (uint64_t)em->nbin * sizeof(float)));
    CUE(dev_malloc(&em->d_s_alt, (uint64_t)em->nbin * sizeof(float)));
    CUE(dev_malloc(&em->d_sT_alt, (uint64_t)em->nbin * sizeof(float)));
    CUE(dev_malloc(&em->d_v, em->model_size * sizeof(float)));
    CUE(dev_malloc(&em->d_vK_prev, (uint64_t)em->nbin * sizeof(float)));
    CUE(dev_malloc(&em->d_n, em->model_size * sizeof(float)));
    CUE(dev_malloc(&em->d_vbg, em->bg_size * sizeof(float)));
    CUE(dev_malloc(&em->d_alpha, (uint64_t)(K + 1) * W * sizeof(float)));
    CUE(dev_malloc(&em->d_xbuf, ((uint64_t)em->nbin + 2) * sizeof(unsigned long long)));
    CUE(cudaMemset(em->d_xbuf, 0, ((uint64_t)em->nbin + 2) * sizeof(unsigned long long)));
    CUE(dev_malloc(&em->d_vdiff, sizeof(float)));
    CUE(dev_malloc(&em->d_vdiff_part, 16 * sizeof(double)));
    tr.mark("model buffers");
    CUE(pinned_scalars_get(&em->h_scal));                                          // one pinned slot: 2 scalars + sum|dv|
    em->h_vdiff = reinterpret_cast<float*>(em->h_scal + 2);
    tr.mark("pinned scalars");
    em->nparts = 1;
    // ---- generic path geometry (only when some sequence needs it): persistent grid of 512-thread CTAs
    if (em->ngen) {
        em->smem_tables = table_bytes <= (size_t)max_optin;
        if (em->smem_tables) {
            em->smem_e = table_bytes; em->smem_m = table_bytes;
            int per_sm = (int)((size_t)(max_optin + 1024) / (table_bytes + 1024));
            if (per_sm < 1) per_sm = 1;
            if (per_sm > 4) per_sm = 4;             // 4 x 512 threads = 2048 = the SM's thread limit
            em->grid_e = em->grid_m = sms * per_sm;
            bool ok = true;
            if (ia->bytes == 2) { ok &= !max_smem_optin(k_estep<uint16_t, true>, table_bytes); ok &= !max_smem_optin(k_mstep<uint16_t, true>, table_bytes); }
            else                { ok &= !max_smem_optin(k_estep<uint32_t, true>, table_bytes); ok &= !max_smem_optin(k_mstep<uint32_t, true>, table_bytes); }
            if (!ok) { fail(BAMM_E_CUDA, "cannot opt in to %zu bytes of shared memory", table_bytes); bamm_em_destroy(em); return BAMM_E_CUDA; }
            em->nparts = (uint32_t)em->grid_m;
        } else {
            em->smem_e = em->smem_m = 0;
            em->grid_e = em->grid_m = sms * 4;
            // global-memory count tables: a few copies (L2-resident) spread the atomics on hot bins
            uint32_t nrep = getenv("BAMM_GEN_REPLICAS") ? (uint32_t)std::max(1, atoi(getenv("BAMM_GEN_REPLICAS"))) : 16u;
            while (nrep > 1 && (uint64_t)nrep * em->nbin * 8 > (128ull << 20)) nrep >>= 1;
            em->gen_nrep = em->nparts = nrep;
            if (W <= 32 && !getenv("BAMM_GEN_NO_ROWS")) CUE(dev_malloc(&em->d_rows, (uint64_t)em->Yn * ((W + 3) & ~3) * sizeof(float)));
            // when a few columns of low words fit shared memory: column ranges x sequence shares (k_mstep_cols)
            const size_t col_bytes = (size_t)em->Yn * 4;
            if (col_bytes <= (size_t)max_optin && !getenv("BAMM_GEN_NO_COLS")) {
                em->gen_nc = (int)std::min<size_t>((size_t)W, (size_t)max_optin / col_bytes);
                em->gen_nsplit = (W + em->gen_nc - 1) / em->gen_nc;
                em->gen_nc = (W + em->gen_nsplit - 1) / em->gen_nsplit;
                const int ngroups = std::max(1, sms / em->gen_nsplit);
                em->grid_m = ngroups * em->gen_nsplit;
                em->smem_m = (size_t)em->gen_nc * col_bytes;
                const bool one = em->gen_nc == 1;
                const bool ok = ia->bytes == 2 ? !(one ? max_smem_optin(k_mstep_cols<uint16_t, true>, em->smem_m) : max_smem_optin(k_mstep_cols<uint16_t, false>, em->smem_m))
                                               : !(one ? max_smem_optin(k_mstep_cols<uint32_t, true>, em->smem_m) : max_smem_optin(k_mstep_cols<uint32_t, false>, em->smem_m));
                if (!ok) { fail(BAMM_E_CUDA, "cannot opt in to %zu bytes of shared memory", em->smem_m); bamm_em_destroy(em); return BAMM_E_CUDA; }
                em->nparts = (uint32_t)ngroups;
            }
        }
    }
    // ---- packed path geometry
    if (em->npk) {
        CUE(dev_malloc(&em->d_scale, (size_t)em->npk * sizeof(float)));
        em->block_pe = BAMM_E_THREADS;              // one CTA per SM: the group tables fill its shared memory
        em->grid_pe = sms;
        // M-step geometry: as many columns per CTA as two 32-bit tables allow, the fewest splits, columns spread evenly
        {
            int nc_max = std::min(32 - K, (int)((size_t)max_optin / ((size_t)em->Yn * 8)));   // K + columns <= 32 bases of one window word
            // without the high-word table in shared memory twice the columns fit: taken when it saves a pass over the windows
            // (orders >= 5 at W = 20: 3 -> 2 column splits; order 6: 20 -> 7)
            const int nc_hg = std::min(32 - K, (int)((size_t)max_optin / ((size_t)em->Yn * 4)));
            em->m_tab.hi_global = 0;
            if (nc_max >= 1 && nc_hg >= 1 && nc_hg <= 14 && !getenv("BAMM_M_NO_HI_GLOBAL") &&       // 14: the instantiated widths (launch_mstep.cu)
                (W + std::min(nc_hg, W) - 1) / std::min(nc_hg, W) < (W + std::min(nc_max, W) - 1) / std::min(nc_max, W)) {
                nc_max = nc_hg; em->m_tab.hi_global = 1;
            }
            if (getenv("BAMM_M_COLS")) nc_max = std::max(1, std::min(nc_max, atoi(getenv("BAMM_M_COLS"))));
            if (nc_max > W) nc_max = W;
            em->m_nsplit = (W + nc_max - 1) / nc_max;
            em->m_nc = (W + em->m_nsplit - 1) / em->m_nsplit;
            em->grid_pl = std::max(1, sms / em->m_nsplit) * em->m_nsplit;
            // table copies for tiny tables (same-address atomics serialise): 32 copies at order 0, 8 at order 1
            {
                const uint32_t nb = (uint32_t)em->m_nc * em->Yn;
                uint32_t nrep = em->Yn <= 4 ? 32u : em->Yn <= 16 ? 8u : 1u;
                if (getenv("BAMM_M_REPLICAS")) nrep = (uint32_t)std::max(1, atoi(getenv("BAMM_M_REPLICAS")));
                while (nrep & (nrep - 1)) nrep &= nrep - 1;                 // power of two
                while (nrep > 1 && (size_t)2 * nrep * (((nb + 30) / 32) * 32 + 1) * 4 > (size_t)max_optin) nrep >>= 1;
                if (em->m_tab.hi_global) nrep = 1;
                em->m_tab.nrep = nrep;
                em->m_tab.rstride = nrep > 1 ? ((nb + 30) / 32) * 32 + 1 : nb;
            }
            // the high table sums at most 257 per sequence and bin (the posteriors of a sequence sum to <= 1)
            if ((uint64_t)em->npk / (uint64_t)(em->grid_pl / em->m_nsplit) >= (1ull << 23)) {
                fail(BAMM_E_INVALID, "too many sequences for one device"); bamm_em_destroy(em); return BAMM_E_INVALID;
            }
            if (mstep_w_dispatch(em, nullptr, nullptr, 0)) { fail(BAMM_E_CUDA, "cannot opt in to shared memory for the packed M-step"); bamm_em_destroy(em); return BAMM_E_CUDA; }
            if (!em->m_tab.hi_global && (uint32_t)em->grid_pl > em->nparts) em->nparts = (uint32_t)em->grid_pl;   // hi_global: one table for all CTAs
        }
        tr.mark("M-step geometry + opt-in");
        // active list: one region per E-step warp, sized as a fraction of the warp's windows (BAMM_LIST_FRAC, 0 = off)
        double frac = getenv("BAMM_LIST_FRAC") ? atof(getenv("BAMM_LIST_FRAC")) : 0.5;
        if (s->nwords >= (1ull << 32)) frac = 0.0;          // list entries address the stream with 32-bit word offsets
        std::vector<uint64_t> reg, win, nsq;
        if (frac > 0.0) {
            em->nregions = (uint32_t)em->grid_pe * (uint32_t)(em->block_pe / 32);
            win.assign(em->nregions, 0); nsq.assign(em->nregions, 0);      // nsq: windows k_emasked may list per region (at most 2W+K+2 per sequence)
            // windows of a sequence k_emasked evaluates (and may list): those over the structural N plus the truncated tail, as the kernels cut them
            auto masked_windows = [&](uint64_t L, bool hasN) -> uint64_t {
                const long long lw1 = (long long)L - W + 1, tl = std::min(std::max((long long)L - 2 * W + 2, 0ll), lw1);
                long long nn = 0;
                if (hasN) { const long long mid = ((long long)L - 1) / 2; nn = std::min(mid + K + 1, tl) - std::min(std::max(mid - W + 1, 0ll), tl); }
                return (uint64_t)(std::max(nn, 0ll) + (lw1 - tl));
            };
            reg.assign((size_t)em->nregions + 1, 0);
            if (em->whole_set && s->minL == s->maxL) {          // equal lengths: sequence i goes to warp i % nregions
                const uint64_t lw1 = s->maxL - (uint64_t)W + 1, per = em->npk / em->nregions, extra = em->npk % em->nregions;
                const uint64_t mper = masked_windows(s->maxL, s->maxL & 1);     // odd length: counted as a both-strand record with its N
                for (uint32_t w = 0; w < em->nregions; w++) { const uint64_t c = per + (w < extra ? 1 : 0); nsq[w] = c * mper; win[w] = c * lw1; }
            } else {
                uint32_t w = 0;
                for (size_t i = 0; i < em->npk; i++) {
                    const uint64_t n = em->whole_set ? i : pk_ids[i];
                    const uint64_t lw1 = s->h_off[n + 1] - s->h_off[n] - (uint64_t)W + 1;
                    win[w] += lw1;
                    nsq[w] += masked_windows(s->h_off[n + 1] - s->h_off[n], s->h_kind[n] == 2);
                    if (++w == em->nregions) w = 0;
                }
            }
            // the list is an accelerator, not a requirement: when memory is short its capacity is halved (down to 1/32 of
            // the windows), below that the M-step scans r
            for (;;) {
                for (uint32_t w = 0; w < em->nregions; w++) {
                    uint64_t cap = (uint64_t)(frac * (double)win[w]) + 256;
                    if (cap > win[w]) cap = win[w];
                    reg[w + 1] = reg[w] + cap;
                }
                const uint64_t total = reg[em->nregions] ? reg[em->nregions] : 1;
                const cudaError_t ea = dev_malloc(&em->d_act, total * sizeof(ActiveEntry));
                if (ea == cudaSuccess) break;
                cudaGetLastError();
                em->d_act = nullptr;
                if (ea != cudaErrorMemoryAllocation) CUE(ea);
                frac *= 0.5;
                if (frac < 1.0 / 32.0) break;
            }
        }
        if (em->d_act) {
            CUE(upload(reg.data(), reg.size() * 8, (void**)&em->d_reg_off));
            CUE(dev_malloc(&em->d_act_cnt, (uint64_t)em->nregions * 8));         // front counts, then back counts
            CUE(cudaMemset(em->d_act_cnt, 0, (uint64_t)em->nregions * 8));
            CUE(dev_malloc(&em->d_overflow, 4));
            CUE(cudaMemset(em->d_overflow, 0, 4));
            // candidate list of the pruned E-step: one region per warp for a fraction of its windows (BAMM_CAND_FRAC). The exact
            // pass lists at most the candidates plus the windows over the N and the truncated tail of every sequence, so a
            // region of the active list that holds all of those cannot overflow there.
            const double cfrac = getenv("BAMM_CAND_FRAC") ? atof(getenv("BAMM_CAND_FRAC")) : 0.42;
            if (cfrac > 0.0 && !getenv("BAMM_NO_SPARSE")) {
                std::vector<uint64_t> creg((size_t)em->nregions + 1, 0);
                bool ok = true;
                for (uint32_t w = 0; w < em->nregions && ok; w++) {
                    const uint64_t masked = nsq[w], rcap = reg[w + 1] - reg[w];
                    uint64_t cap = (uint64_t)(cfrac * (double)win[w]) + 256;
                    if (cap > win[w]) cap = win[w];
                    if (cap + masked > rcap) { if (rcap < masked) ok = false; else cap = rcap - masked; }   // (0: every window of the warp is a masked one)
                    creg[w + 1] = creg[w] + cap;
                }
                em->cand_slots = creg[em->nregions];
                if (ok && dev_malloc(&em->d_cand, (creg[em->nregions] ? creg[em->nregions] : 1) * sizeof(uint32_t)) == cudaSuccess) {
                    CUE(dev_malloc(&em->d_cand_seq, (size_t)em->npk * sizeof(uint2)));
                    CUE(dev_malloc(&em->d_seqacc, (size_t)em->npk * sizeof(ulonglong2)));
                    CUE(upload(creg.data(), creg.size() * 8, (void**)&em->d_creg_off));
                    CUE(dev_malloc(&em->d_eflags, 32));
                    CUE(cudaMemset(em->d_eflags, 0, 32));
                    em->cand_ok = true;
                } else cudaGetLastError();
            }
        }
    }
    tr.mark("active list");
    CUE(dev_malloc(&em->d_part, (uint64_t)em->nparts * em->nbin * sizeof(unsigned long long)));
#undef CUE
    tr.mark("partials");
    *out = em;
    return BAMM_OK;
}

// h_ids / h_r_off of a whole-set object (identity, the set's offsets), made when a host consumer first asks
static void host_lists(bamm_em* em) {
    if (!em->whole_set || !em->h_r_off.empty()) return;
    em->h_ids.resize(em->nsub);
    for (uint64_t i = 0; i < em->nsub; i++) em->h_ids[i] = (uint32_t)i;
    em->h_r_off.assign(em->ss->h_off.begin(), em->ss->h_off.begin() + em->nsub + 1);
}

// group tables of the exact plan(s) from the [j][y] table `d_s` into `d_tab_dst`; bound tables of the pruned path
static int launch_tables(bamm_em* em, const float* d_s, float* d_tab_dst) {
    if (!em->npk) return BAMM_OK;
    for (size_t i = 0; i < em->gplans.size(); i++) {
        const uint32_t total = em->gplans[i].table_bytes >> 2;
        float* dst = (float*)((char*)d_tab_dst + i * em->tab_capacity);
        if (i == 0 && em->sparse && em->K >= 1) {           // + one CTA for the bound levels
            const uint32_t blocks = (total + 1023) / 1024;
            k_em_tables<<<(blocks < 296 ? blocks : 296) + 1, 1024, 0, em->stream>>>(d_s, em->gplans[i], dst, em->blev, em->d_U);
        } else {
            const uint32_t blocks = (total + 255) / 256;
            k_make_group_tables<false><<<blocks < 1184 ? blocks : 1184, 256, 0, em->stream>>>(d_s, em->gplans[i], dst);
        }
        CU(cudaGetLastError());
    }
    if (em->sparse) {
        const uint32_t total = em->bplan.table_bytes >> 2, blocks = (total + 255) / 256;
        k_make_bound_tables<<<blocks < 1184 ? blocks : 1184, 256, 0, em->stream>>>(d_s, em->d_U, em->blev, em->bplan, (uint32_t*)em->d_btab);
        CU(cudaGetLastError());
        em->launches += 1;
    }
    return BAMM_OK;
}

// true when every column j < K of v[K] only depends on the j+1 newest bases (what Motif::updateV produces, Motif.h:126-128)
static bool leading_columns_are_copies(const ModelDims& dims, int K, int W, uint32_t Yn, const float* v_all) {
    const float* vK = v_all + dims.voff[K];
    for (int j = 0; j < K && j < W; j++) {
        const uint32_t period = dims.Y[j + 1];
        for (uint32_t y = period; y < Yn; y++)
            if (vK[(uint64_t)y * W + j] != vK[(uint64_t)(y % period) * W + j]) return false;
    }
    return true;
}

extern "C" int bamm_em_set_model(bamm_em* em, const float* v_all, const float* vbg_all, const float* alpha, float q) {
    REQUIRE(em && v_all && vbg_all && alpha, "NULL argument");
    if (IS_GROUP(em)) return group_set_model(em, v_all, vbg_all, alpha, q);
    REQUIRE(q > 0.0f && q < 1.0f, "q=%g not in (0,1)", (double)q);
    CU(cudaSetDevice(em->device));
    Trace tr("set_model");
    if (em->npk) {
        const bool reduced = leading_columns_are_copies(em->dims, em->K, em->W, em->Yn, v_all) && !getenv("BAMM_NO_REDUCED");
        if (!plan_passes(em->W, em->K, em->K_bg, reduced, em->tab_capacity, em->gplans, em->gfast))
            return fail(BAMM_E_STATE, "no column-group plan fits shared memory");
        // a single-pass plan that leaves room for the plain [j][y] table behind the group tables (the single columns of masked
        // windows then come from shared memory) is preferred when it costs at most one more group
        em->plain_words = 0;
        const size_t plain_bytes = ((size_t)em->nbin + (size_t)em->W) * sizeof(float);      // rows padded by one float in shared memory
        if (em->gplans.size() == 1 && plain_bytes + 4096 <= em->tab_capacity && !getenv("BAMM_NO_PLAIN_SMEM")) {
            std::vector<GroupPlan> pp; std::vector<char> pf;
            if (plan_passes(em->W, em->K, em->K_bg, reduced, em->tab_capacity - plain_bytes, pp, pf) && pp.size() == 1 && pp[0].G <= em->gplans[0].G + 1) {
                em->gplans.swap(pp); em->gfast.swap(pf); em->plain_words = em->nbin;
            }
        }
        // pruned path: worth it when the bound needs at least three lookups fewer than the exact product (BAMM_SPARSE=1: whenever fewer)
        em->sparse = false;
        const bool multi_pass = em->gplans.size() > 1;       // (without the plain table in shared memory the single columns come from global memory)
        if (em->cand_ok && !getenv("BAMM_NO_SPARSE") &&
            make_bound_plan(em->W, em->K, em->K_bg, em->tab_capacity, em->bplan, em->bfast)) {
            // the bound pass costs G1/2 lookups per window (two windows per entry); with column passes the exact product costs
            // the groups of all passes plus a round trip of the partial product per extra pass
            int Gall = 0;
            for (const GroupPlan& g : em->gplans) Gall += g.G;
            // measured (profiles/r2j_sparse_threshold.txt): worth it as soon as the bound is cheaper than the product at all — K=1 W=20
            // (G = 4 against 2 lookups per window): 1.40 -> 0.89 ms per 300k sequences. BAMM_SPARSE_MARGIN asks for more.
            const int need = getenv("BAMM_SPARSE_MARGIN") ? atoi(getenv("BAMM_SPARSE_MARGIN")) : 1;
            em->sparse = (em->bplan.G + 1) / 2 + need <= Gall;
        }
        if (em->gplans.size() > em->tab_passes) {
            CU(cudaStreamSynchronize(em->stream));
            cudaFree(em->d_tab); cudaFree(em->d_tab_alt); em->d_tab = em->d_tab_alt = nullptr; em->tab_passes = 0;
            CU(dev_malloc(&em->d_tab, em->gplans.size() * em->tab_capacity));
            CU(dev_malloc(&em->d_tab_alt, em->gplans.size() * em->tab_capacity));
            em->tab_passes = em->gplans.size();
        }
        const EStepLaunch l = {em->grid_pe, em->block_pe, em->stream};
        for (size_t i = 0; i < em->gplans.size(); i++)
            if (launch_estep_dense(l, true, em->gfast[i] != 0, em->gplans.size() > 1, nullptr, em->gplans[i], nullptr, nullptr, em->plain_words, nullptr, nullptr, nullptr, nullptr))
                return fail(BAMM_E_CUDA, "cannot opt in to %u bytes of shared memory", em->gplans[i].table_bytes);
        if (em->sparse) {
            if (!em->d_btab) {
                CU(dev_malloc(&em->d_btab, em->tab_capacity));
                uint32_t off = 0;
                memset(&em->blev, 0, sizeof(em->blev));
                for (int a = 1; a <= em->K && a < 12; a++) { em->blev.off[a] = off; off += (uint32_t)em->W << (2 * a); }
                CU(dev_malloc(&em->d_U, (size_t)(off ? off : 1) * sizeof(float)));
            }
            // per-warp staging of the sequence words in the exact pass, when shared memory has room left in every pass
            em->stage = !getenv("BAMM_NO_STAGE");
            for (const GroupPlan& g : em->gplans)
                if ((size_t)g.table_bytes + (em->plain_words ? plain_bytes : 0) + estep_stage_bytes(em->block_pe) > em->tab_capacity) em->stage = false;
            if (launch_estep_bound(l, true, em->bfast, nullptr, em->bplan, nullptr, nullptr)) return fail(BAMM_E_CUDA, "cannot opt in to shared memory for the pruned E-step");
            for (size_t i = 0; i < em->gplans.size(); i++)
                if (launch_estep_masked(l, true, em->gfast[i] != 0, nullptr, em->gplans[i], nullptr, nullptr, em->plain_words, nullptr, nullptr, nullptr, nullptr) ||
                    launch_estep_exact(l, true, em->gfast[i] != 0, nullptr, em->gplans[i], nullptr, nullptr, em->plain_words, em->stage, nullptr, nullptr, nullptr, nullptr, nullptr))
                    return fail(BAMM_E_CUDA, "cannot opt in to shared memory for the pruned E-step");
            if (multi_pass && !em->d_cand_part) {      // partial products between the column passes: per candidate slot, per masked window
                CU(dev_malloc(&em->d_cand_part, (size_t)(em->cand_slots ? em->cand_slots : 1) * sizeof(float)));
                CU(dev_malloc(&em->d_mask_part, (size_t)em->npk * (size_t)(2 * em->W + em->K - 1) * sizeof(float)));
            }
        }
    }
    tr.mark("plan + table buffer + opt-in");
    CU(cudaMemcpyAsync(em->d_v, v_all, em->model_size * sizeof(float), cudaMemcpyHostToDevice, em->stream));
    CU(cudaMemcpyAsync(em->d_vbg, vbg_all, em->bg_size * sizeof(float), cudaMemcpyHostToDevice, em->stream));
    CU(cudaMemcpyAsync(em->d_alpha, alpha, (uint64_t)(em->K + 1) * em->W * sizeof(float), cudaMemcpyHostToDevice, em->stream));
    k_make_s<<<64, 256, 0, em->stream>>>(em->dims, em->d_v, em->d_vbg, em->d_s, em->d_sT, em->d_vK_prev);
    CU(cudaGetLastError());
    { int rc = launch_tables(em, em->d_s, em->d_tab); if (rc) return rc; }
    CU(cudaStreamSynchronize(em->stream));
    tr.mark("uploads + s + group tables");
    em->q = q; em->model_set = true; em->s_valid = true; em->r_valid = false; em->llh = 0.0f;
    return BAMM_OK;
}

static ActiveList alist_of(const bamm_em* em) {
    ActiveList al; al.ent = em->d_act; al.scale = em->d_scale; al.reg_off = em->d_reg_off;
    al.cnt = em->d_act_cnt; al.cnt_back = em->d_act_cnt ? em->d_act_cnt + em->nregions : nullptr; al.overflow = em->d_overflow;
    return al;
}
static CandList clist_of(const bamm_em* em) {
    CandList cl; cl.ent = em->d_cand; cl.reg_off = em->d_creg_off; cl.seq = em->d_cand_seq; cl.flags = em->d_eflags;
    return cl;
}
// plan of pass `pass` with the launch-time parameters filled in
static GroupPlan plan_for_launch(const bamm_em* em, size_t pass, float q) {
    GroupPlan gp = em->gplans[pass]; gp.q = q;
    gp.thr0 = FX_HALF_UNIT * (1.0f - q) * 0.999f;
    return gp;
}

static SubsetView view_of(const bamm_em* em) {
    SubsetView sv; sv.seq_off = em->ss->d_off; sv.seq_ids = em->d_gen_ids; sv.r_off = em->d_gen_roff; sv.nsub = em->ngen;
    return sv;
}
static PackedView pview_of(const bamm_em* em) {
    PackedView pv; pv.words = em->ss->d_words; pv.seqs = em->ss->d_pseq; pv.ypatch = em->d_ypatch;
    pv.seq_ids = em->d_pk_ids; pv.r_off = em->d_pk_roff; pv.nlist = em->npk;
    return pv;
}

// row-form E-step of the index-array path (kernels.cuh, k_estep_rows): pads the [y][j] table to rows of W4 floats, then one
// instantiation per padded width
template <typename YT>
static void estep_rows_dispatch(bamm_em* em, const YT* Y, const SubsetView& sv, unsigned long long* scal) {
    const int W4 = (em->W + 3) & ~3;
    const uint32_t total = em->Yn * (uint32_t)W4, blocks = (total + 255) / 256;
    k_pad_rows<<<blocks < 1184 ? blocks : 1184, 256, 0, em->stream>>>(em->d_sT, em->W, W4, em->Yn, em->d_rows);
    em->launches += 1;
#define BAMM_ROWS_CASE(N) case N: k_estep_rows<YT, N><<<em->grid_e, em->block, 0, em->stream>>>(Y, sv, em->W, em->d_rows, em->q, em->d_r, scal); break;
    switch (W4) {
        BAMM_ROWS_CASE(4) BAMM_ROWS_CASE(8) BAMM_ROWS_CASE(12) BAMM_ROWS_CASE(16) BAMM_ROWS_CASE(20) BAMM_ROWS_CASE(24) BAMM_ROWS_CASE(28) BAMM_ROWS_CASE(32)
        default: break;
    }
#undef BAMM_ROWS_CASE
}

static int launch_estep(bamm_em* em, cudaEvent_t* split = nullptr /* 2 events: after the masked windows, after the bounds */) {
    em->launches += 1 + (em->npk ? em->gplans.size() + (em->sparse ? 3 : 0) : 0) + (em->ngen ? 1 : 0);
    unsigned long long* scal = em->d_xbuf + em->nbin;
    k_estep_begin<<<1, 32, 0, em->stream>>>(scal, em->d_overflow, em->d_eflags);
    CU(cudaGetLastError());
    if (em->npk) {
        PackedView pv = pview_of(em);
        const EStepLaunch l = {em->grid_pe, em->block_pe, em->stream};
        const ActiveList al = alist_of(em);
        const bool multi = em->gplans.size() > 1;
        if (em->sparse) {
            // bounds -> exact evaluation of the candidates; the dense kernel only runs when the candidate list overflowed
            const CandList cl = clist_of(em);
            GroupPlan bp = em->bplan; bp.q = em->q;
            bp.thr0 = FX_HALF_UNIT * (1.0f - em->q) * 0.999f * 0.9999f;          // margin: bound and product round differently
            const size_t np = em->gplans.size();
            auto tab_of = [&](size_t pass) { return (const float*)((const char*)em->d_tab + pass * em->tab_capacity); };
            for (size_t pass = 0; pass < np; pass++)
                if (launch_estep_masked(l, false, em->gfast[pass] != 0, &pv, plan_for_launch(em, pass, em->q), tab_of(pass), em->d_s, em->plain_words, &cl,
                                        em->d_seqacc, np > 1 ? em->d_mask_part : nullptr, &al))
                    return fail(BAMM_E_CUDA, "E-step launch failed (masked windows)");
            if (split) CU(cudaEventRecord(split[0], em->stream));
            if (launch_estep_bound(l, false, em->bfast, &pv, bp, em->d_btab, &cl)) return fail(BAMM_E_CUDA, "E-step launch failed (bounds)");
            if (split) CU(cudaEventRecord(split[1], em->stream));
            for (size_t pass = 0; pass < np; pass++)
                if (launch_estep_exact(l, false, em->gfast[pass] != 0, &pv, plan_for_launch(em, pass, em->q), tab_of(pass), em->d_s, em->plain_words, em->stage, &cl,
                                       em->d_seqacc, np > 1 ? em->d_cand_part : nullptr, scal, &al))
                    return fail(BAMM_E_CUDA, "E-step launch failed (candidates)");
            for (size_t pass = 0; pass < np; pass++)
                if (launch_estep_dense(l, false, em->gfast[pass] != 0, np > 1, &pv, plan_for_launch(em, pass, em->q), tab_of(pass), em->d_s, em->plain_words, em->d_r, scal,
                                       &al, em->d_eflags))
                    return fail(BAMM_E_CUDA, "packed E-step launch failed");
            em->launches += 3 * (np - 1);
        } else {
            if (split) { CU(cudaEventRecord(split[0], em->stream)); CU(cudaEventRecord(split[1], em->stream)); }
            for (size_t pass = 0; pass < em->gplans.size(); pass++)
                if (launch_estep_dense(l, false, em->gfast[pass] != 0, multi, &pv, plan_for_launch(em, pass, em->q), (const float*)((const char*)em->d_tab + pass * em->tab_capacity),
                                       em->d_s, em->plain_words, em->d_r, scal, &al, nullptr))
                    return fail(BAMM_E_CUDA, "packed E-step launch failed");
        }
        em->r_scaled = false;
        em->r_mat = !em->sparse;
        CU(cudaGetLastError());
    }
    em->d_s_e = em->d_s; em->d_sT_e = em->d_sT; em->d_tab_e = em->d_tab; em->q_e = em->q;
    em->peer_sums_global = false;
    if (em->ngen) {
        IndexArray& ia = em->ss->index[em->K];
        SubsetView sv = view_of(em);
        if (em->d_rows) {                       // tables beyond shared memory: whole rows per position, columns by shuffle
            if (ia.bytes == 2) estep_rows_dispatch(em, (const uint16_t*)ia.d, sv, scal);
            else               estep_rows_dispatch(em, (const uint32_t*)ia.d, sv, scal);
        } else if (ia.bytes == 2) {
            if (em->smem_tables) k_estep<uint16_t, true><<<em->grid_e, em->block, em->smem_e, em->stream>>>((const uint16_t*)ia.d, sv, em->W, em->Yn, em->d_s, em->q, em->d_r, scal);
            else                 k_estep<uint16_t, false><<<em->grid_e, em->block, 0, em->stream>>>((const uint16_t*)ia.d, sv, em->W, em->Yn, em->d_sT, em->q, em->d_r, scal);
        } else {
            if (em->smem_tables) k_estep<uint32_t, true><<<em->grid_e, em->block, em->smem_e, em->stream>>>((const uint32_t*)ia.d, sv, em->W, em->Yn, em->d_s, em->q, em->d_r, scal);
            else                 k_estep<uint32_t, false><<<em->grid_e, em->block, 0, em->stream>>>((const uint32_t*)ia.d, sv, em->W, em->Yn, em->d_sT, em->q, em->d_r, scal);
        }
        CU(cudaGetLastError());
    }
    return BAMM_OK;
}

// packed M-step kernels (launch_mstep.cu). mode 0: opt in to the shared memory of both kernels, 1: launch the list kernel,
// 2: launch the scan kernel (conditional on the list's overflow flag when there is a list)
static int mstep_w_dispatch(bamm_em* em, const PackedView* pv, const Plan* pl, int mode) {
    const size_t smem = (size_t)(em->m_tab.hi_global ? 1 : 2) * em->m_tab.nrep * em->m_tab.rstride * 4;
    const ActiveList al = alist_of(em);
    return launch_mstep_packed(em->m_nc, mode, em->grid_pl, smem, em->stream, pv, pl, &al, em->nregions, em->m_nsplit, em->m_tab, em->d_part,
                               em->d_r, em->r_scaled ? nullptr : em->d_scale, em->d_act ? em->d_overflow : nullptr);
}

static int launch_mstep_accumulate(bamm_em* em) {
    em->launches += (em->npk ? (em->d_act ? 2 : 1) : 0) + (em->ngen ? 1 : 0);
    CU(cudaMemsetAsync(em->d_part, 0, (uint64_t)em->nparts * em->nbin * sizeof(unsigned long long), em->stream));
    if (em->npk) {
        PackedView pv = pview_of(em);
        Plan pl = em->plan; pl.q = em->q;
        if (em->d_act && getenv("BAMM_DEBUG_LIST")) {
            std::vector<uint32_t> c(2 * (size_t)em->nregions); uint32_t ov = 0;
            cudaStreamSynchronize(em->stream);
            cudaMemcpy(c.data(), em->d_act_cnt, (size_t)em->nregions * 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(&ov, em->d_overflow, 4, cudaMemcpyDeviceToHost);
            uint64_t tot = 0, mx = 0; for (uint32_t x : c) { tot += x; if (x > mx) mx = x; }
            fprintf(stderr, "[bamm] active list: %llu entries (%.4f of r), max region %llu, overflow %u, G=%d fast=%d delta=%d table %u B\n",
                    (unsigned long long)tot, (double)tot / (double)em->rsize, (unsigned long long)mx, ov, em->gplans[0].G, (int)em->gfast[0], em->gplans[0].kd, em->gplans[0].table_bytes);
            if (em->sparse) {
                uint32_t f[4] = {0, 0, 0, 0};
                cudaMemcpy(f, em->d_eflags, 16, cudaMemcpyDeviceToHost);
                fprintf(stderr, "[bamm] pruned E-step: G1=%d fast=%d, %llu candidates (%.4f of r), dense fall-back %u, hold %u\n", em->bplan.G, (int)em->bfast,
                        (unsigned long long)f[2] | ((unsigned long long)f[3] << 32), (double)((unsigned long long)f[2] | ((unsigned long long)f[3] << 32)) / (double)em->rsize, f[0], f[1]);
            }
        }
        // the E-step listed the windows that matter; the scan kernel only does work (device-side decision) if a region
        // overflowed, or when there is no list
        if (em->d_act && mstep_w_dispatch(em, &pv, &pl, 1)) return fail(BAMM_E_CUDA, "list M-step launch failed");
        if (mstep_w_dispatch(em, &pv, &pl, 2)) return fail(BAMM_E_CUDA, "scan M-step launch failed");
    }
    if (em->ngen) {
        IndexArray& ia = em->ss->index[em->K];
        SubsetView sv = view_of(em);
        if (em->gen_nsplit) {
#define BAMM_MCOLS(YT, ONE) k_mstep_cols<YT, ONE><<<em->grid_m, 1024, em->smem_m, em->stream>>>((const YT*)ia.d, sv, em->W, em->Yn, em->d_r, em->d_part, em->gen_nsplit, em->gen_nc)
            if (ia.bytes == 2) { if (em->gen_nc == 1) BAMM_MCOLS(uint16_t, true); else BAMM_MCOLS(uint16_t, false); }
            else               { if (em->gen_nc == 1) BAMM_MCOLS(uint32_t, true); else BAMM_MCOLS(uint32_t, false); }
#undef BAMM_MCOLS
        } else if (ia.bytes == 2) {
            if (em->smem_tables) k_mstep<uint16_t, true><<<em->grid_m, em->block, em->smem_m, em->stream>>>((const uint16_t*)ia.d, sv, em->W, em->Yn, em->d_r, em->d_part, 1u);
            else                 k_mstep<uint16_t, false><<<em->grid_m, em->block, 0, em->stream>>>((const uint16_t*)ia.d, sv, em->W, em->Yn, em->d_r, em->d_part, em->gen_nrep);
        } else {
            if (em->smem_tables) k_mstep<uint32_t, true><<<em->grid_m, em->block, em->smem_m, em->stream>>>((const uint32_t*)ia.d, sv, em->W, em->Yn, em->d_r, em->d_part, 1u);
            else                 k_mstep<uint32_t, false><<<em->grid_m, em->block, 0, em->stream>>>((const uint32_t*)ia.d, sv, em->W, em->Yn, em->d_r, em->d_part, em->gen_nrep);
        }
        CU(cudaGetLastError());
    }
    return BAMM_OK;
}

static int launch_mstep_reduce(bamm_em* em) {
    em->launches += 1;
    if (em->peer_attached) {
        // fused reduce + NVLink push to every rank, then wait-and-sum on this rank (no collective call, no host)
        em->peer_epoch++;
        const uint32_t parity = em->peer_epoch & 1u, words = em->nbin + 2;
        k_reduce_push<<<reduce_grid(words), RP_THREADS, 0, em->stream>>>(em->d_part, em->nparts, em->nbin, em->d_xbuf + em->nbin, em->peer_ptrs,
                                                                   em->peer_rank, em->peer_world, parity, em->peer_epoch, em->d_peer_done);
        CU(cudaGetLastError());
        const size_t slot_bytes = (size_t)2 * em->peer_world * words * sizeof(unsigned long long);
        k_peer_sum<<<(words + 255) / 256, 256, 0, em->stream>>>((const unsigned long long*)em->d_peer_local, (const unsigned int*)(em->d_peer_local + slot_bytes),
                                                                em->peer_world, em->nbin, parity, em->peer_epoch, em->d_xbuf, em->d_peer_wait);
        CU(cudaGetLastError());
        em->launches += 1;
        em->peer_sums_global = true;
        return BAMM_OK;
    }
    k_reduce_parts<<<reduce_grid(em->nbin), RP_THREADS, 0, em->stream>>>(em->d_part, em->nparts, em->nbin, em->d_xbuf);
    CU(cudaGetLastError());
    return BAMM_OK;
}

static int launch_mstep_local(bamm_em* em) {
    int rc = launch_mstep_accumulate(em); if (rc) return rc;
    return launch_mstep_reduce(em);
}

static int launch_update(bamm_em* em) {
    em->launches += 1 + (em->npk ? 1 : 0);
    // the next E-step's tables go to the second buffers, which then become the current ones: the tables of the last E-step stay
    // readable (bamm_em_get_r, bamm_em_get_s)
    // large tables: one thread-block cluster of 8 CTAs instead of one CTA
    if ((uint64_t)em->nbin >= 16384 && em->d_vdiff_part && !getenv("BAMM_NO_CLUSTER_UPDATE"))
        k_update_model_cluster<<<UPDATE_CLUSTER, 1024, 0, em->stream>>>(em->dims, em->d_xbuf, em->d_n, em->d_v, em->d_vK_prev, em->d_vbg, em->d_alpha,
                                                                      em->d_s_alt, em->d_sT_alt, em->d_vdiff, em->d_vdiff_part);
    else
        k_update_model<<<1, 1024, 0, em->stream>>>(em->dims, em->d_xbuf, em->d_n, em->d_v, em->d_vK_prev, em->d_vbg, em->d_alpha, em->d_s_alt, em->d_sT_alt, em->d_vdiff);
    CU(cudaGetLastError());
    std::swap(em->d_s, em->d_s_alt); std::swap(em->d_sT, em->d_sT_alt);
    if (!em->npk) return BAMM_OK;
    const int rc = launch_tables(em, em->d_s, em->d_tab_alt);
    std::swap(em->d_tab, em->d_tab_alt);
    return rc;
}

static int read_scalars(bamm_em* em, bool want_vdiff) {
    CU(cudaMemcpyAsync(em->h_scal, em->d_xbuf + em->nbin, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, em->stream));
    if (want_vdiff) CU(cudaMemcpyAsync(em->h_vdiff, em->d_vdiff, sizeof(float), cudaMemcpyDeviceToHost, em->stream));
    CU(cudaStreamSynchronize(em->stream));
    em->llh = (float)((double)(long long)em->h_scal[0] * SC_INV_D);
    return BAMM_OK;
}

static float q_from_rsum(const bamm_em* em) {     // reference: EM.cpp:515
    const float N1 = (float)((double)(long long)em->h_scal[1] * SC_INV_D);
    return ((float)em->nseq_global - N1 + 1.f) / ((float)em->nseq_global + 2.f);
}

extern "C" int bamm_em_estep_local(bamm_em* em) {
    REQUIRE(em, "em is NULL");
    NOT_FOR_GROUPS(em, "bamm_em_estep_local");
    if (!em->model_set) return fail(BAMM_E_STATE, "bamm_em_set_model has not been called");
    CU(cudaSetDevice(em->device));
    CU(cudaEventRecord(em->ev[0], em->stream));
    int rc = launch_estep(em); if (rc) return rc;
    CU(cudaEventRecord(em->ev[1], em->stream));
    em->r_valid = true;
    return BAMM_OK;
}

extern "C" int bamm_em_estep(bamm_em* em, float* llh) {
    if (IS_GROUP(em)) return group_estep(em, llh);
    int rc = bamm_em_estep_local(em); if (rc) return rc;
    rc = read_scalars(em, false); if (rc) return rc;
    if (llh) *llh = em->llh;
    return BAMM_OK;
}

extern "C" int bamm_em_mstep_local(bamm_em* em) {
    REQUIRE(em, "em is NULL");
    NOT_FOR_GROUPS(em, "bamm_em_mstep_local");
    if (!em->r_valid) return fail(BAMM_E_STATE, "M-step needs the r of an E-step");
    CU(cudaSetDevice(em->device));
    CU(cudaEventRecord(em->ev[2], em->stream));
    return launch_mstep_local(em);
}

extern "C" int bamm_em_finish_iteration(bamm_em* em, int optimize_q, float* llh, float* vdiff) {
    REQUIRE(em, "em is NULL");
    NOT_FOR_GROUPS(em, "bamm_em_finish_iteration");
    CU(cudaSetDevice(em->device));
    int rc = launch_update(em); if (rc) return rc;
    CU(cudaEventRecord(em->ev[3], em->stream));
    if (!optimize_q && !llh && !vdiff) return BAMM_OK;      // stays asynchronous: no host round trip
    rc = read_scalars(em, true); if (rc) return rc;
    if (optimize_q) em->q = q_from_rsum(em);
    if (llh) *llh = em->llh;
    if (vdiff) *vdiff = *em->h_vdiff;
    return BAMM_OK;
}

extern "C" int bamm_em_set_exchange_buffer(bamm_em* em, void* dev_ptr, uint64_t words) {
    REQUIRE(em && dev_ptr, "NULL argument");
    NOT_FOR_GROUPS(em, "bamm_em_set_exchange_buffer");
    REQUIRE(words == (uint64_t)em->nbin + 2, "exchange buffer must hold %llu 64-bit words", (unsigned long long)em->nbin + 2);
    CU(cudaSetDevice(em->device));
    CU(cudaStreamSynchronize(em->stream));
    if (em->own_xbuf) cudaFree(em->d_xbuf);
    em->d_xbuf = (unsigned long long*)dev_ptr; em->own_xbuf = false;
    CU(cudaMemsetAsync(em->d_xbuf, 0, words * 8, em->stream));
    return BAMM_OK;
}

extern "C" int bamm_em_mstep(bamm_em* em) {
    if (IS_GROUP(em)) return group_mstep(em);
    int rc = bamm_em_mstep_local(em); if (rc) return rc;
    rc = launch_update(em); if (rc) return rc;
    CU(cudaEventRecord(em->ev[3], em->stream));
    CU(cudaStreamSynchronize(em->stream));
    return BAMM_OK;
}

extern "C" int bamm_em_optimize_q(bamm_em* em, float* q) {
    REQUIRE(em, "em is NULL");
    if (IS_GROUP(em)) {                       // every shard holds the global sum of the posteriors after an exchange; before one, sum the shards
        float qs = 0.0f;
        for (bamm_em* sh : em->shards) { int rc = bamm_em_optimize_q(sh, &qs); if (rc) return rc; }
        if (!em->shards[0]->peer_sums_global) {
            double n1 = 0.0;
            for (bamm_em* sh : em->shards) n1 += (double)(long long)sh->h_scal[1] * SC_INV_D;
            qs = ((float)em->nsub - (float)n1 + 1.f) / ((float)em->nsub + 2.f);
            for (bamm_em* sh : em->shards) sh->q = qs;
        }
        em->q = qs;
        if (q) *q = qs;
        return BAMM_OK;
    }
    if (!em->r_valid) return fail(BAMM_E_STATE, "optimize_q needs the r of an E-step");
    CU(cudaSetDevice(em->device));
    int rc = read_scalars(em, false); if (rc) return rc;
    em->q = q_from_rsum(em);
    if (q) *q = em->q;
    return BAMM_OK;
}

extern "C" int bamm_em_optimize(bamm_em* em, int optimize_q, float epsilon, int max_iter, int* iterations,
                                float* llh_trace, float* vdiff_trace, float* q_trace) {
    REQUIRE(em, "em is NULL");
    if (IS_GROUP(em)) return group_optimize(em, optimize_q, epsilon, max_iter, iterations, llh_trace, vdiff_trace, q_trace);
    if (!em->model_set) return fail(BAMM_E_STATE, "bamm_em_set_model has not been called");
    REQUIRE(max_iter >= 1, "max_iter must be >= 1");
    CU(cudaSetDevice(em->device));
    bool iterate = true;
    int it = 0;
    float llh_prev;
    // reference: EM.cpp:79-118
    while (iterate && it < max_iter) {
        it++;
        llh_prev = em->llh;
        int rc = launch_estep(em); if (rc) return rc;
        em->r_valid = true;
        rc = launch_mstep_local(em); if (rc) return rc;
        rc = launch_update(em); if (rc) return rc;
        rc = read_scalars(em, true); if (rc) return rc;
        if (optimize_q && it <= 5) em->q = q_from_rsum(em);
        const float v_diff = *em->h_vdiff;
        const float llh_diff = em->llh - llh_prev;
        if (llh_trace) llh_trace[it - 1] = em->llh;
        if (vdiff_trace) vdiff_trace[it - 1] = v_diff;
        if (q_trace) q_trace[it - 1] = em->q;
        if (v_diff < epsilon) iterate = false;
        if (llh_diff < 0 && it > 10) iterate = false;
    }
    if (iterations) *iterations = it;
    return BAMM_OK;
}

static int device_sort_f32(float* d_keys, uint64_t n, bool descending, cudaStream_t st);
// EM::mask (EM.cpp:261-503, --advanceEM) on the device; see mask.cuh. Single GPU, without optimizeQ.
template <typename YT>
static int mask_run(bamm_em* em, const YT* Y, float f, float epsilon, int max_iter, int* iterations, uint64_t* nkept_out, float* cutoff_out) {
    bamm_seqset* s = em->ss;
    const int W = em->W, sms = s->sm_count;
    const uint64_t nsub = em->nsub;
    MaskView mv; mv.seq_off = s->d_off; mv.seq_ids = em->d_m_ids; mv.r_off = em->d_m_roff; mv.nsub = (uint32_t)nsub;
    cudaStream_t st = em->stream;
    // (1) order-0 table s0[y][j] = v[0][y][j] / vbg[0][y] (EM.cpp:271-275) on the host from the device model
    std::vector<float> v0((size_t)em->A * W), vb0(em->A), s0((size_t)em->A * W);
    CU(cudaMemcpyAsync(v0.data(), em->d_v, v0.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(vb0.data(), em->d_vbg, vb0.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int y = 0; y < em->A; y++) for (int j = 0; j < W; j++) s0[(size_t)y * W + j] = v0[(size_t)y * W + j] / vb0[y];
    float* d_s0 = nullptr; float* d_all = nullptr; uint32_t* d_cnt = nullptr;
    int rc = BAMM_OK;
    uint64_t pos_count = 0;
    host_lists(em);
    std::vector<uint64_t> woff(nsub + 1, 0);
    for (uint64_t i = 0; i < nsub; i++) woff[i + 1] = woff[i] + (em->h_r_off[i + 1] - em->h_r_off[i]) - (uint64_t)W + 1;
    pos_count = woff[nsub];
#define CUX(call) do { cudaError_t e2_ = (call); if (e2_ != cudaSuccess) { rc = fail(e2_ == cudaErrorMemoryAllocation ? BAMM_E_NOMEM : BAMM_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e2_)); goto done; } } while (0)
    {
        CUX(dev_malloc(&d_s0, s0.size() * sizeof(float)));
        CUX(cudaMemcpyAsync(d_s0, s0.data(), s0.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        if (!em->d_m_woff) CUX(dev_malloc(&em->d_m_woff, (nsub + 1) * sizeof(uint64_t)));
        CUX(cudaMemcpyAsync(em->d_m_woff, woff.data(), (nsub + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        CUX(cudaMemsetAsync(em->d_r, 0, (em->rsize ? em->rsize : 1) * sizeof(float), st));       // the reference's calloc
        k_mask_phase1<YT><<<sms * 8, 256, 0, st>>>(Y, mv, W, (uint32_t)em->A, d_s0, em->q, em->d_r);
        CUX(cudaGetLastError());
        // (2) threshold: descending sort of every window's r, value at rank floor(float(count) * f)  (EM.cpp:318-334)
        CUX(dev_malloc(&d_all, (pos_count ? pos_count : 1) * sizeof(float)));
        k_mask_gather<<<sms * 8, 256, 0, st>>>(mv, W, em->d_m_woff, em->d_r, d_all);
        CUX(cudaGetLastError());
        rc = device_sort_f32(d_all, pos_count, true, st);
        if (rc) goto done;
        const size_t rank = (size_t)((float)pos_count * f);
        if (rank >= pos_count) { rc = fail(BAMM_E_INVALID, "fraction f=%g selects no threshold", (double)f); goto done; }
        float cutoff = 0.0f;
        CUX(cudaMemcpy(&cutoff, d_all + rank, sizeof(float), cudaMemcpyDeviceToHost));
        cudaFree(d_all); d_all = nullptr;
        CUX(dev_malloc(&d_cnt, (nsub ? nsub : 1) * sizeof(uint32_t)));
        k_mask_select<false><<<sms * 8, 256, 0, st>>>(mv, em->d_m_woff, em->d_r, cutoff, d_cnt, nullptr, nullptr);
        CUX(cudaGetLastError());
        std::vector<uint32_t> cnt(nsub);
        CUX(cudaMemcpyAsync(cnt.data(), d_cnt, nsub * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CUX(cudaStreamSynchronize(st));
        std::vector<uint64_t> seloff(nsub + 1, 0);
        for (uint64_t i = 0; i < nsub; i++) seloff[i + 1] = seloff[i] + cnt[i];
        cudaFree(em->d_m_seloff); cudaFree(em->d_m_sel); em->d_m_seloff = nullptr; em->d_m_sel = nullptr;
        CUX(dev_malloc(&em->d_m_seloff, (nsub + 1) * sizeof(uint64_t)));
        CUX(dev_malloc(&em->d_m_sel, (seloff[nsub] ? seloff[nsub] : 1) * sizeof(uint32_t)));
        CUX(cudaMemcpyAsync(em->d_m_seloff, seloff.data(), (nsub + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        k_mask_select<true><<<sms * 8, 256, 0, st>>>(mv, em->d_m_woff, em->d_r, cutoff, nullptr, em->d_m_seloff, em->d_m_sel);
        CUX(cudaGetLastError());
        if (nkept_out) *nkept_out = seloff[nsub];
        if (cutoff_out) *cutoff_out = cutoff;
        // (3) EM over the kept windows (EM.cpp:363-495): E, M, fold + updateV + next s, the stop rule of optimize()
        int max_optin = 0;
        cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, em->device);
        const size_t tb = (size_t)em->nbin * 4;
        const bool smem = tb <= (size_t)max_optin && em->nparts > 1;
        int grid_m = sms * 2;
        if (smem) {
            grid_m = std::min<int>((int)em->nparts, sms * std::max(1, std::min(4, (int)((size_t)(max_optin + 1024) / (tb + 1024)))));
            CUX(cudaFuncSetAttribute(k_mask_mstep<YT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb));
        }
        unsigned long long* scal = em->d_xbuf + em->nbin;
        bool iterate = true;
        int it = 0;
        float llh_prev;
        em->llh = 0.0f;                                             // EM.h:61: the member starts at 0
        while (iterate && it < max_iter) {
            it++;
            llh_prev = em->llh;
            CUX(cudaMemsetAsync(scal, 0, 2 * sizeof(unsigned long long), st));
            k_mask_estep<YT><<<sms * 8, 256, 0, st>>>(Y, mv, W, em->Yn, em->d_s, em->q, em->d_m_seloff, em->d_m_sel, em->d_r, scal);
            CUX(cudaGetLastError());
            CUX(cudaMemsetAsync(em->d_part, 0, (uint64_t)em->nparts * em->nbin * sizeof(unsigned long long), st));
            if (smem) k_mask_mstep<YT, true><<<grid_m, 512, tb, st>>>(Y, mv, W, em->Yn, em->d_m_seloff, em->d_m_sel, em->d_r, em->d_part);
            else      k_mask_mstep<YT, false><<<grid_m, 512, 0, st>>>(Y, mv, W, em->Yn, em->d_m_seloff, em->d_m_sel, em->d_r, em->d_part);
            CUX(cudaGetLastError());
            em->launches += 2;
            rc = launch_mstep_reduce(em); if (rc) goto done;
            rc = launch_update(em); if (rc) goto done;
            rc = read_scalars(em, true); if (rc) goto done;
            const float v_diff = *em->h_vdiff;
            const float llh_diff = em->llh - llh_prev;
            if (v_diff < epsilon) iterate = false;
            if (llh_diff < 0 && it > 10) iterate = false;
        }
        if (iterations) *iterations = it;
        em->r_valid = true; em->r_scaled = true; em->r_mat = true;
    }
done:
#undef CUX
    cudaFree(d_s0); cudaFree(d_all); cudaFree(d_cnt);
    return rc;
}

extern "C" int bamm_em_mask(bamm_em* em, float f, float epsilon, int max_iter, int* iterations, float* llh, uint64_t* n_kept, float* r_cutoff) {
    REQUIRE(em, "em is NULL");
    NOT_FOR_GROUPS(em, "bamm_em_mask (EM::mask runs on one device: create the object with bamm_set_device_group of one device)");
    if (!em->model_set) return fail(BAMM_E_STATE, "bamm_em_set_model has not been called");
    REQUIRE(max_iter >= 1, "max_iter must be >= 1");
    REQUIRE(f > 0.0f && f < 1.0f, "fraction f=%g not in (0,1)", (double)f);
    REQUIRE(em->W >= 2, "EM::mask needs a motif of at least two columns");     // the reference reads pos_[L] for W = 1
    REQUIRE(em->nsub >= 1, "empty sequence subset");
    REQUIRE(!em->peer_attached, "EM::mask runs on one device");
    CU(cudaSetDevice(em->device));
    IndexArray* ia = nullptr;
    { std::lock_guard<std::mutex> g(em->ss->mu); int rc = seqset_index_locked(em->ss, em->K, &ia); if (rc) return rc; }
    host_lists(em);
    if (!em->d_m_ids) {
        CU(dev_malloc(&em->d_m_ids, em->nsub * sizeof(uint32_t)));
        CU(dev_malloc(&em->d_m_roff, (em->nsub + 1) * sizeof(uint64_t)));
        CU(cudaMemcpy(em->d_m_ids, em->h_ids.data(), em->nsub * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(em->d_m_roff, em->h_r_off.data(), (em->nsub + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
    }
    int rc = ia->bytes == 2 ? mask_run<uint16_t>(em, (const uint16_t*)ia->d, f, epsilon, max_iter, iterations, n_kept, r_cutoff)
                            : mask_run<uint32_t>(em, (const uint32_t*)ia->d, f, epsilon, max_iter, iterations, n_kept, r_cutoff);
    if (rc) return rc;
    if (llh) *llh = em->llh;
    return BAMM_OK;
}

// launches of one iteration with optional event brackets (ev = LOOP_EV events or nullptr):
// start | masked windows | bounds | rest of the E-step | M-step accumulation | reduce + update
constexpr int LOOP_EV = 6;
static int launch_iteration(bamm_em* em, cudaEvent_t* ev) {
    if (ev) CU(cudaEventRecord(ev[0], em->stream));
    int rc = launch_estep(em, ev ? ev + 1 : nullptr); if (rc) return rc;
    if (ev && !em->npk) { CU(cudaEventRecord(ev[1], em->stream)); CU(cudaEventRecord(ev[2], em->stream)); }
    if (ev) CU(cudaEventRecord(ev[3], em->stream));
    rc = launch_mstep_accumulate(em); if (rc) return rc;
    if (ev) CU(cudaEventRecord(ev[4], em->stream));
    rc = launch_mstep_reduce(em); if (rc) return rc;
    rc = launch_update(em); if (rc) return rc;
    if (ev) CU(cudaEventRecord(ev[5], em->stream));
    return BAMM_OK;
}

extern "C" int bamm_em_iterate(bamm_em* em, int n_iter, float* llh_last, float* vdiff_last) {
    REQUIRE(em, "em is NULL");
    if (IS_GROUP(em)) return group_iterate(em, n_iter, llh_last, vdiff_last);
    if (!em->model_set) return fail(BAMM_E_STATE, "bamm_em_set_model has not been called");
    REQUIRE(n_iter >= 0, "n_iter must be >= 0");
    CU(cudaSetDevice(em->device));
    while ((int)em->loop_ev.size() < LOOP_EV * n_iter) { cudaEvent_t e; CU(cudaEventCreate(&e)); em->loop_ev.push_back(e); }
    em->loop_iters = 0;
    for (int it = 0; it < n_iter; it++) {
        int rc = launch_iteration(em, &em->loop_ev[LOOP_EV * it]); if (rc) return rc;
    }
    em->loop_iters = n_iter;
    em->r_valid = n_iter > 0 || em->r_valid;
    int rc = read_scalars(em, true); if (rc) return rc;
    if (llh_last) *llh_last = em->llh;
    if (vdiff_last) *vdiff_last = *em->h_vdiff;
    return BAMM_OK;
}

static int loop_times(bamm_em* em, float out[5], float* total) {
    CU(cudaSetDevice(em->device));
    CU(cudaStreamSynchronize(em->stream));
    for (int k = 0; k < 5; k++) out[k] = 0.0f;
    float x;
    for (int it = 0; it < em->loop_iters; it++) {
        cudaEvent_t* ev = &em->loop_ev[LOOP_EV * it];
        for (int k = 0; k < 5; k++) { CU(cudaEventElapsedTime(&x, ev[k], ev[k + 1])); out[k] += x; }
    }
    *total = 0.0f;
    if (em->loop_iters) CU(cudaEventElapsedTime(total, em->loop_ev[0], em->loop_ev[LOOP_EV * em->loop_iters - 1]));
    return BAMM_OK;
}

extern "C" int bamm_em_loop_timing(bamm_em* em, int* iters, float* estep_ms, float* maccum_ms, float* update_ms, float* total_ms) {
    if (IS_GROUP(em)) return bamm_em_loop_timing(em->shards[0], iters, estep_ms, maccum_ms, update_ms, total_ms);        // identical on every shard (global sums) / shard 0 as the sample
    REQUIRE(em, "em is NULL");
    float t[5], tot;
    int rc = loop_times(em, t, &tot); if (rc) return rc;
    if (iters) *iters = em->loop_iters;
    if (estep_ms) *estep_ms = t[0] + t[1] + t[2];
    if (maccum_ms) *maccum_ms = t[3];
    if (update_ms) *update_ms = t[4];
    if (total_ms) *total_ms = tot;
    return BAMM_OK;
}

extern "C" int bamm_em_loop_timing_estep(bamm_em* em, float* masked_ms, float* bound_ms, float* exact_ms) {
    if (IS_GROUP(em)) return bamm_em_loop_timing_estep(em->shards[0], masked_ms, bound_ms, exact_ms);        // identical on every shard (global sums) / shard 0 as the sample
    REQUIRE(em, "em is NULL");
    float t[5], tot;
    int rc = loop_times(em, t, &tot); if (rc) return rc;
    if (masked_ms) *masked_ms = t[0];
    if (bound_ms) *bound_ms = t[1];
    if (exact_ms) *exact_ms = t[2];
    return BAMM_OK;
}

extern "C" int bamm_em_last_timing(bamm_em* em, float* estep_ms, float* mstep_ms) {
    if (IS_GROUP(em)) return bamm_em_last_timing(em->shards[0], estep_ms, mstep_ms);        // identical on every shard (global sums) / shard 0 as the sample
    REQUIRE(em, "em is NULL");
    CU(cudaSetDevice(em->device));
    CU(cudaStreamSynchronize(em->stream));
    float e = 0, m = 0;
    if (cudaEventElapsedTime(&e, em->ev[0], em->ev[1]) != cudaSuccess) { e = 0; cudaGetLastError(); }
    if (cudaEventElapsedTime(&m, em->ev[2], em->ev[3]) != cudaSuccess) { m = 0; cudaGetLastError(); }
    if (estep_ms) *estep_ms = e;
    if (mstep_ms) *mstep_ms = m;
    return BAMM_OK;
}

extern "C" int bamm_em_get_model(bamm_em* em, float* v_all) {
    if (IS_GROUP(em)) return bamm_em_get_model(em->shards[0], v_all);        // identical on every shard (global sums) / shard 0 as the sample
    REQUIRE(em && v_all, "NULL argument");
    CU(cudaSetDevice(em->device));
    CU(cudaMemcpyAsync(v_all, em->d_v, em->model_size * sizeof(float), cudaMemcpyDeviceToHost, em->stream));
    CU(cudaStreamSynchronize(em->stream));
    return BAMM_OK;
}
extern "C" int bamm_em_get_counts(bamm_em* em, float* n_all) {
    if (IS_GROUP(em)) return bamm_em_get_counts(em->shards[0], n_all);        // identical on every shard (global sums) / shard 0 as the sample
    REQUIRE(em && n_all, "NULL argument");
    CU(cudaSetDevice(em->device));
    CU(cudaMemcpyAsync(n_all, em->d_n, em->model_size * sizeof(float), cudaMemcpyDeviceToHost, em->stream));
    CU(cudaStreamSynchronize(em->stream));
    return BAMM_OK;
}
extern "C" int bamm_em_get_s(bamm_em* em, float* s) {
    if (IS_GROUP(em)) return bamm_em_get_s(em->shards[0], s);        // identical on every shard (global sums) / shard 0 as the sample
    REQUIRE(em && s, "NULL argument");
    CU(cudaSetDevice(em->device));
    std::vector<float> t(em->nbin);
    // the table the last E-step read (what Motif::getS() holds after EM::optimize, EM.cpp:144), or the current one before any
    CU(cudaMemcpyAsync(t.data(), em->r_valid && em->d_s_e ? em->d_s_e : em->d_s, (uint64_t)em->nbin * sizeof(float), cudaMemcpyDeviceToHost, em->stream));
    CU(cudaStreamSynchronize(em->stream));
    for (uint32_t y = 0; y < em->Yn; y++) for (int j = 0; j < em->W; j++) s[(uint64_t)y * em->W + j] = t[(uint64_t)j * em->Yn + y];
    return BAMM_OK;
}
extern "C" int bamm_em_get_q(bamm_em* em, float* q) { REQUIRE(em && q, "NULL argument"); *q = IS_GROUP(em) ? em->shards[0]->q : em->q; return BAMM_OK; }
extern "C" uint64_t bamm_em_r_size(const bamm_em* em) { return em ? em->rsize : 0; }       // a group object carries the sum of its shards
extern "C" int bamm_em_get_r(bamm_em* em, uint64_t first, uint64_t count, float* out) {
    REQUIRE(em && out, "NULL argument");
    if (IS_GROUP(em)) return group_get_r(em, first, count, out);
    REQUIRE(first + count <= em->nsub, "sequence range out of bounds");
    if (!em->r_valid) return fail(BAMM_E_STATE, "no E-step has run");
    CU(cudaSetDevice(em->device));
    if (!em->r_mat && em->npk) {
        // the pruned E-step does not write r: run the dense kernel once with the tables of that E-step (no list, no scalars;
        // the normalisers it stores are bit-identical to the ones already there)
        PackedView pv = pview_of(em);
        const EStepLaunch l = {em->grid_pe, em->block_pe, em->stream};
        ActiveList al = alist_of(em); al.ent = nullptr;
        for (size_t pass = 0; pass < em->gplans.size(); pass++)
            if (launch_estep_dense(l, false, em->gfast[pass] != 0, em->gplans.size() > 1, &pv, plan_for_launch(em, pass, em->q_e),
                                   (const float*)((const char*)em->d_tab_e + pass * em->tab_capacity), em->d_s_e, em->plain_words, em->d_r, nullptr, &al, nullptr))
                return fail(BAMM_E_CUDA, "packed E-step launch failed (materialising r)");
        em->launches += em->gplans.size();
        em->r_mat = true; em->r_scaled = false;
    }
    if (!em->r_scaled) {            // the packed E-step keeps r unnormalised; finish it before it leaves the device
        k_normalise_r<<<em->ss->sm_count * 8, 256, 0, em->stream>>>(pview_of(em), em->W, em->d_scale, em->d_r);
        CU(cudaGetLastError());
        em->r_scaled = true;
    }
    host_lists(em);
    const uint64_t a = em->h_r_off[first], b = em->h_r_off[first + count];
    CU(cudaMemcpyAsync(out, em->d_r + a, (b - a) * sizeof(float), cudaMemcpyDeviceToHost, em->stream));
    CU(cudaStreamSynchronize(em->stream));
    return BAMM_OK;
}

extern "C" int bamm_em_exchange_buffer(bamm_em* em, void** dev_ptr, uint64_t* words) {
    REQUIRE(em && dev_ptr && words, "NULL argument");
    NOT_FOR_GROUPS(em, "bamm_em_exchange_buffer");
    *dev_ptr = em->d_xbuf; *words = (uint64_t)em->nbin + 2;
    return BAMM_OK;
}
extern "C" int bamm_em_set_global_nseq(bamm_em* em, uint64_t n) {
    REQUIRE(em, "em is NULL");
    em->nseq_global = n;
    for (bamm_em* sh : em->shards) sh->nseq_global = n;
    return BAMM_OK;
}
extern "C" int bamm_em_launch_count(bamm_em* em, uint64_t* kernels) {
    REQUIRE(em && kernels, "NULL argument");
    *kernels = em->launches;
    for (bamm_em* sh : em->shards) *kernels += sh->launches;
    return BAMM_OK;
}
extern "C" int bamm_em_estep_info(bamm_em* em, uint64_t info[8]) {
    if (IS_GROUP(em)) return bamm_em_estep_info(em->shards[0], info);        // identical on every shard (global sums) / shard 0 as the sample
    REQUIRE(em && info, "NULL argument");
    CU(cudaSetDevice(em->device));
    CU(cudaStreamSynchronize(em->stream));
    for (int i = 0; i < 8; i++) info[i] = 0;
    info[0] = em->sparse ? 1 : 0;
    info[1] = em->gplans.empty() ? 0 : (uint64_t)em->gplans[0].G;
    info[2] = em->sparse ? (uint64_t)em->bplan.G : 0;
    info[3] = 1;
    if (em->sparse && em->d_eflags) {
        uint32_t f[4] = {0, 0, 0, 0};
        CU(cudaMemcpy(f, em->d_eflags, 16, cudaMemcpyDeviceToHost));
        info[3] = f[0] ? 1 : 0;
        info[4] = (uint64_t)f[2] | ((uint64_t)f[3] << 32);
    }
    if (em->d_act_cnt && em->nregions) {
        std::vector<uint32_t> c(2 * (size_t)em->nregions);
        CU(cudaMemcpy(c.data(), em->d_act_cnt, c.size() * 4, cudaMemcpyDeviceToHost));
        for (uint32_t x : c) info[5] += x;
    }
    info[6] = em->gplans.size();
    info[7] = em->plain_words ? 1 : 0;
    return BAMM_OK;
}
extern "C" int bamm_em_stream(bamm_em* em, void** stream) { REQUIRE(em && stream, "NULL argument"); NOT_FOR_GROUPS(em, "bamm_em_stream"); *stream = (void*)em->stream; return BAMM_OK; }


// ---- NVLink peer exchange ------------------------------------------------------------------------------------------
extern "C" int bamm_em_peer_alloc(bamm_em* em, int rank, int world, void* ipc_handle_out) {
    REQUIRE(em && ipc_handle_out, "NULL argument");
    NOT_FOR_GROUPS(em, "bamm_em_peer_alloc");
    REQUIRE(world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world, "rank %d / world %d out of range (max %d ranks)", rank, world, MAX_PEERS);
    REQUIRE(!em->d_peer_local, "peer buffer already allocated");
    CU(cudaSetDevice(em->device));
    const size_t words = (size_t)em->nbin + 2;
    const size_t slot_bytes = (size_t)2 * world * words * sizeof(unsigned long long);
    const size_t bytes = slot_bytes + MAX_PEERS * sizeof(unsigned int);
    CU(cudaMalloc(&em->d_peer_local, bytes));            // CUDA IPC needs a plain allocation
    CU(cudaMemset(em->d_peer_local, 0, bytes));
    CU(cudaMalloc(&em->d_peer_done, sizeof(unsigned int)));
    CU(cudaMemset(em->d_peer_done, 0, sizeof(unsigned int)));
    CU(cudaMalloc(&em->d_peer_wait, 2 * sizeof(unsigned long long)));
    CU(cudaMemset(em->d_peer_wait, 0, 2 * sizeof(unsigned long long)));
    CU(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, em->d_peer_local));
    static_assert(sizeof(h) == 64, "CUDA IPC handle size");
    memcpy(ipc_handle_out, &h, sizeof(h));
    em->peer_rank = rank; em->peer_world = world;
    return BAMM_OK;
}

extern "C" int bamm_em_peer_wait(bamm_em* em, int reset, double* total_ms, uint64_t* waits) {
    REQUIRE(em && total_ms && waits, "NULL argument");
    if (IS_GROUP(em)) return bamm_em_peer_wait(em->shards[0], reset, total_ms, waits);
    *total_ms = 0.0; *waits = 0;
    if (!em->d_peer_wait) return BAMM_OK;
    CU(cudaSetDevice(em->device));
    CU(cudaStreamSynchronize(em->stream));
    unsigned long long w[2] = {0, 0};
    CU(cudaMemcpy(w, em->d_peer_wait, sizeof(w), cudaMemcpyDeviceToHost));
    *total_ms = (double)w[0] * 1e-6; *waits = w[1];
    if (reset) CU(cudaMemset(em->d_peer_wait, 0, sizeof(w)));
    return BAMM_OK;
}

extern "C" int bamm_em_peer_attach(bamm_em* em, const void* ipc_handles) {
    REQUIRE(em && ipc_handles, "NULL argument");
    NOT_FOR_GROUPS(em, "bamm_em_peer_attach");
    if (!em->d_peer_local) return fail(BAMM_E_STATE, "bamm_em_peer_alloc has not been called");
    CU(cudaSetDevice(em->device));
    const size_t words = (size_t)em->nbin + 2;
    const size_t slot_bytes = (size_t)2 * em->peer_world * words * sizeof(unsigned long long);
    for (int p = 0; p < em->peer_world; p++) {
        unsigned char* base = em->d_peer_local;
        if (p != em->peer_rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const unsigned char*)ipc_handles + (size_t)p * sizeof(h), sizeof(h));
            void* mapped = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return fail(BAMM_E_CUDA, "cudaIpcOpenMemHandle for rank %d failed: %s", p, cudaGetErrorString(e));
            em->peer_mapped[p] = mapped;
            base = (unsigned char*)mapped;
        }
        em->peer_ptrs.slots[p] = (unsigned long long*)base;
        em->peer_ptrs.flags[p] = (unsigned int*)(base + slot_bytes);
    }
    em->peer_attached = true;
    return BAMM_OK;
}
