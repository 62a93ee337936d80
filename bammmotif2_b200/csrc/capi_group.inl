// capi_group.inl — part of capi.cu. Device groups: ONE process drives several devices.
//
// bamm_set_device_group names the devices. An EM object created afterwards is a facade over one ordinary EM object per device:
// the subset is cut into contiguous blocks (whole sequences, balanced by length), block d runs on device d over a copy of the
// sequence set on that device (made on first use with device-to-device copies: bases and patch list travel over NVLink, the
// k-mer classification and the 2-bit packing run again on the copy), the per-iteration count tensor and scalars are exchanged
// by the M-step's own reduction kernel through peer memory (k_reduce_push / k_peer_sum, kernels.cuh) — the path the
// multi-process runs use, with plain peer pointers instead of CUDA-IPC handles. Every shard holds the same sums, takes the same
// decisions (stop rule, q) and ends with the same model bits as a single device would (integer sums, normaliser included).
// The loop calls (bamm_em_optimize, bamm_em_iterate) run one host thread per device; the step-wise calls do the same per call.
#include <thread>

static std::mutex g_group_mu;
static std::vector<int> g_group;                      // devices of the group; empty or one entry: no group

extern "C" int bamm_set_device_group(const int* devices, int n) {
    REQUIRE(n >= 0 && n <= MAX_PEERS, "a device group has at most %d devices", MAX_PEERS);
    REQUIRE(n == 0 || devices, "devices is NULL");
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    for (int i = 0; i < n; i++) {
        REQUIRE(devices[i] >= 0 && devices[i] < count, "device %d does not exist (%d devices)", devices[i], count);
        for (int j = 0; j < i; j++) REQUIRE(devices[i] != devices[j], "device %d is listed twice", devices[i]);
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            if (i == j) continue;
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, devices[i], devices[j]));
            REQUIRE(can, "device %d cannot access the memory of device %d", devices[i], devices[j]);
            CU(cudaSetDevice(devices[i]));
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
            cudaGetLastError();
        }
    std::lock_guard<std::mutex> g(g_group_mu);
    g_group.assign(devices, devices + n);
    if (n) CU(cudaSetDevice(devices[0]));               // objects created from this thread live on the first device
    return BAMM_OK;
}
extern "C" int bamm_get_device_group(int* devices, int cap, int* n) {
    REQUIRE(n, "n is NULL");
    std::lock_guard<std::mutex> g(g_group_mu);
    *n = (int)g_group.size();
    for (int i = 0; i < *n && i < cap && devices; i++) devices[i] = g_group[i];
    return BAMM_OK;
}

static std::vector<int> group_devices() { std::lock_guard<std::mutex> g(g_group_mu); return g_group; }

// a group only pays for itself when every device gets a reasonable block; tiny subsets stay on the set's own device
static bool group_wanted(const bamm_seqset* s, uint64_t nsub) {
    const std::vector<int> devs = group_devices();
    if (devs.size() < 2 || devs[0] != s->device) return false;
    const char* mn = getenv("BAMM_GROUP_MIN_SEQS");
    const uint64_t per = mn ? (uint64_t)atoll(mn) : 2048;
    return nsub >= per * devs.size();
}

// copy of the set on `device` (same sequences, same order, same patch list), made once and owned by the set
static int seqset_replica(bamm_seqset* s, int device, bamm_seqset** out) {
    std::lock_guard<std::mutex> g(s->mu);
    auto it = s->replicas.find(device);
    if (it != s->replicas.end()) { *out = it->second; return BAMM_OK; }
    seqset_copy_done(s);
    REQUIRE(!s->d_zero_pos, "the set still waits for its patch list (bamm_seqset_finish_patches)");
    CU(cudaSetDevice(device));
    bamm_seqset* r = nullptr;
    int rc = seqset_new(s->h_off.data(), s->nseq, s->A, &r);
    if (rc == BAMM_OK) {
        cudaError_t e = s->npos ? cudaMemcpyPeer(r->d_codes, device, s->d_codes, s->device, s->npos) : cudaSuccess;
        r->npatch = s->npatch;
        if (e == cudaSuccess && s->npatch) {
            e = dev_malloc(&r->d_ppos, s->npatch * sizeof(uint64_t));
            if (e == cudaSuccess) e = dev_malloc(&r->d_pkmer, s->npatch * sizeof(uint64_t));
            if (e == cudaSuccess) e = cudaMemcpyPeer(r->d_ppos, device, s->d_ppos, s->device, s->npatch * sizeof(uint64_t));
            if (e == cudaSuccess) e = cudaMemcpyPeer(r->d_pkmer, device, s->d_pkmer, s->device, s->npatch * sizeof(uint64_t));
        }
        if (e != cudaSuccess) { bamm_seqset_destroy(r); r = nullptr; rc = fail(e == cudaErrorMemoryAllocation ? BAMM_E_NOMEM : BAMM_E_CUDA, "copying the sequence set to device %d failed: %s", device, cudaGetErrorString(e)); }
    }
    if (rc == BAMM_OK) { rc = seqset_finish(r, false, false); if (rc) r = nullptr; }   // seqset_finish destroys the set when it fails
    cudaSetDevice(s->device);
    if (rc) return rc;
    s->replicas[device] = r;
    *out = r;
    return BAMM_OK;
}

// runs fn(d) for every shard on its own host thread (shard 0 on the calling thread); first failure wins
template <typename F> static int group_parallel(bamm_em* em, F fn) {
    const size_t n = em->shards.size();
    std::vector<int> rcs(n, BAMM_OK);
    std::vector<std::string> msgs(n);
    std::vector<std::thread> th;
    for (size_t d = 1; d < n; d++)
        th.emplace_back([&, d]() { cudaSetDevice(em->shards[d]->device); rcs[d] = fn(d); if (rcs[d]) msgs[d] = g_err; });
    cudaSetDevice(em->shards[0]->device);
    rcs[0] = fn(0);
    if (rcs[0]) msgs[0] = g_err;
    for (std::thread& t : th) t.join();
    cudaSetDevice(em->shards[0]->device);
    for (size_t d = 0; d < n; d++) if (rcs[d]) return fail(rcs[d], "device %d: %s", em->shards[d]->device, msgs[d].c_str());
    return BAMM_OK;
}

static void group_destroy(bamm_em* em) {
    for (bamm_em* sh : em->shards) if (sh) { cudaSetDevice(sh->device); bamm_em_destroy(sh); }
    if (!em->shards.empty() && em->ss) cudaSetDevice(em->ss->device);
    delete em;
}

static int group_create(bamm_seqset* s, const uint64_t* subset, uint64_t nsub, int W, int K, int K_bg_model, bamm_em** out) {
    const std::vector<int> devs = group_devices();
    const size_t n = devs.size();
    // contiguous blocks of the subset, balanced by the number of windows
    std::vector<uint64_t> ids(nsub);
    uint64_t total = 0;
    for (uint64_t i = 0; i < nsub; i++) {
        const uint64_t q = subset ? subset[i] : i;
        REQUIRE(q < s->nseq, "subset[%llu]=%llu out of range", (unsigned long long)i, (unsigned long long)q);
        ids[i] = q;
        total += s->h_off[q + 1] - s->h_off[q];
    }
    std::vector<uint64_t> first(n + 1, nsub);
    first[0] = 0;
    {
        uint64_t acc = 0; size_t d = 1;
        for (uint64_t i = 0; i < nsub && d < n; i++) {
            acc += s->h_off[ids[i] + 1] - s->h_off[ids[i]];
            while (d < n && acc * n >= total * d) first[d++] = i + 1;
        }
    }
    bamm_em* g = new (std::nothrow) bamm_em();
    if (!g) return fail(BAMM_E_NOMEM, "host allocation failed");
    g->ss = s; g->device = s->device; g->W = W; g->K = K; g->K_bg_model = K_bg_model; g->K_bg = K_bg_model < K ? K_bg_model : K; g->A = s->A;
    g->nsub = nsub; g->nseq_global = nsub;
    g->shards.assign(n, nullptr);
    g->shard_first = first;
    // the copies of the set, then one EM object per device (in parallel: each builds its lists and buffers on its own device)
    std::vector<bamm_seqset*> sets(n, nullptr);
    sets[0] = s;
    for (size_t d = 1; d < n; d++) { int rc = seqset_replica(s, devs[d], &sets[d]); if (rc) { group_destroy(g); return rc; } }
    {
        std::vector<int> rcs(n, BAMM_OK); std::vector<std::string> msgs(n);
        std::vector<std::thread> th;
        auto make = [&](size_t d) {
            cudaSetDevice(devs[d]);
            rcs[d] = bamm_em_create(sets[d], ids.data() + first[d], first[d + 1] - first[d], W, K, K_bg_model, &g->shards[d]);
            if (rcs[d]) msgs[d] = g_err;
        };
        // shards are ordinary objects: the group is switched off while they are made (bamm_em_create would recurse)
        std::vector<int> saved;
        { std::lock_guard<std::mutex> lk(g_group_mu); saved.swap(g_group); }
        for (size_t d = 1; d < n; d++) th.emplace_back(make, d);
        make(0);
        for (std::thread& t : th) t.join();
        { std::lock_guard<std::mutex> lk(g_group_mu); g_group.swap(saved); }
        cudaSetDevice(s->device);
        for (size_t d = 0; d < n; d++) if (rcs[d]) { const int rc = fail(rcs[d], "device %d: %s", devs[d], msgs[d].c_str()); group_destroy(g); return rc; }
    }
    // peer exchange through plain peer pointers: every shard gets a receive buffer on its device, all of them know all buffers
    const size_t words = (size_t)g->shards[0]->nbin + 2;
    const size_t slot_bytes = (size_t)2 * n * words * sizeof(unsigned long long);
    for (size_t d = 0; d < n; d++) {
        bamm_em* sh = g->shards[d];
        cudaSetDevice(sh->device);
        cudaError_t e = cudaMalloc(&sh->d_peer_local, slot_bytes + MAX_PEERS * sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMemset(sh->d_peer_local, 0, slot_bytes + MAX_PEERS * sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMalloc(&sh->d_peer_done, sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMemset(sh->d_peer_done, 0, sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMalloc(&sh->d_peer_wait, 2 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemset(sh->d_peer_wait, 0, 2 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { const int rc = fail(BAMM_E_CUDA, "peer buffers on device %d: %s", sh->device, cudaGetErrorString(e)); group_destroy(g); return rc; }
        sh->peer_rank = (int)d; sh->peer_world = (int)n; sh->nseq_global = nsub;
        g->rsize += sh->rsize;
    }
    for (size_t d = 0; d < n; d++) {
        bamm_em* sh = g->shards[d];
        for (size_t p = 0; p < n; p++) {
            sh->peer_ptrs.slots[p] = (unsigned long long*)g->shards[p]->d_peer_local;
            sh->peer_ptrs.flags[p] = (unsigned int*)(g->shards[p]->d_peer_local + slot_bytes);
        }
        sh->peer_attached = true;
    }
    cudaSetDevice(s->device);
    g->model_size = g->shards[0]->model_size; g->bg_size = g->shards[0]->bg_size; g->nbin = g->shards[0]->nbin; g->Yn = g->shards[0]->Yn;
    *out = g;
    return BAMM_OK;
}

static int group_set_model(bamm_em* em, const float* v_all, const float* vbg_all, const float* alpha, float q) {
    int rc = group_parallel(em, [&](size_t d) { return bamm_em_set_model(em->shards[d], v_all, vbg_all, alpha, q); });
    if (rc) return rc;
    em->q = q; em->model_set = true; em->llh = 0.0f;
    return BAMM_OK;
}

static int group_optimize(bamm_em* em, int optimize_q, float epsilon, int max_iter, int* iterations, float* llh_trace, float* vdiff_trace, float* q_trace) {
    REQUIRE(max_iter >= 1, "max_iter must be >= 1");
    // every shard runs the reference's loop (EM.cpp:79-118) on the same global sums: same trace, same stop, same q everywhere
    std::vector<int> its(em->shards.size(), 0);
    int rc = group_parallel(em, [&](size_t d) {
        return bamm_em_optimize(em->shards[d], optimize_q, epsilon, max_iter, &its[d], d == 0 ? llh_trace : nullptr, d == 0 ? vdiff_trace : nullptr, d == 0 ? q_trace : nullptr);
    });
    if (rc) return rc;
    for (size_t d = 1; d < its.size(); d++)
        if (its[d] != its[0]) return fail(BAMM_E_STATE, "devices %d and %d stopped after different numbers of iterations (%d, %d)", em->shards[0]->device, em->shards[d]->device, its[0], its[d]);
    if (iterations) *iterations = its[0];
    em->q = em->shards[0]->q; em->llh = em->shards[0]->llh; em->r_valid = true;
    return BAMM_OK;
}

static int group_iterate(bamm_em* em, int n_iter, float* llh_last, float* vdiff_last) {
    REQUIRE(n_iter >= 0, "n_iter must be >= 0");
    int rc = group_parallel(em, [&](size_t d) { return bamm_em_iterate(em->shards[d], n_iter, d == 0 ? llh_last : nullptr, d == 0 ? vdiff_last : nullptr); });
    if (rc) return rc;
    em->llh = em->shards[0]->llh; em->r_valid = n_iter > 0 || em->r_valid;
    return BAMM_OK;
}

// E-step on every shard; the log likelihood is the sum of the shards' fixed-point sums (associative: same bits as one device)
static int group_estep(bamm_em* em, float* llh) {
    int rc = group_parallel(em, [&](size_t d) { return bamm_em_estep(em->shards[d], nullptr); });
    if (rc) return rc;
    long long sum = 0;
    for (bamm_em* sh : em->shards) sum += (long long)sh->h_scal[0];
    em->llh = (float)((double)sum * SC_INV_D);
    em->r_valid = true;
    if (llh) *llh = em->llh;
    return BAMM_OK;
}

static int group_mstep(bamm_em* em) {
    if (!em->r_valid) return fail(BAMM_E_STATE, "M-step needs the r of an E-step");
    return group_parallel(em, [&](size_t d) { return bamm_em_mstep(em->shards[d]); });
}

// r of subset sequences [first, first+count): the shards hold contiguous blocks of the subset, in order
static int group_get_r(bamm_em* em, uint64_t first, uint64_t count, float* out) {
    REQUIRE(first + count <= em->nsub, "sequence range out of bounds");
    for (size_t d = 0; d < em->shards.size() && count; d++) {
        const uint64_t a = em->shard_first[d], b = em->shard_first[d + 1];
        if (first >= b || first + count <= a) continue;
        const uint64_t lo = std::max(first, a), hi = std::min(first + count, b);
        bamm_em* sh = em->shards[d];
        cudaSetDevice(sh->device);
        int rc = bamm_em_get_r(sh, lo - a, hi - lo, out);
        cudaSetDevice(em->shards[0]->device);
        if (rc) return rc;
        host_lists(sh);
        out += sh->h_r_off[hi - a] - sh->h_r_off[lo - a];
    }
    return BAMM_OK;
}
