// capi.cu — implementation of include/bamm_b200.h on top of kernels.cuh. Host plumbing only: device memory,
// one CUDA stream per EM object, launches, scalar read-back. No CPU fallback: every compute entry point needs a device.
#include "../../include/bamm_b200.h"
#include "kernels.cuh"
#include "packed.cuh"
#include "negatives.cuh"
#include "stats.cuh"
#include <cub/device/device_scan.cuh>
#include "fasta.cuh"
#include "mask.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace bamm;

// ------------------------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";
static thread_local float g_score_ms = 0.0f;      // device time of the scoring kernels of the last bamm_score_logodds call on this thread
static int fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(e_ == cudaErrorMemoryAllocation ? BAMM_E_NOMEM : BAMM_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define REQUIRE(cond, ...) do { if (!(cond)) return fail(BAMM_E_INVALID, __VA_ARGS__); } while (0)

// Device memory comes from the device's default memory pool (cudaMallocAsync) with an unlimited release threshold: buffers a
// finished object gives back stay mapped, so creating the next sequence set / EM object / scoring call does not pay the
// driver's map-and-zero cost again (measured: the same 1-12 GB allocations vary between 3 ms and 150 ms with cudaMalloc
// after a large cudaFree). cudaFree() returns pool memory to the pool. Falls back to cudaMalloc where pools are missing.
static cudaError_t pool_malloc(void** p, size_t bytes) {
    static thread_local int ready_dev = -1;
    static thread_local bool usable = false;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (ready_dev != dev) {
        int supported = 0;
        usable = !getenv("BAMM_NO_MEMPOOL") && cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev) == cudaSuccess && supported;
        if (usable) {
            cudaMemPool_t pool;
            unsigned long long thr = ~0ull;
            usable = cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess &&
                     cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr) == cudaSuccess;
        }
        cudaGetLastError();
        ready_dev = dev;
    }
    if (!usable) return cudaMalloc(p, bytes);
    e = cudaMallocAsync(p, bytes, 0);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(0);                          // the buffer is used from other (non-blocking) streams right away
}
template <typename T> static cudaError_t dev_malloc(T** p, size_t bytes) { return pool_malloc(reinterpret_cast<void**>(p), bytes); }

// BAMM_TRACE=1: wall-clock phases of the set-up calls on stderr (diagnostics of the end-to-end path)
#include <chrono>
struct Trace {
    bool on; const char* what; std::chrono::steady_clock::time_point t0;
    explicit Trace(const char* w) : on(getenv("BAMM_TRACE") != nullptr), what(w), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* phase) {
        if (getenv("BAMM_SYNC_MARKS")) cudaDeviceSynchronize();
        if (!on) return;
        cudaDeviceSynchronize();
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[bamm trace] %s: %s %.2f ms\n", what, phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};
// Pinned host scalars of the EM objects (log likelihood, sum of posteriors, sum|dv| read back every iteration): slots of one
// process-wide pinned slab — a pinned allocation per object costs milliseconds (FDR creates an EM object per fold).
static std::mutex g_pin_mu;
static unsigned long long* g_pin_slab = nullptr;
static std::vector<int> g_pin_free;
constexpr int PIN_SLOTS = 256, PIN_SLOT_WORDS = 4;
static cudaError_t pinned_scalars_get(unsigned long long** out) {
    {
        std::lock_guard<std::mutex> g(g_pin_mu);
        if (!g_pin_slab) {
            void* p = nullptr;
            if (cudaHostAlloc(&p, (size_t)PIN_SLOTS * PIN_SLOT_WORDS * sizeof(unsigned long long), cudaHostAllocPortable) == cudaSuccess) {
                g_pin_slab = static_cast<unsigned long long*>(p);
                for (int i = PIN_SLOTS - 1; i >= 0; i--) g_pin_free.push_back(i);
            } else cudaGetLastError();
        }
        if (g_pin_slab && !g_pin_free.empty()) {
            *out = g_pin_slab + (size_t)g_pin_free.back() * PIN_SLOT_WORDS;
            g_pin_free.pop_back();
            return cudaSuccess;
        }
    }
    return cudaMallocHost(out, PIN_SLOT_WORDS * sizeof(unsigned long long));
}
static void pinned_scalars_put(unsigned long long* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> g(g_pin_mu);
        if (g_pin_slab && p >= g_pin_slab && p < g_pin_slab + (size_t)PIN_SLOTS * PIN_SLOT_WORDS) {
            g_pin_free.push_back((int)((p - g_pin_slab) / PIN_SLOT_WORDS));
            return;
        }
    }
    cudaFreeHost(p);
}
static uint64_t ipow_u64(uint64_t b, int e) { uint64_t r = 1; while (e-- > 0) r *= b; return r; }

// ------------------------------------------------------------------------------------------- objects
struct IndexArray { void* d = nullptr; int bytes = 2; uint64_t Yn = 0; };

struct bamm_seqset {
    int device = 0;
    int A = 4;
    uint64_t nseq = 0, npos = 0, npatch = 0, maxL = 0, minL = 0;
    uint8_t* d_codes = nullptr;
    uint64_t* d_off = nullptr;
    uint64_t* d_ppos = nullptr;
    uint64_t* d_pkmer = nullptr;
    std::vector<uint64_t> h_off;
    std::map<int, IndexArray> index;   // per order K (built on demand: generic path, get_index, count_kmers)
    std::mutex mu;
    int sm_count = 148;
    // 2-bit packed path (A == 4): per-sequence kind (0 irregular, 1 regular, 2 regular with the structural N)
    std::vector<uint8_t> h_kind;
    uint8_t* d_kind = nullptr;
    PackedSeq* d_pseq = nullptr;
    uint32_t* d_words = nullptr;        // 16 bases per word
    uint64_t nwords = 0, nregular = 0;
    std::map<int, uint16_t*> ypatch;   // per order K: [nseq][K+1] k-mer index at mid..mid+K
    // sets encoded from FASTA text on the device: stored positions of the forward undefined bases, until the patches arrive
    uint64_t* d_zero_pos = nullptr;
    uint64_t n_zero_fwd = 0;
    // bamm_seqset_create: the uploads run on their own stream while the host checks the offsets and the first kernels start
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_codes = nullptr, ev_patches = nullptr;
};
static void seqset_copy_done(bamm_seqset* s) {      // waits for the uploads and drops the stream
    if (s->copy_stream) { cudaStreamSynchronize(s->copy_stream); cudaStreamDestroy(s->copy_stream); s->copy_stream = nullptr; }
    if (s->ev_codes) { cudaEventDestroy(s->ev_codes); s->ev_codes = nullptr; }
    if (s->ev_patches) { cudaEventDestroy(s->ev_patches); s->ev_patches = nullptr; }
}

struct bamm_em {
    bamm_seqset* ss = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> loop_ev;       // 4 per iteration of the last bamm_em_iterate call: E | M accumulate | reduce+update
    int loop_iters = 0;
    bool own_xbuf = true;
    int W = 0, K = 0, K_bg_model = 0, K_bg = 0, A = 4;
    uint32_t Yn = 0;            // A^(K+1)
    uint32_t nbin = 0;          // W * Yn
    uint64_t nsub = 0, rsize = 0, nseq_global = 0;
    uint64_t model_size = 0, bg_size = 0;
    ModelDims dims;
    bool smem_tables = true;    // generic path: tables fit in shared memory
    int grid_e = 0, grid_m = 0, block = 512;
    size_t smem_e = 0, smem_m = 0;
    // the subset is split into a packed list (regular 4-letter sequences) and a generic list (everything else)
    uint32_t ngen = 0, npk = 0;
    uint32_t* d_gen_ids = nullptr;  uint64_t* d_gen_roff = nullptr;
    uint32_t* d_pk_ids = nullptr;   uint64_t* d_pk_roff = nullptr;
    Plan plan;                  // W, K, Yn for the packed M-step / scoring kernels
    std::vector<GroupPlan> gplans;   // column groups of the packed E-step, one plan per column pass (chosen in set_model)
    std::vector<char> gfast;         // per pass: every group's bit field sits below bit 32 of the window word
    size_t tab_capacity = 0;    // bytes available for the group tables (= opt-in shared memory)
    float* d_tab = nullptr;     // group tables, concatenated; tab_capacity bytes per pass
    size_t tab_passes = 0;      // passes d_tab has room for
    // active list (windows that survive the M-step's fixed-point rounding), one region per E-step warp
    uint32_t nregions = 0;
    ActiveEntry* d_act = nullptr;
    float* d_scale = nullptr;   // 1/normaliser per packed-list sequence (the packed E-step leaves r unnormalised)
    bool r_scaled = true;       // r already holds normalised values
    uint64_t launches = 0;      // kernels launched by this object (bench.py reports them)
    // NVLink peer exchange (multi-GPU): local receive buffer [2][world][nbin+2] + flags, peers' buffers mapped through CUDA IPC
    int peer_rank = 0, peer_world = 0;
    bool peer_attached = false;
    unsigned char* d_peer_local = nullptr;      // slots, then flags
    void* peer_mapped[MAX_PEERS] = {nullptr};   // cudaIpcOpenMemHandle results (to close)
    PeerPtrs peer_ptrs;
    uint32_t peer_epoch = 0;
    unsigned int* d_peer_done = nullptr;
    int m_nc = 0, m_nsplit = 1; // packed M-step: columns per CTA and number of column splits (packed.cuh, "column split")
    MTables m_tab = {1, 0};     // table copies per CTA (orders 0 and 1) and their stride in words
    int grid_pl = 0;            // CTAs of the packed M-step kernels (a multiple of m_nsplit)
    uint32_t *d_act_cnt = nullptr, *d_overflow = nullptr;
    uint64_t* d_reg_off = nullptr;
    uint16_t* d_ypatch = nullptr;   // owned by the seqset
    int grid_pe = 0, block_pe = 1024;
    size_t smem_pe = 0;
    uint32_t nparts = 1;
    // device
    uint32_t* d_seq_ids = nullptr;
    uint64_t* d_r_off = nullptr;
    float* d_r = nullptr;
    float* d_s = nullptr;         // [j][y]
    float* d_sT = nullptr;        // the same table in the reference's [y][j] order (row gathers of the patched k-mers)
    float* d_v = nullptr;         // all orders
    float* d_vK_prev = nullptr;
    float* d_n = nullptr;         // all orders (float, reference layout)
    float* d_vbg = nullptr;
    float* d_alpha = nullptr;
    unsigned long long* d_part = nullptr;   // per-CTA partial count tables
    unsigned long long* d_xbuf = nullptr;   // [nbin] counts + [2] scalars  (the multi-GPU exchange buffer)
    float* d_vdiff = nullptr;
    double* d_vdiff_part = nullptr;         // per-CTA partial sums of the clustered model update
    // EM::mask workspaces (created by the first bamm_em_mask call)
    uint32_t* d_m_ids = nullptr; uint64_t* d_m_roff = nullptr; uint64_t* d_m_woff = nullptr;
    uint64_t* d_m_seloff = nullptr; uint32_t* d_m_sel = nullptr;
    std::vector<uint32_t> h_ids;            // subset -> seqset index (all sequences of the subset, in order)
    bool whole_set = false;                 // identity subset of an all-regular set: h_ids / h_r_off are made on first use (host_lists)
    // host (pinned)
    unsigned long long* h_scal = nullptr;   // 2 scalars
    float* h_vdiff = nullptr;
    std::vector<uint64_t> h_r_off;
    float q = 0.3f;
    float llh = 0.0f;
    bool model_set = false, s_valid = false, r_valid = false;
    float t_e = 0, t_m = 0;
};

// ------------------------------------------------------------------------------------------- library / device
extern "C" int bamm_version(void) { return 10000 * 0 + 100 * 1 + 0; }
extern "C" const char* bamm_last_error(void) { return g_err; }
extern "C" int bamm_device_count(int* count) {
    REQUIRE(count, "count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return fail(BAMM_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    return BAMM_OK;
}
extern "C" int bamm_set_device(int device) { CU(cudaSetDevice(device)); return BAMM_OK; }
extern "C" int bamm_device_info(int* sm_count, int* cc_major, int* cc_minor, uint64_t* total_mem) {
    int dev; CU(cudaGetDevice(&dev));
    cudaDeviceProp p; CU(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = (uint64_t)p.totalGlobalMem;
    return BAMM_OK;
}

// ------------------------------------------------------------------------------------------- seqset
extern "C" void bamm_seqset_destroy(bamm_seqset* s);
// host bookkeeping + device allocation of codes / offsets (codes are filled by the caller: H2D copy or a device kernel)
// upload_codes: host codes to copy (bamm_seqset_create); the copy is started first, on the set's copy stream, and the host
// pass over the offsets runs while the bases travel. ev_codes marks codes + offsets in place.
// own_offsets: the vector `offsets` points into; it is moved into the set instead of copied (10^7 records of a negative set)
static int seqset_new(const uint64_t* offsets, uint64_t nseq, int A, bamm_seqset** out, const uint8_t* upload_codes = nullptr,
                      std::vector<uint64_t>* own_offsets = nullptr) {
    REQUIRE(out, "out is NULL");
    *out = nullptr;
    REQUIRE(offsets, "offsets is NULL");
    REQUIRE(A >= 2 && A <= 6, "alphabet size %d not in [2,6]", A);
    REQUIRE(offsets[0] == 0, "offsets[0] must be 0");
    REQUIRE(nseq < (1ull << 32), "too many sequences");
    if (!upload_codes)
        for (uint64_t n = 0; n < nseq; n++) REQUIRE(offsets[n + 1] >= offsets[n], "offsets not monotone at %llu", (unsigned long long)n);
    const uint64_t npos = offsets[nseq];
    bamm_seqset* s = new (std::nothrow) bamm_seqset();
    if (!s) return fail(BAMM_E_NOMEM, "host allocation failed");
    int dev; cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { delete s; return fail(BAMM_E_CUDA, "no CUDA device: %s", cudaGetErrorString(e)); }
    s->device = dev; s->A = A; s->nseq = nseq; s->npos = npos; s->npatch = 0;
    cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, dev);
#define CUS(call) do { cudaError_t e2_ = (call); if (e2_ != cudaSuccess) { bamm_seqset_destroy(s); \
    return fail(e2_ == cudaErrorMemoryAllocation ? BAMM_E_NOMEM : BAMM_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e2_)); } } while (0)
    CUS(dev_malloc(&s->d_codes, npos ? npos : 1));
    CUS(dev_malloc(&s->d_off, (nseq + 1) * sizeof(uint64_t)));
    if (upload_codes) {
        CUS(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        CUS(cudaEventCreateWithFlags(&s->ev_codes, cudaEventDisableTiming));
        CUS(cudaEventCreateWithFlags(&s->ev_patches, cudaEventDisableTiming));
        CUS(cudaMemcpyAsync(s->d_off, offsets, (nseq + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s->copy_stream));
        if (npos) CUS(cudaMemcpyAsync(s->d_codes, upload_codes, npos, cudaMemcpyHostToDevice, s->copy_stream));
        CUS(cudaEventRecord(s->ev_codes, s->copy_stream));
    } else {
        CUS(cudaMemcpy(s->d_off, offsets, (nseq + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
    }
    uint64_t maxL = 0, minL = ~0ull, bad = ~0ull;
    for (uint64_t n = 0; n < nseq; n++) {
        if (offsets[n + 1] < offsets[n]) { bad = n; break; }
        uint64_t L = offsets[n + 1] - offsets[n];
        if (L > maxL) maxL = L;
        if (L < minL) minL = L;
    }
    if (bad != ~0ull) { bamm_seqset_destroy(s); return fail(BAMM_E_INVALID, "offsets not monotone at %llu", (unsigned long long)bad); }
    s->maxL = maxL; s->minL = nseq ? minL : 0;
    if (own_offsets) s->h_off.swap(*own_offsets);
    else s->h_off.assign(offsets, offsets + nseq + 1);
    *out = s;
    return BAMM_OK;
}

// classification + 2-bit packing on the device (no host pass over the bases); d_codes and the patch list are in place
// known_regular: the caller vouches that every code is in 1..4 (a set the library sampled itself): no classification pass
// validate_patches: the patch list came from the caller and is checked first (strictly increasing, inside the set)
static int seqset_finish(bamm_seqset* s, bool known_regular = false, bool validate_patches = false) {
    const uint64_t nseq = s->nseq, npatch = s->npatch;
    Trace tr("seqset_finish");
    s->h_kind.assign(nseq, s->A == 4 && known_regular ? 1 : 0);
    const bool classify = s->A == 4 && nseq && !known_regular;
    uint32_t* d_cover0 = nullptr;
    if (classify) {
        CUS(dev_malloc(&s->d_kind, nseq));
        CUS(dev_malloc(&d_cover0, nseq * sizeof(uint32_t)));
    }
    if (s->ev_codes) { cudaError_t ew = cudaStreamWaitEvent(0, s->ev_codes, 0); if (ew != cudaSuccess) { cudaFree(d_cover0); CUS(ew); } }
    if (classify) {       // needs the bases only: runs while the patch list is still on its way
        cudaMemsetAsync(d_cover0, 0, nseq * sizeof(uint32_t), 0);
        k_classify<<<s->sm_count * 8, 256>>>(s->d_codes, s->d_off, nseq, s->d_kind);
    }
    if (s->ev_patches) { cudaError_t ew = cudaStreamWaitEvent(0, s->ev_patches, 0); if (ew != cudaSuccess) { cudaFree(d_cover0); CUS(ew); } }
    if (validate_patches && npatch) {
        uint32_t* d_bad = nullptr; uint32_t bad = 0;
        cudaError_t ev = dev_malloc(&d_bad, sizeof(uint32_t));
        if (ev == cudaSuccess) ev = cudaMemsetAsync(d_bad, 0, sizeof(uint32_t), 0);
        if (ev == cudaSuccess) {
            k_validate_patches<<<(unsigned)((npatch + 255) / 256), 256>>>(s->d_ppos, npatch, s->npos, d_bad);
            ev = cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
        }
        cudaFree(d_bad);
        if (ev != cudaSuccess || bad) cudaFree(d_cover0);
        CUS(ev);
        if (bad) { bamm_seqset_destroy(s); return fail(BAMM_E_INVALID, bad & 1u ? "patch position out of range" : "patch positions must be strictly increasing"); }
    }
    if (s->A == 4 && nseq && known_regular) {
        CUS(dev_malloc(&s->d_kind, nseq));
        CUS(cudaMemset(s->d_kind, 1, nseq));
    }
    if (s->A == 4 && nseq) {
        uint32_t* d_cover = d_cover0;
        if (!known_regular) {
        if (npatch) k_check_patches<<<(unsigned)((npatch + 255) / 256), 256>>>(s->d_ppos, npatch, s->d_off, nseq, s->d_kind, d_cover);
        k_finish_kinds<<<(unsigned)((nseq + 255) / 256), 256>>>(s->d_off, nseq, d_cover, s->d_kind);
        cudaError_t ec = cudaMemcpy(s->h_kind.data(), s->d_kind, nseq, cudaMemcpyDeviceToHost);
        cudaFree(d_cover);
        CUS(ec);
        }
        tr.mark("classify + kinds D2H");
        // packed-stream layout on the device: word counts -> exclusive scan -> PackedSeq records (no host pass, no upload)
        if (known_regular) s->nregular = nseq;
        else for (uint64_t n = 0; n < nseq; n++) s->nregular += s->h_kind[n] != 0;
        if (s->nregular) {
            unsigned long long* d_wc = nullptr; void* d_tmp = nullptr; size_t tmp_bytes = 0;
            CUS(dev_malloc(&d_wc, (nseq + 1) * 2 * sizeof(unsigned long long)));
            unsigned long long* d_scan = d_wc + nseq + 1;
            CUS(cudaMemset(d_wc + nseq, 0, sizeof(unsigned long long)));            // sentinel: the scan's last entry is the total
            k_word_counts<<<(unsigned)((nseq + 255) / 256), 256>>>(s->d_off, nseq, s->d_kind, d_wc);
            cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_wc, d_scan, (int)(nseq + 1));
            cudaError_t es = dev_malloc(&d_tmp, tmp_bytes ? tmp_bytes : 16);
            if (es == cudaSuccess) es = cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_wc, d_scan, (int)(nseq + 1));
            unsigned long long w = 0;
            if (es == cudaSuccess) es = cudaMemcpy(&w, d_scan + nseq, sizeof(w), cudaMemcpyDeviceToHost);
            if (es == cudaSuccess) es = dev_malloc(&s->d_pseq, nseq * sizeof(PackedSeq));
            if (es == cudaSuccess) {
                k_fill_pseq<<<(unsigned)((nseq + 255) / 256), 256>>>(s->d_off, nseq, s->d_kind, d_scan, s->d_pseq);
                es = cudaGetLastError();
            }
            cudaFree(d_tmp); cudaFree(d_wc);
            CUS(es);
            s->nwords = w;
            tr.mark("layout scan");
            CUS(dev_malloc(&s->d_words, (w + 16) * sizeof(uint32_t)));
            tr.mark("words alloc");    // slack: the rolling fetch of the last sequence runs a few words ahead
            k_pack<<<s->sm_count * 8, 256>>>(s->d_codes, s->d_off, nseq, s->d_kind, s->d_pseq, s->d_words);
            CUS(cudaGetLastError());
            CUS(cudaDeviceSynchronize());
            tr.mark("pack kernel");
        }
    }
    return BAMM_OK;
}

extern "C" int bamm_seqset_create(const uint8_t* codes, const uint64_t* offsets, uint64_t nseq, int A,
                                  const uint64_t* patch_pos, const uint64_t* patch_kmer, uint64_t npatch,
                                  bamm_seqset** out) {
    REQUIRE(out, "out is NULL");
    *out = nullptr;
    REQUIRE(codes && offsets, "codes/offsets is NULL");
    REQUIRE(npatch == 0 || (patch_pos && patch_kmer), "patch arrays are NULL");
    bamm_seqset* s = nullptr;
    Trace tr("seqset_create");
    // uploads on the copy stream: offsets, bases, then the patch list; the host pass over the offsets, the classification of
    // the bases and the patch checks (on the device: 11 entries per both-strand sequence) start as soon as their input is there
    { int rc = seqset_new(offsets, nseq, A, &s, codes); if (rc) return rc; }
    tr.mark("alloc + host checks (uploads in flight)");
    s->npatch = npatch;
    if (npatch) {
        CUS(dev_malloc(&s->d_ppos, npatch * sizeof(uint64_t)));
        CUS(dev_malloc(&s->d_pkmer, npatch * sizeof(uint64_t)));
        CUS(cudaMemcpyAsync(s->d_ppos, patch_pos, npatch * sizeof(uint64_t), cudaMemcpyHostToDevice, s->copy_stream));
        CUS(cudaMemcpyAsync(s->d_pkmer, patch_kmer, npatch * sizeof(uint64_t), cudaMemcpyHostToDevice, s->copy_stream));
    }
    CUS(cudaEventRecord(s->ev_patches, s->copy_stream));
    { int rc = seqset_finish(s, false, true); if (rc) return rc; }
    seqset_copy_done(s);                           // the caller's buffers are free again
    tr.mark("uploads + classify + pack");
    *out = s;
    return BAMM_OK;
}
#undef CUS

extern "C" void bamm_seqset_destroy(bamm_seqset* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    for (auto& kv : s->index) cudaFree(kv.second.d);
    for (auto& kv : s->ypatch) cudaFree(kv.second);
    cudaFree(s->d_kind); cudaFree(s->d_pseq); cudaFree(s->d_words); cudaFree(s->d_zero_pos);
    seqset_copy_done(s);
    cudaFree(s->d_codes); cudaFree(s->d_off); cudaFree(s->d_ppos); cudaFree(s->d_pkmer);
    delete s;
}

extern "C" int bamm_seqset_info(const bamm_seqset* s, uint64_t* nseq, uint64_t* npos, int* A) {
    REQUIRE(s, "seqset is NULL");
    if (nseq) *nseq = s->nseq;
    if (npos) *npos = s->npos;
    if (A) *A = s->A;
    return BAMM_OK;
}

static int seqset_index_locked(bamm_seqset* s, int K, IndexArray** out) {
    REQUIRE(K >= 0 && K <= 10, "order K=%d not in [0,10]", K);   // the reference hashes at most 11-mers (Sequence.cpp:36)
    auto it = s->index.find(K);
    if (it != s->index.end()) { if (out) *out = &it->second; return BAMM_OK; }
    const uint64_t Yn = ipow_u64((uint64_t)s->A, K + 1);
    REQUIRE(Yn <= (1ull << 31), "A^(K+1) too large");
    IndexArray ia; ia.Yn = Yn; ia.bytes = (Yn <= 65536) ? 2 : 4;
    CU(cudaSetDevice(s->device));
    CU(dev_malloc(&ia.d, (s->npos ? s->npos : 1) * (uint64_t)ia.bytes));
    const int block = 256;
    const int grid = s->sm_count * 8;
    if (s->nseq) {
        if (ia.bytes == 2) k_build_index<uint16_t><<<grid, block>>>(s->d_codes, s->d_off, s->nseq, s->A, K, Yn, (uint16_t*)ia.d);
        else               k_build_index<uint32_t><<<grid, block>>>(s->d_codes, s->d_off, s->nseq, s->A, K, Yn, (uint32_t*)ia.d);
    }
    if (s->npatch) {
        const int pg = (int)((s->npatch + 255) / 256);
        if (ia.bytes == 2) k_patch_index<uint16_t><<<pg, 256>>>(s->d_ppos, s->d_pkmer, s->npatch, Yn, (uint16_t*)ia.d);
        else               k_patch_index<uint32_t><<<pg, 256>>>(s->d_ppos, s->d_pkmer, s->npatch, Yn, (uint32_t*)ia.d);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { cudaFree(ia.d); return fail(BAMM_E_CUDA, "index build failed: %s", cudaGetErrorString(e)); }
    auto ins = s->index.emplace(K, ia);
    if (out) *out = &ins.first->second;
    return BAMM_OK;
}

extern "C" int bamm_seqset_index(bamm_seqset* s, int K) {
    REQUIRE(s, "seqset is NULL");
    std::lock_guard<std::mutex> g(s->mu);
    return seqset_index_locked(s, K, nullptr);
}

// per-order k-mer index at the K+1 positions after the structural N of every kind-2 sequence
static int seqset_ypatch_locked(bamm_seqset* s, int K, uint16_t** out) {
    auto it = s->ypatch.find(K);
    if (it != s->ypatch.end()) { *out = it->second; return BAMM_OK; }
    const uint64_t Yn = ipow_u64(4, K + 1);
    REQUIRE(Yn <= 65536, "order too high for the packed path");
    uint16_t* d = nullptr;
    CU(cudaSetDevice(s->device));
    const uint64_t bytes = (s->nseq ? s->nseq : 1) * (uint64_t)(K + 1) * sizeof(uint16_t);
    CU(dev_malloc(&d, bytes));
    CU(cudaMemset(d, 0, bytes));
    if (s->npatch) {
        k_make_ypatch<<<(unsigned)((s->npatch + 255) / 256), 256>>>(s->d_ppos, s->d_pkmer, s->npatch, s->d_off, s->nseq, s->d_kind, K, Yn, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { cudaFree(d); return fail(BAMM_E_CUDA, "ypatch build failed: %s", cudaGetErrorString(e)); }
    }
    s->ypatch[K] = d;
    *out = d;
    return BAMM_OK;
}

extern "C" int bamm_seqset_get_index(bamm_seqset* s, int K, uint32_t* out) {
    REQUIRE(s && out, "NULL argument");
    IndexArray* ia;
    { std::lock_guard<std::mutex> g(s->mu); int rc = seqset_index_locked(s, K, &ia); if (rc) return rc; }
    CU(cudaSetDevice(s->device));
    if (ia->bytes == 4) { CU(cudaMemcpy(out, ia->d, s->npos * 4, cudaMemcpyDeviceToHost)); return BAMM_OK; }
    std::vector<uint16_t> tmp(s->npos);
    CU(cudaMemcpy(tmp.data(), ia->d, s->npos * 2, cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < s->npos; i++) out[i] = tmp[i];
    return BAMM_OK;
}

extern "C" int bamm_seqset_count_kmers(bamm_seqset* s, int K, uint64_t* n_all) {
    REQUIRE(s && n_all, "NULL argument");
    IndexArray* ia;
    { std::lock_guard<std::mutex> g(s->mu); int rc = seqset_index_locked(s, K, &ia); if (rc) return rc; }
    CU(cudaSetDevice(s->device));
    unsigned long long* d_cnt;
    CU(dev_malloc(&d_cnt, ia->Yn * 8));
    CU(cudaMemset(d_cnt, 0, ia->Yn * 8));
    if (s->npos) {
        const uint32_t Yn = (uint32_t)ia->Yn;
        const bool smem = ia->Yn <= 12288 && s->npos / ((uint64_t)s->sm_count * 8) < 0xffffffffull;     // 48 KB of 32-bit bins
        const size_t sh = smem ? (size_t)Yn * 4 : 0;
        const int grid = s->sm_count * 8;
        if (ia->bytes == 2) {
            if (smem) k_count_kmers<uint16_t, true><<<grid, 256, sh>>>((const uint16_t*)ia->d, s->npos, Yn, d_cnt);
            else      k_count_kmers<uint16_t, false><<<grid, 256>>>((const uint16_t*)ia->d, s->npos, Yn, d_cnt);
        } else {
            if (smem) k_count_kmers<uint32_t, true><<<grid, 256, sh>>>((const uint32_t*)ia->d, s->npos, Yn, d_cnt);
            else      k_count_kmers<uint32_t, false><<<grid, 256>>>((const uint32_t*)ia->d, s->npos, Yn, d_cnt);
        }
    }
    std::vector<uint64_t> top(ia->Yn);
    cudaError_t e = cudaMemcpy(top.data(), d_cnt, ia->Yn * 8, cudaMemcpyDeviceToHost);
    cudaFree(d_cnt);
    if (e != cudaSuccess) return fail(BAMM_E_CUDA, "k-mer count failed: %s", cudaGetErrorString(e));
    // every position contributes once to every order and y_{k-1} = y_k % A^k, so lower orders are folds
    std::vector<uint64_t> off(K + 2, 0);
    for (int k = 0; k <= K; k++) off[k + 1] = off[k] + ipow_u64(s->A, k + 1);
    memset(n_all, 0, off[K + 1] * sizeof(uint64_t));
    memcpy(n_all + off[K], top.data(), ia->Yn * 8);
    for (int k = K; k > 0; k--) {
        const uint64_t Yk = ipow_u64(s->A, k), Yk1 = Yk * s->A;
        for (uint64_t y = 0; y < Yk1; y++) n_all[off[k - 1] + y % Yk] += n_all[off[k] + y];
    }
    return BAMM_OK;
}

// ------------------------------------------------------------------------------------------- EM
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
    cudaSetDevice(em->device);
    if (em->stream) cudaStreamSynchronize(em->stream);
    for (int p = 0; p < MAX_PEERS; p++) if (em->peer_mapped[p]) cudaIpcCloseMemHandle(em->peer_mapped[p]);
    cudaFree(em->d_peer_local); cudaFree(em->d_peer_done);
    cudaFree(em->d_act); cudaFree(em->d_scale); cudaFree(em->d_act_cnt); cudaFree(em->d_overflow); cudaFree(em->d_reg_off);
    cudaFree(em->d_gen_ids); cudaFree(em->d_gen_roff); cudaFree(em->d_pk_ids); cudaFree(em->d_pk_roff); cudaFree(em->d_tab);
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

static int estep_packed_dispatch(bamm_em* em, const PackedView* pv, size_t pass, bool optin_only);
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

extern "C" int bamm_em_create(bamm_seqset* s, const uint64_t* subset, uint64_t nsub, int W, int K, int K_bg_model,
                              bamm_em** out) {
    REQUIRE(out, "out is NULL");
    *out = nullptr;
    REQUIRE(s, "seqset is NULL");
    REQUIRE(W >= 1 && W <= 32, "motif width W=%d not in [1,32]", W);
    REQUIRE(K >= 0 && K <= 10 && K_bg_model >= 0 && K_bg_model <= 10, "order out of range");
    Trace tr("em_create");
    if (!subset) nsub = s->nseq;
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
    if (getenv("BAMM_TABLE_BYTES")) em->tab_capacity = std::min(em->tab_capacity, (size_t)atol(getenv("BAMM_TABLE_BYTES")));
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
                em->m_tab.nrep = nrep;
                em->m_tab.rstride = nrep > 1 ? ((nb + 30) / 32) * 32 + 1 : nb;
            }
            // the high table sums at most 257 per sequence and bin (the posteriors of a sequence sum to <= 1)
            if ((uint64_t)em->npk / (uint64_t)(em->grid_pl / em->m_nsplit) >= (1ull << 23)) {
                fail(BAMM_E_INVALID, "too many sequences for one device"); bamm_em_destroy(em); return BAMM_E_INVALID;
            }
            if (mstep_w_dispatch(em, nullptr, nullptr, 0)) { fail(BAMM_E_CUDA, "cannot opt in to shared memory for the packed M-step"); bamm_em_destroy(em); return BAMM_E_CUDA; }
            if ((uint32_t)em->grid_pl > em->nparts) em->nparts = (uint32_t)em->grid_pl;
        }
        tr.mark("M-step geometry + opt-in");
        // active list: one region per E-step warp, sized as a fraction of the warp's windows (BAMM_LIST_FRAC, 0 = off)
        double frac = getenv("BAMM_LIST_FRAC") ? atof(getenv("BAMM_LIST_FRAC")) : 0.5;
        std::vector<uint64_t> reg;
        if (frac > 0.0) {
            em->nregions = (uint32_t)em->grid_pe * (uint32_t)(em->block_pe / 32);
            std::vector<uint64_t> win(em->nregions, 0);
            reg.assign((size_t)em->nregions + 1, 0);
            if (em->whole_set && s->minL == s->maxL) {          // equal lengths: sequence i goes to warp i % nregions
                const uint64_t lw1 = s->maxL - (uint64_t)W + 1, per = em->npk / em->nregions, extra = em->npk % em->nregions;
                for (uint32_t w = 0; w < em->nregions; w++) win[w] = (per + (w < extra ? 1 : 0)) * lw1;
            } else {
                uint32_t w = 0;
                for (size_t i = 0; i < em->npk; i++) {
                    const uint64_t n = em->whole_set ? i : pk_ids[i];
                    win[w] += s->h_off[n + 1] - s->h_off[n] - (uint64_t)W + 1;
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

static int launch_tuple_table(bamm_em* em) {
    if (!em->npk) return BAMM_OK;
    for (size_t i = 0; i < em->gplans.size(); i++) {
        const uint32_t total = em->gplans[i].table_bytes >> 2;
        const uint32_t blocks = (total + 255) / 256;
        k_make_group_tables<false><<<blocks < 1184 ? blocks : 1184, 256, 0, em->stream>>>(em->d_s, em->gplans[i], (float*)((char*)em->d_tab + i * em->tab_capacity));
        CU(cudaGetLastError());
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
    REQUIRE(q > 0.0f && q < 1.0f, "q=%g not in (0,1)", (double)q);
    CU(cudaSetDevice(em->device));
    Trace tr("set_model");
    if (em->npk) {
        const bool reduced = leading_columns_are_copies(em->dims, em->K, em->W, em->Yn, v_all) && !getenv("BAMM_NO_REDUCED");
        if (!plan_passes(em->W, em->K, em->K_bg, reduced, em->tab_capacity, em->gplans, em->gfast))
            return fail(BAMM_E_STATE, "no column-group plan fits shared memory");
        if (em->gplans.size() > em->tab_passes) {
            CU(cudaStreamSynchronize(em->stream));
            cudaFree(em->d_tab); em->d_tab = nullptr; em->tab_passes = 0;
            CU(dev_malloc(&em->d_tab, em->gplans.size() * em->tab_capacity));
            em->tab_passes = em->gplans.size();
        }
        for (size_t i = 0; i < em->gplans.size(); i++)
            if (estep_packed_dispatch(em, nullptr, i, true)) return fail(BAMM_E_CUDA, "cannot opt in to %u bytes of shared memory", em->gplans[i].table_bytes);
    }
    tr.mark("plan + table buffer + opt-in");
    CU(cudaMemcpyAsync(em->d_v, v_all, em->model_size * sizeof(float), cudaMemcpyHostToDevice, em->stream));
    CU(cudaMemcpyAsync(em->d_vbg, vbg_all, em->bg_size * sizeof(float), cudaMemcpyHostToDevice, em->stream));
    CU(cudaMemcpyAsync(em->d_alpha, alpha, (uint64_t)(em->K + 1) * em->W * sizeof(float), cudaMemcpyHostToDevice, em->stream));
    k_make_s<<<64, 256, 0, em->stream>>>(em->dims, em->d_v, em->d_vbg, em->d_s, em->d_sT, em->d_vK_prev);
    CU(cudaGetLastError());
    { int rc = launch_tuple_table(em); if (rc) return rc; }
    CU(cudaStreamSynchronize(em->stream));
    tr.mark("uploads + s + group tables");
    em->q = q; em->model_set = true; em->s_valid = true; em->r_valid = false; em->llh = 0.0f;
    return BAMM_OK;
}

// k_estep_packed is instantiated for every group count the planner can choose, in both extraction modes;
// optin_only sets the shared-memory attribute instead of launching.
static ActiveList alist_of(const bamm_em* em) {
    ActiveList al; al.ent = em->d_act; al.scale = em->d_scale; al.reg_off = em->d_reg_off;
    al.cnt = em->d_act_cnt; al.cnt_back = em->d_act_cnt ? em->d_act_cnt + em->nregions : nullptr; al.overflow = em->d_overflow;
    return al;
}
template <int G, bool FAST, bool MULTI> static int estep_packed_one(bamm_em* em, const PackedView* pv, size_t pass, bool optin_only) {
    GroupPlan gp = em->gplans[pass]; gp.q = em->q;
    gp.thr0 = FX_HALF_UNIT * (1.0f - em->q) * 0.999f;
    if (optin_only) return max_smem_optin(k_estep_packed<G, FAST, MULTI>, gp.table_bytes);
    k_estep_packed<G, FAST, MULTI><<<em->grid_pe, em->block_pe, gp.table_bytes, em->stream>>>(*pv, gp, (const float*)((const char*)em->d_tab + pass * em->tab_capacity),
                                                                                       em->d_s, em->d_sT, em->d_r, em->d_xbuf + em->nbin, alist_of(em));
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
template <int G> static int estep_packed_g(bamm_em* em, const PackedView* pv, size_t pass, bool optin_only) {
    const bool multi = em->gplans.size() > 1;
    if (em->gfast[pass]) return multi ? estep_packed_one<G, true, true>(em, pv, pass, optin_only) : estep_packed_one<G, true, false>(em, pv, pass, optin_only);
    return multi ? estep_packed_one<G, false, true>(em, pv, pass, optin_only) : estep_packed_one<G, false, false>(em, pv, pass, optin_only);
}
static int estep_packed_dispatch(bamm_em* em, const PackedView* pv, size_t pass, bool optin_only) {
    switch (em->gplans[pass].G) {
#define BAMM_CASE(g) case g: return estep_packed_g<g>(em, pv, pass, optin_only);
        BAMM_CASE(1) BAMM_CASE(2) BAMM_CASE(3) BAMM_CASE(4) BAMM_CASE(5) BAMM_CASE(6) BAMM_CASE(7) BAMM_CASE(8)
        BAMM_CASE(9) BAMM_CASE(10) BAMM_CASE(11) BAMM_CASE(12) BAMM_CASE(13) BAMM_CASE(14) BAMM_CASE(15) BAMM_CASE(16)
#undef BAMM_CASE
        default: return -1;
    }
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

static int launch_estep(bamm_em* em) {
    em->launches += (em->npk ? em->gplans.size() : 0) + (em->ngen ? 1 : 0);
    unsigned long long* scal = em->d_xbuf + em->nbin;
    CU(cudaMemsetAsync(scal, 0, 2 * sizeof(unsigned long long), em->stream));
    if (em->npk) {
        PackedView pv = pview_of(em);
        if (em->d_overflow) CU(cudaMemsetAsync(em->d_overflow, 0, 4, em->stream));
        for (size_t pass = 0; pass < em->gplans.size(); pass++)
            if (estep_packed_dispatch(em, &pv, pass, false)) return fail(BAMM_E_CUDA, "packed E-step launch failed");
        em->r_scaled = false;
        CU(cudaGetLastError());
    }
    if (em->ngen) {
        IndexArray& ia = em->ss->index[em->K];
        SubsetView sv = view_of(em);
        if (ia.bytes == 2) {
            if (em->smem_tables) k_estep<uint16_t, true><<<em->grid_e, em->block, em->smem_e, em->stream>>>((const uint16_t*)ia.d, sv, em->W, em->Yn, em->d_s, em->q, em->d_r, scal);
            else                 k_estep<uint16_t, false><<<em->grid_e, em->block, 0, em->stream>>>((const uint16_t*)ia.d, sv, em->W, em->Yn, em->d_s, em->q, em->d_r, scal);
        } else {
            if (em->smem_tables) k_estep<uint32_t, true><<<em->grid_e, em->block, em->smem_e, em->stream>>>((const uint32_t*)ia.d, sv, em->W, em->Yn, em->d_s, em->q, em->d_r, scal);
            else                 k_estep<uint32_t, false><<<em->grid_e, em->block, 0, em->stream>>>((const uint32_t*)ia.d, sv, em->W, em->Yn, em->d_s, em->q, em->d_r, scal);
        }
        CU(cudaGetLastError());
    }
    return BAMM_OK;
}

// packed M-step kernels: one instantiation per column count of a CTA. mode 0: opt in to the shared memory of both kernels,
// 1: launch the list kernel, 2: launch the scan kernel (conditional on the list's overflow flag when there is a list)
template <int NC> static int mstep_w_one(bamm_em* em, const PackedView* pv, const Plan* pl, int mode) {
    const size_t smem = (size_t)2 * em->m_tab.nrep * em->m_tab.rstride * 4;
    if (mode == 0) return max_smem_optin(k_mstep_list_w<NC>, smem) | max_smem_optin(k_mstep_scan_w<NC>, smem);
    if (mode == 1) k_mstep_list_w<NC><<<em->grid_pl, 1024, smem, em->stream>>>(*pv, *pl, alist_of(em), em->nregions, em->m_nsplit, em->m_tab, em->d_part);
    else k_mstep_scan_w<NC><<<em->grid_pl, 1024, smem, em->stream>>>(*pv, *pl, em->d_r, em->r_scaled ? nullptr : em->d_scale,
                                                                    em->d_act ? em->d_overflow : nullptr, em->m_nsplit, em->m_tab, em->d_part);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
static int mstep_w_dispatch(bamm_em* em, const PackedView* pv, const Plan* pl, int mode) {
    switch (em->m_nc) {
#define BAMM_CASE(w) case w: return mstep_w_one<w>(em, pv, pl, mode);
        BAMM_CASE(1) BAMM_CASE(2) BAMM_CASE(3) BAMM_CASE(4) BAMM_CASE(5) BAMM_CASE(6) BAMM_CASE(7) BAMM_CASE(8)
        BAMM_CASE(9) BAMM_CASE(10) BAMM_CASE(11) BAMM_CASE(12) BAMM_CASE(13) BAMM_CASE(14) BAMM_CASE(15) BAMM_CASE(16)
        BAMM_CASE(17) BAMM_CASE(18) BAMM_CASE(19) BAMM_CASE(20) BAMM_CASE(21) BAMM_CASE(22) BAMM_CASE(23) BAMM_CASE(24)
        BAMM_CASE(25) BAMM_CASE(26) BAMM_CASE(27) BAMM_CASE(28) BAMM_CASE(29) BAMM_CASE(30) BAMM_CASE(31) BAMM_CASE(32)
#undef BAMM_CASE
        default: return -1;
    }
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
        }
        // the E-step listed the windows that matter; the scan kernel only does work (device-side decision) if a region
        // overflowed, or when there is no list
        if (em->d_act && mstep_w_dispatch(em, &pv, &pl, 1)) return fail(BAMM_E_CUDA, "list M-step launch failed");
        if (mstep_w_dispatch(em, &pv, &pl, 2)) return fail(BAMM_E_CUDA, "scan M-step launch failed");
    }
    if (em->ngen) {
        IndexArray& ia = em->ss->index[em->K];
        SubsetView sv = view_of(em);
        if (ia.bytes == 2) {
            if (em->smem_tables) k_mstep<uint16_t, true><<<em->grid_m, em->block, em->smem_m, em->stream>>>((const uint16_t*)ia.d, sv, em->W, em->Yn, em->d_r, em->d_part);
            else                 k_mstep<uint16_t, false><<<em->grid_m, em->block, 0, em->stream>>>((const uint16_t*)ia.d, sv, em->W, em->Yn, em->d_r, em->d_part);
        } else {
            if (em->smem_tables) k_mstep<uint32_t, true><<<em->grid_m, em->block, em->smem_m, em->stream>>>((const uint32_t*)ia.d, sv, em->W, em->Yn, em->d_r, em->d_part);
            else                 k_mstep<uint32_t, false><<<em->grid_m, em->block, 0, em->stream>>>((const uint32_t*)ia.d, sv, em->W, em->Yn, em->d_r, em->d_part);
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
        k_reduce_push<<<(words + 255) / 256, 256, 0, em->stream>>>(em->d_part, em->nparts, em->nbin, em->d_xbuf + em->nbin, em->peer_ptrs,
                                                                   em->peer_rank, em->peer_world, parity, em->peer_epoch, em->d_peer_done);
        CU(cudaGetLastError());
        const size_t slot_bytes = (size_t)2 * em->peer_world * words * sizeof(unsigned long long);
        k_peer_sum<<<(words + 255) / 256, 256, 0, em->stream>>>((const unsigned long long*)em->d_peer_local, (const unsigned int*)(em->d_peer_local + slot_bytes),
                                                                em->peer_world, em->nbin, parity, em->peer_epoch, em->d_xbuf);
        CU(cudaGetLastError());
        em->launches += 1;
        return BAMM_OK;
    }
    k_reduce_parts<<<(em->nbin + 255) / 256, 256, 0, em->stream>>>(em->d_part, em->nparts, em->nbin, em->d_xbuf);
    CU(cudaGetLastError());
    return BAMM_OK;
}

static int launch_mstep_local(bamm_em* em) {
    int rc = launch_mstep_accumulate(em); if (rc) return rc;
    return launch_mstep_reduce(em);
}

static int launch_update(bamm_em* em) {
    em->launches += 1 + (em->npk ? 1 : 0);
    // large tables: one thread-block cluster of 8 CTAs instead of one CTA
    if ((uint64_t)em->nbin >= 16384 && em->d_vdiff_part && !getenv("BAMM_NO_CLUSTER_UPDATE"))
        k_update_model_cluster<<<UPDATE_CLUSTER, 1024, 0, em->stream>>>(em->dims, em->d_xbuf, em->d_n, em->d_v, em->d_vK_prev, em->d_vbg, em->d_alpha,
                                                                      em->d_s, em->d_sT, em->d_vdiff, em->d_vdiff_part);
    else
        k_update_model<<<1, 1024, 0, em->stream>>>(em->dims, em->d_xbuf, em->d_n, em->d_v, em->d_vK_prev, em->d_vbg, em->d_alpha, em->d_s, em->d_sT, em->d_vdiff);
    CU(cudaGetLastError());
    return launch_tuple_table(em);
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
    if (!em->model_set) return fail(BAMM_E_STATE, "bamm_em_set_model has not been called");
    CU(cudaSetDevice(em->device));
    CU(cudaEventRecord(em->ev[0], em->stream));
    int rc = launch_estep(em); if (rc) return rc;
    CU(cudaEventRecord(em->ev[1], em->stream));
    em->r_valid = true;
    return BAMM_OK;
}

extern "C" int bamm_em_estep(bamm_em* em, float* llh) {
    int rc = bamm_em_estep_local(em); if (rc) return rc;
    rc = read_scalars(em, false); if (rc) return rc;
    if (llh) *llh = em->llh;
    return BAMM_OK;
}

extern "C" int bamm_em_mstep_local(bamm_em* em) {
    REQUIRE(em, "em is NULL");
    if (!em->r_valid) return fail(BAMM_E_STATE, "M-step needs the r of an E-step");
    CU(cudaSetDevice(em->device));
    CU(cudaEventRecord(em->ev[2], em->stream));
    return launch_mstep_local(em);
}

extern "C" int bamm_em_finish_iteration(bamm_em* em, int optimize_q, float* llh, float* vdiff) {
    REQUIRE(em, "em is NULL");
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
    REQUIRE(words == (uint64_t)em->nbin + 2, "exchange buffer must hold %llu 64-bit words", (unsigned long long)em->nbin + 2);
    CU(cudaSetDevice(em->device));
    CU(cudaStreamSynchronize(em->stream));
    if (em->own_xbuf) cudaFree(em->d_xbuf);
    em->d_xbuf = (unsigned long long*)dev_ptr; em->own_xbuf = false;
    CU(cudaMemsetAsync(em->d_xbuf, 0, words * 8, em->stream));
    return BAMM_OK;
}

extern "C" int bamm_em_mstep(bamm_em* em) {
    int rc = bamm_em_mstep_local(em); if (rc) return rc;
    rc = launch_update(em); if (rc) return rc;
    CU(cudaEventRecord(em->ev[3], em->stream));
    CU(cudaStreamSynchronize(em->stream));
    return BAMM_OK;
}

extern "C" int bamm_em_optimize_q(bamm_em* em, float* q) {
    REQUIRE(em, "em is NULL");
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
        em->r_valid = true; em->r_scaled = true;
    }
done:
#undef CUX
    cudaFree(d_s0); cudaFree(d_all); cudaFree(d_cnt);
    return rc;
}

extern "C" int bamm_em_mask(bamm_em* em, float f, float epsilon, int max_iter, int* iterations, float* llh, uint64_t* n_kept, float* r_cutoff) {
    REQUIRE(em, "em is NULL");
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

// launches of one iteration with optional event brackets (ev4 = 4 events or nullptr)
static int launch_iteration(bamm_em* em, cudaEvent_t* ev4) {
    if (ev4) CU(cudaEventRecord(ev4[0], em->stream));
    int rc = launch_estep(em); if (rc) return rc;
    if (ev4) CU(cudaEventRecord(ev4[1], em->stream));
    rc = launch_mstep_accumulate(em); if (rc) return rc;
    if (ev4) CU(cudaEventRecord(ev4[2], em->stream));
    rc = launch_mstep_reduce(em); if (rc) return rc;
    rc = launch_update(em); if (rc) return rc;
    if (ev4) CU(cudaEventRecord(ev4[3], em->stream));
    return BAMM_OK;
}

extern "C" int bamm_em_iterate(bamm_em* em, int n_iter, float* llh_last, float* vdiff_last) {
    REQUIRE(em, "em is NULL");
    if (!em->model_set) return fail(BAMM_E_STATE, "bamm_em_set_model has not been called");
    REQUIRE(n_iter >= 0, "n_iter must be >= 0");
    CU(cudaSetDevice(em->device));
    while ((int)em->loop_ev.size() < 4 * n_iter) { cudaEvent_t e; CU(cudaEventCreate(&e)); em->loop_ev.push_back(e); }
    em->loop_iters = 0;
    for (int it = 0; it < n_iter; it++) {
        int rc = launch_iteration(em, &em->loop_ev[4 * it]); if (rc) return rc;
    }
    em->loop_iters = n_iter;
    em->r_valid = n_iter > 0 || em->r_valid;
    int rc = read_scalars(em, true); if (rc) return rc;
    if (llh_last) *llh_last = em->llh;
    if (vdiff_last) *vdiff_last = *em->h_vdiff;
    return BAMM_OK;
}

extern "C" int bamm_em_loop_timing(bamm_em* em, int* iters, float* estep_ms, float* maccum_ms, float* update_ms, float* total_ms) {
    REQUIRE(em, "em is NULL");
    CU(cudaSetDevice(em->device));
    CU(cudaStreamSynchronize(em->stream));
    float e = 0, m = 0, u = 0, t = 0, x;
    for (int it = 0; it < em->loop_iters; it++) {
        cudaEvent_t* ev = &em->loop_ev[4 * it];
        CU(cudaEventElapsedTime(&x, ev[0], ev[1])); e += x;
        CU(cudaEventElapsedTime(&x, ev[1], ev[2])); m += x;
        CU(cudaEventElapsedTime(&x, ev[2], ev[3])); u += x;
    }
    if (em->loop_iters) CU(cudaEventElapsedTime(&t, em->loop_ev[0], em->loop_ev[4 * em->loop_iters - 1]));
    if (iters) *iters = em->loop_iters;
    if (estep_ms) *estep_ms = e;
    if (maccum_ms) *maccum_ms = m;
    if (update_ms) *update_ms = u;
    if (total_ms) *total_ms = t;
    return BAMM_OK;
}

extern "C" int bamm_em_last_timing(bamm_em* em, float* estep_ms, float* mstep_ms) {
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
    REQUIRE(em && v_all, "NULL argument");
    CU(cudaSetDevice(em->device));
    CU(cudaMemcpyAsync(v_all, em->d_v, em->model_size * sizeof(float), cudaMemcpyDeviceToHost, em->stream));
    CU(cudaStreamSynchronize(em->stream));
    return BAMM_OK;
}
extern "C" int bamm_em_get_counts(bamm_em* em, float* n_all) {
    REQUIRE(em && n_all, "NULL argument");
    CU(cudaSetDevice(em->device));
    CU(cudaMemcpyAsync(n_all, em->d_n, em->model_size * sizeof(float), cudaMemcpyDeviceToHost, em->stream));
    CU(cudaStreamSynchronize(em->stream));
    return BAMM_OK;
}
extern "C" int bamm_em_get_s(bamm_em* em, float* s) {
    REQUIRE(em && s, "NULL argument");
    CU(cudaSetDevice(em->device));
    std::vector<float> t(em->nbin);
    CU(cudaMemcpyAsync(t.data(), em->d_s, (uint64_t)em->nbin * sizeof(float), cudaMemcpyDeviceToHost, em->stream));
    CU(cudaStreamSynchronize(em->stream));
    for (uint32_t y = 0; y < em->Yn; y++) for (int j = 0; j < em->W; j++) s[(uint64_t)y * em->W + j] = t[(uint64_t)j * em->Yn + y];
    return BAMM_OK;
}
extern "C" int bamm_em_get_q(bamm_em* em, float* q) { REQUIRE(em && q, "NULL argument"); *q = em->q; return BAMM_OK; }
extern "C" uint64_t bamm_em_r_size(const bamm_em* em) { return em ? em->rsize : 0; }
extern "C" int bamm_em_get_r(bamm_em* em, uint64_t first, uint64_t count, float* out) {
    REQUIRE(em && out, "NULL argument");
    REQUIRE(first + count <= em->nsub, "sequence range out of bounds");
    if (!em->r_valid) return fail(BAMM_E_STATE, "no E-step has run");
    CU(cudaSetDevice(em->device));
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
    *dev_ptr = em->d_xbuf; *words = (uint64_t)em->nbin + 2;
    return BAMM_OK;
}
extern "C" int bamm_em_set_global_nseq(bamm_em* em, uint64_t n) { REQUIRE(em, "em is NULL"); em->nseq_global = n; return BAMM_OK; }
extern "C" int bamm_em_launch_count(bamm_em* em, uint64_t* kernels) { REQUIRE(em && kernels, "NULL argument"); *kernels = em->launches; return BAMM_OK; }
extern "C" int bamm_em_stream(bamm_em* em, void** stream) { REQUIRE(em && stream, "NULL argument"); *stream = (void*)em->stream; return BAMM_OK; }


// ---- NVLink peer exchange ------------------------------------------------------------------------------------------
extern "C" int bamm_em_peer_alloc(bamm_em* em, int rank, int world, void* ipc_handle_out) {
    REQUIRE(em && ipc_handle_out, "NULL argument");
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
    CU(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, em->d_peer_local));
    static_assert(sizeof(h) == 64, "CUDA IPC handle size");
    memcpy(ipc_handle_out, &h, sizeof(h));
    em->peer_rank = rank; em->peer_world = world;
    return BAMM_OK;
}

extern "C" int bamm_em_peer_attach(bamm_em* em, const void* ipc_handles) {
    REQUIRE(em && ipc_handles, "NULL argument");
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

// ------------------------------------------------------------------------------------------- scoring
template <int G, bool FAST>
static int score_zoops_one(const GroupPlan& gp, int sms, cudaStream_t st, const PackedView& pv, const float* d_tab, const float* d_s, float two_eps,
                           float* d_zoops, unsigned long long* d_z, const uint32_t* d_out, size_t plain_bytes) {
    const size_t smem = (size_t)gp.table_bytes + plain_bytes;
    if (cudaFuncSetAttribute(k_score_zoops_packed<G, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    k_score_zoops_packed<G, FAST><<<sms, 1024, smem, st>>>(pv, gp, d_tab, d_s, two_eps, d_zoops, d_z, d_out);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
static int score_zoops_dispatch(const GroupPlan& gp, bool fast, int sms, cudaStream_t st, const PackedView& pv, const float* d_tab, const float* d_s,
                                float two_eps, float* d_zoops, unsigned long long* d_z, const uint32_t* d_out, size_t plain_bytes) {
    switch (gp.G) {
#define BAMM_CASE(g) case g: return fast ? score_zoops_one<g, true>(gp, sms, st, pv, d_tab, d_s, two_eps, d_zoops, d_z, d_out, plain_bytes) \
                                         : score_zoops_one<g, false>(gp, sms, st, pv, d_tab, d_s, two_eps, d_zoops, d_z, d_out, plain_bytes);
        BAMM_CASE(1) BAMM_CASE(2) BAMM_CASE(3) BAMM_CASE(4) BAMM_CASE(5) BAMM_CASE(6) BAMM_CASE(7) BAMM_CASE(8)
        BAMM_CASE(9) BAMM_CASE(10) BAMM_CASE(11) BAMM_CASE(12) BAMM_CASE(13) BAMM_CASE(14) BAMM_CASE(15) BAMM_CASE(16)
#undef BAMM_CASE
        default: return -1;
    }
}

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
                if (score_zoops_dispatch(zplan, zfast, s->sm_count, st, pv, d_ztab, d_s, two_eps, d_zoops, d_z, identity_out ? nullptr : d_pout, tb)) {
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
    REQUIRE(n < (1ull << 31), "too many scores for one sort call");
    float* d_alt = nullptr; void* d_tmp = nullptr; size_t tmp_bytes = 0;
    cudaError_t e = dev_malloc(&d_alt, n * sizeof(float));
    if (e != cudaSuccess) return fail(BAMM_E_NOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e));
    cub::DoubleBuffer<float> buf(d_keys, d_alt);
    if (descending) cub::DeviceRadixSort::SortKeysDescending(nullptr, tmp_bytes, buf, (int)n, 0, 32, st);
    else            cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, buf, (int)n, 0, 32, st);
    e = dev_malloc(&d_tmp, tmp_bytes ? tmp_bytes : 16);
    if (e == cudaSuccess) {
        if (descending) e = cub::DeviceRadixSort::SortKeysDescending(d_tmp, tmp_bytes, buf, (int)n, 0, 32, st);
        else            e = cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, buf, (int)n, 0, 32, st);
    }
    if (e == cudaSuccess && buf.Current() != d_keys) e = cudaMemcpyAsync(d_keys, buf.Current(), n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_alt); cudaFree(d_tmp);
    if (e != cudaSuccess) return fail(BAMM_E_CUDA, "device sort failed: %s", cudaGetErrorString(e));
    return BAMM_OK;
}

extern "C" int bamm_sort_scores(float* scores, uint64_t n, int descending) {
    REQUIRE(scores || n == 0, "scores is NULL");
    if (n < 2) return BAMM_OK;
    float* d = nullptr;
    CU(dev_malloc(&d, n * sizeof(float)));
    cudaError_t e = cudaMemcpy(d, scores, n * sizeof(float), cudaMemcpyHostToDevice);
    int rc = e == cudaSuccess ? device_sort_f32(d, n, descending != 0, 0) : fail(BAMM_E_CUDA, "H2D failed: %s", cudaGetErrorString(e));
    if (!rc) { e = cudaMemcpy(scores, d, n * sizeof(float), cudaMemcpyDeviceToHost); if (e != cudaSuccess) rc = fail(BAMM_E_CUDA, "D2H failed: %s", cudaGetErrorString(e)); }
    cudaFree(d);
    return rc;
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
        const size_t nTop = (size_t)std::min(100, (int)nneg / 10);
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
