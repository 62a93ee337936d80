// microbench.cu — shared-memory lookup / atomic / shuffle throughput on one B200, full occupancy.
// Answers the design questions of DESIGN.md §5: what does a random table lookup, a random native integer
// shared atomic (with / without return), a float CAS-loop atomic and a warp shuffle cost per SM per clock?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench tools/microbench.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t xs(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

enum Mode { LDS_RAND, LDS_CF, LDS128_RAND, LDS64_RAND, ATOMS_NORET_RAND, ATOMS_RET_RAND, ATOMS_NORET_CF, ATOMS_RET_CF, FATOM_RAND, FATOM_CF, SHFL, ALU_ONLY, LDS_RAND_U16, NMODES };
const char* names[] = {"lds32 random", "lds32 conflict-free", "lds128 random(16B)", "lds64 random(8B)", "atoms.add.u32 noret random", "atoms.add.u32 ret random",
                       "atoms.add.u32 noret conflict-free", "atoms.add.u32 ret conflict-free", "atomicAdd(float) CAS random", "atomicAdd(float) CAS conflict-free",
                       "shfl.idx", "alu only (index gen)", "lds.u16 random"};

template <int MODE>
__global__ void __launch_bounds__(512) bench(uint32_t words, int iters, unsigned long long* out, long long* cyc) {
    extern __shared__ uint32_t tbl[];
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) tbl[i] = i * 2654435761u >> 8;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t s = (blockIdx.x * 7919u + threadIdx.x) * 2654435761u + 12345u;
    uint32_t acc = 0; float facc = 0.f;
    const uint32_t mask = words - 1;            // words is a power of two
    uint32_t rr[16];                            // random bases held in registers: 2 ALU ops per generated index
#pragma unroll
    for (int u = 0; u < 16; u++) rr[u] = xs(s);
    long long t0 = clock64();
    for (int it = 0; it < iters; it += 16) {
        const uint32_t step = (uint32_t)it * 40503u;
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const uint32_t r = rr[u] + step;
            if (MODE == LDS_RAND) acc += tbl[r & mask];
            else if (MODE == LDS_RAND_U16) acc += ((uint16_t*)tbl)[r & (2 * words - 1)];
            else if (MODE == LDS_CF) acc += tbl[((r & mask) & ~31u) | lane];
            else if (MODE == LDS128_RAND) { uint4 v = ((uint4*)tbl)[(r & mask) >> 2]; acc += v.x ^ v.y ^ v.z ^ v.w; }
            else if (MODE == LDS64_RAND) { uint2 v = ((uint2*)tbl)[(r & mask) >> 1]; acc += v.x ^ v.y; }
            else if (MODE == ATOMS_NORET_RAND) atomicAdd(&tbl[r & mask], r);
            else if (MODE == ATOMS_RET_RAND) acc += atomicAdd(&tbl[r & mask], r);
            else if (MODE == ATOMS_NORET_CF) atomicAdd(&tbl[((r & mask) & ~31u) | lane], r);
            else if (MODE == ATOMS_RET_CF) acc += atomicAdd(&tbl[((r & mask) & ~31u) | lane], r);
            else if (MODE == FATOM_RAND) atomicAdd((float*)&tbl[r & mask], 1.0f);
            else if (MODE == FATOM_CF) atomicAdd((float*)&tbl[((r & mask) & ~31u) | lane], 1.0f);
            else if (MODE == SHFL) acc += __shfl_sync(0xffffffffu, r, (lane + u) & 31);
            else if (MODE == ALU_ONLY) acc += r & mask;
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0xdeadbeef || facc == 1.5f) out[0] = acc + tbl[acc & mask];
}

template <int MODE> int run(uint32_t words, int blocks_per_sm, int sms, unsigned long long* d_out, long long* d_cyc) {
    const int iters = 1 << 16;
    const int grid = sms * blocks_per_sm;
    size_t smem = (size_t)words * 4;
    CK(cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int w = 0; w < 3; w++) bench<MODE><<<grid, 512, smem>>>(words, iters, d_out, d_cyc);   // warm-up (clock ramp)
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    bench<MODE><<<grid, 512, smem>>>(words, iters, d_out, d_cyc);
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    static long long h_cyc[4096];
    CK(cudaMemcpy(h_cyc, d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0; double mean = 0;
    for (int i = 0; i < grid; i++) { if (h_cyc[i] > mx) mx = h_cyc[i]; mean += h_cyc[i]; }
    mean /= grid;
    // per SM: blocks_per_sm * 512 threads * iters lane-ops in `mean` cycles (blocks on an SM run concurrently)
    double lane_ops_per_clk = (double)blocks_per_sm * 512.0 * iters / mean;
    printf("%-38s table %6u KB  ctas/SM %d  %8.3f ms  lane-ops/clk/SM %7.2f  (wavefront-equiv/clk %5.2f)  clk %.0f MHz\n",
           names[MODE], words * 4 / 1024, blocks_per_sm, ms, lane_ops_per_clk, lane_ops_per_clk / 32.0, mean / (ms * 1e3));
    return 0;
}

int main() {
    int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    printf("device %s  SMs %d  smem optin %zu\n", p.name, p.multiProcessorCount, (size_t)p.sharedMemPerBlockOptin);
    unsigned long long* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, 8)); CK(cudaMalloc(&d_cyc, 4096 * 8));
    int sms = p.multiProcessorCount;
    for (uint32_t words : {4096u, 16384u, 32768u}) {     // 16 KB (4 CTA/SM), 64 KB (3), 128 KB (1)
        int bps = words == 4096 ? 4 : (words == 16384 ? 3 : 1);
        run<ALU_ONLY>(words, bps, sms, d_out, d_cyc);
        run<LDS_RAND>(words, bps, sms, d_out, d_cyc);
        run<LDS_RAND_U16>(words, bps, sms, d_out, d_cyc);
        run<LDS_CF>(words, bps, sms, d_out, d_cyc);
        run<LDS64_RAND>(words, bps, sms, d_out, d_cyc);
        run<LDS128_RAND>(words, bps, sms, d_out, d_cyc);
        run<ATOMS_NORET_RAND>(words, bps, sms, d_out, d_cyc);
        run<ATOMS_RET_RAND>(words, bps, sms, d_out, d_cyc);
        run<ATOMS_NORET_CF>(words, bps, sms, d_out, d_cyc);
        run<ATOMS_RET_CF>(words, bps, sms, d_out, d_cyc);
        run<FATOM_RAND>(words, bps, sms, d_out, d_cyc);
        run<FATOM_CF>(words, bps, sms, d_out, d_cyc);
        run<SHFL>(words, bps, sms, d_out, d_cyc);
    }
    return 0;
}
