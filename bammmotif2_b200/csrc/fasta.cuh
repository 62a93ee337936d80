// fasta.cuh — sequence encoding on the device.
// reference: Sequence::Sequence + appendRevComp (src/init/Sequence.cpp:4-43, 91-99), Alphabet::getCode /
// getComplementCode (src/init/Alphabet.h:35-45, Alphabet.cpp:10-55), the reader SequenceSet::readFASTA
// (src/init/SequenceSet.cpp:67-225). The host keeps what is inherently sequential and tiny — finding the lines of the
// file, the headers, and the libc rand() draws for undefined bases — and ships the raw text; every base is touched only
// here: text byte -> code through the alphabet's 256-entry table, forward | 0 | reverse complement layout, base counts,
// and the code neighbourhoods of the undefined bases that the host needs for its rand()-dependent k-mer hashes.
#pragma once
#include "kernels.cuh"

namespace bamm {

struct FastaSeg {           // one sequence line of the file
    uint64_t text_off;      // first byte in the text
    uint32_t len;           // bytes (bases) of the line
    uint32_t rec;           // record (sequence) index
    uint64_t dst;           // bases of the record before this line
};

// lut: base -> code (0 = undefined), comp: code -> complement code (the reference maps code 0 to the CHARACTER 'N' = 78,
// Alphabet.cpp:51). One warp per line; forward codes and, on both strands, the mirrored complement. Forward undefined
// bases are appended (unordered) to zero_pos as stored positions; base counts skip them (SequenceSet.cpp:99-108).
__global__ void __launch_bounds__(256)
k_fasta_encode(const uint8_t* __restrict__ text, const FastaSeg* __restrict__ segs, uint64_t nseg, const uint64_t* __restrict__ rec_off,
               const uint32_t* __restrict__ rec_L0, int single_strand, const uint8_t* __restrict__ lut_g, const uint8_t* __restrict__ comp_g,
               int A, uint8_t* __restrict__ codes, unsigned long long* __restrict__ base_counts, unsigned long long* __restrict__ zero_pos,
               unsigned long long zero_cap, unsigned long long* __restrict__ zero_cnt) {
    __shared__ uint8_t lut[256], comp[256];
    __shared__ uint32_t cnt_sh[8];
    lut[threadIdx.x] = lut_g[threadIdx.x]; comp[threadIdx.x] = comp_g[threadIdx.x];
    if (threadIdx.x < 8) cnt_sh[threadIdx.x] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    uint32_t mycnt[6] = {0, 0, 0, 0, 0, 0};
    for (uint64_t sidx = warp; sidx < nseg; sidx += nwarps) {
        const FastaSeg sg = segs[sidx];
        const uint64_t base = rec_off[sg.rec];
        const uint64_t L0 = rec_L0[sg.rec];
        const uint8_t* __restrict__ src = text + sg.text_off;
        for (uint32_t t = lane; t < sg.len; t += 32) {
            const uint8_t c = lut[src[t]];
            const uint64_t i = sg.dst + t;
            codes[base + i] = c;
            if (!single_strand) codes[base + 2 * L0 - i] = comp[c];
            if (c) mycnt[c - 1]++;
            else {
                const unsigned long long slot = atomicAdd(zero_cnt, 1ull);
                if (slot < zero_cap) zero_pos[slot] = base + i;
            }
        }
        if (!single_strand && sg.dst == 0 && lane == 0) codes[base + L0] = 0;      // the structural N between the strands
    }
    for (int a = 0; a < A; a++) if (mycnt[a]) atomicAdd(&cnt_sh[a], mycnt[a]);
    __syncthreads();
    if (threadIdx.x < (unsigned)A && cnt_sh[threadIdx.x]) atomicAdd(&base_counts[threadIdx.x], (unsigned long long)cnt_sh[threadIdx.x]);
}

// 21 codes around every undefined base (positions z-10 .. z+10 of ITS record, 0xff outside the record): what the host needs
// to hash the 11-mers that contain it (Sequence.cpp:35-41)
__global__ void k_zero_windows(const uint8_t* __restrict__ codes, const uint64_t* __restrict__ zpos, const uint64_t* __restrict__ zbeg,
                               const uint64_t* __restrict__ zend, uint64_t nz, uint8_t* __restrict__ win /* [nz][21] */) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nz * 21) return;
    const uint64_t k = t / 21; const int d = (int)(t % 21) - 10;
    const long long pos = (long long)zpos[k] + d;
    win[t] = (pos >= (long long)zbeg[k] && pos < (long long)zend[k]) ? codes[pos] : (uint8_t)0xff;
}

}  // namespace bamm
