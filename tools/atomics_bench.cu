// Scratch micro-benchmark: random-access op throughput on a region of given size (L2-resident or not).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
template <int OP>
__global__ void k(uint64_t* tab, uint64_t slots, uint64_t n_per_thread, uint64_t* sink) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint64_t i = 0; i < n_per_thread; i += 8) {
        uint64_t idx[8];
#pragma unroll
        for (int j = 0; j < 8; j++) idx[j] = __umul64hi(mix(tid * n_per_thread + i + j + 12345), slots);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (OP == 0) acc += __ldcg(tab + idx[j]);
            if (OP == 1) atomicAdd((unsigned int*)(tab + idx[j]), 1u);
            if (OP == 2) atomicAdd((unsigned long long*)(tab + idx[j]), 1ull);
            if (OP == 3) acc += atomicCAS((unsigned long long*)(tab + idx[j]), 0ull, 1ull);
            if (OP == 4) acc += atomicAdd((unsigned int*)(tab + idx[j]), 1u);
            if (OP == 5) ((uint32_t*)tab)[idx[j] * 2] = (uint32_t)i;
            if (OP == 6) { uint64_t v = __ldcg(tab + idx[j]); if (v != 77) atomicAdd((unsigned long long*)(tab + idx[j]), 1ull); }
            if (OP == 7) acc += atomicCAS((unsigned int*)(tab + idx[j]), 0u, 1u);
        }
    }
    if (acc == 0x1234567) *sink = acc;
}
int main(int argc, char** argv) {
    const char* names[] = {"ld.cg 8B", "red.add.u32", "red.add.u64", "cas.u64 (ret)", "atom.add.u32 (ret)", "st 4B", "ld+red.u64", "cas.u32 (ret)"};
    uint64_t* sink; cudaMalloc(&sink, 8);
    double sizes_mb[] = {8, 32, 64, 4096};
    for (double mb : sizes_mb) {
        uint64_t slots = (uint64_t)(mb * 1e6 / 8);
        uint64_t* tab; cudaMalloc(&tab, slots * 8); cudaMemset(tab, 0, slots * 8);
        for (int op = 0; op < 8; op++) {
            const int blocks = 148 * 8, threads = 256; const uint64_t npt = 512;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                switch (op) {
                    case 0: k<0><<<blocks, threads>>>(tab, slots, npt, sink); break;
                    case 1: k<1><<<blocks, threads>>>(tab, slots, npt, sink); break;
                    case 2: k<2><<<blocks, threads>>>(tab, slots, npt, sink); break;
                    case 3: k<3><<<blocks, threads>>>(tab, slots, npt, sink); break;
                    case 4: k<4><<<blocks, threads>>>(tab, slots, npt, sink); break;
                    case 5: k<5><<<blocks, threads>>>(tab, slots, npt, sink); break;
                    case 6: k<6><<<blocks, threads>>>(tab, slots, npt, sink); break;
                    case 7: k<7><<<blocks, threads>>>(tab, slots, npt, sink); break;
                }
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double ops = (double)blocks * threads * npt;
            printf("%6.0f MB  %-20s %8.2f Gops/s\n", mb, names[op], ops / ms / 1e6);
        }
        cudaFree(tab);
    }
    cudaError_t e = cudaDeviceSynchronize(); if (e) printf("err %s\n", cudaGetErrorString(e));
    return 0;
}
