// Micro-benchmarks behind DESIGN.md's kernel choices: what a B200 sustains for the access
// patterns of the dedupe stage (random 8-byte loads / reductions / CAS into an L2-sized array).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u32 mix32(u32 h) { h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13; h *= 0xC2B2AE3Du; h ^= h >> 16; return h; }

template <int MODE, int ILP>
__global__ void k(u64 *arr, u32 mask, int iters, u64 *sink)
{
    u32 id = blockIdx.x * blockDim.x + threadIdx.x;
    u64 acc = 0;
    for (int it = 0; it < iters; ++it) {
        u32 idx[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) idx[j] = mix32(id * 2654435761u + (it * ILP + j) * 40503u) & mask;
        if (MODE == 0 || MODE == 2 || MODE == 3) {
            u64 v[ILP];
#pragma unroll
            for (int j = 0; j < ILP; ++j) v[j] = __ldcg(&arr[idx[j]]);
#pragma unroll
            for (int j = 0; j < ILP; ++j) acc += v[j];
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) atomicAdd(&arr[(idx[j] * 16u + (id & 15u)) & mask], 1ull);   // RED, separate line
        }
        if (MODE == 3) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) acc += atomicCAS(&arr[idx[j]], ~0ull, 5ull);
        }
        if (MODE == 4) {   // 32 lanes -> consecutive 8 B (coalesced) RED
#pragma unroll
            for (int j = 0; j < ILP; ++j) atomicAdd(&arr[((idx[j] & ~31u) + (threadIdx.x & 31)) & mask], 1ull);
        }
    }
    if (acc == 0x1234567) *sink = acc;
}

template <int MODE, int ILP>
void run(const char *name, u64 *arr, u32 n, int blocks, int threads, int iters, u64 *sink)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, ILP><<<blocks, threads>>>(arr, n - 1, iters, sink);
    cudaEventRecord(e0);
    k<MODE, ILP><<<blocks, threads>>>(arr, n - 1, iters, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * ILP;
    printf("%-28s arr=%4u MiB blocks=%5d thr=%d ilp=%d: %7.1f us  %7.2f Gops/s\n", name, (unsigned)((u64)n * 8 >> 20), blocks, threads, ILP, ms * 1e3, ops / ms / 1e6);
}

int main()
{
    u64 *arr, *sink;
    const u32 nmax = 1u << 27;   // 1 GiB
    cudaMalloc(&arr, (size_t)nmax * 8); cudaMalloc(&sink, 8);
    cudaMemset(arr, 0, (size_t)nmax * 8);
    for (u32 n : {1u << 21, 1u << 23, 1u << 27}) {   // 16 MiB, 64 MiB, 1 GiB
        for (int blocks : {148 * 4, 148 * 16}) {
            run<0, 1>("random ld.cg 8B", arr, n, blocks, 128, 64, sink);
            run<0, 4>("random ld.cg 8B", arr, n, blocks, 128, 16, sink);
            run<1, 1>("random RED.64", arr, n, blocks, 128, 64, sink);
            run<1, 4>("random RED.64", arr, n, blocks, 128, 16, sink);
            run<2, 1>("ld + RED (2 lines)", arr, n, blocks, 128, 64, sink);
            run<2, 4>("ld + RED (2 lines)", arr, n, blocks, 128, 16, sink);
            run<3, 1>("ld + CAS same word", arr, n, blocks, 128, 64, sink);
            run<4, 4>("coalesced RED.64", arr, n, blocks, 128, 16, sink);
        }
    }
    return 0;
}
