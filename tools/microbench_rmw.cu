// What a B200 sustains for the update kernel's table access pattern: read one random 16-byte slot
// (a 32-byte sector moves) and write 8 bytes back into it.  This -- not streaming bandwidth -- is the
// hardware bound of k_apply_chunk when the voxel table is far larger than L2 (profiles/README.md).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench_rmw tools/microbench_rmw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix64(u64 x) { x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31; return x; }

template <int ILP, bool WRITE>
__global__ void k(ulonglong2 *t, u64 mask, int iters, u64 *sink)
{
    const u64 id = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    u64 acc = 0;
    for (int it = 0; it < iters; ++it) {
        u64 idx[ILP]; ulonglong2 v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) idx[j] = mix64(id * 0x9E3779B97F4A7C15ull + (u64)(it * ILP + j)) & mask;
#pragma unroll
        for (int j = 0; j < ILP; ++j) v[j] = __ldcg(&t[idx[j]]);
#pragma unroll
        for (int j = 0; j < ILP; ++j) { acc += v[j].x; if (WRITE) t[idx[j]].y = v[j].y + 1; }
    }
    if (acc == 0x1234567) *sink = acc;
}

template <int ILP, bool WRITE>
void run(ulonglong2 *t, u64 n, int blocks, int threads, int iters, u64 *sink)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP, WRITE><<<blocks, threads>>>(t, n - 1, iters, sink);
    cudaEventRecord(e0);
    k<ILP, WRITE><<<blocks, threads>>>(t, n - 1, iters, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * iters * ILP;
    printf("%s table=%6llu MiB blocks=%5d thr=%d ilp=%d: %8.1f us %7.2f Gops/s  (%.0f GB/s of 32-B sectors%s)\n", WRITE ? "read16+write8" : "read16       ",
           (unsigned long long)(n * 16 >> 20), blocks, threads, ILP, ms * 1e3, ops / ms / 1e6, ops / ms / 1e6 * 32 * (WRITE ? 2 : 1), WRITE ? ", read + write-back" : "");
}

int main()
{
    ulonglong2 *t; u64 *sink;
    const u64 nmax = 1ull << 28;   // 4 GiB of 16-byte slots
    if (cudaMalloc(&t, nmax * 16) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    cudaMemset(t, 0, nmax * 16);
    for (u64 n : {1ull << 22, 1ull << 26, 1ull << 28}) {   // 64 MiB (L2), 1 GiB, 4 GiB
        for (int blocks : {148 * 8, 148 * 32}) {
            run<1, false>(t, n, blocks, 256, 16, sink);
            run<4, false>(t, n, blocks, 256, 4, sink);
            run<1, true>(t, n, blocks, 256, 16, sink);
            run<4, true>(t, n, blocks, 256, 4, sink);
        }
    }
    return 0;
}
