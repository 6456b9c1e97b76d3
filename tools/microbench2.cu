// Micro-benchmark 2: reductions into 128-byte counter lines shared by 16 concurrent "frames"
// (the chunk dedupe layout) versus private lines, after the lines were zeroed by plain stores.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u32 mix32(u32 h) { h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13; h *= 0xC2B2AE3Du; h ^= h >> 16; return h; }

__global__ void k_zero(u64 *arr, u64 n) { for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) arr[i] = 0; }

// MODE 0: word = frame (16 frames share each line); 1: all frames use word 0 of frame-private lines;
// 2: like 0 but 32-bit atomics; 3: like 0 but each line visited by ONE frame only (lines partitioned)
template <int MODE>
__global__ void k(u64 *arr, u32 line_mask, int samples_per_thread, int n_lines_used)
{
    const u32 frame = blockIdx.y;
    const u32 id = blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < samples_per_thread; ++it) {
        u32 h = mix32(id * 2654435761u + it * 40503u);
        u32 line = (h % (u32)n_lines_used);                     // same voxel set for every frame
        line = mix32(line * 0x9E3779B1u) & line_mask;           // scattered over the table
        if (MODE == 0) atomicAdd(&arr[(u64)line * 16 + frame], 1ull);
        if (MODE == 1) atomicAdd(&arr[(u64)((line + frame * 7919u) & line_mask) * 16], 1ull);
        if (MODE == 2) atomicAdd(reinterpret_cast<u32 *>(&arr[(u64)line * 16 + frame]), 1u);
        if (MODE == 3) atomicAdd(&arr[(u64)((line & ~15u) | frame) * 16 + frame], 1ull);
    }
}

template <int MODE>
void run(const char *name, u64 *arr, u32 n_lines, int n_used, int bx, int spt, bool zero_first)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
        if (zero_first) k_zero<<<148 * 8, 256>>>(arr, (u64)n_lines * 16);
        cudaEventRecord(e0);
        k<MODE><<<dim3(bx, 16), 128>>>(arr, n_lines - 1, spt, n_used);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    double ops = (double)bx * 16 * 128 * spt;
    printf("%-44s lines=%7u used=%6d zero=%d: %7.1f us %7.2f Gops/s (%.0f ops)\n", name, n_lines, n_used, (int)zero_first, best * 1e3, ops / best / 1e6, ops);
}

int main()
{
    u64 *arr;
    const u32 n_lines = 1u << 18;      // 32 MiB of 128-byte lines
    cudaMalloc(&arr, (size_t)n_lines * 128);
    cudaMemset(arr, 0, (size_t)n_lines * 128);
    for (int zero = 0; zero < 2; ++zero) {
        // 16 frames x 64 blocks x 128 thr x 9 = 1.18 M REDs, ~90k distinct voxels
        run<0>("16 frames share lines, 64-bit", arr, n_lines, 90000, 64, 9, zero);
        run<2>("16 frames share lines, 32-bit", arr, n_lines, 90000, 64, 9, zero);
        run<1>("frame-private lines, 64-bit", arr, n_lines, 90000, 64, 9, zero);
        run<3>("lines partitioned by frame, 64-bit", arr, n_lines, 90000, 64, 9, zero);
        run<0>("16 frames share lines, 64-bit, 4x work", arr, n_lines, 90000, 256, 9, zero);
    }
    return 0;
}
