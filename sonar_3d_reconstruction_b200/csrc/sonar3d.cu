// sonar3d.cu -- sm_100a kernels + C-ABI (include/sonar3d.h) of the sonar -> voxel hot path.
//
// Reference behaviour: luckkim123/sonar_3d_reconstruction scripts/3d_mapper.py
//   :387-483 process_sonar_ray   -> k_first_hit + k_expand   (K1, K2+K3 of SURVEY 2.2)
//   :485-567 process_sonar_image -> per-frame dedupe scratch + k_apply (K3, K4)
//   :83-115  update_voxel        -> apply_one()
//   :117-188 queries / export    -> k_query, k_export (K5, K6)
//
// Data layout in HBM (DESIGN.md has the full account):
//   voxel table   Slot[cap]      16 B {packed key, fp64 log-odds}, open addressing, linear probing
//   frame scratch SEntry[G][C]   16 B {packed key, n_occ<<32 | n_free}, one sub-table per in-flight frame
//   touch list    u32[G][C]      scratch slots first touched this frame (so apply/reset are O(unique))
// No tensor cores: the path has no dense contraction; it is integer/fp64 scatter work.
#include "../../include/sonar3d.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr u64 EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;
constexpr int KEY_BITS = 21;
constexpr int KEY_BIAS = 1 << 20;
constexpr u32 ERR_KEYRANGE = 1u, ERR_TABLEFULL = 2u, ERR_SCRATCH = 4u;

struct __align__(16) Slot { u64 key; double val; };
struct __align__(16) SEntry { u64 key; u64 cnt; };

struct DevParams {
    double res, inv_res, lo_occ, lo_free, lo_min, lo_max, a_thr, a_ratio, zmin;
    int adaptive, zfilter, thr;
};

struct DevTables {
    int H, W, n_beams, nv_max, free_step, occ_window;
    const int *beam_col;
    const double *cos_b, *sin_b, *range_m;
    const int *nv_free, *nv_occ;
    const double *cos_va, *sin_va;
    const short *col_to_beam;  // [W] processed-beam index of a column, -1 = not processed
};

struct MapCtr {       // device-resident map counters
    u64 count;        // live voxels
    int kmin[3], kmax[3];
    u32 err;
    u32 pad;
};

struct FrameCtr { u32 list_count; u32 ticket; };   // per in-flight frame slot

struct DevStats { u64 n_occ, n_free, n_voxels, n_samples; };  // == s3d_frame_stats
static_assert(sizeof(DevStats) == sizeof(s3d_frame_stats), "stats layout");

__host__ __device__ __forceinline__ u64 mix64(u64 x)
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

__host__ __device__ __forceinline__ u64 pack_key(int i, int j, int k)
{
    return ((u64)(u32)(i + KEY_BIAS) << (2 * KEY_BITS)) | ((u64)(u32)(j + KEY_BIAS) << KEY_BITS) |
           (u64)(u32)(k + KEY_BIAS);
}

__host__ __device__ __forceinline__ void unpack_key(u64 key, int &i, int &j, int &k)
{
    const u32 m = (1u << KEY_BITS) - 1;
    i = (int)((key >> (2 * KEY_BITS)) & m) - KEY_BIAS;
    j = (int)((key >> KEY_BITS) & m) - KEY_BIAS;
    k = (int)(key & m) - KEY_BIAS;
}

// floor(w / res) exactly as IEEE division would give it (3d_mapper.py:63-65), but with the
// division taken only when the reciprocal product lands within 1e-6 of an integer.
__device__ __forceinline__ bool voxel_index(double w, double res, double inv_res, int &out)
{
    double q = w * inv_res;
    if (!(fabs(q) < (double)(KEY_BIAS - 1))) return false;   // also rejects NaN/inf
    if (fabs(q - rint(q)) < 1e-6) q = __ddiv_rn(w, res);
    out = (int)floor(q);
    return true;
}

// ------------------------------------------------------------------------------------ K1
// First above-threshold range bin per processed beam (3d_mapper.py:406-409).  The image is
// read once, row-major and coalesced (16 B per thread); per-beam minima are kept in shared
// memory per row strip and merged with one atomicMin per (strip, beam) that found a hit.
constexpr int FH_ROWS = 16;
constexpr int FH_THREADS = 256;

template <int VEC>
__global__ void __launch_bounds__(FH_THREADS)
k_first_hit(const uint8_t *__restrict__ imgs, size_t img_stride, DevTables tab, int thr,
            int *__restrict__ first_hit)
{
    extern __shared__ int s_min[];   // [n_beams]
    const int g = blockIdx.y;
    const uint8_t *img = imgs + (size_t)g * img_stride;
    const int H = tab.H, W = tab.W;
    const int r0 = blockIdx.x * FH_ROWS;
    const int r1 = min(r0 + FH_ROWS, H);
    for (int b = threadIdx.x; b < tab.n_beams; b += FH_THREADS) s_min[b] = 0x7f7f7f7f;
    __syncthreads();
    const int vec_per_row = (W + VEC - 1) / VEC;
    const int n_items = (r1 - r0) * vec_per_row;
    for (int it = threadIdx.x; it < n_items; it += FH_THREADS) {
        const int r = r0 + it / vec_per_row;
        const int c0 = (it % vec_per_row) * VEC;
        const uint8_t *p = img + (size_t)r * W + c0;
        if (VEC == 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
            const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                    const int px = (int)((w[q] >> (8 * bb)) & 0xffu);
                    if (px > thr) {
                        const int beam = tab.col_to_beam[c0 + 4 * q + bb];
                        if (beam >= 0) atomicMin(&s_min[beam], r);
                    }
                }
            }
        } else {
            const int px = (int)__ldg(p);
            if (px > thr) {
                const int beam = tab.col_to_beam[c0];
                if (beam >= 0) atomicMin(&s_min[beam], r);
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < tab.n_beams; b += FH_THREADS) {
        const int m = s_min[b];
        if (m < H) atomicMin(&first_hit[g * tab.n_beams + b], m);
    }
}

// ------------------------------------------------------------------------------------ K2+K3
// One block per (processed beam, in-flight frame).  The block lists the beam's range samples
// (free: every free_step-th bin before the first hit; occupied: above-threshold bins in the
// occ_window bins from the first hit), prefix-sums their fan sizes (2*nv+1), and its threads
// then walk the flattened (range sample, vertical step) space: sonar-frame point from the
// host trig tables, Sonar->Map transform, z filter, voxel key, and a count bump in the
// frame's dedupe scratch table.
constexpr int EX_THREADS = 128;

struct ExpandArgs {
    const uint8_t *imgs; size_t img_stride;
    const double *T;             // [G][16]
    DevTables tab; DevParams p;
    const int *first_hit;        // [G][n_beams]
    SEntry *scratch; u32 scratch_mask; size_t scratch_stride;   // per frame slot
    u32 *lists; u32 list_cap;
    FrameCtr *fctr;              // [G]
    DevStats *stats;             // [G]
    MapCtr *mc;
};

__device__ __forceinline__ void scratch_add(SEntry *S, u32 mask, u64 key, u64 inc, u32 *list, u32 list_cap,
                                            FrameCtr *fc, MapCtr *mc)
{
    u32 slot = (u32)mix64(key) & mask;
    for (u32 probe = 0; probe <= mask; ++probe) {
        u64 cur = __ldcg(&S[slot].key);
        if (cur == EMPTY_KEY) {
            cur = atomicCAS(&S[slot].key, EMPTY_KEY, key);
            if (cur == EMPTY_KEY) {
                const u32 idx = atomicAdd(&fc->list_count, 1u);
                if (idx < list_cap) list[idx] = slot; else atomicOr(&mc->err, ERR_SCRATCH);
                cur = key;
            }
        }
        if (cur == key) { atomicAdd(&S[slot].cnt, inc); return; }
        slot = (slot + 1) & mask;
    }
    atomicOr(&mc->err, ERR_SCRATCH);
}

__global__ void __launch_bounds__(EX_THREADS)
k_expand(ExpandArgs a)
{
    extern __shared__ int s_dyn[];
    const DevTables &tab = a.tab;
    const int H = tab.H, W = tab.W;
    const int max_c = (H + tab.free_step - 1) / tab.free_step + tab.occ_window;
    int *s_off = s_dyn;                 // [max_c + 1] exclusive prefix of fan sizes
    int *s_rn = s_dyn + max_c + 1;      // [max_c] r | (occupied << 30); fan half-width in s_nv
    int *s_nv = s_rn + max_c;           // [max_c]
    __shared__ double s_T[12];
    __shared__ int s_total;

    const int beam = blockIdx.x, g = blockIdx.y;
    const uint8_t *img = a.imgs + (size_t)g * a.img_stride;
    const int col = tab.beam_col[beam];
    int fh = a.first_hit[g * tab.n_beams + beam];
    fh = fh < H ? fh : H;                                          // no hit -> whole ray is free (:412-413)
    const int nfc = (fh + tab.free_step - 1) / tab.free_step;      // range(0, fh, free_step) (:420)
    const int noc = fh < H ? min(tab.occ_window, H - fh) : 0;      // range(fh, min(fh+50, H)) (:451)
    const int nc = nfc + noc;
    if (threadIdx.x < 12) s_T[threadIdx.x] = a.T[g * 16 + threadIdx.x];
    for (int c = threadIdx.x; c < nc; c += EX_THREADS) {
        int r, nv, occ;
        if (c < nfc) { r = c * tab.free_step; nv = tab.nv_free[r]; occ = 0; }
        else {
            r = fh + (c - nfc); occ = 1;
            nv = ((int)img[(size_t)r * W + col] > a.p.thr) ? tab.nv_occ[r] : 0;   // :452
        }
        s_rn[c] = r | (occ << 30);
        s_nv[c] = nv;
        s_off[c] = nv > 0 ? 2 * nv + 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x < 32) {          // warp 0: exclusive scan of <= a few hundred fan sizes
        int carry = 0;
        for (int base = 0; base < nc; base += 32) {
            const int c = base + threadIdx.x;
            const int v = c < nc ? s_off[c] : 0;
            int incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)threadIdx.x >= d) incl += t;
            }
            if (c < nc) s_off[c] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (threadIdx.x == 0) { s_off[nc] = carry; s_total = carry; }
    }
    __syncthreads();
    const int total = s_total;
    const double cb = tab.cos_b[beam], sb = tab.sin_b[beam];
    SEntry *S = a.scratch + (size_t)g * a.scratch_stride;
    u32 *list = a.lists + (size_t)g * a.list_cap;
    FrameCtr *fc = a.fctr + g;
    int emitted = 0;
    for (int w = threadIdx.x; w < total; w += EX_THREADS) {
        int lo = 0, hi = nc;                      // largest c with s_off[c] <= w
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_off[mid] <= w) lo = mid; else hi = mid;
        }
        const int c = lo;
        const int rn = s_rn[c], nv = s_nv[c];
        const int r = rn & 0x3fffffff, occ = rn >> 30;
        const int vi = w - s_off[c];              // v_step + nv
        const int ti = nv * nv - 1 + vi;
        const double cv = __ldg(&tab.cos_va[ti]), sv = __ldg(&tab.sin_va[ti]);
        const double range = __ldg(&tab.range_m[r]);
        // sonar frame, X fwd / Y right / Z down, products rounded left to right (:434-436)
        const double rc = __dmul_rn(range, cv);
        const double xs = __dmul_rn(rc, cb);
        const double ys = -__dmul_rn(rc, sb);
        const double zs = __dmul_rn(range, sv);
        // T @ [x,y,z,1] as numpy evaluates it: (t0*x + t2*z) + (t1*y + t3) (:440)
        double wv[3];
#pragma unroll
        for (int q = 0; q < 3; ++q)
            wv[q] = __dadd_rn(__dadd_rn(__dmul_rn(s_T[4 * q], xs), __dmul_rn(s_T[4 * q + 2], zs)),
                              __dadd_rn(__dmul_rn(s_T[4 * q + 1], ys), s_T[4 * q + 3]));
        if (a.p.zfilter && wv[2] < a.p.zmin) continue;          // :443, :478
        ++emitted;
        int ki, kj, kk;
        if (!(voxel_index(wv[0], a.p.res, a.p.inv_res, ki) && voxel_index(wv[1], a.p.res, a.p.inv_res, kj) &&
              voxel_index(wv[2], a.p.res, a.p.inv_res, kk))) {
            atomicOr(&a.mc->err, ERR_KEYRANGE);
            continue;
        }
        scratch_add(S, a.scratch_mask, pack_key(ki, kj, kk), occ ? (1ull << 32) : 1ull, list, a.list_cap, fc, a.mc);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) emitted += __shfl_xor_sync(0xffffffffu, emitted, d);
    if ((threadIdx.x & 31) == 0 && emitted) atomicAdd(&a.stats[g].n_samples, (u64)emitted);
}

// ------------------------------------------------------------------------------------ K4
// update_voxel (3d_mapper.py:83-115) on one table slot.
__device__ __forceinline__ double apply_one(double L, double upd, bool adaptive, const DevParams &p)
{
    if (adaptive && p.adaptive && upd > 0.0) {                  // :95
        const double prob = 1.0 / (1.0 + exp(-L));              // :97
        if (prob <= p.a_thr) upd *= (prob / p.a_thr) * p.a_ratio;   // :100-102
    }
    L += upd;                                                   // :107
    L = fmin(fmax(L, p.lo_min), p.lo_max);                      // :110
    return L;
}

// find-or-insert; returns slot index or ~0 on failure.  `fresh` says the key was inserted.
__device__ __forceinline__ u64 table_find_or_insert(Slot *table, u64 mask, u64 key, bool &fresh, double &val)
{
    u64 slot = (mix64(key) >> 8) & mask;
    fresh = false;
    for (u32 probe = 0; probe < (1u << 20); ++probe) {
        const ulonglong2 raw = __ldcg(reinterpret_cast<const ulonglong2 *>(&table[slot]));
        u64 cur = raw.x;
        if (cur == key) { val = __longlong_as_double((long long)raw.y); return slot; }
        if (cur == EMPTY_KEY) {
            cur = atomicCAS(&table[slot].key, EMPTY_KEY, key);
            if (cur == EMPTY_KEY) { fresh = true; val = 0.0; return slot; }     // :105-106
            if (cur == key) { val = __ldcg(&table[slot].val); return slot; }
        }
        slot = (slot + 1) & mask;
    }
    return ~0ull;
}

struct LocalAcc { int n_occ, n_free, n_new; int kmin[3], kmax[3]; };

__device__ __forceinline__ void acc_init(LocalAcc &a)
{
    a.n_occ = a.n_free = a.n_new = 0;
    for (int q = 0; q < 3; ++q) { a.kmin[q] = INT_MAX; a.kmax[q] = INT_MIN; }
}

// warp-reduce the accumulators; lane 0 publishes (bounds only when they extend the box)
__device__ __forceinline__ void acc_publish(LocalAcc &a, DevStats *st, MapCtr *mc)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a.n_occ += __shfl_xor_sync(0xffffffffu, a.n_occ, d);
        a.n_free += __shfl_xor_sync(0xffffffffu, a.n_free, d);
        a.n_new += __shfl_xor_sync(0xffffffffu, a.n_new, d);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            a.kmin[q] = min(a.kmin[q], __shfl_xor_sync(0xffffffffu, a.kmin[q], d));
            a.kmax[q] = max(a.kmax[q], __shfl_xor_sync(0xffffffffu, a.kmax[q], d));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        if (st) {
            if (a.n_occ) atomicAdd(&st->n_occ, (u64)a.n_occ);
            if (a.n_free) atomicAdd(&st->n_free, (u64)a.n_free);
        }
        if (a.n_new) atomicAdd(&mc->count, (u64)a.n_new);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (a.kmin[q] < __ldcg(&mc->kmin[q])) atomicMin(&mc->kmin[q], a.kmin[q]);
            if (a.kmax[q] > __ldcg(&mc->kmax[q])) atomicMax(&mc->kmax[q], a.kmax[q]);
        }
    }
}

constexpr int AP_THREADS = 256;

// One thread per voxel touched by the frame (3d_mapper.py:557-567): consume and reset the
// scratch entry, form the per-voxel mean update, then read-modify-write the table slot.
__global__ void __launch_bounds__(AP_THREADS)
k_apply(SEntry *__restrict__ S, const u32 *__restrict__ list, FrameCtr *fc, DevStats *st,
        Slot *table, u64 tmask, DevParams p, MapCtr *mc)
{
    const u32 n = fc->list_count;
    LocalAcc acc; acc_init(acc);
    for (u32 i = blockIdx.x * AP_THREADS + threadIdx.x; i < n; i += gridDim.x * AP_THREADS) {
        const u32 s = list[i];
        const ulonglong2 e = *reinterpret_cast<const ulonglong2 *>(&S[s]);
        *reinterpret_cast<ulonglong2 *>(&S[s]) = make_ulonglong2(EMPTY_KEY, 0ull);   // ready for the next frame
        const u64 key = e.x;
        const u32 n_occ = (u32)(e.y >> 32), n_free = (u32)(e.y & 0xffffffffu);
        // mean of the per-sample deltas, summed one by one as the reference does (:546, :559)
        double sum = 0.0;
        for (u32 q = 0; q < n_free; ++q) sum += p.lo_free;
        for (u32 q = 0; q < n_occ; ++q) sum += p.lo_occ;
        const double avg = sum / (double)(n_occ + n_free);
        const bool occ_typed = n_occ > 0;                                            // :544-545
        bool fresh; double L;
        const u64 slot = table_find_or_insert(table, tmask, key, fresh, L);
        if (slot == ~0ull) { atomicOr(&mc->err, ERR_TABLEFULL); continue; }
        L = apply_one(L, avg, occ_typed, p);
        table[slot].val = L;
        if (occ_typed) ++acc.n_occ; else ++acc.n_free;
        if (fresh) ++acc.n_new;
        int ki, kj, kk; unpack_key(key, ki, kj, kk);
        acc.kmin[0] = min(acc.kmin[0], ki); acc.kmax[0] = max(acc.kmax[0], ki);
        acc.kmin[1] = min(acc.kmin[1], kj); acc.kmax[1] = max(acc.kmax[1], kj);
        acc.kmin[2] = min(acc.kmin[2], kk); acc.kmax[2] = max(acc.kmax[2], kk);
    }
    acc_publish(acc, st, mc);
    // last block out: snapshot len(voxels) (:592) and re-arm the frame slot
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const u32 t = atomicAdd(&fc->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        st->n_voxels = atomicAdd(&mc->count, 0ull);
        fc->list_count = 0;
        fc->ticket = 0;
    }
}

// ------------------------------------------------------------------------------ store kernels
__global__ void k_fill_slots(Slot *t, u64 n)
{
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        *reinterpret_cast<ulonglong2 *>(&t[i]) = make_ulonglong2(EMPTY_KEY, 0ull);
}

__global__ void k_fill_scratch(SEntry *t, u64 n)
{
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        *reinterpret_cast<ulonglong2 *>(&t[i]) = make_ulonglong2(EMPTY_KEY, 0ull);
}

__global__ void k_rehash(const Slot *__restrict__ old_t, u64 old_n, Slot *new_t, u64 new_mask, MapCtr *mc)
{
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < old_n; i += (u64)gridDim.x * blockDim.x) {
        const ulonglong2 raw = *reinterpret_cast<const ulonglong2 *>(&old_t[i]);
        if (raw.x == EMPTY_KEY) continue;
        bool fresh; double v;
        const u64 slot = table_find_or_insert(new_t, new_mask, raw.x, fresh, v);
        if (slot == ~0ull) { atomicOr(&mc->err, ERR_TABLEFULL); continue; }
        new_t[slot].val = __longlong_as_double((long long)raw.y);
    }
}

// update_voxel for explicit (key, delta, adaptive) triples with unique keys per launch
__global__ void k_apply_direct(const u64 *__restrict__ keys, const double *__restrict__ delta,
                               const uint8_t *__restrict__ adaptive, u64 n, Slot *table, u64 tmask,
                               DevParams p, MapCtr *mc)
{
    LocalAcc acc; acc_init(acc);
    const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (i < n) {
        const u64 key = keys[i];
        bool fresh; double L;
        const u64 slot = table_find_or_insert(table, tmask, key, fresh, L);
        if (slot == ~0ull) atomicOr(&mc->err, ERR_TABLEFULL);
        else {
            table[slot].val = apply_one(L, delta[i], adaptive[i] != 0, p);
            if (fresh) ++acc.n_new;
            int ki, kj, kk; unpack_key(key, ki, kj, kk);
            acc.kmin[0] = acc.kmax[0] = ki; acc.kmin[1] = acc.kmax[1] = kj; acc.kmin[2] = acc.kmax[2] = kk;
        }
    }
    acc_publish(acc, nullptr, mc);
}

__global__ void k_load(const u64 *__restrict__ keys, const double *__restrict__ vals, u64 n, Slot *table,
                       u64 tmask, MapCtr *mc)
{
    LocalAcc acc; acc_init(acc);
    const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (i < n) {
        const u64 key = keys[i];
        bool fresh; double L;
        const u64 slot = table_find_or_insert(table, tmask, key, fresh, L);
        if (slot == ~0ull) atomicOr(&mc->err, ERR_TABLEFULL);
        else {
            table[slot].val = vals[i];
            if (fresh) ++acc.n_new;
            int ki, kj, kk; unpack_key(key, ki, kj, kk);
            acc.kmin[0] = acc.kmax[0] = ki; acc.kmin[1] = acc.kmax[1] = kj; acc.kmin[2] = acc.kmax[2] = kk;
        }
    }
    acc_publish(acc, nullptr, mc);
}

__global__ void k_query(const u64 *__restrict__ keys, u64 n, const Slot *__restrict__ table, u64 tmask,
                        double *__restrict__ out, uint8_t *__restrict__ found)
{
    const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 key = keys[i];
    u64 slot = (mix64(key) >> 8) & tmask;
    double v = 0.0; uint8_t f = 0;
    for (u64 probe = 0; probe <= tmask; ++probe) {
        const ulonglong2 raw = *reinterpret_cast<const ulonglong2 *>(&table[slot]);
        if (raw.x == key) { v = __longlong_as_double((long long)raw.y); f = 1; break; }
        if (raw.x == EMPTY_KEY) break;
        slot = (slot + 1) & tmask;
    }
    out[i] = v; found[i] = f;
}

// ------------------------------------------------------------------------------------ K5
struct ExportOut { double *xyz, *prob, *L; int8_t *cls; int *ijk; u64 *counts /*[4]: free, unknown, occupied, staged*/; };

__global__ void __launch_bounds__(256)
k_export(const Slot *__restrict__ table, u64 n_slots, double res, double thr_occ, double thr_free, u32 class_mask,
         ExportOut o)
{
    __shared__ u32 s_cnt[3];
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 n_iter = (n_slots + stride - 1) / stride;      // uniform trip count keeps ballots full-warp
    for (u64 it = 0; it < n_iter; ++it) {
        const u64 i = it * stride + blockIdx.x * (u64)blockDim.x + threadIdx.x;
        bool live = false; int cls = 0; u64 key = 0; double L = 0.0;
        if (i < n_slots) {
            const ulonglong2 raw = __ldcs(reinterpret_cast<const ulonglong2 *>(&table[i]));
            if (raw.x != EMPTY_KEY) {
                live = true; key = raw.x; L = __longlong_as_double((long long)raw.y);
                cls = L < thr_free ? S3D_CLASS_FREE : (L > thr_occ ? S3D_CLASS_OCCUPIED : S3D_CLASS_UNKNOWN);
            }
        }
        const bool sel = live && ((class_mask >> cls) & 1u);
        const u32 lane = threadIdx.x & 31;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const u32 b = __ballot_sync(0xffffffffu, live && cls == c);
            if (lane == 0 && b) atomicAdd(&s_cnt[c], __popc(b));
        }
        const u32 bal = __ballot_sync(0xffffffffu, sel);
        if (bal) {
            u64 base = 0;
            if (lane == 0) base = atomicAdd(&o.counts[3], (u64)__popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (sel) {
                const u64 dst = base + __popc(bal & ((1u << lane) - 1));
                int ki, kj, kk; unpack_key(key, ki, kj, kk);
                o.ijk[3 * dst] = ki; o.ijk[3 * dst + 1] = kj; o.ijk[3 * dst + 2] = kk;
                o.xyz[3 * dst] = __dmul_rn((double)ki + 0.5, res);       // key_to_world (:78-80)
                o.xyz[3 * dst + 1] = __dmul_rn((double)kj + 0.5, res);
                o.xyz[3 * dst + 2] = __dmul_rn((double)kk + 0.5, res);
                o.prob[dst] = 1.0 / (1.0 + exp(-L));                     // :150
                o.L[dst] = L;
                o.cls[dst] = (int8_t)cls;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 3 && s_cnt[threadIdx.x]) atomicAdd(&o.counts[threadIdx.x], (u64)s_cnt[threadIdx.x]);
}

__global__ void k_pack_xyzi32(const double *__restrict__ xyz, const double *__restrict__ prob, u64 n,
                              float4 *__restrict__ out)
{
    const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4((float)xyz[3 * i], (float)xyz[3 * i + 1], (float)xyz[3 * i + 2], (float)prob[i]);
}

__global__ void k_reset_ctr(MapCtr *mc)
{
    mc->count = 0; mc->err = 0;
    for (int q = 0; q < 3; ++q) { mc->kmin[q] = INT_MAX; mc->kmax[q] = INT_MIN; }
}

// ------------------------------------------------------------------------------------ host
thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? S3D_ENOMEM : S3D_ECUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                               \
    } while (0)

u64 next_pow2(u64 v) { u64 p = 1; while (p < v) p <<= 1; return p; }

template <typename T> struct DevBuf {
    T *p = nullptr; size_t n = 0;
    int ensure(size_t want) {
        if (want <= n) return 0;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        size_t cap = std::max<size_t>(want, 16);
        cudaError_t e = cudaMalloc(&p, cap * sizeof(T));
        if (e != cudaSuccess) return fail(S3D_ENOMEM, "cudaMalloc(%zu bytes): %s", cap * sizeof(T), cudaGetErrorString(e));
        n = cap;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

} // namespace

struct s3d_map {
    int device = 0;
    cudaStream_t stream = nullptr;
    int n_sm = 148;
    // voxel table
    Slot *table = nullptr; u64 cap = 0;
    MapCtr *mc = nullptr;            // device
    MapCtr *mc_host = nullptr;       // pinned mirror
    u64 count_known = 0;             // exact count at last sync
    u64 max_touched = 0;             // largest per-frame unique-voxel count seen
    // params / tables
    bool have_params = false, have_tables = false;
    DevParams p{};
    DevTables tab{};
    DevBuf<int> d_beam_col, d_nv_free, d_nv_occ;
    DevBuf<double> d_cos_b, d_sin_b, d_range, d_cos_va, d_sin_va;
    DevBuf<short> d_col_to_beam;
    u64 samples_max = 0;             // worst-case samples per frame for these tables
    // in-flight frame slots
    int G = 16;
    DevBuf<SEntry> scratch; u64 scratch_cap = 0;   // per frame slot
    DevBuf<u32> lists;
    DevBuf<int> first_hit;
    FrameCtr *fctr = nullptr;
    DevBuf<DevStats> stats; DevStats *stats_host = nullptr; size_t stats_host_n = 0;
    // staging
    DevBuf<uint8_t> img_dev; DevBuf<double> T_dev;
    uint8_t *img_pinned = nullptr; size_t img_pinned_n = 0;
    double *T_pinned = nullptr; size_t T_pinned_n = 0;
    DevBuf<u64> io_keys; DevBuf<double> io_vals; DevBuf<uint8_t> io_flags;
    // export staging
    DevBuf<double> ex_xyz, ex_prob, ex_L; DevBuf<int8_t> ex_cls; DevBuf<int> ex_ijk; DevBuf<float4> ex_f32;
    u64 *ex_counts = nullptr; u64 *ex_counts_host = nullptr; u64 ex_n = 0; bool ex_valid = false;
};

namespace {

int set_device(s3d_map *m) { CU(cudaSetDevice(m->device)); return 0; }

int launch_fill_table(s3d_map *m, Slot *t, u64 n)
{
    const int blocks = (int)std::min<u64>((n + 255) / 256, (u64)m->n_sm * 16);
    k_fill_slots<<<blocks, 256, 0, m->stream>>>(t, n);
    CU(cudaGetLastError());
    return 0;
}

// read back the map counters (sync) and turn device error flags into return codes
int sync_counters(s3d_map *m)
{
    CU(cudaMemcpyAsync(m->mc_host, m->mc, sizeof(MapCtr), cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    m->count_known = m->mc_host->count;
    const u32 err = m->mc_host->err;
    if (err) {
        u32 zero = 0;   // clear so that the map stays usable after the caller handles the error
        cudaMemcpyAsync(&m->mc->err, &zero, sizeof zero, cudaMemcpyHostToDevice, m->stream);
        cudaStreamSynchronize(m->stream);
        if (err & ERR_TABLEFULL) return fail(S3D_ETABLEFULL, "voxel table full (capacity %llu slots)", (unsigned long long)m->cap);
        if (err & ERR_SCRATCH) return fail(S3D_ESCRATCH, "per-frame dedupe scratch overflow");
        if (err & ERR_KEYRANGE) return fail(S3D_EKEYRANGE, "voxel key outside +-2^20 or non-finite coordinate");
    }
    return 0;
}

int grow_table(s3d_map *m, u64 new_cap)
{
    new_cap = next_pow2(new_cap);
    if (new_cap <= m->cap) return 0;
    Slot *nt = nullptr;
    cudaError_t e = cudaMalloc(&nt, new_cap * sizeof(Slot));
    if (e != cudaSuccess) return fail(S3D_ETABLEFULL, "cannot grow voxel table to %llu slots: %s",
                                      (unsigned long long)new_cap, cudaGetErrorString(e));
    int rc = launch_fill_table(m, nt, new_cap);
    if (rc) return rc;
    if (m->table) {
        const int blocks = (int)std::min<u64>((m->cap + 255) / 256, (u64)m->n_sm * 16);
        k_rehash<<<blocks, 256, 0, m->stream>>>(m->table, m->cap, nt, new_cap - 1, m->mc);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(m->stream));
        CU(cudaFree(m->table));
    }
    m->table = nt; m->cap = new_cap;
    m->ex_valid = false;
    return 0;
}

// keep load factor <= 1/2 for `extra` more voxels on top of the last known count
int ensure_room(s3d_map *m, u64 extra)
{
    const u64 need = 2 * (m->count_known + extra);
    if (need > m->cap) return grow_table(m, std::max(need, m->cap * 2));
    return 0;
}

template <typename T> int upload(DevBuf<T> &b, const T *src, size_t n, cudaStream_t s)
{
    int rc = b.ensure(n); if (rc) return rc;
    CU(cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
    return 0;
}

int ensure_frame_slots(s3d_map *m)
{
    // scratch sized for the worst-case number of samples a frame can emit (every sample a new key)
    const u64 want = next_pow2(std::max<u64>(1024, m->samples_max + m->samples_max / 2));
    if (want != m->scratch_cap || !m->scratch.p) {
        CU(cudaStreamSynchronize(m->stream));
        m->scratch_cap = want;
        int rc = m->scratch.ensure((size_t)want * m->G); if (rc) return rc;
        rc = m->lists.ensure((size_t)want * m->G); if (rc) return rc;
        const u64 n = m->scratch.n;
        const int blocks = (int)std::min<u64>((n + 255) / 256, (u64)m->n_sm * 16);
        k_fill_scratch<<<blocks, 256, 0, m->stream>>>(m->scratch.p, n);
        CU(cudaGetLastError());
        CU(cudaMemsetAsync(m->fctr, 0, sizeof(FrameCtr) * m->G, m->stream));
    }
    int rc = m->first_hit.ensure((size_t)m->G * std::max(1, m->tab.n_beams)); if (rc) return rc;
    return 0;
}

// The device pipeline for n frames whose images / transforms already sit in device memory.
int run_frames(s3d_map *m, const uint8_t *imgs_dev, int64_t n, const double *T_dev, DevStats *stats_dev)
{
    const DevTables &tab = m->tab;
    const size_t img_stride = (size_t)tab.H * tab.W;
    CU(cudaMemsetAsync(stats_dev, 0, sizeof(DevStats) * (size_t)n, m->stream));
    if (tab.n_beams == 0 || tab.H == 0) {
        // nothing to expand; num_voxels still has to be reported
        for (int64_t f = 0; f < n; ++f)
            CU(cudaMemcpyAsync(&stats_dev[f].n_voxels, &m->mc->count, sizeof(u64), cudaMemcpyDeviceToDevice, m->stream));
        return 0;
    }
    const bool vec16 = (tab.W % 16 == 0) && ((uintptr_t)imgs_dev % 16 == 0);
    const int max_c = (tab.H + tab.free_step - 1) / tab.free_step + tab.occ_window;
    const size_t ex_smem = sizeof(int) * (size_t)(3 * max_c + 1);
    const int apply_blocks = m->n_sm * 2;
    for (int64_t base = 0; base < n; base += m->G) {
        const int g = (int)std::min<int64_t>(m->G, n - base);
        CU(cudaMemsetAsync(m->first_hit.p, 0x7f, sizeof(int) * (size_t)g * tab.n_beams, m->stream));
        dim3 g1((tab.H + FH_ROWS - 1) / FH_ROWS, g);
        const size_t fh_smem = sizeof(int) * (size_t)tab.n_beams;
        if (vec16)
            k_first_hit<16><<<g1, FH_THREADS, fh_smem, m->stream>>>(imgs_dev + base * img_stride, img_stride, tab, m->p.thr, m->first_hit.p);
        else
            k_first_hit<1><<<g1, FH_THREADS, fh_smem, m->stream>>>(imgs_dev + base * img_stride, img_stride, tab, m->p.thr, m->first_hit.p);
        ExpandArgs a;
        a.imgs = imgs_dev + base * img_stride; a.img_stride = img_stride;
        a.T = T_dev + base * 16;
        a.tab = tab; a.p = m->p;
        a.first_hit = m->first_hit.p;
        a.scratch = m->scratch.p; a.scratch_mask = (u32)(m->scratch_cap - 1); a.scratch_stride = m->scratch_cap;
        a.lists = m->lists.p; a.list_cap = (u32)m->scratch_cap;
        a.fctr = m->fctr; a.stats = stats_dev + base; a.mc = m->mc;
        k_expand<<<dim3(tab.n_beams, g), EX_THREADS, ex_smem, m->stream>>>(a);
        for (int f = 0; f < g; ++f)
            k_apply<<<apply_blocks, AP_THREADS, 0, m->stream>>>(
                m->scratch.p + (size_t)f * m->scratch_cap, m->lists.p + (size_t)f * m->scratch_cap, m->fctr + f,
                stats_dev + base + f, m->table, m->cap - 1, m->p, m->mc);
    }
    CU(cudaGetLastError());
    m->ex_valid = false;
    return 0;
}

int check_ready(s3d_map *m)
{
    if (!m) return fail(S3D_EINVAL, "null map");
    if (!m->have_params) return fail(S3D_EINVAL, "s3d_set_params has not been called");
    if (!m->have_tables) return fail(S3D_EINVAL, "s3d_set_tables has not been called");
    return 0;
}

// room for n more frames: each frame can add at most its unique-voxel count; bounded by the
// largest count seen so far (x1.5), or by the worst-case sample count before any frame ran.
int reserve_frames(s3d_map *m, int64_t n)
{
    u64 per_frame = m->max_touched ? m->max_touched + m->max_touched / 2 + 1024 : m->samples_max;
    per_frame = std::min<u64>(per_frame, m->samples_max);
    return ensure_room(m, per_frame * (u64)n);
}

int finish_stats(s3d_map *m, const DevStats *stats_dev, int64_t n, s3d_frame_stats *out)
{
    if (m->stats_host_n < (size_t)n) {
        if (m->stats_host) cudaFreeHost(m->stats_host);
        m->stats_host = nullptr; m->stats_host_n = 0;
        CU(cudaMallocHost(&m->stats_host, sizeof(DevStats) * (size_t)n));
        m->stats_host_n = (size_t)n;
    }
    CU(cudaMemcpyAsync(m->stats_host, stats_dev, sizeof(DevStats) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    int rc = sync_counters(m);
    for (int64_t f = 0; f < n; ++f) {
        const DevStats &s = m->stats_host[f];
        m->max_touched = std::max<u64>(m->max_touched, s.n_occ + s.n_free);
        if (out) {
            out[f].num_occupied = (int64_t)s.n_occ; out[f].num_free = (int64_t)s.n_free;
            out[f].num_voxels = (int64_t)s.n_voxels; out[f].num_samples = (int64_t)s.n_samples;
        }
    }
    return rc;
}

int pack_keys_host(const int32_t *ijk, int64_t n, std::vector<u64> &out)
{
    out.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const int a = ijk[3 * i], b = ijk[3 * i + 1], c = ijk[3 * i + 2];
        if (a < -KEY_BIAS || a >= KEY_BIAS || b < -KEY_BIAS || b >= KEY_BIAS || c < -KEY_BIAS || c >= KEY_BIAS)
            return fail(S3D_EKEYRANGE, "key (%d,%d,%d) outside +-2^20", a, b, c);
        out[(size_t)i] = pack_key(a, b, c);
    }
    return 0;
}

} // namespace

// ======================================================================================= C-ABI
extern "C" {

const char *s3d_last_error(void) { return g_err.c_str(); }
int s3d_abi_version(void) { return S3D_ABI_VERSION; }

int s3d_create(int device, uint64_t initial_capacity, s3d_map **out)
{
    if (!out) return fail(S3D_EINVAL, "out is null");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(S3D_ECUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n_dev) return fail(S3D_EINVAL, "device %d out of range (0..%d)", device, n_dev - 1);
    s3d_map *m = new s3d_map();
    m->device = device;
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    m->n_sm = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&m->mc, sizeof(MapCtr)));
    CU(cudaMallocHost(&m->mc_host, sizeof(MapCtr)));
    CU(cudaMalloc(&m->fctr, sizeof(FrameCtr) * m->G));
    CU(cudaMemsetAsync(m->fctr, 0, sizeof(FrameCtr) * m->G, m->stream));
    CU(cudaMalloc(&m->ex_counts, sizeof(u64) * 4));
    CU(cudaMallocHost(&m->ex_counts_host, sizeof(u64) * 4));
    k_reset_ctr<<<1, 1, 0, m->stream>>>(m->mc);
    int rc = grow_table(m, initial_capacity ? initial_capacity : (1ull << 22));
    if (rc) { s3d_destroy(m); return rc; }
    CU(cudaStreamSynchronize(m->stream));
    *out = m;
    return 0;
}

int s3d_destroy(s3d_map *m)
{
    if (!m) return 0;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->table) cudaFree(m->table);
    if (m->mc) cudaFree(m->mc);
    if (m->mc_host) cudaFreeHost(m->mc_host);
    if (m->fctr) cudaFree(m->fctr);
    if (m->ex_counts) cudaFree(m->ex_counts);
    if (m->ex_counts_host) cudaFreeHost(m->ex_counts_host);
    if (m->stats_host) cudaFreeHost(m->stats_host);
    if (m->img_pinned) cudaFreeHost(m->img_pinned);
    if (m->T_pinned) cudaFreeHost(m->T_pinned);
    m->d_beam_col.release(); m->d_nv_free.release(); m->d_nv_occ.release();
    m->d_cos_b.release(); m->d_sin_b.release(); m->d_range.release(); m->d_cos_va.release(); m->d_sin_va.release();
    m->d_col_to_beam.release(); m->scratch.release(); m->lists.release(); m->first_hit.release(); m->stats.release();
    m->img_dev.release(); m->T_dev.release(); m->io_keys.release(); m->io_vals.release(); m->io_flags.release();
    m->ex_xyz.release(); m->ex_prob.release(); m->ex_L.release(); m->ex_cls.release(); m->ex_ijk.release(); m->ex_f32.release();
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return 0;
}

int s3d_set_params(s3d_map *m, const s3d_params *q)
{
    if (!m || !q) return fail(S3D_EINVAL, "null argument");
    if (!(q->resolution > 0.0) || !std::isfinite(q->resolution)) return fail(S3D_EINVAL, "resolution must be > 0");
    DevParams &p = m->p;
    p.res = q->resolution; p.inv_res = 1.0 / q->resolution;
    p.lo_occ = q->log_odds_occupied; p.lo_free = q->log_odds_free;
    p.lo_min = q->log_odds_min; p.lo_max = q->log_odds_max;
    p.a_thr = q->adaptive_threshold; p.a_ratio = q->adaptive_max_ratio;
    p.zmin = q->z_filter_min;
    p.adaptive = q->adaptive_update; p.zfilter = q->z_filter_enabled;
    p.thr = std::max(-1, std::min(255, q->intensity_threshold));
    m->have_params = true;
    m->ex_valid = false;
    return 0;
}

int s3d_set_tables(s3d_map *m, const s3d_tables *t)
{
    if (!m || !t) return fail(S3D_EINVAL, "null argument");
    if (t->H < 0 || t->W < 0 || t->n_beams < 0 || t->n_beams > t->W || t->nv_max < 0)
        return fail(S3D_EINVAL, "bad table shape");
    if (t->free_step < 1 || t->occ_window < 0) return fail(S3D_EINVAL, "bad free_step/occ_window");
    if (t->H >= (1 << 30)) return fail(S3D_EINVAL, "H too large");
    int rc = set_device(m); if (rc) return rc;
    CU(cudaStreamSynchronize(m->stream));
    const size_t nb = (size_t)t->n_beams, H = (size_t)t->H;
    const size_t nfan = (size_t)t->nv_max * ((size_t)t->nv_max + 2);
    std::vector<short> c2b((size_t)t->W, (short)-1);
    for (size_t b = 0; b < nb; ++b) {
        const int col = t->beam_col[b];
        if (col < 0 || col >= t->W) return fail(S3D_EINVAL, "beam_col[%zu]=%d out of range", b, col);
        c2b[(size_t)col] = (short)b;
    }
    if (nb > 32767) return fail(S3D_EINVAL, "too many processed beams");
    u64 free_sum = 0, occ_best = 0, occ_run = 0;
    for (size_t r = 0; r < H; ++r) {
        if (t->nv_free[r] < 0 || t->nv_free[r] > t->nv_max || t->nv_occ[r] < 0 || t->nv_occ[r] > t->nv_max)
            return fail(S3D_EINVAL, "nv table entry out of range at r=%zu", r);
        if (r % (size_t)t->free_step == 0 && t->nv_free[r] > 0) free_sum += 2 * (u64)t->nv_free[r] + 1;
        occ_run += t->nv_occ[r] > 0 ? 2 * (u64)t->nv_occ[r] + 1 : 0;
        if (r >= (size_t)t->occ_window) {
            const int o = t->nv_occ[r - (size_t)t->occ_window];
            occ_run -= o > 0 ? 2 * (u64)o + 1 : 0;
        }
        occ_best = std::max(occ_best, occ_run);
    }
    m->samples_max = std::max<u64>(1, (free_sum + occ_best) * (u64)nb);
    if ((rc = upload(m->d_beam_col, t->beam_col, nb, m->stream))) return rc;
    if ((rc = upload(m->d_cos_b, t->cos_b, nb, m->stream))) return rc;
    if ((rc = upload(m->d_sin_b, t->sin_b, nb, m->stream))) return rc;
    if ((rc = upload(m->d_range, t->range_m, H, m->stream))) return rc;
    if ((rc = upload(m->d_nv_free, t->nv_free, H, m->stream))) return rc;
    if ((rc = upload(m->d_nv_occ, t->nv_occ, H, m->stream))) return rc;
    if ((rc = upload(m->d_cos_va, t->cos_va, nfan, m->stream))) return rc;
    if ((rc = upload(m->d_sin_va, t->sin_va, nfan, m->stream))) return rc;
    if ((rc = upload(m->d_col_to_beam, c2b.data(), c2b.size(), m->stream))) return rc;
    CU(cudaStreamSynchronize(m->stream));     // host vectors go out of scope
    DevTables &d = m->tab;
    d.H = t->H; d.W = t->W; d.n_beams = t->n_beams; d.nv_max = t->nv_max;
    d.free_step = t->free_step; d.occ_window = t->occ_window;
    d.beam_col = m->d_beam_col.p; d.cos_b = m->d_cos_b.p; d.sin_b = m->d_sin_b.p; d.range_m = m->d_range.p;
    d.nv_free = m->d_nv_free.p; d.nv_occ = m->d_nv_occ.p; d.cos_va = m->d_cos_va.p; d.sin_va = m->d_sin_va.p;
    d.col_to_beam = m->d_col_to_beam.p;
    m->have_tables = true;
    return ensure_frame_slots(m);
}

int s3d_ingest_batch_dev(s3d_map *m, const uint8_t *images_dev, int64_t n, const double *T_dev,
                         s3d_frame_stats *out, s3d_frame_stats *stats_dev)
{
    int rc = check_ready(m); if (rc) return rc;
    if (n < 0) return fail(S3D_EINVAL, "n < 0");
    if (n == 0) return 0;
    if ((rc = set_device(m))) return rc;
    if ((rc = reserve_frames(m, n))) return rc;
    DevStats *sd = reinterpret_cast<DevStats *>(stats_dev);
    if (!sd) { if ((rc = m->stats.ensure((size_t)n))) return rc; sd = m->stats.p; }
    if ((rc = run_frames(m, images_dev, n, T_dev, sd))) return rc;
    if (out) return finish_stats(m, sd, n, out);
    return 0;
}

int s3d_ingest_batch(s3d_map *m, const uint8_t *images, int64_t n, const double *T, s3d_frame_stats *out)
{
    int rc = check_ready(m); if (rc) return rc;
    if (n < 0) return fail(S3D_EINVAL, "n < 0");
    if (n == 0) return 0;
    if (!images || !T) return fail(S3D_EINVAL, "null input");
    if ((rc = set_device(m))) return rc;
    if ((rc = reserve_frames(m, n))) return rc;
    const size_t img_bytes = (size_t)m->tab.H * m->tab.W;
    if ((rc = m->stats.ensure((size_t)n))) return rc;
    // frames travel in pieces of `piece` frames so the device staging stays bounded
    const int64_t piece = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)((256u << 20) / std::max<size_t>(1, img_bytes))));
    if ((rc = m->img_dev.ensure(std::max<size_t>(16, img_bytes * (size_t)piece)))) return rc;
    if ((rc = m->T_dev.ensure(16 * (size_t)piece))) return rc;
    for (int64_t base = 0; base < n; base += piece) {
        const int64_t k = std::min<int64_t>(piece, n - base);
        if (img_bytes) CU(cudaMemcpyAsync(m->img_dev.p, images + (size_t)base * img_bytes, img_bytes * (size_t)k, cudaMemcpyHostToDevice, m->stream));
        CU(cudaMemcpyAsync(m->T_dev.p, T + base * 16, sizeof(double) * 16 * (size_t)k, cudaMemcpyHostToDevice, m->stream));
        if ((rc = run_frames(m, m->img_dev.p, k, m->T_dev.p, m->stats.p + base))) return rc;
    }
    return finish_stats(m, m->stats.p, n, out);
}

int s3d_ingest(s3d_map *m, const uint8_t *image, const double T[16], s3d_frame_stats *out)
{
    return s3d_ingest_batch(m, image, 1, T, out);
}

int s3d_reserve(s3d_map *m, uint64_t n_voxels)
{
    if (!m) return fail(S3D_EINVAL, "null map");
    int rc = set_device(m); if (rc) return rc;
    if ((rc = sync_counters(m))) return rc;
    return ensure_room(m, n_voxels);
}

int s3d_sync(s3d_map *m)
{
    if (!m) return fail(S3D_EINVAL, "null map");
    int rc = set_device(m); if (rc) return rc;
    return sync_counters(m);
}

void *s3d_stream(s3d_map *m) { return m ? (void *)m->stream : nullptr; }
uint64_t s3d_capacity(s3d_map *m) { return m ? m->cap : 0; }

int s3d_apply_updates(s3d_map *m, const int32_t *ijk, const double *delta, const uint8_t *adaptive, int64_t n)
{
    if (!m || !m->have_params) return fail(S3D_EINVAL, "map not ready (params)");
    if (n <= 0) return n == 0 ? 0 : fail(S3D_EINVAL, "n < 0");
    int rc = set_device(m); if (rc) return rc;
    std::vector<u64> keys;
    if ((rc = pack_keys_host(ijk, n, keys))) return rc;
    // round r holds the r-th occurrence of every key, so that each launch sees unique keys and
    // equal keys are applied in array order
    std::vector<int> round((size_t)n, 0);
    int n_rounds = 1;
    if (n > 1) {
        std::unordered_map<u64, int> seen;
        seen.reserve((size_t)n * 2);
        for (int64_t i = 0; i < n; ++i) { const int r = seen[keys[(size_t)i]]++; round[(size_t)i] = r; n_rounds = std::max(n_rounds, r + 1); }
    }
    std::vector<int64_t> start((size_t)n_rounds + 1, 0);
    for (int64_t i = 0; i < n; ++i) ++start[(size_t)round[(size_t)i] + 1];
    for (int r = 0; r < n_rounds; ++r) start[(size_t)r + 1] += start[(size_t)r];
    std::vector<u64> k2((size_t)n); std::vector<double> d2((size_t)n); std::vector<uint8_t> a2((size_t)n);
    std::vector<int64_t> cur(start.begin(), start.end() - 1);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t dst = cur[(size_t)round[(size_t)i]]++;
        k2[(size_t)dst] = keys[(size_t)i]; d2[(size_t)dst] = delta[i]; a2[(size_t)dst] = adaptive ? adaptive[i] : 1;
    }
    if ((rc = sync_counters(m))) return rc;
    if ((rc = ensure_room(m, start[1]))) return rc;     // at most one new voxel per distinct key
    if ((rc = upload(m->io_keys, k2.data(), (size_t)n, m->stream))) return rc;
    if ((rc = upload(m->io_vals, d2.data(), (size_t)n, m->stream))) return rc;
    if ((rc = upload(m->io_flags, a2.data(), (size_t)n, m->stream))) return rc;
    for (int r = 0; r < n_rounds; ++r) {
        const u64 cnt = (u64)(start[(size_t)r + 1] - start[(size_t)r]);
        if (!cnt) continue;
        k_apply_direct<<<(unsigned)((cnt + 255) / 256), 256, 0, m->stream>>>(
            m->io_keys.p + start[(size_t)r], m->io_vals.p + start[(size_t)r], m->io_flags.p + start[(size_t)r], cnt,
            m->table, m->cap - 1, m->p, m->mc);
    }
    CU(cudaGetLastError());
    m->ex_valid = false;
    return sync_counters(m);
}

int s3d_load(s3d_map *m, const int32_t *ijk, const double *log_odds, int64_t n)
{
    if (!m) return fail(S3D_EINVAL, "null map");
    if (n <= 0) return n == 0 ? 0 : fail(S3D_EINVAL, "n < 0");
    int rc = set_device(m); if (rc) return rc;
    std::vector<u64> keys;
    if ((rc = pack_keys_host(ijk, n, keys))) return rc;
    {
        std::unordered_map<u64, int> seen;
        seen.reserve((size_t)n * 2);
        for (int64_t i = 0; i < n; ++i)
            if (seen[keys[(size_t)i]]++) return fail(S3D_EINVAL, "s3d_load: duplicate key at index %lld", (long long)i);
    }
    if ((rc = sync_counters(m))) return rc;
    if ((rc = ensure_room(m, (u64)n))) return rc;
    if ((rc = upload(m->io_keys, keys.data(), (size_t)n, m->stream))) return rc;
    if ((rc = upload(m->io_vals, log_odds, (size_t)n, m->stream))) return rc;
    k_load<<<(unsigned)(((u64)n + 255) / 256), 256, 0, m->stream>>>(m->io_keys.p, m->io_vals.p, (u64)n, m->table, m->cap - 1, m->mc);
    CU(cudaGetLastError());
    m->ex_valid = false;
    return sync_counters(m);
}

int s3d_query(s3d_map *m, const int32_t *ijk, int64_t n, double *log_odds, uint8_t *found)
{
    if (!m) return fail(S3D_EINVAL, "null map");
    if (n <= 0) return n == 0 ? 0 : fail(S3D_EINVAL, "n < 0");
    int rc = set_device(m); if (rc) return rc;
    std::vector<u64> keys((size_t)n);
    std::vector<uint8_t> in_range((size_t)n, 1);
    for (int64_t i = 0; i < n; ++i) {
        const int a = ijk[3 * i], b = ijk[3 * i + 1], c = ijk[3 * i + 2];
        if (a < -KEY_BIAS || a >= KEY_BIAS || b < -KEY_BIAS || b >= KEY_BIAS || c < -KEY_BIAS || c >= KEY_BIAS) {
            in_range[(size_t)i] = 0; keys[(size_t)i] = pack_key(0, 0, 0);   // cannot be stored => absent
        } else keys[(size_t)i] = pack_key(a, b, c);
    }
    if ((rc = upload(m->io_keys, keys.data(), (size_t)n, m->stream))) return rc;
    if ((rc = m->io_vals.ensure((size_t)n))) return rc;
    if ((rc = m->io_flags.ensure((size_t)n))) return rc;
    k_query<<<(unsigned)(((u64)n + 255) / 256), 256, 0, m->stream>>>(m->io_keys.p, (u64)n, m->table, m->cap - 1, m->io_vals.p, m->io_flags.p);
    CU(cudaGetLastError());
    std::vector<uint8_t> f((size_t)n);
    CU(cudaMemcpyAsync(log_odds, m->io_vals.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    CU(cudaMemcpyAsync(f.data(), m->io_flags.p, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    for (int64_t i = 0; i < n; ++i) {
        if (!in_range[(size_t)i]) { log_odds[i] = 0.0; f[(size_t)i] = 0; }
        if (found) found[i] = f[(size_t)i];
    }
    return 0;
}

int s3d_count(s3d_map *m, uint64_t *count)
{
    if (!m || !count) return fail(S3D_EINVAL, "null argument");
    int rc = set_device(m); if (rc) return rc;
    rc = sync_counters(m);
    *count = m->count_known;
    return rc;
}

int s3d_bounds(s3d_map *m, int32_t kmin[3], int32_t kmax[3])
{
    if (!m) return fail(S3D_EINVAL, "null map");
    int rc = set_device(m); if (rc) return rc;
    rc = sync_counters(m);
    for (int q = 0; q < 3; ++q) { kmin[q] = m->mc_host->kmin[q]; kmax[q] = m->mc_host->kmax[q]; }
    return rc;
}

int s3d_clear(s3d_map *m)
{
    if (!m) return fail(S3D_EINVAL, "null map");
    int rc = set_device(m); if (rc) return rc;
    if ((rc = launch_fill_table(m, m->table, m->cap))) return rc;
    k_reset_ctr<<<1, 1, 0, m->stream>>>(m->mc);
    CU(cudaGetLastError());
    m->max_touched = 0;
    m->ex_valid = false;
    return sync_counters(m);
}

int s3d_export_begin(s3d_map *m, double thr_occ, double thr_free, uint32_t class_mask, uint64_t counts[3], uint64_t *n_out)
{
    if (!m || !m->have_params) return fail(S3D_EINVAL, "map not ready (params)");
    int rc = set_device(m); if (rc) return rc;
    if ((rc = sync_counters(m))) return rc;
    const size_t n = (size_t)std::max<u64>(1, m->count_known);
    if ((rc = m->ex_xyz.ensure(3 * n)) || (rc = m->ex_prob.ensure(n)) || (rc = m->ex_L.ensure(n)) ||
        (rc = m->ex_cls.ensure(n)) || (rc = m->ex_ijk.ensure(3 * n))) return rc;
    CU(cudaMemsetAsync(m->ex_counts, 0, sizeof(u64) * 4, m->stream));
    ExportOut o{m->ex_xyz.p, m->ex_prob.p, m->ex_L.p, m->ex_cls.p, m->ex_ijk.p, m->ex_counts};
    const int blocks = (int)std::min<u64>((m->cap + 255) / 256, (u64)m->n_sm * 8);
    k_export<<<blocks, 256, 0, m->stream>>>(m->table, m->cap, m->p.res, thr_occ, thr_free, class_mask, o);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(m->ex_counts_host, m->ex_counts, sizeof(u64) * 4, cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    if (counts) for (int c = 0; c < 3; ++c) counts[c] = m->ex_counts_host[c];
    m->ex_n = m->ex_counts_host[3];
    m->ex_valid = true;
    if (n_out) *n_out = m->ex_n;
    return 0;
}

int s3d_export_read(s3d_map *m, double *xyz, double *prob, int8_t *cls, int32_t *ijk, uint64_t n)
{
    if (!m || !m->ex_valid) return fail(S3D_EINVAL, "no staged export (call s3d_export_begin; the map must not change in between)");
    if (n != m->ex_n) return fail(S3D_EINVAL, "n=%llu does not match the staged export (%llu)", (unsigned long long)n, (unsigned long long)m->ex_n);
    if (n == 0) return 0;
    int rc = set_device(m); if (rc) return rc;
    if (xyz) CU(cudaMemcpyAsync(xyz, m->ex_xyz.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, m->stream));
    if (prob) CU(cudaMemcpyAsync(prob, m->ex_prob.p, sizeof(double) * n, cudaMemcpyDeviceToHost, m->stream));
    if (cls) CU(cudaMemcpyAsync(cls, m->ex_cls.p, n, cudaMemcpyDeviceToHost, m->stream));
    if (ijk) CU(cudaMemcpyAsync(ijk, m->ex_ijk.p, sizeof(int) * 3 * n, cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    return 0;
}

int s3d_export_read_xyzi32(s3d_map *m, float *xyzi, uint64_t n)
{
    if (!m || !m->ex_valid) return fail(S3D_EINVAL, "no staged export (call s3d_export_begin)");
    if (n != m->ex_n) return fail(S3D_EINVAL, "n does not match the staged export");
    if (n == 0) return 0;
    int rc = set_device(m); if (rc) return rc;
    if ((rc = m->ex_f32.ensure((size_t)n))) return rc;
    k_pack_xyzi32<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(m->ex_xyz.p, m->ex_prob.p, n, m->ex_f32.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(xyzi, m->ex_f32.p, sizeof(float4) * n, cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    return 0;
}

int s3d_dump(s3d_map *m, int32_t *ijk, double *log_odds, uint64_t cap, uint64_t *n_out)
{
    if (!m || !m->have_params) return fail(S3D_EINVAL, "map not ready (params)");
    uint64_t n = 0;
    const double inf = std::numeric_limits<double>::infinity();
    int rc = s3d_export_begin(m, inf, -inf, 7u, nullptr, &n);   // everything is UNKNOWN => all staged
    if (rc) return rc;
    if (n_out) *n_out = n;
    const u64 k = std::min<u64>(n, cap);
    if (k == 0) return 0;
    if (ijk) CU(cudaMemcpyAsync(ijk, m->ex_ijk.p, sizeof(int) * 3 * k, cudaMemcpyDeviceToHost, m->stream));
    if (log_odds) CU(cudaMemcpyAsync(log_odds, m->ex_L.p, sizeof(double) * k, cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    return 0;
}

} // extern "C"
