// sonar3d.cu -- sm_100a kernels + C-ABI (include/sonar3d.h) of the sonar -> voxel hot path.
//
// Reference behaviour: luckkim123/sonar_3d_reconstruction scripts/3d_mapper.py
//   :387-483 process_sonar_ray   -> k_expand (first-hit scan, expansion, keys, block combiner; K1+K2+K3 of SURVEY 2.2)
//   :485-567 process_sonar_image -> chunk dedupe table + k_apply_chunk (K3, K4)
//   :83-115  update_voxel        -> apply_one()
//   :117-188 queries / export    -> k_query, k_export (K5, K6)
//
// Data layout in HBM (DESIGN.md has the full account):
//   voxel table    Slot[cap]    16 B {packed key, fp64 log-odds}, open addressing, probed by 32-byte sector
//   chunk dedupe   u64 keys[C] + lanes[C][16]: one entry per voxel touched by a chunk of up to 16
//                  consecutive frames, one {n_occ, n_free} counter lane per frame (u32 = 16+16 bits
//                  normally, u64 = 32+32 bits after an overflow)
// No tensor cores: the path has no dense contraction; it is integer/fp64 scatter work.
#include "../../include/sonar3d.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <deque>
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
#ifndef S3D_GF
#define S3D_GF S3D_CHUNK_FRAMES
#endif
constexpr int GF = S3D_GF;                                  // frames per chunk = counter lanes per dedupe entry
static_assert(GF == S3D_CHUNK_FRAMES && (GF == 16 || GF == 32), "frames per chunk: 16 or 32 (one mask bit per frame)");
constexpr int N_CHUNK_BUF = 4;                              // chunk dedupe buffers (a single map cycles through 3 of them)
constexpr u32 ERR_KEYRANGE = 1u, ERR_TABLEFULL = 2u;        // fatal
constexpr u32 ERR_ROUTE_FULL = 4u, ERR_ROUTE_TIMEOUT = 8u;  // fatal (routed map): a peer inbox overflowed / a peer never signalled
constexpr u32 ERR_INTERNAL = 32u;                           // a bounds guard tripped (should be impossible): reported, not faulted
constexpr u32 ERR_VERIFY = 16u;                             // S3D_VERIFY_FAST: an accepted fp32 estimate differed from the fp64 key
constexpr u32 ABORT_SCRATCH = 1u, ABORT_TABLE = 2u;         // retryable: the host enlarges and re-runs the chunk
constexpr u32 ABORT_NARROW = 4u;                            // retryable: a 16-bit sample count overflowed -> wide lanes

// A dedupe entry holds one counter lane per frame of the chunk: {n_occ, n_free} of the voxel in
// that frame.  Narrow lanes (u32 = 16 + 16 bits) are the normal format -- half the bytes of the
// wide ones; a voxel that collects more than 65535 samples of one kind in one frame (voxels far
// larger than the range resolution) makes the chunk re-run with wide lanes (u64 = 32 + 32 bits).
template <typename CT> struct Lane;
template <> struct Lane<u32> {
    static constexpr bool narrow = true;
    static __device__ __forceinline__ u32 n_occ(u32 c) { return c >> 16; }
    static __device__ __forceinline__ u32 n_free(u32 c) { return c & 0xffffu; }
    static __device__ __forceinline__ u32 make(u32 n_occ, u32 n_free) { return (n_occ << 16) | n_free; }
    static __device__ __forceinline__ bool overflows(u32 old, u32 inc)
    { return (((old & 0xffffu) + (inc & 0xffffu)) | ((old >> 16) + (inc >> 16))) > 0xffffu; }
};
template <> struct Lane<u64> {
    static constexpr bool narrow = false;
    static __device__ __forceinline__ u32 n_occ(u64 c) { return (u32)(c >> 32); }
    static __device__ __forceinline__ u32 n_free(u64 c) { return (u32)(c & 0xffffffffu); }
    static __device__ __forceinline__ u64 make(u32 n_occ, u32 n_free) { return ((u64)n_occ << 32) | (u64)n_free; }
    static __device__ __forceinline__ bool overflows(u64, u64) { return false; }
};

struct __align__(16) Slot { u64 key; double val; };

struct DevParams {
    double res, inv_res, lo_occ, lo_free, lo_min, lo_max, a_thr, a_ratio, zmin;
    double l_skip;   // log-odds above which prob > a_thr for sure (+inf = never skip the exp)
    double scale0;   // factor of an occupied update on a voxel with L == 0 (prob == 0.5 exactly): (0.5 / a_thr) * a_ratio, or 1 when 0.5 > a_thr
    int adaptive, zfilter, thr;
    u32 fr_hi_lo, fr_hi_span;   // high words bounding the fractional parts the fast quantiser accepts
};

struct DevTables {
    int H, W, n_beams, nv_max, free_step, occ_window;
    const int *beam_col;
    const double *cos_b, *sin_b, *range_m;
    const int *nv_free, *nv_occ;
    const double *cos_va, *sin_va;
    const float2 *csva32;     // {cos, sin}(va) rounded to fp32 (fast path)
    int col_step;             // beam_col[b] == b * col_step for every beam (0 = irregular: no strip staging)
    int n_trig;               // entries of the cos_va / sin_va / csva32 tables
};

struct MapCtr {       // device-resident map counters
    u64 count;        // live voxels
    u64 abort_seq;    // first chunk (sequence number) that asked for a retry; ~0 = none
    int kmin[3], kmax[3];
    u32 err;          // fatal flags
    u32 abort;        // retry flags; while set, the kernels of chunk abort_seq and of every later chunk
                      // return without side effects (earlier chunks run to completion)
    u32 last_new;     // voxels inserted by the last applied chunk
    u32 last_unique;  // voxels touched by the last applied chunk
    u64 life_count;   // debug counters: keys in the lifetime sample-count table
    u64 life_max;     // debug counters: largest lifetime sample count of any voxel (3d_mapper.py:578)
    u64 route_sent;   // routed map: 16-byte records this rank has written into peers' inboxes (NVLink traffic)
    u64 probes;       // voxel-table probes so far: one per voxel per applied chunk
};

struct ChunkCtr {     // working counters of the chunk in flight (chunks are serialised on the stream)
    u32 n_unique;     // dedupe entries created for the chunk (k_expand, k_shard_merge)
    u32 ticket;       // last-block-out election (k_apply_chunk)
    u32 xticket;      // last-block-out election of k_expand (routed map)
    u32 mticket;      // last-block-out election of k_route_merge (routed map)
    u32 neu[GF];      // voxels first inserted at frame f of the chunk
    // debug counters (3d_mapper.py:549-551, :575-585), only written when they are enabled
    u32 dmax[GF];     // largest per-voxel sample count of frame f
    u32 dgt10[GF];    // voxels with more than 10 samples in frame f
    u32 life_new;     // keys first inserted into the lifetime table by this chunk
    u32 pad_;
    u64 dlife[GF];    // largest lifetime sample count reached by a voxel touched in frame f
};

// == s3d_frame_stats
struct DevStats { u64 n_occ, n_free, n_voxels, n_samples, max_in_frame, n_gt10, max_total, reserved; };
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

// floor(w / res) exactly as IEEE division would give it (3d_mapper.py:63-65).  The reciprocal
// product decides whenever it is not within 1e-6 of an integer (its error is < 1e-9 for
// in-range keys); only the rare near-integer case pays for the division.
__device__ __forceinline__ bool voxel_index(double w, double res, double inv_res, int &out)
{
    const double q = w * inv_res;
    const int k = __double2int_rd(q);                 // floor; saturates, NaN -> 0
    const double fr = q - (double)k;
    if (fr > 1e-6 && fr < 1.0 - 1e-6 && k > -KEY_BIAS && k < KEY_BIAS - 1) { out = k; return true; }
    if (!(fabs(q) < (double)(KEY_BIAS - 1))) return false;   // out of range, NaN or inf
    out = (int)floor(__ddiv_rn(w, res));
    return true;
}

// owning rank of a voxel when the map is sharded over `world` GPUs
__device__ __forceinline__ u32 key_owner(u64 key, u32 world) { return (u32)((mix64(key) >> 40) % world); }

// cheap 32-bit mix for the chunk dedupe table (the voxel table keeps the stronger mix64)
__device__ __forceinline__ u32 mix32(u64 key)
{
    u32 h = (u32)key ^ ((u32)(key >> 32) * 0x9E3779B1u);
    h ^= h >> 15; h *= 0x85EBCA77u;
    h ^= h >> 13; h *= 0xC2B2AE3Du;
    h ^= h >> 16;
    return h;
}

// ------------------------------------------------------------------------------------ K1+K2+K3
// One block per (group of EX_WARPS adjacent processed beams, frame of the chunk); warp w lists
// and walks beam w.  Samples are merged twice before they reach HBM-side structures:
//   1. in a block-local combiner in shared memory (open addressing, 32-bit voxel key relative
//      to the voxel of the sonar origin, 16+16-bit occupied/free counts): adjacent range bins,
//      vertical steps and beams fall into the same voxels, so a block's ~5 k samples collapse
//      to ~1 k entries at shared-memory atomic cost;
//   2. the combiner is flushed, entry by entry, into the chunk's dedupe table in global memory
//      (one entry per voxel touched by the chunk, one counter lane per frame).
// (measured at cfg2 / cfg3, frames/s: 16 warps x 2 blocks per SM with a 8192-entry combiner 221 k / 29.0 k; 8 x 4 with 4096
// entries 211 k / 29.0 k; 4 x 8 with 2048 151 k / 22.8 k; 32 x 1 with 16384 177 k)
#ifndef S3D_EX_WARPS
#define S3D_EX_WARPS 16
#endif
constexpr int EX_WARPS = S3D_EX_WARPS;       // warps per block
constexpr int EX_MAXB = 16;                  // beams per tile at most (a 16-byte image strip at bearing step 1)
constexpr int EX_THREADS = EX_WARPS * 32;
#ifndef S3D_EX_BPS
#define S3D_EX_BPS 2
#endif
constexpr int EX_BPS = S3D_EX_BPS;           // resident blocks per SM the kernel is compiled for
#ifndef S3D_LT_BITS
#define S3D_LT_BITS 13        // log2 of the block combiner's entries
#endif
#ifndef S3D_EX_ILP
#define S3D_EX_ILP 2
#endif
constexpr int EX_ILP = S3D_EX_ILP;           // samples per lane per pass
constexpr int EX_PASS = 32 * EX_ILP;         // samples per warp per pass
constexpr int EX_ROUND = (S3D_LT_BITS >= 12 ? 12 : 6) / EX_ILP;   // passes between two block-wide flush votes
constexpr int EX_ROUND_SAMPLES = EX_WARPS * EX_PASS * EX_ROUND;
constexpr int LT_BITS = S3D_LT_BITS;
constexpr int LT_CAP = 1 << LT_BITS;         // combiner entries
constexpr int LT_LIMIT = LT_CAP * 3 / 4;     // load bound
constexpr int LT_MAX_SAMPLES = 0xFFFF;       // 16-bit counts cannot overflow between two flushes
constexpr u32 LT_EMPTY = 0xFFFFFFFFu;
constexpr int LK_BITS = 10;                  // local key: 3 x 10 bits around the sonar-origin voxel
constexpr int LK_HALF = 1 << (LK_BITS - 1);
constexpr int FL_ILP = 4;                    // combiner entries per thread per flush batch
constexpr u32 SCRATCH_PROBE_LIMIT = 512;
static_assert(EX_ROUND_SAMPLES <= LT_LIMIT, "a round must fit the combiner");

struct __align__(16) Fan { int off; u32 code; float rho; u32 pad_; };   // code = r | nv << 16 | occupied << 31; rho = range / res
constexpr int TMA_BOX_ROWS = 256;            // rows per TMA box (the hardware limit of a box dimension)
constexpr int TMA_STRIP_BYTES = 16;          // bytes per image row a tile stages: its beams' columns
constexpr float MAGICF = 12582912.0f;        // 1.5 * 2^23: adding it (round down) leaves floor(q) in the low mantissa bits

// ---- routed map (one process per GPU): a rank expands its slice of the beams and its combiner
// flush writes the entries of voxels it does not own straight into the owner's inbox over
// NVLink peer memory (plain 16-byte stores, slots handed out by a local atomic) -- the exchange
// is part of the expansion kernel, there is no separate pack / all-to-all / unpack.  Every rank
// exports one "exchange block": a header of flags followed by ROUTE_DEPTH (chunk parity) x world (source)
// inbox regions of `cap` records.  Ordering: after its k_expand of chunk c a rank publishes, per
// peer, the record count and the sequence number c+1 (system-scope fence in between); the owner
// waits for all sources, merges the records into its dedupe table, and acknowledges, which
// lets the sources reuse that parity for chunk c + ROUTE_DEPTH.
constexpr int ROUTE_MAX_WORLD = 64;
constexpr int ROUTE_DEPTH = 16;           // inbox regions per source: chunk c uses region c % ROUTE_DEPTH ("parity")
struct __align__(16) RouteRec { u64 key; u32 counts; u32 frame; };      // counts = n_occ << 16 | n_free
struct RouteHdr {
    u64 flag_count[ROUTE_DEPTH][ROUTE_MAX_WORLD];   // [parity][source]: records the source wrote for the chunk
    u64 flag_seq[ROUTE_DEPTH][ROUTE_MAX_WORLD];     //                   ... and the chunk's sequence number + 1
    u64 ack_seq[ROUTE_DEPTH][ROUTE_MAX_WORLD];      // [parity][owner]: that owner has merged my records of sequence number - 1
};
struct RouteCtx {
    u32 world, rank, parity;
    u64 cap;                              // records per (parity, source) inbox region
    unsigned char *const *peer;           // [world] exchange blocks of all ranks (device array; peer[rank] is mine)
    u32 *cursor;                          // [ROUTE_DEPTH][ROUTE_MAX_WORLD] records written so far per (parity, owner); local
    u64 seq;                              // sequence number of the chunk among the routed chunks (same on every rank)
    u64 timeout_ns;                       // a peer silent for this long raises ERR_ROUTE_TIMEOUT instead of hanging
};

__device__ __forceinline__ u64 global_ns() { u64 t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// S3D_TRACE=1 (development): first-block-in / last-block-out GPU timestamps of the pipeline kernels
// of each chunk, slot[0] = earliest start, slot[1] = latest end (ns, %globaltimer)
__device__ __forceinline__ void trace_begin(u64 *slot) { if (slot && threadIdx.x == 0) atomicMin(slot, global_ns()); }
__device__ __forceinline__ void trace_end(u64 *slot) { if (slot && threadIdx.x == 0) atomicMax(slot + 1, global_ns()); }

// spin until *w >= want (a word a peer writes into this rank's exchange block)
__device__ __forceinline__ void route_wait_word(const u64 *word, u64 want, u64 timeout_ns, MapCtr *mc)
{
    const volatile u64 *w = word;
    const u64 t0 = global_ns();
    while (*w < want) {
        if (global_ns() - t0 > timeout_ns) { atomicOr(&mc->err, ERR_ROUTE_TIMEOUT); break; }
        __nanosleep(100);
    }
}

// publish to every peer how many records it got from this rank for the chunk, then the chunk's
// sequence number + 1 (threads o < world of one block; all record stores are fenced before)
__device__ __forceinline__ void route_signal(const RouteCtx &rt, MapCtr *mc)
{
    const u32 o = threadIdx.x;
    if (o >= rt.world || o == rt.rank) return;
    u32 *cur = &rt.cursor[rt.parity * ROUTE_MAX_WORLD + o];
    const u64 cnt = min((u64)atomicAdd(cur, 0u), rt.cap);
    if (mc) atomicAdd(&mc->route_sent, cnt);
    *cur = 0;                                              // this parity's next use is ROUTE_DEPTH chunks away, on this stream
    RouteHdr *h = reinterpret_cast<RouteHdr *>(rt.peer[o]);
    __threadfence_system();
    *(volatile u64 *)&h->flag_count[rt.parity][rt.rank] = cnt;
    __threadfence_system();
    *(volatile u64 *)&h->flag_seq[rt.parity][rt.rank] = rt.seq + 1;
}

__device__ __forceinline__ RouteRec *route_inbox(const RouteCtx &rt, u32 owner, u32 parity, u32 source)
{
    return reinterpret_cast<RouteRec *>(rt.peer[owner] + sizeof(RouteHdr)) + ((size_t)parity * rt.world + source) * rt.cap;
}

// one record to `owner` (the rare per-sample path; the flush hands out slots per warp)
__device__ __forceinline__ void route_send_one(const RouteCtx &rt, MapCtr *mc, u32 owner, u64 key, u32 counts, u32 frame)
{
    const u32 pos = atomicAdd(&rt.cursor[rt.parity * ROUTE_MAX_WORLD + owner], 1u);
    if (pos >= rt.cap) { atomicOr(&mc->err, ERR_ROUTE_FULL); return; }
    RouteRec r; r.key = key; r.counts = counts; r.frame = frame;
    route_inbox(rt, owner, rt.parity, rt.rank)[pos] = r;
}

struct ExpandArgs {
    const uint8_t *imgs; size_t img_stride;
    const double *T;             // [g][16]
    DevTables tab; DevParams p;
    u64 *skeys; void *scnt; u32 smask;  // chunk dedupe table: keys[C], counter lanes[C][GF] (u32 or u64 lanes)
    ChunkCtr *cc;
    DevStats *stats;             // [g]
    MapCtr *mc;
    int beam_lo, beam_hi;        // processed beams [beam_lo, beam_hi) are expanded (a rank's slice when sharded)
    int bpb;                     // beams per tile, 1..EX_MAXB
    int n_frames;                // frames of the chunk covered by this launch
    int use_tma;                 // the tile's image strip is staged with one 2-D TMA box per 256 rows (else plain loads)
    int fast32_ok;               // 0: every sample takes the fp64 path (S3D_NO_FAST32)
    int box_rows;                // rows per TMA box: min(256, H)
    u32 own_rank, own_world;     // own_world > 1: keep only the voxels this rank owns (replicated expansion)
    RouteCtx rt;                 // rt.world > 1: voxels of other owners are routed to them (routed map)
    u64 seq;                     // chunk sequence number (for abort bookkeeping)
    u64 *trace;                  // S3D_TRACE: {start, end} of this launch, or null
    u64 *marks;                  // S3D_MARKS (fault localisation): mapped host counters {blocks started, blocks finished}, or null
};

__device__ __forceinline__ void raise_abort(MapCtr *mc, u32 why, u64 seq)
{
    atomicOr(&mc->abort, why);
    atomicMin(&mc->abort_seq, seq);
}

// The dedupe table is probed by 32-byte sector: a bucket is 4 consecutive keys, loaded with two
// 16-byte ld.cg (one sector), so a lookup is one L2 round trip unless the bucket is full of other
// keys.
struct Bucket { ulonglong2 a, b; };

__device__ __forceinline__ Bucket load_bucket(const u64 *skeys, u32 base)
{
    Bucket k;
    k.a = __ldcg(reinterpret_cast<const ulonglong2 *>(skeys + base));
    k.b = __ldcg(reinterpret_cast<const ulonglong2 *>(skeys + base + 2));
    return k;
}

// find `key` in the dedupe table or create its entry; `k` holds the already loaded home bucket.
// Returns the slot, or ~0u if the table is over-loaded.  `created` says this call made the entry.
__device__ __forceinline__ u32 dedupe_slot(u64 *skeys, u32 smask, u64 key, u32 base, Bucket k, bool &created)
{
    created = false;
    for (u32 probe = 0; probe < SCRATCH_PROBE_LIMIT; ++probe) {
        const u64 w[4] = {k.a.x, k.a.y, k.b.x, k.b.y};
        int hit = -1;
#pragma unroll
        for (int j = 3; j >= 0; --j) if (w[j] == key) hit = j;
        if (hit >= 0) return base + (u32)hit;
        int hole = -1;
#pragma unroll
        for (int j = 3; j >= 0; --j) if (w[j] == EMPTY_KEY) hole = j;
        if (hole >= 0) {
            const u64 cur = atomicCAS(&skeys[base + hole], EMPTY_KEY, key);
            if (cur == EMPTY_KEY) { created = true; return base + (u32)hole; }
            if (cur == key) return base + (u32)hole;
            k = load_bucket(skeys, base);            // another key took the hole: look at the bucket again
            continue;
        }
        base = (base + 4) & smask;                   // bucket full of other keys
        k = load_bucket(skeys, base);
    }
    return ~0u;
}

__device__ __forceinline__ u32 dedupe_home(u64 key, u32 smask) { return mix32(key) & smask & ~3u; }

// add n_occ / n_free samples to lane g of the entry of `key`; returns 1 if this call created the entry
// CHECK: watch narrow lanes for overflow (costs waiting for the atomic's return value); off when
// the host has proved from the geometry tables that no voxel can collect 2^16 samples in a frame
template <typename CT, bool CHECK>
__device__ __forceinline__ u32 dedupe_add(u64 *skeys, CT *scnt, u32 smask, MapCtr *mc, u64 seq, u64 key, u32 home,
                                          const Bucket &k, int g, u32 n_occ, u32 n_free)
{
    bool created;
    const u32 slot = dedupe_slot(skeys, smask, key, home, k, created);
    if (slot == ~0u) { raise_abort(mc, ABORT_SCRATCH, seq); return 0u; }   // the host enlarges the table and retries
    const CT inc = Lane<CT>::make(n_occ, n_free);
    if (Lane<CT>::narrow && CHECK) {
        const CT old = atomicAdd(&scnt[(size_t)slot * GF + g], inc);
        if (Lane<CT>::overflows(old, inc)) raise_abort(mc, ABORT_NARROW, seq);   // re-run with wide lanes
    } else {
        atomicAdd(&scnt[(size_t)slot * GF + g], inc);
    }
    return created ? 1u : 0u;
}

// a sample that does not fit the block combiner (more than 2^9 voxels from the sonar origin, or
// a transform too large for the fast quantiser) goes to the dedupe table on its own
template <typename CT, bool CHECK, bool ROUTE>
__device__ __noinline__ void commit_direct(u64 *skeys, CT *scnt, u32 smask, ChunkCtr *cc, MapCtr *mc, u64 seq, u64 key,
                                           int g, bool occ, RouteCtx rt)
{
    if (ROUTE && rt.world > 1) {
        const u32 owner = key_owner(key, rt.world);
        if (owner != rt.rank) { route_send_one(rt, mc, owner, key, occ ? 0x10000u : 1u, (u32)g); return; }
    }
    const u32 home = dedupe_home(key, smask);
    if (dedupe_add<CT, CHECK>(skeys, scnt, smask, mc, seq, key, home, load_bucket(skeys, home), g, occ ? 1u : 0u, occ ? 0u : 1u))
        atomicAdd(&cc->n_unique, 1u);
}

// floor(w / res) for samples of a block whose transform is known to keep |w / res| < 2^30:
// the reciprocal product, floored by a round-down add of 1.5 * 2^52 (the integer lands in the
// low word), decides unless it is within 1e-6 of an integer; then the IEEE division does.
__device__ __forceinline__ bool quantise(const DevParams &p, double w, bool fast, int &k)
{
    if (fast) {
        const double MAGIC = 6755399441055744.0;
        const double q = __dmul_rn(w, p.inv_res);
        const double t = __dadd_rd(q, MAGIC);
        const double fr = __dadd_rn(q, -__dadd_rn(t, -MAGIC));        // q - floor(q), exact
        k = __double2loint(t);
        if ((u32)(__double2hiint(fr) - p.fr_hi_lo) < p.fr_hi_span) return true;
    }
    return voxel_index(w, p.res, p.inv_res, k);
}

// the key range voxel_index accepts: -2^20 < k < 2^20 - 1 on every axis
__device__ __forceinline__ bool key_in_range(int ki, int kj, int kk)
{
    const u32 span = 2u * KEY_BIAS - 2u;
    return (u32)(ki + KEY_BIAS - 1) < span && (u32)(kj + KEY_BIAS - 1) < span && (u32)(kk + KEY_BIAS - 1) < span;
}

// Block-wide: move every combiner entry into the chunk dedupe table and leave the combiner
// empty.  All threads of the block call it after a barrier that follows the last insert;
// live[0 .. *s_count) lists the slots in use (appended as the entries were created).
template <typename CT, bool CHECK, bool ROUTE>
__device__ __forceinline__ void flush_combiner(const ExpandArgs &a, u32 *tkey, u32 *tcnt, unsigned short *live,
                                               volatile u32 *s_count, const int (&o)[3], int g, u32 &emitted)
{
    const int tid = threadIdx.x, lane = tid & 31;
    int n = (int)*s_count;                       // live[0..n) lists the combiner slots in use
    if (n > LT_CAP) { atomicOr(&a.mc->err, ERR_INTERNAL); n = LT_CAP; }
    // FL_ILP entries per thread at a time: home buckets are fetched together, then resolved
    for (int base = 0; base < n; base += EX_THREADS * FL_ILP) {
        u64 key[FL_ILP]; u32 inc[FL_ILP], home[FL_ILP]; bool ok[FL_ILP]; Bucket bk[FL_ILP];
        u32 dst[FL_ILP];                      // routed map: owner the entry is sent to, or ~0u
        u32 made = 0;
#pragma unroll
        for (int j = 0; j < FL_ILP; ++j) {
            const int e = base + j * EX_THREADS + tid;
            ok[j] = false; key[j] = 0; inc[j] = 0; home[j] = 0; dst[j] = ~0u;
            bk[j].a = bk[j].b = make_ulonglong2(0ull, 0ull);
            int idx = e < n ? (int)live[e] : -1;
            if (idx >= LT_CAP) { atomicOr(&a.mc->err, ERR_INTERNAL); idx = -1; }
            if (idx >= 0) {
                const u32 lk = tkey[idx], c = tcnt[idx];
                tkey[idx] = LT_EMPTY; tcnt[idx] = 0;
                const int ki = (int)(lk & (2 * LK_HALF - 1)) - LK_HALF + o[0];
                const int kj = (int)((lk >> LK_BITS) & (2 * LK_HALF - 1)) - LK_HALF + o[1];
                const int kk = (int)(lk >> (2 * LK_BITS)) - LK_HALF + o[2];
                if (!key_in_range(ki, kj, kk)) atomicOr(&a.mc->err, ERR_KEYRANGE);
                else {
                    key[j] = pack_key(ki, kj, kk);
                    // replicated expansion: every rank computes every sample and keeps the voxels it owns
                    const u32 owner = (ROUTE && a.rt.world > 1) ? key_owner(key[j], a.rt.world) : a.rt.rank;
                    if (owner != a.rt.rank) {
                        dst[j] = owner; inc[j] = c;
                        emitted += (c >> 16) + (c & 0xffffu);
                    } else if (a.own_world <= 1 || key_owner(key[j], a.own_world) == a.own_rank) {
                        ok[j] = true;
                        inc[j] = c;
                        emitted += (c >> 16) + (c & 0xffffu);
                        home[j] = dedupe_home(key[j], a.smask);
                        bk[j] = load_bucket(a.skeys, home[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < FL_ILP; ++j)
            if (ok[j]) made += dedupe_add<CT, CHECK>(a.skeys, static_cast<CT *>(a.scnt), a.smask, a.mc, a.seq, key[j], home[j], bk[j],
                                              g, inc[j] >> 16, inc[j] & 0xffffu);
        // entries created for the chunk: one reduction per warp per batch (nobody waits for it)
        made = __reduce_add_sync(0xffffffffu, made);
        if (lane == 0 && made) atomicAdd(&a.cc->n_unique, made);
        // routed map: entries of other owners go into their inboxes; lanes with the same owner
        // take consecutive slots from one local atomic, the records are plain peer-memory stores
        if (ROUTE && a.rt.world > 1) {
#pragma unroll
            for (int j = 0; j < FL_ILP; ++j) {
                const bool send = dst[j] != ~0u;
                if (!__any_sync(0xffffffffu, send)) continue;
                const u32 peers = __match_any_sync(0xffffffffu, dst[j]);
                const int leader = __ffs(peers) - 1;
                u32 at = 0;
                if (send && lane == leader) at = atomicAdd(&a.rt.cursor[a.rt.parity * ROUTE_MAX_WORLD + dst[j]], (u32)__popc(peers));
                at = __shfl_sync(0xffffffffu, at, leader);
                if (send) {
                    const u32 pos = at + __popc(peers & ((1u << lane) - 1));
                    if (pos >= a.rt.cap) atomicOr(&a.mc->err, ERR_ROUTE_FULL);
                    else {
                        RouteRec r; r.key = key[j]; r.counts = inc[j]; r.frame = (u32)g;
                        route_inbox(a.rt, dst[j], a.rt.parity, a.rt.rank)[pos] = r;
                    }
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) *s_count = 0;
    __syncthreads();
}

// Routed map: the last block of the grid to get here tells the peers that their records of this
// chunk are complete.  Block-wide; every block calls it exactly once.
template <bool ROUTE>
__device__ __forceinline__ void expand_finish(const ExpandArgs &a)
{
    trace_end(a.trace);
    if (a.marks && threadIdx.x == 0) atomicAdd_system(a.marks + 1, 1ull);
    if (!ROUTE || a.rt.world <= 1) return;
    __shared__ bool s_last;
    __threadfence_system();                     // this thread's record stores, before the ticket
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&a.cc->xticket, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    route_signal(a.rt, a.mc);
    if (threadIdx.x == 0) a.cc->xticket = 0;
}

// dynamic shared memory of k_expand: the combiner (keys, counts, live list) -- whose first bytes double
// as the landing zone of the tile's image strip while the fans are listed -- and the fan lists
__host__ __device__ inline size_t expand_smem_bytes(int H, int free_step, int occ_window, int bpb)
{
    const int max_f = (H + free_step - 1) / free_step;
    return sizeof(u32) * 2 * LT_CAP + sizeof(unsigned short) * LT_CAP +
           sizeof(Fan) * (size_t)bpb * (size_t)(max_f + occ_window + 1);
}
__host__ __device__ inline int strip_box_rows(int H) { return H < TMA_BOX_ROWS ? (H > 0 ? H : 1) : TMA_BOX_ROWS; }
__host__ __device__ inline size_t strip_smem_bytes(int H)
{
    const int br = strip_box_rows(H);
    return (size_t)((H + br - 1) / br) * br * TMA_STRIP_BYTES;
}

// ---- mbarrier / TMA (sm_100a PTX)
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one 2-D box {TMA_STRIP_BYTES x TMA_BOX_ROWS} of the chunk's frames (a [rows][W] byte tensor) -> shared memory
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, u64 *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// K1 + K2 + K3.  A *tile* is (up to EX_MAXB adjacent processed beams, one frame); blocks walk the
// launch's tiles with a grid stride.  Per tile:
//   stage    the tile's 16-byte-wide image strip [H rows] is brought into shared memory with one
//            2-D TMA box per 256 rows (cp.async.bulk.tensor + mbarrier; the strip lands on the
//            combiner's memory, which is idle until the walk) -- the frames are read here, once,
//            in full sectors.  Shapes that cannot be described as a strip fall back to plain loads;
//   list     warp w finds its beam's first above-threshold bin (:406-409) and lists the beam's
//            fans in walking order;
//   walk     the flattened (fan, vertical step) space of the tile's beams is cut into passes dealt
//            round-robin to the warps.  Per sample the voxel index relative to the voxel of the sonar
//            origin is first estimated in fp32: q = rho * (a_b * cos(va) + c * sin(va)) + f0 per axis
//            (rho = range / res; a_b, c, f0 are per-beam / per-frame constants prepared in fp64).  The
//            estimate is accepted when its fractional part is further than a proven error band from
//            both neighbouring integers (and the z filter's threshold, if any); the rare rest, and
//            every sample of a frame whose transform does not satisfy the bound's premises, is
//            evaluated in fp64 in numpy's operation order (:434-440, :63-65), which is what decides
//            in the reference.  So every key is the reference's key; the fp32 path only skips work;
//   combine / flush  as described above.
// ROUTE: compiled with the routed-map code (records of remote owners go to their inboxes); the
// single-map instantiation does not carry it.  VERIFY (tests): every sample also takes the fp64
// path and a fast-path key that differs from it raises ERR_VERIFY.
template <typename CT, bool CHECK, bool ROUTE, bool VERIFY>
__global__ void __launch_bounds__(EX_THREADS, EX_BPS)
k_expand(const ExpandArgs a, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(128) unsigned char s_raw[];
    u32 *tkey = reinterpret_cast<u32 *>(s_raw);
    u32 *tcnt = tkey + LT_CAP;
    unsigned short *live = reinterpret_cast<unsigned short *>(tcnt + LT_CAP);
    Fan *fans_all = reinterpret_cast<Fan *>(live + LT_CAP);
    const DevTables &tab = a.tab;
    const int H = tab.H, W = tab.W;
    const int max_f = (H + tab.free_step - 1) / tab.free_step;        // free candidates per beam
    const int nf_max = max_f + tab.occ_window;                        // fans per beam
    __shared__ __align__(16) double s_T[12];
    __shared__ __align__(8) u64 s_bar;
    __shared__ int s_o[3], s_fast, s_fast32, s_tot[EX_MAXB], s_nfan[EX_MAXB], s_pfirst[EX_MAXB + 1];
    __shared__ double s_cb[EX_MAXB], s_sb[EX_MAXB];
    __shared__ float s_ab[EX_MAXB][4], s_c[3], s_f0[3], s_kband, s_zq, s_zslack;
    __shared__ u32 s_count, s_abort;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u32 lt_mask = (1u << lane) - 1;
    // (a retry asked for by this chunk or an earlier one; a later chunk's flag does not stop this one)
    if (tid == 0) {
        s_abort = (__ldcg(&a.mc->abort) != 0u && a.seq >= __ldcg(&a.mc->abort_seq)) ? 1u : 0u;
        s_count = 0;
        if (a.use_tma) mbar_init(&s_bar, 1);
    }
    trace_begin(a.trace);
    if (a.marks && tid == 0) atomicAdd_system(a.marks, 1ull);
    for (int i = tid; i < LT_CAP / 4; i += EX_THREADS) {
        reinterpret_cast<uint4 *>(tkey)[i] = make_uint4(LT_EMPTY, LT_EMPTY, LT_EMPTY, LT_EMPTY);
        reinterpret_cast<uint4 *>(tcnt)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (a.use_tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (s_abort) {                      // stay side-effect free (block-uniform)
        expand_finish<ROUTE>(a);
        return;
    }
    const int n_slice = a.beam_hi - a.beam_lo;
    const int tiles_x = (n_slice + a.bpb - 1) / a.bpb;
    const int n_tiles = tiles_x * a.n_frames;
    const int n_box = (H + a.box_rows - 1) / a.box_rows;
    u32 tma_parity = 0;
    volatile u32 *v_count = &s_count;
    volatile u32 *v_tkey = tkey;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int g = tile / tiles_x;
        const int beam0 = a.beam_lo + (tile - g * tiles_x) * a.bpb;
        const int nb = min(a.bpb, a.beam_hi - beam0);                 // beams of this tile
        const uint8_t *img = a.imgs + (size_t)g * a.img_stride;
        // ---- stage: image strip by TMA, frame constants
        if (tid == 0 && a.use_tma) {
            // (the combiner region was last touched through the generic proxy: order it before the async proxy's writes)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s_bar, (u32)(n_box * a.box_rows * TMA_STRIP_BYTES));
            for (int q = 0; q < n_box; ++q)
                tma_load_2d(s_raw + (size_t)q * a.box_rows * TMA_STRIP_BYTES, &tmap, (beam0 * tab.col_step) & ~(TMA_STRIP_BYTES - 1),
                            g * H + q * a.box_rows, &s_bar);
        }
        if (tid < 12) s_T[tid] = a.T[g * 16 + tid];
        if (tid == 32) {
            // voxel of the sonar origin (the combiner's key origin); whether every sample of this frame
            // is certain to stay below 2^30 voxels from zero (then the fp64 quantiser's reciprocal
            // shortcut is exact); and whether the fp32 estimate's premises hold: finite transform,
            // every sample within the combiner's +-2^9 voxels of the origin
            const double *T = a.T + g * 16;
            const double reach = H > 0 ? tab.range_m[H - 1] : 0.0;
            bool fast = true;
            int o[3] = {0, 0, 0};
            double l2max = 0.0, l1max = 0.0;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const double l1 = fabs(T[4 * q]) + fabs(T[4 * q + 1]) + fabs(T[4 * q + 2]);
                const double bound = l1 * reach + fabs(T[4 * q + 3]);
                if (!(bound * a.p.inv_res < 1073741824.0)) fast = false;
                if (!voxel_index(T[4 * q + 3], a.p.res, a.p.inv_res, o[q])) fast = false;
                l1max = fmax(l1max, l1);
                l2max = fmax(l2max, sqrt(T[4 * q] * T[4 * q] + T[4 * q + 1] * T[4 * q + 1] + T[4 * q + 2] * T[4 * q + 2]));
            }
            s_o[0] = o[0]; s_o[1] = o[1]; s_o[2] = o[2];
            s_fast = fast;
            const bool fast32 = fast && (reach * a.p.inv_res * l2max + 2.0 < (double)(LK_HALF - 1)) && a.fast32_ok;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                s_c[q] = (float)T[4 * q + 2];
                s_f0[q] = (float)(T[4 * q + 3] * a.p.inv_res - (double)o[q]);
            }
            // error band of the estimate, per unit of rho (derivation in DESIGN.md): every rounding is
            // <= 2^-24 relative, five of them act on the rotated term and one on the result
            s_kband = (float)(l1max * 7.5e-7);
            const double zq = a.p.zmin * a.p.inv_res - (double)o[2];
            s_zq = (float)zq;
            s_zslack = (float)(fabs(zq) * 1.3e-7);
            s_fast32 = (fast32 && (!a.p.zfilter || fabs(zq) < 1e9)) ? 1 : 0;
        }
        if (a.use_tma) { mbar_wait(&s_bar, tma_parity); tma_parity ^= 1u; }
        __syncthreads();

        // ---- list: per warp, first hit and fan list of its beams
        for (int b = warp; b < nb; b += EX_WARPS) {
            Fan *fans = fans_all + (size_t)b * (nf_max + 1);
            const int beam = beam0 + b;
            const int col = tab.beam_col[beam];
            const double cb = tab.cos_b[beam], sb = tab.sin_b[beam];
            const unsigned char *strip = s_raw + (((beam0 + b) * tab.col_step) & (TMA_STRIP_BYTES - 1));   // this beam's column of the staged strip
            auto pix = [&](int r) -> int {
                return a.use_tma ? (int)strip[r * TMA_STRIP_BYTES] : (int)__ldg(&img[(size_t)r * W + col]);
            };
            // first above-threshold range bin of this beam (:406-409): 128 rows per step, lanes = rows
            int fh = H;                                                    // no hit -> whole ray is free (:412-413)
            for (int r0 = 0; r0 < H && fh == H; r0 += 128) {
                bool hit[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = r0 + q * 32 + lane;
                    hit[q] = r < H && pix(r) > a.p.thr;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const u32 mk = __ballot_sync(0xffffffffu, hit[q]);
                    if (mk && fh == H) fh = r0 + q * 32 + __ffs(mk) - 1;
                }
            }
            const int nfc = (fh + tab.free_step - 1) / tab.free_step;      // range(0, fh, free_step) (:420)
            const int noc = fh < H ? min(tab.occ_window, H - fh) : 0;      // range(fh, min(fh+50, H)) (:451)
            // fans in walking order, empty ones dropped: free = every free_step-th bin before the first
            // hit (:420; bins below min_range have nv == 0), then occupied = above-threshold bins in
            // the window from the first hit (:451-452)
            int run = 0, nfan = 0;
            const int nfc32 = (nfc + 31) & ~31;
            for (int base = 0; base < nfc32 + noc; base += 32) {
                int r = 0, nv = 0;
                u32 occ_bit = 0u;
                if (base < nfc32) {                                         // warp-uniform
                    const int c = base + lane;
                    if (c < nfc) { r = c * tab.free_step; nv = tab.nv_free[r]; }
                } else {
                    const int c = base - nfc32 + lane;
                    if (c < noc) {
                        r = fh + c;
                        nv = (pix(r) > a.p.thr) ? tab.nv_occ[r] : 0;       // :452, :456
                    }
                    occ_bit = 0x80000000u;
                }
                const u32 m = __ballot_sync(0xffffffffu, nv > 0);
                const int v = nv > 0 ? 2 * nv + 1 : 0;
                int incl = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if (nv > 0) {
                    Fan f; f.off = run + incl - v; f.code = (u32)r | ((u32)nv << 16) | occ_bit;
                    f.rho = (float)(tab.range_m[r] * a.p.inv_res); f.pad_ = 0u;
                    fans[nfan + __popc(m & lt_mask)] = f;
                }
                nfan += __popc(m);
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) {
                fans[nfan].off = run;
                s_tot[b] = run; s_nfan[b] = nfan; s_cb[b] = cb; s_sb[b] = sb;
            }
            if (lane < 3) s_ab[b][lane] = (float)(s_T[4 * lane] * cb - s_T[4 * lane + 1] * sb);
        }
        __syncthreads();
        if (a.use_tma) {                 // the strip sat on the combiner: make it a combiner again
            for (int i = tid; i < (int)(strip_smem_bytes(H) / 16); i += EX_THREADS) {
                if (i < LT_CAP / 4) reinterpret_cast<uint4 *>(tkey)[i] = make_uint4(LT_EMPTY, LT_EMPTY, LT_EMPTY, LT_EMPTY);
                else reinterpret_cast<uint4 *>(tkey)[i] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        if (tid == 0) {                  // first pass of each beam (exclusive prefix of the pass counts)
            int run = 0;
            for (int b = 0; b < nb; ++b) { s_pfirst[b] = run; run += (s_tot[b] + EX_PASS - 1) / EX_PASS; }
            for (int b = nb; b <= EX_MAXB; ++b) s_pfirst[b] = run;
        }
        __syncthreads();
        const bool fast = s_fast != 0, fast32 = s_fast32 != 0;
        const int o[3] = {s_o[0], s_o[1], s_o[2]};
        const float cz[3] = {s_c[0], s_c[1], s_c[2]}, f0[3] = {s_f0[0], s_f0[1], s_f0[2]};
        const float kband = s_kband, zq = s_zq, zslack = s_zslack;

        // ---- walk: the flattened (fan, vertical step) space of each beam is cut into passes of EX_PASS
        // consecutive samples; the passes of the tile's beams, in beam order, are dealt to the warps
        // round-robin, so the warps finish together whatever the beams hold.  The block walks in rounds
        // of EX_ROUND passes per warp; between rounds it votes on flushing the combiner, until what is
        // left certainly fits (then the warps run free to the end).
        u32 emitted = 0;
        const int n_pass = s_pfirst[EX_MAXB];        // passes of the whole tile
        const int my_pfirst = lane < nb ? s_pfirst[lane] : INT_MAX;
        int since = 0;                   // samples the block has walked since the last flush (>= combiner entries)
        for (int p_lo = 0; p_lo < n_pass; p_lo += EX_WARPS * EX_ROUND) {
            int rem = (n_pass - p_lo) * EX_PASS;                            // samples still to walk (upper bound)
            const bool free_run = since + rem <= LT_LIMIT;                  // block-uniform
            const int p_hi = free_run ? n_pass : min(n_pass, p_lo + EX_WARPS * EX_ROUND);
            for (int p = p_lo + warp; p < p_hi; p += EX_WARPS) {
                // which beam, and where in it
                const int b = __popc(__ballot_sync(0xffffffffu, my_pfirst <= p)) - 1;     // last beam that starts at or before p
                const int base = (p - s_pfirst[b]) * EX_PASS;               // first sample of the pass, in its beam
                const int total = s_tot[b], nfan = s_nfan[b];
                const float ab[3] = {s_ab[b][0], s_ab[b][1], s_ab[b][2]};
                const Fan *fans = fans_all + (size_t)b * (nf_max + 1);
                // fan that holds sample `base`: two-level search over the beam's fan offsets (32 probes each)
                int c0;
                {
                    const int S = (nfan + 31) / 32;
                    const int i1 = min(lane * S, nfan);
                    const int seg = __popc(__ballot_sync(0xffffffffu, fans[i1].off <= base && lane * S < nfan)) - 1;
                    const int i2 = seg * S + lane;
                    c0 = seg * S + __popc(__ballot_sync(0xffffffffu, lane < S && i2 < nfan && fans[min(i2, nfan)].off <= base)) - 1;
                }
                // fans that start inside this pass: bit (start - base) of a 64-bit mask (a fan has >= 3 samples)
                const int kf = c0 + 1 + lane;
                const int rel = (kf <= nfan ? fans[kf].off : INT_MAX) - base;
                // (word j of the mask covers samples base + 32 j .. base + 32 j + 31; a lane looks at one later fan,
                // and a fan has >= 3 samples, so 32 lanes see every start inside 96 samples -- the pass is 128 at most:
                // lanes also take fan c0 + 33 + lane)
                u32 fm[EX_ILP];
                {
                    const int kf2 = kf + 32;
                    const int rel2 = EX_ILP > 2 ? (kf2 <= nfan ? fans[kf2].off : INT_MAX) - base : INT_MAX;
#pragma unroll
                    for (int j = 0; j < EX_ILP; ++j) {
                        const bool in1 = rel > 0 && rel >= 32 * j && rel < 32 * j + 32;
                        const bool in2 = rel2 >= 32 * j && rel2 < 32 * j + 32;
                        fm[j] = __reduce_or_sync(0xffffffffu, (in1 ? 1u << (rel - 32 * j) : 0u) | (in2 ? 1u << (rel2 - 32 * j) : 0u));
                    }
                }
                const u32 upto = 0xffffffffu >> (31 - lane);                // bits 0..lane
                bool made[EX_ILP]; u32 made_at[EX_ILP];
#pragma unroll
                for (int j = 0; j < EX_ILP; ++j) {
                    made[j] = false; made_at[j] = 0;
                    const int w = base + j * 32 + lane;
                    if (w >= total) continue;
                    int fpos = c0 + __popc(fm[j] & upto);
#pragma unroll
                    for (int q = 0; q < j; ++q) fpos += __popc(fm[q]);
                    if (fpos > nfan) { atomicOr(&a.mc->err, ERR_INTERNAL); continue; }
                    const Fan f = fans[fpos];
                    const int nv = (int)((f.code >> 16) & 0x7fffu);
                    const bool occ = (f.code >> 31) != 0u;
                    const int ti = nv * nv - 1 + (w - f.off);               // row nv, entry v_step + nv
                    if ((u32)ti >= (u32)tab.n_trig) { atomicOr(&a.mc->err, ERR_INTERNAL); continue; }
                    bool need_exact = true, to_comb = false;
                    u32 lk = 0;
                    if (fast32) {
                        // fp32 estimate of the voxel index relative to the sonar-origin voxel
                        const float2 cs = __ldg(&tab.csva32[ti]);
                        const float band = fmaf(f.rho, kband, 3e-7f);
                        bool ok = true;
                        u32 tb[3];
                        float q2 = 0.f;
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const float v = fmaf(cz[q], cs.y, __fmul_rn(ab[q], cs.x));
                            const float qq = fmaf(f.rho, v, f0[q]);
                            const float t = __fadd_rd(qq, MAGICF);                       // floor(qq) + MAGICF, exactly
                            const float fr = __fadd_rn(qq, -__fadd_rn(t, -MAGICF));      // qq - floor(qq), exact
                            ok = ok && (fabsf(fr - 0.5f) < 0.5f - band);
                            tb[q] = __float_as_uint(t);
                            if (q == 2) q2 = qq;
                        }
                        bool drop = false;
                        if (a.p.zfilter) {                                               // :443, :478
                            const float dz = q2 - zq;
                            if (fabsf(dz) < band + zslack) ok = false; else drop = dz < 0.f;
                        }
                        constexpr u32 KOFF = (u32)LK_HALF - 0x4B400000u;
                        lk = ((tb[0] + KOFF) & (2 * LK_HALF - 1)) | (((tb[1] + KOFF) & (2 * LK_HALF - 1)) << LK_BITS) |
                             (((tb[2] + KOFF) & (2 * LK_HALF - 1)) << (2 * LK_BITS));
                        need_exact = !ok;
                        to_comb = ok && !drop;
                    }
                    if (need_exact || VERIFY) {
                        // the reference's arithmetic: sonar frame, X fwd / Y right / Z down, products rounded left to right (:434-436)
                        const double range = tab.range_m[f.code & 0xffffu];
                        const double cv = __ldg(&tab.cos_va[ti]), sv = __ldg(&tab.sin_va[ti]);
                        const double cb = s_cb[b], sb = s_sb[b];
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
                        bool x_comb = false; u32 x_lk = 0; bool x_direct = false; u64 x_key = 0;
                        if (!(a.p.zfilter && wv[2] < a.p.zmin)) {           // :443, :478
                            int ki, kj, kk;
                            if (!(quantise(a.p, wv[0], fast, ki) && quantise(a.p, wv[1], fast, kj) && quantise(a.p, wv[2], fast, kk))) {
                                atomicOr(&a.mc->err, ERR_KEYRANGE);
                            } else {
                                const u32 d0 = (u32)(ki - o[0] + LK_HALF), d1 = (u32)(kj - o[1] + LK_HALF), d2 = (u32)(kk - o[2] + LK_HALF);
                                if (fast && (d0 | d1 | d2) < (u32)(2 * LK_HALF)) {
                                    x_comb = true; x_lk = d0 | (d1 << LK_BITS) | (d2 << (2 * LK_BITS));
                                } else if (!key_in_range(ki, kj, kk)) atomicOr(&a.mc->err, ERR_KEYRANGE);
                                else { x_direct = true; x_key = pack_key(ki, kj, kk); }
                            }
                        }
                        if (VERIFY && !need_exact && (to_comb != x_comb || (to_comb && lk != x_lk))) atomicOr(&a.mc->err, ERR_VERIFY);
                        if (need_exact) {
                            to_comb = x_comb; lk = x_lk;
                            if (x_direct && !(a.own_world > 1 && key_owner(x_key, a.own_world) != a.own_rank)) {
                                ++emitted;
                                commit_direct<CT, CHECK, ROUTE>(a.skeys, static_cast<CT *>(a.scnt), a.smask, a.cc, a.mc, a.seq, x_key, g, occ, a.rt);
                            }
                        }
                    }
                    if (to_comb) {
                        // block combiner: find-or-insert the 30-bit local key, bump its 16+16-bit counts
                        u32 h = (lk * 0x9E3779B1u) >> (32 - LT_BITS);
                        for (;;) {
                            u32 cur = v_tkey[h];
                            if (cur == lk) break;
                            if (cur == LT_EMPTY) {
                                cur = atomicCAS(&tkey[h], LT_EMPTY, lk);
                                if (cur == LT_EMPTY) { made[j] = true; made_at[j] = h; break; }
                                if (cur == lk) break;
                            }
                            h = (h + 1) & (LT_CAP - 1);
                        }
                        atomicAdd(&tcnt[h], occ ? 0x10000u : 1u);
                    }
                }
                // combiner slots created by this pass join the live list: one shared-memory atomic per pass
                u32 mk[EX_ILP], n_made = 0;
#pragma unroll
                for (int j = 0; j < EX_ILP; ++j) { mk[j] = __ballot_sync(0xffffffffu, made[j]); n_made += __popc(mk[j]); }
                if (n_made) {
                    u32 at = 0;
                    if (lane == 0) at = atomicAdd(&s_count, n_made);
                    at = __shfl_sync(0xffffffffu, at, 0);
#pragma unroll
                    for (int j = 0; j < EX_ILP; ++j) {
                        if (made[j]) live[at + __popc(mk[j] & lt_mask)] = (unsigned short)made_at[j];
                        at += __popc(mk[j]);
                    }
                }
            }
            if (free_run) break;
            since += (p_hi - p_lo) * EX_PASS;
            rem -= (p_hi - p_lo) * EX_PASS;
            // flush if the next round could overflow the combiner's entries or its 16-bit counts.  Every
            // thread votes with the entry count it sees on arrival; the last one to arrive sees the final one.
            const int next = min(rem, EX_ROUND_SAMPLES);
            const int want = (*v_count + (u32)next > (u32)LT_LIMIT) || (since + next > LT_MAX_SAMPLES);
            if (__syncthreads_or(want)) {
                flush_combiner<CT, CHECK, ROUTE>(a, tkey, tcnt, live, v_count, o, g, emitted);
                since = 0;
            }
        }
        __syncthreads();
        if (*v_count > 0u) flush_combiner<CT, CHECK, ROUTE>(a, tkey, tcnt, live, v_count, o, g, emitted);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) emitted += __shfl_xor_sync(0xffffffffu, emitted, d);
        if (lane == 0 && emitted) atomicAdd(&a.stats[g].n_samples, (u64)emitted);
        // (every thread orders its generic-proxy accesses to the combiner before the async proxy's next strip)
        if (a.use_tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();                 // the tile's shared state is reused by the next one
    }
    expand_finish<ROUTE>(a);
}

// ------------------------------------------------------------------------------------ K4
// update_voxel (3d_mapper.py:83-115) on one table slot.
__device__ __forceinline__ double apply_one(double L, double upd, bool adaptive, const DevParams &p)
{
    if (adaptive && p.adaptive && upd > 0.0) {                  // :95
        // :97 prob = 1/(1+exp(-L)).  Two cases need no exp: L == 0 gives exactly 0.5, and L well
        // above logit(threshold) gives prob > threshold, i.e. no scaling at all.
        if (L == 0.0) {
            upd *= p.scale0;                                    // prob == 0.5 exactly: the factor is a constant (1 if 0.5 > threshold)
        } else if (!(L > p.l_skip)) {
            const double prob = 1.0 / (1.0 + exp(-L));
            if (prob <= p.a_thr) upd *= (prob / p.a_thr) * p.a_ratio;   // :100-102
        }
    }
    L += upd;                                                   // :107
    L = L < p.lo_min ? p.lo_min : (L > p.lo_max ? p.lo_max : L);   // :110 (L is never NaN)
    return L;
}

// find-or-insert; returns slot index or ~0 on failure.  `fresh` says the key was inserted.
// Probing goes by 32-byte sector = 2 slots (one L2 round trip sees both), starting at the even
// slot of the home pair; every reader of the table (k_query) walks the same sequence.
__device__ __forceinline__ u64 table_home(u64 key, u64 mask) { return (mix64(key) >> 8) & mask & ~1ull; }

__device__ __forceinline__ u64 table_find_or_insert(Slot *table, u64 mask, u64 key, bool &fresh, double &val)
{
    u64 pair = table_home(key, mask);
    fresh = false;
    for (u32 probe = 0; probe < (1u << 20); ++probe) {
        const ulonglong2 s0 = __ldcg(reinterpret_cast<const ulonglong2 *>(&table[pair]));
        const ulonglong2 s1 = __ldcg(reinterpret_cast<const ulonglong2 *>(&table[pair + 1]));
        if (s0.x == key) { val = __longlong_as_double((long long)s0.y); return pair; }
        if (s1.x == key) { val = __longlong_as_double((long long)s1.y); return pair + 1; }
        if (s0.x == EMPTY_KEY || s1.x == EMPTY_KEY) {
            const u64 slot = s0.x == EMPTY_KEY ? pair : pair + 1;
            const u64 cur = atomicCAS(&table[slot].key, EMPTY_KEY, key);
            if (cur == EMPTY_KEY) { fresh = true; val = 0.0; return slot; }     // :105-106
            if (cur == key) { val = __ldcg(&table[slot].val); return slot; }
            continue;                                   // lost the hole to another key: same pair again
        }
        pair = (pair + 2) & mask;
    }
    return ~0ull;
}

struct LocalAcc { int n_new; int kmin[3], kmax[3]; };

__device__ __forceinline__ void acc_init(LocalAcc &a)
{
    a.n_new = 0;
    for (int q = 0; q < 3; ++q) { a.kmin[q] = INT_MAX; a.kmax[q] = INT_MIN; }
}

__device__ __forceinline__ void acc_key(LocalAcc &a, u64 key)
{
    int ki, kj, kk; unpack_key(key, ki, kj, kk);
    a.kmin[0] = min(a.kmin[0], ki); a.kmax[0] = max(a.kmax[0], ki);
    a.kmin[1] = min(a.kmin[1], kj); a.kmax[1] = max(a.kmax[1], kj);
    a.kmin[2] = min(a.kmin[2], kk); a.kmax[2] = max(a.kmax[2], kk);
}

// warp-reduce; lane 0 publishes (bounds only when they extend the box).  add_count: also bump
// the live-voxel counter (the chunk path does that once, from its per-frame totals).
__device__ __forceinline__ void acc_publish(LocalAcc &a, MapCtr *mc, bool add_count)
{
    a.n_new = __reduce_add_sync(0xffffffffu, a.n_new);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        a.kmin[q] = __reduce_min_sync(0xffffffffu, a.kmin[q]);
        a.kmax[q] = __reduce_max_sync(0xffffffffu, a.kmax[q]);
    }
    if ((threadIdx.x & 31) == 0) {
        if (add_count && a.n_new) atomicAdd(&mc->count, (u64)a.n_new);
        int lo[3], hi[3];                                       // (read together: one round trip, not six)
#pragma unroll
        for (int q = 0; q < 3; ++q) { lo[q] = __ldcg(&mc->kmin[q]); hi[q] = __ldcg(&mc->kmax[q]); }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (a.kmin[q] < lo[q]) atomicMin(&mc->kmin[q], a.kmin[q]);
            if (a.kmax[q] > hi[q]) atomicMax(&mc->kmax[q], a.kmax[q]);
        }
    }
}

constexpr int AP_THREADS = 256;              // = entries per batch
constexpr int AP_WARPS = AP_THREADS / 32;
constexpr int AP_SCAN = 2;                   // dedupe slots per thread per scan step (independent loads)
constexpr int AP_TILE = AP_THREADS * AP_SCAN;   // dedupe slots a block scans per step
constexpr int AP_LCAP = 1024;                // live list (ring): less than a batch left over + one tile
constexpr int AP_PK_ROUNDS = 7;              // batches between reductions of the packed per-frame byte counters: 32 lanes x 7 < 256
constexpr int SUMT = 64;                     // entries of the sequential-sum tables
constexpr int AP_MF = 8;                     // mean table: n_free < AP_MF (any n_occ < SUMT), plus the pure-free row
constexpr int MEAN_WORDS = (AP_MF + 1) * SUMT;
static_assert(AP_LCAP >= AP_THREADS + AP_TILE && (AP_LCAP & (AP_LCAP - 1)) == 0, "list holds a leftover batch plus one tile");
// staged counter lanes of the entries of a batch: one padded row per entry (stride of 5 / 9
// sixteen-byte words: conflict-free 128-bit stores)
template <typename CT> __host__ __device__ constexpr int ap_row_words() { return (int)(sizeof(CT) * GF / 4) + 4; }
template <typename CT> __host__ __device__ constexpr size_t apply_smem_bytes() { return (size_t)AP_THREADS * ap_row_words<CT>() * 4; }

// sum of n_free copies of lo_free followed by n_occ copies of lo_occ, added one by one as the
// reference's `sum += log_odds` does (3d_mapper.py:546).  tab[0][n] / tab[1][n] hold the
// running sums of n copies of lo_free / lo_occ, tab[2][n] / tab[3][n] those sums divided by n;
// only counts >= SUMT or mixed voxels take the loop.
__device__ __forceinline__ double seq_avg(u32 n_free, u32 n_occ, const double (*tab)[SUMT], const DevParams &p)
{
    if (n_occ == 0 && n_free < SUMT) return tab[2][n_free];      // pure voxels: mean precomputed (sum/n)
    if (n_free == 0 && n_occ < SUMT) return tab[3][n_occ];
    double sum;
    if (n_free < SUMT) sum = tab[0][n_free];
    else { sum = tab[0][SUMT - 1]; for (u32 q = SUMT - 1; q < n_free; ++q) sum += p.lo_free; }
    if (n_occ) {
        if (n_free == 0 && n_occ < SUMT) sum = tab[1][n_occ];
        else for (u32 q = 0; q < n_occ; ++q) sum += p.lo_occ;
    }
    return sum / (double)(n_occ + n_free);                       // :559
}

// find-or-insert with the home pair already loaded (the loads were issued together with the
// entry's counter lanes, so a warp has all of its table sectors in flight at once)
__device__ __forceinline__ u64 table_resolve(Slot *table, u64 mask, u64 key, u64 pair, ulonglong2 s0, ulonglong2 s1,
                                             bool &fresh, double &val)
{
    fresh = false;
    for (u32 probe = 0; probe < (1u << 20); ++probe) {
        if (s0.x == key) { val = __longlong_as_double((long long)s0.y); return pair; }
        if (s1.x == key) { val = __longlong_as_double((long long)s1.y); return pair + 1; }
        if (s0.x == EMPTY_KEY || s1.x == EMPTY_KEY) {
            const u64 slot = s0.x == EMPTY_KEY ? pair : pair + 1;
            const u64 cur = atomicCAS(&table[slot].key, EMPTY_KEY, key);
            if (cur == EMPTY_KEY) { fresh = true; val = 0.0; return slot; }                       // :105-106
            if (cur == key) { val = __ldcg(&table[slot].val); return slot; }
        } else {
            pair = (pair + 2) & mask;
        }
        s0 = __ldcg(reinterpret_cast<const ulonglong2 *>(&table[pair]));
        s1 = __ldcg(reinterpret_cast<const ulonglong2 *>(&table[pair + 1]));
    }
    return ~0ull;
}

// spread the low 4 bits of x into the low bits of the 4 bytes of a word
__device__ __forceinline__ u32 spread4(u32 x) { return ((x & 0xFu) * 0x00204081u) & 0x01010101u; }

#ifdef S3D_AP_PHASES
// build-time instrumentation (tools only): clock64 spans of thread 0 of every block, summed per phase
__device__ unsigned long long g_ap_phase[16];
#define AP_PH(i) do { if (tid == 0) { const long long t1_ = clock64(); ph_acc[i] += (unsigned long long)(t1_ - ph_t0); ph_t0 = t1_; } } while (0)
#else
#define AP_PH(i) do { } while (0)
#endif

struct ApplyArgs {
    u64 *skeys; void *scnt; u32 n_slots; int g;
    ChunkCtr *cc; DevStats *st;
    Slot *table; u64 tmask;
    DevParams p; const double *sum_tab;
    MapCtr *mc; u64 table_limit; u64 seq; u64 *trace;
    u64 *marks;                 // S3D_MARKS: mapped host counters {blocks started, blocks finished}, or null
    // debug counters (DEBUG instantiation only)
    Slot *life;                 // lifetime sample counts: same geometry as the voxel table, val = u64 count
    ulonglong2 *last; u32 *last_n; u32 last_cap;   // {key, samples} of the chunk's last frame
};

// K4.  The chunk's dedupe table is walked by blocks in tiles of AP_TILE slots:
//   scan     every thread reads AP_SCAN slots' keys; live ones are appended to the block's list;
//   batch    whenever the list holds AP_THREADS entries (and at the end), every thread takes one:
//            the entry's counter lanes and the home sector of its voxel in the table are requested
//            together (six independent 16-byte loads per thread -- the whole block's table sectors
//            are in flight at once), the entry is wiped for the next chunk, the voxel is found or
//            inserted (one probe per voxel per chunk), lanes + L are staged in shared memory;
//   sort     the batch is ordered by the number of frames that touched each voxel (counting sort,
//            16 bins), so that the warps of the update step are homogeneous: the per-voxel frame
//            chains are sequential, and a warp takes as long as its longest chain;
//   update   one thread per staged entry walks only the frames that touched the voxel, in order:
//            per-voxel mean of the sample deltas (3d_mapper.py:557-559), then update_voxel
//            (:562-567); L goes back to the table with one 8-byte store.
// num_occupied / num_free per frame are kept as packed byte counters in registers (one add per
// entry per 4 frames) and reduced per warp when they could overflow and at the end.
// Frames stay strictly ordered per voxel, which is all the reference's sequential semantics
// require (voxels are independent of each other).
template <typename CT, bool DEBUG>
__global__ void __launch_bounds__(AP_THREADS, 3)
k_apply_chunk(const ApplyArgs a)
{
    trace_begin(a.trace);
    MapCtr *mc = a.mc; ChunkCtr *cc = a.cc;
    // a retry was asked for by this chunk or an earlier one: stay side-effect free.  (An older
    // chunk keeps running when a later chunk -- expanded concurrently -- raises the flag.)
    // (the four counter loads of this prologue are independent: one round trip instead of three)
    const u32 abort0 = __ldcg(&mc->abort);
    const u64 abort_seq0 = __ldcg(&mc->abort_seq);
    const u64 count0 = __ldcg(&mc->count);
    const u32 uniq0 = __ldcg(&cc->n_unique);
    if (abort0 != 0u && a.seq >= abort_seq0) return;
    // The gate: the chunk may be applied only if the table keeps its load bound even when every
    // voxel of the chunk is new; otherwise nothing of it touches the table and the host grows the
    // table and re-runs the chunk.  Every block takes the same decision from the same numbers
    // (the previous chunk has finished on this stream, so `count` is final until our last block).
    {
        u64 load = count0;
        if (DEBUG) load = max(load, __ldcg(&mc->life_count));
        if (load + uniq0 > a.table_limit) {
            if (blockIdx.x == 0 && threadIdx.x == 0) raise_abort(mc, ABORT_TABLE, a.seq);
            return;
        }
    }
    if (a.marks && threadIdx.x == 0) atomicAdd_system(a.marks + 2, 1ull);
    extern __shared__ __align__(16) unsigned char s_dyn[];          // staged counter lanes [AP_THREADS][ROWW]
    __shared__ u64 s_lkey[AP_LCAP];                                 // live list: key ...
    __shared__ u32 s_lslot[AP_LCAP];                                // ... and dedupe slot
    __shared__ double s_L[AP_THREADS];
    __shared__ u64 s_tslot[AP_THREADS];
    __shared__ u64 s_life[DEBUG ? AP_THREADS : 1], s_lifeslot[DEBUG ? AP_THREADS : 1];
    __shared__ u32 s_mask[AP_THREADS];
    __shared__ unsigned short s_ord[AP_THREADS];
    __shared__ u32 s_hist[GF + 1], s_tail;
    __shared__ u32 s_occ[GF], s_free[GF], s_new[GF];
    __shared__ u32 s_dmax[GF], s_dgt10[GF], s_lifenew;
    __shared__ u64 s_dlife[GF];
    __shared__ double s_mean[MEAN_WORDS];   // per-voxel mean of a frame's sample deltas by (n_free, n_occ): one lookup per update
    __shared__ bool s_last;
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 lt_mask = (1u << lane) - 1;
    const DevParams &p = a.p;
#ifdef S3D_AP_PHASES
    unsigned long long ph_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long ph_t0 = clock64();
#endif
    if (tid < GF) { s_occ[tid] = 0; s_free[tid] = 0; s_new[tid] = 0; s_dmax[tid] = 0; s_dgt10[tid] = 0; s_dlife[tid] = 0; }
    if (tid <= GF) s_hist[tid] = 0;
    if (tid == 0) { s_lifenew = 0; s_tail = 0; }
    for (int q = tid; q < MEAN_WORDS; q += AP_THREADS) s_mean[q] = a.sum_tab[4 * SUMT + q];
    const double (*gsum)[SUMT] = reinterpret_cast<const double (*)[SUMT]>(a.sum_tab);    // running sums, for counts beyond the table
    __syncthreads();
    constexpr int ROWW = ap_row_words<CT>();
    constexpr int NV = (int)(sizeof(CT) * GF / 16);
    u32 *rows = reinterpret_cast<u32 *>(s_dyn);
    CT *scnt = static_cast<CT *>(a.scnt);
    AP_PH(0);
    LocalAcc acc; acc_init(acc);
    u32 pk_occ[GF / 4], pk_free[GF / 4];                          // byte f%4 of word f/4: voxels updated as occupied / free in frame f
#pragma unroll
    for (int w = 0; w < GF / 4; ++w) { pk_occ[w] = 0; pk_free[w] = 0; }
    u32 pk_rounds = 0;
    u32 head = 0;                                                 // block-uniform: list entries already taken

    // (a byte of a lane counts at most one voxel per batch, so after <= AP_PK_ROUNDS batches the warp's sum of a
    // byte still fits the byte: the packed words are reduced whole, then lane f takes frame f's byte)
    auto flush_packed = [&]() {
        u32 xo = 0, xf = 0;
#pragma unroll
        for (int w = 0; w < GF / 4; ++w) {
            const u32 ro = __reduce_add_sync(0xffffffffu, pk_occ[w]);
            const u32 rf = __reduce_add_sync(0xffffffffu, pk_free[w]);
            if ((int)(lane >> 2) == w) { xo = ro; xf = rf; }
            pk_occ[w] = 0; pk_free[w] = 0;
        }
        if (lane < (u32)GF) {
            const u32 so = (xo >> (8 * (lane & 3))) & 0xffu, sf = (xf >> (8 * (lane & 3))) & 0xffu;
            if (so) atomicAdd(&s_occ[lane], so);
            if (sf) atomicAdd(&s_free[lane], sf);
        }
        pk_rounds = 0;
    };

    // one batch: threads < n take the entries at the head of the list.  Block-wide.
    auto batch = [&](u32 n) {
        const bool have = tid < n;
        u32 nfr = 0, rank = 0;                                     // frames that touched my entry; my rank among the entries with as many
        if (have) {
            const u32 li = (head + tid) & (AP_LCAP - 1);
            const u64 key = s_lkey[li];
            const u32 s = s_lslot[li];
            // ---- all loads of the entry first: counter lanes + home sector of the voxel (+ lifetime table)
            uint4 *cp = reinterpret_cast<uint4 *>(scnt + (size_t)s * GF);
            uint4 raw[NV];
#pragma unroll
            for (int q = 0; q < NV; ++q) raw[q] = __ldcg(cp + q);
            const u64 pair = table_home(key, a.tmask);
            const ulonglong2 t0 = __ldcg(reinterpret_cast<const ulonglong2 *>(&a.table[pair]));
            const ulonglong2 t1 = __ldcg(reinterpret_cast<const ulonglong2 *>(&a.table[pair + 1]));
            ulonglong2 l0 = make_ulonglong2(0ull, 0ull), l1 = l0;
            if (DEBUG) {
                l0 = __ldcg(reinterpret_cast<const ulonglong2 *>(&a.life[pair]));
                l1 = __ldcg(reinterpret_cast<const ulonglong2 *>(&a.life[pair + 1]));
            }
            // ---- the entry is ready for the next chunk
            AP_PH(2);
            a.skeys[s] = EMPTY_KEY;
#pragma unroll
            for (int q = 0; q < NV; ++q) cp[q] = make_uint4(0u, 0u, 0u, 0u);
            // ---- frames that saw the voxel at all / occupied (occupied has priority, :544-545); stage the lanes
            u32 m_occ = 0, m_any = 0;
            u32 *my_row = rows + (size_t)tid * ROWW;
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                reinterpret_cast<uint4 *>(my_row)[q] = raw[q];
                const u32 w[4] = {raw[q].x, raw[q].y, raw[q].z, raw[q].w};
                if (sizeof(CT) == 4) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        m_any |= (w[j] != 0u ? 1u : 0u) << (4 * q + j);
                        m_occ |= ((w[j] >> 16) != 0u ? 1u : 0u) << (4 * q + j);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        m_any |= ((w[2 * j] | w[2 * j + 1]) != 0u ? 1u : 0u) << (2 * q + j);
                        m_occ |= (w[2 * j + 1] != 0u ? 1u : 0u) << (2 * q + j);
                    }
                }
            }
            AP_PH(3);
            u64 slot = ~0ull; double L = 0.0;
            if (m_any) {
                bool fresh;
                slot = table_resolve(a.table, a.tmask, key, pair, t0, t1, fresh, L);
                if (slot == ~0ull) { atomicOr(&mc->err, ERR_TABLEFULL); m_any = 0; }
                else {
                    if (fresh) atomicAdd(&s_new[__ffs(m_any) - 1], 1u);       // len(voxels) grows at the first frame that touched it
                    if (DEBUG) {
                        bool lfresh; double lv;
                        const u64 lslot = table_resolve(a.life, a.tmask, key, pair, l0, l1, lfresh, lv);
                        if (lslot == ~0ull) atomicOr(&mc->err, ERR_TABLEFULL);
                        else if (lfresh) atomicAdd(&s_lifenew, 1u);
                        s_lifeslot[tid] = lslot;
                        s_life[tid] = (lslot == ~0ull || lfresh) ? 0ull : (u64)__double_as_longlong(lv);
                        const CT last = reinterpret_cast<const CT *>(my_row)[a.g - 1];   // frame_update_counts of the chunk's last frame
                        if (last != 0) {
                            const u32 at = atomicAdd(a.last_n, 1u);
                            if (at < a.last_cap) a.last[at] = make_ulonglong2(key, (u64)(Lane<CT>::n_occ(last) + Lane<CT>::n_free(last)));
                        }
                    }
                    acc_key(acc, key);
                    // num_occupied / num_free of every frame that touched the voxel (:562-567)
                    const u32 m_free = m_any & ~m_occ;
#pragma unroll
                    for (int w = 0; w < GF / 4; ++w) { pk_occ[w] += spread4(m_occ >> (4 * w)); pk_free[w] += spread4(m_free >> (4 * w)); }
                }
            }
            s_L[tid] = L; s_tslot[tid] = slot; s_mask[tid] = m_any;
            nfr = (u32)__popc(m_any);
            AP_PH(4);
        }
        {
            // rank among the batch's entries with the same frame count (one shared atomic per distinct count per warp)
            const u32 peers = __match_any_sync(0xffffffffu, have ? nfr : 0xffffu);
            u32 base = 0;
            if (have && lane == (u32)__ffs(peers) - 1) base = atomicAdd(&s_hist[nfr], (u32)__popc(peers));
            base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
            rank = base + __popc(peers & lt_mask);
        }
        if (++pk_rounds == (u32)AP_PK_ROUNDS) flush_packed();
        __syncthreads();
        AP_PH(5);
        {
            // longest chains first: an entry goes behind all entries with more frames.  Lane l holds the number of
            // entries with l + 1 frames; a warp suffix scan gives every count's offset at once.
            u32 suf = lane < (u32)GF ? s_hist[lane + 1] : 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 t = __shfl_down_sync(0xffffffffu, suf, d);
                if (lane + d < 32u) suf += t;
            }
            const u32 longer = __shfl_sync(0xffffffffu, suf, nfr & 31u);     // entries with more than nfr frames
            if (have) s_ord[rank + (nfr < 32u ? longer : 0u)] = (unsigned short)tid;
        }
        __syncthreads();
        AP_PH(6);
        if (tid <= GF) s_hist[tid] = 0;
        if (have) {
            const u32 e = s_ord[tid];
            u32 todo = s_mask[e];
            if (todo) {
                double L = s_L[e];
                u64 life = DEBUG ? s_life[e] : 0ull;
                const CT *row = reinterpret_cast<const CT *>(rows + (size_t)e * ROWW);
                // (measured and rejected: unrolling the chain over all GF frames with the lanes and means fetched up
                // front -- the chains are short on average, and the extra work cost more than the shorter steps saved)
                while (todo) {                                                  // only the frames that touched the voxel, in order
                    const int f = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const CT cf = row[f];
                    const u32 n_occ = Lane<CT>::n_occ(cf), n_free = Lane<CT>::n_free(cf);
                    // mean of the frame's sample deltas for this voxel (:557-559): tabulated by (n_free, n_occ), so that
                    // free, occupied and mixed voxels of a warp take the same few instructions
                    const bool small = n_free < (u32)AP_MF && n_occ < (u32)SUMT;
                    const u32 ti = small ? n_free * (u32)SUMT + n_occ : (u32)(AP_MF * SUMT) + n_free;
                    double upd;
                    if (small || (n_occ == 0u && n_free < (u32)SUMT)) upd = s_mean[ti];
                    else upd = seq_avg(n_free, n_occ, gsum, p);
                    L = apply_one(L, upd, n_occ > 0u, p);                         // :562-567
                    if (DEBUG) {
                        const u32 nn = n_occ + n_free;                          // frame_update_counts[key] (:550)
                        life += nn;                                             // voxel_update_counts[key] (:551)
                        atomicMax(&s_dmax[f], nn);
                        if (nn > 10u) atomicAdd(&s_dgt10[f], 1u);
                        atomicMax(&s_dlife[f], life);
                    }
                }
                a.table[s_tslot[e]].val = L;
                if (DEBUG && s_lifeslot[e] != ~0ull) a.life[s_lifeslot[e]].val = __longlong_as_double((long long)life);
            }
        }
        head += n;
        AP_PH(7);
        __syncthreads();
        AP_PH(8);
    };

    const u32 n_tiles = a.n_slots / AP_TILE;
    u32 tile = blockIdx.x;
    bool more = tile < n_tiles;
    while (more) {
        const u32 s0 = tile * AP_TILE;
        u64 k[AP_SCAN];
#pragma unroll
        for (int j = 0; j < AP_SCAN; ++j) k[j] = __ldcg(&a.skeys[s0 + j * AP_THREADS + tid]);
#pragma unroll
        for (int j = 0; j < AP_SCAN; ++j) {
            const bool live = k[j] != EMPTY_KEY;
            const u32 m = __ballot_sync(0xffffffffu, live);
            u32 at = 0;
            if (lane == 0 && m) at = atomicAdd(&s_tail, (u32)__popc(m));
            at = __shfl_sync(0xffffffffu, at, 0);
            if (live) {
                const u32 li = (at + __popc(m & lt_mask)) & (AP_LCAP - 1);
                s_lkey[li] = k[j]; s_lslot[li] = s0 + j * AP_THREADS + tid;
            }
        }
        tile += gridDim.x;
        more = tile < n_tiles;
        __syncthreads();
        AP_PH(1);
        u32 avail = s_tail - head;
        bool ran = false;
        while (avail >= (u32)AP_THREADS || (!more && avail > 0u)) {   // full batches; the last one may be ragged
            const u32 n = min(avail, (u32)AP_THREADS);
            batch(n);
            avail -= n;
            ran = true;
        }
        // Nobody may append to the list before every thread has read the tail above: a warp that ran ahead into the
        // next scan would move `s_tail` under a slower warp, the two would disagree on `avail`, and the block would
        // split over the barriers inside batch().  A batch ends with a barrier; a tile too sparse for one needs its own.
        // (found with tools/stress_growth.py: sparse tiles are what the short last chunk of a call produces)
        if (!ran) __syncthreads();
    }
    flush_packed();
    acc_publish(acc, mc, false);
    __syncthreads();
    if (tid < (u32)a.g) {
        const int f = tid;
        if (s_occ[f]) atomicAdd(&a.st[f].n_occ, (u64)s_occ[f]);
        if (s_free[f]) atomicAdd(&a.st[f].n_free, (u64)s_free[f]);
        if (s_new[f]) atomicAdd(&cc->neu[f], s_new[f]);
        if (DEBUG) {
            if (s_dmax[f]) atomicMax(&cc->dmax[f], s_dmax[f]);
            if (s_dgt10[f]) atomicAdd(&cc->dgt10[f], s_dgt10[f]);
            if (s_dlife[f]) atomicMax(&cc->dlife[f], s_dlife[f]);
        }
    }
    if (DEBUG && tid == 0 && s_lifenew) atomicAdd(&cc->life_new, s_lifenew);
    // last block out: len(voxels) after each frame (:592), publish the new count, re-arm the chunk
    __syncthreads();
    trace_end(a.trace);
    if (a.marks && tid == 0) atomicAdd_system(a.marks + 3, 1ull);
    if (tid == 0) {
        __threadfence();
        const u32 t = atomicAdd(&cc->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && tid < 32u) {
        // (one warp, lane f = frame f: the per-frame totals are fetched together -- a single L2 round trip, where a
        // loop over the frames would pay one per frame with every other block already gone -- and turned into running
        // values by a warp scan)
        __threadfence();
        const bool mine = (int)lane < a.g;
        u32 neu = mine ? atomicAdd(&cc->neu[lane], 0u) : 0u;
        u64 dl = 0; u32 dm = 0, dg = 0;
        if (DEBUG && mine) { dl = atomicMax(&cc->dlife[lane], 0ull); dm = atomicMax(&cc->dmax[lane], 0u); dg = atomicAdd(&cc->dgt10[lane], 0u); }
        const u32 uniq = atomicAdd(&cc->n_unique, 0u);
        const u32 life_new = DEBUG ? atomicAdd(&cc->life_new, 0u) : 0u;
        u64 lmax = DEBUG ? max(__ldcg(&mc->life_max), dl) : 0ull;
        u32 incl = neu;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, d);
            const u64 tl = __shfl_up_sync(0xffffffffu, lmax, d);
            if ((int)lane >= d) { incl += t; lmax = max(lmax, tl); }
        }
        if (mine) {
            a.st[lane].n_voxels = count0 + incl;                            // len(voxels) after frame f (:592)
            cc->neu[lane] = 0;
            if (DEBUG) {
                a.st[lane].max_in_frame = dm;
                a.st[lane].n_gt10 = dg;
                a.st[lane].max_total = lmax;                                // max(voxel_update_counts.values()) (:578)
                cc->dmax[lane] = 0; cc->dgt10[lane] = 0; cc->dlife[lane] = 0;
            }
        }
        const u32 total = __shfl_sync(0xffffffffu, incl, 31);
        const u64 lmax_all = __shfl_sync(0xffffffffu, lmax, 31);
        if (lane == 0) {
            if (DEBUG) {
                mc->life_max = lmax_all;
                mc->life_count = __ldcg(&mc->life_count) + life_new;
                cc->life_new = 0;
            }
            mc->last_new = total;
            mc->last_unique = uniq;
            mc->probes += uniq;
            atomicExch(&mc->count, count0 + total);
            cc->n_unique = 0; cc->ticket = 0;
        }
    }
#ifdef S3D_AP_PHASES
    AP_PH(9);
    if (tid == 0) {
        for (int i = 0; i < 10; ++i) atomicAdd(&g_ap_phase[i], ph_acc[i]);
        atomicAdd(&g_ap_phase[15], 1ull);
    }
#endif
}

// ------------------------------------------------------------------------------ sharded map
// The map shards by a hash of the voxel key: owner = (mix64(key) >> 40) % world.  A rank expands
// its slice of the beams into its local dedupe table, drains the table into per-owner runs of
// 17-word records {key, counter lane of each of the 16 frames}, the runs travel by all-to-all,
// and the owner merges what it receives (integer adds, so the result does not depend on how the
// beams were split) before the ordinary apply kernel runs on its shard of the table.
constexpr int REC_WORDS = 1 + GF;
static_assert(REC_WORDS == S3D_RECORD_WORDS, "record layout of the NCCL route");

// pass 1: how many of the chunk's dedupe entries go to each owner
__global__ void k_shard_count(const u64 *__restrict__ skeys, u32 n_slots, u32 world, u32 *owner_count)
{
    __shared__ u32 s_cnt[64];
    if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x) {
        const u64 key = skeys[i];
        if (key != EMPTY_KEY) atomicAdd(&s_cnt[key_owner(key, world)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < world && s_cnt[threadIdx.x]) atomicAdd(&owner_count[threadIdx.x], s_cnt[threadIdx.x]);
}

// pass 2: drain the dedupe table into the send buffer, records grouped by owner (wire format:
// the packed key and one (n_occ << 32 | n_free) word per frame, whatever the lane format).
// owner_base[o] = first record of owner o (exclusive prefix of the counts), owner_fill[o] = cursor.
template <typename CT>
__global__ void k_shard_pack(u64 *__restrict__ skeys, CT *__restrict__ scnt, u32 n_slots, u32 world,
                             const u32 *__restrict__ owner_base, u32 *owner_fill, u64 *__restrict__ send)
{
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += gridDim.x * blockDim.x) {
        const u64 key = skeys[s];
        if (key == EMPTY_KEY) continue;
        const u32 o = key_owner(key, world);
        const u32 dst = owner_base[o] + atomicAdd(&owner_fill[o], 1u);
        u64 *rec = send + (size_t)dst * REC_WORDS;
        rec[0] = key;
        CT *cp = scnt + (size_t)s * GF;
#pragma unroll
        for (int f = 0; f < GF; ++f) {
            const CT c = cp[f];
            rec[1 + f] = ((u64)Lane<CT>::n_occ(c) << 32) | (u64)Lane<CT>::n_free(c);
            cp[f] = 0;
        }
        skeys[s] = EMPTY_KEY;
    }
}

__global__ void k_shard_reset_cc(ChunkCtr *cc) { cc->n_unique = 0; cc->ticket = 0; }

// owner side: merge received records into the (empty) dedupe table
template <typename CT>
__global__ void k_shard_merge(const u64 *__restrict__ recv, u64 n_rec, u64 *skeys, CT *scnt, u32 smask, ChunkCtr *cc,
                              MapCtr *mc, u64 seq)
{
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n_rec; i += (u64)gridDim.x * blockDim.x) {
        const u64 *rec = recv + i * REC_WORDS;
        const u64 key = rec[0];
        const u32 base = dedupe_home(key, smask);
        bool created;
        const u32 slot = dedupe_slot(skeys, smask, key, base, load_bucket(skeys, base), created);
        if (slot == ~0u) { raise_abort(mc, ABORT_SCRATCH, seq); continue; }
        if (created) atomicAdd(&cc->n_unique, 1u);
#pragma unroll
        for (int f = 0; f < GF; ++f) {
            const u64 w = rec[1 + f];
            if (!w) continue;
            const u32 n_occ = (u32)(w >> 32), n_free = (u32)(w & 0xffffffffu);
            if (Lane<CT>::narrow) {
                if ((n_occ | n_free) > 0xffffu) { raise_abort(mc, ABORT_NARROW, seq); continue; }
                const CT inc = Lane<CT>::make(n_occ, n_free);
                const CT old = atomicAdd(&scnt[(size_t)slot * GF + f], inc);
                if (Lane<CT>::overflows(old, inc)) raise_abort(mc, ABORT_NARROW, seq);
            } else {
                atomicAdd(&scnt[(size_t)slot * GF + f], Lane<CT>::make(n_occ, n_free));
            }
        }
    }
}

// ---- routed map: signalling and the owner-side merge
// One small block: thread s waits until word[s] >= want (skipping this rank).  Kept out of the
// big kernels on purpose.  Measured on B200: a merge kernel that spins in all of its blocks while
// it waits for the sources deadlocks the rank -- its blocks sit on every SM, an SM cannot switch to
// the large shared-memory carve-out k_expand needs while other blocks are resident, so this
// rank's own k_expand never starts, never signals, and the peers spin for ever too.  One spinning
// block ties up one SM at most.
__global__ void k_route_wait(const u64 *words, u32 world, u32 rank, u64 want, u64 timeout_ns, MapCtr *mc, u64 *trace)
{
    trace_begin(trace);
    const u32 s = threadIdx.x;
    if (s < world && s != rank) route_wait_word(words + s, want, timeout_ns, mc);
    __threadfence_system();
    __syncthreads();
    trace_end(trace);
}

// a rank whose beam slice is empty still has to tell its peers that nothing is coming
__global__ void k_route_signal(RouteCtx rt) { route_signal(rt, nullptr); }

// Owner side (its own stream: it runs beside this rank's own k_expand of the chunk -- both add into
// the same dedupe table with atomics -- and beside the apply of the chunk before): merge the
// records of all sources into this rank's dedupe table of the chunk; the last block out
// acknowledges to the sources (they may reuse the parity).
template <typename CT, bool CHECK>
__global__ void k_route_merge(RouteCtx rt, u64 *skeys, CT *scnt, u32 smask, ChunkCtr *cc, MapCtr *mc, u64 seq, u64 *trace)
{
    __shared__ bool s_last;
    __shared__ u64 s_first[ROUTE_MAX_WORLD + 1];       // exclusive prefix of the sources' record counts
    trace_begin(trace);
    const RouteHdr *h = reinterpret_cast<const RouteHdr *>(rt.peer[rt.rank]);
    // (every source has published the chunk: k_route_wait ran before this kernel on this stream)
    __threadfence_system();
    if (threadIdx.x == 0) {
        u64 run = 0;
        for (u32 src = 0; src < rt.world; ++src) {
            s_first[src] = run;
            if (src != rt.rank) run += *(const volatile u64 *)&h->flag_count[rt.parity][src];
        }
        s_first[rt.world] = run;
    }
    __syncthreads();
    u32 made = 0;
    if (!__ldcg(&mc->abort)) {
        const u64 total = s_first[rt.world];
        for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
            u32 src = 0;
            while (s_first[src + 1] <= i) ++src;                     // world <= 64: a short walk
            const RouteRec *in = route_inbox(rt, rt.rank, rt.parity, src) + (i - s_first[src]);
            const ulonglong2 raw = __ldcv(reinterpret_cast<const ulonglong2 *>(in));   // written by a peer: do not cache
            const u64 key = raw.x;
            const u32 counts = (u32)raw.y, frame = (u32)(raw.y >> 32);
            const u32 home = dedupe_home(key, smask);
            made += dedupe_add<CT, CHECK>(skeys, scnt, smask, mc, seq, key, home, load_bucket(skeys, home),
                                          (int)(frame & (GF - 1)), counts >> 16, counts & 0xffffu);
        }
    }
    made = __reduce_add_sync(0xffffffffu, made);
    if ((threadIdx.x & 31) == 0 && made) atomicAdd(&cc->n_unique, made);
    __syncthreads();
    trace_end(trace);
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(&cc->mticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < rt.world && threadIdx.x != rt.rank) {
        RouteHdr *ph = reinterpret_cast<RouteHdr *>(rt.peer[threadIdx.x]);
        __threadfence_system();
        *(volatile u64 *)&ph->ack_seq[rt.parity][rt.rank] = rt.seq + 1;
    }
    if (threadIdx.x == 0) cc->mticket = 0;
}


// ------------------------------------------------------------------------------ store kernels
__global__ void k_fill_slots(Slot *t, u64 n)
{
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        *reinterpret_cast<ulonglong2 *>(&t[i]) = make_ulonglong2(EMPTY_KEY, 0ull);
}

__global__ void k_fill_u64(u64 *t, u64 n, u64 v)
{
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) t[i] = v;
}

__global__ void k_clear_abort(MapCtr *mc, ChunkCtr *cc)
{
    mc->abort = 0; mc->abort_seq = ~0ull;
    for (int b = 0; b < N_CHUNK_BUF; ++b) {
        cc[b].n_unique = 0; cc[b].ticket = 0; cc[b].xticket = 0; cc[b].mticket = 0;
        cc[b].life_new = 0;
        for (int f = 0; f < GF; ++f) { cc[b].neu[f] = 0; cc[b].dmax[f] = 0; cc[b].dgt10[f] = 0; cc[b].dlife[f] = 0; }
    }
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
            // (bounds: the reference extends them with the caller's raw point, :113-115 -- the host keeps those)
            table[slot].val = apply_one(L, delta[i], adaptive[i] != 0, p);
            if (fresh) ++acc.n_new;
        }
    }
    acc_publish(acc, mc, true);
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
            table[slot].val = vals[i];                     // a direct dict write: no bounds (:113-115 are update_voxel only)
            if (fresh) ++acc.n_new;
        }
    }
    acc_publish(acc, mc, true);
}

__global__ void k_query(const u64 *__restrict__ keys, u64 n, const Slot *__restrict__ table, u64 tmask,
                        double *__restrict__ out, uint8_t *__restrict__ found)
{
    const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 key = keys[i];
    u64 slot = table_home(key, tmask);
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

// CUBE_LIST payloads (scripts/3d_mapper_node.py:448-527): the voxel centres of each class as one contiguous run of
// geometry_msgs/Point (3 x float64); base[c] = first point of class c (from a counting pass), cursor[c] = fill
__global__ void __launch_bounds__(256)
k_export_grouped(const Slot *__restrict__ table, u64 n_slots, double res, double thr_occ, double thr_free,
                 const u64 *__restrict__ base, u64 *cursor, double *__restrict__ xyz)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 n_iter = (n_slots + stride - 1) / stride;      // uniform trip count keeps ballots full-warp
    const u32 lane = threadIdx.x & 31;
    for (u64 it = 0; it < n_iter; ++it) {
        const u64 i = it * stride + blockIdx.x * (u64)blockDim.x + threadIdx.x;
        int cls = -1; u64 key = 0;
        if (i < n_slots) {
            const ulonglong2 raw = __ldcs(reinterpret_cast<const ulonglong2 *>(&table[i]));
            if (raw.x != EMPTY_KEY) {
                key = raw.x;
                const double L = __longlong_as_double((long long)raw.y);
                cls = L < thr_free ? S3D_CLASS_FREE : (L > thr_occ ? S3D_CLASS_OCCUPIED : S3D_CLASS_UNKNOWN);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const u32 b = __ballot_sync(0xffffffffu, cls == c);
            if (!b) continue;
            u64 at = 0;
            if (lane == (u32)__ffs(b) - 1) at = atomicAdd(&cursor[c], (u64)__popc(b));
            at = __shfl_sync(0xffffffffu, at, __ffs(b) - 1);
            if (cls == c) {
                const u64 dst = base[c] + at + __popc(b & ((1u << lane) - 1));
                int ki, kj, kk; unpack_key(key, ki, kj, kk);
                xyz[3 * dst] = __dmul_rn((double)ki + 0.5, res);         // key_to_world (:78-80)
                xyz[3 * dst + 1] = __dmul_rn((double)kj + 0.5, res);
                xyz[3 * dst + 2] = __dmul_rn((double)kk + 0.5, res);
            }
        }
    }
}

__global__ void k_pack_xyzi32(const double *__restrict__ xyz, const double *__restrict__ prob, u64 n,
                              float4 *__restrict__ out)
{
    const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4((float)xyz[3 * i], (float)xyz[3 * i + 1], (float)xyz[3 * i + 2], (float)prob[i]);
}

// 16-bit sonar frames (mono16 / 16UC1): the node's `(img / 256).astype(uint8)`
// (scripts/3d_mapper_node.py:308-310) folded into the upload -- the high byte of every pixel
__global__ void k_mono16_to_u8(const uint16_t *__restrict__ in, uint8_t *__restrict__ out, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            const ushort4 v = *reinterpret_cast<const ushort4 *>(in + i);      // i is a multiple of 4: 8-byte aligned
            *reinterpret_cast<uchar4 *>(out + i) = make_uchar4((unsigned char)(v.x >> 8), (unsigned char)(v.y >> 8),
                                                               (unsigned char)(v.z >> 8), (unsigned char)(v.w >> 8));
        } else {
            for (size_t q = i; q < n; ++q) out[q] = (uint8_t)(in[q] >> 8);
        }
    }
}

__global__ void k_reset_ctr(MapCtr *mc)
{
    // (life_count / life_max stay: reset_map does not clear voxel_update_counts, 3d_mapper.py:644-650)
    mc->count = 0; mc->err = 0; mc->abort = 0; mc->abort_seq = ~0ull; mc->last_new = 0; mc->last_unique = 0;
    for (int q = 0; q < 3; ++q) { mc->kmin[q] = INT_MAX; mc->kmax[q] = INT_MIN; }
}

__global__ void k_init_life_ctr(MapCtr *mc) { mc->life_count = 0; mc->life_max = 0; mc->route_sent = 0; mc->probes = 0; }

// compact the lifetime table into {key, count} pairs (debug view voxel_update_counts)
__global__ void k_dump_life(const Slot *__restrict__ life, u64 n_slots, ulonglong2 *out, u64 out_cap, u64 *cursor)
{
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n_slots; i += (u64)gridDim.x * blockDim.x) {
        const ulonglong2 raw = *reinterpret_cast<const ulonglong2 *>(&life[i]);
        if (raw.x == EMPTY_KEY) continue;
        const u64 at = atomicAdd(cursor, 1ull);
        if (at < out_cap) out[at] = raw;
    }
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

struct Job {                         // one s3d_ingest_batch* call: frames [0, n) of caller-owned device buffers
    u64 id = 0;
    const uint8_t *imgs = nullptr; const double *T = nullptr; DevStats *stats = nullptr;
    int64_t n = 0, next = 0;         // frames [0, next) have been enqueued
    u64 last_seq = ~0ull;            // sequence number of the chunk that holds its last frame (once enqueued)
    cudaEvent_t done_ev = nullptr;   // recorded behind the apply of that chunk (asynchronous batches)
};

// staging of one asynchronous host batch (s3d_ingest_submit / s3d_ingest_collect); two of them, so
// that the copy and the kernels of one batch overlap the tail of the batch before
struct Staging {
    uint8_t *img = nullptr; size_t img_n = 0; double *T = nullptr; size_t T_n = 0;
    DevStats *stats = nullptr, *stats_host = nullptr; size_t stats_n = 0;
    u64 last_job = 0; int64_t n = 0; bool busy = false;
};

struct InFlight { u64 seq = ~0ull; u64 job = 0; int64_t base = 0; int g = 0; };

struct s3d_map {
    int device = 0;
    cudaStream_t stream = nullptr;
    int n_sm = 148;
    // voxel table
    Slot *table = nullptr; u64 cap = 0;
    MapCtr *mc = nullptr;            // device
    MapCtr *mc_host = nullptr;       // pinned mirror
    u64 count_known = 0;             // live voxels at the last sync / chunk snapshot
    u64 unique_est = 0;              // voxels touched by a recent chunk (upper bound of what it can insert)
    static constexpr int RING = 40;  // per-chunk counter snapshots (bounded launch-ahead)
    MapCtr *snap_host = nullptr;     // pinned [RING]
    cudaEvent_t snap_ev[RING] = {};
    InFlight inflight[RING];
    u64 chunk_seq = 0;               // chunks enqueued so far
    u64 snap_floor = 0;              // snapshots older than this predate the last exact sync
    std::deque<Job> jobs; u64 job_seq = 0;
    u64 n_retries = 0, n_grows = 0;
    // sharded map (multi-GPU): this rank's identity and beam slice, exchange staging
    int shard_rank = 0, shard_world = 1, beam_lo = 0, beam_hi = -1;
    bool shard_filter = false;       // replicated expansion: s3d_ingest* keeps only the voxels this rank owns
    // routed map: exchange block (flags + inboxes) of this rank, the peers' blocks, local cursors
    bool route_on = false;
    bool route_by_chunk = true;      // routed map: ranks take turns expanding whole chunks (S3D_ROUTE_SPLIT=beams: beam slices)
    unsigned char *xblock = nullptr; size_t xblock_bytes = 0; u64 route_cap = 0;
    std::vector<unsigned char *> peer_ptr; std::vector<void *> ipc_opened;
    DevBuf<unsigned char *> d_peers; DevBuf<u32> route_cursor;
    u64 route_seq = 0;               // chunks routed so far (the same on every rank)
    u64 route_sent_seen = 0;         // MapCtr::route_sent at the last s3d_profile_read
    u64 probes_seen = 0;             // MapCtr::probes at the last s3d_profile_read
    u64 route_timeout_ns = 30000000000ull;   // S3D_ROUTE_TIMEOUT_MS
    int lookahead_env = 0;           // S3D_LOOKAHEAD (experiments)
    u64 scratch_env = 0;             // S3D_SCRATCH_CAP: first size of the chunk dedupe tables (tests force retries with a tiny one)
    int bpb_env = 0;                 // S3D_BEAMS_PER_BLOCK (experiments)
    // S3D_TRACE: per chunk 5 x {start, end}: ack wait, expand, flag wait, merge, apply
    DevBuf<u64> trace; static constexpr u64 TRACE_CHUNKS = 4096; static constexpr int TRACE_W = 10;
    DevBuf<u64> send_buf; DevBuf<u32> owner_ctr;      // owner_ctr: [3][64] count / base / fill
    u32 *owner_host = nullptr;                        // pinned [64]
    // params / tables
    bool have_params = false, have_tables = false;
    DevParams p{};
    DevTables tab{};
    DevBuf<int> d_beam_col, d_nv_free, d_nv_occ;
    DevBuf<double> d_cos_b, d_sin_b, d_range, d_cos_va, d_sin_va;
    DevBuf<float2> d_csva32;
    bool tma_ok = true, fast32_ok = true, verify_fast = false;   // S3D_NO_TMA, S3D_NO_FAST32, S3D_VERIFY_FAST
    bool serial = false;             // S3D_SERIAL_KERNELS: every pipeline kernel on one stream (exclusive per-kernel timings)
    bool debug_sync = false;         // S3D_DEBUG_SYNC: synchronise and check after every pipeline launch (fault localisation)
    bool one_xstream = false;        // S3D_ONE_XSTREAM: consecutive chunks are expanded one after the other (they still overlap the apply)
    u64 *marks_host = nullptr, *marks_dev = nullptr;   // S3D_MARKS: mapped host counters of started / finished blocks per kernel
    u64 samples_max = 0;             // worst-case samples per frame for these tables
    // chunk working set
    // chunk dedupe table: one allocation {counters[C][GF], keys[C], list[C]} so that a single L2
    // access-policy window can keep it resident between the kernels of a chunk
    DevBuf<uint8_t> spool; u64 scratch_cap = 0;
    bool wide = false;               // counter lanes: u32 (16+16 bits) normally, u64 after a count overflowed
    bool narrow_safe = false;        // proved on the host: no voxel can collect 2^16 samples of a kind in one frame
    std::vector<double> h_range; std::vector<int> h_nv_free, h_nv_occ; double half_aperture = 0.0;
    u64 *skeys = nullptr; void *scnt = nullptr;                      // buffer 0 (also the sharded path's)
    // NBUF chunk buffers and two expand streams: k_expand of chunks c+1 and c+2 overlap each other
    // and k_apply_chunk of chunk c (a chunk alone does not fill the GPU)
    static constexpr int NBUF = N_CHUNK_BUF;
    struct ChunkBuf { u64 *skeys = nullptr; void *scnt = nullptr; ChunkCtr *cc = nullptr;
                      cudaEvent_t expanded = nullptr, freed = nullptr, merged = nullptr; bool used = false; };
    ChunkBuf buf[NBUF];
    cudaStream_t xstream = nullptr, xstream2 = nullptr;  // expand streams (chunks alternate)
    cudaEvent_t x_ev = nullptr;      // orders work queued on xstream before xstream2
    cudaStream_t snap_stream = nullptr;  // per-chunk counter snapshots (device -> pinned host)
    cudaStream_t mstream = nullptr;      // routed map: owner-side merges
    cudaStream_t ctl_stream = nullptr;   // small reads for s3d_ingest_collect (never behind queued chunks)
    DevBuf<double> sum_tab;          // [4][SUMT] running sums of n copies of lo_free / lo_occ and their means, then the (n_free, n_occ) mean table
    ChunkCtr *cc = nullptr;
    DevBuf<DevStats> stats; DevStats *stats_host = nullptr; size_t stats_host_n = 0;
    // staging
    DevBuf<uint8_t> img_dev; DevBuf<double> T_dev; DevBuf<uint16_t> img16_dev;
    Staging stg[2]; int stg_next = 0;
    std::vector<cudaEvent_t> job_ev_pool;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> copy_ev;
    DevBuf<u64> io_keys; DevBuf<double> io_vals; DevBuf<uint8_t> io_flags;
    // export staging
    DevBuf<double> ex_xyz, ex_prob, ex_L; DevBuf<int8_t> ex_cls; DevBuf<int> ex_ijk; DevBuf<float4> ex_f32;
    u64 *ex_counts = nullptr; u64 *ex_counts_host = nullptr; u64 ex_n = 0; bool ex_valid = false;
    u64 mk_n = ~0ull;                // points staged by s3d_export_markers (~0 = none)
    // debug counters: lifetime sample counts per voxel key + the last frame's per-key counts
    bool debug_on = false;
    Slot *life = nullptr;            // same capacity / probing as the voxel table; val holds a u64 count
    DevBuf<ulonglong2> dbg_last; u32 *dbg_last_n = nullptr;
    int apply_bps = 3;               // k_apply_chunk blocks per SM at most (S3D_APPLY_BPS; with the balanced grid: cfg2 equal for 2 and 3, cfg3 update kernel 301 -> 258 us)
    // measurement
    bool prof_on = false;
    std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
    struct Span { size_t e0, e1; int kind; };
    std::vector<Span> spans;
    s3d_profile prof{};
    u64 launches = 0;
};

namespace {

int set_device(s3d_map *m) { CU(cudaSetDevice(m->device)); return 0; }

// With lazy module loading the first launch of a kernel may have to wait for the device to go
// idle; a routed map keeps kernels spinning on peer flags, so every pipeline kernel is loaded
// up front.
template <typename F> int preload(F f) { cudaFuncAttributes at; CU(cudaFuncGetAttributes(&at, f)); return 0; }
int preload_pipeline_kernels()
{
    int rc;
    if ((rc = preload(k_expand<u32, false, true, false>)) || (rc = preload(k_expand<u32, true, true, false>)) ||
        (rc = preload(k_expand<u64, false, true, false>)) || (rc = preload(k_expand<u32, false, false, false>)) ||
        (rc = preload(k_apply_chunk<u32, false>)) || (rc = preload(k_apply_chunk<u64, false>)) ||
        (rc = preload(k_route_signal)) || (rc = preload(k_route_wait)) ||
        (rc = preload(k_route_merge<u32, true>)) || (rc = preload(k_route_merge<u32, false>)) ||
        (rc = preload(k_route_merge<u64, false>)) ||
        (rc = preload(k_fill_slots)) || (rc = preload(k_fill_u64)) || (rc = preload(k_rehash)) || (rc = preload(k_clear_abort)))
        return rc;
    return 0;
}

// CUDA-event bracket around one kernel group, on the launching stream
size_t prof_mark(s3d_map *m, cudaStream_t st)
{
    if (m->ev_used == m->ev_pool.size()) {
        cudaEvent_t e; cudaEventCreate(&e);
        m->ev_pool.push_back(e);
    }
    cudaEventRecord(m->ev_pool[m->ev_used], st);
    return m->ev_used++;
}

int prof_collect(s3d_map *m)
{
    if (m->spans.empty()) { m->ev_used = 0; return 0; }
    CU(cudaStreamSynchronize(m->xstream));
    CU(cudaStreamSynchronize(m->xstream2));
    CU(cudaStreamSynchronize(m->stream));
    for (const auto &sp : m->spans) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, m->ev_pool[sp.e0], m->ev_pool[sp.e1]));
        m->prof.ms[sp.kind] += ms;
    }
    m->spans.clear(); m->ev_used = 0;
    return 0;
}

int launch_fill_table(s3d_map *m, Slot *t, u64 n)
{
    const int blocks = (int)std::min<u64>((n + 255) / 256, (u64)m->n_sm * 16);
    k_fill_slots<<<blocks, 256, 0, m->stream>>>(t, n);
    CU(cudaGetLastError());
    return 0;
}

int pump(s3d_map *m, bool drain);

int fatal_from_flags(s3d_map *m, u32 err)
{
    if (!err) return 0;
    u32 zero = 0;   // clear so that the map stays usable after the caller handles the error
    cudaMemcpyAsync(&m->mc->err, &zero, sizeof zero, cudaMemcpyHostToDevice, m->stream);
    cudaStreamSynchronize(m->stream);
    if (err & ERR_INTERNAL) return fail(S3D_ECUDA, "internal: a bounds guard of the expansion kernel tripped");
    if (err & ERR_VERIFY) return fail(S3D_ECUDA, "S3D_VERIFY_FAST: an accepted fp32 voxel-index estimate differed from the fp64 key");
    if (err & ERR_TABLEFULL) return fail(S3D_ETABLEFULL, "voxel table full (capacity %llu slots)", (unsigned long long)m->cap);
    if (err & ERR_ROUTE_FULL) return fail(S3D_EROUTE, "routed map: a peer inbox overflowed (%llu records per pair; export a larger one)",
                                          (unsigned long long)m->route_cap);
    if (err & ERR_ROUTE_TIMEOUT) return fail(S3D_EROUTE, "routed map: a peer rank did not signal within %.1f s (every rank must ingest the same frames)",
                                             (double)m->route_timeout_ns * 1e-9);
    return fail(S3D_EKEYRANGE, "voxel key outside +-2^20 or non-finite coordinate");
}

// Finish every queued frame (re-running chunks that asked for a retry), then read back the map
// counters and turn fatal device flags into return codes.  Every synchronous entry point
// starts or ends here.
int sync_counters(s3d_map *m)
{
    int rc = pump(m, true);
    if (rc) return rc;
    CU(cudaMemcpyAsync(m->mc_host, m->mc, sizeof(MapCtr), cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    m->count_known = m->mc_host->count;
    m->snap_floor = m->chunk_seq;
    return fatal_from_flags(m, m->mc_host->err);
}

int grow_table(s3d_map *m, u64 new_cap)
{
    new_cap = next_pow2(new_cap);
    if (new_cap <= m->cap) return 0;
    if (getenv("S3D_LOG")) fprintf(stderr, "[s3d] grow_table: %llu -> %llu at chunk %llu\n", (unsigned long long)m->cap, (unsigned long long)new_cap,
                                   (unsigned long long)m->chunk_seq);
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
        ++m->n_grows;
    }
    if (m->debug_on) {
        // the lifetime table keeps the voxel table's geometry
        Slot *nl = nullptr;
        e = cudaMalloc(&nl, new_cap * sizeof(Slot));
        if (e != cudaSuccess) { cudaFree(nt); return fail(S3D_ETABLEFULL, "cannot grow the debug-counter table to %llu slots: %s",
                                                          (unsigned long long)new_cap, cudaGetErrorString(e)); }
        if ((rc = launch_fill_table(m, nl, new_cap))) return rc;
        if (m->life) {
            const int blocks = (int)std::min<u64>((m->cap + 255) / 256, (u64)m->n_sm * 16);
            k_rehash<<<blocks, 256, 0, m->stream>>>(m->life, m->cap, nl, new_cap - 1, m->mc);
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(m->stream));
            CU(cudaFree(m->life));
        }
        m->life = nl;
    }
    m->table = nt; m->cap = new_cap;
    m->ex_valid = false;
    return 0;
}

// the device gate: a chunk is applied only if count + unique(chunk) stays within this
u64 table_limit(const s3d_map *m) { return m->cap / 2; }

// room for `extra` more voxels on top of the last known count (stream must be idle)
int ensure_room(s3d_map *m, u64 extra)
{
    u64 cap = m->cap;
    while (m->count_known + extra > cap / 2) cap *= 2;
    return cap > m->cap ? grow_table(m, cap) : 0;
}

template <typename T> int upload(DevBuf<T> &b, const T *src, size_t n, cudaStream_t s)
{
    int rc = b.ensure(n); if (rc) return rc;
    CU(cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
    return 0;
}

size_t lane_bytes(const s3d_map *m) { return m->wide ? sizeof(u64) : sizeof(u32); }

// Upper bound of the samples of one kind that one frame can put into one voxel, from the tables:
// every sample of range bin r lies exactly range_m[r] from the sonar origin, a voxel spans at
// most sqrt(3)*res in distance (so only a few consecutive bins reach it), and on the arc of one
// (beam, bin) fan consecutive samples are a chord 2*r*sin(ha/(2*nv)) apart.  When the bound stays
// below 2^16 the narrow counter lanes cannot overflow and k_expand skips the overflow watch.
void update_count_bound(s3d_map *m)
{
    m->narrow_safe = false;
    if (!m->have_params || m->h_range.empty()) return;
    const double diam = std::sqrt(3.0) * m->p.res * (1.0 + 1e-9);
    const size_t H = m->h_range.size();
    const double rr = H > 1 ? m->h_range[1] - m->h_range[0] : 0.0;
    const int step = m->tab.free_step > 0 ? m->tab.free_step : 1;
    const double bins = rr > 0.0 ? std::floor(diam / rr) + 1.0 : (double)H;
    const double bins_occ = std::min<double>({(double)m->tab.occ_window, bins, (double)H});
    const double bins_free = std::min<double>((double)((H + step - 1) / step), rr > 0.0 ? std::floor(diam / (rr * step)) + 1.0 : (double)H);
    auto per_fan = [&](int nv, double r) -> double {
        if (nv <= 0) return 0.0;
        const double all = 2.0 * nv + 1.0;
        const double chord = 2.0 * r * std::sin(m->half_aperture / (2.0 * nv));
        return chord > 0.0 ? std::min(all, std::floor(diam / chord) + 1.0) : all;
    };
    double occ = 0.0, fre = 0.0;
    for (size_t r = 0; r < H; ++r) {
        occ = std::max(occ, per_fan(m->h_nv_occ[r], m->h_range[r]));
        if (r % (size_t)step == 0) fre = std::max(fre, per_fan(m->h_nv_free[r], m->h_range[r]));
    }
    const double nb = (double)m->tab.n_beams;
    m->narrow_safe = nb * bins_occ * occ <= 65535.0 && nb * bins_free * fre <= 65535.0;
}

// (re)allocate and wipe the chunk dedupe tables; the streams must be idle
int ensure_scratch(s3d_map *m, u64 want_cap, bool wipe, bool force_realloc = false)
{
    want_cap = next_pow2(std::max<u64>(want_cap, 1u << 12));
    if (want_cap > (1ull << 31)) return fail(S3D_ENOMEM, "chunk dedupe table would exceed 2^31 entries");
    const bool realloc = want_cap > m->scratch_cap || force_realloc;
    if (getenv("S3D_LOG")) fprintf(stderr, "[s3d] ensure_scratch: %llu -> %llu realloc=%d wipe=%d at chunk %llu\n", (unsigned long long)m->scratch_cap,
                                   (unsigned long long)want_cap, (int)realloc, (int)wipe, (unsigned long long)m->chunk_seq);
    const size_t cnt_bytes = lane_bytes(m) * (size_t)std::max<u64>(want_cap, m->scratch_cap) * GF;
    if (realloc) {
        want_cap = std::max<u64>(want_cap, m->scratch_cap);
        const size_t key_bytes = sizeof(u64) * (size_t)want_cap;
        const size_t one = cnt_bytes + key_bytes;
        const size_t pool = s3d_map::NBUF * one;
        CU(cudaStreamSynchronize(m->xstream));
        CU(cudaStreamSynchronize(m->xstream2));
        CU(cudaStreamSynchronize(m->stream));
        int rc = m->spool.ensure(pool); if (rc) return rc;
        for (int b = 0; b < s3d_map::NBUF; ++b) {
            uint8_t *base = m->spool.p + (size_t)b * one;
            m->buf[b].scnt = base;
            m->buf[b].skeys = reinterpret_cast<u64 *>(base + cnt_bytes);
            m->buf[b].used = false;
        }
        m->scnt = m->buf[0].scnt; m->skeys = m->buf[0].skeys;
        m->scratch_cap = want_cap;
    }
    if (realloc || wipe) {
        const int blocks = (int)std::min<u64>((m->scratch_cap + 255) / 256, (u64)m->n_sm * 16);
        CU(cudaStreamSynchronize(m->xstream));
        CU(cudaStreamSynchronize(m->xstream2));
        for (int b = 0; b < s3d_map::NBUF; ++b) {
            k_fill_u64<<<blocks, 256, 0, m->stream>>>(m->buf[b].skeys, m->scratch_cap, EMPTY_KEY);
            CU(cudaGetLastError());
            CU(cudaMemsetAsync(m->buf[b].scnt, 0, cnt_bytes, m->stream));
            m->buf[b].used = false;
        }
        CU(cudaStreamSynchronize(m->stream));
    }
    return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tma_encoder()
{
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

template <typename CT, bool CHECK>
void launch_expand_t(s3d_map *m, const ExpandArgs &a, const CUtensorMap &tmap, int grid, size_t smem, cudaStream_t st)
{
    if (a.rt.world > 1) k_expand<CT, CHECK, true, false><<<grid, EX_THREADS, smem, st>>>(a, tmap);
    else if (m->verify_fast) k_expand<CT, CHECK, false, true><<<grid, EX_THREADS, smem, st>>>(a, tmap);
    else k_expand<CT, CHECK, false, false><<<grid, EX_THREADS, smem, st>>>(a, tmap);
}

void launch_expand(s3d_map *m, ExpandArgs &a, int n_beams, int g, cudaStream_t st)
{
    const DevTables &t = a.tab;
    a.n_frames = g;
    a.fast32_ok = m->fast32_ok ? 1 : 0;
    a.box_rows = strip_box_rows(t.H);
    // The image strip of a tile goes through TMA when the frames can be described as a 2-D byte tensor
    // whose 16-byte column strips hold whole groups of processed beams: regular beam columns, 16 | W,
    // 16-byte aligned frames, and a strip that fits the combiner's memory.
    const int per_strip = (t.col_step > 0 && TMA_STRIP_BYTES % t.col_step == 0) ? TMA_STRIP_BYTES / t.col_step : 0;
    bool tma = m->tma_ok && tma_encoder() && per_strip > 0 && t.W % 16 == 0 && (reinterpret_cast<uintptr_t>(a.imgs) % 16 == 0) &&
               a.img_stride % 16 == 0 && strip_smem_bytes(t.H) <= (sizeof(u32) * 2 + sizeof(unsigned short)) * LT_CAP &&
               (u64)g * (u64)t.H < (1ull << 31);
    // beams per tile: all warps of a block share the passes of its beams, so fewer beams per tile means
    // shorter tiles.  A small slice (a rank of a routed map) is cut finer so that the launch still has a
    // few hundred tiles.  With strips, a tile must not straddle two of them.
    // (measured at cfg2: 4 beams per tile -- 1024 tiles per 16 frames -- beats 8, 2 and 1)
    int bpb = m->bpb_env > 0 ? std::min(m->bpb_env, EX_MAXB) : std::max(1, std::min(EX_WARPS / 2, n_beams / 16));
    if (tma) {
        int q = per_strip;
        while (q > 1 && q > bpb) q /= 2;                 // largest power-of-two fraction of a strip that is <= the target
        if (per_strip % q != 0 || a.beam_lo % q != 0) tma = false; else bpb = q;
    }
    size_t smem = expand_smem_bytes(t.H, t.free_step, t.occ_window, bpb);
    // (the fan lists grow with H: keep EX_BPS blocks resident per SM -- 227 KB less 2 KB of static + reserved memory per block)
    const size_t smem_fit = (size_t)(227 * 1024) / EX_BPS - 2048;
    while (smem > smem_fit && bpb > 1) { bpb = tma ? bpb / 2 : bpb - 1; smem = expand_smem_bytes(t.H, t.free_step, t.occ_window, bpb); }
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    if (tma) {
        const cuuint64_t dims[2] = {(cuuint64_t)t.W, (cuuint64_t)g * (cuuint64_t)t.H};
        const cuuint64_t strides[1] = {(cuuint64_t)t.W};
        const cuuint32_t box[2] = {(cuuint32_t)TMA_STRIP_BYTES, (cuuint32_t)a.box_rows};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = tma_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(a.imgs), dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) tma = false;
    }
    a.use_tma = tma ? 1 : 0;
    a.bpb = bpb;
    const int tiles = ((n_beams + bpb - 1) / bpb) * g;
    const int grid = std::max(1, std::min(tiles, m->n_sm * EX_BPS));
    if (m->wide) launch_expand_t<u64, false>(m, a, tmap, grid, smem, st);
    else if (m->narrow_safe) launch_expand_t<u32, false>(m, a, tmap, grid, smem, st);
    else launch_expand_t<u32, true>(m, a, tmap, grid, smem, st);
    m->launches += 1;
}

u64 *trace_slot(s3d_map *m, int which)
{
    if (!m->trace.p || m->chunk_seq >= s3d_map::TRACE_CHUNKS) return nullptr;
    return m->trace.p + m->chunk_seq * s3d_map::TRACE_W + 2 * which;
}

void launch_apply(s3d_map *m, u64 *skeys, void *scnt, int g, ChunkCtr *cc, DevStats *st, cudaStream_t stream)
{
    // one tile of AP_TILE slots per block and step; up to `apply_bps` blocks per SM
    const u64 tiles = m->scratch_cap / AP_TILE;
    // (the fewest blocks that keep the busiest block's tile count: 1024 tiles on 296 block slots are 4 per block
    // either way, and 256 blocks leave more of the SMs to the expansion running beside)
    u64 nb = std::max<u64>(1, std::min<u64>(tiles, (u64)m->n_sm * (u64)m->apply_bps));
    const u64 per = (tiles + nb - 1) / nb;
    if (per > 0) nb = std::max<u64>(1, (tiles + per - 1) / per);
    const int blocks = (int)nb;
    ApplyArgs a;
    a.skeys = skeys; a.scnt = scnt; a.n_slots = (u32)m->scratch_cap; a.g = g; a.cc = cc; a.st = st;
    a.table = m->table; a.tmask = m->cap - 1; a.p = m->p; a.sum_tab = m->sum_tab.p; a.mc = m->mc;
    a.table_limit = table_limit(m); a.seq = m->chunk_seq; a.trace = trace_slot(m, 4);
    a.marks = m->marks_dev;
    a.life = m->life; a.last = m->dbg_last.p; a.last_n = m->dbg_last_n; a.last_cap = (u32)std::min<size_t>(m->dbg_last.n, 0xffffffffu);
    if (m->debug_on) {
        cudaMemsetAsync(m->dbg_last_n, 0, sizeof(u32), stream);      // the list restarts with every chunk: it ends up holding the last frame
        if (m->wide) k_apply_chunk<u64, true><<<blocks, AP_THREADS, apply_smem_bytes<u64>(), stream>>>(a);
        else k_apply_chunk<u32, true><<<blocks, AP_THREADS, apply_smem_bytes<u32>(), stream>>>(a);
    } else {
        if (m->wide) k_apply_chunk<u64, false><<<blocks, AP_THREADS, apply_smem_bytes<u64>(), stream>>>(a);
        else k_apply_chunk<u32, false><<<blocks, AP_THREADS, apply_smem_bytes<u32>(), stream>>>(a);
    }
    m->launches += 1;
}

int enqueue_chunk(s3d_map *m, const Job &j, int64_t base, int g)
{
    const DevTables &tab = m->tab;
    const size_t img_stride = (size_t)tab.H * tab.W;
    const uint8_t *imgs = j.imgs + (size_t)base * img_stride;
    // measured: a single map runs 2 % faster on 3 buffers than on 4 (L2); a routed map, which queues
    // deeper, loses 8 % on 3.  Switching happens only between drained states (s3d_route_enable syncs).
    const int n_buf = (m->route_on && m->shard_world > 1) ? s3d_map::NBUF : s3d_map::NBUF - 1;
    s3d_map::ChunkBuf &cb = m->buf[m->chunk_seq % (u64)n_buf];
    cudaStream_t xs = (m->chunk_seq & 1) ? m->xstream2 : m->xstream, as = m->stream;
    if (m->one_xstream) xs = m->xstream;
    if (m->serial) {                 // measurement mode: no overlap, so the event spans are kernel durations
        xs = as;
        CU(cudaStreamWaitEvent(as, m->x_ev, 0));     // (submit_frames zeroed the counters on the expand stream)
    }
    // ---- expand stream: first hits + expansion into this chunk's dedupe buffer.  It may run while
    // earlier chunks are still being expanded or applied; it only waits for its own buffer to be drained.
    if (cb.used) CU(cudaStreamWaitEvent(xs, cb.freed, 0));
    const size_t e1 = m->prof_on ? prof_mark(m, xs) : 0;
    ExpandArgs a;
    a.imgs = imgs; a.img_stride = img_stride;
    a.T = j.T + base * 16;
    a.tab = tab; a.p = m->p;
    a.skeys = cb.skeys; a.scnt = cb.scnt; a.smask = (u32)(m->scratch_cap - 1);
    a.cc = cb.cc; a.stats = j.stats + base; a.mc = m->mc;
    a.seq = m->chunk_seq;
    a.trace = trace_slot(m, 1);
    a.marks = m->marks_dev;
    a.beam_lo = 0; a.beam_hi = tab.n_beams;
    a.own_rank = (u32)m->shard_rank; a.own_world = m->shard_filter ? (u32)m->shard_world : 1u;
    a.rt = RouteCtx{1u, 0u, 0u, 0ull, nullptr, nullptr, 0ull, 0ull};
    const bool routed = m->route_on && m->shard_world > 1;
    if (routed) {
        if (m->route_by_chunk) {
            // split by chunks: rank (c mod world) expands every beam of chunk c -- a launch as large as a single
            // GPU's, so the expansion scales with the ranks -- and the others only tell their peers that nothing is
            // coming from them; voxels of other owners travel to them
            const bool mine = (int)(m->route_seq % (u64)m->shard_world) == m->shard_rank;
            a.beam_lo = 0; a.beam_hi = mine ? tab.n_beams : 0;
        } else {
            // split by beams: this rank's contiguous slice of the processed beams of every chunk
            a.beam_lo = (int)((int64_t)tab.n_beams * m->shard_rank / m->shard_world);
            a.beam_hi = (int)((int64_t)tab.n_beams * (m->shard_rank + 1) / m->shard_world);
        }
        a.own_world = 1u;
        a.rt = RouteCtx{(u32)m->shard_world, (u32)m->shard_rank, (u32)(m->route_seq % ROUTE_DEPTH), m->route_cap, m->d_peers.p,
                        m->route_cursor.p, m->route_seq, m->route_timeout_ns};
        // the owners must have merged what this rank sent them ROUTE_DEPTH chunks ago (same parity)
        if (m->route_seq >= (u64)ROUTE_DEPTH) {
            const RouteHdr *h = reinterpret_cast<const RouteHdr *>(m->xblock);
            k_route_wait<<<1, ROUTE_MAX_WORLD, 0, xs>>>(h->ack_seq[a.rt.parity], a.rt.world, a.rt.rank, m->route_seq - ROUTE_DEPTH + 1,
                                                       m->route_timeout_ns, m->mc, trace_slot(m, 0));
            m->launches += 1;
        }
    }
    // (routed: the kernel's last block publishes the record counts to the peers)
    if (a.beam_hi > a.beam_lo) launch_expand(m, a, a.beam_hi - a.beam_lo, g, xs);
    else if (routed) { k_route_signal<<<1, ROUTE_MAX_WORLD, 0, xs>>>(a.rt); m->launches += 1; }
    if (m->debug_sync) {
        cudaError_t e = cudaStreamSynchronize(xs);
        if (e != cudaSuccess) return fail(S3D_ECUDA, "k_expand of chunk %llu (frames %lld.., g=%d, tma=%d, bpb=%d): %s", (unsigned long long)m->chunk_seq,
                                          (long long)base, g, a.use_tma, a.bpb, cudaGetErrorString(e));
    }
    const size_t e2 = m->prof_on ? prof_mark(m, xs) : 0;
    CU(cudaEventRecord(cb.expanded, xs));
    if (routed) {
        // ---- merge stream: wait for the sources (one small block), merge their records (beside our own k_expand), acknowledge
        cudaStream_t ms = m->mstream;
        if (cb.used) CU(cudaStreamWaitEvent(ms, cb.freed, 0));
        {
            const RouteHdr *h = reinterpret_cast<const RouteHdr *>(m->xblock);
            k_route_wait<<<1, ROUTE_MAX_WORLD, 0, ms>>>(h->flag_seq[a.rt.parity], a.rt.world, a.rt.rank, m->route_seq + 1,
                                                       m->route_timeout_ns, m->mc, trace_slot(m, 2));
            m->launches += 1;
        }
        const int mb = m->n_sm * 2;
        if (m->wide)
            k_route_merge<u64, false><<<mb, 256, 0, ms>>>(a.rt, cb.skeys, static_cast<u64 *>(cb.scnt), a.smask, cb.cc, m->mc,
                                                         m->chunk_seq, trace_slot(m, 3));
        else if (m->narrow_safe)
            k_route_merge<u32, false><<<mb, 256, 0, ms>>>(a.rt, cb.skeys, static_cast<u32 *>(cb.scnt), a.smask, cb.cc, m->mc,
                                                         m->chunk_seq, trace_slot(m, 3));
        else
            k_route_merge<u32, true><<<mb, 256, 0, ms>>>(a.rt, cb.skeys, static_cast<u32 *>(cb.scnt), a.smask, cb.cc, m->mc,
                                                        m->chunk_seq, trace_slot(m, 3));
        CU(cudaEventRecord(cb.merged, ms));
        m->launches += 1;
        ++m->route_seq;
    }
    // ---- apply stream: the gate and the chunk's frames, in order, into the voxel table
    CU(cudaStreamWaitEvent(as, cb.expanded, 0));
    if (routed) CU(cudaStreamWaitEvent(as, cb.merged, 0));
    const size_t e3 = m->prof_on ? prof_mark(m, as) : 0;
    launch_apply(m, cb.skeys, cb.scnt, g, cb.cc, j.stats + base, as);
    CU(cudaGetLastError());
    if (m->debug_sync) {
        cudaError_t e = cudaStreamSynchronize(as);
        if (e != cudaSuccess) return fail(S3D_ECUDA, "k_apply_chunk of chunk %llu (frames %lld.., g=%d, cap=%llu, scratch=%llu): %s", (unsigned long long)m->chunk_seq,
                                          (long long)base, g, (unsigned long long)m->cap, (unsigned long long)m->scratch_cap, cudaGetErrorString(e));
    }
    CU(cudaEventRecord(cb.freed, as));
    cb.used = true;
    if (m->prof_on) {
        const size_t e4 = prof_mark(m, as);
        m->spans.push_back({e1, e2, S3D_K_EXPAND});
        m->spans.push_back({e3, e4, S3D_K_APPLY});
        m->prof.launches[S3D_K_EXPAND] += 1; m->prof.launches[S3D_K_APPLY] += 1;
        m->prof.frames += (u64)g;
    }
    // counter snapshot for the host (growth estimates, retry flags): copied on its own stream so
    // that the copy engine's latency is not between this chunk's apply and the next chunk's kernels
    const int ri = (int)(m->chunk_seq % s3d_map::RING);
    CU(cudaStreamWaitEvent(m->snap_stream, cb.freed, 0));
    CU(cudaMemcpyAsync(&m->snap_host[ri], m->mc, sizeof(MapCtr), cudaMemcpyDeviceToHost, m->snap_stream));
    CU(cudaEventRecord(m->snap_ev[ri], m->snap_stream));
    m->inflight[ri] = InFlight{m->chunk_seq, j.id, base, g};
    ++m->chunk_seq;
    m->ex_valid = false; m->mk_n = ~0ull;
    return 0;
}

// A chunk asked for a retry (dedupe table too loaded, or the voxel table would pass its load
// bound).  Nothing of that chunk or of any later one has touched the voxel table.  Enlarge what
// was short, wipe the chunk working set and rewind the job queue to the chunk's first frame.
int recover(s3d_map *m)
{
    CU(cudaStreamSynchronize(m->xstream));          // later chunks may still be expanding
    CU(cudaStreamSynchronize(m->xstream2));
    CU(cudaStreamSynchronize(m->mstream));
    CU(cudaMemcpyAsync(m->mc_host, m->mc, sizeof(MapCtr), cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    const MapCtr mc = *m->mc_host;
    if (!mc.abort) return 0;
    if (m->route_on && m->shard_world > 1)
        return fail(S3D_EROUTE, "routed map: chunk %llu needs a re-run (flags %u: 1 = dedupe table, 2 = voxel table, 4 = counter width); "
                                "peers cannot replay it -- reserve capacity (s3d_reserve) before ingesting",
                    (unsigned long long)mc.abort_seq, mc.abort);
    if (getenv("S3D_LOG")) fprintf(stderr, "[s3d] recover: abort=%u abort_seq=%llu chunk_seq=%llu count=%llu cap=%llu scratch=%llu\n", mc.abort,
                                   (unsigned long long)mc.abort_seq, (unsigned long long)m->chunk_seq, (unsigned long long)mc.count,
                                   (unsigned long long)m->cap, (unsigned long long)m->scratch_cap);
    const InFlight inf = m->inflight[mc.abort_seq % s3d_map::RING];
    if (inf.seq != mc.abort_seq) return fail(S3D_ECUDA, "internal: lost track of chunk %llu", (unsigned long long)mc.abort_seq);
    ++m->n_retries;
    m->count_known = mc.count;
    m->snap_floor = m->chunk_seq;
    int rc;
    if (mc.abort & ABORT_TABLE) {
        if ((rc = grow_table(m, m->cap * 2))) return rc;
    }
    const bool widen = (mc.abort & ABORT_NARROW) && !m->wide;
    if (widen) m->wide = true;
    if ((rc = ensure_scratch(m, (mc.abort & ABORT_SCRATCH) ? m->scratch_cap * 2 : m->scratch_cap, true, widen))) return rc;
    k_clear_abort<<<1, 1, 0, m->stream>>>(m->mc, m->cc);
    CU(cudaGetLastError());
    bool hit = false;
    for (Job &j : m->jobs) {
        if (j.id == inf.job) { j.next = inf.base; hit = true; }
        else if (hit) j.next = 0;
        if (hit && j.next < j.n)
            CU(cudaMemsetAsync(j.stats + j.next, 0, sizeof(DevStats) * (size_t)(j.n - j.next), m->stream));
    }
    if (!hit) return fail(S3D_ECUDA, "internal: retry for an unknown job");
    // The cleared retry flag and the zeroed counters were queued on the apply stream; the re-run's k_expand goes to
    // an expand stream that is not ordered behind it.  Without this wait a block of the re-run could still see the
    // old flag and leave without expanding its tile -- frames lost without a trace (found with tools/stress_growth.py).
    CU(cudaStreamSynchronize(m->stream));
    return 0;
}

// Enqueue the queued frames chunk by chunk with a bounded launch-ahead; with `drain`, also wait
// until everything is applied (re-running chunks that asked for a retry).
int pump(s3d_map *m, bool drain)
{
    // Chunks the host may queue behind the one being enqueued.  A routed map queues deeper than
    // its chunk buffers (buffer reuse is ordered on the device by events), which keeps the host's
    // wait for an old snapshot and its launch latency off the critical path of the exchange; a
    // single map runs best with a short queue (more chunks in flight only fight over L2).
    // (split by chunks: `world` chunks are being expanded at a time, one per rank, so the queue must reach past them)
    const bool routed_q = m->route_on && m->shard_world > 1;
    const int LOOKAHEAD = m->lookahead_env > 0 ? std::min(m->lookahead_env, s3d_map::RING - 2)
                                               : (routed_q ? (m->route_by_chunk ? std::min(s3d_map::RING - 2, std::max(6, 2 * m->shard_world + 2)) : 5) : 2);
    // (an inbox region is reused on the expand stream that used it before; a chunk buffer may change
    // streams, its reuse is ordered by the `freed` event behind the apply that drained it)
    static_assert(ROUTE_DEPTH % 2 == 0, "an inbox region is always reused on the same expand stream");
    for (;;) {
        Job *job = nullptr;
        for (Job &j : m->jobs) if (j.next < j.n) { job = &j; break; }
        if (job) {
            const u64 q = m->chunk_seq;
            if (q >= m->snap_floor + 1 + LOOKAHEAD) {
                // wait for the counters of chunk q-1-LOOKAHEAD: bounds the launch-ahead and is
                // where retries and growth needs are noticed early
                const int ri = (int)((q - 1 - LOOKAHEAD) % s3d_map::RING);
                {
                    cudaError_t e_ = cudaEventSynchronize(m->snap_ev[ri]);
                    if (e_ != cudaSuccess) {
                        if (m->marks_host)
                            return fail(S3D_ECUDA, "%s while waiting for chunk %llu of %llu queued; blocks started/finished: k_expand %llu/%llu, k_apply_chunk %llu/%llu",
                                        cudaGetErrorString(e_), (unsigned long long)(q - 1 - LOOKAHEAD), (unsigned long long)q,
                                        (unsigned long long)m->marks_host[0], (unsigned long long)m->marks_host[1],
                                        (unsigned long long)m->marks_host[2], (unsigned long long)m->marks_host[3]);
                        return fail(S3D_ECUDA, "cudaEventSynchronize(m->snap_ev[ri]): %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__);
                    }
                }
                const MapCtr &sn = m->snap_host[ri];
                if (sn.abort) { int rc = recover(m); if (rc) return rc; continue; }
                m->count_known = sn.count;
                m->unique_est = std::max<u64>(sn.last_unique, m->unique_est - m->unique_est / 8);
            }
            // grow ahead of the gate when the chunks in flight could reach it (saves a retry)
            // (a routed map cannot re-run a chunk, so it keeps twice the margin)
            const u64 margin = (u64)(LOOKAHEAD + 2) * m->unique_est * (m->route_on ? 2 : 1);
            if (m->count_known + margin > table_limit(m) || 3 * m->unique_est > 2 * m->scratch_cap) {
                CU(cudaMemcpyAsync(m->mc_host, m->mc, sizeof(MapCtr), cudaMemcpyDeviceToHost, m->stream));
                CU(cudaStreamSynchronize(m->stream));
                if (m->mc_host->abort) { int rc = recover(m); if (rc) return rc; continue; }
                m->count_known = m->mc_host->count;
                m->snap_floor = m->chunk_seq;
                int rc;
                if (m->count_known + margin > table_limit(m) && (rc = grow_table(m, m->cap * 2))) return rc;
                if (3 * m->unique_est > 2 * m->scratch_cap && (rc = ensure_scratch(m, m->scratch_cap * 2, false))) return rc;
            }
            const int g = (int)std::min<int64_t>(GF, job->n - job->next);
            if (m->prof_on && m->ev_used > 4096) { int rc = prof_collect(m); if (rc) return rc; }
            int rc = enqueue_chunk(m, *job, job->next, g);
            if (rc) return rc;
            job->next += g;
            if (job->next == job->n) {
                job->last_seq = m->chunk_seq - 1;
                if (job->done_ev) CU(cudaEventRecord(job->done_ev, m->stream));
            }
            continue;
        }
        if (!drain || m->jobs.empty()) return 0;
        CU(cudaMemcpyAsync(m->mc_host, m->mc, sizeof(MapCtr), cudaMemcpyDeviceToHost, m->stream));
        CU(cudaStreamSynchronize(m->stream));
        if (m->mc_host->abort) { int rc = recover(m); if (rc) return rc; continue; }
        m->count_known = m->mc_host->count;
        m->unique_est = std::max<u64>(m->unique_est, m->mc_host->last_unique);
        m->snap_floor = m->chunk_seq;
        for (Job &j : m->jobs) if (j.done_ev) m->job_ev_pool.push_back(j.done_ev);
        m->jobs.clear();
        return 0;
    }
}

// queue n frames whose images / transforms sit in device memory
int submit_frames(s3d_map *m, const uint8_t *imgs_dev, int64_t n, const double *T_dev, DevStats *stats_dev,
                  bool want_done_event = false)
{
    // zeroed on the expand stream: k_expand is the first writer (num_samples)
    CU(cudaMemsetAsync(stats_dev, 0, sizeof(DevStats) * (size_t)n, m->xstream));
    CU(cudaEventRecord(m->x_ev, m->xstream));
    CU(cudaStreamWaitEvent(m->xstream2, m->x_ev, 0));
    if (m->tab.n_beams == 0 || m->tab.H == 0) {
        // nothing to expand; len(voxels) still has to be reported
        int rc = pump(m, true); if (rc) return rc;
        CU(cudaStreamSynchronize(m->xstream));
        for (int64_t f = 0; f < n; ++f)
            CU(cudaMemcpyAsync(&stats_dev[f].n_voxels, &m->mc->count, sizeof(u64), cudaMemcpyDeviceToDevice, m->stream));
        return 0;
    }
    Job j;
    j.id = ++m->job_seq; j.imgs = imgs_dev; j.T = T_dev; j.stats = stats_dev; j.n = n; j.next = 0;
    if (want_done_event) {
        if (m->job_ev_pool.empty()) { CU(cudaEventCreateWithFlags(&j.done_ev, cudaEventDisableTiming)); }
        else { j.done_ev = m->job_ev_pool.back(); m->job_ev_pool.pop_back(); }
    }
    m->jobs.push_back(j);
    return pump(m, false);
}

// Block until every frame of the jobs up to `job_id` is applied (re-running chunks that asked for a
// retry), without draining what was queued behind them.
int wait_jobs_through(s3d_map *m, u64 job_id)
{
    for (;;) {
        int rc = pump(m, false); if (rc) return rc;
        Job *job = nullptr;
        for (Job &j : m->jobs) if (j.id == job_id) { job = &j; break; }
        if (!job) return 0;                                  // a full drain has confirmed it already
        if (job->done_ev) CU(cudaEventSynchronize(job->done_ev));
        else CU(cudaStreamSynchronize(m->stream));
        // (its own stream: the snapshot stream may be queued behind chunks of later batches)
        CU(cudaMemcpyAsync(m->mc_host, m->mc, sizeof(MapCtr), cudaMemcpyDeviceToHost, m->ctl_stream));
        CU(cudaStreamSynchronize(m->ctl_stream));
        if (m->mc_host->abort && m->mc_host->abort_seq <= job->last_seq) {
            if ((rc = recover(m))) return rc;                // the chunk is queued again; go round
            continue;
        }
        if (m->mc_host->err) return fatal_from_flags(m, m->mc_host->err);
        break;
    }
    while (!m->jobs.empty() && m->jobs.front().id <= job_id) {
        if (m->jobs.front().done_ev) m->job_ev_pool.push_back(m->jobs.front().done_ev);
        m->jobs.pop_front();
    }
    return 0;
}

int check_ready(s3d_map *m)
{
    if (!m) return fail(S3D_EINVAL, "null map");
    if (!m->have_params) return fail(S3D_EINVAL, "s3d_set_params has not been called");
    if (!m->have_tables) return fail(S3D_EINVAL, "s3d_set_tables has not been called");
    return 0;
}

void stats_to_abi(const DevStats &s, s3d_frame_stats &o)
{
    o.num_occupied = (int64_t)s.n_occ; o.num_free = (int64_t)s.n_free;
    o.num_voxels = (int64_t)s.n_voxels; o.num_samples = (int64_t)s.n_samples;
    o.max_samples_per_voxel = (int64_t)s.max_in_frame; o.num_voxels_gt10 = (int64_t)s.n_gt10;
    o.max_total_samples = (int64_t)s.max_total; o.reserved = 0;
}

int finish_stats(s3d_map *m, const DevStats *stats_dev, int64_t n, s3d_frame_stats *out)
{
    if (out && m->stats_host_n < (size_t)n) {
        if (m->stats_host) cudaFreeHost(m->stats_host);
        m->stats_host = nullptr; m->stats_host_n = 0;
        CU(cudaMallocHost(&m->stats_host, sizeof(DevStats) * (size_t)n));
        m->stats_host_n = (size_t)n;
    }
    // Fast path (every single-frame call takes it): all frames are enqueued, so the map counters
    // and the per-frame counters are copied behind the last apply and one synchronisation ends the
    // call.  Only a chunk that asked for a retry falls through to the general drain.
    bool all_enqueued = !m->jobs.empty();
    for (const Job &j : m->jobs) if (j.next < j.n) all_enqueued = false;
    if (all_enqueued) {
        CU(cudaMemcpyAsync(m->mc_host, m->mc, sizeof(MapCtr), cudaMemcpyDeviceToHost, m->stream));
        if (out) CU(cudaMemcpyAsync(m->stats_host, stats_dev, sizeof(DevStats) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
        CU(cudaStreamSynchronize(m->stream));
        if (!m->mc_host->abort) {
            m->count_known = m->mc_host->count;
            m->unique_est = std::max<u64>(m->unique_est, m->mc_host->last_unique);
            m->snap_floor = m->chunk_seq;
            for (Job &j : m->jobs) if (j.done_ev) m->job_ev_pool.push_back(j.done_ev);
            m->jobs.clear();
            int rc = fatal_from_flags(m, m->mc_host->err);
            if (rc) return rc;
            if (out) for (int64_t f = 0; f < n; ++f) stats_to_abi(m->stats_host[f], out[f]);
            return 0;
        }
    }
    int rc = sync_counters(m);
    if (rc) return rc;
    if (!out) return 0;
    CU(cudaMemcpyAsync(m->stats_host, stats_dev, sizeof(DevStats) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    for (int64_t f = 0; f < n; ++f) stats_to_abi(m->stats_host[f], out[f]);
    return 0;
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
    CU(cudaMallocHost(&m->snap_host, sizeof(MapCtr) * s3d_map::RING));
    for (int i = 0; i < s3d_map::RING; ++i) CU(cudaEventCreateWithFlags(&m->snap_ev[i], cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&m->xstream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&m->xstream2, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&m->snap_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&m->mstream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&m->ctl_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&m->x_ev, cudaEventDisableTiming));
    { const char *e = getenv("S3D_WIDE_LANES"); m->wide = e && atoi(e) != 0; }
    { const char *e = getenv("S3D_LOOKAHEAD"); if (e) m->lookahead_env = atoi(e); }
    { const char *e = getenv("S3D_BEAMS_PER_BLOCK"); if (e) m->bpb_env = atoi(e); }
    { const char *e = getenv("S3D_APPLY_BPS"); if (e && atoi(e) > 0) m->apply_bps = std::min(atoi(e), 8); }
    { const char *e = getenv("S3D_NO_TMA"); if (e && atoi(e) != 0) m->tma_ok = false; }
    { const char *e = getenv("S3D_ROUTE_SPLIT"); if (e && !strcmp(e, "beams")) m->route_by_chunk = false; }
    { const char *e = getenv("S3D_SERIAL_KERNELS"); if (e && atoi(e) != 0) m->serial = true; }
    { const char *e = getenv("S3D_DEBUG_SYNC"); if (e && atoi(e) != 0) m->debug_sync = true; }
    { const char *e = getenv("S3D_ONE_XSTREAM"); if (e && atoi(e) != 0) m->one_xstream = true; }
    if (const char *e = getenv("S3D_MARKS")) if (atoi(e) != 0) {
        CU(cudaHostAlloc(&m->marks_host, 8 * sizeof(u64), cudaHostAllocMapped));
        memset(m->marks_host, 0, 8 * sizeof(u64));
        CU(cudaHostGetDevicePointer(&m->marks_dev, m->marks_host, 0));
    }
    { const char *e = getenv("S3D_NO_FAST32"); if (e && atoi(e) != 0) m->fast32_ok = false; }
    { const char *e = getenv("S3D_VERIFY_FAST"); if (e && atoi(e) != 0) m->verify_fast = true; }
    { const char *e = getenv("S3D_SCRATCH_CAP"); if (e && atoll(e) > 0) m->scratch_env = (u64)atoll(e); }
    if (const char *e = getenv("S3D_TRACE")) if (atoi(e) != 0) {
        if (m->trace.ensure(s3d_map::TRACE_CHUNKS * s3d_map::TRACE_W)) return S3D_ENOMEM;
        std::vector<u64> init(s3d_map::TRACE_CHUNKS * s3d_map::TRACE_W);
        for (size_t i = 0; i < init.size(); ++i) init[i] = (i & 1) ? 0ull : ~0ull;
        CU(cudaMemcpy(m->trace.p, init.data(), init.size() * sizeof(u64), cudaMemcpyHostToDevice));
    }
    CU(cudaMalloc(&m->cc, sizeof(ChunkCtr) * s3d_map::NBUF));
    CU(cudaMemsetAsync(m->cc, 0, sizeof(ChunkCtr) * s3d_map::NBUF, m->stream));
    for (int b = 0; b < s3d_map::NBUF; ++b) {
        m->buf[b].cc = m->cc + b;
        CU(cudaEventCreateWithFlags(&m->buf[b].expanded, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&m->buf[b].freed, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&m->buf[b].merged, cudaEventDisableTiming));
    }
    CU(cudaMalloc(&m->ex_counts, sizeof(u64) * 4));
    CU(cudaMallocHost(&m->ex_counts_host, sizeof(u64) * 4));
    k_reset_ctr<<<1, 1, 0, m->stream>>>(m->mc);
    k_init_life_ctr<<<1, 1, 0, m->stream>>>(m->mc);
    CU(cudaMalloc(&m->dbg_last_n, sizeof(u32)));
    CU(cudaMemsetAsync(m->dbg_last_n, 0, sizeof(u32), m->stream));
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
#ifdef S3D_AP_PHASES
    {
        unsigned long long ph[16];
        if (cudaMemcpyFromSymbol(ph, g_ap_phase, sizeof ph) == cudaSuccess && ph[15]) {
            unsigned long long tot = 0;
            for (int i = 0; i < 10; ++i) tot += ph[i];
            fprintf(stderr, "[s3d] k_apply_chunk phases over %llu blocks (%% of thread-0 time; cycles per block):", ph[15]);
            for (int i = 0; i < 10; ++i) fprintf(stderr, " p%d %.1f%% %llu;", i, 100.0 * (double)ph[i] / (double)tot, ph[i] / ph[15]);
            fprintf(stderr, "\n");
            memset(ph, 0, sizeof ph);
            cudaMemcpyToSymbol(g_ap_phase, ph, sizeof ph);
        }
    }
#endif
    if (m->table) cudaFree(m->table);
    if (m->life) cudaFree(m->life);
    if (m->dbg_last_n) cudaFree(m->dbg_last_n);
    m->dbg_last.release();
    if (m->mc) cudaFree(m->mc);
    if (m->mc_host) cudaFreeHost(m->mc_host);
    if (m->snap_host) cudaFreeHost(m->snap_host);
    for (int i = 0; i < s3d_map::RING; ++i) if (m->snap_ev[i]) cudaEventDestroy(m->snap_ev[i]);
    if (m->cc) cudaFree(m->cc);
    for (int b = 0; b < s3d_map::NBUF; ++b) {
        if (m->buf[b].expanded) cudaEventDestroy(m->buf[b].expanded);
        if (m->buf[b].freed) cudaEventDestroy(m->buf[b].freed);
        if (m->buf[b].merged) cudaEventDestroy(m->buf[b].merged);
    }
    if (m->mstream) { cudaStreamSynchronize(m->mstream); cudaStreamDestroy(m->mstream); }
    if (m->ctl_stream) { cudaStreamSynchronize(m->ctl_stream); cudaStreamDestroy(m->ctl_stream); }
    if (m->xstream) { cudaStreamSynchronize(m->xstream); cudaStreamDestroy(m->xstream); }
    if (m->xstream2) { cudaStreamSynchronize(m->xstream2); cudaStreamDestroy(m->xstream2); }
    if (m->x_ev) cudaEventDestroy(m->x_ev);
    if (m->snap_stream) { cudaStreamSynchronize(m->snap_stream); cudaStreamDestroy(m->snap_stream); }
    if (m->owner_host) cudaFreeHost(m->owner_host);
    m->send_buf.release(); m->owner_ctr.release();
    for (void *q : m->ipc_opened) cudaIpcCloseMemHandle(q);
    if (m->xblock) cudaFree(m->xblock);
    m->d_peers.release(); m->route_cursor.release(); m->trace.release();
    for (cudaEvent_t e : m->copy_ev) cudaEventDestroy(e);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    if (m->ex_counts) cudaFree(m->ex_counts);
    if (m->ex_counts_host) cudaFreeHost(m->ex_counts_host);
    if (m->stats_host) cudaFreeHost(m->stats_host);
    m->d_beam_col.release(); m->d_nv_free.release(); m->d_nv_occ.release();
    m->d_cos_b.release(); m->d_sin_b.release(); m->d_range.release(); m->d_cos_va.release(); m->d_sin_va.release();
    m->d_csva32.release();
    m->spool.release(); m->sum_tab.release(); m->stats.release();
    m->img_dev.release(); m->T_dev.release(); m->img16_dev.release();
    for (Staging &s : m->stg) {
        if (s.img) cudaFree(s.img);
        if (s.T) cudaFree(s.T);
        if (s.stats) cudaFree(s.stats);
        if (s.stats_host) cudaFreeHost(s.stats_host);
    }
    for (Job &j : m->jobs) if (j.done_ev) cudaEventDestroy(j.done_ev);
    for (cudaEvent_t e : m->job_ev_pool) cudaEventDestroy(e); m->io_keys.release(); m->io_vals.release(); m->io_flags.release();
    m->ex_xyz.release(); m->ex_prob.release(); m->ex_L.release(); m->ex_cls.release(); m->ex_ijk.release(); m->ex_f32.release();
    for (cudaEvent_t e : m->ev_pool) cudaEventDestroy(e);
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
    p.l_skip = (p.a_thr > 1e-3 && p.a_thr < 1.0 - 1e-3) ? std::log(p.a_thr / (1.0 - p.a_thr)) + 1e-6
                                                         : std::numeric_limits<double>::infinity();
    p.scale0 = (0.5 <= p.a_thr) ? (0.5 / p.a_thr) * p.a_ratio : 1.0;
    p.thr = std::max(-1, std::min(255, q->intensity_threshold));
    {
        // fast quantiser: fractional parts in (1e-6, 1 - 1e-6) are decided by the reciprocal product
        const double lo = 1e-6, hi = 1.0 - 1e-6;
        u64 blo, bhi;
        memcpy(&blo, &lo, 8); memcpy(&bhi, &hi, 8);
        p.fr_hi_lo = (u32)(blo >> 32) + 1u;
        p.fr_hi_span = (u32)(bhi >> 32) - p.fr_hi_lo;
    }
    {
        // running sums, one addition at a time, exactly as `sum += log_odds` accumulates them
        int rc = set_device(m); if (rc) return rc;
        if ((rc = pump(m, true))) return rc;
        double tab[4][SUMT];
        tab[0][0] = tab[1][0] = tab[2][0] = tab[3][0] = 0.0;
        for (int n = 1; n < SUMT; ++n) {
            tab[0][n] = tab[0][n - 1] + p.lo_free; tab[1][n] = tab[1][n - 1] + p.lo_occ;
            tab[2][n] = tab[0][n] / (double)n; tab[3][n] = tab[1][n] / (double)n;
        }
        // ... and the mean of n_free free deltas followed by n_occ occupied ones (:546, :559) for the small counts
        std::vector<double> all((size_t)4 * SUMT + MEAN_WORDS, 0.0);
        memcpy(all.data(), &tab[0][0], sizeof tab);
        for (int nf = 0; nf < SUMT; ++nf)
            for (int no = 0; no < SUMT; ++no) {
                if (nf >= AP_MF && no != 0) continue;
                if (nf + no == 0) continue;
                double sum = 0.0;
                for (int q = 0; q < nf; ++q) sum += p.lo_free;
                for (int q = 0; q < no; ++q) sum += p.lo_occ;
                const double mean = sum / (double)(nf + no);
                if (nf < AP_MF) all[(size_t)4 * SUMT + (size_t)nf * SUMT + no] = mean;
                if (no == 0) all[(size_t)4 * SUMT + (size_t)AP_MF * SUMT + nf] = mean;
            }
        if ((rc = upload(m->sum_tab, all.data(), all.size(), m->stream))) return rc;
        CU(cudaStreamSynchronize(m->stream));
    }
    m->have_params = true;
    m->ex_valid = false;
    update_count_bound(m);
    return 0;
}

int s3d_set_tables(s3d_map *m, const s3d_tables *t)
{
    if (!m || !t) return fail(S3D_EINVAL, "null argument");
    if (t->H < 0 || t->W < 0 || t->n_beams < 0 || t->n_beams > t->W || t->nv_max < 0)
        return fail(S3D_EINVAL, "bad table shape");
    if (t->free_step < 1 || t->occ_window < 0) return fail(S3D_EINVAL, "bad free_step/occ_window");
    if (t->H >= (1 << 16) || t->nv_max >= (1 << 15)) return fail(S3D_EINVAL, "H or nv_max too large");
    const size_t ex_smem = expand_smem_bytes(t->H, t->free_step, t->occ_window, 1);
    if (ex_smem > 200 * 1024) return fail(S3D_EINVAL, "image height %d needs %zu bytes of shared memory per block", t->H, ex_smem);
    int rc = set_device(m); if (rc) return rc;
    if ((rc = sync_counters(m))) return rc;
    const size_t nb = (size_t)t->n_beams, H = (size_t)t->H;
    const size_t nfan = (size_t)t->nv_max * ((size_t)t->nv_max + 2);
    for (size_t b = 0; b < nb; ++b) {
        const int col = t->beam_col[b];
        if (col < 0 || col >= t->W) return fail(S3D_EINVAL, "beam_col[%zu]=%d out of range", b, col);
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
    std::vector<float2> cs32(std::max<size_t>(nfan, 1));
    for (size_t i = 0; i < nfan; ++i) cs32[i] = make_float2((float)t->cos_va[i], (float)t->sin_va[i]);   // fp32 copy for the fast path
    if ((rc = upload(m->d_csva32, cs32.data(), cs32.size(), m->stream))) return rc;
    int col_step = nb > 1 ? t->beam_col[1] - t->beam_col[0] : 1;
    for (size_t b = 0; b < nb && col_step > 0; ++b) if (t->beam_col[b] != (int)b * col_step) col_step = 0;
    CU(cudaStreamSynchronize(m->stream));     // host vectors go out of scope
    DevTables &d = m->tab;
    d.H = t->H; d.W = t->W; d.n_beams = t->n_beams; d.nv_max = t->nv_max;
    d.free_step = t->free_step; d.occ_window = t->occ_window;
    d.beam_col = m->d_beam_col.p; d.cos_b = m->d_cos_b.p; d.sin_b = m->d_sin_b.p; d.range_m = m->d_range.p;
    d.nv_free = m->d_nv_free.p; d.nv_occ = m->d_nv_occ.p; d.cos_va = m->d_cos_va.p; d.sin_va = m->d_sin_va.p;
    d.csva32 = m->d_csva32.p; d.col_step = col_step > 0 ? col_step : 0; d.n_trig = (int)nfan;
    m->have_tables = true;
    {
        const int cap = 200 * 1024;
        CU(cudaFuncSetAttribute(k_expand<u32, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        CU(cudaFuncSetAttribute(k_expand<u32, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        CU(cudaFuncSetAttribute(k_expand<u64, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        CU(cudaFuncSetAttribute(k_expand<u32, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        CU(cudaFuncSetAttribute(k_expand<u32, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        CU(cudaFuncSetAttribute(k_expand<u64, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        CU(cudaFuncSetAttribute(k_expand<u32, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        CU(cudaFuncSetAttribute(k_expand<u32, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        CU(cudaFuncSetAttribute(k_expand<u64, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    }
    CU(cudaFuncSetAttribute(k_apply_chunk<u32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)apply_smem_bytes<u32>()));
    CU(cudaFuncSetAttribute(k_apply_chunk<u32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)apply_smem_bytes<u32>()));
    CU(cudaFuncSetAttribute(k_apply_chunk<u64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)apply_smem_bytes<u64>()));
    CU(cudaFuncSetAttribute(k_apply_chunk<u64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)apply_smem_bytes<u64>()));
    m->h_range.assign(t->range_m, t->range_m + H);
    m->h_nv_free.assign(t->nv_free, t->nv_free + H);
    m->h_nv_occ.assign(t->nv_occ, t->nv_occ + H);
    m->half_aperture = t->nv_max > 0 ? std::atan2(t->sin_va[nfan - 1], t->cos_va[nfan - 1]) : 0.0;   // va of v_step = +nv_max
    update_count_bound(m);
    // first guess for the chunk dedupe table; it doubles on demand (retry) from here
    // (a rank of a sharded map holds about 1/world of the entries; 1.5x margin, it doubles on demand)
    // (a rank of a routed map cannot re-run a chunk -- its peers have moved on -- so it starts from half the
    // worst case, every sample of 16 frames a voxel of its own, split over the ranks; from then on the table
    // doubles ahead of need from the counts of the chunks before, as on a single map)
    u64 want = std::min<u64>(1u << 20, std::max<u64>(1u << 14, m->samples_max / 4));
    if (m->shard_world > 1) want = std::min<u64>(1u << 22, std::max<u64>(1u << 16, m->samples_max * GF / (2 * (u64)m->shard_world)));
    return ensure_scratch(m, m->scratch_env ? m->scratch_env : want, false);
}

int s3d_ingest_batch_dev(s3d_map *m, const uint8_t *images_dev, int64_t n, const double *T_dev,
                         s3d_frame_stats *out, s3d_frame_stats *stats_dev)
{
    int rc = check_ready(m); if (rc) return rc;
    if (n < 0) return fail(S3D_EINVAL, "n < 0");
    if (n == 0) return 0;
    if ((rc = set_device(m))) return rc;
    DevStats *sd = reinterpret_cast<DevStats *>(stats_dev);
    if (!sd) {
        if ((rc = pump(m, true))) return rc;              // the internal stats buffer may be reallocated
        if ((rc = m->stats.ensure((size_t)n))) return rc;
        sd = m->stats.p;
    }
    if ((rc = submit_frames(m, images_dev, n, T_dev, sd))) return rc;
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
    if ((rc = pump(m, true))) return rc;                  // staging buffers are about to be reused
    const size_t img_bytes = (size_t)m->tab.H * m->tab.W;
    if ((rc = m->stats.ensure((size_t)n))) return rc;
    // The whole call is staged in device memory (up to 1 GiB at a time), so a chunk that has
    // to be re-run still finds its frames.  Copies go in sub-pieces on a second stream and
    // overlap the kernels of the previous sub-piece.
    const int64_t stage = std::max<int64_t>(GF, std::min<int64_t>(n, (int64_t)((1ull << 30) / std::max<size_t>(1, img_bytes))));
    // (the first piece is one chunk, so that the first kernel starts after 16 frames' worth of copy)
    auto piece = [](int64_t s0) -> int64_t { return s0 == 0 ? GF : 2 * GF; };
    if ((rc = m->img_dev.ensure(std::max<size_t>(16, img_bytes * (size_t)stage)))) return rc;
    if ((rc = m->T_dev.ensure(16 * (size_t)stage))) return rc;
    for (int64_t base = 0; base < n; base += stage) {
        const int64_t k = std::min<int64_t>(stage, n - base);
        if (base > 0 && (rc = pump(m, true))) return rc;
        CU(cudaMemcpyAsync(m->T_dev.p, T + base * 16, sizeof(double) * 16 * (size_t)k, cudaMemcpyHostToDevice, m->copy_stream));
        // all copies of this stage are queued first (they run back to back on the copy stream),
        // then the frames are submitted piece by piece, each behind the event of its own copy
        size_t ei = 0;
        for (int64_t s0 = 0; s0 < k; s0 += piece(s0), ++ei) {
            const int64_t kk = std::min<int64_t>(piece(s0), k - s0);
            if (ei == m->copy_ev.size()) {
                cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                m->copy_ev.push_back(e);
            }
            if (img_bytes)
                CU(cudaMemcpyAsync(m->img_dev.p + (size_t)s0 * img_bytes, images + (size_t)(base + s0) * img_bytes,
                                   img_bytes * (size_t)kk, cudaMemcpyHostToDevice, m->copy_stream));
            CU(cudaEventRecord(m->copy_ev[ei], m->copy_stream));
        }
        ei = 0;
        for (int64_t s0 = 0; s0 < k; s0 += piece(s0), ++ei) {
            const int64_t kk = std::min<int64_t>(piece(s0), k - s0);
            CU(cudaStreamWaitEvent(m->xstream, m->copy_ev[ei], 0));
            CU(cudaStreamWaitEvent(m->xstream2, m->copy_ev[ei], 0));
            if ((rc = submit_frames(m, m->img_dev.p + (size_t)s0 * img_bytes, kk, m->T_dev.p + s0 * 16, m->stats.p + base + s0))) return rc;
        }
    }
    return finish_stats(m, m->stats.p, n, out);
}

int s3d_ingest_submit(s3d_map *m, const uint8_t *images, int64_t n, const double *T, int *ticket)
{
    int rc = check_ready(m); if (rc) return rc;
    if (!images || !T || !ticket) return fail(S3D_EINVAL, "null argument");
    if (n < 1) return fail(S3D_EINVAL, "n < 1");
    if ((rc = set_device(m))) return rc;
    const size_t img_bytes = (size_t)m->tab.H * m->tab.W;
    if (img_bytes * (size_t)n > (1ull << 30)) return fail(S3D_EINVAL, "an asynchronous batch is limited to 1 GiB of frames; split it");
    Staging &s = m->stg[m->stg_next];
    if (s.busy) return fail(S3D_EINVAL, "both staging slots hold uncollected batches: call s3d_ingest_collect first");
    // (re)size this slot's buffers; nothing of this slot is in flight
    if (s.img_n < std::max<size_t>(16, img_bytes * (size_t)n)) {
        if (s.img) cudaFree(s.img);
        s.img = nullptr; s.img_n = 0;
        CU(cudaMalloc(&s.img, std::max<size_t>(16, img_bytes * (size_t)n))); s.img_n = std::max<size_t>(16, img_bytes * (size_t)n);
    }
    if (s.T_n < 16 * (size_t)n) {
        if (s.T) cudaFree(s.T);
        s.T = nullptr; s.T_n = 0;
        CU(cudaMalloc(&s.T, sizeof(double) * 16 * (size_t)n)); s.T_n = 16 * (size_t)n;
    }
    if (s.stats_n < (size_t)n) {
        if (s.stats) cudaFree(s.stats);
        if (s.stats_host) cudaFreeHost(s.stats_host);
        s.stats = nullptr; s.stats_host = nullptr; s.stats_n = 0;
        CU(cudaMalloc(&s.stats, sizeof(DevStats) * (size_t)n));
        CU(cudaMallocHost(&s.stats_host, sizeof(DevStats) * (size_t)n));
        s.stats_n = (size_t)n;
    }
    auto piece = [](int64_t s0) -> int64_t { return s0 == 0 ? GF : 2 * GF; };
    CU(cudaMemcpyAsync(s.T, T, sizeof(double) * 16 * (size_t)n, cudaMemcpyHostToDevice, m->copy_stream));
    size_t ei = 0;
    for (int64_t s0 = 0; s0 < n; s0 += piece(s0), ++ei) {
        const int64_t kk = std::min<int64_t>(piece(s0), n - s0);
        if (ei == m->copy_ev.size()) {
            cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            m->copy_ev.push_back(e);
        }
        if (img_bytes)
            CU(cudaMemcpyAsync(s.img + (size_t)s0 * img_bytes, images + (size_t)s0 * img_bytes, img_bytes * (size_t)kk,
                               cudaMemcpyHostToDevice, m->copy_stream));
        CU(cudaEventRecord(m->copy_ev[ei], m->copy_stream));
    }
    ei = 0;
    for (int64_t s0 = 0; s0 < n; s0 += piece(s0), ++ei) {
        const int64_t kk = std::min<int64_t>(piece(s0), n - s0);
        CU(cudaStreamWaitEvent(m->xstream, m->copy_ev[ei], 0));
        CU(cudaStreamWaitEvent(m->xstream2, m->copy_ev[ei], 0));
        if ((rc = submit_frames(m, s.img + (size_t)s0 * img_bytes, kk, s.T + s0 * 16, s.stats + s0, s0 + kk == n))) return rc;
    }
    s.last_job = m->job_seq; s.n = n; s.busy = true;
    *ticket = m->stg_next;
    m->stg_next ^= 1;
    return 0;
}

int s3d_ingest_collect(s3d_map *m, int ticket, s3d_frame_stats *out)
{
    if (!m) return fail(S3D_EINVAL, "null map");
    if (ticket < 0 || ticket > 1 || !m->stg[ticket].busy) return fail(S3D_EINVAL, "no batch is pending under ticket %d", ticket);
    int rc = set_device(m); if (rc) return rc;
    Staging &s = m->stg[ticket];
    if ((rc = wait_jobs_through(m, s.last_job))) { s.busy = false; return rc; }
    if (out) {
        CU(cudaMemcpyAsync(s.stats_host, s.stats, sizeof(DevStats) * (size_t)s.n, cudaMemcpyDeviceToHost, m->ctl_stream));
        CU(cudaStreamSynchronize(m->ctl_stream));
        for (int64_t f = 0; f < s.n; ++f) stats_to_abi(s.stats_host[f], out[f]);
    }
    s.busy = false;
    return 0;
}

int s3d_ingest_batch_mono16(s3d_map *m, const uint16_t *images, int64_t n, const double *T, s3d_frame_stats *out)
{
    int rc = check_ready(m); if (rc) return rc;
    if (n < 0) return fail(S3D_EINVAL, "n < 0");
    if (n == 0) return 0;
    if (!images || !T) return fail(S3D_EINVAL, "null input");
    if ((rc = set_device(m))) return rc;
    if ((rc = pump(m, true))) return rc;                  // staging buffers are about to be reused
    const size_t img_px = (size_t)m->tab.H * m->tab.W;
    if ((rc = m->stats.ensure((size_t)n))) return rc;
    // staged in pieces of up to 256 MiB of 16-bit pixels; the conversion runs on the copy stream
    // right behind its copy, and the frames are submitted behind the conversion
    const int64_t stage = std::max<int64_t>(GF, std::min<int64_t>(n, (int64_t)((1ull << 27) / std::max<size_t>(1, img_px))));
    if ((rc = m->img16_dev.ensure(std::max<size_t>(16, img_px * (size_t)stage)))) return rc;
    if ((rc = m->img_dev.ensure(std::max<size_t>(16, img_px * (size_t)stage)))) return rc;
    if ((rc = m->T_dev.ensure(16 * (size_t)stage))) return rc;
    if (m->copy_ev.empty()) { cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); m->copy_ev.push_back(e); }
    for (int64_t base = 0; base < n; base += stage) {
        const int64_t k = std::min<int64_t>(stage, n - base);
        if (base > 0 && (rc = pump(m, true))) return rc;
        CU(cudaMemcpyAsync(m->T_dev.p, T + base * 16, sizeof(double) * 16 * (size_t)k, cudaMemcpyHostToDevice, m->copy_stream));
        if (img_px) {
            CU(cudaMemcpyAsync(m->img16_dev.p, images + (size_t)base * img_px, sizeof(uint16_t) * img_px * (size_t)k,
                               cudaMemcpyHostToDevice, m->copy_stream));
            const size_t px = img_px * (size_t)k;
            const int blocks = (int)std::min<size_t>((px / 4 + 255) / 256 + 1, (size_t)m->n_sm * 8);
            k_mono16_to_u8<<<blocks, 256, 0, m->copy_stream>>>(m->img16_dev.p, m->img_dev.p, px);
            CU(cudaGetLastError());
            m->launches += 1;
        }
        CU(cudaEventRecord(m->copy_ev[0], m->copy_stream));
        CU(cudaStreamWaitEvent(m->xstream, m->copy_ev[0], 0));
        CU(cudaStreamWaitEvent(m->xstream2, m->copy_ev[0], 0));
        if ((rc = submit_frames(m, m->img_dev.p, k, m->T_dev.p, m->stats.p + base))) return rc;
    }
    return finish_stats(m, m->stats.p, n, out);
}

int s3d_ingest(s3d_map *m, const uint8_t *image, const double T[16], s3d_frame_stats *out)
{
    return s3d_ingest_batch(m, image, 1, T, out);
}

int s3d_shard_config(s3d_map *m, int rank, int world)
{
    if (!m) return fail(S3D_EINVAL, "null map");
    if (world < 1 || world > 64 || rank < 0 || rank >= world) return fail(S3D_EINVAL, "bad rank/world");
    int rc = set_device(m); if (rc) return rc;
    if ((rc = sync_counters(m))) return rc;
    m->shard_rank = rank; m->shard_world = world;
    if (!m->owner_host) CU(cudaMallocHost(&m->owner_host, sizeof(u32) * 64));
    return m->owner_ctr.ensure(3 * 64);
}

int s3d_shard_filter(s3d_map *m, int on)
{
    if (!m) return fail(S3D_EINVAL, "null map");
    int rc = set_device(m); if (rc) return rc;
    if ((rc = sync_counters(m))) return rc;
    m->shard_filter = on != 0;
    return 0;
}

int s3d_route_export(s3d_map *m, uint64_t records_per_pair, unsigned char *handle)
{
    if (!m || !handle) return fail(S3D_EINVAL, "null argument");
    if (!m->owner_host) return fail(S3D_EINVAL, "s3d_shard_config has not been called");
    if (records_per_pair < 1024) records_per_pair = 1024;
    int rc = set_device(m); if (rc) return rc;
    if ((rc = sync_counters(m))) return rc;
    if (m->route_on) return fail(S3D_EINVAL, "routing is enabled; disable it before exporting a new exchange block");
    if (m->xblock) { cudaFree(m->xblock); m->xblock = nullptr; }
    m->route_cap = records_per_pair;
    m->xblock_bytes = sizeof(RouteHdr) + sizeof(RouteRec) * ROUTE_DEPTH * (size_t)m->shard_world * (size_t)records_per_pair;
    cudaError_t e = cudaMalloc(&m->xblock, m->xblock_bytes);
    if (e != cudaSuccess) return fail(S3D_ENOMEM, "exchange block of %zu bytes: %s", m->xblock_bytes, cudaGetErrorString(e));
    CU(cudaMemset(m->xblock, 0, sizeof(RouteHdr)));
    m->route_seq = 0;
    memset(handle, 0, S3D_ROUTE_HANDLE_BYTES);
    memcpy(handle, &m->xblock, sizeof(void *));                       // same-process peers use the pointer itself
    cudaIpcMemHandle_t ipc;
    e = cudaIpcGetMemHandle(&ipc, m->xblock);
    if (e == cudaSuccess) memcpy(handle + 8, &ipc, sizeof ipc);
    else cudaGetLastError();                                          // no IPC on this platform: same-process use only
    static_assert(sizeof(cudaIpcMemHandle_t) + 8 <= S3D_ROUTE_HANDLE_BYTES, "handle size");
    return 0;
}

int s3d_route_attach(s3d_map *m, const unsigned char *handles, int same_process)
{
    if (!m || !handles) return fail(S3D_EINVAL, "null argument");
    if (!m->xblock) return fail(S3D_EINVAL, "s3d_route_export has not been called");
    int rc = set_device(m); if (rc) return rc;
    if ((rc = sync_counters(m))) return rc;
    for (void *q : m->ipc_opened) cudaIpcCloseMemHandle(q);
    m->ipc_opened.clear();
    const int world = m->shard_world;
    m->peer_ptr.assign((size_t)world, nullptr);
    for (int o = 0; o < world; ++o) {
        const unsigned char *h = handles + (size_t)o * S3D_ROUTE_HANDLE_BYTES;
        if (o == m->shard_rank) { m->peer_ptr[(size_t)o] = m->xblock; continue; }
        if (same_process) {
            unsigned char *ptr = nullptr;
            memcpy(&ptr, h, sizeof ptr);
            m->peer_ptr[(size_t)o] = ptr;
        } else {
            cudaIpcMemHandle_t ipc;
            memcpy(&ipc, h + 8, sizeof ipc);
            void *ptr = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return fail(S3D_ECUDA, "cudaIpcOpenMemHandle(rank %d): %s", o, cudaGetErrorString(e));
            m->ipc_opened.push_back(ptr);
            m->peer_ptr[(size_t)o] = static_cast<unsigned char *>(ptr);
        }
        if (!m->peer_ptr[(size_t)o]) return fail(S3D_EINVAL, "rank %d: empty exchange handle", o);
    }
    if ((rc = m->d_peers.ensure((size_t)world))) return rc;
    CU(cudaMemcpy(m->d_peers.p, m->peer_ptr.data(), sizeof(unsigned char *) * (size_t)world, cudaMemcpyHostToDevice));
    if ((rc = m->route_cursor.ensure(ROUTE_DEPTH * ROUTE_MAX_WORLD))) return rc;
    CU(cudaMemset(m->route_cursor.p, 0, sizeof(u32) * ROUTE_DEPTH * ROUTE_MAX_WORLD));
    { const char *e = getenv("S3D_ROUTE_TIMEOUT_MS"); if (e && atof(e) > 0) m->route_timeout_ns = (u64)(atof(e) * 1e6); }
    return 0;
}

int s3d_route_enable(s3d_map *m, int on)
{
    if (!m) return fail(S3D_EINVAL, "null map");
    int rc = set_device(m); if (rc) return rc;
    if ((rc = sync_counters(m))) return rc;
    if (on && m->shard_world > 1 && m->peer_ptr.size() != (size_t)m->shard_world)
        return fail(S3D_EINVAL, "s3d_route_attach has not been called");
    if (on && m->shard_filter) return fail(S3D_EINVAL, "routing and the replicated-expansion filter exclude each other");
    if (on && (rc = preload_pipeline_kernels())) return rc;
    if (on && !m->wide && !m->narrow_safe && m->have_tables) {
        // a count overflow would need a re-run that the peers cannot replay: start with wide lanes
        m->wide = true;
        if ((rc = ensure_scratch(m, m->scratch_cap, true, true))) return rc;
    }
    m->route_on = on != 0;
    return 0;
}

int s3d_shard_owner(const int32_t *ijk, int64_t n, int world, int32_t *owner)
{
    if (world < 1) return fail(S3D_EINVAL, "world < 1");
    for (int64_t i = 0; i < n; ++i)
        owner[i] = (int32_t)((mix64(pack_key(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2])) >> 40) % (u64)world);
    return 0;
}

int s3d_shard_expand(s3d_map *m, const uint8_t *images_dev, const double *T_dev, int g, s3d_frame_stats *stats_dev,
                     const void **records_dev, uint64_t *counts)
{
    int rc = check_ready(m); if (rc) return rc;
    if (g < 1 || g > GF) return fail(S3D_EINVAL, "a sharded chunk holds 1..%d frames", GF);
    if (!m->owner_host) return fail(S3D_EINVAL, "s3d_shard_config has not been called");
    if ((rc = set_device(m))) return rc;
    if ((rc = pump(m, true))) return rc;
    const DevTables &tab = m->tab;
    const u32 world = (u32)m->shard_world;
    // this rank's contiguous slice of the processed beams
    const int lo = (int)((int64_t)tab.n_beams * m->shard_rank / m->shard_world);
    const int hi = (int)((int64_t)tab.n_beams * (m->shard_rank + 1) / m->shard_world);
    const size_t img_stride = (size_t)tab.H * tab.W;
    DevStats *st = reinterpret_cast<DevStats *>(stats_dev);
    for (;;) {
        CU(cudaMemsetAsync(st, 0, sizeof(DevStats) * (size_t)g, m->stream));
        u32 n_unique = 0;
        if (tab.n_beams > 0 && tab.H > 0 && hi > lo) {
            ExpandArgs a;
            a.imgs = images_dev; a.img_stride = img_stride; a.T = T_dev;
            a.tab = tab; a.p = m->p;
            a.skeys = m->skeys; a.scnt = m->scnt; a.smask = (u32)(m->scratch_cap - 1);
            a.cc = m->cc; a.stats = st; a.mc = m->mc;
            a.seq = m->chunk_seq;                                // the owner gates growth, not the expander
            a.trace = nullptr; a.marks = nullptr;
            a.beam_lo = lo; a.beam_hi = hi;
            a.own_rank = 0; a.own_world = 1;
            a.rt = RouteCtx{1u, 0u, 0u, 0ull, nullptr, nullptr, 0ull, 0ull};
            launch_expand(m, a, hi - lo, g, m->stream);
        }
        // counts per owner -> host (the all-to-all needs the split sizes), and the retry flag
        CU(cudaMemsetAsync(m->owner_ctr.p, 0, sizeof(u32) * 3 * 64, m->stream));
        k_shard_count<<<m->n_sm * 2, 256, 0, m->stream>>>(m->skeys, (u32)m->scratch_cap, world, m->owner_ctr.p);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(m->owner_host, m->owner_ctr.p, sizeof(u32) * 64, cudaMemcpyDeviceToHost, m->stream));
        CU(cudaMemcpyAsync(m->mc_host, m->mc, sizeof(MapCtr), cudaMemcpyDeviceToHost, m->stream));
        CU(cudaStreamSynchronize(m->stream));
        ++m->chunk_seq; m->snap_floor = m->chunk_seq;
        if (m->mc_host->abort) {             // dedupe table too small, or a count overflowed: enlarge / widen, wipe, redo
            ++m->n_retries;
            const bool widen = (m->mc_host->abort & ABORT_NARROW) && !m->wide;
            if (widen) m->wide = true;
            if ((rc = ensure_scratch(m, (m->mc_host->abort & ABORT_SCRATCH) ? m->scratch_cap * 2 : m->scratch_cap, true, widen))) return rc;
            k_clear_abort<<<1, 1, 0, m->stream>>>(m->mc, m->cc);
            CU(cudaGetLastError());
            continue;
        }
        if ((rc = fatal_from_flags(m, m->mc_host->err))) return rc;
        u32 base[64];
        for (u32 o = 0; o < world; ++o) { base[o] = n_unique; n_unique += m->owner_host[o]; counts[o] = m->owner_host[o]; }
        if ((rc = m->send_buf.ensure((size_t)std::max<u32>(n_unique, 1) * REC_WORDS))) return rc;
        CU(cudaMemcpyAsync(m->owner_ctr.p + 64, base, sizeof(u32) * 64, cudaMemcpyHostToDevice, m->stream));
        if (m->wide)
            k_shard_pack<u64><<<m->n_sm * 2, 256, 0, m->stream>>>(m->skeys, static_cast<u64 *>(m->scnt), (u32)m->scratch_cap, world,
                                                                 m->owner_ctr.p + 64, m->owner_ctr.p + 128, m->send_buf.p);
        else
            k_shard_pack<u32><<<m->n_sm * 2, 256, 0, m->stream>>>(m->skeys, static_cast<u32 *>(m->scnt), (u32)m->scratch_cap, world,
                                                                 m->owner_ctr.p + 64, m->owner_ctr.p + 128, m->send_buf.p);
        k_shard_reset_cc<<<1, 1, 0, m->stream>>>(m->cc);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(m->stream));
        m->launches += 3;
        if (2ull * n_unique > m->scratch_cap && (rc = ensure_scratch(m, m->scratch_cap * 2, false))) return rc;
        *records_dev = m->send_buf.p;
        return 0;
    }
}

int s3d_shard_apply(s3d_map *m, const void *records_dev, uint64_t n_records, int g, s3d_frame_stats *stats_dev)
{
    int rc = check_ready(m); if (rc) return rc;
    if (g < 1 || g > GF) return fail(S3D_EINVAL, "a sharded chunk holds 1..%d frames", GF);
    if ((rc = set_device(m))) return rc;
    if ((rc = pump(m, true))) return rc;
    // every received record is at most one new voxel: exact room check on the host
    CU(cudaMemcpyAsync(m->mc_host, m->mc, sizeof(MapCtr), cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    m->count_known = m->mc_host->count;
    while (m->count_known + n_records > table_limit(m)) if ((rc = grow_table(m, m->cap * 2))) return rc;
    if (2 * n_records > m->scratch_cap && (rc = ensure_scratch(m, 2 * n_records, false))) return rc;
    DevStats *st = reinterpret_cast<DevStats *>(stats_dev);
    for (;;) {
        CU(cudaMemsetAsync(st, 0, sizeof(DevStats) * (size_t)g, m->stream));
        if (n_records) {
            const int blocks = (int)std::min<u64>((n_records + 255) / 256, (u64)m->n_sm * 8);
            const u64 *rec = reinterpret_cast<const u64 *>(records_dev);
            if (m->wide)
                k_shard_merge<u64><<<blocks, 256, 0, m->stream>>>(rec, n_records, m->skeys, static_cast<u64 *>(m->scnt),
                                                                 (u32)(m->scratch_cap - 1), m->cc, m->mc, m->chunk_seq);
            else
                k_shard_merge<u32><<<blocks, 256, 0, m->stream>>>(rec, n_records, m->skeys, static_cast<u32 *>(m->scnt),
                                                                 (u32)(m->scratch_cap - 1), m->cc, m->mc, m->chunk_seq);
            CU(cudaGetLastError());
            // a retryable flag here (dedupe table over-loaded, or a sample count beyond 16 bits) is
            // handled before anything touches the voxel table: the records are still in `recv`
            CU(cudaMemcpyAsync(m->mc_host, m->mc, sizeof(MapCtr), cudaMemcpyDeviceToHost, m->stream));
            CU(cudaStreamSynchronize(m->stream));
            if (m->mc_host->abort) {
                ++m->n_retries;
                const bool widen = (m->mc_host->abort & ABORT_NARROW) && !m->wide;
                if (widen) m->wide = true;
                if ((rc = ensure_scratch(m, (m->mc_host->abort & ABORT_SCRATCH) ? m->scratch_cap * 2 : m->scratch_cap, true, widen))) return rc;
                k_clear_abort<<<1, 1, 0, m->stream>>>(m->mc, m->cc);
                CU(cudaGetLastError());
                continue;
            }
        }
        break;
    }
    launch_apply(m, m->skeys, m->scnt, g, m->cc, st, m->stream);
    CU(cudaGetLastError());
    m->launches += 1;
    ++m->chunk_seq; m->snap_floor = m->chunk_seq;
    m->ex_valid = false;
    return sync_counters(m);
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
    if ((rc = pump(m, true))) return rc;                 // a queued chunk may be waiting for a retry
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
    if ((rc = sync_counters(m))) return rc;              // queued frames (and pending retries) belong to the map being cleared
    if ((rc = launch_fill_table(m, m->table, m->cap))) return rc;
    k_reset_ctr<<<1, 1, 0, m->stream>>>(m->mc);
    CU(cudaGetLastError());
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

int s3d_export_markers(s3d_map *m, double thr_occ, double thr_free, uint64_t counts[3])
{
    if (!m || !m->have_params || !counts) return fail(S3D_EINVAL, "map not ready (params) or null argument");
    uint64_t n_staged = 0;
    int rc = s3d_export_begin(m, thr_occ, thr_free, 0u, counts, &n_staged);      // counting pass: nothing staged
    if (rc) return rc;
    const u64 total = counts[0] + counts[1] + counts[2];
    m->ex_valid = false;
    m->mk_n = total;
    if (total == 0) return 0;
    if ((rc = m->ex_xyz.ensure(3 * (size_t)total))) return rc;
    u64 h[6] = {0, counts[S3D_CLASS_FREE], counts[S3D_CLASS_FREE] + counts[S3D_CLASS_UNKNOWN], 0, 0, 0};   // bases, cursors
    DevBuf<u64> d; if ((rc = d.ensure(6))) return rc;
    CU(cudaMemcpyAsync(d.p, h, sizeof h, cudaMemcpyHostToDevice, m->stream));
    const int blocks = (int)std::min<u64>((m->cap + 255) / 256, (u64)m->n_sm * 8);
    k_export_grouped<<<blocks, 256, 0, m->stream>>>(m->table, m->cap, m->p.res, thr_occ, thr_free, d.p, d.p + 3, m->ex_xyz.p);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(m->stream));
    d.release();
    return 0;
}

int s3d_export_read_markers(s3d_map *m, double *xyz, uint64_t n_total)
{
    if (!m || !xyz) return fail(S3D_EINVAL, "null argument");
    if (n_total != m->mk_n) return fail(S3D_EINVAL, "n_total does not match the staged markers (call s3d_export_markers; the map must not change in between)");
    if (n_total == 0) return 0;
    int rc = set_device(m); if (rc) return rc;
    CU(cudaMemcpyAsync(xyz, m->ex_xyz.p, sizeof(double) * 3 * n_total, cudaMemcpyDeviceToHost, m->stream));
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

int s3d_extend_bounds(s3d_map *m, const int32_t kmin[3], const int32_t kmax[3])
{
    if (!m || !kmin || !kmax) return fail(S3D_EINVAL, "null argument");
    int rc = set_device(m); if (rc) return rc;
    if ((rc = sync_counters(m))) return rc;
    MapCtr h = *m->mc_host;
    for (int q = 0; q < 3; ++q) { h.kmin[q] = std::min(h.kmin[q], (int)kmin[q]); h.kmax[q] = std::max(h.kmax[q], (int)kmax[q]); }
    CU(cudaMemcpyAsync(m->mc->kmin, h.kmin, sizeof h.kmin, cudaMemcpyHostToDevice, m->stream));
    CU(cudaMemcpyAsync(m->mc->kmax, h.kmax, sizeof h.kmax, cudaMemcpyHostToDevice, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    return 0;
}

int s3d_reset_bounds(s3d_map *m)
{
    if (!m) return fail(S3D_EINVAL, "null map");
    int rc = set_device(m); if (rc) return rc;
    if ((rc = sync_counters(m))) return rc;
    const int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    CU(cudaMemcpyAsync(m->mc->kmin, lo, sizeof lo, cudaMemcpyHostToDevice, m->stream));
    CU(cudaMemcpyAsync(m->mc->kmax, hi, sizeof hi, cudaMemcpyHostToDevice, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    return 0;
}

int s3d_debug_counters(s3d_map *m, int on)
{
    if (!m) return fail(S3D_EINVAL, "null map");
    int rc = set_device(m); if (rc) return rc;
    if ((rc = sync_counters(m))) return rc;
    if (on && m->route_on) return fail(S3D_EINVAL, "debug counters are not available on a routed map");
    if (on && !m->life) {
        cudaError_t e = cudaMalloc(&m->life, m->cap * sizeof(Slot));
        if (e != cudaSuccess) return fail(S3D_ENOMEM, "debug-counter table of %llu slots: %s", (unsigned long long)m->cap, cudaGetErrorString(e));
        if ((rc = launch_fill_table(m, m->life, m->cap))) return rc;
        CU(cudaStreamSynchronize(m->stream));
    }
    if (on && (rc = m->dbg_last.ensure(std::max<size_t>(m->scratch_cap, 1u << 12)))) return rc;
    m->debug_on = on != 0;
    return 0;
}

namespace {
int unpack_pairs(s3d_map *m, const ulonglong2 *dev, u64 k, int32_t *ijk, uint64_t *counts)
{
    std::vector<ulonglong2> h((size_t)k);
    CU(cudaMemcpyAsync(h.data(), dev, sizeof(ulonglong2) * (size_t)k, cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    for (u64 i = 0; i < k; ++i) {
        int a, b, c; unpack_key(h[(size_t)i].x, a, b, c);
        if (ijk) { ijk[3 * i] = a; ijk[3 * i + 1] = b; ijk[3 * i + 2] = c; }
        if (counts) counts[i] = h[(size_t)i].y;
    }
    return 0;
}
}

int s3d_debug_last_frame(s3d_map *m, int32_t *ijk, uint64_t *counts, uint64_t cap, uint64_t *n_out)
{
    if (!m || !n_out) return fail(S3D_EINVAL, "null argument");
    *n_out = 0;
    if (!m->debug_on) return 0;
    int rc = set_device(m); if (rc) return rc;
    if ((rc = sync_counters(m))) return rc;
    u32 n = 0;
    CU(cudaMemcpyAsync(&n, m->dbg_last_n, sizeof n, cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    n = (u32)std::min<u64>(n, m->dbg_last.n);
    *n_out = n;
    const u64 k = std::min<u64>(n, cap);
    return k ? unpack_pairs(m, m->dbg_last.p, k, ijk, counts) : 0;
}

int s3d_debug_totals(s3d_map *m, int32_t *ijk, uint64_t *counts, uint64_t cap, uint64_t *n_out)
{
    if (!m || !n_out) return fail(S3D_EINVAL, "null argument");
    *n_out = 0;
    if (!m->life) return 0;
    int rc = set_device(m); if (rc) return rc;
    if ((rc = sync_counters(m))) return rc;
    const u64 n = m->mc_host->life_count;
    *n_out = n;
    const u64 k = std::min<u64>(n, cap);
    if (!k) return 0;
    DevBuf<ulonglong2> outb; DevBuf<u64> cur;
    if ((rc = outb.ensure((size_t)k)) || (rc = cur.ensure(1))) { outb.release(); cur.release(); return rc; }
    cudaMemsetAsync(cur.p, 0, sizeof(u64), m->stream);
    const int blocks = (int)std::min<u64>((m->cap + 255) / 256, (u64)m->n_sm * 8);
    k_dump_life<<<blocks, 256, 0, m->stream>>>(m->life, m->cap, outb.p, k, cur.p);
    rc = unpack_pairs(m, outb.p, k, ijk, counts);
    outb.release(); cur.release();
    return rc;
}

int s3d_trace_read(s3d_map *m, uint64_t *out, uint64_t max_chunks, uint64_t *n_chunks)
{
    if (!m || !out || !n_chunks) return fail(S3D_EINVAL, "null argument");
    *n_chunks = 0;
    if (!m->trace.p) return 0;
    int rc = set_device(m); if (rc) return rc;
    if ((rc = sync_counters(m))) return rc;
    const u64 n = std::min<u64>({max_chunks, m->chunk_seq, s3d_map::TRACE_CHUNKS});
    CU(cudaMemcpy(out, m->trace.p, n * s3d_map::TRACE_W * sizeof(u64), cudaMemcpyDeviceToHost));
    *n_chunks = n;
    return 0;
}

int s3d_profile_enable(s3d_map *m, int on)
{
    if (!m) return fail(S3D_EINVAL, "null map");
    int rc = set_device(m); if (rc) return rc;
    if ((rc = prof_collect(m))) return rc;
    m->prof_on = on != 0;
    return 0;
}

int s3d_profile_read(s3d_map *m, s3d_profile *out)
{
    if (!m || !out) return fail(S3D_EINVAL, "null argument");
    int rc = set_device(m); if (rc) return rc;
    if ((rc = prof_collect(m))) return rc;
    CU(cudaStreamSynchronize(m->stream));
    m->prof.total_launches = m->launches;
    m->prof.retries = m->n_retries; m->prof.grows = m->n_grows;
    {
        MapCtr h;
        CU(cudaMemcpy(&h, m->mc, sizeof h, cudaMemcpyDeviceToHost));
        m->prof.route_records_sent = h.route_sent - m->route_sent_seen;
        m->route_sent_seen = h.route_sent;
        m->prof.voxel_probes = h.probes - m->probes_seen;
        m->probes_seen = h.probes;
    }
    *out = m->prof;
    m->prof = s3d_profile{};
    m->launches = 0; m->n_retries = 0; m->n_grows = 0;
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
