/*
 * sonar3d.h -- C-ABI of the B200-native sonar -> voxel log-odds hot path.
 *
 * The reference (luckkim123/sonar_3d_reconstruction) is pure Python and has no FFI of its
 * own: its operator API for this path is the Python class pair SonarTo3DMapper / SimpleOctree
 * in scripts/3d_mapper.py.  The entry points below are what a binding for exactly those
 * methods needs; each one names the reference lines it replaces.  The Python mirror of the
 * reference classes (sonar_3d_reconstruction_b200/mapper.py) is the only caller, via ctypes.
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: every function returns 0 on success and a negative S3D_E* code on failure;
 * s3d_last_error() gives the message of the calling thread's last failure.  All pointers are
 * plain host pointers unless the name says `_dev`.  Buffers are caller-owned and borrowed for
 * the duration of the call only.  Voxel keys cross the ABI as int32 triples (i, j, k), the
 * reference's tuple keys (scripts/3d_mapper.py:53-66); each axis must lie in
 * [-2^20, 2^20) (S3D_KEY_LIMIT) -- the table packs 3 x 21 bits into one 64-bit word.
 * One map = one CUDA device + one stream; calls on one map must come from one thread at a
 * time (the reference is single-threaded, scripts/3d_mapper_node.py:536).
 */
#ifndef SONAR3D_H
#define SONAR3D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S3D_ABI_VERSION 2
#define S3D_KEY_LIMIT (1 << 20)

enum {
    S3D_OK = 0,
    S3D_EINVAL = -1,      /* bad argument / tables or params not set */
    S3D_ECUDA = -2,       /* CUDA runtime error (message has the detail) */
    S3D_ENOMEM = -3,      /* device or host allocation failed */
    S3D_EKEYRANGE = -4,   /* a voxel key fell outside +-S3D_KEY_LIMIT (or a coordinate was NaN) */
    S3D_ETABLEFULL = -5,  /* the voxel table could not grow enough; map state is unspecified */
    S3D_ESCRATCH = -6,    /* per-frame dedupe scratch overflow (internal sizing bug) */
    S3D_EROUTE = -7       /* routed (multi-GPU) map: inbox overflow, silent peer, or a chunk that would need a re-run */
};

typedef struct s3d_map s3d_map; /* opaque */

/* Live tunables.  The reference reads these from public mutable attributes at update time
 * (SimpleOctree: scripts/3d_mapper.py:33-51; mapper: :259-270), so the host passes them
 * again whenever they change. */
typedef struct s3d_params {
    double resolution;          /* SimpleOctree.resolution, divisor of world_to_key (:63-65) */
    double log_odds_occupied;   /* :42  delta of an occupied sample (:481) */
    double log_odds_free;       /* :43  delta of a free sample (:446) */
    double log_odds_min;        /* :44  clamp (:110) */
    double log_odds_max;        /* :45 */
    double adaptive_threshold;  /* :50 */
    double adaptive_max_ratio;  /* :51 */
    double z_filter_min;        /* :269, test at :443/:478 */
    int32_t adaptive_update;    /* :49 */
    int32_t z_filter_enabled;   /* :270 */
    int32_t intensity_threshold;/* integer t with (pixel > t) == (pixel > self.intensity_threshold)
                                   for every uint8 pixel (:407, :452); -1 = every pixel passes,
                                   255 = none */
    int32_t reserved;
} s3d_params;

/* Host-built geometry tables for one image shape.  All trigonometry is evaluated on the host
 * with the reference's own expressions so that voxel keys are bit-exact;
 * the device only multiplies and adds. */
typedef struct s3d_tables {
    int32_t H, W;               /* range bins (rows), bearings (cols): polar_image.shape (:508) */
    int32_t n_beams;            /* processed beams: range(0, W, max(1, W//256)) that pass the FOV gate (:528-535) */
    int32_t nv_max;             /* largest fan half-width in nv_free/nv_occ */
    int32_t free_step;          /* 10  (:419) */
    int32_t occ_window;         /* 50  (:451) */
    const int32_t *beam_col;    /* [n_beams] image column of each processed beam */
    const double *cos_b;        /* [n_beams] cos(bearing_angles[col]) (:434) */
    const double *sin_b;        /* [n_beams] sin(bearing_angles[col]) (:435) */
    const double *range_m;      /* [H] r_idx * (max_range / H) (:404, :421) */
    const int32_t *nv_free;     /* [H] max(1, int(range*tan(ha)/(res*4))) (:427); 0 = bin skipped (range < min_range, :422) */
    const int32_t *nv_occ;      /* [H] max(2, int(range*tan(ha)/(res*1.5))) (:463); 0 = bin skipped (:456) */
    const double *cos_va;       /* [nv_max*(nv_max+2)] cos((v/max(1,nv))*ha); row nv starts at nv*nv-1, v=-nv..nv (:430,:466) */
    const double *sin_va;       /* same layout, sin */
} s3d_tables;

/* Per-frame counters = the integer entries of the reference's stats dict (:587-595). */
typedef struct s3d_frame_stats {
    int64_t num_occupied;  /* voxels updated with type 'occupied' this frame (:564) */
    int64_t num_free;      /* voxels updated with type 'free' (:567) */
    int64_t num_voxels;    /* len(octree.voxels) after the frame (:592) */
    int64_t num_samples;   /* samples emitted by all rays after the z filter (len of the :542 loop) */
    /* the reference's every-10th-frame [DEBUG] statistics (:575-585); filled only while
     * s3d_debug_counters is on, 0 otherwise */
    int64_t max_samples_per_voxel; /* max(frame_update_counts.values()) (:576) */
    int64_t num_voxels_gt10;       /* voxels with more than 10 samples this frame (:585) */
    int64_t max_total_samples;     /* max(voxel_update_counts.values()) after the frame (:578) */
    int64_t reserved;
} s3d_frame_stats;

/* ---- lifetime -------------------------------------------------------------------------- */

/* Create a map on CUDA device `device` with room for `initial_capacity` voxel slots (rounded
 * up to a power of two; 0 = default).  The table grows by rehashing when needed.
 * Replaces SimpleOctree.__init__ (:25-51). */
int s3d_create(int device, uint64_t initial_capacity, s3d_map **out);
int s3d_destroy(s3d_map *map);
const char *s3d_last_error(void);
int s3d_abi_version(void);

int s3d_set_params(s3d_map *map, const s3d_params *params);
int s3d_set_tables(s3d_map *map, const s3d_tables *tables);

/* ---- ingest: SonarTo3DMapper.process_sonar_image (:485-595) ----------------------------- */

/* One frame.  `image` is uint8[H*W] row-major (rows = range), `T` the row-major 4x4
 * T_sonar_to_world (:521).  Runs first-hit scan, free/occupied expansion, transform, key
 * quantisation, per-frame dedupe and the adaptive clamped log-odds update on the device and
 * returns the counters.  Synchronous. */
int s3d_ingest(s3d_map *map, const uint8_t *image, const double T[16], s3d_frame_stats *out);

/* `n` frames in sequence order (frame f is fully applied before frame f+1, as N successive
 * process_sonar_image calls would).  images: uint8[n*H*W]; T: double[n*16]; out: [n] or NULL.
 * Synchronous: returns when all frames are applied and `out` is filled. */
int s3d_ingest_batch(s3d_map *map, const uint8_t *images, int64_t n, const double *T,
                     s3d_frame_stats *out);

/* Asynchronous form of s3d_ingest_batch: s3d_ingest_submit queues the copies and kernels of a host
 * batch and returns at once with a ticket (0 or 1: two batches may be pending, so the upload and the
 * kernels of one overlap the tail of the one before); s3d_ingest_collect waits for that batch only
 * and returns its per-frame counters.  `images` and `T` must stay valid and unchanged until the batch
 * is collected.  Batches are applied in submission order; collect them in that order too. */
int s3d_ingest_submit(s3d_map *map, const uint8_t *images, int64_t n, const double *T, int *ticket);
int s3d_ingest_collect(s3d_map *map, int ticket, s3d_frame_stats *out);
/* 16-bit frames (ROS encodings mono16 / 16UC1): uint16[n][H][W] host images.  The node converts
 * them with `(img / 256).astype(uint8)` before it calls the mapper (scripts/3d_mapper_node.py:308-310);
 * here the raw 16-bit pixels are uploaded and the high byte is taken on the device, then the frames
 * go through the same path as s3d_ingest_batch. */
int s3d_ingest_batch_mono16(s3d_map *map, const uint16_t *images, int64_t n, const double *T, s3d_frame_stats *out);
/* Same with inputs already resident in device memory (bench `value` path; multi-GPU path).
 * Asynchronous on the map's stream when out == NULL; s3d_sync() or any synchronous call
 * orders after it.  `stats_dev`, if not NULL, receives n s3d_frame_stats in device memory. */
int s3d_ingest_batch_dev(s3d_map *map, const uint8_t *images_dev, int64_t n, const double *T_dev,
                         s3d_frame_stats *out, s3d_frame_stats *stats_dev);

/* Pre-size the table for `n_voxels` more voxels (load <= 1/2) so that it does not have to
 * rehash-grow later, e.g. inside a timed region. */
int s3d_reserve(s3d_map *map, uint64_t n_voxels);
int s3d_sync(s3d_map *map);
/* The map's cudaStream_t (for CUDA-event timing on the launching stream). */
void *s3d_stream(s3d_map *map);

/* ---- store: SimpleOctree (:83-125, :190-194) -------------------------------------------- */

/* update_voxel (:83-115) for n (key, delta, adaptive) triples, applied in array order (equal
 * keys are applied one after the other exactly as n successive calls would). */
int s3d_apply_updates(s3d_map *map, const int32_t *ijk, const double *delta,
                      const uint8_t *adaptive, int64_t n);
/* get_log_odds (:117-120): absent -> 0.0 and found = 0; never inserts. */
int s3d_query(s3d_map *map, const int32_t *ijk, int64_t n, double *log_odds, uint8_t *found);
/* len(voxels) (:592) */
int s3d_count(s3d_map *map, uint64_t *count);
/* voxels.items() (:147, :173): up to `cap` (key, log-odds) pairs in table order. */
int s3d_dump(s3d_map *map, int32_t *ijk, double *log_odds, uint64_t cap, uint64_t *n_out);
/* Bulk insert/overwrite of (key, log-odds) pairs: restore of a dump. */
int s3d_load(s3d_map *map, const int32_t *ijk, const double *log_odds, int64_t n);
/* clear (:190-194) */
int s3d_clear(s3d_map *map);
/* Integer bounding box of every key ever updated; min/max_bounds (:113-115) follow as
 * (k + 0.5) * resolution.  empty map: kmin > kmax. */
int s3d_bounds(s3d_map *map, int32_t kmin[3], int32_t kmax[3]);
uint64_t s3d_capacity(s3d_map *map);
/* Restore of a checkpoint (s3d_dump + s3d_bounds -> s3d_load + s3d_extend_bounds): widen the key
 * bounding box.  s3d_load and s3d_apply_updates themselves never touch it -- the reference extends
 * min/max_bounds only inside update_voxel, with the caller's point (:113-115), and not at all for
 * direct writes into `voxels`. */
int s3d_extend_bounds(s3d_map *map, const int32_t kmin[3], const int32_t kmax[3]);
/* Forget the box (the reference lets callers assign min_bounds / max_bounds, :38-40; the host
 * mirror folds the device's box into its own pair and restarts the device's). */
int s3d_reset_bounds(s3d_map *map);

/* ---- debug counters: voxel_update_counts / frame_update_counts (:307-308, :549-551, :575-585) ---- */

/* While on, the update kernel also keeps a lifetime sample count per voxel key (a second table
 * of the voxel table's size; it survives s3d_clear like voxel_update_counts survives reset_map,
 * :644-650) and fills the three debug fields of s3d_frame_stats.  Off by default. */
int s3d_debug_counters(s3d_map *map, int on);
/* frame_update_counts after the last ingested frame: (key, samples) pairs, any order. */
int s3d_debug_last_frame(s3d_map *map, int32_t *ijk, uint64_t *counts, uint64_t cap, uint64_t *n_out);
/* voxel_update_counts: (key, lifetime samples) pairs, any order. */
int s3d_debug_totals(s3d_map *map, int32_t *ijk, uint64_t *counts, uint64_t cap, uint64_t *n_out);

/* ---- export: get_occupied_voxels / get_all_voxels_classified (:127-188) ------------------ */

#define S3D_CLASS_FREE 0
#define S3D_CLASS_UNKNOWN 1
#define S3D_CLASS_OCCUPIED 2

/* Scan the table once on the device.  A voxel is FREE if L < thr_free (:177), else OCCUPIED
 * if L > thr_occ (:148, :179), else UNKNOWN.  Voxels whose class bit is set in `class_mask`
 * are compacted into a device staging buffer (centre = (k+0.5)*resolution (:78-80),
 * probability = 1/(1+exp(-L)) (:150, :175)).  counts[3] receives the per-class totals,
 * *n_out the number staged.  Pass thr_free = -inf for the occupied-only query. */
int s3d_export_begin(s3d_map *map, double thr_occ, double thr_free, uint32_t class_mask,
                     uint64_t counts[3], uint64_t *n_out);
/* Copy the staged result to host arrays (any may be NULL): xyz double[n*3], prob double[n],
 * cls int8[n], ijk int32[n*3].  n must equal *n_out of the last s3d_export_begin. */
int s3d_export_read(s3d_map *map, double *xyz, double *prob, int8_t *cls, int32_t *ijk, uint64_t n);
/* Same result as the PointCloud2 payload of the reference node: little-endian float32
 * x, y, z, intensity=probability, 16-byte stride (scripts/3d_mapper_node.py:419-443). */
int s3d_export_read_xyzi32(s3d_map *map, float *xyzi, uint64_t n);
/* The node's classified display (publish_marker_array, scripts/3d_mapper_node.py:448-527) fills one
 * CUBE_LIST marker per class with the voxel centres as geometry_msgs/Point (3 x float64).
 * s3d_export_markers classifies as s3d_export_begin does and stages all centres grouped by class --
 * [FREE | UNKNOWN | OCCUPIED], counts[c] points each -- so that each marker's `points` is one contiguous
 * block; s3d_export_read_markers copies the 3 * (counts[0] + counts[1] + counts[2]) doubles out
 * (SURVEY.md section 8f, row n1). */
int s3d_export_markers(s3d_map *map, double thr_occ, double thr_free, uint64_t counts[3]);
int s3d_export_read_markers(s3d_map *map, double *xyz, uint64_t n_total);

/* ---- sharded map (one process per GPU) ---------------------------------- */

/* The map shards by a hash of the voxel key: owner = (mix64(packed key) >> 40) % world.  Each
 * rank expands its slice of the processed beams of every frame, the per-(voxel, frame) integer
 * counts travel to the owning rank by all-to-all (the caller moves the bytes, e.g. with NCCL
 * through torch.distributed), and the owner merges and applies them.  Integer merges make the
 * N-rank result identical to the 1-rank result.  A record is 1 + S3D_CHUNK_FRAMES uint64: the packed key and
 * the (n_occ << 32 | n_free) counter of each of the frames of the chunk. */
#define S3D_CHUNK_FRAMES 16
#define S3D_RECORD_WORDS (1 + S3D_CHUNK_FRAMES)

int s3d_shard_config(s3d_map *map, int rank, int world);
/* Replicated expansion (the alternative to routing): with the filter on, s3d_ingest* on every rank
 * expands ALL beams of every frame but keeps only the samples whose voxel this rank owns.  No
 * exchange is needed -- each rank ends up with exactly the per-voxel counts of its shard -- at
 * the price of repeating the (cheap) expansion arithmetic on every rank.  Counters returned by
 * s3d_ingest* are then per-shard partials (sum over ranks = the reference's counters). */
int s3d_shard_filter(s3d_map *map, int on);
/* Routed map (the fused form of the exchange): every rank expands its slice of the processed
 * beams of every frame, and the expansion kernel itself writes the (voxel, frame, counts)
 * records of voxels owned by other ranks into the owners' inboxes over NVLink peer memory; the
 * owner merges them before it applies the chunk.  No host synchronisation and no separate
 * all-to-all per chunk: device-side sequence flags order sources and owners.  Set-up is
 * collective: every rank exports an exchange block (handle = S3D_ROUTE_HANDLE_BYTES opaque bytes,
 * to be all-gathered by the caller), attaches the handles of all ranks (rank order), then
 * enables routing; from then on s3d_ingest* on every rank must be called with the same frames.
 * The per-frame counters are per-shard partials (sum over ranks = the reference's counters).
 * A routed map cannot re-run a chunk: reserve table capacity up front (s3d_reserve); the
 * library grows the table ahead of need, and reports S3D_EROUTE if a chunk still hits the gate. */
#define S3D_ROUTE_HANDLE_BYTES 80
int s3d_route_export(s3d_map *map, uint64_t records_per_pair, unsigned char *handle);
/* handles: world x S3D_ROUTE_HANDLE_BYTES.  same_process != 0: the peers are maps of this process
 * on the same device (single-GPU tests); otherwise the blocks are opened with CUDA IPC. */
int s3d_route_attach(s3d_map *map, const unsigned char *handles, int same_process);
int s3d_route_enable(s3d_map *map, int on);
/* owner rank of each key (host helper; same function the device uses) */
int s3d_shard_owner(const int32_t *ijk, int64_t n, int world, int32_t *owner);
/* Expand g <= 16 frames (device-resident images / transforms) on this rank's beam slice and
 * pack the dedupe entries by owner.  *records_dev points at world runs of records laid out
 * back to back (valid until the next shard call); counts[o] is the length of run o.
 * stats_dev[g] receives this rank's num_samples.  Synchronous. */
int s3d_shard_expand(s3d_map *map, const uint8_t *images_dev, const double *T_dev, int g,
                     s3d_frame_stats *stats_dev, const void **records_dev, uint64_t *counts);
/* Merge n_records received records (device memory) and apply the chunk's g frames in order to
 * this rank's shard; stats_dev[g] receives the shard's num_occupied / num_free / num_voxels
 * (sum over ranks = the reference's counters).  Synchronous. */
int s3d_shard_apply(s3d_map *map, const void *records_dev, uint64_t n_records, int g,
                    s3d_frame_stats *stats_dev);

/* ---- measurement (bench.py) ----------------------------------------------------------------- */

#define S3D_K_FIRST_HIT 0
#define S3D_K_EXPAND 1
#define S3D_K_APPLY 2
#define S3D_K_COUNT 3

/* Per-kernel device time measured with CUDA events on the map's stream, and launch counts.
 * Enabling costs two event records per kernel group; the numbers accumulate until read. */
typedef struct s3d_profile {
    double ms[S3D_K_COUNT];          /* summed device time of each kernel class */
    uint64_t launches[S3D_K_COUNT];  /* launches of each kernel class */
    uint64_t frames;                 /* frames ingested while profiling was on */
    uint64_t total_launches;         /* pipeline kernels launched since the last read */
    uint64_t retries;                /* chunks re-run after the device gate asked for more room */
    uint64_t grows;                  /* voxel-table rehash-grows */
    uint64_t route_records_sent;     /* routed map: 16-byte records written into peers' inboxes over NVLink since the last read */
    uint64_t voxel_probes;           /* voxel-table find-or-insert probes since the last read: one per voxel per applied chunk */
} s3d_profile;

int s3d_profile_enable(s3d_map *map, int on);
/* Development aid (maps created with S3D_TRACE=1 in the environment): GPU timestamps
 * (%globaltimer, ns) of the pipeline kernels of the first chunks, 10 uint64 per chunk =
 * {start, end} x {ack wait, expand, flag wait, merge, apply}; untouched slots read ~0 / 0. */
int s3d_trace_read(s3d_map *map, uint64_t *out, uint64_t max_chunks, uint64_t *n_chunks);
/* Synchronises the stream, fills `out`, resets the accumulators. */
int s3d_profile_read(s3d_map *map, s3d_profile *out);

#ifdef __cplusplus
}
#endif
#endif /* SONAR3D_H */
