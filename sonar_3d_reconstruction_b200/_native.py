"""ctypes binding of libsonar3d.so (include/sonar3d.h).  No CPU fallback: if the CUDA library
is missing or no GPU is visible, construction fails loudly."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("S3D_LIB_PATH") or os.path.join(_HERE, "csrc", "libsonar3d.so")   # (override: kernel-variant experiments)

# every symbol include/sonar3d.h declares (tests check the list against the header and the .so)
SYMBOLS = (
    "s3d_create", "s3d_destroy", "s3d_last_error", "s3d_abi_version", "s3d_set_params", "s3d_set_tables",
    "s3d_ingest", "s3d_ingest_batch", "s3d_ingest_batch_dev", "s3d_reserve", "s3d_sync", "s3d_stream",
    "s3d_apply_updates", "s3d_query", "s3d_count", "s3d_dump", "s3d_load", "s3d_clear", "s3d_bounds",
    "s3d_capacity", "s3d_export_begin", "s3d_export_read", "s3d_export_read_xyzi32",
    "s3d_profile_enable", "s3d_profile_read",
    "s3d_shard_config", "s3d_shard_filter", "s3d_shard_owner", "s3d_shard_expand", "s3d_shard_apply",
    "s3d_route_export", "s3d_route_attach", "s3d_route_enable", "s3d_trace_read", "s3d_ingest_batch_mono16",
    "s3d_ingest_submit", "s3d_ingest_collect",
    "s3d_export_markers", "s3d_export_read_markers", "s3d_extend_bounds", "s3d_reset_bounds", "s3d_debug_counters", "s3d_debug_last_frame", "s3d_debug_totals",
)


class NativeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsonar3d error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("resolution", C.c_double), ("log_odds_occupied", C.c_double), ("log_odds_free", C.c_double),
                ("log_odds_min", C.c_double), ("log_odds_max", C.c_double), ("adaptive_threshold", C.c_double),
                ("adaptive_max_ratio", C.c_double), ("z_filter_min", C.c_double), ("adaptive_update", C.c_int32),
                ("z_filter_enabled", C.c_int32), ("intensity_threshold", C.c_int32), ("reserved", C.c_int32)]


class Tables(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("n_beams", C.c_int32), ("nv_max", C.c_int32),
                ("free_step", C.c_int32), ("occ_window", C.c_int32),
                ("beam_col", C.POINTER(C.c_int32)), ("cos_b", C.POINTER(C.c_double)), ("sin_b", C.POINTER(C.c_double)),
                ("range_m", C.POINTER(C.c_double)), ("nv_free", C.POINTER(C.c_int32)), ("nv_occ", C.POINTER(C.c_int32)),
                ("cos_va", C.POINTER(C.c_double)), ("sin_va", C.POINTER(C.c_double))]


class FrameStats(C.Structure):
    _fields_ = [("num_occupied", C.c_int64), ("num_free", C.c_int64), ("num_voxels", C.c_int64),
                ("num_samples", C.c_int64), ("max_samples_per_voxel", C.c_int64), ("num_voxels_gt10", C.c_int64),
                ("max_total_samples", C.c_int64), ("reserved", C.c_int64)]


STATS_WORDS = 8          # int64 words per s3d_frame_stats
STATS_BYTES = 8 * STATS_WORDS


class Profile(C.Structure):
    _fields_ = [("ms", C.c_double * 3), ("launches", C.c_uint64 * 3), ("frames", C.c_uint64),
                ("total_launches", C.c_uint64), ("retries", C.c_uint64), ("grows", C.c_uint64),
                ("route_records_sent", C.c_uint64), ("voxel_probes", C.c_uint64)]


KERNEL_NAMES = ("k_first_hit", "k_expand", "k_apply")

STATS_DTYPE = np.dtype([("num_occupied", "<i8"), ("num_free", "<i8"), ("num_voxels", "<i8"), ("num_samples", "<i8"),
                        ("max_samples_per_voxel", "<i8"), ("num_voxels_gt10", "<i8"), ("max_total_samples", "<i8"),
                        ("reserved", "<i8")])

_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raise (never fall back) if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  sonar_3d_reconstruction_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u8p, dp, i32p, i8p, u64p = (C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_int8), C.POINTER(C.c_uint64))
    sp = C.POINTER(FrameStats)
    L.s3d_last_error.restype = C.c_char_p
    L.s3d_abi_version.restype = C.c_int
    L.s3d_create.argtypes = [C.c_int, C.c_uint64, C.POINTER(vp)]
    L.s3d_destroy.argtypes = [vp]
    L.s3d_set_params.argtypes = [vp, C.POINTER(Params)]
    L.s3d_set_tables.argtypes = [vp, C.POINTER(Tables)]
    L.s3d_ingest.argtypes = [vp, vp, dp, sp]
    L.s3d_ingest_batch.argtypes = [vp, vp, C.c_int64, dp, vp]
    L.s3d_ingest_batch_dev.argtypes = [vp, vp, C.c_int64, vp, vp, vp]
    L.s3d_reserve.argtypes = [vp, C.c_uint64]
    L.s3d_sync.argtypes = [vp]
    L.s3d_stream.argtypes = [vp]
    L.s3d_stream.restype = vp
    L.s3d_apply_updates.argtypes = [vp, i32p, dp, u8p, C.c_int64]
    L.s3d_query.argtypes = [vp, i32p, C.c_int64, dp, u8p]
    L.s3d_count.argtypes = [vp, u64p]
    L.s3d_dump.argtypes = [vp, i32p, dp, C.c_uint64, u64p]
    L.s3d_load.argtypes = [vp, i32p, dp, C.c_int64]
    L.s3d_clear.argtypes = [vp]
    L.s3d_bounds.argtypes = [vp, i32p, i32p]
    L.s3d_capacity.argtypes = [vp]
    L.s3d_capacity.restype = C.c_uint64
    L.s3d_export_begin.argtypes = [vp, C.c_double, C.c_double, C.c_uint32, u64p, u64p]
    L.s3d_export_read.argtypes = [vp, dp, dp, i8p, i32p, C.c_uint64]
    L.s3d_export_read_xyzi32.argtypes = [vp, C.POINTER(C.c_float), C.c_uint64]
    L.s3d_profile_enable.argtypes = [vp, C.c_int]
    L.s3d_profile_read.argtypes = [vp, C.POINTER(Profile)]
    L.s3d_shard_config.argtypes = [vp, C.c_int, C.c_int]
    L.s3d_shard_filter.argtypes = [vp, C.c_int]
    L.s3d_route_export.argtypes = [vp, C.c_uint64, C.c_char_p]
    L.s3d_route_attach.argtypes = [vp, C.c_char_p, C.c_int]
    L.s3d_route_enable.argtypes = [vp, C.c_int]
    L.s3d_ingest_submit.argtypes = [vp, vp, C.c_int64, dp, C.POINTER(C.c_int)]
    L.s3d_ingest_collect.argtypes = [vp, C.c_int, vp]
    L.s3d_ingest_batch_mono16.argtypes = [vp, C.POINTER(C.c_uint16), C.c_int64, dp, vp]
    L.s3d_trace_read.argtypes = [vp, u64p, C.c_uint64, u64p]
    L.s3d_shard_owner.argtypes = [i32p, C.c_int64, C.c_int, i32p]
    L.s3d_shard_expand.argtypes = [vp, vp, vp, C.c_int, vp, C.POINTER(vp), u64p]
    L.s3d_shard_apply.argtypes = [vp, vp, C.c_uint64, C.c_int, vp]
    L.s3d_export_markers.argtypes = [vp, C.c_double, C.c_double, u64p]
    L.s3d_export_read_markers.argtypes = [vp, dp, C.c_uint64]
    L.s3d_extend_bounds.argtypes = [vp, i32p, i32p]
    L.s3d_reset_bounds.argtypes = [vp]
    L.s3d_debug_counters.argtypes = [vp, C.c_int]
    L.s3d_debug_last_frame.argtypes = [vp, i32p, u64p, C.c_uint64, u64p]
    L.s3d_debug_totals.argtypes = [vp, i32p, u64p, C.c_uint64, u64p]
    for name in SYMBOLS:
        getattr(L, name)          # AttributeError here = header / library mismatch
    _lib = L
    return L


def shard_owner(ijk: np.ndarray, world: int) -> np.ndarray:
    """Owner rank of each voxel key, evaluated by the library's host helper."""
    ijk = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
    out = np.empty(len(ijk), dtype=np.int32)
    _check(load_library().s3d_shard_owner(_ptr(ijk, C.c_int32), len(ijk), int(world), _ptr(out, C.c_int32)))
    return out


def _check(rc: int):
    if rc != 0:
        raise NativeError(rc, load_library().s3d_last_error().decode("utf-8", "replace"))


def _ptr(a: Optional[np.ndarray], typ):
    return None if a is None else a.ctypes.data_as(C.POINTER(typ))


class NativeMap:
    """One GPU-resident voxel hash table + the per-frame kernels (opaque s3d_map handle)."""

    CLASS_FREE, CLASS_UNKNOWN, CLASS_OCCUPIED = 0, 1, 2

    def __init__(self, device: int = 0, capacity: int = 0):
        self._lib = load_library()
        h = C.c_void_p()
        _check(self._lib.s3d_create(int(device), int(capacity), C.byref(h)))
        self._h = h
        self.device = int(device)
        self._keep = None       # keeps the table arrays alive for the duration of set_tables

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.s3d_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration ------------------------------------------------------------------
    def set_params(self, p: Params):
        _check(self._lib.s3d_set_params(self._h, C.byref(p)))

    def set_tables(self, t: "HostTables"):
        s = Tables()
        s.H, s.W, s.n_beams, s.nv_max = t.H, t.W, len(t.beam_col), t.nv_max
        s.free_step, s.occ_window = t.free_step, t.occ_window
        s.beam_col = _ptr(t.beam_col, C.c_int32)
        s.cos_b, s.sin_b = _ptr(t.cos_b, C.c_double), _ptr(t.sin_b, C.c_double)
        s.range_m = _ptr(t.range_m, C.c_double)
        s.nv_free, s.nv_occ = _ptr(t.nv_free, C.c_int32), _ptr(t.nv_occ, C.c_int32)
        s.cos_va, s.sin_va = _ptr(t.cos_va, C.c_double), _ptr(t.sin_va, C.c_double)
        _check(self._lib.s3d_set_tables(self._h, C.byref(s)))

    # -- ingest -------------------------------------------------------------------------
    def ingest(self, image_u8: np.ndarray, T: np.ndarray) -> FrameStats:
        st = FrameStats()
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        _check(self._lib.s3d_ingest(self._h, image_u8.ctypes.data, _ptr(T, C.c_double), C.byref(st)))
        return st

    def ingest_batch(self, images_u8: np.ndarray, T: np.ndarray) -> np.ndarray:
        n = int(images_u8.shape[0])
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(n, 16)
        out = np.zeros(n, dtype=STATS_DTYPE)
        _check(self._lib.s3d_ingest_batch(self._h, images_u8.ctypes.data, n, _ptr(T, C.c_double), out.ctypes.data))
        return out

    def ingest_submit(self, images_u8: np.ndarray, T: np.ndarray) -> int:
        """Queue a host batch without waiting; the arrays must stay alive and unchanged until ingest_collect."""
        n = int(images_u8.shape[0])
        ticket = C.c_int(-1)
        _check(self._lib.s3d_ingest_submit(self._h, images_u8.ctypes.data, n, _ptr(T, C.c_double), C.byref(ticket)))
        return ticket.value

    def ingest_collect(self, ticket: int, n: int) -> np.ndarray:
        out = np.zeros(n, dtype=STATS_DTYPE)
        _check(self._lib.s3d_ingest_collect(self._h, int(ticket), out.ctypes.data))
        return out

    def ingest_batch_mono16(self, images_u16: np.ndarray, T: np.ndarray) -> np.ndarray:
        """uint16[n, H, W] frames; the device keeps the high byte of every pixel (the node's img / 256)."""
        n = int(images_u16.shape[0])
        images_u16 = np.ascontiguousarray(images_u16, dtype=np.uint16)
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(n, 16)
        out = np.zeros(n, dtype=STATS_DTYPE)
        _check(self._lib.s3d_ingest_batch_mono16(self._h, images_u16.ctypes.data_as(C.POINTER(C.c_uint16)), n,
                                                 _ptr(T, C.c_double), out.ctypes.data))
        return out

    def ingest_batch_dev(self, images_ptr: int, n: int, T_ptr: int, want_stats: bool = True,
                         stats_dev_ptr: int = 0) -> Optional[np.ndarray]:
        out = np.zeros(n, dtype=STATS_DTYPE) if want_stats else None
        _check(self._lib.s3d_ingest_batch_dev(self._h, images_ptr, int(n), T_ptr,
                                              out.ctypes.data if want_stats else None, stats_dev_ptr or None))
        return out

    def reserve(self, n_voxels: int):
        _check(self._lib.s3d_reserve(self._h, int(n_voxels)))

    def sync(self):
        _check(self._lib.s3d_sync(self._h))

    @property
    def stream(self) -> int:
        return int(self._lib.s3d_stream(self._h) or 0)

    @property
    def capacity(self) -> int:
        return int(self._lib.s3d_capacity(self._h))

    # -- sharded map --------------------------------------------------------------------
    CHUNK_FRAMES = 16                   # include/sonar3d.h S3D_CHUNK_FRAMES (a build-time constant of the library: 16 or 32)
    RECORD_WORDS = 1 + CHUNK_FRAMES     # S3D_RECORD_WORDS

    def shard_config(self, rank: int, world: int):
        _check(self._lib.s3d_shard_config(self._h, int(rank), int(world)))
        self._world = int(world)

    def shard_filter(self, on: bool):
        _check(self._lib.s3d_shard_filter(self._h, int(bool(on))))

    ROUTE_HANDLE_BYTES = 80

    def route_export(self, records_per_pair: int) -> bytes:
        """Allocate this rank's exchange block (flags + inboxes); returns the opaque handle to all-gather."""
        buf = C.create_string_buffer(self.ROUTE_HANDLE_BYTES)
        _check(self._lib.s3d_route_export(self._h, int(records_per_pair), buf))
        return buf.raw

    def route_attach(self, handles: bytes, same_process: bool = False):
        """handles: the world's handles back to back, in rank order."""
        _check(self._lib.s3d_route_attach(self._h, bytes(handles), int(bool(same_process))))

    def trace_read(self, max_chunks: int = 4096) -> np.ndarray:
        """S3D_TRACE=1: uint64[n_chunks, 5, 2] = {start, end} ns of ack wait, expand, flag wait, merge, apply."""
        out = np.zeros(max_chunks * 10, dtype=np.uint64)
        n = C.c_uint64(0)
        _check(self._lib.s3d_trace_read(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), int(max_chunks), C.byref(n)))
        return out[: n.value * 10].reshape(-1, 5, 2)

    def route_enable(self, on: bool = True):
        _check(self._lib.s3d_route_enable(self._h, int(bool(on))))

    def shard_expand(self, images_ptr: int, T_ptr: int, g: int, stats_dev_ptr: int):
        """-> (device pointer of the packed records, [count per owner])"""
        rec = C.c_void_p()
        counts = (C.c_uint64 * 64)()
        _check(self._lib.s3d_shard_expand(self._h, images_ptr, T_ptr, int(g), stats_dev_ptr, C.byref(rec), counts))
        return int(rec.value or 0), [int(counts[o]) for o in range(self._world)]

    def shard_apply(self, records_ptr: int, n_records: int, g: int, stats_dev_ptr: int):
        _check(self._lib.s3d_shard_apply(self._h, records_ptr or None, int(n_records), int(g), stats_dev_ptr))

    # -- measurement --------------------------------------------------------------------
    def profile_enable(self, on: bool = True):
        _check(self._lib.s3d_profile_enable(self._h, int(bool(on))))

    def profile_read(self) -> dict:
        p = Profile()
        _check(self._lib.s3d_profile_read(self._h, C.byref(p)))
        return {"ms": {k: p.ms[i] for i, k in enumerate(KERNEL_NAMES)},
                "launches": {k: int(p.launches[i]) for i, k in enumerate(KERNEL_NAMES)},
                "frames": int(p.frames), "total_launches": int(p.total_launches),
                "retries": int(p.retries), "grows": int(p.grows), "route_records_sent": int(p.route_records_sent),
                "voxel_probes": int(p.voxel_probes)}

    # -- store --------------------------------------------------------------------------
    def apply_updates(self, ijk: np.ndarray, delta: np.ndarray, adaptive: np.ndarray):
        ijk = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
        delta = np.ascontiguousarray(delta, dtype=np.float64).reshape(-1)
        adaptive = np.ascontiguousarray(adaptive, dtype=np.uint8).reshape(-1)
        _check(self._lib.s3d_apply_updates(self._h, _ptr(ijk, C.c_int32), _ptr(delta, C.c_double),
                                           _ptr(adaptive, C.c_uint8), len(ijk)))

    def query(self, ijk: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        ijk = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
        L = np.zeros(len(ijk), dtype=np.float64)
        found = np.zeros(len(ijk), dtype=np.uint8)
        _check(self._lib.s3d_query(self._h, _ptr(ijk, C.c_int32), len(ijk), _ptr(L, C.c_double), _ptr(found, C.c_uint8)))
        return L, found.astype(bool)

    def count(self) -> int:
        n = C.c_uint64()
        _check(self._lib.s3d_count(self._h, C.byref(n)))
        return int(n.value)

    def dump(self) -> Tuple[np.ndarray, np.ndarray]:
        cap = self.count()
        ijk = np.empty((cap, 3), dtype=np.int32)
        L = np.empty(cap, dtype=np.float64)
        n = C.c_uint64()
        _check(self._lib.s3d_dump(self._h, _ptr(ijk, C.c_int32), _ptr(L, C.c_double), cap, C.byref(n)))
        k = min(cap, int(n.value))
        return ijk[:k], L[:k]

    def load(self, ijk: np.ndarray, log_odds: np.ndarray):
        ijk = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
        log_odds = np.ascontiguousarray(log_odds, dtype=np.float64).reshape(-1)
        _check(self._lib.s3d_load(self._h, _ptr(ijk, C.c_int32), _ptr(log_odds, C.c_double), len(ijk)))

    def clear(self):
        _check(self._lib.s3d_clear(self._h))

    def extend_bounds(self, kmin: np.ndarray, kmax: np.ndarray):
        kmin = np.ascontiguousarray(kmin, dtype=np.int32).reshape(3)
        kmax = np.ascontiguousarray(kmax, dtype=np.int32).reshape(3)
        _check(self._lib.s3d_extend_bounds(self._h, _ptr(kmin, C.c_int32), _ptr(kmax, C.c_int32)))

    def reset_bounds(self):
        _check(self._lib.s3d_reset_bounds(self._h))

    # -- debug counters -------------------------------------------------------------------
    def debug_counters(self, on: bool = True):
        _check(self._lib.s3d_debug_counters(self._h, int(bool(on))))

    def _debug_pairs(self, fn) -> Tuple[np.ndarray, np.ndarray]:
        n = C.c_uint64()
        _check(fn(self._h, None, None, 0, C.byref(n)))
        cap = int(n.value)
        ijk = np.empty((cap, 3), dtype=np.int32)
        cnt = np.empty(cap, dtype=np.uint64)
        if cap:
            _check(fn(self._h, _ptr(ijk, C.c_int32), _ptr(cnt, C.c_uint64), cap, C.byref(n)))
        k = min(cap, int(n.value))
        return ijk[:k], cnt[:k]

    def debug_last_frame(self) -> Tuple[np.ndarray, np.ndarray]:
        """(keys int32[n,3], samples uint64[n]) of the last ingested frame (frame_update_counts)."""
        return self._debug_pairs(self._lib.s3d_debug_last_frame)

    def debug_totals(self) -> Tuple[np.ndarray, np.ndarray]:
        """(keys int32[n,3], lifetime samples uint64[n]) (voxel_update_counts)."""
        return self._debug_pairs(self._lib.s3d_debug_totals)

    def bounds(self) -> Tuple[np.ndarray, np.ndarray]:
        kmin = np.zeros(3, dtype=np.int32)
        kmax = np.zeros(3, dtype=np.int32)
        _check(self._lib.s3d_bounds(self._h, _ptr(kmin, C.c_int32), _ptr(kmax, C.c_int32)))
        return kmin, kmax

    def export_markers(self, thr_occ: float, thr_free: float):
        """Voxel centres grouped by class: {class id: float64[n_c, 3]} (views of one buffer)."""
        counts = (C.c_uint64 * 3)()
        _check(self._lib.s3d_export_markers(self._h, float(thr_occ), float(thr_free), counts))
        n = [int(c) for c in counts]
        buf = np.empty((sum(n), 3), dtype=np.float64)
        if sum(n):
            _check(self._lib.s3d_export_read_markers(self._h, _ptr(buf, C.c_double), sum(n)))
        a, b = n[0], n[0] + n[1]
        return {self.CLASS_FREE: buf[:a], self.CLASS_UNKNOWN: buf[a:b], self.CLASS_OCCUPIED: buf[b:]}

    # -- export -------------------------------------------------------------------------
    def export(self, thr_occ: float, thr_free: float, class_mask: int, want=("xyz", "prob", "cls")):
        counts = (C.c_uint64 * 3)()
        n = C.c_uint64()
        _check(self._lib.s3d_export_begin(self._h, float(thr_occ), float(thr_free), int(class_mask), counts, C.byref(n)))
        n = int(n.value)
        out = {"counts": [int(c) for c in counts], "n": n}
        xyz = np.empty((n, 3), dtype=np.float64) if "xyz" in want else None
        prob = np.empty(n, dtype=np.float64) if "prob" in want else None
        cls = np.empty(n, dtype=np.int8) if "cls" in want else None
        ijk = np.empty((n, 3), dtype=np.int32) if "ijk" in want else None
        if n and (xyz is not None or prob is not None or cls is not None or ijk is not None):
            _check(self._lib.s3d_export_read(self._h, _ptr(xyz, C.c_double), _ptr(prob, C.c_double),
                                             _ptr(cls, C.c_int8), _ptr(ijk, C.c_int32), n))
        if "xyzi32" in want:
            blob = np.empty((n, 4), dtype=np.float32)
            if n:
                _check(self._lib.s3d_export_read_xyzi32(self._h, blob.ctypes.data_as(C.POINTER(C.c_float)), n))
            out["xyzi32"] = blob
        out.update(xyz=xyz, prob=prob, cls=cls, ijk=ijk)
        return out
