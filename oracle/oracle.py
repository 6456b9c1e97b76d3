"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

`OracleMapper` exposes the C restatement in ``sonar_oracle.c`` behind the same
method names as the reference's ``SonarTo3DMapper`` (scripts/3d_mapper.py:197)
so that parity tests read like calls into the reference.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg may import this module; the shipped package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Any, Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsonar_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "sonar_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s", "libsonar_oracle.so"])
    return _LIB_PATH


class _Config(C.Structure):
    _fields_ = [
        ("horizontal_fov_deg", C.c_double), ("vertical_aperture_deg", C.c_double),
        ("max_range", C.c_double), ("min_range", C.c_double),
        ("intensity_threshold", C.c_double),
        ("image_width", C.c_int32), ("image_height", C.c_int32),
        ("sonar_position", C.c_double * 3), ("sonar_orientation", C.c_double * 3),
        ("voxel_resolution", C.c_double), ("min_probability", C.c_double),
        ("dynamic_expansion", C.c_int32), ("adaptive_update", C.c_int32),
        ("adaptive_threshold", C.c_double), ("adaptive_max_ratio", C.c_double),
        ("log_odds_occupied", C.c_double), ("log_odds_free", C.c_double),
        ("log_odds_min", C.c_double), ("log_odds_max", C.c_double),
        ("z_filter_min", C.c_double), ("z_filter_enabled", C.c_int32), ("_pad", C.c_int32),
    ]


class _Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ("frame_count", "processed_count", "num_occupied", "num_free", "num_voxels", "num_samples")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip8, i64p, i32p, u8p = (C.POINTER(C.c_double), C.POINTER(C.c_int8), C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int32), C.POINTER(C.c_uint8))
        L.so_default_config.argtypes = [C.POINTER(_Config)]
        L.so_create.restype = C.c_void_p
        L.so_create.argtypes = [C.POINTER(_Config)]
        L.so_destroy.argtypes = [C.c_void_p]
        L.so_ingest_T.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, dp, C.POINTER(_Stats)]
        L.so_ingest.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, dp, dp, C.POINTER(_Stats)]
        L.so_first_hits.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, dp, i32p]
        L.so_expand_frame.restype = C.c_int64
        L.so_expand_frame.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, dp, dp, ip8, C.c_int64]
        L.so_bearing_table.argtypes = [C.c_void_p, dp]
        L.so_bearing_count.argtypes = [C.c_void_p]
        L.so_set_beam_slice.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.so_sonar_to_base.argtypes = [C.c_void_p, dp]
        L.so_set_octree_params.argtypes = [C.c_void_p] + [C.c_double] * 4 + [C.c_int] + [C.c_double] * 2
        L.so_update_voxel.argtypes = [C.c_void_p, dp, C.c_double, C.c_int]
        L.so_world_to_key_n.argtypes = [C.c_void_p, dp, C.c_int64, i64p]
        L.so_get_log_odds.restype = C.c_double
        L.so_get_log_odds.argtypes = [C.c_void_p] + [C.c_double] * 3
        L.so_get_probability.restype = C.c_double
        L.so_get_probability.argtypes = [C.c_void_p] + [C.c_double] * 3
        L.so_num_voxels.restype = C.c_int64
        L.so_num_voxels.argtypes = [C.c_void_p]
        L.so_dump.restype = C.c_int64
        L.so_dump.argtypes = [C.c_void_p, i64p, dp, C.c_int64]
        L.so_get_occupied.restype = C.c_int64
        L.so_get_occupied.argtypes = [C.c_void_p, C.c_double, dp, dp, C.c_int64]
        L.so_classify.restype = C.c_int64
        L.so_classify.argtypes = [C.c_void_p, C.c_double, ip8, dp, dp, C.c_int64]
        L.so_bounds.argtypes = [C.c_void_p, dp, dp]
        L.so_reset.argtypes = [C.c_void_p]
        L.so_clear_octree.argtypes = [C.c_void_p]
        L.so_transform_from_rpy.argtypes = [dp, dp, dp]
        L.so_transform_from_odometry.argtypes = [dp, dp, dp]
        L.so_compose.argtypes = [dp, dp, dp]
        _lib = L
    return _lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _as_u8(image) -> np.ndarray:
    img = np.ascontiguousarray(image)
    if img.dtype != np.uint8:
        raise TypeError("oracle expects uint8 images")
    if img.ndim != 2:
        raise ValueError("image must be 2-D (range bins x bearings)")
    return img


_CONFIG_KEYS = ("horizontal_fov", "vertical_aperture", "max_range", "min_range", "intensity_threshold",
                "image_width", "image_height", "sonar_position", "sonar_orientation", "voxel_resolution",
                "min_probability", "dynamic_expansion", "adaptive_update", "adaptive_threshold",
                "adaptive_max_ratio", "log_odds_occupied", "log_odds_free", "log_odds_min", "log_odds_max",
                "z_filter_min", "z_filter_enabled")


class OracleMapper:
    """C oracle behind the reference's SonarTo3DMapper method names."""

    def __init__(self, config: Optional[Dict[str, Any]] = None):
        L = lib()
        c = _Config()
        L.so_default_config(C.byref(c))
        for k, v in (config or {}).items():
            if k not in _CONFIG_KEYS:
                continue
            if k == "horizontal_fov":
                c.horizontal_fov_deg = float(v)
            elif k == "vertical_aperture":
                c.vertical_aperture_deg = float(v)
            elif k in ("sonar_position", "sonar_orientation"):
                for a in range(3):
                    getattr(c, k)[a] = float(v[a])
            elif k in ("image_width", "image_height", "dynamic_expansion", "adaptive_update", "z_filter_enabled"):
                setattr(c, k, int(v))
            else:
                setattr(c, k, float(v))
        self._cfg = c
        self._h = C.c_void_p(L.so_create(C.byref(c)))
        self.voxel_resolution = c.voxel_resolution
        self.min_probability = c.min_probability
        self.intensity_threshold = c.intensity_threshold
        self.frame_count = 0
        self.last_num_samples = 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.so_destroy(h)

    # -- ingest ---------------------------------------------------------------
    def process_sonar_image(self, polar_image, robot_position, robot_orientation) -> Dict[str, Any]:
        img = _as_u8(polar_image)
        H, W = img.shape
        pos = np.asarray(robot_position, dtype=np.float64)
        quat = np.asarray(robot_orientation, dtype=np.float64)
        st = _Stats()
        lib().so_ingest(self._h, img.ctypes.data_as(C.POINTER(C.c_uint8)), H, W, _dptr(pos), _dptr(quat), C.byref(st))
        return self._stats(st)

    def process_sonar_image_T(self, polar_image, T_sonar_to_world) -> Dict[str, Any]:
        """Same as above with the composed 4x4 supplied (3d_mapper.py:521)."""
        img = _as_u8(polar_image)
        H, W = img.shape
        T = np.ascontiguousarray(T_sonar_to_world, dtype=np.float64).reshape(16)
        st = _Stats()
        lib().so_ingest_T(self._h, img.ctypes.data_as(C.POINTER(C.c_uint8)), H, W, _dptr(T), C.byref(st))
        return self._stats(st)

    def _stats(self, st):
        self.frame_count = int(st.frame_count)
        self.last_num_samples = int(st.num_samples)
        return {"frame_count": int(st.frame_count), "processed_count": int(st.processed_count),
                "num_occupied": int(st.num_occupied), "num_free": int(st.num_free),
                "num_voxels": int(st.num_voxels), "num_samples": int(st.num_samples)}

    # -- per-stage probes -----------------------------------------------------
    def first_hits(self, polar_image, T=None) -> np.ndarray:
        img = _as_u8(polar_image)
        H, W = img.shape
        step = max(1, W // 256)
        out = np.zeros(len(range(0, W, step)), dtype=np.int32)
        T = np.eye(4).reshape(16) if T is None else np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        lib().so_first_hits(self._h, img.ctypes.data_as(C.POINTER(C.c_uint8)), H, W, _dptr(T),
                            out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out

    def expand_frame(self, polar_image, T):
        """World-frame samples of one frame, in the reference's emission order."""
        img = _as_u8(polar_image)
        H, W = img.shape
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        u8 = img.ctypes.data_as(C.POINTER(C.c_uint8))
        n = lib().so_expand_frame(self._h, u8, H, W, _dptr(T), None, None, 0)
        xyz = np.empty((n, 3), dtype=np.float64)
        occ = np.empty(n, dtype=np.int8)
        lib().so_expand_frame(self._h, u8, H, W, _dptr(T), _dptr(xyz), occ.ctypes.data_as(C.POINTER(C.c_int8)), n)
        return xyz, occ

    def set_beam_slice(self, lo: int, hi: int):
        """Expand only processed beams [lo, hi) (not in the reference; used to test the sharded path)."""
        lib().so_set_beam_slice(self._h, int(lo), int(hi))

    def world_to_key(self, xyz) -> np.ndarray:
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        out = np.empty((len(xyz), 3), dtype=np.int64)
        lib().so_world_to_key_n(self._h, _dptr(xyz), len(xyz), out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out

    @property
    def bearing_angles(self) -> np.ndarray:
        out = np.empty(lib().so_bearing_count(self._h), dtype=np.float64)
        lib().so_bearing_table(self._h, _dptr(out))
        return out

    @property
    def T_sonar_to_base(self) -> np.ndarray:
        out = np.empty(16, dtype=np.float64)
        lib().so_sonar_to_base(self._h, _dptr(out))
        return out.reshape(4, 4)

    # -- store ----------------------------------------------------------------
    def set_octree_params(self, log_odds_occupied, log_odds_free, log_odds_min, log_odds_max,
                          adaptive_update, adaptive_threshold, adaptive_max_ratio):
        lib().so_set_octree_params(self._h, log_odds_occupied, log_odds_free, log_odds_min, log_odds_max,
                                   int(adaptive_update), adaptive_threshold, adaptive_max_ratio)

    def update_voxel(self, point, log_odds_update, adaptive=True):
        p = np.asarray(point, dtype=np.float64)
        lib().so_update_voxel(self._h, _dptr(p), float(log_odds_update), int(bool(adaptive)))

    def get_log_odds(self, x, y, z) -> float:
        return lib().so_get_log_odds(self._h, x, y, z)

    def get_probability(self, x, y, z) -> float:
        return lib().so_get_probability(self._h, x, y, z)

    def num_voxels(self) -> int:
        return int(lib().so_num_voxels(self._h))

    def dump(self):
        """(keys int64[n,3], log_odds float64[n]) in dict insertion order."""
        n = self.num_voxels()
        keys = np.empty((n, 3), dtype=np.int64)
        L = np.empty(n, dtype=np.float64)
        lib().so_dump(self._h, keys.ctypes.data_as(C.POINTER(C.c_int64)), _dptr(L), n)
        return keys, L

    def bounds(self):
        mn, mx = np.empty(3), np.empty(3)
        lib().so_bounds(self._h, _dptr(mn), _dptr(mx))
        return mn, mx

    # -- export ---------------------------------------------------------------
    def get_point_cloud(self, include_free: bool = False) -> Dict[str, Any]:
        n = self.num_voxels()
        if include_free:
            cls = np.empty(n, dtype=np.int8)
            pts = np.empty((n, 3))
            pr = np.empty(n)
            lib().so_classify(self._h, self.min_probability, cls.ctypes.data_as(C.POINTER(C.c_int8)),
                              _dptr(pts), _dptr(pr), n)
            out = {}
            for name, c in (("free", 0), ("unknown", 1), ("occupied", 2)):
                sel = cls == c
                out[name] = list(zip(pts[sel], pr[sel]))
            mn, mx = self.bounds()
            dyn = bool(self._cfg.dynamic_expansion)
            out.update(num_voxels=n, num_occupied=len(out["occupied"]), num_free=len(out["free"]),
                       num_unknown=len(out["unknown"]),
                       bounds={"min": mn if dyn else None, "max": mx if dyn else None})
            return out
        pts = np.empty((n, 3))
        pr = np.empty(n)
        k = lib().so_get_occupied(self._h, self.min_probability, _dptr(pts), _dptr(pr), n)
        return {"points": pts[:k].copy(), "probabilities": pr[:k].copy(), "num_voxels": n, "num_occupied": int(k)}

    def reset_map(self):
        lib().so_reset(self._h)
        self.frame_count = 0


def transform_from_rpy(position, rpy) -> np.ndarray:
    p = np.asarray(position, dtype=np.float64)
    r = np.asarray(rpy, dtype=np.float64)
    T = np.empty(16)
    lib().so_transform_from_rpy(_dptr(p), _dptr(r), _dptr(T))
    return T.reshape(4, 4)


def transform_from_odometry(position, quaternion) -> np.ndarray:
    p = np.asarray(position, dtype=np.float64)
    q = np.asarray(quaternion, dtype=np.float64)
    T = np.empty(16)
    lib().so_transform_from_odometry(_dptr(p), _dptr(q), _dptr(T))
    return T.reshape(4, 4)


def compose(A, B) -> np.ndarray:
    A = np.ascontiguousarray(A, dtype=np.float64).reshape(16)
    B = np.ascontiguousarray(B, dtype=np.float64).reshape(16)
    out = np.empty(16)
    lib().so_compose(_dptr(A), _dptr(B), _dptr(out))
    return out.reshape(4, 4)
