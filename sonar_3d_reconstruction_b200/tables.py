"""Host-side geometry tables for the expansion kernel.

Every transcendental the reference evaluates per sample -- cos/sin of the bearing, cos/sin of
the vertical angle, tan of the half aperture (scripts/3d_mapper.py:426-436, :462-471) -- has
its argument in a small finite set, so the tables are built here with numpy using the
reference's own expressions and uploaded once per image shape.  The device then only does
IEEE multiplies/adds, which is what makes the voxel keys bit-exact.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

FREE_SAMPLING_STEP = 10   # scripts/3d_mapper.py:419
OCCUPIED_WINDOW = 50      # scripts/3d_mapper.py:451
MAX_BEAMS_DIVISOR = 256   # scripts/3d_mapper.py:528
MAX_FAN = 20000           # guard against absurd (range / resolution) ratios


@dataclass
class HostTables:
    H: int
    W: int
    nv_max: int
    free_step: int
    occ_window: int
    beam_col: np.ndarray   # int32[n_beams]
    cos_b: np.ndarray      # float64[n_beams]
    sin_b: np.ndarray
    range_m: np.ndarray    # float64[H]
    nv_free: np.ndarray    # int32[H]
    nv_occ: np.ndarray     # int32[H]
    cos_va: np.ndarray     # float64[nv_max*(nv_max+2)]
    sin_va: np.ndarray

    @property
    def n_beams(self) -> int:
        return len(self.beam_col)

    def samples_upper_bound(self) -> int:
        """Worst-case samples one frame can emit (all free bins + the densest occupied window)."""
        free = int(np.where(self.nv_free[::self.free_step] > 0, 2 * self.nv_free[::self.free_step] + 1, 0).sum())
        fan = np.where(self.nv_occ > 0, 2 * self.nv_occ + 1, 0).astype(np.int64)
        if len(fan) == 0:
            return 0
        c = np.concatenate([[0], np.cumsum(fan)])
        w = min(self.occ_window, len(fan))
        occ = int((c[w:] - c[:-w]).max()) if w > 0 else 0
        return (free + occ) * self.n_beams


def fan_offset(nv: int) -> int:
    """Start of row `nv` (v_step = -nv .. nv) in the cos_va / sin_va tables."""
    return nv * nv - 1


def build_tables(bearing_angles: np.ndarray, horizontal_fov: float, vertical_aperture: float, max_range: float,
                 min_range: float, voxel_resolution: float, H: int, W: int) -> HostTables:
    """Tables for an H x W polar image.  Arguments are the mapper's attributes, radians where
    the reference stores radians (horizontal_fov, vertical_aperture: scripts/3d_mapper.py:257-258)."""
    # processed beams and the FOV gate (:528-535, :382-385)
    step = max(1, W // MAX_BEAMS_DIVISOR)
    half_fov = horizontal_fov / 2
    cols = [b for b in range(0, W, step) if abs(bearing_angles[b]) <= half_fov]
    beam_col = np.asarray(cols, dtype=np.int32)
    ang = np.asarray([bearing_angles[b] for b in cols], dtype=np.float64)
    cos_b, sin_b = np.cos(ang), np.sin(ang)                          # :434-435

    # per range bin (:404, :421-427, :453-463)
    rr = max_range / H if H > 0 else 0.0
    range_m = np.arange(H, dtype=np.int64) * rr
    half_ap = vertical_aperture / 2                                   # :416
    spread = range_m * np.tan(half_ap)                                # :426
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        nv_free = np.maximum(1, np.trunc(spread / (voxel_resolution * 4))).astype(np.int64)     # :427
        nv_occ = np.maximum(2, np.trunc(spread / (voxel_resolution * 1.5))).astype(np.int64)    # :463
    too_near = range_m < min_range                                    # :422, :456
    nv_free[too_near] = 0
    nv_occ[too_near] = 0
    beyond = range_m > max_range                                      # :458 (break; ranges are monotone)
    if beyond.any():
        nv_occ[np.argmax(beyond):] = 0
    nv_max = int(max(nv_free.max(initial=0), nv_occ.max(initial=0)))
    if nv_max > MAX_FAN:
        raise ValueError(f"vertical fan of {2 * nv_max + 1} samples per range bin: voxel_resolution "
                         f"{voxel_resolution} is too fine for max_range {max_range}")

    # vertical fan (:429-436, :465-471): va = (v_step / max(1, nv)) * half_aperture
    nvs = np.repeat(np.arange(1, nv_max + 1, dtype=np.int64), 2 * np.arange(1, nv_max + 1) + 1)
    starts = np.arange(1, nv_max + 1, dtype=np.int64) ** 2 - 1
    v = np.arange(len(nvs), dtype=np.int64) - np.repeat(starts, 2 * np.arange(1, nv_max + 1) + 1) - nvs
    va = (v / np.maximum(1, nvs)) * half_ap
    cos_va, sin_va = np.cos(va), np.sin(va)

    c = np.ascontiguousarray
    return HostTables(H=int(H), W=int(W), nv_max=nv_max, free_step=FREE_SAMPLING_STEP, occ_window=OCCUPIED_WINDOW,
                      beam_col=c(beam_col), cos_b=c(cos_b, dtype=np.float64), sin_b=c(sin_b, dtype=np.float64),
                      range_m=c(range_m, dtype=np.float64), nv_free=c(nv_free.astype(np.int32)),
                      nv_occ=c(nv_occ.astype(np.int32)), cos_va=c(cos_va, dtype=np.float64),
                      sin_va=c(sin_va, dtype=np.float64))
