"""Synthetic Oculus-M750D-shaped frames and poses (the KIRO bags are not available offline).

Used by bench.py, the golden-vector script and the tests.  The image layout is the
reference's: ``uint8[H, W]`` with rows = range bins and columns = bearings
(scripts/3d_mapper.py:508).  The generator follows SURVEY.md section 8(d):
exponential background speckle clipped below the threshold, a seabed return per beam
lasting ``seabed_thickness`` metres, optional sparse above-threshold spikes (they
exercise early first hits and hits below ``min_range``), and a lawnmower track whose
poses are jittered so that nothing is aligned with the voxel lattice.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

# Named workloads of BASELINE.json `configs` (sensor + map parameters only).
CONFIGS: Dict[str, Dict] = {
    # configs[0]: library defaults of the reference (3d_mapper.py:220-250)
    "cfg1": dict(H=500, W=512, frames=100, step_m=0.05, seabed_depth=4.0,
                 config=dict(horizontal_fov=130.0, vertical_aperture=20.0, max_range=10.0, min_range=0.5,
                             intensity_threshold=35, voxel_resolution=0.05)),
    # configs[1]: KIRO water-tank YAML values (config/3d_mapper.yaml:12-56) at 0.05 m voxels
    "cfg2": dict(H=500, W=512, frames=2000, step_m=0.05, seabed_depth=4.0,
                 config=dict(horizontal_fov=70.0, vertical_aperture=20.0, max_range=10.0, min_range=1.0,
                             intensity_threshold=120, sonar_position=[0.0, 0.0, -0.1],
                             sonar_orientation=[0.0, float(np.radians(60.0)), 0.0],
                             voxel_resolution=0.05, min_probability=0.7, dynamic_expansion=True,
                             z_filter_min=-6.3, z_filter_enabled=True, adaptive_update=True,
                             adaptive_threshold=0.5, adaptive_max_ratio=0.3, log_odds_occupied=0.5,
                             log_odds_free=-0.1, log_odds_min=-10.0, log_odds_max=7.0)),
    # configs[2]: high-resolution hash-table stress
    "cfg3": dict(H=2000, W=1024, frames=10000, step_m=0.05, seabed_depth=4.0,
                 config=dict(horizontal_fov=130.0, vertical_aperture=20.0, max_range=10.0, min_range=0.5,
                             intensity_threshold=35, voxel_resolution=0.02)),
    # configs[3]: large seabed survey (cfg-1 sensor, long track)
    "cfg4": dict(H=500, W=512, frames=100000, step_m=0.25, seabed_depth=4.0,
                 config=dict(horizontal_fov=130.0, vertical_aperture=20.0, max_range=10.0, min_range=0.5,
                             intensity_threshold=35, voxel_resolution=0.05)),
}


def make_frame(rng: np.random.Generator, H: int, W: int, *, fov_deg: float = 130.0, max_range: float = 10.0,
               threshold: int = 35, seabed_depth: float = 4.0, seabed_thickness: float = 0.6,
               spike_prob: float = 1e-3, range_sigma: float = 0.03) -> np.ndarray:
    """One polar intensity image, uint8[H, W] (rows = range, cols = bearing)."""
    thr = int(threshold)
    bg = rng.exponential(8.0, size=(H, W))
    img = np.minimum(bg, max(thr - 1, 0)).astype(np.uint8)
    rr = max_range / H
    bearings = np.linspace(-np.radians(fov_deg) / 2, np.radians(fov_deg) / 2, W)
    start_m = seabed_depth / np.cos(bearings / 2) + rng.normal(0.0, range_sigma, size=W)
    start = np.clip((start_m / rr).astype(np.int64), 0, H)
    stop = np.clip(start + max(1, int(round(seabed_thickness / rr))), 0, H)
    rows = np.arange(H)[:, None]
    seabed = (rows >= start[None, :]) & (rows < stop[None, :])
    lo = min(thr + 1, 254)
    vals = rng.integers(lo, 255, size=(H, W), dtype=np.int64).astype(np.uint8)
    img = np.where(seabed, vals, img)
    if spike_prob > 0:
        spikes = rng.random(size=(H, W)) < spike_prob
        img = np.where(spikes, vals, img)
    return np.ascontiguousarray(img, dtype=np.uint8)


def _quat_from_rpy(roll: float, pitch: float, yaw: float) -> np.ndarray:
    cy, sy = np.cos(yaw * 0.5), np.sin(yaw * 0.5)
    cp, sp = np.cos(pitch * 0.5), np.sin(pitch * 0.5)
    cr, sr = np.cos(roll * 0.5), np.sin(roll * 0.5)
    q = np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                  cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy])
    return q / np.linalg.norm(q)


def make_poses(rng: np.random.Generator, n: int, *, step_m: float = 0.05, leg_m: float = 20.0,
               lane_m: float = 2.0, pos_sigma: float = 1e-3, att_sigma_deg: float = 0.5
               ) -> Tuple[np.ndarray, np.ndarray]:
    """Lawnmower track: positions float64[n,3], unit quaternions xyzw float64[n,4]."""
    pos = np.zeros((n, 3))
    quat = np.zeros((n, 4))
    per_leg = max(1, int(round(leg_m / step_m)))
    for i in range(n):
        leg, k = divmod(i, per_leg)
        forward = (leg % 2 == 0)
        x = k * step_m if forward else (per_leg - 1 - k) * step_m
        y = leg * lane_m
        yaw = 0.0 if forward else np.pi
        pos[i] = np.array([x, y, 0.0]) + rng.normal(0.0, pos_sigma, size=3)
        roll, pitch = np.radians(rng.normal(0.0, att_sigma_deg, size=2))
        quat[i] = _quat_from_rpy(roll, pitch, yaw + np.radians(rng.normal(0.0, att_sigma_deg)))
    return pos, quat


def make_sequence(name_or_spec, n_frames: int, seed: int = 0, distinct_images: int = 0, cycle: bool = True):
    """Frames + poses for a named workload.

    Returns (images uint8[n,H,W], positions[n,3], quaternions[n,4], config dict).  With
    ``distinct_images`` > 0 only that many distinct images are generated and cycled
    (poses stay distinct) -- generating thousands of 256 KB frames on the host is slow
    and the kernels' work depends on the pose, not on which speckle realisation is used.
    ``cycle=False`` returns only the distinct images (frame f uses image f % len(images)).
    """
    spec = CONFIGS[name_or_spec] if isinstance(name_or_spec, str) else name_or_spec
    cfg = dict(spec["config"])
    rng = np.random.default_rng(seed)
    H, W = spec["H"], spec["W"]
    k = n_frames if distinct_images <= 0 else min(n_frames, distinct_images)
    base = np.stack([make_frame(rng, H, W, fov_deg=cfg.get("horizontal_fov", 130.0),
                                max_range=cfg.get("max_range", 10.0),
                                threshold=cfg.get("intensity_threshold", 35),
                                seabed_depth=spec.get("seabed_depth", 4.0)) for _ in range(k)])
    images = base if (k == n_frames or not cycle) else base[np.arange(n_frames) % k]
    pos, quat = make_poses(rng, n_frames, step_m=spec.get("step_m", 0.05))
    return images, pos, quat, cfg
