"""Property tests over the configuration space (SURVEY.md section 4 iii): threshold, min/max range, voxel size,
aperture, field of view, mount, z filter, adaptive rule on/off, log-odds deltas and clamps, image shape.

* CPU (`-m "not gpu"`, build container only -- it needs /root/reference): the C oracle against the UNMODIFIED
  reference imported by path, on hypothesis-drawn configurations.  This widens the oracle's pin beyond the
  committed golden vectors.
* GPU (`-m gpu`): the CUDA path against the oracle on seeded random configurations of the same space, through
  the single-frame call and the batched call.
Bar: per-frame counters and voxel-key sets exact, |d log-odds| <= 1e-5 (observed <= 1e-12)."""
import contextlib
import importlib.util
import io
import os

import numpy as np
import pytest

from helpers import assert_same_map

REF_PATH = os.environ.get("S3D_REFERENCE", "/root/reference/scripts/3d_mapper.py")


def _draw_case(rng):
    """One configuration + a short posed sequence.  Everything is drawn from `rng` (a numpy Generator)."""
    from sonar_3d_reconstruction_b200 import synthetic
    H = int(rng.integers(24, 90))
    W = int(rng.choice([1, 7, 16, 33, 48, 64, 300]))
    max_range = float(rng.uniform(2.0, 12.0))
    cfg = dict(
        horizontal_fov=float(rng.uniform(20.0, 160.0)), vertical_aperture=float(rng.uniform(2.0, 40.0)),
        max_range=max_range, min_range=float(rng.uniform(0.0, 0.4 * max_range)),
        intensity_threshold=(int(rng.integers(0, 255)) if rng.random() < 0.8 else float(rng.uniform(0, 255))),
        sonar_position=[float(x) for x in rng.normal(0, 0.3, 3)],
        sonar_orientation=[float(x) for x in rng.uniform(-np.pi, np.pi, 3)],
        voxel_resolution=float(rng.choice([0.03, 0.05, 0.11, 0.25, 0.7])),
        min_probability=float(rng.uniform(0.05, 0.95)), dynamic_expansion=bool(rng.random() < 0.9),
        adaptive_update=bool(rng.random() < 0.6), adaptive_threshold=float(rng.uniform(0.2, 0.8)),
        adaptive_max_ratio=float(rng.uniform(0.05, 1.0)),
        log_odds_occupied=float(rng.choice([1.5, 0.5, 0.85, 2.2])), log_odds_free=float(rng.choice([-2.0, -0.1, -0.4, -1.0])),
        log_odds_min=float(rng.uniform(-12.0, -1.0)), log_odds_max=float(rng.uniform(1.0, 12.0)),
        z_filter_enabled=bool(rng.random() < 0.5), z_filter_min=float(rng.uniform(-6.0, 1.0)),
    )
    n = int(rng.integers(2, 5))
    thr = int(np.floor(cfg["intensity_threshold"]))
    images = np.stack([synthetic.make_frame(rng, H, W, fov_deg=cfg["horizontal_fov"], max_range=max_range,
                                            threshold=max(1, min(thr, 250)), seabed_depth=float(rng.uniform(0.3, 0.9)) * max_range,
                                            spike_prob=float(rng.choice([0.0, 1e-3, 2e-2]))) for _ in range(n)])
    pos, quat = synthetic.make_poses(rng, n, step_m=float(rng.uniform(0.0, 0.3)))
    pos = pos + rng.normal(0, 2.0, 3)
    if rng.random() < 0.3:                       # the reference does not normalise quaternions (:346-366)
        quat = quat * rng.uniform(0.9, 1.1)
    return cfg, images, pos, quat


def _run(mapper, images, pos, quat):
    out = []
    with contextlib.redirect_stdout(io.StringIO()):
        for f in range(len(images)):
            st = mapper.process_sonar_image(images[f], list(pos[f]), list(quat[f]))
            out.append([st["num_occupied"], st["num_free"], st["num_voxels"]])
    return out


@pytest.mark.skipif(not os.path.exists(REF_PATH), reason="needs the unmodified reference (build container)")
def test_oracle_equals_reference_over_the_config_space():
    from hypothesis import HealthCheck, given, settings, strategies as st
    from oracle.oracle import OracleMapper
    spec = importlib.util.spec_from_file_location("reference_3d_mapper", REF_PATH)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    @settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(st.integers(min_value=0, max_value=2**31 - 1))
    def check(seed):
        cfg, images, pos, quat = _draw_case(np.random.default_rng(seed))
        r, o = ref.SonarTo3DMapper(dict(cfg)), OracleMapper(dict(cfg))
        assert _run(r, images, pos, quat) == _run(o, images, pos, quat), cfg
        kr = np.array(list(r.octree.voxels.keys()), dtype=np.int64).reshape(-1, 3)
        vr = np.array([float(v) for v in r.octree.voxels.values()], dtype=np.float64)
        ko, vo = o.dump()
        assert np.array_equal(kr, ko), "dict insertion order differs"            # the oracle even keeps the order
        assert np.abs(vr - vo).max(initial=0.0) <= 1e-12
        pr, po = r.get_point_cloud(False), o.get_point_cloud(False)
        assert pr["num_occupied"] == po["num_occupied"]

    check()


@pytest.mark.gpu
def test_cuda_equals_oracle_over_the_config_space():
    import sonar_3d_reconstruction_b200 as s3d
    from oracle.oracle import OracleMapper
    for seed in range(24):
        cfg, images, pos, quat = _draw_case(np.random.default_rng(1000 + seed))
        o = OracleMapper(dict(cfg))
        want = _run(o, images, pos, quat)
        g = s3d.SonarTo3DMapper(dict(cfg))
        assert _run(g, images, pos, quat) == want, (seed, cfg)
        assert assert_same_map(*g.octree.voxels.to_arrays(), *o.dump(), 1e-5, f"seed {seed}") <= 1e-9
        b = s3d.SonarTo3DMapper(dict(cfg))
        got = [[s["num_occupied"], s["num_free"], s["num_voxels"]] for s in b.process_sonar_images(images, pos, quat)]
        assert got == want, (seed, "batched")
        assert_same_map(*b.octree.voxels.to_arrays(), *g.octree.voxels.to_arrays(), 0.0, f"seed {seed} batched")
        assert g.get_point_cloud()["num_occupied"] == o.get_point_cloud()["num_occupied"]
