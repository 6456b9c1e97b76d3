"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the
host tables follow the reference's arithmetic, the drop-in shim loads the way the reference
node loads it, and the product path refuses to run without a GPU (no CPU fallback)."""
import importlib.util
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import golden_config, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    from sonar_3d_reconstruction_b200 import _native
    from sonar_3d_reconstruction_b200.build import build_native
    so = build_native()
    header = open(os.path.join(ROOT, "include", "sonar3d.h")).read()
    declared = set(re.findall(r"\b(s3d_[a-z0-9_]+)\s*\(", header))
    declared -= {"s3d_map"}
    exported = set(re.findall(r" T (s3d_\w+)", subprocess.check_output(["nm", "-D", so], text=True)))
    assert declared == exported, (declared ^ exported)
    assert set(_native.SYMBOLS) == declared
    lib = _native.load_library()                  # dlopen + argtypes; needs no GPU
    assert lib.s3d_abi_version() == 2
    assert "S3D_ABI_VERSION 2" in header


def test_struct_layouts_match_header():
    import ctypes as C
    from sonar_3d_reconstruction_b200 import _native
    assert C.sizeof(_native.Params) == 8 * 8 + 4 * 4
    assert C.sizeof(_native.FrameStats) == 64 and _native.STATS_DTYPE.itemsize == 64
    assert C.sizeof(_native.Tables) == 6 * 4 + 8 * 8
    assert C.sizeof(_native.Profile) == 3 * 8 + 3 * 8 + 6 * 8


@pytest.mark.skipif(_cuda(), reason="only meaningful on a host without a GPU")
def test_no_cpu_fallback():
    import sonar_3d_reconstruction_b200 as pkg
    from sonar_3d_reconstruction_b200._native import NativeError
    with pytest.raises(NativeError, match="no CPU fallback"):
        pkg.SonarTo3DMapper({})
    with pytest.raises(NativeError):
        pkg.SimpleOctree(0.05)


def test_package_never_touches_the_oracle():
    pkg_dir = os.path.join(ROOT, "sonar_3d_reconstruction_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("CPU oracle", ""), f
    shim = open(os.path.join(ROOT, "scripts", "3d_mapper.py")).read()
    assert "oracle" not in shim


def test_drop_in_shim_loads_like_the_node_does():
    """scripts/3d_mapper_node.py:33-42: exec the file named 3d_mapper.py, read SonarTo3DMapper."""
    path = os.path.join(ROOT, "scripts", "3d_mapper.py")
    spec = importlib.util.spec_from_file_location("mapper_3d", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    M = mod.SonarTo3DMapper
    for name in ("process_sonar_image", "get_point_cloud", "reset_map", "create_transform_matrix",
                 "quaternion_to_matrix", "create_odometry_transform", "is_bearing_in_valid_fov"):
        assert callable(getattr(M, name))
    for name in ("world_to_key", "key_to_world", "update_voxel", "get_log_odds", "get_probability",
                 "get_occupied_voxels", "get_all_voxels_classified", "clear"):
        assert callable(getattr(mod.SimpleOctree, name))


def test_tables_follow_reference_arithmetic():
    from sonar_3d_reconstruction_b200 import tables
    for name in ("seq_cfg1_default", "seq_kiro_yaml", "seq_wide_step4"):
        g = load_golden(name)
        cfg = dict(horizontal_fov=130.0, vertical_aperture=20.0, max_range=10.0, min_range=0.5, voxel_resolution=0.05)
        cfg.update(golden_config(g))
        H, W = g["images"].shape[1:]
        fov, ap = np.radians(cfg["horizontal_fov"]), np.radians(cfg["vertical_aperture"])
        t = tables.build_tables(g["bearing_angles"], fov, ap, cfg["max_range"], cfg["min_range"],
                                cfg["voxel_resolution"], H, W)
        step = max(1, W // 256)
        assert t.beam_col.tolist() == list(range(0, W, step))
        assert np.array_equal(t.cos_b, np.cos(g["bearing_angles"][::step]))
        rr = cfg["max_range"] / H
        half, th = ap / 2, np.tan(ap / 2)
        for r in (0, 1, 10, 57, H // 2, H - 1):
            rm = r * rr
            assert t.range_m[r] == rm
            want_f = 0 if rm < cfg["min_range"] else max(1, int(rm * th / (cfg["voxel_resolution"] * 4)))
            want_o = 0 if rm < cfg["min_range"] else max(2, int(rm * th / (cfg["voxel_resolution"] * 1.5)))
            assert (t.nv_free[r], t.nv_occ[r]) == (want_f, want_o)
        for nv in (1, 2, t.nv_max):
            for v in (-nv, 0, nv):
                va = (v / max(1, nv)) * half
                i = tables.fan_offset(nv) + v + nv
                assert t.cos_va[i] == np.cos(va) and t.sin_va[i] == np.sin(va)
        assert len(t.cos_va) == t.nv_max * (t.nv_max + 2)
        assert t.samples_upper_bound() > 0


def test_threshold_to_int_equivalence():
    from sonar_3d_reconstruction_b200.mapper import _threshold_to_int
    px = np.arange(256)
    for thr in (-5, -1, 0, 0.5, 34.999, 35, 35.0, 99.5, 254, 255, 300, float("nan"), np.float64(120.0), np.uint8(30)):
        t = _threshold_to_int(thr)
        with np.errstate(invalid="ignore"):
            assert np.array_equal(px > t, px > thr), thr


def test_rekey_of_voxel_centre_is_identity():
    """update_voxel re-keys the voxel centre (3d_mapper.py:92); the device skips that step."""
    k = np.concatenate([np.arange(-(1 << 20), -(1 << 20) + 4096), np.arange(-4096, 4096), np.arange((1 << 20) - 4096, 1 << 20),
                        np.random.default_rng(0).integers(-(1 << 20), 1 << 20, size=200000)])
    for res in (0.02, 0.03, 0.05, 0.07, 0.1, 0.12, 0.15, 0.2, 1.0 / 3.0):
        centre = (k + 0.5) * res
        assert np.array_equal(np.floor(centre / res).astype(np.int64), k), res


def test_synthetic_generator_shapes_and_determinism():
    from sonar_3d_reconstruction_b200 import synthetic
    a = synthetic.make_sequence("cfg2", 3, seed=7)
    b = synthetic.make_sequence("cfg2", 3, seed=7)
    assert a[0].shape == (3, 500, 512) and a[0].dtype == np.uint8
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert np.allclose(np.linalg.norm(a[2], axis=1), 1.0)
    assert (a[0] > a[3]["intensity_threshold"]).any()


def test_vectorised_pose_composition_is_bit_identical_or_disabled():
    """compose_transforms may evaluate all poses at once only where that reproduces the
    pose-by-pose evaluation (scripts/3d_mapper.py:346-380, :521) bit for bit."""
    from sonar_3d_reconstruction_b200 import synthetic
    from sonar_3d_reconstruction_b200.mapper import SonarTo3DMapper as M

    class Host:
        quaternion_to_matrix = M.quaternion_to_matrix
        create_odometry_transform = M.create_odometry_transform
        _compose_scalar = M._compose_scalar
        _compose_vector = M._compose_vector
        compose_transforms = M.compose_transforms

    h = Host()
    h.T_sonar_to_base = M.create_transform_matrix(h, np.array([0.0, 0.0, -0.1]), np.array([0.0, np.radians(60.0), 0.0]))
    pos, quat = synthetic.make_poses(np.random.default_rng(3), 700)
    want = h._compose_scalar(pos, quat)
    got = h.compose_transforms(pos, quat)
    assert np.array_equal(got, want)                 # whichever path was chosen, the result is the scalar one
    one = h.compose_transforms(pos[:1], quat[:1])
    assert np.array_equal(one[0], want[0])


def test_reference_arm_of_the_bench_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) runs without a GPU and
    keeps the line contract: impl, metric/unit of our arm, cpu_baseline and a zero-copy e2e object."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-frames-per-step", "4", "--ref-port", "--frames-per-step", "16", "--distinct-images", "8"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sonar_frames_per_s" and d["unit"] == "frames/s"
    assert d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # with the unmodified reference shipped next to the repo (baseline/_ref, copied by build() in the build
    # container) the same arm times that file instead
    if os.path.exists(os.path.join(ROOT, "baseline", "_ref", "3d_mapper.py")):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                            "--frames-per-step", "4", "--distinct-images", "4"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        d = json.loads([l for l in r.stdout.strip().splitlines() if l.startswith("{")][0])
        assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["port_value"] > d["value"] > 0
