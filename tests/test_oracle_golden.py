"""CPU oracle vs the recorded outputs of the unmodified reference (tests/golden/).

This is what pins the oracle (tier rule 3): voxel keys, insertion order and per-frame
counters must be bit-exact; log-odds within 1e-12 (observed 0.0 in the build container).
"""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, assert_same_map, golden_config, load_golden
from oracle.oracle import OracleMapper, compose, transform_from_odometry, transform_from_rpy

SEQS = ["seq_cfg1_default", "seq_kiro_yaml", "seq_small_noadapt", "seq_wide_step4", "seq_overlap_clamp"]


def test_selftest_known_answers():
    """The reference's own __main__ sequence (scripts/3d_mapper.py:653-683), SURVEY.md section 4 table."""
    with open(os.path.join(GOLDEN, "selftest_known_answers.json")) as f:
        ka = json.load(f)
    assert ka["stats"] == [[3716, 26277, 29993], [3716, 26277, 57949], [3716, 26277, 84325]]
    assert ka["sha_keys"] == "9ea27058f23cda54" and ka["sha_vals"] == "0443b04dd46e9b62"
    m = OracleMapper({"voxel_resolution": 0.1, "min_probability": 0.6, "intensity_threshold": 30})
    img = np.zeros((500, 512), dtype=np.uint8)
    img[100:150, 200:300] = 100
    img[300:350, 100:150] = 150
    stats = []
    for i in range(3):
        st = m.process_sonar_image(img, [i * 0.1, 0, 0], [0, 0, 0, 1])
        stats.append([st["num_occupied"], st["num_free"], st["num_voxels"]])
    assert stats == ka["stats"]
    keys, vals = m.dump()
    order = np.lexsort(keys.T[::-1])
    assert hashlib.sha256(keys[order].tobytes()).hexdigest()[:16] == ka["sha_keys"]
    assert hashlib.sha256(vals[order].tobytes()).hexdigest()[:16] == ka["sha_vals"]
    assert float(vals.sum()) == ka["sum_logodds"]
    assert keys.min(0).tolist() == ka["key_min"] and keys.max(0).tolist() == ka["key_max"]
    pc = m.get_point_cloud()
    assert pc["num_occupied"] == ka["num_occupied"] == 5710
    pcf = m.get_point_cloud(True)
    assert [pcf["num_occupied"], pcf["num_free"], pcf["num_unknown"]] == ka["counts"]
    mn, mx = m.bounds()
    assert mn.tolist() == ka["min_bounds"] and mx.tolist() == ka["max_bounds"]


@pytest.mark.parametrize("name", SEQS)
def test_sequences(name):
    g = load_golden(name)
    m = OracleMapper(golden_config(g))
    for f in range(len(g["images"])):
        st = m.process_sonar_image(g["images"][f], g["positions"][f], g["quaternions"][f])
        assert [st["num_occupied"], st["num_free"], st["num_voxels"]] == g["stats"][f].tolist(), f"frame {f}"
        if f"ckpt{f}_keys" in g:
            k, v = m.dump()
            assert np.array_equal(k, g[f"ckpt{f}_keys"])            # same insertion order too
            assert np.abs(v - g[f"ckpt{f}_logodds"]).max() <= 1e-12
    keys, vals = m.dump()
    assert np.array_equal(keys, g["keys"].astype(np.int64))
    assert np.abs(vals - g["logodds"]).max() <= 1e-12
    pc = m.get_point_cloud()
    assert np.array_equal(pc["points"], g["pc_points"])
    assert np.abs(pc["probabilities"] - g["pc_prob"]).max() <= 1e-15
    pcf = m.get_point_cloud(True)
    assert [pcf["num_occupied"], pcf["num_free"], pcf["num_unknown"]] == g["counts"].tolist()
    mn, mx = m.bounds()
    assert np.array_equal(mn, g["min_bounds"]) and np.array_equal(mx, g["max_bounds"])
    assert np.array_equal(m.bearing_angles, g["bearing_angles"])
    assert np.abs(m.T_sonar_to_base - g["T_sonar_to_base"]).max() <= 1e-15


def test_edge_frames():
    g = load_golden("edge_frames")
    cfg = golden_config(g)
    for name in ("nohit", "allhit", "hit_at_zero", "late_hit", "random", "equal_thr"):
        m = OracleMapper(cfg)
        st = m.process_sonar_image(g[f"{name}__image"], g["positions"][0], g["quaternions"][0])
        assert [[st["num_occupied"], st["num_free"], st["num_voxels"]]] == g[f"{name}__stats"].tolist(), name
        k, v = m.dump()
        assert np.array_equal(k, g[f"{name}__keys"].reshape(-1, 3)), name
        assert np.abs(v - g[f"{name}__logodds"]).max(initial=0.0) <= 1e-12
    m = OracleMapper(dict(cfg, intensity_threshold=99.5))
    st = m.process_sonar_image(g["onebeam__image"], g["positions"][0], g["quaternions"][0])
    assert [[st["num_occupied"], st["num_free"], st["num_voxels"]]] == g["onebeam__stats"].tolist()
    k, v = m.dump()
    assert_same_map(k, v, g["onebeam__keys"], g["onebeam__logodds"], atol=1e-12)


def test_stage_vectors():
    """first hit per beam, world xyz of every sample (bit-exact), keys (bit-exact)."""
    g = load_golden("stage_vectors")
    m = OracleMapper(golden_config(g))
    fh = m.first_hits(g["image"], g["T"])
    assert np.array_equal(fh, g["first_hits"])
    xyz, occ = m.expand_frame(g["image"], g["T"])
    assert xyz.shape == g["xyz"].shape
    assert np.array_equal(occ, g["occupied"])
    assert np.array_equal(xyz, g["xyz"])          # same rounding order as numpy's 4x4 @ 4 here
    assert np.array_equal(m.world_to_key(xyz), g["keys"].astype(np.int64))
    # the pose -> 4x4 chain (3d_mapper.py:346-380, :521); BLAS-dependent last bit => 2 ulp allowance
    T = compose(transform_from_odometry(g["position"], g["quaternion"]), m.T_sonar_to_base)
    assert np.allclose(T, g["T"], rtol=0, atol=4e-16 * max(1.0, np.abs(g["T"]).max()))


def test_store_vectors():
    g = load_golden("store_vectors")
    m = OracleMapper({"voxel_resolution": 0.05, "adaptive_max_ratio": 0.3})
    for p, u, a in zip(g["points"], g["updates"], g["adaptive"]):
        m.update_voxel(p, float(u), bool(a))
    k, v = m.dump()
    assert np.array_equal(k, g["keys"].astype(np.int64))
    assert np.abs(v - g["logodds"]).max() <= 1e-12
    lo = np.array([m.get_log_odds(*p) for p in g["query_points"]])
    pr = np.array([m.get_probability(*p) for p in g["query_points"]])
    assert np.abs(lo - g["query_logodds"]).max() <= 1e-12
    assert np.abs(pr - g["query_prob"]).max() <= 1e-15
    assert m.num_voxels() == len(g["keys"])           # queries never insert
    mn, mx = m.bounds()
    assert np.array_equal(mn, g["min_bounds"]) and np.array_equal(mx, g["max_bounds"])
    edge = json.loads(str(g["edge_thr_json"]))
    for p, key in ((1.0, "p1"), (0.0, "p0"), (0.5, "p05")):
        m.min_probability = p
        assert m.get_point_cloud()["num_occupied"] == edge[key]
    m.min_probability = 0.6
    pc = m.get_point_cloud()
    assert np.array_equal(pc["points"], g["occ_points"])
    m.min_probability = 0.7
    pcf = m.get_point_cloud(True)
    assert [pcf["num_free"], pcf["num_unknown"], pcf["num_occupied"]] == g["cls_counts"].tolist()


def test_rpy_transform_matches_numpy_expression():
    rng = np.random.default_rng(0)
    for _ in range(50):
        pos, rpy = rng.normal(size=3), rng.uniform(-3, 3, size=3)
        cr, sr, cp, sp, cy, sy = (np.cos(rpy[0]), np.sin(rpy[0]), np.cos(rpy[1]), np.sin(rpy[1]),
                                  np.cos(rpy[2]), np.sin(rpy[2]))
        R = np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                      [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                      [-sp, cp * sr, cp * cr]])
        T = transform_from_rpy(pos, rpy)
        assert np.array_equal(T[:3, :3], R) and np.array_equal(T[:3, 3], pos)
