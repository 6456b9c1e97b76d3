"""CUDA path (through the Python mirror -> C-ABI -> sm_100a kernels) vs the golden vectors of
the unmodified reference and vs the CPU oracle.  Bar (BASELINE.json north_star): voxel-key sets
bit-exact, per-frame counters exact, |d log-odds| <= 1e-5 per voxel."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, assert_same_map, golden_config, load_golden, sort_by_key

pytestmark = pytest.mark.gpu

LOGODDS_ATOL = 1e-5          # north_star tolerance
SEQS = ["seq_cfg1_default", "seq_kiro_yaml", "seq_small_noadapt", "seq_wide_step4", "seq_overlap_clamp"]


@pytest.fixture(scope="module")
def s3d():
    import sonar_3d_reconstruction_b200 as pkg
    return pkg


def _stats3(st):
    return [st["num_occupied"], st["num_free"], st["num_voxels"]]


def test_selftest_known_answers(s3d):
    """The reference's own __main__ sequence (scripts/3d_mapper.py:653-683)."""
    with open(os.path.join(GOLDEN, "selftest_known_answers.json")) as f:
        ka = json.load(f)
    m = s3d.SonarTo3DMapper({"voxel_resolution": 0.1, "min_probability": 0.6, "intensity_threshold": 30})
    img = np.zeros((500, 512), dtype=np.uint8)
    img[100:150, 200:300] = 100
    img[300:350, 100:150] = 150
    stats = [_stats3(m.process_sonar_image(img, [i * 0.1, 0, 0], [0, 0, 0, 1])) for i in range(3)]
    assert stats == ka["stats"]
    keys, vals = m.octree.voxels.to_arrays()
    ks, vs = sort_by_key(keys, vals)
    assert hashlib.sha256(ks.tobytes()).hexdigest()[:16] == ka["sha_keys"]
    assert abs(float(vs.sum()) - ka["sum_logodds"]) <= 1e-6
    assert float(vs.min()) == ka["min_logodds"] and abs(float(vs.max()) - ka["max_logodds"]) <= LOGODDS_ATOL
    pc = m.get_point_cloud()
    assert pc["num_occupied"] == ka["num_occupied"] and pc["num_voxels"] == 84325
    assert pc["points"].shape == (5710, 3) and pc["probabilities"].shape == (5710,)
    pcf = m.get_point_cloud(include_free=True)
    assert [pcf["num_occupied"], pcf["num_free"], pcf["num_unknown"]] == ka["counts"]
    assert pcf["bounds"]["min"].tolist() == ka["min_bounds"] and pcf["bounds"]["max"].tolist() == ka["max_bounds"]
    assert (pcf["frame_count"], pcf["processed_count"]) == (3, 3)


@pytest.mark.parametrize("name", SEQS)
def test_golden_sequences(s3d, name):
    g = load_golden(name)
    m = s3d.SonarTo3DMapper(golden_config(g))
    for f in range(len(g["images"])):
        st = m.process_sonar_image(g["images"][f], list(g["positions"][f]), list(g["quaternions"][f]))
        assert _stats3(st) == g["stats"][f].tolist(), f"frame {f}"
        if f"ckpt{f}_keys" in g:
            k, v = m.octree.voxels.to_arrays()
            assert_same_map(k, v, g[f"ckpt{f}_keys"], g[f"ckpt{f}_logodds"], LOGODDS_ATOL, f"{name} ckpt {f}")
    keys, vals = m.octree.voxels.to_arrays()
    err = assert_same_map(keys, vals, g["keys"], g["logodds"], LOGODDS_ATOL, name)
    assert err <= 1e-9, f"log-odds drifted more than fp64 rounding explains: {err}"
    # export: same occupied set, centres bit-exact, probabilities to fp64 rounding
    pc = m.get_point_cloud()
    res = m.voxel_resolution
    kg, pg, qg = sort_by_key(np.floor(pc["points"] / res), pc["points"], pc["probabilities"])
    kr, pr, qr = sort_by_key(np.floor(g["pc_points"] / res), g["pc_points"], g["pc_prob"])
    assert np.array_equal(kg, kr) and np.array_equal(pg, pr)
    assert np.abs(qg - qr).max(initial=0.0) <= 1e-9
    pcf = m.get_point_cloud(True)
    assert [pcf["num_occupied"], pcf["num_free"], pcf["num_unknown"]] == g["counts"].tolist()
    assert len(pcf["occupied"]) == g["counts"][0]
    assert np.array_equal(m.octree.min_bounds, g["min_bounds"]) and np.array_equal(m.octree.max_bounds, g["max_bounds"])
    assert np.array_equal(m.bearing_angles, g["bearing_angles"])
    assert np.array_equal(m.T_sonar_to_base, g["T_sonar_to_base"])


def test_edge_frames(s3d):
    g = load_golden("edge_frames")
    cfg = golden_config(g)
    for name in ("nohit", "allhit", "hit_at_zero", "late_hit", "random", "equal_thr"):
        m = s3d.SonarTo3DMapper(cfg)
        st = m.process_sonar_image(g[f"{name}__image"], list(g["positions"][0]), list(g["quaternions"][0]))
        assert [_stats3(st)] == g[f"{name}__stats"].tolist(), name
        k, v = m.octree.voxels.to_arrays()
        assert_same_map(k, v, g[f"{name}__keys"], g[f"{name}__logodds"], LOGODDS_ATOL, name)
    m = s3d.SonarTo3DMapper(dict(cfg, intensity_threshold=99.5))       # float threshold, W = 1
    st = m.process_sonar_image(g["onebeam__image"], list(g["positions"][0]), list(g["quaternions"][0]))
    assert [_stats3(st)] == g["onebeam__stats"].tolist()
    k, v = m.octree.voxels.to_arrays()
    assert_same_map(k, v, g["onebeam__keys"], g["onebeam__logodds"], LOGODDS_ATOL, "onebeam")
    # empty map exports
    e = s3d.SonarTo3DMapper(cfg)
    pc = e.get_point_cloud()
    assert pc["points"].shape == (0, 3) and pc["probabilities"].shape == (0,) and pc["num_voxels"] == 0
    pcf = e.get_point_cloud(True)
    assert len(pcf["occupied"]) == len(pcf["free"]) == len(pcf["unknown"]) == 0
    assert np.isinf(pcf["bounds"]["min"]).all() and np.isinf(pcf["bounds"]["max"]).all()


def test_stage_vectors(s3d):
    """Sample count and the key multiset of one frame against process_sonar_ray / world_to_key."""
    g = load_golden("stage_vectors")
    cfg = golden_config(g)
    m = s3d.SonarTo3DMapper(cfg)
    st = m.process_sonar_image(g["image"], list(g["position"]), list(g["quaternion"]))
    assert m.last_num_samples == len(g["xyz"])
    uk, inv = np.unique(g["keys"].astype(np.int64), axis=0, return_inverse=True)
    occ_any = np.zeros(len(uk), dtype=bool)
    np.logical_or.at(occ_any, inv.reshape(-1), g["occupied"].astype(bool))
    assert st["num_occupied"] == int(occ_any.sum()) and st["num_free"] == int((~occ_any).sum())
    k, _ = m.octree.voxels.to_arrays()
    assert np.array_equal(sort_by_key(k)[0], sort_by_key(uk)[0])


def test_store_vectors(s3d):
    """SimpleOctree driven directly (scripts/3d_mapper.py:83-188)."""
    g = load_golden("store_vectors")
    oc = s3d.SimpleOctree(resolution=0.05, dynamic_expansion=True)
    oc.adaptive_max_ratio = 0.3
    n_single = 64
    for p, u, a in zip(g["points"][:n_single], g["updates"][:n_single], g["adaptive"][:n_single]):
        oc.update_voxel(p, float(u), adaptive=bool(a))
    oc.update_voxels(g["points"][n_single:], g["updates"][n_single:], g["adaptive"][n_single:])
    k, v = oc.voxels.to_arrays()
    assert_same_map(k, v, g["keys"], g["logodds"], 1e-12, "store")
    lo = np.array([oc.get_log_odds(*p) for p in g["query_points"]])
    pr = np.array([oc.get_probability(*p) for p in g["query_points"]])
    assert np.abs(lo - g["query_logodds"]).max() <= 1e-12 and np.abs(pr - g["query_prob"]).max() <= 1e-12
    assert len(oc.voxels) == len(g["keys"])                       # queries never insert
    assert np.array_equal(oc.min_bounds, g["min_bounds"]) and np.array_equal(oc.max_bounds, g["max_bounds"])
    edge = json.loads(str(g["edge_thr_json"]))
    assert len(oc.get_occupied_voxels(1.0)) == edge["p1"]
    assert len(oc.get_occupied_voxels(0.0)) == edge["p0"]
    assert len(oc.get_occupied_voxels(0.5)) == edge["p05"]
    occ = oc.get_occupied_voxels(0.6)
    pts = np.array([p for p, _ in occ]).reshape(-1, 3)
    assert np.array_equal(sort_by_key(np.floor(pts / 0.05), pts)[1], sort_by_key(np.floor(g["occ_points"] / 0.05), g["occ_points"])[1])
    cls = oc.get_all_voxels_classified(0.7)
    assert [len(cls["free"]), len(cls["unknown"]), len(cls["occupied"])] == g["cls_counts"].tolist()
    # voxels view
    key0 = tuple(int(x) for x in g["keys"][0])
    assert key0 in oc.voxels and abs(oc.voxels.get(key0) - g["logodds"][0]) <= 1e-12
    assert (9999, 9999, 9999) not in oc.voxels and oc.voxels.get((9999, 9999, 9999), 0.0) == 0.0
    assert dict(oc.voxels.items())[key0] == oc.voxels.get(key0)
    oc.clear()
    assert len(oc.voxels) == 0 and np.isinf(oc.min_bounds).all()


@pytest.mark.parametrize("cfg_name,n_frames", [("cfg1", 4), ("cfg2", 6), ("cfg3", 2)])
def test_vs_oracle_full_size(s3d, cfg_name, n_frames):
    """BASELINE.json configs at full image size against the CPU oracle on the same seeded input."""
    from oracle.oracle import OracleMapper
    from sonar_3d_reconstruction_b200 import synthetic
    images, pos, quat, cfg = synthetic.make_sequence(cfg_name, n_frames, seed=0)
    gpu, cpu = s3d.SonarTo3DMapper(cfg), OracleMapper(cfg)
    for f in range(n_frames):
        a = gpu.process_sonar_image(images[f], list(pos[f]), list(quat[f]))
        b = cpu.process_sonar_image(images[f], pos[f], quat[f])
        assert _stats3(a) == _stats3(b), f"{cfg_name} frame {f}"
        assert gpu.last_num_samples == b["num_samples"]
    kg, vg = gpu.octree.voxels.to_arrays()
    kc, vc = cpu.dump()
    err = assert_same_map(kg, vg, kc, vc, LOGODDS_ATOL, cfg_name)
    assert err <= 1e-9
    assert gpu.get_point_cloud()["num_occupied"] == cpu.get_point_cloud()["num_occupied"]


def test_batch_equals_sequential_and_growth(s3d):
    """One s3d_ingest_batch call == the same frames one call each; a tiny initial table forces
    several rehash-grows on the way and must not change anything."""
    from sonar_3d_reconstruction_b200 import synthetic
    spec = dict(H=160, W=200, config=dict(voxel_resolution=0.06, intensity_threshold=45, max_range=8.0), step_m=0.03)
    images, pos, quat, cfg = synthetic.make_sequence(spec, 40, seed=3)
    a = s3d.SonarTo3DMapper(cfg)
    b = s3d.SonarTo3DMapper(dict(cfg, table_capacity=1024))
    sa = [_stats3(a.process_sonar_image(images[f], list(pos[f]), list(quat[f]))) for f in range(len(images))]
    sb = [_stats3(s) for s in b.process_sonar_images(images, pos, quat)]
    assert sa == sb
    ka, va = a.octree.voxels.to_arrays()
    kb, vb = b.octree.voxels.to_arrays()
    assert_same_map(ka, va, kb, vb, 0.0, "batch vs sequential")
    assert b.frame_count == a.frame_count == 40


def test_async_batches_equal_the_blocking_call_and_survive_growth(s3d):
    """Two batches pending at a time (submit k+1 before collecting k), tiny initial table so that chunks
    of pending batches are re-run after rehash-grows: same counters and map as one blocking call."""
    from sonar_3d_reconstruction_b200 import synthetic
    spec = dict(H=160, W=200, config=dict(voxel_resolution=0.06, intensity_threshold=45, max_range=8.0), step_m=0.03)
    images, pos, quat, cfg = synthetic.make_sequence(spec, 150, seed=8)
    a = s3d.SonarTo3DMapper(cfg)
    sa = [_stats3(x) for x in a.process_sonar_images(images, pos, quat)]
    b = s3d.SonarTo3DMapper(dict(cfg, table_capacity=1024))
    sb, pending = [], None
    for f0 in range(0, 150, 37):                                   # ragged batch sizes
        h = b.process_sonar_images_async(images[f0:f0 + 37], pos[f0:f0 + 37], quat[f0:f0 + 37])
        if pending is not None:
            sb += [_stats3(x) for x in pending.result()]
        pending = h
    sb += [_stats3(x) for x in pending.result()]
    assert sa == sb
    assert_same_map(*a.octree.voxels.to_arrays(), *b.octree.voxels.to_arrays(), 0.0, "async vs blocking")
    assert b.frame_count == 150


def test_dump_load_roundtrip_and_xyzi32(s3d):
    from sonar_3d_reconstruction_b200 import synthetic
    images, pos, quat, cfg = synthetic.make_sequence(dict(H=120, W=96, config=dict(voxel_resolution=0.1)), 5, seed=9)
    m = s3d.SonarTo3DMapper(cfg)
    m.process_sonar_images(images, pos, quat)
    k, v = m.octree.voxels.to_arrays()
    other = s3d.SimpleOctree(resolution=0.1)
    other._push_params()
    other._native.load(k, v)
    k2, v2 = other.voxels.to_arrays()
    assert_same_map(k, v, k2, v2, 0.0, "dump -> load")
    pc = m.get_point_cloud()
    blob = m.get_point_cloud_xyzi32()
    assert blob.dtype == np.float32 and blob.shape == (pc["num_occupied"], 4)
    want = np.concatenate([pc["points"], pc["probabilities"][:, None]], axis=1).astype(np.float32)
    assert np.array_equal(blob[np.lexsort(blob.T[::-1])], want[np.lexsort(want.T[::-1])])
    m.reset_map()
    assert m.frame_count == 0 and m.get_point_cloud()["num_voxels"] == 0


def _oracle_parity(s3d, images, pos, quat, cfg, what, gpu_cfg=None):
    from oracle.oracle import OracleMapper
    gpu, cpu = s3d.SonarTo3DMapper(dict(cfg, **(gpu_cfg or {}))), OracleMapper(cfg)
    for f in range(len(images)):
        a = gpu.process_sonar_image(images[f], list(pos[f]), list(quat[f]))
        b = cpu.process_sonar_image(images[f], pos[f], quat[f])
        assert _stats3(a) == _stats3(b), f"{what} frame {f}"
        assert gpu.last_num_samples == b["num_samples"]
    kg, vg = gpu.octree.voxels.to_arrays()
    kc, vc = cpu.dump()
    assert assert_same_map(kg, vg, kc, vc, LOGODDS_ATOL, what) <= 1e-9
    return gpu


def test_wide_counter_lanes(s3d, monkeypatch):
    """The 32+32-bit counter-lane format (the fallback of the 16+16-bit one) gives the same map."""
    from sonar_3d_reconstruction_b200 import synthetic
    monkeypatch.setenv("S3D_WIDE_LANES", "1")
    spec = dict(H=160, W=200, config=dict(voxel_resolution=0.06, intensity_threshold=45, max_range=8.0), step_m=0.03)
    images, pos, quat, cfg = synthetic.make_sequence(spec, 20, seed=11)
    _oracle_parity(s3d, images, pos, quat, cfg, "wide lanes")


def test_sample_count_beyond_16_bits_switches_to_wide_lanes(s3d):
    """511 beams x 50 bins x 5 vertical steps = 127750 occupied samples of one frame in ONE voxel
    (10 m voxels around a 2 m sonar): the 16-bit count overflows, the chunk is re-run with wide
    lanes, and the result still equals the reference's."""
    H, W = 100, 511
    img = np.zeros((H, W), dtype=np.uint8)
    img[10:, :] = 255
    cfg = dict(voxel_resolution=10.0, max_range=2.0, min_range=0.1, intensity_threshold=35)
    images = np.stack([img, img, img])
    pos = np.array([[5.0, 5.0, 5.0], [5.01, 5.0, 5.0], [5.02, 5.0, 5.0]])
    quat = np.array([[0.0, 0.0, 0.0, 1.0]] * 3)
    gpu = _oracle_parity(s3d, images, pos, quat, cfg, "count overflow")
    assert gpu.last_num_samples > 65535


def test_samples_far_from_the_sonar_origin(s3d):
    """Fine voxels and long range: samples more than 512 voxels from the sonar origin do not fit
    the block combiner's 10-bit local coordinates and take the direct path to the dedupe table."""
    from sonar_3d_reconstruction_b200 import synthetic
    spec = dict(H=120, W=48, seabed_depth=8.5, config=dict(voxel_resolution=0.012, intensity_threshold=40, max_range=10.0,
                                                            vertical_aperture=6.0), step_m=0.02)
    images, pos, quat, cfg = synthetic.make_sequence(spec, 3, seed=13)
    _oracle_parity(s3d, images, pos, quat, cfg, "far samples")


def test_mono16_frames_take_the_high_byte_on_the_device(s3d):
    """16-bit frames: same map as the node's (img / 256).astype(uint8) followed by the 8-bit path."""
    from sonar_3d_reconstruction_b200 import synthetic
    spec = dict(H=150, W=130, config=dict(voxel_resolution=0.07, intensity_threshold=45, max_range=8.0), step_m=0.03)
    images, pos, quat, cfg = synthetic.make_sequence(spec, 19, seed=21)
    rng = np.random.default_rng(5)
    img16 = (images.astype(np.uint16) << 8) | rng.integers(0, 256, size=images.shape, dtype=np.uint16)
    assert np.array_equal((img16 / 256).astype(np.uint8), images)
    a, b = s3d.SonarTo3DMapper(cfg), s3d.SonarTo3DMapper(cfg)
    sa = [_stats3(x) for x in a.process_sonar_images(images, pos, quat)]
    sb = [_stats3(x) for x in b.process_sonar_images_mono16(img16, pos, quat)]
    assert sa == sb
    assert_same_map(*a.octree.voxels.to_arrays(), *b.octree.voxels.to_arrays(), 0.0, "mono16 vs 8-bit")


@pytest.mark.parametrize("mode", ["replicate", "route"])
def test_sharded_single_rank_equals_plain(s3d, mode):
    """The sharded code paths (route: expand -> pack by owner -> merge -> apply) with world = 1
    must reproduce the plain mapper exactly; the 2-rank exchange is covered on CPU (gloo) and by
    tools/sharded_check.py under torchrun on 2+ GPUs."""
    from sonar_3d_reconstruction_b200 import synthetic
    from sonar_3d_reconstruction_b200.sharded import ShardedSonarMapper
    spec = dict(H=160, W=200, config=dict(voxel_resolution=0.06, intensity_threshold=45, max_range=8.0), step_m=0.03)
    images, pos, quat, cfg = synthetic.make_sequence(spec, 37, seed=4)
    a = s3d.SonarTo3DMapper(cfg)
    sa = [_stats3(s) for s in a.process_sonar_images(images, pos, quat)]
    b = ShardedSonarMapper(dict(cfg, table_capacity=4096), group=None, mode=mode)
    sb = [_stats3(s) for s in b.process_sonar_images(images, pos, quat)]
    assert sa == sb
    ka, va = a.octree.voxels.to_arrays()
    kb, vb = b.gather_map()
    assert_same_map(ka, va, kb, vb, 0.0, "sharded(world=1) vs plain")
    assert b.get_point_cloud()["num_occupied"] == a.get_point_cloud()["num_occupied"]
    assert b.num_voxels() == len(ka)


@pytest.mark.parametrize("world", [2, 3])
def test_routed_map_ranks_in_one_process(world):
    """The fused exchange (records written into the owner's inbox by the expansion kernel,
    device-side sequence flags, owner-side merge) with two / three ranks on this one GPU (53 frames:
    three full chunks and a ragged one); multi-process runs over CUDA IPC are exercised by
    tools/sharded_check.py under torchrun."""
    import subprocess
    import sys
    root = os.path.dirname(GOLDEN.rstrip("/").rsplit("/", 1)[0])
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32", S3D_ROUTE_TIMEOUT_MS="20000", S3D_LOCAL_WORLD=str(world))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "route_local_check.py")], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "route_local_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_cfg1_100_frames_single_and_batched_calls(s3d):
    """north_star's parity target: BASELINE configs[0] (512 beams x 500 bins, library defaults) x 100
    posed frames -- once through process_sonar_image (the reference's own call, one frame per
    C-ABI call) and once through process_sonar_images (16-frame chunks, three in flight) -- against
    the CPU oracle: per-frame counters exact, key set exact, |dL| <= 1e-5 (observed ~1e-15).
    Spec: scripts/3d_mapper.py:485-595."""
    from oracle.oracle import OracleMapper
    from sonar_3d_reconstruction_b200 import synthetic
    n = 100
    images, pos, quat, cfg = synthetic.make_sequence("cfg1", n, seed=0)
    cpu = OracleMapper(cfg)
    want = [cpu.process_sonar_image(images[f], pos[f], quat[f]) for f in range(n)]
    kc, vc = cpu.dump()
    one = s3d.SonarTo3DMapper(cfg)
    for f in range(n):
        a = one.process_sonar_image(images[f], list(pos[f]), list(quat[f]))
        assert _stats3(a) == _stats3(want[f]), f"single-frame call, frame {f}"
        assert one.last_num_samples == want[f]["num_samples"]
    assert assert_same_map(*one.octree.voxels.to_arrays(), kc, vc, LOGODDS_ATOL, "cfg1 x 100, single-frame calls") <= 1e-9
    batch = s3d.SonarTo3DMapper(cfg)
    got = batch.process_sonar_images(images, pos, quat)
    assert [_stats3(x) for x in got] == [_stats3(x) for x in want]
    assert assert_same_map(*batch.octree.voxels.to_arrays(), kc, vc, LOGODDS_ATOL, "cfg1 x 100, batched call") <= 1e-9
    assert batch.get_point_cloud()["num_occupied"] == cpu.get_point_cloud()["num_occupied"]


def test_retry_of_a_later_chunk_leaves_earlier_chunks_whole(s3d, monkeypatch):
    """A chunk that asks for a retry (here: its dedupe table, forced tiny, overflows during k_expand) must
    not stop the chunks before it, which are still being applied on the other streams: chunk 0 is small
    and fits, chunks 1..3 overflow the 4096-entry table (twice: 4096 -> 8192 -> 16384) while chunk 0 is
    in flight.  Counters and map must still equal the oracle's."""
    from oracle.oracle import OracleMapper
    from sonar_3d_reconstruction_b200 import synthetic
    monkeypatch.setenv("S3D_SCRATCH_CAP", "4096")
    spec = dict(H=200, W=128, seabed_depth=3.0, config=dict(voxel_resolution=0.1, intensity_threshold=40), step_m=0.05)
    from sonar_3d_reconstruction_b200._native import NativeMap
    cf = NativeMap.CHUNK_FRAMES
    images, pos, quat, cfg = synthetic.make_sequence(spec, 4 * cf, seed=17)
    images = images.copy()
    images[:cf] = 0
    images[:cf, 12, :] = 255                        # chunk 0: first hit at 0.6 m, a few hundred voxels per frame
    for rep in range(3):                            # the race is timing dependent: a few tries
        gpu, cpu = s3d.SonarTo3DMapper(cfg), OracleMapper(cfg)
        got = gpu.process_sonar_images(images, pos, quat)
        want = [cpu.process_sonar_image(images[f], pos[f], quat[f]) for f in range(len(images))]
        assert [_stats3(x) for x in got] == [_stats3(x) for x in want], f"try {rep}"
        assert assert_same_map(*gpu.octree.voxels.to_arrays(), *cpu.dump(), LOGODDS_ATOL, "retry") <= 1e-9
        # the same through two pending asynchronous batches
        gpu2 = s3d.SonarTo3DMapper(cfg)
        cut = cf * 2 + 8
        h1 = gpu2.process_sonar_images_async(images[:cut], pos[:cut], quat[:cut])
        h2 = gpu2.process_sonar_images_async(images[cut:], pos[cut:], quat[cut:])
        got2 = h1.result() + h2.result()
        assert [_stats3(x) for x in got2] == [_stats3(x) for x in want], f"async try {rep}"
        assert assert_same_map(*gpu2.octree.voxels.to_arrays(), *cpu.dump(), LOGODDS_ATOL, "retry async") <= 1e-9


def test_debug_counters_reproduce_the_reference_stdout(s3d, capsys):
    """with debug_counters=True the every-10th-frame [DEBUG] block (scripts/3d_mapper.py:574-585),
    "Map reset" (:650) and the two debug dicts equal the unmodified reference's, frame by frame and as one batch."""
    g = load_golden("debug_counters")
    cfg = dict(golden_config(g), debug_counters=True)
    cut = int(g["reset_after"])
    img, pos, quat = g["images"], g["positions"], g["quaternions"]
    m = s3d.SonarTo3DMapper(cfg)
    capsys.readouterr()
    for f in range(cut):
        m.process_sonar_image(img[f], list(pos[f]), list(quat[f]))
    m.reset_map()
    for f in range(cut, len(img)):
        m.process_sonar_image(img[f], list(pos[f]), list(quat[f]))
    assert capsys.readouterr().out == str(g["stdout"])
    want_frame = dict(zip(map(tuple, g["frame_keys"].tolist()), g["frame_counts"].tolist()))
    want_total = dict(zip(map(tuple, g["total_keys"].tolist()), g["total_counts"].tolist()))
    assert dict(m.frame_update_counts) == want_frame
    assert dict(m.voxel_update_counts) == want_total
    assert max(m.voxel_update_counts.values()) == max(want_total.values()) and len(m.frame_update_counts) == len(want_frame)
    # the batched call prints the same text
    b = s3d.SonarTo3DMapper(cfg)
    b.process_sonar_images(img[:cut], pos[:cut], quat[:cut])
    b.reset_map()
    b.process_sonar_images(img[cut:], pos[cut:], quat[cut:])
    assert capsys.readouterr().out == str(g["stdout"])
    assert dict(b.voxel_update_counts) == want_total and dict(b.frame_update_counts) == want_frame
    # off by default: silent, dicts empty
    q = s3d.SonarTo3DMapper(golden_config(g))
    for f in range(10):
        q.process_sonar_image(img[f], list(pos[f]), list(quat[f]))
    assert capsys.readouterr().out == "" and len(q.voxel_update_counts) == 0 and not q.frame_update_counts


def test_api_fidelity_lists_bounds_and_float_images(s3d):
    """Return types and attribute behaviour of the reference: exports are real lists (:151, :178-182),
    min/max_bounds are assignable (:38-40) and extended by later updates (:113-115), update_voxel bounds
    follow the raw point, non-uint8 images are thresholded like uint8 ones (:407)."""
    from sonar_3d_reconstruction_b200 import synthetic
    images, pos, quat, cfg = synthetic.make_sequence(dict(H=120, W=96, config=dict(voxel_resolution=0.1)), 3, seed=2)
    m = s3d.SonarTo3DMapper(cfg)
    m.process_sonar_images(images[:2], pos[:2], quat[:2])
    occ = m.octree.get_occupied_voxels(0.6)
    assert type(occ + occ) is list and len(occ + occ) == 2 * len(occ)
    occ.append((np.zeros(3), 0.5))
    p0, q0 = occ[0]
    assert isinstance(p0, np.ndarray) and p0.shape == (3,) and isinstance(q0, float)
    cls = m.octree.get_all_voxels_classified(0.7)
    assert isinstance(cls["free"], list) and isinstance(cls["unknown"] + cls["occupied"], list)
    # bounds: assign, then extend
    mn0, mx0 = m.octree.min_bounds.copy(), m.octree.max_bounds.copy()
    m.octree.min_bounds = np.array([1e6, 1e6, 1e6])
    m.octree.max_bounds = np.array([-1e6, -1e6, -1e6])
    assert np.array_equal(m.octree.min_bounds, [1e6] * 3) and np.array_equal(m.octree.max_bounds, [-1e6] * 3)
    m.octree.update_voxel(np.array([0.06, 0.06, 0.06]), 0.4)
    assert np.array_equal(m.octree.max_bounds, [0.06] * 3) and np.array_equal(m.octree.min_bounds, [0.06] * 3)
    m.process_sonar_image(images[2], list(pos[2]), list(quat[2]))
    only = s3d.SonarTo3DMapper(cfg)
    only.process_sonar_image(images[2], list(pos[2]), list(quat[2]))
    assert np.array_equal(m.octree.min_bounds, np.minimum(0.06, only.octree.min_bounds))
    assert np.array_equal(m.octree.max_bounds, np.maximum(0.06, only.octree.max_bounds))
    assert (mn0 < mx0).all()
    # float image == the uint8 image it came from
    a, b = s3d.SonarTo3DMapper(cfg), s3d.SonarTo3DMapper(cfg)
    sa = a.process_sonar_image(images[0], list(pos[0]), list(quat[0]))
    sb = b.process_sonar_image(images[0].astype(np.float32) + 0.25, list(pos[0]), list(quat[0]))
    assert _stats3(sa) == _stats3(sb)
    assert_same_map(*a.octree.voxels.to_arrays(), *b.octree.voxels.to_arrays(), 0.0, "float image")


@pytest.mark.parametrize("cfg_name,n_frames", [("cfg1", 6), ("cfg2", 8), ("cfg3", 2)])
def test_fp32_estimate_never_disagrees_with_the_fp64_key(s3d, monkeypatch, cfg_name, n_frames):
    """k_expand accepts the fp32 voxel-index estimate only outside a proven error band; with
    S3D_VERIFY_FAST=1 every sample also takes the reference's fp64 arithmetic and an accepted estimate
    that differs raises an error (millions of samples per config).  The same frames without TMA
    staging and without the fp32 path must give the same map."""
    from oracle.oracle import OracleMapper
    from sonar_3d_reconstruction_b200 import synthetic
    images, pos, quat, cfg = synthetic.make_sequence(cfg_name, n_frames, seed=5)
    monkeypatch.setenv("S3D_VERIFY_FAST", "1")
    v = s3d.SonarTo3DMapper(cfg)
    sv = [_stats3(x) for x in v.process_sonar_images(images, pos, quat)]     # raises on a mismatch
    monkeypatch.delenv("S3D_VERIFY_FAST")
    cpu = OracleMapper(cfg)
    want = [_stats3(cpu.process_sonar_image(images[f], pos[f], quat[f])) for f in range(n_frames)]
    assert sv == want
    assert assert_same_map(*v.octree.voxels.to_arrays(), *cpu.dump(), LOGODDS_ATOL, cfg_name) <= 1e-9
    monkeypatch.setenv("S3D_NO_TMA", "1")
    monkeypatch.setenv("S3D_NO_FAST32", "1")
    plain = s3d.SonarTo3DMapper(cfg)
    assert [_stats3(x) for x in plain.process_sonar_images(images, pos, quat)] == want
    assert_same_map(*v.octree.voxels.to_arrays(), *plain.octree.voxels.to_arrays(), 0.0, "fp64-only, plain loads")


def test_many_short_calls_equal_one_call(s3d):
    """Streaming in calls that are not a multiple of the chunk size (24 frames: a full chunk and a half-full one,
    whose dedupe table is sparse) must give what one call gives -- per-frame counters and the whole map.  A few
    hundred chunks back to back with nothing drained in between: the shape that exposed a block-level race in the
    update kernel (tools/stress_growth.py is the long version)."""
    import torch
    from sonar_3d_reconstruction_b200 import synthetic
    n, step = 2400, 24
    base, pos, quat, cfg = synthetic.make_sequence("cfg2", n, seed=1, distinct_images=250, cycle=False)
    d_img = torch.from_numpy(np.ascontiguousarray(base)).cuda()[torch.arange(n, device="cuda") % len(base)]
    H, W = base.shape[1:]
    maps, stats = [], []
    for calls in (step, n, step):
        m = s3d.SonarTo3DMapper(cfg)
        m._check_width(W); m._sync_device_config(H, W)
        nat = m.octree._native
        d_T = torch.from_numpy(np.ascontiguousarray(m.compose_transforms(pos, quat).reshape(n, 16))).cuda()
        st = torch.zeros((n, 8), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        for f0 in range(0, n, calls):
            k = min(calls, n - f0)
            nat.ingest_batch_dev(d_img.data_ptr() + f0 * H * W, k, d_T.data_ptr() + f0 * 128, want_stats=False,
                                 stats_dev_ptr=st.data_ptr() + f0 * 64)
        nat.sync()
        stats.append(st.cpu().numpy()[:, :4])
        maps.append(sort_by_key(*m.octree.voxels.to_arrays()))
        m.close()
    for i in (0, 2):
        assert np.array_equal(stats[i], stats[1]), f"per-frame counters, run {i}"
        assert np.array_equal(maps[i][0], maps[1][0]) and np.array_equal(maps[i][1], maps[1][1]), f"map, run {i}"
