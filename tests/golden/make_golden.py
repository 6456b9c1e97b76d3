#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden.py

The reference module is loaded by file path exactly as its own ROS2 node does
(scripts/3d_mapper_node.py:33-42).  Every fixture stores its inputs (images, poses,
config) next to the reference's outputs, so the tests do not depend on RNG stability.
The reference has no tests or golden files of its own; these
recorded outputs are the parity pin of both the CPU oracle and the CUDA path.
"""
import contextlib
import hashlib
import importlib.util
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from sonar_3d_reconstruction_b200 import synthetic  # noqa: E402

REF_PATH = os.environ.get("S3D_REFERENCE", "/root/reference/scripts/3d_mapper.py")


def load_reference():
    spec = importlib.util.spec_from_file_location("reference_3d_mapper", REF_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def dump_octree(octree):
    keys = np.array(list(octree.voxels.keys()), dtype=np.int64).reshape(-1, 3)
    vals = np.array([float(v) for v in octree.voxels.values()], dtype=np.float64)
    return keys, vals


def run_sequence(ref, config, images, pos, quat, checkpoints=()):
    """Feed a posed sequence through the reference; record everything a test compares."""
    mapper = ref.SonarTo3DMapper(dict(config))
    stats, ckpt = [], {}
    for f in range(len(images)):
        st = quiet(mapper.process_sonar_image, images[f], list(pos[f]), list(quat[f]))
        stats.append([st["num_occupied"], st["num_free"], st["num_voxels"]])
        if f in checkpoints:
            k, v = dump_octree(mapper.octree)
            ckpt[f"ckpt{f}_keys"] = k.astype(np.int32)
            ckpt[f"ckpt{f}_logodds"] = v
    keys, vals = dump_octree(mapper.octree)
    pc = mapper.get_point_cloud(False)
    pcf = mapper.get_point_cloud(True)
    out = dict(
        images=np.asarray(images), positions=np.asarray(pos, dtype=np.float64),
        quaternions=np.asarray(quat, dtype=np.float64),
        config_json=np.array(json.dumps(config)),
        stats=np.array(stats, dtype=np.int64),
        keys=keys.astype(np.int32), logodds=vals,      # dict insertion order
        pc_points=np.asarray(pc["points"], dtype=np.float64).reshape(-1, 3),
        pc_prob=np.asarray(pc["probabilities"], dtype=np.float64),
        counts=np.array([pcf["num_occupied"], pcf["num_free"], pcf["num_unknown"]], dtype=np.int64),
        min_bounds=np.asarray(mapper.octree.min_bounds, dtype=np.float64),
        max_bounds=np.asarray(mapper.octree.max_bounds, dtype=np.float64),
        T_sonar_to_base=np.asarray(mapper.T_sonar_to_base, dtype=np.float64),
        bearing_angles=np.asarray(mapper.bearing_angles, dtype=np.float64),
    )
    out.update(ckpt)
    return out


def seq_inputs(H, W, n, seed, cfg, step_m=0.05, spike_prob=1e-3, seabed_depth=4.0):
    rng = np.random.default_rng(seed)
    images = np.stack([synthetic.make_frame(rng, H, W, fov_deg=cfg.get("horizontal_fov", 130.0),
                                            max_range=cfg.get("max_range", 10.0),
                                            threshold=cfg.get("intensity_threshold", 35),
                                            seabed_depth=seabed_depth, spike_prob=spike_prob)
                       for _ in range(n)])
    pos, quat = synthetic.make_poses(rng, n, step_m=step_m)
    return images, pos, quat


def case_selftest(ref):
    """The reference's own __main__ sequence (scripts/3d_mapper.py:653-683)."""
    cfg = {"voxel_resolution": 0.1, "min_probability": 0.6, "intensity_threshold": 30}
    img = np.zeros((500, 512), dtype=np.uint8)
    img[100:150, 200:300] = 100
    img[300:350, 100:150] = 150
    images = np.stack([img] * 3)
    pos = np.array([[i * 0.1, 0.0, 0.0] for i in range(3)])
    quat = np.array([[0.0, 0.0, 0.0, 1.0]] * 3)
    out = run_sequence(ref, cfg, images, pos, quat)
    keys, vals = out["keys"].astype(np.int64), out["logodds"]
    order = np.lexsort(keys.T[::-1])
    summary = dict(
        stats=out["stats"].tolist(), counts=out["counts"].tolist(),
        num_occupied=int(len(out["pc_prob"])),
        sum_logodds=float(vals.sum()), min_logodds=float(vals.min()), max_logodds=float(vals.max()),
        distinct_logodds=int(len(np.unique(vals))),
        key_min=keys.min(0).tolist(), key_max=keys.max(0).tolist(),
        min_bounds=out["min_bounds"].tolist(), max_bounds=out["max_bounds"].tolist(),
        sha_keys=hashlib.sha256(keys[order].tobytes()).hexdigest()[:16],
        sha_vals=hashlib.sha256(vals[order].tobytes()).hexdigest()[:16],
    )
    with open(os.path.join(HERE, "selftest_known_answers.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print("selftest", summary["stats"], summary["sha_keys"], summary["sha_vals"])


def case_sequences(ref):
    kiro = dict(synthetic.CONFIGS["cfg2"]["config"])
    kiro["voxel_resolution"] = 0.15           # the YAML exactly as shipped
    cases = {
        # library defaults, full M750D frame (configs[0] shape), 2 frames
        "seq_cfg1_default": (dict(), (500, 512, 2, 0), {}),
        # KIRO YAML as shipped: z-filter, pitch 60 deg, 70 deg FOV, +0.5/-0.1, clamp 7
        "seq_kiro_yaml": (kiro, (500, 512, 4, 1), {}),
        # small frame, all beams processed (W < 256), adaptive off, coarse voxels
        "seq_small_noadapt": (dict(voxel_resolution=0.1, adaptive_update=False, intensity_threshold=50,
                                   log_odds_occupied=0.7, log_odds_free=-0.4, log_odds_max=3.5,
                                   log_odds_min=-2.0),
                              (120, 96, 8, 2), dict(step_m=0.02)),
        # wide image (bearing step 4), narrow aperture, other thresholds, long overlap
        "seq_wide_step4": (dict(voxel_resolution=0.08, intensity_threshold=60, min_range=0.8,
                                vertical_aperture=12.0, horizontal_fov=90.0, max_range=8.0,
                                adaptive_threshold=0.4, adaptive_max_ratio=0.5,
                                sonar_position=[0.2, -0.1, -0.3], sonar_orientation=[0.05, 1.2, -0.1]),
                           (300, 1024, 3, 3), {}),
        # many overlapping frames on a small image: exercises clamping and the adaptive branch
        "seq_overlap_clamp": (dict(voxel_resolution=0.12, intensity_threshold=40, log_odds_max=2.0,
                                   log_odds_min=-3.0, z_filter_enabled=True, z_filter_min=-4.2),
                              (150, 64, 30, 4), dict(step_m=0.01)),
    }
    for name, (cfg, (H, W, n, seed), kw) in cases.items():
        full = dict(cfg)
        images, pos, quat = seq_inputs(H, W, n, seed, full, **kw)
        ck = (0,) if n > 1 else ()
        out = run_sequence(ref, full, images, pos, quat, checkpoints=ck)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, out["stats"][-1].tolist(), out["counts"].tolist())


def case_edges(ref):
    """Edge images: no hit anywhere, hit everywhere, hit at bin 0, H not a multiple of 10,
    W = 1, a non-normalised quaternion, float threshold."""
    rng = np.random.default_rng(7)
    H, W = 95, 40
    base_cfg = dict(voxel_resolution=0.1, max_range=6.0, min_range=0.3, intensity_threshold=35)
    imgs = {
        "nohit": np.full((H, W), 10, np.uint8),
        "allhit": np.full((H, W), 200, np.uint8),
        "hit_at_zero": np.where(np.arange(H)[:, None] == 0, 255, 0).astype(np.uint8) * np.ones((1, W), np.uint8),
        "late_hit": np.where(np.arange(H)[:, None] >= H - 3, 90, 5).astype(np.uint8) * np.ones((1, W), np.uint8),
        "random": rng.integers(0, 256, size=(H, W)).astype(np.uint8),
        "equal_thr": np.full((H, W), 35, np.uint8),
    }
    out = {}
    pos = np.array([[0.013, -0.021, 0.007]])
    quat = np.array([[0.02, -0.03, 0.5, 0.8]])        # deliberately not unit length
    for name, img in imgs.items():
        r = run_sequence(ref, base_cfg, img[None], pos, quat)
        for k in ("stats", "keys", "logodds", "counts", "min_bounds", "max_bounds"):
            out[f"{name}__{k}"] = r[k]
        out[f"{name}__image"] = img
    one = rng.integers(0, 256, size=(60, 1)).astype(np.uint8)
    r = run_sequence(ref, dict(base_cfg, intensity_threshold=99.5), one[None], pos, quat)
    for k in ("stats", "keys", "logodds", "counts"):
        out[f"onebeam__{k}"] = r[k]
    out["onebeam__image"] = one
    out["positions"] = pos
    out["quaternions"] = quat
    out["config_json"] = np.array(json.dumps(base_cfg))
    np.savez_compressed(os.path.join(HERE, "edge_frames.npz"), **out)
    print("edges", {k: out[k].tolist() for k in out if k.endswith("__stats")})


def case_stages(ref):
    """Per-stage vectors from the reference's own functions: first hit per beam, the
    sample list of process_sonar_ray (world xyz + type), world_to_key of every sample."""
    cfg = dict(voxel_resolution=0.07, intensity_threshold=45, min_range=0.4, max_range=7.0,
               z_filter_enabled=True, z_filter_min=-3.9)
    H, W = 140, 80
    cfg.update(image_width=W, image_height=H)   # bearing table is built for image_width (:295-299)
    images, pos, quat = seq_inputs(H, W, 1, 11, cfg, spike_prob=3e-3, seabed_depth=2.5)
    img = images[0]
    m = ref.SonarTo3DMapper(dict(cfg))
    T = m.create_odometry_transform(list(pos[0]), list(quat[0])) @ m.T_sonar_to_base
    first_hits, xyz, occ, beam_of = [], [], [], []
    step = max(1, W // 256)
    for b in range(0, W, step):
        prof = img[:, b]
        fh = -1
        for r_idx, v in enumerate(prof):
            if v > m.intensity_threshold:
                fh = r_idx
                break
        first_hits.append(fh)
        for point, lo, typ in m.process_sonar_ray(m.bearing_angles[b], prof, T):
            xyz.append(np.array(point, dtype=np.float64))
            occ.append(1 if typ == "occupied" else 0)
            beam_of.append(b)
    xyz = np.array(xyz).reshape(-1, 3)
    keys = np.array([m.octree.world_to_key(*p) for p in xyz], dtype=np.int64).reshape(-1, 3)
    np.savez_compressed(os.path.join(HERE, "stage_vectors.npz"), image=img, position=pos[0], quaternion=quat[0],
                        T=np.asarray(T), config_json=np.array(json.dumps(cfg)),
                        first_hits=np.array(first_hits, dtype=np.int32), xyz=xyz,
                        occupied=np.array(occ, dtype=np.int8), beam=np.array(beam_of, dtype=np.int32),
                        keys=keys.astype(np.int32))
    print("stages", len(xyz), "samples")


def case_store(ref):
    """SimpleOctree driven directly (scripts/3d_mapper.py:19-194): scripted update_voxel
    calls with adaptive on/off, clamping at both ends, queries of present/absent voxels."""
    rng = np.random.default_rng(5)
    oc = ref.SimpleOctree(resolution=0.05, dynamic_expansion=True)
    oc.adaptive_max_ratio = 0.3
    n = 4000
    pts = rng.uniform(-1.0, 1.0, size=(n, 3)) * np.array([0.6, 0.6, 0.3])
    upd = rng.choice([1.5, -2.0, 0.45, -0.1, 3.7, -6.0, 0.0], size=n)
    adp = rng.integers(0, 2, size=n).astype(np.int8)
    for i in range(n):
        oc.update_voxel(pts[i], float(upd[i]), adaptive=bool(adp[i]))
    keys, vals = dump_octree(oc)
    q = rng.uniform(-1.0, 1.0, size=(200, 3)) * np.array([0.7, 0.7, 0.4])
    q_lo = np.array([oc.get_log_odds(*p) for p in q])
    q_pr = np.array([oc.get_probability(*p) for p in q])
    occ = oc.get_occupied_voxels(0.6)
    cls = oc.get_all_voxels_classified(0.7)
    edge_thr = dict(p1=len(oc.get_occupied_voxels(1.0)), p0=len(oc.get_occupied_voxels(0.0)),
                    p05=len(oc.get_occupied_voxels(0.5)))
    np.savez_compressed(os.path.join(HERE, "store_vectors.npz"), points=pts, updates=upd, adaptive=adp,
                        keys=keys.astype(np.int32), logodds=vals, query_points=q, query_logodds=q_lo,
                        query_prob=q_pr, occ_points=np.array([o[0] for o in occ]).reshape(-1, 3),
                        occ_prob=np.array([o[1] for o in occ]),
                        cls_counts=np.array([len(cls["free"]), len(cls["unknown"]), len(cls["occupied"])]),
                        edge_thr_json=np.array(json.dumps(edge_thr)),
                        min_bounds=oc.min_bounds, max_bounds=oc.max_bounds)
    print("store", len(keys), "voxels", edge_thr)


def case_debug(ref):
    """stdout of the reference ([DEBUG] block every 10th frame, :574-585; "Map reset", :650) and its two
    debug dicts over 25 frames + reset_map + 12 frames (voxel_update_counts survives the reset)."""
    cfg = dict(voxel_resolution=0.25, intensity_threshold=45, max_range=8.0)
    images, pos, quat = seq_inputs(120, 96, 37, 31, cfg, step_m=0.04)
    mapper = ref.SonarTo3DMapper(dict(cfg))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        for f in range(25):
            mapper.process_sonar_image(images[f], list(pos[f]), list(quat[f]))
        mapper.reset_map()
        for f in range(25, 37):
            mapper.process_sonar_image(images[f], list(pos[f]), list(quat[f]))
    fk = np.array(list(mapper.frame_update_counts.keys()), dtype=np.int32).reshape(-1, 3)
    fv = np.array(list(mapper.frame_update_counts.values()), dtype=np.int64)
    vk = np.array(list(mapper.voxel_update_counts.keys()), dtype=np.int32).reshape(-1, 3)
    vv = np.array(list(mapper.voxel_update_counts.values()), dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "debug_counters.npz"), images=images, positions=pos, quaternions=quat,
                        config_json=np.array(json.dumps(cfg)), stdout=np.array(buf.getvalue()),
                        frame_keys=fk, frame_counts=fv, total_keys=vk, total_counts=vv, reset_after=np.array(25))
    print("debug", len(buf.getvalue().splitlines()), "stdout lines;", len(fk), "frame keys,", len(vk), "total keys")


def case_node(ref):
    """Transcripts of the fake node (tests/fake_node.py: the node's exact call sequence) on the reference:
    one session publishing PointCloud2, one publishing the classified CUBE_LIST markers."""
    sys.path.insert(0, os.path.dirname(HERE))
    import fake_node
    cls = fake_node.load_mapper_class(REF_PATH)
    cfg = fake_node.node_config(fake_node.NODE_PARAMS)
    images, pos, quat = seq_inputs(160, 128, 23, 41, cfg, step_m=0.05)
    rng = np.random.default_rng(7)
    images16 = {f: (images[f].astype(np.uint16) << 8) | rng.integers(0, 256, size=images[f].shape, dtype=np.uint16)
                for f in (3, 11)}
    out = dict(images=images, positions=pos, quaternions=quat,
               frames16=np.array(sorted(images16)), images16=np.stack([images16[f] for f in sorted(images16)]))
    for name, free in (("cloud", False), ("markers", True)):
        tr = quiet(fake_node.run_session, cls, images, images16, pos, quat, free)
        out[f"{name}__log"] = np.array(json.dumps(tr["log"]))
        out[f"{name}__n_clouds"] = np.array(len(tr["clouds"]))
        for i, c in enumerate(tr["clouds"]):
            out[f"{name}__cloud{i}"] = c
        out[f"{name}__n_markers"] = np.array(len(tr["markers"]))
        for i, m in enumerate(tr["markers"]):
            out[f"{name}__marker{i}_meta"] = np.array(json.dumps({k: {"scale": v["scale"], "rgba": v["rgba"]} for k, v in m.items()}))
            for k, v in m.items():
                out[f"{name}__marker{i}_{k}"] = v["points"]
        print("node", name, tr["log"][-1])
    np.savez_compressed(os.path.join(HERE, "node_sessions.npz"), **out)


if __name__ == "__main__":
    ref = load_reference()
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            globals()["case_" + name](ref)
        sys.exit(0)
    case_selftest(ref)
    case_stages(ref)
    case_store(ref)
    case_edges(ref)
    case_sequences(ref)
    case_debug(ref)
    case_node(ref)
    print("numpy", np.__version__)
