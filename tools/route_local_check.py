#!/usr/bin/env python3
"""Routed (fused-exchange) map with TWO ranks inside ONE process on ONE GPU: both maps live on
device 0, attach each other's exchange block by pointer, and are fed chunk by chunk in turn.
The union of the two shards must equal the plain single-map result bit for bit.  Needs more
hardware queues than the default 8 (a waiting kernel of one map must never sit in front of the
other map's work): run with CUDA_DEVICE_MAX_CONNECTIONS=32 (tests/test_gpu_parity.py does)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    import torch
    from helpers import assert_same_map
    from sonar_3d_reconstruction_b200 import SonarTo3DMapper, synthetic
    world = int(os.environ.get("S3D_LOCAL_WORLD", "2"))
    n = int(os.environ.get("S3D_CHECK_FRAMES", "53"))
    name = os.environ.get("S3D_CHECK_WORKLOAD", "")
    spec = name or dict(H=160, W=200, config=dict(voxel_resolution=0.06, intensity_threshold=45, max_range=8.0), step_m=0.03)
    images, pos, quat, cfg = synthetic.make_sequence(spec, n, seed=4)
    H, W = images.shape[1:]
    plain = SonarTo3DMapper(cfg)
    ref_stats = plain.process_sonar_images(images, pos, quat)
    k1, L1 = plain.octree.voxels.to_arrays()

    maps = [SonarTo3DMapper(cfg) for _ in range(world)]
    T = maps[0].compose_transforms(pos, quat).reshape(n, 16)
    d_img = torch.from_numpy(images).cuda()
    d_T = torch.from_numpy(np.ascontiguousarray(T)).cuda()
    stats = [torch.zeros((n, 8), dtype=torch.int64, device="cuda") for _ in range(world)]
    handles = []
    for r, m in enumerate(maps):
        m._check_width(W)
        m._sync_device_config(H, W)
        nat = m.octree._native
        nat.shard_config(r, world)
        nat.reserve(4 * len(k1) // world + 100000)          # a routed map cannot re-run a chunk
        handles.append(nat.route_export(1 << 21))
    for m in maps:
        m.octree._native.route_attach(b"".join(handles), same_process=True)
    for m in maps:
        m.octree._native.route_enable(True)
    torch.cuda.synchronize()
    from sonar_3d_reconstruction_b200._native import NativeMap
    cf = NativeMap.CHUNK_FRAMES
    for f0 in range(0, n, cf):                               # one chunk per rank, in turn
        g = min(cf, n - f0)
        for r, m in enumerate(maps):
            m.octree._native.ingest_batch_dev(d_img.data_ptr() + f0 * H * W, g, d_T.data_ptr() + f0 * 128,
                                              want_stats=False, stats_dev_ptr=stats[r].data_ptr() + f0 * 64)
    for m in maps:
        m.octree._native.sync()
    tot = sum(s.cpu().numpy() for s in stats)
    for f in range(n):
        want = (ref_stats[f]["num_occupied"], ref_stats[f]["num_free"], ref_stats[f]["num_voxels"])
        assert tuple(int(x) for x in tot[f, :3]) == want, (f, tot[f], want)
    keys, vals = zip(*[m.octree.voxels.to_arrays() for m in maps])
    sizes = [len(k) for k in keys]
    err = assert_same_map(np.concatenate(keys), np.concatenate(vals), k1, L1, 0.0, "routed (one process) vs plain")
    print(f"route_local_check ok: world={world} frames={n} voxels={len(k1)} shards={sizes} max|dL|={err}")


if __name__ == "__main__":
    main()
