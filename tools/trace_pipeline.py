#!/usr/bin/env python3
"""GPU timeline of the chunk pipeline from in-kernel %globaltimer stamps (S3D_TRACE=1).
    python tools/trace_pipeline.py                       # one GPU
    torchrun --nproc-per-node N tools/trace_pipeline.py  # routed map over N GPUs
Prints, for a steady-state window of chunks on rank 0: start offsets and durations of
ack wait / expand / flag wait / merge / apply, and the chunk period."""
import os
import sys

os.environ["S3D_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from sonar_3d_reconstruction_b200 import SonarTo3DMapper, synthetic


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n = int(os.environ.get("S3D_TRACE_FRAMES", "1024"))
    wl = os.environ.get("S3D_TRACE_WORKLOAD", "cfg2")
    torch.cuda.set_device(local)
    images, pos, quat, cfg = synthetic.make_sequence(wl, n, seed=1, distinct_images=250)
    cfg = dict(cfg, device=local, table_capacity=1 << 25)
    if world > 1:
        import torch.distributed as dist
        from sonar_3d_reconstruction_b200.sharded import ShardedSonarMapper
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        sh = ShardedSonarMapper(cfg, group=dist.group.WORLD, mode=os.environ.get("S3D_SHARD_MODE", "fused"))
        m, native = sh.mapper, sh.backend.native
    else:
        m = SonarTo3DMapper(cfg)
        native = m.octree._native
    H, W = images.shape[1:]
    m._check_width(W); m._sync_device_config(H, W)
    T = m.compose_transforms(pos, quat).reshape(n, 16)
    d_img = torch.from_numpy(images).cuda(); d_T = torch.from_numpy(np.ascontiguousarray(T)).cuda()
    st = torch.zeros((n, 8), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    native.ingest_batch_dev(d_img.data_ptr(), n, d_T.data_ptr(), want_stats=False, stats_dev_ptr=st.data_ptr())
    native.sync()
    tr = native.trace_read().astype(np.float64)
    if world > 1:
        dist.barrier()
    if rank == 0:
        names = ["ack_wait", "expand", "flag_wait", "merge", "apply"]
        lo, hi = len(tr) // 2, min(len(tr) // 2 + 8, len(tr))
        t0 = tr[lo, 1, 0]
        for c in range(lo, hi):
            parts = []
            for k, nm in enumerate(names):
                s, e = tr[c, k]
                if e == 0:
                    continue
                parts.append(f"{nm} +{(s - t0) / 1e3:7.1f} {(e - s) / 1e3:6.1f}us")
            print(f"chunk {c}: " + " | ".join(parts))
        per = np.diff(tr[lo:len(tr) - 2, 4, 1]).mean() / 1e3
        dur = {nm: float(np.mean([(tr[c, k, 1] - tr[c, k, 0]) / 1e3 for c in range(lo, len(tr) - 2) if tr[c, k, 1] > 0] or [0]))
               for k, nm in enumerate(names)}
        print(f"world={world} chunk period {per:.1f} us = {16e6 / per:.0f} frames/s; mean durations us: "
              + ", ".join(f"{k} {v:.1f}" for k, v in dur.items()))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
