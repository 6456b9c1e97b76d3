#!/usr/bin/env python3
"""Sharded map fed from pinned host images, step by step with progress prints (development probe; torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from sonar_3d_reconstruction_b200 import synthetic
from sonar_3d_reconstruction_b200.sharded import ShardedSonarMapper
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_step, steps = 250, 6
images, pos, quat, cfg = synthetic.make_sequence("cfg2", n_step * steps, seed=1, distinct_images=250)
cfg = dict(cfg, device=local, table_capacity=1 << 24)
n_maps = int(os.environ.get("S3D_N_MAPS", "1"))
maps = [ShardedSonarMapper(cfg, group=dist.group.WORLD) for _ in range(n_maps)]
sh = maps[-1]
pinned = torch.from_numpy(images).pin_memory().numpy()
dist.barrier()
for s in range(steps):
    f0 = s * n_step
    t0 = time.perf_counter()
    try:
        out = sh.process_sonar_images(pinned[f0:f0 + n_step], pos[f0:f0 + n_step], quat[f0:f0 + n_step])
    except Exception as e:
        print(f"rank {rank} step {s} FAILED after {time.perf_counter() - t0:.1f}s: {e}", flush=True)
        p = sh.backend.native
        sys.exit(1)
    print(f"rank {rank} step {s}: {(time.perf_counter() - t0) * 1e3:.2f} ms voxels {out[-1]['num_voxels']}", flush=True)
dist.destroy_process_group()
