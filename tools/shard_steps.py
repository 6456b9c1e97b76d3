#!/usr/bin/env python3
"""Per-step wall times of the sharded device-resident arm (development probe; torchrun --nproc-per-node N)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from sonar_3d_reconstruction_b200 import synthetic
from sonar_3d_reconstruction_b200.sharded import ShardedSonarMapper

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_step, steps = 250, 11
images, pos, quat, cfg = synthetic.make_sequence("cfg2", n_step * steps, seed=1, distinct_images=250)
cfg = dict(cfg, device=local)
if os.environ.get("S3D_RESERVE"):
    cfg["table_capacity"] = 1 << 25
sh = ShardedSonarMapper(cfg, group=dist.group.WORLD, mode=os.environ.get("S3D_SHARD_MODE", "fused"))
sh.mapper._check_width(512); sh.mapper._sync_device_config(500, 512)
T = sh.mapper.compose_transforms(pos, quat)
d_img, d_T = sh.backend.upload(images, T)
nat = sh.backend.native
times = []
for s in range(steps):
    f0 = s * n_step
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = torch.empty((n_step, 8), dtype=torch.int64, device="cuda")
    nat.ingest_batch_dev(d_img[f0:f0 + n_step].data_ptr(), n_step, d_T[f0:f0 + n_step].data_ptr(), want_stats=False, stats_dev_ptr=st.data_ptr())
    t1 = time.perf_counter()
    nat.sync()
    t2 = time.perf_counter()
    dist.all_reduce(st); torch.cuda.synchronize()
    t3 = time.perf_counter()
    times.append((t1 - t0, t2 - t1, t3 - t2))
p = nat.profile_read()
if rank == 0:
    for s, (a, b, c) in enumerate(times):
        print(f"step {s}: enqueue {a*1e3:7.2f} ms  drain {b*1e3:7.2f} ms  allreduce {c*1e3:6.2f} ms")
    print("retries", p["retries"], "grows", p["grows"], "cap", nat.capacity)
dist.destroy_process_group()
