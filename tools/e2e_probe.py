#!/usr/bin/env python3
"""Where does the time of SonarTo3DMapper.process_sonar_images go?  (development probe, one GPU)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sonar_3d_reconstruction_b200 import SonarTo3DMapper, synthetic

n_step, steps = 250, 8
images, pos, quat, cfg = synthetic.make_sequence("cfg2", n_step * (steps + 3), seed=1, distinct_images=250)
m = SonarTo3DMapper(dict(cfg, table_capacity=1 << 25))
pinned = torch.from_numpy(images).pin_memory().numpy()
nat = m.octree._native
for s in range(3):
    f0 = s * n_step
    m.process_sonar_images(pinned[f0:f0 + n_step], pos[f0:f0 + n_step], quat[f0:f0 + n_step])
acc = dict(compose=0.0, sync_cfg=0.0, ingest=0.0, total=0.0)
for s in range(3, 3 + steps):
    f0 = s * n_step
    t0 = time.perf_counter()
    T = m.compose_transforms(pos[f0:f0 + n_step], quat[f0:f0 + n_step])
    t1 = time.perf_counter()
    m._sync_device_config(500, 512)
    t2 = time.perf_counter()
    st = nat.ingest_batch(np.ascontiguousarray(pinned[f0:f0 + n_step]), T)
    t3 = time.perf_counter()
    acc["compose"] += t1 - t0; acc["sync_cfg"] += t2 - t1; acc["ingest"] += t3 - t2
t0 = time.perf_counter()
for s in range(3 + steps, 3 + steps):
    pass
# whole API
m2 = SonarTo3DMapper(dict(cfg, table_capacity=1 << 25))
for s in range(3):
    f0 = s * n_step
    m2.process_sonar_images(pinned[f0:f0 + n_step], pos[f0:f0 + n_step], quat[f0:f0 + n_step])
t0 = time.perf_counter()
for s in range(3, 3 + steps):
    f0 = s * n_step
    m2.process_sonar_images(pinned[f0:f0 + n_step], pos[f0:f0 + n_step], quat[f0:f0 + n_step])
acc["total"] = time.perf_counter() - t0
# raw H2D bandwidth of one step's frames
d = torch.empty((n_step, 500, 512), dtype=torch.uint8, device="cuda")
src = torch.from_numpy(pinned[:n_step])
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    d.copy_(src, non_blocking=True)
torch.cuda.synchronize()
h2d = (time.perf_counter() - t0) / 10
print({k: round(v / steps * 1e3, 3) for k, v in acc.items()}, "ms per 250-frame step; raw H2D of a step %.3f ms (%.1f GB/s)" % (h2d * 1e3, src.numel() / h2d / 1e9),
      "is_pinned", src.is_pinned())
