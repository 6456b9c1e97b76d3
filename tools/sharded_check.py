#!/usr/bin/env python3
"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/sharded_check.py

Every rank feeds the same synthetic sequence to a ShardedSonarMapper (NCCL all-to-all of the
per-voxel counts); rank 0 also runs the plain single-GPU mapper and the CPU oracle on a prefix
and asserts that the sharded map is identical (keys exact, log-odds exact on the GPU pair)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from helpers import assert_same_map
    from sonar_3d_reconstruction_b200 import SonarTo3DMapper, synthetic
    from sonar_3d_reconstruction_b200.sharded import ShardedSonarMapper
    n = int(os.environ.get("S3D_CHECK_FRAMES", "70"))
    images, pos, quat, cfg = synthetic.make_sequence("cfg2", n, seed=2)
    cfg = dict(cfg, device=local)
    plain_ref = None
    for mode in ("fused", "replicate", "route"):
        sh = ShardedSonarMapper(cfg, group=dist.group.WORLD, mode=mode)
        stats = sh.process_sonar_images(images, pos, quat)
        keys, L = sh.gather_map()
        pc = sh.get_point_cloud()
        if rank == 0:
            if plain_ref is None:
                plain = SonarTo3DMapper(cfg)
                ps = plain.process_sonar_images(images, pos, quat)
                plain_ref = (ps, *plain.octree.voxels.to_arrays(), plain.get_point_cloud()["num_occupied"])
            ps, k1, L1, n_occ = plain_ref
            for a, b in zip(stats, ps):
                assert (a["num_occupied"], a["num_free"], a["num_voxels"]) == (b["num_occupied"], b["num_free"], b["num_voxels"])
            err = assert_same_map(keys, L, k1, L1, 0.0, f"{world}-GPU sharded ({mode}) vs 1-GPU")
            assert pc["num_occupied"] == n_occ
            print(f"sharded_check ok: mode={mode} world={world} frames={n} voxels={len(k1)} max|dL|={err} "
                  f"exchange={sh.last_exchange_bytes / 1e6:.1f} MB sent by rank 0")
        del sh
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
