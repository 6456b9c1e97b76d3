#!/usr/bin/env python3
"""Stress of the growth / retry path: a fresh map of the default size swallows tens of thousands of frames
queued back to back (device-resident), so the voxel table rehash-grows many times and chunks are re-run while
others are in flight.  Repeats with new maps; checks the final voxel count against the first repetition."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from sonar_3d_reconstruction_b200 import SonarTo3DMapper, synthetic


def main():
    reps = int(os.environ.get("STRESS_REPS", "20"))
    n = int(os.environ.get("STRESS_FRAMES", "25000"))
    step = int(os.environ.get("STRESS_STEP", "5000"))
    wl = os.environ.get("STRESS_WORKLOAD", "cfg2")
    base, pos, quat, cfg = synthetic.make_sequence(wl, n, seed=1, distinct_images=250, cycle=False)
    d_base = torch.from_numpy(np.ascontiguousarray(base)).cuda()
    pad = int(os.environ.get("STRESS_PAD", "0"))       # extra frames behind the sequence (are reads past the end the fault?)
    d_all = d_base[torch.arange(n + pad, device="cuda") % len(base)]
    d_img = d_all[:n]
    H, W = base.shape[1:]
    first = None
    for r in range(reps):
        m = SonarTo3DMapper(cfg)
        m._check_width(W); m._sync_device_config(H, W)
        nat = m.octree._native
        T = np.ascontiguousarray(m.compose_transforms(pos, quat).reshape(n, 16))
        d_T = torch.from_numpy(T).cuda()
        st = torch.zeros((n, 8), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        t0 = time.time()
        warm = int(os.environ.get("STRESS_WARM", "0"))
        for f0 in range(0, n, step):
            k = min(step, n - f0)
            if warm and f0 == warm:          # as bench.py does after its warm-up steps: pre-size for the rest
                nat.sync()
                rate = (int(st[warm - 1, 2]) - int(st[warm // 2, 2])) / max(1, warm - warm // 2)
                nat.reserve(int(1.3 * rate * (n - warm)) + 100000)
                nat.profile_read()
            try:
                nat.ingest_batch_dev(d_img.data_ptr() + f0 * H * W, k, d_T.data_ptr() + f0 * 128, want_stats=False,
                                     stats_dev_ptr=st.data_ptr() + f0 * 64)
            except Exception as e:
                print(f"FAILED in rep {r} at frames {f0}..: {e}", flush=True)
                raise
        nat.sync()
        prof = nat.profile_read()
        last = st[-1].cpu().numpy()
        sig = (int(last[2]), int(st[:, 0].sum()), int(st[:, 1].sum()))
        print(f"rep {r}: {time.time() - t0:.2f} s voxels {sig[0]} grows {prof['grows']} retries {prof['retries']} cap {nat.capacity}", flush=True)
        cur = st.cpu().numpy()
        if first is None:
            first, first_st = sig, cur
        if sig != first:
            bad = np.nonzero((cur[:, :4] != first_st[:, :4]).any(axis=1))[0]
            print(f"MISMATCH in rep {r}: {len(bad)} frames differ; first at frame {bad[0]} (frame % step = {bad[0] % step}): "
                  f"{cur[bad[0], :4].tolist()} vs {first_st[bad[0], :4].tolist()}; next {bad[1:6].tolist()}", flush=True)
            d = cur[bad, :4] - first_st[bad, :4]
            print("  columns that differ (occ, free, voxels, samples):", (d != 0).sum(axis=0).tolist(), "; frame%step histogram:",
                  np.bincount(bad % step, minlength=min(step, 64))[:64].tolist(), flush=True)
        assert sig == first, (sig, first)
        m.close()
        del m, nat
    print("stress_growth ok")


if __name__ == "__main__":
    main()
