#!/usr/bin/env python3
"""bench.py -- sonar frames/s and voxel log-odds updates/s of the B200 hot path.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
A *step* is one pass of the hot path over one batch of `--frames-per-step` posed synthetic
M750D-shaped frames (KIRO bags are not available offline).

N = 1   headline workload = BASELINE.json configs[1] ("cfg2": KIRO water-tank YAML values at 0.05 m
        voxels, tilt 60 deg) -- the configuration the unmodified Python reference can also be timed on
        within minutes.  The largest single-GPU configuration, configs[2] ("cfg3": 1024 x 2000 frames,
        0.02 m voxels, hash-table stress), is measured in the same run with the same fields and
        printed as the `cfg3` object of the same line.
N > 1   configs[3] ("cfg4": cfg-1 sensor, 0.25 m per frame survey): ONE map sharded by voxel-key hash
        over the ranks, the same frames on every rank (strong scaling), checked against a single-GPU
        map of the same frames before timing.

  value     frames/s with the frames already resident in HBM (CUDA events on the map's stream)
  e2e       frames/s through the reference-facing Python API with pinned HOST buffers
            (host->device copies and the stats read-back inside the timed region)
  parity    the first frames of the timed sequence against the CPU oracle: per-frame counters exact,
            key set exact, |dL| <= 1e-5; a mismatch makes the run fail
  roofline / cpu_baseline: see DESIGN.md section "Measurement".

`--impl reference` times the reference's own CPU implementation of the path on the same workload:
the unmodified scripts/3d_mapper.py (shipped untracked as baseline/_ref/3d_mapper.py) when it is there,
else the C restatement oracle/sonar_oracle.c.
"""
from __future__ import annotations

import argparse
import contextlib
import hashlib
import importlib.util
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from sonar_3d_reconstruction_b200 import synthetic  # noqa: E402

METRIC = "sonar_frames_per_s"
UNIT = "frames/s"
REF_FILE = os.path.join(ROOT, "baseline", "_ref", "3d_mapper.py")
STATS_WORDS = 8            # int64 words per s3d_frame_stats
DESCR = {
    "cfg1": "cfg1: library defaults, 500x512 frames (range x bearing), 0.05 m voxels",
    "cfg2": "cfg2: KIRO water-tank-shaped synthetic sequence, 500x512 frames (range x bearing), tilt 60 deg, 0.05 m voxels",
    "cfg3": "cfg3: high-resolution hash-table stress, 2000x1024 frames (range x bearing), 0.02 m voxels",
    "cfg4": "cfg4: large seabed survey, cfg-1 sensor (500x512 frames), 0.25 m per frame, 0.05 m voxels",
}


def note(msg):
    if os.environ.get("BENCH_DEBUG"):
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(workload):
    """DRAM bytes per k_apply_chunk launch from the committed cold-cache ncu capture (profiles/r2_traffic.json:
    `ncu --cache-control all`, 16 frames per launch), or None."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(workload, {}).get("k_apply_chunk", {}).get("dram_bytes_per_launch")
    return None


# tools/microbench_rmw.cu on this pool's B200 (profiles/r2_microbench_rmw.txt): random 16-byte slot reads with an
# 8-byte write-back into a table far larger than L2 -- the access pattern of the update kernel -- sustain
# 14-19 G slots/s (about 1.0-1.2 TB/s of 32-byte sectors); the lower figures belong to the grid sizes a kernel
# sharing the GPU can use.  Reported beside the streaming-bandwidth fraction, not instead of it.
RANDOM_RMW_GSLOTS_PER_S = 17.0


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(gpu_index: int):
    """Run this process on the CPUs of the NUMA node the GPU hangs off, so that the pinned host
    frames are allocated next to it (host->device copies from the far socket run at about half the
    bandwidth).  Returns a short description for the JSON line; any failure leaves the affinity alone."""
    try:
        import torch
        p = torch.cuda.get_device_properties(gpu_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return f"{bdf}: no NUMA information"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return f"{bdf}: node {node} has no CPU this process may use"
        os.sched_setaffinity(0, allowed)
        return f"{bdf}: bound to NUMA node {node} ({len(allowed)} CPUs)"
    except Exception as e:                                       # not fatal: measure unbound
        return f"unbound ({type(e).__name__})"


# ------------------------------------------------------------------------------- workload
class Workload:
    """A posed frame sequence: `distinct` generated images cycled over n_frames distinct poses
    (the kernels' work depends on the pose, not on which speckle realisation is used)."""

    def __init__(self, name: str, n_frames: int, seed: int, distinct: int, step: int = 0):
        spec = synthetic.CONFIGS[name]
        self.name, self.H, self.W, self.n = name, spec["H"], spec["W"], n_frames
        self.k = min(n_frames, distinct) if distinct > 0 else n_frames
        if step > 0 and self.k < n_frames:
            while step % self.k:          # a whole number of cycles per step: every step's host buffer holds the same images
                self.k -= 1
        images, self.pos, self.quat, self.cfg = synthetic.make_sequence(name, n_frames, seed=seed, distinct_images=self.k,
                                                                        cycle=False)
        self.base = np.ascontiguousarray(images)

    def image(self, f: int) -> np.ndarray:
        return self.base[f % self.k]

    def images(self, f0: int, f1: int) -> np.ndarray:
        return self.base[np.arange(f0, f1) % self.k]


def load_reference_class():
    """The unmodified reference module, loaded by file path exactly as its ROS2 node does
    (scripts/3d_mapper_node.py:37-42).  None when baseline/_ref/ was not shipped."""
    if not os.path.exists(REF_FILE):
        return None
    spec = importlib.util.spec_from_file_location("reference_3d_mapper", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.SonarTo3DMapper


def time_cpu(make_mapper, wl: Workload, cfg: dict, budget_s: float, max_frames: int, first: int = 0):
    """frames/s and updates/s of a CPU mapper on frames [first, ...) of the workload, bounded by time."""
    m = make_mapper(cfg)
    upd = n = 0
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        for f in range(first, min(wl.n, first + max_frames)):
            st = m.process_sonar_image(wl.image(f), list(wl.pos[f]), list(wl.quat[f]))
            upd += st["num_occupied"] + st["num_free"]
            n += 1
            if time.perf_counter() - t0 > budget_s:
                break
    dt = time.perf_counter() - t0
    return n / dt, upd / dt, n


def cpu_baseline(wl: Workload, cfg: dict, ref_budget_s: float, port_frames: int):
    """The reference's CPU path on this box's host cores: the unmodified Python reference (single-threaded
    pure Python: 1 core) on a bounded sample, with the C restatement beside it."""
    from oracle.oracle import OracleMapper
    cfg = {k: v for k, v in cfg.items() if k not in ("device", "table_capacity")}
    port_fps, port_ups, port_n = time_cpu(OracleMapper, wl, cfg, 25.0, port_frames)
    out = {"unit": UNIT, "cores": 1, "host_cpus": os.cpu_count(),
           "port_value": port_fps, "port_voxel_updates_per_s": port_ups,
           "port_sample": f"first {port_n} frames through oracle/sonar_oracle.c (single-threaded C restatement)"}
    ref_cls = load_reference_class()
    if ref_cls is not None:
        fps, ups, n = time_cpu(ref_cls, wl, cfg, ref_budget_s, 10000)
        out.update(value=fps, kind="reference", voxel_updates_per_s=ups,
                   sample=f"first {n} frame(s) of the same sequence through the unmodified reference "
                          f"(baseline/_ref/3d_mapper.py = scripts/3d_mapper.py, SonarTo3DMapper.process_sonar_image, "
                          f"single-threaded pure Python + numpy)")
    else:
        out.update(value=port_fps, kind="port", voxel_updates_per_s=port_ups, sample=out["port_sample"] +
                   "; baseline/_ref/3d_mapper.py was not shipped, so the unmodified reference could not be timed")
    return out


def parity_check(gpu_stats: np.ndarray, keys: np.ndarray, L: np.ndarray, wl: Workload, cfg: dict, n: int):
    """First n frames of the sequence through the CPU oracle (itself pinned on the unmodified reference's
    outputs, tests/test_oracle_golden.py) against what the GPU produced for the same frames."""
    from oracle.oracle import OracleMapper
    cpu = OracleMapper({k: v for k, v in cfg.items() if k not in ("device", "table_capacity")})
    want = np.zeros((n, 4), dtype=np.int64)
    for f in range(n):
        st = cpu.process_sonar_image(wl.image(f), wl.pos[f], wl.quat[f])
        want[f] = [st["num_occupied"], st["num_free"], st["num_voxels"], st["num_samples"]]
    counters_ok = bool(np.array_equal(want, gpu_stats[:n, :4]))
    kc, vc = cpu.dump()
    kg = np.asarray(keys, dtype=np.int64).reshape(-1, 3)
    og, oc = np.lexsort(kg.T[::-1]), np.lexsort(kc.T[::-1])
    keys_ok = kg.shape == kc.shape and bool(np.array_equal(kg[og], kc[oc]))
    max_dl = float(np.abs(np.asarray(L)[og] - vc[oc]).max()) if keys_ok and len(kg) else (0.0 if keys_ok else float("nan"))
    ok = counters_ok and keys_ok and max_dl <= 1e-5
    return {"frames": n, "ok": bool(ok), "counters_exact": counters_ok, "keys_exact": bool(keys_ok), "max_dL": max_dl,
            "voxels": int(len(kc)), "sum_logodds": float(vc.sum()),
            "key_hash": hashlib.sha256(kc[oc].tobytes()).hexdigest()[:16],
            "checker": "oracle/sonar_oracle.c (CPU), tolerance 1e-5 on log-odds, everything else exact"}


# ------------------------------------------------------------------------------- single GPU
def run_config(args, name: str, fps_step: int, local_rank: int, barrier, ref_budget_s: float, parity_frames: int,
               port_frames: int, unreserved: bool):
    """All measurements of one workload on one GPU; returns the fields of a bench line."""
    import torch
    from sonar_3d_reconstruction_b200 import SonarTo3DMapper

    dev = f"cuda:{local_rank}"
    n_total = (args.steps + args.warmup) * fps_step
    wl = Workload(name, n_total, args.seed, args.distinct_images, fps_step)
    H, W = wl.H, wl.W
    cfg = dict(wl.cfg, device=local_rank)
    if args.no_adaptive:
        cfg["adaptive_update"] = False
    img_bytes = H * W

    d_base = torch.from_numpy(wl.base).to(dev)
    d_img = d_base[torch.arange(n_total, device=dev) % wl.k]      # every timed step reads its own frames from HBM

    def fresh(table_capacity=0, env=None):
        old = {}
        for k_, v_ in (env or {}).items():
            old[k_] = os.environ.get(k_)
            os.environ[k_] = v_
        try:
            m = SonarTo3DMapper(dict(cfg, table_capacity=table_capacity) if table_capacity else cfg)
        finally:
            for k_, v_ in old.items():
                if v_ is None:
                    os.environ.pop(k_, None)
                else:
                    os.environ[k_] = v_
        m._check_width(W)
        m._sync_device_config(H, W)
        return m

    mapper = fresh()
    native = mapper.octree._native
    T_all = np.ascontiguousarray(mapper.compose_transforms(wl.pos, wl.quat).reshape(n_total, 16))
    d_T = torch.from_numpy(T_all).to(dev)
    d_stats = torch.zeros((n_total, STATS_WORDS), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    def step_dev(nat, s, stats):
        f0 = s * fps_step
        nat.ingest_batch_dev(d_img.data_ptr() + f0 * img_bytes, fps_step, d_T.data_ptr() + f0 * 128,
                             want_stats=False, stats_dev_ptr=stats.data_ptr() + f0 * 8 * STATS_WORDS)

    def timed_pass(nat, stats, reserve):
        for s in range(args.warmup):
            step_dev(nat, s, stats)
        nat.sync()
        if reserve:
            # pre-size the table for the timed frames from the growth rate seen in warm-up, as a user
            # who knows the survey length would
            st_w = stats[: args.warmup * fps_step].cpu().numpy()
            rate = (int(st_w[-1, 2]) - int(st_w[len(st_w) // 2, 2])) / max(1, len(st_w) - len(st_w) // 2)
            nat.reserve(int(1.3 * rate * fps_step * args.steps) + 100000)
        nat.profile_read()
        stream = torch.cuda.ExternalStream(nat.stream, device=local_rank)
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for s in range(args.warmup, args.warmup + args.steps):
            step_dev(nat, s, stats)
        ev1.record(stream)
        nat.sync()
        barrier()
        clocks = sampler.stop()
        return ev0.elapsed_time(ev1), nat.profile_read(), clocks

    # ------------------------------------------------------------- value: inputs resident in HBM
    note(f"{name}: timed pass, {n_total} frames resident")
    ms_total, prof, clocks = timed_pass(native, d_stats, reserve=True)
    note(f"{name}: timed pass done, table {native.capacity} slots")
    st_all = d_stats.cpu().numpy()
    st = st_all[args.warmup * fps_step:]
    n_frames = args.steps * fps_step
    updates = int((st[:, 0] + st[:, 1]).sum())
    samples = int(st[:, 3].sum())
    n_voxels = int(st[-1, 2])
    cap = native.capacity
    mapper.close()
    del mapper, native

    # ------------------------------------------------------------- parity of the timed sequence's first frames
    par = None
    note(f"{name}: parity")
    if parity_frames > 0:
        n_par = min(parity_frames, n_total)
        pm = fresh()
        pst = torch.zeros((n_par, STATS_WORDS), dtype=torch.int64, device=dev)
        pm.octree._native.ingest_batch_dev(d_img.data_ptr(), n_par, d_T.data_ptr(), want_stats=False, stats_dev_ptr=pst.data_ptr())
        pm.octree._native.sync()
        pst = pst.cpu().numpy()
        assert np.array_equal(pst[:, :4], st_all[:n_par, :4]), "two GPU passes over the same frames disagree"
        keys, L = pm.octree.voxels.to_arrays()
        par = parity_check(pst, keys, L, wl, cfg, n_par)
        pm.close()
        del pm

    # ------------------------------------------------------------- exclusive per-kernel times (no overlap)
    note(f"{name}: exclusive pass")
    sm = fresh(table_capacity=cap, env={"S3D_SERIAL_KERNELS": "1"})
    sn = sm.octree._native
    s_stats = torch.zeros((n_total, STATS_WORDS), dtype=torch.int64, device=dev)
    k_steps = min(args.steps, max(1, 4096 // fps_step))
    for s in range(args.warmup):
        step_dev(sn, s, s_stats)
    sn.sync()
    sn.profile_enable(True)
    sn.profile_read()
    for s in range(args.warmup, args.warmup + k_steps):
        step_dev(sn, s, s_stats)
    sn.sync()
    sprof = sn.profile_read()
    sn.profile_enable(False)
    sst = s_stats[args.warmup * fps_step:(args.warmup + k_steps) * fps_step].cpu().numpy()
    s_updates, s_frames = int((sst[:, 0] + sst[:, 1]).sum()), k_steps * fps_step
    sm.close()
    del sm, sn, s_stats

    # ------------------------------------------------------------- no pre-sizing: rehash-grows inside the timed region
    unres = None
    note(f"{name}: unreserved / e2e")
    if unreserved:
        um = fresh()
        u_stats = torch.zeros((n_total, STATS_WORDS), dtype=torch.int64, device=dev)
        u_ms, u_prof, _ = timed_pass(um.octree._native, u_stats, reserve=False)
        assert int(u_stats[-1, 2]) == n_voxels, (int(u_stats[-1, 2]), n_voxels)
        unres = {"value": n_frames / (u_ms * 1e-3), "unit": UNIT, "table_grows_timed": u_prof["grows"],
                 "chunk_retries": u_prof["retries"], "table_slots_end": um.octree._native.capacity,
                 "note": "same timed frames, table not pre-sized: every rehash-grow (allocate, re-insert, free) is inside the timed region"}
        um.close()
        del um, u_stats

    # ------------------------------------------------------------- e2e: public API, host buffers
    e2e = None
    if not args.no_e2e:
        pin = torch.empty((2, fps_step, H, W), dtype=torch.uint8).pin_memory()      # two steps of pinned host frames
        pinned = pin.numpy()
        # When the generated images cycle with the step length both buffers hold every step's frames, and the
        # timed region contains exactly the copies and the calls; otherwise the step's frames are gathered into
        # its buffer first (numpy, inside the timed region: it only makes the number smaller)
        same_cycle = fps_step % wl.k == 0
        for slot in range(2):
            pinned[slot] = wl.images(0, fps_step)

        def args_of(s, slot):
            f0 = s * fps_step
            if not same_cycle:
                pinned[slot] = wl.images(f0, f0 + fps_step)
            return pinned[slot], wl.pos[f0:f0 + fps_step], wl.quat[f0:f0 + fps_step]

        # (a) the blocking batch call, one step per call
        m2 = fresh(table_capacity=cap)
        for s in range(args.warmup):
            m2.process_sonar_images(*args_of(s, s & 1))
        barrier()
        t0 = time.perf_counter()
        for s in range(args.warmup, args.warmup + args.steps):
            out = m2.process_sonar_images(*args_of(s, s & 1))
        torch.cuda.synchronize()
        blocking_s = time.perf_counter() - t0
        assert out[-1]["num_voxels"] == n_voxels, (out[-1]["num_voxels"], n_voxels)
        m2.close()
        del m2
        # (b) the asynchronous form of the call, two steps pending at a time: step s+1 is uploaded and expanded
        # while step s finishes; every step still copies its frames from pinned host memory and reads its
        # per-frame result back inside the timed region
        m3 = fresh(table_capacity=cap)
        for s in range(args.warmup):
            m3.process_sonar_images_async(*args_of(s, s & 1)).result()
        barrier()
        t0 = time.perf_counter()
        pending = None
        for s in range(args.warmup, args.warmup + args.steps):
            h = m3.process_sonar_images_async(*args_of(s, s & 1))
            if pending is not None:
                out = pending.result()
            pending = h
        out = pending.result()
        torch.cuda.synchronize()
        async_s = time.perf_counter() - t0
        assert out[-1]["num_voxels"] == n_voxels, (out[-1]["num_voxels"], n_voxels)      # both arms built the same map
        m3.close()
        del m3
        # (c) the reference's own call: one frame per process_sonar_image (the only call the ROS2 node makes)
        n_single = min(args.single_frames, fps_step)
        m4 = fresh(table_capacity=cap)
        for f in range(min(8, n_total)):
            m4.process_sonar_image(wl.image(f), list(wl.pos[f]), list(wl.quat[f]))
        f0 = args.warmup * fps_step
        frames = [np.ascontiguousarray(wl.image(f)) for f in range(f0, f0 + n_single)]
        t0 = time.perf_counter()
        for i, f in enumerate(range(f0, f0 + n_single)):
            m4.process_sonar_image(frames[i], list(wl.pos[f]), list(wl.quat[f]))
        single_s = time.perf_counter() - t0
        m4.close()
        del m4
        e2e = {"value": n_frames / async_s, "unit": UNIT,
               "h2d_bytes_per_step": fps_step * (H * W + 128), "d2h_bytes_per_step": fps_step * 8 * STATS_WORDS,
               "api": f"SonarTo3DMapper.process_sonar_images_async, two {fps_step}-frame steps pending at a time "
                      "(pinned host images, poses on host; H2D of every frame and D2H of every step's per-frame "
                      "counters inside the timed region)",
               "blocking_value": n_frames / blocking_s,
               "blocking_api": "SonarTo3DMapper.process_sonar_images, one step per call",
               "single_frame_value": n_single / single_s,
               "single_frame_api": f"SonarTo3DMapper.process_sonar_image, the reference's own call, {n_single} frames one "
                                   "call each from pageable host arrays (H2D of the frame, both kernels, D2H of the counters, "
                                   "one synchronisation per call)",
               "host": args.numa}
        del pin, pinned

    # ------------------------------------------------------------- the line
    peak, peak_src = load_peaks()
    secs = ms_total * 1e-3
    alg_bytes = n_frames * H * W + 24 * updates                 # image read once + 16 B slot read + 8 B write per update
    whole = alg_bytes / secs / 1e9
    kern = {}
    for k_ in ("k_expand", "k_apply"):
        if sprof["launches"][k_]:
            us = sprof["ms"][k_] / sprof["launches"][k_] * 1e3
            ab = (s_frames * H * W if k_ == "k_expand" else 24 * s_updates) / sprof["launches"][k_]
            kern[k_] = {"us_per_launch_exclusive": us, "launches": sprof["launches"][k_],
                        "alg_bytes_per_launch": ab, "gbs": ab / (us * 1e-6) / 1e9, "frac": ab / (us * 1e-6) / 1e9 / peak}
    ka = kern.get("k_apply", {"gbs": 0.0, "frac": 0.0})
    if "k_apply" in kern:
        probes_per_s = sprof["voxel_probes"] / (sprof["ms"]["k_apply"] * 1e-3)
        kern["k_apply"]["voxel_probes_per_launch"] = sprof["voxel_probes"] / sprof["launches"]["k_apply"]
        kern["k_apply"]["table_probes_per_s"] = probes_per_s
        kern["k_apply"]["frac_of_random_rmw_bound"] = probes_per_s / (RANDOM_RMW_GSLOTS_PER_S * 1e9)
        kern["k_apply"]["random_rmw_bound"] = (f"{RANDOM_RMW_GSLOTS_PER_S} G slot read-modify-writes/s measured on this pool's B200 "
                                               "(tools/microbench_rmw.cu, profiles/r2_microbench_rmw.txt): the kernel touches one random "
                                               "32-byte sector of a table >> L2 per voxel per launch, so this, not the streaming peak, bounds it")
    line = {
        "metric": METRIC, "value": n_frames / secs, "unit": UNIT,
        "ms_per_step": ms_total / args.steps,
        "voxel_updates_per_s": updates / secs, "samples_per_s": samples / secs,
        "config": {"workload": DESCR[name], "frames_per_step": fps_step, "frames_timed": n_frames,
                   "updates_per_frame": updates / n_frames, "samples_per_frame": samples / n_frames,
                   "map_voxels_end": n_voxels, "table_slots": cap,
                   "chunk_retries": prof["retries"], "table_grows_timed": prof["grows"],
                   "l2": f"inputs streamed once: every step reads its own frames from HBM (timed input "
                         f"{n_frames * H * W / 2**20:.0f} MiB > L2), voxel table {cap * 16 / 2**20:.0f} MiB",
                   "adaptive_update": bool(cfg.get("adaptive_update", True)),
                   "parallelism": "single GPU"},
        "gpu_launches": prof["total_launches"],
        "roofline": {"bound": "hbm", "kernel": "k_apply_chunk -- the update kernel north_star's >= 50 % target names",
                     "achieved": ka["gbs"], "peak": peak, "unit": "GB/s", "frac": ka["frac"],
                     "traffic": load_traffic(name),
                     "how": "algorithmic bytes = 24 B (16 B slot read + 8 B write-back) x voxel updates of one 16-frame "
                            "launch / that kernel's exclusive launch duration (CUDA events on the launching stream, "
                            "kernels serialised on one stream for this measurement: S3D_SERIAL_KERNELS)",
                     "peak_source": peak_src,
                     "whole_path": {"achieved": whole, "frac": whole / peak,
                                    "how": "(H*W + 24*U) bytes per frame / CUDA-event wall time of the timed region, "
                                           "k_expand of two chunks overlapping k_apply_chunk of a third"},
                     "kernels": kern},
        "clocks": clocks,
    }
    if par is not None:
        line["parity"] = par
    if unres is not None:
        line["unreserved"] = unres
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(wl, cfg, ref_budget_s, port_frames)
    del d_img, d_base, d_T, d_stats
    torch.cuda.empty_cache()
    return line


def run_single(args, local_rank):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: sonar_3d_reconstruction_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    args.numa = bind_to_gpu_numa(local_rank)

    def barrier():
        torch.cuda.synchronize()

    fps_step = args.frames_per_step or (4000 if args.workload != "cfg3" else args.cfg3_frames_per_step)
    head = run_config(args, args.workload, fps_step, local_rank, barrier,
                      ref_budget_s=18.0 if args.workload != "cfg3" else 1.0, parity_frames=args.parity_frames if args.workload != "cfg3"
                      else min(args.parity_frames, 12), port_frames=args.cpu_frames if args.workload != "cfg3" else 24,
                      unreserved=args.workload == "cfg3")
    line = {"metric": METRIC, "value": head.pop("value"), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic"}
    head.pop("metric"); head.pop("unit")
    line.update(head)
    if args.workload == "cfg2" and not args.no_cfg3:
        # the largest single-GPU configuration, same fields (its CPU legs are shorter: one reference frame takes ~18 s)
        sub = run_config(args, "cfg3", args.cfg3_frames_per_step, local_rank, barrier, ref_budget_s=1.0,
                         parity_frames=min(args.parity_frames, 12), port_frames=24, unreserved=True)
        sub["steps"], sub["warmup"] = args.steps, args.warmup
        line["cfg3"] = sub
    ok = all(p is None or p.get("ok", True) for p in (line.get("parity"), line.get("cfg3", {}).get("parity")))
    print(json.dumps(line))
    if not ok:
        print("bench.py: PARITY FAILED (see the `parity` objects of the line above)", file=sys.stderr)
        sys.exit(3)


# ------------------------------------------------------------------------------- N GPUs
def run_sharded(args, rank, world, local_rank):
    """N > 1: ONE map sharded by voxel-key hash over the ranks (strong scaling: the same frames on every
    rank, rank c mod N expands 16-frame chunk c -- or, S3D_ROUTE_SPLIT=beams, every rank 1/N of the beams of
    every chunk -- and every rank owns 1/N of the voxels; the expansion kernel writes the records of remote
    owners into their inboxes over NVLink peer memory)."""
    import torch
    import torch.distributed as dist

    from sonar_3d_reconstruction_b200 import SonarTo3DMapper
    from sonar_3d_reconstruction_b200.sharded import ShardedSonarMapper

    torch.cuda.set_device(local_rank)
    args.numa = bind_to_gpu_numa(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = f"cuda:{local_rank}"

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    name = args.workload if args.workload_given else "cfg4"
    fps_step = args.frames_per_step or 2000
    n_total = (args.steps + args.warmup) * fps_step
    wl = Workload(name, n_total, args.seed, args.distinct_images, fps_step)
    H, W = wl.H, wl.W
    cfg = dict(wl.cfg, device=local_rank)
    if args.no_adaptive:
        cfg["adaptive_update"] = False
    img_bytes = H * W
    d_base = torch.from_numpy(wl.base).to(dev)
    d_img = d_base[torch.arange(n_total, device=dev) % wl.k]
    sh = ShardedSonarMapper(cfg, group=dist.group.WORLD, mode=args.shard_mode)
    sh.mapper._check_width(W)
    sh.mapper._sync_device_config(H, W)
    T_all = np.ascontiguousarray(sh.mapper.compose_transforms(wl.pos, wl.quat).reshape(n_total, 16))
    d_T = torch.from_numpy(T_all).to(dev)
    native = sh.backend.native

    # ---- parity before timing: the N-rank map of the first frames must equal a single-GPU map of the same
    # frames bit for bit (integer merges), and that one the CPU oracle's
    n_par = min(args.parity_frames, n_total, 160)
    par = None
    if n_par > 0:
        stats_sh = sh.process_device_batch(d_img[:n_par], d_T[:n_par]).cpu().numpy()
        ks, ls = sh.gather_map()
        if rank == 0:
            plain = SonarTo3DMapper(cfg)
            plain._check_width(W)
            plain._sync_device_config(H, W)
            pst = torch.zeros((n_par, STATS_WORDS), dtype=torch.int64, device=dev)
            plain.octree._native.ingest_batch_dev(d_img.data_ptr(), n_par, d_T.data_ptr(), want_stats=False, stats_dev_ptr=pst.data_ptr())
            plain.octree._native.sync()
            pst = pst.cpu().numpy()
            kp, lp = plain.octree.voxels.to_arrays()
            ks64, kp64 = np.asarray(ks, dtype=np.int64), np.asarray(kp, dtype=np.int64)
            o1, o2 = np.lexsort(ks64.T[::-1]), np.lexsort(kp64.T[::-1])
            same_keys = len(ks64) == len(kp64) and bool(np.array_equal(ks64[o1], kp64[o2]))
            same_L = same_keys and bool(np.array_equal(np.asarray(ls)[o1], np.asarray(lp)[o2]))
            same_stats = bool(np.array_equal(stats_sh[:, :4], pst[:, :4]))
            par = parity_check(pst, kp, lp, wl, cfg, n_par)
            par["sharded_vs_single_gpu"] = {"frames": n_par, "keys_identical": same_keys, "logodds_bit_identical": same_L,
                                            "counters_identical": same_stats}
            par["ok"] = bool(par["ok"] and same_keys and same_L and same_stats)
            plain.close()
            del plain
        sh.reset_map()

    def step_dev(s, stats, o0):
        f0 = s * fps_step
        native.ingest_batch_dev(d_img.data_ptr() + f0 * img_bytes, fps_step, d_T.data_ptr() + f0 * 128,
                                want_stats=False, stats_dev_ptr=stats.data_ptr() + o0 * 8 * STATS_WORDS)

    fused_like = sh.mode in ("fused", "replicate")
    w_stats = torch.zeros((max(1, args.warmup) * fps_step, STATS_WORDS), dtype=torch.int64, device=dev)
    if fused_like:
        for s in range(args.warmup):
            step_dev(s, w_stats, s * fps_step)
        native.sync()
        dist.all_reduce(w_stats)
        st_w = w_stats.cpu().numpy()
    else:
        st_w = torch.cat([sh.process_device_batch(d_img[s * fps_step:(s + 1) * fps_step], d_T[s * fps_step:(s + 1) * fps_step])
                          for s in range(args.warmup)], dim=0).cpu().numpy()
    n_frames = args.steps * fps_step
    rate = (int(st_w[-1, 2]) - int(st_w[len(st_w) // 2, 2])) / max(1, len(st_w) - len(st_w) // 2)
    native.reserve(int(1.5 * rate * n_frames / world) + 100000)
    cap = native.capacity
    native.profile_read()
    sampler = ClockSampler(local_rank)
    sampler.start()                       # before the barrier: spawning nvidia-smi must not skew the ranks' start
    stream = torch.cuda.ExternalStream(native.stream, device=local_rank)
    d_stats = torch.zeros((n_frames, STATS_WORDS), dtype=torch.int64, device=dev)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if fused_like:
        # the steps are queued back to back on the map's streams (no host sync in between), then
        # one sync and one all-reduce of the per-shard counters -- all inside the timed region
        ev0.record(stream)
        for s in range(args.warmup, args.warmup + args.steps):
            step_dev(s, d_stats, (s - args.warmup) * fps_step)
        native.sync()
        dist.all_reduce(d_stats)
        ev1.record()
        torch.cuda.synchronize()
    else:
        ev0.record()
        d_stats = torch.cat([sh.process_device_batch(d_img[s * fps_step:(s + 1) * fps_step], d_T[s * fps_step:(s + 1) * fps_step])
                             for s in range(args.warmup, args.warmup + args.steps)], dim=0)
        ev1.record()
        torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    prof = native.profile_read()
    st = d_stats.cpu().numpy()
    updates, samples, n_voxels = int((st[:, 0] + st[:, 1]).sum()), int(st[:, 3].sum()), int(st[-1, 2])
    rec_sent = prof.get("route_records_sent", 0)

    # ---- the same timed frames on rank 0 alone (plain mapper): what one GPU does with this workload
    single = None
    if not args.no_single_ref:
        if rank == 0:
            one = SonarTo3DMapper(cfg)
            one._check_width(W)
            one._sync_device_config(H, W)
            on = one.octree._native
            o_stats = torch.zeros((n_total, STATS_WORDS), dtype=torch.int64, device=dev)

            def one_step(s):
                on.ingest_batch_dev(d_img.data_ptr() + s * fps_step * img_bytes, fps_step, d_T.data_ptr() + s * fps_step * 128,
                                    want_stats=False, stats_dev_ptr=o_stats.data_ptr() + s * fps_step * 8 * STATS_WORDS)
            for s in range(args.warmup):
                one_step(s)
            on.sync()
            on.reserve(int(1.5 * rate * n_frames) + 100000)
            ostream = torch.cuda.ExternalStream(on.stream, device=local_rank)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ostream)
            for s in range(args.warmup, args.warmup + args.steps):
                one_step(s)
            e1.record(ostream)
            on.sync()
            one_ms = e0.elapsed_time(e1)
            assert int(o_stats[-1, 2]) == n_voxels, (int(o_stats[-1, 2]), n_voxels)
            single = {"value": n_frames / (one_ms * 1e-3), "unit": UNIT,
                      "note": "the same timed frames through one unsharded map on rank 0's GPU, after the sharded run"}
            one.close()
            del one, on, o_stats
        barrier()

    # ---- BASELINE config 5's ablation: the same timed steps with the adaptive rule switched off
    ablation = None
    if fused_like and not args.no_ablation:
        off = ShardedSonarMapper(dict(cfg, adaptive_update=not bool(cfg.get("adaptive_update", True)), table_capacity=cap),
                                 group=dist.group.WORLD, mode=args.shard_mode)
        off.mapper._check_width(W)
        off.mapper._sync_device_config(H, W)
        onat = off.backend.native
        a_stats = torch.zeros((n_total, STATS_WORDS), dtype=torch.int64, device=dev)

        def off_step(s):
            onat.ingest_batch_dev(d_img.data_ptr() + s * fps_step * img_bytes, fps_step, d_T.data_ptr() + s * fps_step * 128,
                                  want_stats=False, stats_dev_ptr=a_stats.data_ptr() + s * fps_step * 8 * STATS_WORDS)
        for s in range(args.warmup):
            off_step(s)
        onat.sync()
        onat.reserve(int(1.5 * rate * n_frames / world) + 100000)
        astream = torch.cuda.ExternalStream(onat.stream, device=local_rank)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(astream)
        for s in range(args.warmup, args.warmup + args.steps):
            off_step(s)
        onat.sync()
        a1.record()
        torch.cuda.synchronize()
        barrier()
        t_off = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t_off, op=dist.ReduceOp.MAX)
        ablation = {"adaptive_update": not bool(cfg.get("adaptive_update", True)), "value": n_frames / (float(t_off[0]) * 1e-3), "unit": UNIT,
                    "note": "BASELINE config 5: the same timed frames with the adaptive rule toggled (batched ingest, "
                            f"{fps_step} frames per call)"}
        off.mapper.close()
        del off, onat, a_stats

    e2e_s = float("nan")
    if not args.no_e2e:
        sh2 = ShardedSonarMapper(dict(cfg, table_capacity=cap), group=dist.group.WORLD, mode=args.shard_mode)
        pin = torch.empty((fps_step, H, W), dtype=torch.uint8).pin_memory()
        pinned = pin.numpy()
        pinned[:] = wl.images(0, fps_step)
        same_cycle = fps_step % wl.k == 0
        barrier()                              # page-locking takes a different time on every rank

        def args_of(s):
            f0 = s * fps_step
            if not same_cycle:
                pinned[:] = wl.images(f0, f0 + fps_step)
            return pinned, wl.pos[f0:f0 + fps_step], wl.quat[f0:f0 + fps_step]

        for s in range(args.warmup):
            sh2.process_sonar_images(*args_of(s))
        barrier()
        t0 = time.perf_counter()
        for s in range(args.warmup, args.warmup + args.steps):
            out = sh2.process_sonar_images(*args_of(s))
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        assert out[-1]["num_voxels"] == n_voxels
    t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
    tot = torch.tensor([prof["total_launches"], rec_sent], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, e2e_s = float(t[0]), float(t[1])
    if rank == 0:
        peak, peak_src = load_peaks()
        secs = ms_total * 1e-3
        alg_bytes = n_frames * H * W + 24 * updates
        achieved = alg_bytes / secs / 1e9
        nvl_bytes = 16.0 * float(tot[1])
        line = {
            "metric": METRIC, "value": n_frames / secs, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "voxel_updates_per_s": updates / secs, "samples_per_s": samples / secs,
            "config": {"workload": f"{DESCR[name]}; one map sharded by voxel-key hash over {world} GPUs",
                       "frames_per_step": fps_step, "frames_timed": n_frames, "updates_per_frame": updates / n_frames,
                       "map_voxels_end": n_voxels, "table_slots_per_rank": cap,
                       "parallelism": (f"shard{world} ({sh.mode}, " +
                                       (f"split by beams: each rank expands beams/{world} of every frame"
                                        if os.environ.get("S3D_ROUTE_SPLIT") == "beams" else
                                        f"split by chunks: rank c mod {world} expands every beam of 16-frame chunk c") +
                                       "); every rank owns 1/N of the voxels; the "
                                       "expansion kernel writes (voxel, frame, counts) records of remote owners into their "
                                       "inboxes over NVLink peer memory (CUDA IPC), device-side flags, the owner merges "
                                       "and applies; no host sync or separate all-to-all per chunk")
                       if sh.mode == "fused" else f"shard{world} ({sh.mode})",
                       "nvlink_bytes_timed_all_ranks": nvl_bytes,
                       "nvlink_bytes_per_frame": nvl_bytes / n_frames,
                       "nvlink_GBps_per_rank": nvl_bytes / world / secs / 1e9,
                       "l2": "inputs streamed once: every step reads its own frames from HBM"},
            "gpu_launches": int(tot[0]),
            "roofline": {"bound": "hbm", "achieved": achieved / world, "peak": peak, "unit": "GB/s",
                         "frac": achieved / world / peak, "traffic": None,
                         "kernel": "whole sharded step (expand+route, merge, apply), per GPU: (H*W + 24*U) bytes per frame / wall time",
                         "peak_source": peak_src},
            "clocks": clocks,
        }
        if single is not None:
            line["single_gpu_same_workload"] = single
            line["speedup_vs_one_gpu_same_workload"] = line["value"] / single["value"]
        if par is not None:
            line["parity"] = par
        if ablation is not None:
            line["adaptive_ablation"] = ablation
        if not args.no_e2e:
            line["e2e"] = {"value": n_frames / e2e_s, "unit": UNIT,
                           "h2d_bytes_per_step": fps_step * H * W + world * fps_step * 128, "d2h_bytes_per_step": fps_step * 8 * STATS_WORDS,
                           "api": "ShardedSonarMapper.process_sonar_images (pinned host images on every rank; each rank "
                                  "uploads 1/N of the frames and the ranks all-gather them over NVLink)",
                           "host": args.numa}
        print(json.dumps(line))
        if par is not None and not par["ok"]:
            print("bench.py: PARITY FAILED (see the `parity` object of the line above)", file=sys.stderr)
    failed = torch.tensor([1 if (rank == 0 and par is not None and not par["ok"]) else 0], device=dev)
    dist.all_reduce(failed)
    dist.destroy_process_group()
    if int(failed[0]):
        sys.exit(3)


# ------------------------------------------------------------------------------- CPU arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation on the box's host cores (rank 0 only): the unmodified
    scripts/3d_mapper.py when baseline/_ref/ was shipped, else the C restatement.  The path is single-threaded
    by construction (frames are sequential, the reference is pure Python), so it uses 1 core.  A step is a
    bounded sample of the GPU arm's step: its first `--ref-frames-per-step` frames."""
    if rank != 0:
        return
    from oracle.oracle import OracleMapper
    name = args.workload if (args.workload_given or args.gpus == 1) else "cfg4"
    ref_cls = None if args.ref_port else load_reference_class()
    kind = "reference" if ref_cls is not None else "port"
    make = ref_cls if ref_cls is not None else OracleMapper
    per_step = args.ref_frames_per_step or (1 if kind == "reference" else 40)
    gpu_step = args.frames_per_step or ((4000 if args.gpus == 1 else 2000) if name != "cfg3" else args.cfg3_frames_per_step)
    n_total = (args.steps + args.warmup) * gpu_step
    wl = Workload(name, n_total, args.seed, args.distinct_images, gpu_step)
    cfg = dict(wl.cfg)
    m = make(cfg)
    updates = 0
    with contextlib.redirect_stdout(io.StringIO()):
        for s in range(args.warmup):
            for f in range(s * gpu_step, s * gpu_step + per_step):
                m.process_sonar_image(wl.image(f), list(wl.pos[f]), list(wl.quat[f]))
        t0 = time.perf_counter()
        for s in range(args.warmup, args.warmup + args.steps):
            for f in range(s * gpu_step, s * gpu_step + per_step):
                st = m.process_sonar_image(wl.image(f), list(wl.pos[f]), list(wl.quat[f]))
                updates += st["num_occupied"] + st["num_free"]
        dt = time.perf_counter() - t0
    fps = args.steps * per_step / dt
    sample = (f"{args.steps} x {per_step} frames of {name}: the first {per_step} frame(s) of each {gpu_step}-frame step of the GPU arm "
              f"(same seed, same poses), through " +
              ("the unmodified reference (baseline/_ref/3d_mapper.py = scripts/3d_mapper.py), single-threaded pure Python + numpy"
               if kind == "reference" else "oracle/sonar_oracle.c (C restatement of scripts/3d_mapper.py)"))
    cb = {"value": fps, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample, "host_cpus": os.cpu_count(),
          "voxel_updates_per_s": updates / dt}
    if kind == "reference":
        # the C restatement beside it, same frames
        o = OracleMapper(cfg)
        t1 = time.perf_counter()
        n = 0
        for s in range(args.warmup, args.warmup + args.steps):
            for f in range(s * gpu_step, s * gpu_step + per_step):
                o.process_sonar_image(wl.image(f), wl.pos[f], wl.quat[f])
                n += 1
        cb["port_value"] = n / (time.perf_counter() - t1)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "voxel_updates_per_s": updates / dt,
        "config": {"workload": DESCR[name] + (f"; one map sharded by voxel-key hash over {args.gpus} GPUs" if args.gpus > 1 else ""),
                   "frames_per_step": per_step,
                   "note": f"bounded sample: the first {per_step} frame(s) of each of the GPU arm's {gpu_step}-frame steps"},
        "cpu_baseline": cb,
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(synthetic.CONFIGS))
    ap.add_argument("--frames-per-step", type=int, default=0, help="default: 4000 at N = 1 (an asynchronous batch is limited to 1 GiB of frames), 2000 at N > 1 (cfg3: --cfg3-frames-per-step)")
    ap.add_argument("--cfg3-frames-per-step", type=int, default=500)
    ap.add_argument("--ref-frames-per-step", type=int, default=0)
    ap.add_argument("--ref-port", action="store_true", help="--impl reference: time the C restatement even if baseline/_ref exists")
    ap.add_argument("--distinct-images", type=int, default=250)
    ap.add_argument("--cpu-frames", type=int, default=400)
    ap.add_argument("--parity-frames", type=int, default=400)
    ap.add_argument("--single-frames", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg3", action="store_true", help="N = 1: skip the cfg3 object of the line")
    ap.add_argument("--no-single-ref", action="store_true", help="N > 1: skip the single-GPU run of the same frames on rank 0")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer arm (profiling runs only)")
    ap.add_argument("--no-ablation", action="store_true", help="N > 1: skip the adaptive on/off pass")
    ap.add_argument("--no-adaptive", action="store_true",
                    help="ablation (BASELINE config 5): adaptive_update off -- every update is applied unscaled")
    ap.add_argument("--shard-mode", default="fused", choices=["fused", "replicate", "route"],
                    help="N > 1: how the sharded map moves data (sonar_3d_reconstruction_b200/sharded.py)")
    args = ap.parse_args()
    args.workload_given = args.workload is not None
    if args.workload is None:
        args.workload = "cfg2"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        run_sharded(args, rank, world, local_rank)
    else:
        run_single(args, local_rank)


if __name__ == "__main__":
    main()
