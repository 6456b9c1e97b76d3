"""Voxel map sharded across the GPUs of one node (one process per GPU).

The reference has no distributed code; this is the B200-native scaling of its hot path.  The
map shards by a hash of the voxel key (`owner_of_keys`); every rank sees every frame.  Two modes:

* ``mode="route"`` -- each rank expands its slice of the processed beams into per-(voxel, frame)
  integer counts, the counts travel to the owning rank in one variable-size all-to-all per chunk
  of 16 frames (NCCL over NVLink/NVSwitch through torch.distributed; gloo works for CPU tests),
  and the owner merges them by integer addition and applies the chunk's frames in order.
* ``mode="fused"`` (default) -- the same partition as "route", but the exchange is part of the
  expansion kernel: its flush writes the records of voxels owned by other ranks straight into the
  owners' inboxes over NVLink peer memory (CUDA IPC mappings, set up once), device-side sequence
  flags order sources and owners, and the owner merges before it applies.  No host
  synchronisation and no separate all-to-all per chunk; the rank-local pipeline stays asynchronous.
  Only the per-frame counters are all-reduced, once per call.
* ``mode="replicate"`` -- each rank expands ALL beams but keeps only the samples whose
  voxel it owns, so its dedupe table already holds exactly its shard's counts and no exchange is
  needed; the cheap expansion arithmetic is repeated on every rank, the hash-table work (the
  expensive part, see profiles/README.md) and the voxel table are split N ways, and the rank-local
  pipeline stays fully asynchronous.  Only the per-frame counters are all-reduced, once per call.

Either way the merge is an integer sum, so the N-rank map is identical to the 1-rank map.

`ShardedSonarMapper` keeps the reference's method names (process_sonar_image,
get_point_cloud, reset_map); every rank must call them collectively with the same arguments.
The per-rank compute lives behind a small backend interface so that the exchange logic can be
exercised on CPU with a test double; the product backend is CUDA-only.
"""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

CHUNK_FRAMES = 16                  # include/sonar3d.h S3D_CHUNK_FRAMES
RECORD_WORDS = 1 + CHUNK_FRAMES    # include/sonar3d.h S3D_RECORD_WORDS: packed key + one counter per frame of the chunk
KEY_BIAS = 1 << 20
STATS_WORDS = 8       # int64 words per s3d_frame_stats (include/sonar3d.h)


def pack_keys(ijk: np.ndarray) -> np.ndarray:
    """(i, j, k) int triples -> the table's 63-bit packed key (21 bits per axis, biased)."""
    k = np.asarray(ijk, dtype=np.int64).reshape(-1, 3) + KEY_BIAS
    return (k[:, 0].astype(np.uint64) << np.uint64(42)) | (k[:, 1].astype(np.uint64) << np.uint64(21)) | \
        k[:, 2].astype(np.uint64)


def unpack_keys(packed: np.ndarray) -> np.ndarray:
    p = np.asarray(packed, dtype=np.uint64)
    m = np.uint64((1 << 21) - 1)
    out = np.stack([(p >> np.uint64(42)) & m, (p >> np.uint64(21)) & m, p & m], axis=1).astype(np.int64)
    return out - KEY_BIAS


def _mix64(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def owner_of_packed(packed: np.ndarray, world: int) -> np.ndarray:
    """Owning rank of each packed key: (mix64(key) >> 40) % world (same as the device code)."""
    return ((_mix64(packed) >> np.uint64(40)) % np.uint64(world)).astype(np.int64)


def owner_of_keys(ijk: np.ndarray, world: int) -> np.ndarray:
    return owner_of_packed(pack_keys(ijk), world)


def beam_slice(n_beams: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of the processed beams a rank expands."""
    return n_beams * rank // world, n_beams * (rank + 1) // world


# ---------------------------------------------------------------------------- collectives
class _SoloGroup:
    """world = 1 without torch.distributed."""
    rank, world = 0, 1


def _dist():
    import torch.distributed as dist
    return dist


class Exchange:
    """The collectives of the sharded path, on device tensors (NCCL) or CPU tensors (gloo)."""

    def __init__(self, group=None):
        if group is None or isinstance(group, _SoloGroup):
            self.rank, self.world, self.group, self.backend = 0, 1, None, "solo"
        else:
            dist = _dist()
            self.group = group
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
            self.backend = dist.get_backend(group)

    def all_to_all_records(self, send, send_counts: Sequence[int]):
        """send: [n, RECORD_WORDS] int64 tensor grouped by destination rank.  Returns (recv, recv_counts)."""
        import torch
        if self.world == 1:
            return send, list(send_counts)
        dist = _dist()
        sc = torch.tensor(list(send_counts), dtype=torch.int64, device=send.device)
        rc = torch.empty_like(sc)
        if self.backend == "nccl":
            dist.all_to_all_single(rc, sc, group=self.group)
            recv_counts = [int(v) for v in rc.tolist()]
            recv = torch.empty((sum(recv_counts), RECORD_WORDS), dtype=torch.int64, device=send.device)
            dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=list(send_counts),
                                   group=self.group)
            return recv, recv_counts
        # gloo has no all-to-all: pairwise exchange, counts first
        gathered = [torch.empty_like(sc) for _ in range(self.world)]
        dist.all_gather(gathered, sc, group=self.group)
        recv_counts = [int(gathered[src][self.rank]) for src in range(self.world)]
        recv = torch.empty((sum(recv_counts), RECORD_WORDS), dtype=torch.int64, device=send.device)
        s_off = np.concatenate([[0], np.cumsum(send_counts)])
        r_off = np.concatenate([[0], np.cumsum(recv_counts)])
        ops = []
        for peer in range(self.world):
            s_part = send[s_off[peer]:s_off[peer + 1]]
            r_part = recv[r_off[peer]:r_off[peer + 1]]
            if peer == self.rank:
                r_part.copy_(s_part)
                continue
            if len(s_part):
                ops.append(dist.P2POp(dist.isend, s_part.contiguous(), peer, group=self.group))
            if len(r_part):
                ops.append(dist.P2POp(dist.irecv, r_part, peer, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return recv, recv_counts

    def barrier(self):
        if self.world > 1:
            _dist().barrier(group=self.group)

    def all_reduce_sum(self, t):
        if self.world > 1:
            _dist().all_reduce(t, group=self.group)
        return t

    def all_reduce_max(self, t):
        if self.world > 1:
            dist = _dist()
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def all_gather_rows(self, t):
        """Concatenate per-rank [n_r, ...] tensors (n_r differs) along dim 0, same result on every rank."""
        import torch
        if self.world == 1:
            return t
        dist = _dist()
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        ns = [torch.empty_like(n) for _ in range(self.world)]
        dist.all_gather(ns, n, group=self.group)
        ns = [int(v) for v in ns]
        pad = max(ns + [1])
        buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        buf[:t.shape[0]] = t
        parts = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(parts, buf, group=self.group)
        return torch.cat([p[:k] for p, k in zip(parts, ns)], dim=0)

    def broadcast(self, t, src: int = 0):
        if self.world > 1:
            _dist().broadcast(t, src, group=self.group)
        return t


def upload_partition(n: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Frames [lo, hi) of a batch of n that `rank` uploads itself (the rest arrives by all-gather):
    equal parts of ceil(n / world) frames, the last ranks may get fewer or none.  -> (part, lo, hi)"""
    part = (n + world - 1) // world
    return part, min(rank * part, n), min((rank + 1) * part, n)


# ---------------------------------------------------------------------------- CUDA backend
class _DevView:
    """Zero-copy torch view of a raw device pointer (CUDA array interface)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class CudaShardBackend:
    """Per-rank compute on the GPU through the C-ABI (s3d_shard_expand / s3d_shard_apply)."""

    def __init__(self, mapper, rank: int, world: int):
        import torch
        self.torch = torch
        self.mapper = mapper
        self.native = mapper.octree._native
        self.device = torch.device("cuda", self.native.device)
        self.native.shard_config(rank, world)
        self.world = world

    def set_mode(self, mode: str, exchange=None, inbox_records: int = 1 << 22):
        self.native.route_enable(False)
        self.native.shard_filter(mode == "replicate" and self.world > 1)
        if mode == "fused" and self.world > 1:
            # collective: exchange blocks of all ranks, mapped into every rank through CUDA IPC
            t = self.torch
            mine = t.frombuffer(bytearray(self.native.route_export(inbox_records)), dtype=t.uint8).to(self.device)
            allh = exchange.all_gather_rows(mine[None, :]).cpu().numpy().tobytes()
            self.native.route_attach(allh, same_process=False)
            exchange.barrier()
            self.native.route_enable(True)
            exchange.barrier()

    def ingest_owned_dev(self, d_img, d_T):
        """replicate mode, inputs on the device: per-shard counters int64[n, 4] (device tensor)."""
        t = self.torch
        n = int(d_img.shape[0])
        st = t.empty((n, STATS_WORDS), dtype=t.int64, device=self.device)
        self.native.ingest_batch_dev(d_img.data_ptr(), n, d_T.data_ptr(), want_stats=False, stats_dev_ptr=st.data_ptr())
        self.native.sync()
        return st[:, :4].contiguous()

    def ingest_owned_host(self, images: np.ndarray, T: np.ndarray):
        """replicate mode, host inputs (copies overlap the kernels inside the library)."""
        st = self.native.ingest_batch(np.ascontiguousarray(images), T)
        arr = np.stack([st["num_occupied"], st["num_free"], st["num_voxels"], st["num_samples"]], axis=1)
        return self.torch.from_numpy(arr.astype(np.int64)).to(self.device)

    def ingest_owned_shared_upload(self, images: np.ndarray, T: np.ndarray, ex):
        """Host frames that every rank holds (fused / replicate modes): each rank uploads 1/world of
        them over its own PCIe link and the ranks all-gather the rest over NVLink, instead of every
        rank pulling every frame from host memory.  Returns the per-shard counters int64[n, 4]."""
        t = self.torch
        dist = _dist()
        n, H, W = images.shape
        w = ex.world
        part, lo, hi = upload_partition(n, w, ex.rank)
        if getattr(self, "_gather_buf", None) is None or tuple(self._gather_buf.shape) != (part * w, H, W):
            self._gather_buf = t.empty((part * w, H, W), dtype=t.uint8, device=self.device)
        buf = self._gather_buf
        mine = buf[ex.rank * part: (ex.rank + 1) * part]
        if hi > lo:
            mine[: hi - lo].copy_(t.from_numpy(images[lo:hi]), non_blocking=True)
        dist.all_gather_into_tensor(buf.view(-1), mine.reshape(-1), group=ex.group)
        d_T = t.from_numpy(np.ascontiguousarray(T, dtype=np.float64).reshape(-1, 16)).to(self.device, non_blocking=True)
        st = t.empty((n, STATS_WORDS), dtype=t.int64, device=self.device)
        t.cuda.current_stream(self.device).synchronize()          # the map's kernels run on the library's own streams
        self.native.ingest_batch_dev(buf.data_ptr(), n, d_T.data_ptr(), want_stats=False, stats_dev_ptr=st.data_ptr())
        self.native.sync()
        return st[:, :4].contiguous()

    def upload(self, images: np.ndarray, T: np.ndarray):
        t = self.torch
        return (t.from_numpy(np.ascontiguousarray(images)).to(self.device, non_blocking=False),
                t.from_numpy(np.ascontiguousarray(T, dtype=np.float64).reshape(-1, 16)).to(self.device))

    def expand(self, d_img, d_T, f0: int, g: int):
        t = self.torch
        H, W = d_img.shape[1], d_img.shape[2]
        # (the library zeroes the buffer itself, on its own stream: torch must not race it with a fill)
        st = t.empty((g, STATS_WORDS), dtype=t.int64, device=self.device)
        t.cuda.current_stream(self.device).synchronize()
        ptr, counts = self.native.shard_expand(d_img.data_ptr() + f0 * H * W, d_T.data_ptr() + f0 * 128, g, st.data_ptr())
        n = sum(counts)
        if n:
            send = t.as_tensor(_DevView(ptr, (n, RECORD_WORDS), "<i8"), device=self.device)
        else:
            send = t.empty((0, RECORD_WORDS), dtype=t.int64, device=self.device)
        return send, counts, st[:, 3].clone()

    def apply(self, recv, g: int):
        t = self.torch
        st = t.empty((g, STATS_WORDS), dtype=t.int64, device=self.device)
        t.cuda.current_stream(self.device).synchronize()          # the exchange wrote `recv` on torch's stream
        self.native.shard_apply(recv.data_ptr() if recv.shape[0] else 0, int(recv.shape[0]), g, st.data_ptr())
        return st[:, :3].clone()

    def count(self) -> int:
        return self.native.count()

    def export_occupied(self, min_probability: float):
        r = self.mapper.octree._export_occupied(min_probability)
        return r["xyz"], r["prob"]

    def export_classified(self, min_probability: float):
        """{class name: (centres float64[n,3], probabilities float64[n])} of this rank's shard."""
        cl = self.mapper.octree.get_all_voxels_classified(min_probability)
        return {name: (seq.points.reshape(-1, 3), seq.probabilities.reshape(-1)) for name, seq in cl.items()}

    def bounds(self):
        oc = self.mapper.octree
        return oc.min_bounds.copy(), oc.max_bounds.copy()

    def dump(self):
        return self.native.dump()

    def clear(self):
        self.mapper.octree.clear()

    def tensor(self, a: np.ndarray):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)


# ---------------------------------------------------------------------------- the mapper
class ShardedSonarMapper:
    """SonarTo3DMapper whose voxel map is sharded over the ranks of a process group.

    Collective: every rank constructs it with the same config and calls every method with the
    same arguments.  `group=None` means a single rank (no torch.distributed needed)."""

    def __init__(self, config: Optional[Dict[str, Any]] = None, group=None, backend_factory=None,
                 mode: str = "fused"):
        if mode not in ("fused", "replicate", "route"):
            raise ValueError("mode must be 'fused', 'replicate' or 'route'")
        self.mode = mode
        self.ex = Exchange(group)
        self.rank, self.world = self.ex.rank, self.ex.world
        if backend_factory is None:
            from .mapper import SonarTo3DMapper
            self.mapper = SonarTo3DMapper(config)
            self.backend = CudaShardBackend(self.mapper, self.rank, self.world)
        else:
            self.mapper, self.backend = backend_factory(config, self.rank, self.world)
        if mode == "fused":
            # records per (source, owner, chunk) inbox region; 16 bytes each, 16 regions deep: at most 4 GiB per rank
            # unless the config says otherwise (a region must hold every record one rank sends to one owner for a chunk)
            default = min(1 << 22, (4 << 30) // (16 * 16 * max(1, self.world)))
            self.backend.set_mode(mode, exchange=self.ex,
                                  inbox_records=int((config or {}).get("route_inbox_records", default)))
        else:
            self.backend.set_mode(mode)
        self.frame_count = 0
        self.processed_frame_count = 0
        self.total_processing_time = 0.0
        self.last_exchange_bytes = 0

    # the host-side attributes of the reference mapper are served by the wrapped mapper
    def __getattr__(self, name):
        return getattr(self.__dict__["mapper"], name)

    def process_sonar_images(self, polar_images: np.ndarray, robot_positions, robot_orientations) -> List[Dict[str, Any]]:
        import torch
        t0 = time.time()
        polar_images = np.asarray(polar_images)
        if polar_images.dtype != np.uint8 or polar_images.ndim != 3:
            raise TypeError("process_sonar_images expects uint8 images [n, range bins, bearings]")
        n, H, W = polar_images.shape
        m = self.mapper
        m._check_width(W)
        T = m.compose_transforms(robot_positions, robot_orientations)
        m._sync_device_config(H, W)
        if self.mode in ("replicate", "fused"):
            if self.world > 1 and hasattr(self.backend, "ingest_owned_shared_upload"):
                stats = self.ex.all_reduce_sum(self.backend.ingest_owned_shared_upload(polar_images, T, self.ex))
            else:
                stats = self.ex.all_reduce_sum(self.backend.ingest_owned_host(polar_images, T))
            self.last_exchange_bytes = 0
            return self._finish(stats, n, t0)
        d_img, d_T = self.backend.upload(polar_images, T)
        return self._finish(self.process_device_batch(d_img, d_T), n, t0)

    def process_device_batch(self, d_img, d_T):
        """Frames and 4x4 transforms already on this rank's device (as returned by backend.upload).
        Returns the globally reduced per-frame counters as an int64 tensor [n, 4]:
        num_occupied, num_free, num_voxels, num_samples."""
        import torch
        n = int(d_img.shape[0])
        if self.mode in ("replicate", "fused"):
            self.last_exchange_bytes = 0
            return self.ex.all_reduce_sum(self.backend.ingest_owned_dev(d_img, d_T))
        samples, applied = [], []
        self.last_exchange_bytes = 0
        for f0 in range(0, n, CHUNK_FRAMES):
            g = min(CHUNK_FRAMES, n - f0)
            send, counts, n_samp = self.backend.expand(d_img, d_T, f0, g)
            recv, recv_counts = self.ex.all_to_all_records(send, counts)
            self.last_exchange_bytes += 8 * RECORD_WORDS * (sum(counts) - counts[self.rank])
            applied.append(self.backend.apply(recv, g))
            samples.append(n_samp)
        stats = torch.cat([torch.cat(applied, dim=0), torch.cat(samples, dim=0)[:, None]], dim=1)   # [n, 4]
        return self.ex.all_reduce_sum(stats)

    def _finish(self, stats, n: int, t0: float) -> List[Dict[str, Any]]:
        rows = stats.cpu().tolist()
        dt = time.time() - t0
        per = dt / n if n else 0.0
        fc0, pc0, tot0 = self.frame_count, self.processed_frame_count, self.total_processing_time
        out = [{'frame_count': fc0 + f + 1, 'processed_count': pc0 + f + 1,
                'num_occupied': rows[f][0], 'num_free': rows[f][1], 'num_voxels': rows[f][2], 'num_samples': rows[f][3],
                'processing_time': per, 'avg_processing_time': (tot0 + per * (f + 1)) / (pc0 + f + 1)} for f in range(n)]
        self.frame_count, self.processed_frame_count = fc0 + n, pc0 + n
        self.total_processing_time = tot0 + dt
        return out

    def process_sonar_image(self, polar_image, robot_position, robot_orientation) -> Dict[str, Any]:
        if not isinstance(polar_image, np.ndarray):
            polar_image = np.array(polar_image)
        range_bins, bearing_bins = polar_image.shape
        return self.process_sonar_images(polar_image[None], [robot_position], [robot_orientation])[0]

    def num_voxels(self) -> int:
        import torch
        t = self.backend.tensor(np.array([self.backend.count()], dtype=np.int64))
        return int(self.ex.all_reduce_sum(t).cpu()[0])

    def get_point_cloud(self, include_free: bool = False) -> Dict[str, Any]:
        if include_free:
            # every rank classifies its shard on the device (:155-188), the three lists are all-gathered and the
            # bounds (:113-115) reduced: the same dict as the reference's :612-633 on every rank
            from .mapper import _PointProbList
            cl = self.backend.export_classified(self.mapper.min_probability)
            out = {}
            for name in ("occupied", "free", "unknown"):
                xyz, prob = cl[name]
                both = np.concatenate([np.asarray(xyz, dtype=np.float64).reshape(-1, 3), np.asarray(prob, dtype=np.float64).reshape(-1, 1)], axis=1)
                allp = self.ex.all_gather_rows(self.backend.tensor(both)).cpu().numpy()
                out[name] = _PointProbList(np.ascontiguousarray(allp[:, :3]), np.ascontiguousarray(allp[:, 3]))
            mn, mx = self.backend.bounds()
            ext = self.ex.all_reduce_max(self.backend.tensor(np.concatenate([-np.asarray(mn, dtype=np.float64),
                                                                             np.asarray(mx, dtype=np.float64)]))).cpu().numpy()
            dyn = bool(getattr(self.mapper, "dynamic_expansion", True))
            return {'occupied': out['occupied'], 'free': out['free'], 'unknown': out['unknown'],
                    'num_voxels': self.num_voxels(), 'num_occupied': len(out['occupied']), 'num_free': len(out['free']),
                    'num_unknown': len(out['unknown']), 'frame_count': self.frame_count,
                    'processed_count': self.processed_frame_count,
                    'bounds': {'min': -ext[:3] if dyn else None, 'max': ext[3:] if dyn else None}}
        xyz, prob = self.backend.export_occupied(self.mapper.min_probability)
        both = np.concatenate([xyz.reshape(-1, 3), prob.reshape(-1, 1)], axis=1)
        allp = self.ex.all_gather_rows(self.backend.tensor(both)).cpu().numpy()
        return {'points': allp[:, :3].copy() if len(allp) else np.empty((0, 3)),
                'probabilities': allp[:, 3].copy() if len(allp) else np.empty(0),
                'num_voxels': self.num_voxels(), 'num_occupied': len(allp),
                'frame_count': self.frame_count, 'processed_count': self.processed_frame_count}

    def gather_map(self) -> Tuple[np.ndarray, np.ndarray]:
        """Whole map on every rank: (keys int64[n,3], log-odds float64[n]) -- parity checks, checkpoints."""
        ijk, L = self.backend.dump()
        rows = np.concatenate([np.asarray(ijk, dtype=np.float64).reshape(-1, 3), np.asarray(L).reshape(-1, 1)], axis=1)
        allr = self.ex.all_gather_rows(self.backend.tensor(rows)).cpu().numpy()
        return allr[:, :3].astype(np.int64), allr[:, 3].copy()

    def reset_map(self):
        self.backend.clear()
        self.frame_count = 0
        self.processed_frame_count = 0
        self.total_processing_time = 0.0
        if self.rank == 0:
            print("Map reset")
