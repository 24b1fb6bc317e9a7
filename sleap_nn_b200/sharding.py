"""Frame sharding over the GPUs of one box.

Frames are independent, so the hot path shards with NO collective: rank r of W owns a contiguous
frame range, runs the 2-launch chain over it batch by batch and appends each batch's instances to
a packed per-rank table on the device (`snb_pack_instances`, no host sync).  The only exchange is
the end-of-shard gather of the variable-length results: one all-gather of the per-rank row / frame
counts, then one all-gather of each payload padded to the largest rank.  Works with any
`torch.distributed` backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).

The reference has no multi-GPU inference path; its single-process equivalent is the per-sample
list concatenation of `group_scored_batch` (sleap_nn/inference/streaming.py:187-255).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Iterator, List, Optional, Tuple

import torch
import torch.distributed as dist


# ----------------------------------------------------------------------------- partitioning
def frame_shard(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of `n_frames` for `rank`; the first n % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if n_frames < 0:
        raise ValueError("n_frames must be >= 0")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def batch_ranges(start: int, stop: int, batch: int) -> Iterator[Tuple[int, int]]:
    """[s, e) sub-ranges of at most `batch` frames covering [start, stop); the last one may be ragged."""
    if batch <= 0:
        raise ValueError("batch must be positive")
    s = start
    while s < stop:
        e = min(s + batch, stop)
        yield s, e
        s = e


# ----------------------------------------------------------------------------- host locality
def _parse_cpulist(text: str) -> List[int]:
    """"0-3,8,10-11" -> [0, 1, 2, 3, 8, 10, 11] (the format of sysfs `local_cpulist`)."""
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(device_index: int, sysfs_root: str = "/sys/bus/pci/devices") -> List[int]:
    """CPUs on the NUMA node the GPU's PCIe root hangs off (sysfs `local_cpulist`); [] when unknown."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"{sysfs_root}/{bdf}/local_cpulist") as f:
            return _parse_cpulist(f.read())
    except Exception:
        return []


def bind_host_to_gpu(device_index: int, min_cpus: int = 2) -> Optional[List[int]]:
    """Pin this process to the CPUs next to its GPU so that pinned staging buffers are first-touched on that node.

    With one rank per GPU each rank streams ~55 GB/s from pinned host memory; pages that sit on the other socket
    cross the inter-socket link first, which is what stops the host-buffer path scaling past 4 GPUs.  Returns the
    CPU list applied, or None when nothing was changed (no sysfs entry, or the intersection with the CPUs this
    process may use is smaller than `min_cpus` - a container cpuset must not be narrowed to nothing).
    """
    import os

    if not hasattr(os, "sched_getaffinity"):
        return None
    local = set(gpu_local_cpus(device_index))
    allowed = set(os.sched_getaffinity(0))
    use = sorted(local & allowed)
    if len(use) < min_cpus or len(use) == len(allowed):
        return None
    try:
        os.sched_setaffinity(0, use)
    except OSError:
        return None
    return use


# ----------------------------------------------------------------------------- packed results
class SharedBatchQueue:
    """One queue of batch indices for all ranks of a job whose inputs live in HOST memory.

    Frames shard with no data-path collective, but a FIXED equal split is paced by the slowest
    feed: on this pool's 8-GPU boxes the GPUs sit behind two host bridges of unequal speed (20.8 vs 35.7 GB/s per GPU
    when all eight copy at once, `tools/h2d_scaling_probe.py`), and the ranks behind the fast one finish early.  Here
    every rank pulls the next batch index when it has a free slot: one atomic `add` on torch.distributed's store per
    batch (~0.1 ms against ~6 ms for a cfg3 batch over PCIe) - 41.4 k instead of 35.0 k frames/s at N=8
    (`bench.py`'s e2e leg speaks the same protocol inline).  Without a process group it is a plain counter.

        q = SharedBatchQueue(n_batches)            # collective: every rank constructs it (one barrier)
        while (k := q.next()) is not None: submit(batch k)
    """

    def __init__(self, total: int, key: str = "snb_batch_queue", group=None):
        if total < 0:
            raise ValueError("total must be >= 0")
        self.total, self.key, self.taken = int(total), str(key), 0
        self._store = None
        self._local = 0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            self._store = dist.distributed_c10d._get_default_store()
            if dist.get_rank(group) == 0:
                self._store.set(self.key, "0")
            dist.barrier(group)  # nobody pulls before the counter exists

    def next(self) -> Optional[int]:
        """The next batch index nobody else has, or None when the queue is empty."""
        if self._store is None:
            k, self._local = self._local, self._local + 1
        else:
            k = int(self._store.add(self.key, 1)) - 1
        if k >= self.total:
            return None
        self.taken += 1
        return k


@dataclass
class PackedInstances:
    """Instances of a run of frames in packed form (row i belongs to frame `frame[i]`).

    counts[f - first_frame] = number of instances of frame f (zero-instance frames included).
    """

    first_frame: int
    counts: torch.Tensor   # (n_frames,) i32
    frame: torch.Tensor    # (rows,) i32 global frame index
    xy: torch.Tensor       # (rows, N, 2) f32, NaN = missing node
    val: torch.Tensor      # (rows, N) f32
    score: torch.Tensor    # (rows,) f32

    @property
    def n_frames(self) -> int:
        return int(self.counts.shape[0])

    @property
    def rows(self) -> int:
        return int(self.frame.shape[0])

    def to_lists(self):
        """Per-frame CPU lists like `PAFScorer.predict`'s first three outputs."""
        c = self.counts.cpu().tolist()
        xy, val, sc = self.xy.cpu(), self.val.cpu(), self.score.cpu()
        out, o = ([], [], []), 0
        for n in c:
            out[0].append(xy[o:o + n]); out[1].append(val[o:o + n]); out[2].append(sc[o:o + n])
            o += n
        return out


def _all_gather_var(t: torch.Tensor, sizes: List[int], group=None) -> torch.Tensor:
    """All-gather of a tensor whose dim-0 length differs per rank (`sizes[r]` rows on rank r)."""
    world = len(sizes)
    mx = max(sizes) if sizes else 0
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)


def gather_packed(local: PackedInstances, group=None) -> PackedInstances:
    """Concatenate every rank's packed results in rank (= frame) order on every rank.

    Two rounds: per-rank (rows, frames, first_frame) triples, then the five payload tensors padded to
    the largest rank.  With one rank (or no initialised process group) this is the identity.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    dev = local.frame.device
    meta = torch.tensor([local.rows, local.n_frames, local.first_frame], dtype=torch.int64, device=dev)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    rows = [int(m[0]) for m in metas]
    frames = [int(m[1]) for m in metas]
    firsts = [int(m[2]) for m in metas]
    for r in range(1, world):  # shards must tile the frame range in rank order
        if frames[r] and frames[r - 1] and firsts[r] != firsts[r - 1] + frames[r - 1]:
            raise ValueError(f"rank {r} shard starts at frame {firsts[r]}, expected {firsts[r - 1] + frames[r - 1]}")
    return PackedInstances(
        first_frame=min((f for f, n in zip(firsts, frames) if n), default=local.first_frame),
        counts=_all_gather_var(local.counts, frames, group),
        frame=_all_gather_var(local.frame, rows, group),
        xy=_all_gather_var(local.xy, rows, group),
        val=_all_gather_var(local.val, rows, group),
        score=_all_gather_var(local.score, rows, group),
    )


# ----------------------------------------------------------------------------- the per-rank runner
class ShardRunner:
    """Run `BottomUpPostproc` over this rank's frame shard and pack the results on the device.

    source(s, e) -> (cms, pafs) CUDA tensors for global frames [s, e) (e - s == batch except for the
    last, ragged batch, which is padded with frames that yield no peaks).  Nothing synchronises with
    the host until `finish()`.
    """

    def __init__(self, pipe, n_frames: int, rank: int = 0, world: int = 1, rows_cap: Optional[int] = None):
        from sleap_nn_b200 import _native as N

        self._N = N
        self.pipe, self.rank, self.world = pipe, rank, world
        self.start, self.stop = frame_shard(n_frames, rank, world)
        n_local = self.stop - self.start
        n_padded = -(-n_local // pipe.batch) * pipe.batch if n_local else 0
        dev, Nn = pipe.device, pipe.n_nodes
        # default: room for every frame filling its instance table (so the packed table cannot overflow before the
        # per-frame one does); pass rows_cap to bound the memory of very long shards of sparse frames
        self.rows_cap = int(rows_cap if rows_cap is not None else max(n_padded, 1) * pipe.caps["inst_cap"])
        with torch.cuda.device(dev):
            self.cursor = torch.zeros((3,), dtype=torch.int64, device=dev)
            self.o_xy = torch.empty((self.rows_cap, Nn, 2), dtype=torch.float32, device=dev)
            self.o_val = torch.empty((self.rows_cap, Nn), dtype=torch.float32, device=dev)
            self.o_score = torch.empty((self.rows_cap,), dtype=torch.float32, device=dev)
            self.o_frame = torch.empty((self.rows_cap,), dtype=torch.int32, device=dev)
            self.o_count = torch.zeros((max(n_padded, 1),), dtype=torch.int32, device=dev)
        self.launches = 0

    def run(self, source: Callable[[int, int], Tuple[torch.Tensor, torch.Tensor]]) -> "ShardRunner":
        with torch.cuda.device(self.pipe.device):  # the C side launches on the CURRENT device
            return self._run(source)

    def _run(self, source):
        N, pipe = self._N, self.pipe
        for s, e in batch_ranges(self.start, self.stop, pipe.batch):
            cms, pafs = source(s, e)
            if e - s != pipe.batch:  # ragged tail: pad with empty frames (no value above the threshold)
                pad = pipe.batch - (e - s)
                cms = torch.cat([cms, torch.zeros((pad,) + tuple(cms.shape[1:]), dtype=cms.dtype, device=cms.device)])
                pafs = torch.cat([pafs, torch.zeros((pad,) + tuple(pafs.shape[1:]), dtype=pafs.dtype, device=pafs.device)])
            res = pipe(cms, pafs)
            res.wait()
            N.check(N.lib.snb_pack_instances(
                N.ptr(res.n_instances), pipe.batch, pipe.caps["inst_cap"], pipe.n_nodes, N.ptr(res.instances),
                N.ptr(res.peak_scores), N.ptr(res.instance_scores), s, N.ptr(self.cursor), self.rows_cap,
                N.ptr(self.o_xy), N.ptr(self.o_val), N.ptr(self.o_score), N.ptr(self.o_frame), N.ptr(self.o_count),
                N.ptr(res.status), N.stream_ptr(pipe.device)), "snb_pack_instances")
            self.launches += pipe.launches_per_call + 1
        return self

    def finish(self) -> PackedInstances:
        """The one host sync of the shard: read the cursor and status, slice the packed tables."""
        N = self._N
        cur = self.cursor.cpu()
        status = int(self.pipe.buf["status"].item())
        if status:
            self.pipe.buf["status"].zero_()
        if status & N.STATUS_LSAP_INFEASIBLE:
            raise ValueError("cost matrix is infeasible")
        if status & N.STATUS_INSTANCE_OVERFLOW and int(cur[0]) > self.rows_cap:
            raise RuntimeError(f"sharded bottom-up run produced {int(cur[0])} instances, more than rows_cap={self.rows_cap} "
                               "(the packed per-rank table): pass a larger rows_cap to ShardRunner")
        if status:
            raise RuntimeError(f"sharded bottom-up run overflowed a fixed-capacity table (status 0x{status:x}): raise "
                               "peak_cap / cand_cap / match_cap / inst_cap of the pipeline")
        rows, n_local = int(cur[0]), self.stop - self.start
        return PackedInstances(self.start, self.o_count[:n_local].clone(), self.o_frame[:rows], self.o_xy[:rows],
                               self.o_val[:rows], self.o_score[:rows])
