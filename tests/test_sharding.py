"""Frame sharding: partition properties, the end-of-shard gather of
variable-length results over world-size-2 gloo on CPU, and (GPU) the per-rank runner vs the
per-batch pipeline."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sleap_nn_b200.sharding import PackedInstances, SharedBatchQueue, batch_ranges, frame_shard, gather_packed


@pytest.mark.parametrize("n,world", [(0, 1), (1, 2), (7, 2), (64, 8), (100_000, 8), (100_003, 4), (5, 8)])
def test_frame_shard_tiles_the_range(n, world):
    spans = [frame_shard(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0 and a0 <= a1 and b0 <= b1
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_frame_shard_rejects_bad_ranks():
    with pytest.raises(ValueError):
        frame_shard(10, 2, 2)
    with pytest.raises(ValueError):
        frame_shard(-1, 0, 1)


@pytest.mark.parametrize("start,stop,batch", [(0, 0, 4), (3, 4, 64), (0, 128, 64), (5, 200, 64)])
def test_batch_ranges_cover_and_are_ragged_only_at_the_end(start, stop, batch):
    r = list(batch_ranges(start, stop, batch))
    assert [b for _, b in r[:-1]] == [a for a, _ in r[1:]]
    assert (r[0][0], r[-1][1]) == (start, stop) if r else start == stop
    assert all(e - s == batch for s, e in r[:-1]) and all(0 < e - s <= batch for s, e in r)


def _fake_packed(first, counts, n_nodes=3):
    """Deterministic packed results for frames [first, first+len(counts)): values encode (frame, row)."""
    frame = np.repeat(np.arange(first, first + len(counts)), counts).astype(np.int32)
    rows = len(frame)
    xy = (frame[:, None, None] * 100.0 + np.arange(n_nodes)[None, :, None] + np.array([0.25, 0.5])[None, None, :]).astype(np.float32)
    if rows:
        xy[::3, 0] = np.nan  # missing nodes travel as NaN
    val = (frame[:, None] + np.arange(n_nodes)[None, :] / 10.0).astype(np.float32)
    score = (frame * 0.5).astype(np.float32)
    t = torch.from_numpy
    return PackedInstances(first, t(np.asarray(counts, np.int32)), t(frame), t(xy), t(val), t(score))


def _gather_worker(rank, world, port, counts_by_rank, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_total = sum(len(c) for c in counts_by_rank)
        s, e = frame_shard(n_total, rank, world)
        assert e - s == len(counts_by_rank[rank])
        got = gather_packed(_fake_packed(s, counts_by_rank[rank]))
        torch.save({k: getattr(got, k) for k in ("first_frame", "counts", "frame", "xy", "val", "score")},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("counts_by_rank", [
    [[2, 0, 3, 1], [1, 1, 0]],        # ragged rows, 7 frames over 2 ranks (4 + 3)
    [[0, 0], [0]],                    # nothing found anywhere
    [[5], []],                        # one frame in total: rank 1 owns an empty shard
])
def test_gather_packed_world2_gloo(tmp_path, counts_by_rank):
    world = 2
    mp.spawn(_gather_worker, args=(world, _free_port(), counts_by_rank, str(tmp_path)), nprocs=world, join=True)
    flat = [c for cs in counts_by_rank for c in cs]
    want = _fake_packed(0, flat)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        assert got["first_frame"] == 0
        np.testing.assert_array_equal(got["counts"].numpy(), want.counts.numpy())
        np.testing.assert_array_equal(got["frame"].numpy(), want.frame.numpy())
        np.testing.assert_array_equal(got["xy"].numpy(), want.xy.numpy())  # NaN == NaN
        np.testing.assert_array_equal(got["val"].numpy(), want.val.numpy())
        np.testing.assert_array_equal(got["score"].numpy(), want.score.numpy())
        assert sorted(got["frame"].tolist()) == got["frame"].tolist()  # global frame order


def _queue_worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import time

        q = SharedBatchQueue(total)
        mine = []
        while (k := q.next()) is not None:
            mine.append(k)
            time.sleep(0.002 * (1 + 3 * rank))  # rank 1 is the slow feed: it must end up with fewer batches
        assert q.next() is None and q.taken == len(mine)
        torch.save(mine, os.path.join(out_dir, f"q{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [0, 1, 37])
def test_shared_batch_queue_world2_gloo(tmp_path, total):
    """Every batch index is handed out exactly once across the ranks, and the faster rank takes more of them."""
    world = 2
    mp.spawn(_queue_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(os.path.join(str(tmp_path), f"q{r}.pt")) for r in range(world)]
    assert sorted(got[0] + got[1]) == list(range(total))
    assert all(g == sorted(g) for g in got)
    if total >= 30:
        assert len(got[0]) > len(got[1]) > 0


def test_shared_batch_queue_without_a_process_group():
    q = SharedBatchQueue(3)
    assert [q.next(), q.next(), q.next(), q.next(), q.next()] == [0, 1, 2, None, None] and q.taken == 3


def test_gather_packed_is_identity_without_a_process_group():
    p = _fake_packed(4, [1, 2])
    assert gather_packed(p) is p
    lists = p.to_lists()
    assert [len(x) for x in lists[0]] == [1, 2] and lists[2][1].tolist() == [2.5, 2.5]


@pytest.mark.gpu
def test_shard_runner_matches_per_batch_pipeline():
    """Two 'ranks' run in turn on one GPU over a ragged 11-frame range; packed == per-batch results."""
    from sleap_nn_b200 import synthetic
    from sleap_nn_b200.pipeline import BottomUpPostproc
    from sleap_nn_b200.sharding import ShardRunner

    dev = torch.device("cuda", 0)
    Nn, n_inst, hw, stride, Bt, n_frames = 5, 2, (256, 256), 2, 4, 11
    edges = synthetic.chain_edges(Nn)
    poses = synthetic.random_poses(3, n_frames, n_inst, Nn, hw, edges, margin=60.0, step=24.0)
    cms, pafs = synthetic.render_batch(poses, hw, stride, edges, dev, seed=3)
    pipe = BottomUpPostproc(Nn, edges, Bt, tuple(cms.shape[-2:]), cms_stride=stride, pafs_stride=stride, device=dev)
    source = lambda s, e: (cms[s:e], pafs[s:e])
    parts = []
    for rank in range(2):
        parts.append(ShardRunner(pipe, n_frames, rank, 2).run(source).finish())
    assert [p.first_frame for p in parts] == [0, 6] and [p.n_frames for p in parts] == [6, 5]
    got = ([], [], [])
    for p in parts:
        for k, lst in enumerate(p.to_lists()):
            got[k].extend(lst)
        assert p.frame.cpu().tolist() == np.repeat(np.arange(p.first_frame, p.first_frame + p.n_frames),
                                                   p.counts.cpu().numpy()).tolist()
    # reference: the same frames through the plain per-batch call (last batch padded the same way)
    for s in range(0, n_frames, Bt):
        e = min(s + Bt, n_frames)
        c, p = cms[s:e], pafs[s:e]
        if e - s < Bt:
            c = torch.cat([c, torch.zeros((Bt - (e - s),) + tuple(c.shape[1:]), device=dev)])
            p = torch.cat([p, torch.zeros((Bt - (e - s),) + tuple(p.shape[1:]), device=dev)])
        inst, pv, sc = pipe(c, p).to_lists()
        for b in range(e - s):
            np.testing.assert_array_equal(got[0][s + b].numpy(), inst[b].numpy())
            np.testing.assert_array_equal(got[1][s + b].numpy(), pv[b].numpy())
            np.testing.assert_array_equal(got[2][s + b].numpy(), sc[b].numpy())
    assert sum(len(x) for x in got[0]) == n_frames * n_inst


def test_host_locality_helpers_are_safe_without_a_gpu():
    """bind_host_to_gpu must never narrow the process to nothing: no sysfs entry / no GPU -> no change."""
    import os

    from sleap_nn_b200.sharding import _parse_cpulist, bind_host_to_gpu, gpu_local_cpus

    assert _parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert _parse_cpulist("") == []
    assert _parse_cpulist("5") == [5]
    before = os.sched_getaffinity(0)
    assert gpu_local_cpus(0, sysfs_root="/nonexistent") == []
    assert bind_host_to_gpu(0) is None
    assert os.sched_getaffinity(0) == before
