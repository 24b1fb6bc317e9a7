"""PAF scoring + grouping - same API as sleap_nn/inference/ops/paf.py, computed by CUDA kernels.

Every public name of the reference module is provided with the same signature, defaults,
return dtypes and nesting (lists of per-sample tensors; `match_*` and `group_*` return CPU
tensors exactly as the reference does).  Host code below only marshals: it concatenates the
per-sample lists, allocates outputs and slices results; all arithmetic - candidate
enumeration, line sampling, scoring, the assignment problems and instance assembly - runs in
sleap_nn_b200/csrc/paf.cu through the C ABI.  Citations `paf.py:NN` are into the reference.
"""

from __future__ import annotations

from collections import deque
from typing import Dict, List, Optional, Sequence, Text, Tuple, Union

import attr
import attrs
import numpy as np
import torch

from sleap_nn_b200 import _native as N
from sleap_nn_b200.inference.utils import interp1d  # noqa: F401  (API parity with the reference import)


# ------------------------------------------------------------------------------ value types
@attrs.define(auto_attribs=True, frozen=True)
class PeakID:
    """(node_ind, peak_ind) key of a peak; paf.py:37-49."""

    node_ind: int
    peak_ind: int


@attrs.define(auto_attribs=True, frozen=True)
class EdgeType:
    """(src_node_ind, dst_node_ind) of a skeleton edge; paf.py:52-64."""

    src_node_ind: int
    dst_node_ind: int


@attrs.define(auto_attribs=True)
class EdgeConnection:
    """A matched connection (src_peak_ind, dst_peak_ind, score); paf.py:67-81."""

    src_peak_ind: int
    dst_peak_ind: int
    score: float


# ------------------------------------------------------------------------------ marshalling
_T_TABLES: Dict[Tuple[int, str], torch.Tensor] = {}


def _t_table(n_points: int, dev: torch.device) -> torch.Tensor:
    """torch.linspace(0, 1, n) evaluated on the HOST (defines the exact fp32 sample positions)."""
    key = (int(n_points), str(dev))
    if key not in _T_TABLES:
        _T_TABLES[key] = torch.linspace(0, 1, steps=int(n_points), dtype=torch.float32).to(dev)
    return _T_TABLES[key]


def _edges_tensor(edges, dev) -> torch.Tensor:
    """list of tuples / tensor (E,2) -> int32 device tensor (read at call time, never cached)."""
    if isinstance(edges, torch.Tensor):
        e = edges.detach().to(device=dev, dtype=torch.int32)
    else:
        e = torch.tensor([[int(a), int(b)] for a, b in edges], dtype=torch.int32, device=dev)
    return e.reshape(-1, 2).contiguous()


def _edge_type_tensor(edge_types, dev) -> torch.Tensor:
    return torch.tensor([[int(et.src_node_ind), int(et.dst_node_ind)] for et in edge_types], dtype=torch.int32,
                        device=dev).reshape(-1, 2)


class _Frames:
    """Concatenated per-sample tensors + CSR starts/counts on the device."""

    def __init__(self, tensors: Sequence[torch.Tensor], dev, dtype, tail=()):
        lens = [int(t.shape[0]) for t in tensors]
        self.lens = lens
        self.total = sum(lens)
        if tensors:
            self.data = torch.cat([torch.as_tensor(t).detach().to(device=dev, dtype=dtype).reshape((-1,) + tuple(tail))
                                   for t in tensors]).contiguous()
        else:
            self.data = torch.zeros((0,) + tuple(tail), dtype=dtype, device=dev)
        starts = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=starts[1:])
        self.starts_host = starts
        self.start = torch.tensor(starts[:-1], dtype=torch.int32, device=dev)
        self.count = torch.tensor(lens, dtype=torch.int32, device=dev)


def _status_check(status: torch.Tensor, what: str) -> None:
    s = int(status.item())
    if s & N.STATUS_LSAP_INFEASIBLE:
        raise ValueError("cost matrix is infeasible")  # scipy's message, paf.py:589
    if s & N.STATUS_BAD_INDEX:
        raise IndexError(f"{what}: index out of range")
    if s & N.STATUS_ASM_MISMATCH:  # make_predicted_instances' sanity assert (paf.py:866-873)
        raise AssertionError("both peaks of a connection should have been assigned to the same instance")
    if s & N.STATUS_ASM_MISSING:   # ... its dict lookup of a destination peak that is in no kept instance
        raise KeyError("destination peak of a scored connection is not assigned to an instance")
    if s:
        raise RuntimeError(f"{what}: device status 0x{s:x}")


def _prepare(chan_frames: _Frames, edges_t: torch.Tensor, n_nodes: int, dev):
    B = len(chan_frames.lens)
    E = int(edges_t.shape[0])
    node_start = torch.empty((B, n_nodes + 1), dtype=torch.int32, device=dev)
    node_peaks = torch.empty((max(chan_frames.total, 1),), dtype=torch.int32, device=dev)
    edge_off = torch.empty((B, E + 1), dtype=torch.int32, device=dev)
    match_off = torch.empty((B, E + 1), dtype=torch.int32, device=dev)
    N.check(
        N.lib.snb_paf_prepare(N.ptr(chan_frames.data), N.ptr(chan_frames.start), 0, N.ptr(chan_frames.count), B,
                              N.ptr(edges_t), int(n_nodes), E, N.ptr(node_start), N.ptr(node_peaks), N.ptr(edge_off),
                              N.ptr(match_off), N.stream_ptr(dev)),
        "snb_paf_prepare",
    )
    return node_start, node_peaks, edge_off, match_off


def _score_frames(pafs: Optional[torch.Tensor], peaks: Sequence[torch.Tensor], chans: Sequence[torch.Tensor], edges,
                  n_nodes: int, n_points: int, stride, max_edge_length: float, weight: float, dev):
    """prepare -> (host reads candidate counts) -> score.  Returns per-sample lists on `dev`."""
    B = len(chans)
    edges_t = _edges_tensor(edges, dev)
    E = int(edges_t.shape[0])
    chan_f = _Frames(chans, dev, torch.int32)
    xy_f = _Frames(peaks, dev, torch.float32, tail=(2,)) if peaks is not None else None
    node_start, node_peaks, edge_off, _ = _prepare(chan_f, edges_t, n_nodes, dev)
    m_per = edge_off[:, E].cpu().numpy().astype(np.int64) if B else np.zeros(0, np.int64)
    starts = np.zeros(B + 1, np.int64)
    np.cumsum(m_per, out=starts[1:])
    total = int(starts[-1])
    cand_edge = torch.empty((total,), dtype=torch.int32, device=dev)
    cand_epi = torch.empty((total, 2), dtype=torch.int64, device=dev)
    cand_score = torch.empty((total,), dtype=torch.float32, device=dev)
    if total:
        cand_start = torch.tensor(starts[:-1], dtype=torch.int32, device=dev)
        status = torch.zeros((1,), dtype=torch.int32, device=dev)
        if pafs is not None:
            pb, py, px, pc = pafs.stride()
            H, W = int(pafs.shape[1]), int(pafs.shape[2])
        else:
            pb = py = px = pc = 0
            H = W = 1
        N.check(
            N.lib.snb_paf_score_t(N.ptr(pafs), N.dtype_code(pafs.dtype) if pafs is not None else 0, pb, py, px, pc, H, W,
                                N.ptr(_t_table(n_points, dev)), int(n_points),
                                float(stride), float(max_edge_length), float(weight),
                                N.ptr(xy_f.data) if xy_f is not None else None, N.ptr(chan_f.start), 0, B,
                                N.ptr(edges_t), int(n_nodes), E, N.ptr(node_start), N.ptr(node_peaks),
                                N.ptr(edge_off), N.ptr(cand_start), 0, int(m_per.max()), N.ptr(cand_edge),
                                N.ptr(cand_epi), N.ptr(cand_score), N.ptr(status), N.stream_ptr(dev)),
            "snb_paf_score_t",
        )
    split = lambda t: [t[starts[b]:starts[b + 1]] for b in range(B)]
    return split(cand_edge), split(cand_epi), split(cand_score)


# ------------------------------------------------------------------------------ candidates / lines
def get_connection_candidates(
    peak_channel_inds_sample: torch.Tensor, skeleton_edges: torch.Tensor, n_nodes: int
) -> Tuple[torch.Tensor, torch.Tensor]:
    """All (source peak, destination peak) pairs per skeleton edge; paf.py:84-130.

    Returns `(edge_inds (M,) int32, edge_peak_inds (M,2) int64)` in edge-major, source-major,
    destination-minor order (identical to the reference whenever its argsort is stable).
    """
    dev = N.compute_device(peak_channel_inds_sample)
    out_dev = peak_channel_inds_sample.device
    with torch.cuda.device(dev):
        e, p, _ = _score_frames(None, None, [peak_channel_inds_sample], skeleton_edges, n_nodes, 1, 1, 0.0, 0.0, dev)
    return e[0].to(out_dev), p[0].to(out_dev)


def make_line_subs(
    peaks_sample: torch.Tensor,
    edge_peak_inds: torch.Tensor,
    edge_inds: torch.Tensor,
    n_line_points: int,
    pafs_stride: int,
    pafs_hw: tuple,
) -> torch.Tensor:
    """Line sample subscripts `(M, n_line_points, 2, 3)` int32 `[row, col, channel]`; paf.py:133-234."""
    dev = N.compute_device(peaks_sample)
    out_dev = peaks_sample.device
    pk = peaks_sample.detach().to(device=dev, dtype=torch.float32).contiguous()
    epi = edge_peak_inds.detach().to(device=dev, dtype=torch.int64).contiguous()
    ei = edge_inds.detach().to(device=dev, dtype=torch.int32).contiguous()
    M = int(epi.shape[0])
    out = torch.empty((M, int(n_line_points), 2, 3), dtype=torch.int32, device=dev)
    if M * int(n_line_points):
        with torch.cuda.device(dev):
            status = torch.zeros((1,), dtype=torch.int32, device=dev)
            N.check(
                N.lib.snb_line_subs(N.ptr(pk), int(pk.shape[0]), N.ptr(epi), N.ptr(ei), M,
                                    N.ptr(_t_table(n_line_points, dev)), int(n_line_points), float(pafs_stride),
                                    int(pafs_hw[0]), int(pafs_hw[1]), N.ptr(out), N.ptr(status), N.stream_ptr(dev)),
                "snb_line_subs",
            )
            if not peaks_sample.is_cuda:
                _status_check(status, "make_line_subs")
    return out.to(out_dev)


def get_paf_lines(
    pafs_sample: torch.Tensor,
    peaks_sample: torch.Tensor,
    edge_peak_inds: torch.Tensor,
    edge_inds: torch.Tensor,
    n_line_points: int,
    pafs_stride: int,
) -> torch.Tensor:
    """PAF vectors at the line samples, `(M, n_line_points, 2)`; paf.py:237-287.

    `pafs_sample` is `(height, width, 2 * n_edges)` and may be a strided view.
    """
    dev = N.compute_device(pafs_sample, peaks_sample)
    out_dev = pafs_sample.device
    pafs = pafs_sample.detach().to(device=dev, dtype=torch.float32)
    subs = make_line_subs(peaks_sample.to(dev), edge_peak_inds, edge_inds, n_line_points, pafs_stride,
                          pafs.shape[:2]).contiguous()
    n_sub = subs.numel() // 3
    out = torch.empty(tuple(subs.shape[:-1]), dtype=torch.float32, device=dev)
    if n_sub:
        with torch.cuda.device(dev):
            status = torch.zeros((1,), dtype=torch.int32, device=dev)
            py, px, pc = pafs.stride()
            N.check(
                N.lib.snb_paf_gather(N.ptr(pafs), py, px, pc, int(pafs.shape[0]), int(pafs.shape[1]),
                                     int(pafs.shape[2]), N.ptr(subs), n_sub, N.ptr(out), N.ptr(status),
                                     N.stream_ptr(dev)),
                "snb_paf_gather",
            )
            if not pafs_sample.is_cuda:
                _status_check(status, "get_paf_lines")
    return out.to(device=out_dev, dtype=pafs_sample.dtype)


def compute_distance_penalty(
    spatial_vec_lengths: torch.Tensor,
    max_edge_length: float,
    dist_penalty_weight: float = 1.0,
) -> torch.Tensor:
    """`min(max_edge_length / length - 1, 0) * weight`, any shape; paf.py:290-332."""
    dev = N.compute_device(spatial_vec_lengths)
    x = spatial_vec_lengths.detach().to(device=dev, dtype=torch.float32).contiguous()
    out = torch.empty_like(x)
    if x.numel():
        with torch.cuda.device(dev):
            N.check(N.lib.snb_distance_penalty(N.ptr(x), x.numel(), float(max_edge_length), float(dist_penalty_weight),
                                               N.ptr(out), N.stream_ptr(dev)), "snb_distance_penalty")
    return out.to(spatial_vec_lengths.device)


def score_paf_lines(
    paf_lines_sample: torch.Tensor,
    peaks_sample: torch.Tensor,
    edge_peak_inds_sample: torch.Tensor,
    max_edge_length: float,
    dist_penalty_weight: float = 1.0,
) -> torch.Tensor:
    """Mean PAF . unit(dst - src) over the line samples plus the distance penalty; paf.py:335-410."""
    dev = N.compute_device(paf_lines_sample, peaks_sample)
    out_dev = paf_lines_sample.device
    lines = paf_lines_sample.detach().to(device=dev, dtype=torch.float32).contiguous()
    pk = peaks_sample.detach().to(device=dev, dtype=torch.float32).contiguous()
    epi = edge_peak_inds_sample.detach().to(device=dev, dtype=torch.int64).contiguous()
    M = int(epi.shape[0])
    out = torch.empty((M,), dtype=torch.float32, device=dev)
    if M:
        with torch.cuda.device(dev):
            status = torch.zeros((1,), dtype=torch.int32, device=dev)
            N.check(
                N.lib.snb_score_lines(N.ptr(lines), N.ptr(pk), int(pk.shape[0]), N.ptr(epi), M, int(lines.shape[1]),
                                      float(max_edge_length), float(dist_penalty_weight), N.ptr(out), N.ptr(status),
                                      N.stream_ptr(dev)),
                "snb_score_lines",
            )
            if not paf_lines_sample.is_cuda:
                _status_check(status, "score_paf_lines")
    return out.to(out_dev)


def score_paf_lines_batch(
    pafs: torch.Tensor,
    peaks: torch.Tensor,
    peak_channel_inds: torch.Tensor,
    skeleton_edges: torch.Tensor,
    n_line_points: int,
    pafs_stride: int,
    max_edge_length_ratio: float,
    dist_penalty_weight: float,
    n_nodes: int,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Candidates + line scores for every sample of a batch; paf.py:413-497.

    `pafs` is `(n_samples, height, width, 2 * n_edges)` (typically a permuted view, read in
    place); `peaks` / `peak_channel_inds` are per-sample lists.  Returns three per-sample lists:
    edge_inds (int32), edge_peak_inds (int64, indices into that sample's peaks), line_scores.
    """
    max_edge_length = max_edge_length_ratio * max(pafs.shape[-1], pafs.shape[-2], pafs.shape[-3]) * pafs_stride
    dev = N.compute_device(pafs, *list(peaks))
    out_dev = pafs.device
    # fp16 / bf16 PAFs (an autocast backbone) are sampled in place: the taps are exact in fp32, no up-cast copy
    p = pafs.detach().to(device=dev) if pafs.dtype in N._DTYPES else pafs.detach().to(device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        e, ep, sc = _score_frames(p, list(peaks), list(peak_channel_inds), skeleton_edges, n_nodes, n_line_points,
                                  pafs_stride, max_edge_length, dist_penalty_weight, dev)
    if out_dev != dev:
        e, ep, sc = ([t.to(out_dev) for t in x] for x in (e, ep, sc))
    return e, ep, sc


# ------------------------------------------------------------------------------ matching
def _match_frames(edge_inds: Sequence[torch.Tensor], edge_peak_inds: Sequence[torch.Tensor],
                  line_scores: Sequence[torch.Tensor], n_edges: int):
    dev = N.compute_device(*list(line_scores))
    B = len(edge_inds)
    E = int(n_edges)
    with torch.cuda.device(dev):
        ce = _Frames(edge_inds, dev, torch.int32)
        cp = _Frames(edge_peak_inds, dev, torch.int64, tail=(2,))
        cs = _Frames(line_scores, dev, torch.float32)
        empty = lambda dt: [torch.zeros((0,), dtype=dt) for _ in range(B)]
        if B == 0 or E == 0 or ce.total == 0:
            return empty(torch.int32), empty(torch.int32), empty(torch.int32), empty(torch.float32)
        max_id = int(cp.data.max().item())
        status = torch.zeros((1,), dtype=torch.int32, device=dev)
        dims = torch.zeros((B * E, 2), dtype=torch.int32, device=dev)
        st = N.stream_ptr(dev)
        common = (N.ptr(ce.data), N.ptr(cp.data), N.ptr(cs.data), N.ptr(ce.start), N.ptr(ce.count), B, E, max_id)
        N.check(N.lib.snb_match_generic(0, *common, N.ptr(dims), None, None, None, None, None, 0, None, None, None,
                                        None, N.ptr(status), st), "snb_match_generic(count)")
        d = dims.cpu().numpy().astype(np.int64)
        cells = d[:, 0] * d[:, 1]
        n_match = np.minimum(d[:, 0], d[:, 1])
        cost_off = np.zeros(B * E + 1, np.int64)
        np.cumsum(cells, out=cost_off[1:])
        m_off = np.zeros(B * E + 1, np.int64)
        np.cumsum(n_match, out=m_off[1:])
        total = int(m_off[-1])
        max_dim = int(d.max()) if d.size else 0
        ws_bytes = N.lib.snb_lsap_workspace_bytes(max(max_dim, 1)) * B * E
        cost = torch.empty((max(int(cost_off[-1]), 1),), dtype=torch.float64, device=dev)
        cell_src = torch.empty((max(int(cost_off[-1]), 1),), dtype=torch.int32, device=dev)
        ws = torch.empty(((ws_bytes + 7) // 8,), dtype=torch.int64, device=dev)
        m_edge = torch.empty((total,), dtype=torch.int32, device=dev)
        m_src = torch.empty((total,), dtype=torch.int32, device=dev)
        m_dst = torch.empty((total,), dtype=torch.int32, device=dev)
        m_score = torch.empty((total,), dtype=torch.float32, device=dev)
        if total:
            cost_off_t = torch.tensor(cost_off[:-1], dtype=torch.int64, device=dev)
            m_start_t = torch.tensor(m_off[:-1], dtype=torch.int32, device=dev)
            N.check(N.lib.snb_match_generic(1, *common, N.ptr(dims), N.ptr(cost_off_t), N.ptr(cost), N.ptr(cell_src),
                                            N.ptr(m_start_t), N.ptr(ws), max(max_dim, 1), N.ptr(m_edge), N.ptr(m_src),
                                            N.ptr(m_dst), N.ptr(m_score), N.ptr(status), st),
                    "snb_match_generic(solve)")
        # match_candidates_* return CPU tensors (paf.py:549-551, 602-611)
        m_edge, m_src, m_dst, m_score = (t.cpu() for t in (m_edge, m_src, m_dst, m_score))
        _status_check(status, "match_candidates")
    per = m_off[::E]  # frame boundaries inside the (frame, edge)-ordered match list
    split = lambda t: [t[per[b]:per[b + 1]] for b in range(B)]
    return split(m_edge), split(m_src), split(m_dst), split(m_score)


def match_candidates_sample(
    edge_inds_sample: torch.Tensor,
    edge_peak_inds_sample: torch.Tensor,
    line_scores_sample: torch.Tensor,
    n_edges: int,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Per-edge optimal assignment of candidates (Hungarian semantics of scipy); paf.py:500-619.

    Returns CPU tensors `(match_edge_inds, match_src_peak_inds, match_dst_peak_inds,
    match_line_scores)`; the src/dst indices are ranks among the edge's distinct peak ids.
    Raises ValueError("cost matrix is infeasible") exactly where scipy would.
    """
    e, s, d, sc = _match_frames([edge_inds_sample], [edge_peak_inds_sample], [line_scores_sample], n_edges)
    return e[0], s[0], d[0], sc[0]


def match_candidates_batch(
    edge_inds: torch.Tensor,
    edge_peak_inds: torch.Tensor,
    line_scores: torch.Tensor,
    n_edges: int,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """`match_candidates_sample` over per-sample lists; paf.py:622-702."""
    return _match_frames(list(edge_inds), list(edge_peak_inds), list(line_scores), n_edges)


# ------------------------------------------------------------------------------ assembly
def toposort_edges(edge_types: List[EdgeType]) -> Tuple[int]:
    """Edge visiting order for assembly; paf.py:890-912 (one-time host setup).

    Root = first node, in order of appearance, with no incoming edge (networkx's first
    topological generation); breadth-first tree edges from that root, each reported by its
    first index in `edge_types`.  Edges outside that tree are absent, as in the reference.
    """
    edges = [(et.src_node_ind, et.dst_node_ind) for et in edge_types]
    succ: Dict[int, List[int]] = {}
    indeg: Dict[int, int] = {}
    for a, b in edges:
        for n in (a, b):
            succ.setdefault(n, [])
            indeg.setdefault(n, 0)
        if b not in succ[a]:
            succ[a].append(b)
            indeg[b] += 1
    roots = [n for n in succ if indeg[n] == 0]
    if not roots:
        if not succ:
            raise StopIteration  # next() on an empty topological sort, as the reference
        raise ValueError("Graph contains a cycle or graph changed during iteration")
    seen, todo, order = {roots[0]}, deque([roots[0]]), []
    while todo:
        a = todo.popleft()
        for b in succ[a]:
            if b not in seen:
                seen.add(b)
                order.append(edges.index((a, b)))
                todo.append(b)
    return tuple(order)


def _assemble_tables(peaks: _Frames, vals: _Frames, chans: _Frames, m_edge: _Frames, m_src: _Frames, m_dst: _Frames,
                     m_score: _Frames, n_nodes: int, sorted_edge_inds, edges_t: torch.Tensor, min_instance_peaks,
                     min_line_scores: float, dev):
    """snb_assemble into padded device tables (B, inst_cap, ...) + per-frame counts; no host synchronisation."""
    B = len(chans.lens)
    if isinstance(min_instance_peaks, float):
        min_instance_peaks = int(min_instance_peaks * n_nodes) if min_instance_peaks > 0 else 0  # paf.py:791-802
    node_start, node_peaks, _, _ = _prepare(chans, edges_t[:0], n_nodes, dev)
    sorted_t = torch.tensor([int(i) for i in sorted_edge_inds], dtype=torch.int32, device=dev)
    max_p = max(chans.lens) if chans.lens else 0
    inst_cap = max(max_p, 1)  # every instance owns >= 1 peak
    ws = torch.empty((B * 4 * max(max_p, 1),), dtype=torch.int32, device=dev)
    inst_xy = torch.empty((B, inst_cap, n_nodes, 2), dtype=torch.float32, device=dev)
    inst_val = torch.empty((B, inst_cap, n_nodes), dtype=torch.float32, device=dev)
    inst_score = torch.empty((B, inst_cap), dtype=torch.float32, device=dev)
    n_inst = torch.zeros((B,), dtype=torch.int32, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    N.check(
        N.lib.snb_assemble(N.ptr(peaks.data), N.ptr(vals.data), N.ptr(chans.data), N.ptr(chans.start), 0,
                           N.ptr(chans.count), B, N.ptr(node_start), N.ptr(node_peaks), int(n_nodes), N.ptr(edges_t),
                           N.ptr(sorted_t), int(sorted_t.numel()), N.ptr(m_edge.data), N.ptr(m_src.data),
                           N.ptr(m_dst.data), N.ptr(m_score.data), N.ptr(m_edge.start), 0, N.ptr(m_edge.count),
                           int(min_instance_peaks), float(min_line_scores), N.ptr(ws), max(max_p, 1), inst_cap,
                           N.ptr(inst_xy), N.ptr(inst_val), N.ptr(inst_score), N.ptr(n_inst), N.ptr(status),
                           N.stream_ptr(dev)),
        "snb_assemble",
    )
    return inst_xy, inst_val, inst_score, n_inst, status


def _assemble(peaks: _Frames, vals: _Frames, chans: _Frames, m_edge: _Frames, m_src: _Frames, m_dst: _Frames,
              m_score: _Frames, n_nodes: int, sorted_edge_inds, edges_t: torch.Tensor, min_instance_peaks,
              min_line_scores: float, dev):
    B = len(chans.lens)
    inst_xy, inst_val, inst_score, n_inst, status = _assemble_tables(peaks, vals, chans, m_edge, m_src, m_dst, m_score,
                                                                     n_nodes, sorted_edge_inds, edges_t,
                                                                     min_instance_peaks, min_line_scores, dev)
    n = n_inst.cpu().numpy()
    _status_check(status, "group_instances")
    xy, val, sc = inst_xy.cpu(), inst_val.cpu(), inst_score.cpu()
    return ([xy[b, : n[b]].clone() for b in range(B)], [val[b, : n[b]].clone() for b in range(B)],
            [sc[b, : n[b]].clone() for b in range(B)])


def _group_frames(peaks, peak_vals, peak_channel_inds, match_edge_inds, match_src_peak_inds, match_dst_peak_inds,
                  match_line_scores, n_nodes, sorted_edge_inds, edge_types, min_instance_peaks, min_line_scores,
                  tables: bool = False):
    dev = N.compute_device(*[t for t in peaks if isinstance(t, torch.Tensor)])
    as_t = lambda seq: [torch.as_tensor(x) for x in seq]
    with torch.cuda.device(dev):
        pk = _Frames(as_t(peaks), dev, torch.float32, tail=(2,))
        pv = _Frames(as_t(peak_vals), dev, torch.float32)
        pc = _Frames(as_t(peak_channel_inds), dev, torch.int32)
        me = _Frames(as_t(match_edge_inds), dev, torch.int32)
        ms = _Frames(as_t(match_src_peak_inds), dev, torch.int32)
        md = _Frames(as_t(match_dst_peak_inds), dev, torch.int32)
        msc = _Frames(as_t(match_line_scores), dev, torch.float32)
        edges_t = _edge_type_tensor(edge_types, dev)
        fn = _assemble_tables if tables else _assemble
        return fn(pk, pv, pc, me, ms, md, msc, int(n_nodes), sorted_edge_inds, edges_t, min_instance_peaks,
                  min_line_scores, dev)


def group_instances_sample(
    peaks_sample: torch.Tensor,
    peak_scores_sample: torch.Tensor,
    peak_channel_inds_sample: torch.Tensor,
    match_edge_inds_sample: torch.Tensor,
    match_src_peak_inds_sample: torch.Tensor,
    match_dst_peak_inds_sample: torch.Tensor,
    match_line_scores_sample: torch.Tensor,
    n_nodes: int,
    sorted_edge_inds: Tuple[int],
    edge_types: List[EdgeType],
    min_instance_peaks: int,
    min_line_scores: float = 0.25,
) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Group one sample's matched connections into instances; paf.py:915-1038.

    Returns numpy float32 arrays `(instances (I,N,2), peak_scores (I,N), instance_scores (I,))`,
    NaN where a node is missing.
    """
    xy, val, sc = _group_frames([peaks_sample], [peak_scores_sample], [peak_channel_inds_sample],
                                [match_edge_inds_sample], [match_src_peak_inds_sample], [match_dst_peak_inds_sample],
                                [match_line_scores_sample], n_nodes, sorted_edge_inds, edge_types, min_instance_peaks,
                                min_line_scores)
    return xy[0].numpy(), val[0].numpy(), sc[0].numpy()


def group_instances_batch(
    peaks: torch.Tensor,
    peak_vals: torch.Tensor,
    peak_channel_inds: torch.Tensor,
    match_edge_inds: torch.Tensor,
    match_src_peak_inds: torch.Tensor,
    match_dst_peak_inds: torch.Tensor,
    match_line_scores: torch.Tensor,
    n_nodes: int,
    sorted_edge_inds: Tuple[int],
    edge_types: List[EdgeType],
    min_instance_peaks: int,
    min_line_scores: float = 0.25,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """`group_instances_sample` over per-sample lists; returns lists of CPU tensors; paf.py:1041-1149."""
    return _group_frames(list(peaks), list(peak_vals), list(peak_channel_inds), list(match_edge_inds),
                         list(match_src_peak_inds), list(match_dst_peak_inds), list(match_line_scores), n_nodes,
                         sorted_edge_inds, edge_types, min_instance_peaks, min_line_scores)


def _flatten_connections(connections: Dict[EdgeType, List[EdgeConnection]]):
    """dict API -> the array form the assembly kernel consumes.

    Peak ids are arbitrary ints in the dict API; each (node, peak_ind) key gets a dense rank
    within its node (ascending peak_ind), which is what the kernel's (node, rank) addressing
    expects.  Pure bookkeeping - no grouping decision is taken on the host.
    """
    edge_types = list(connections.keys())
    by_node: Dict[int, set] = {}
    for et, conns in connections.items():
        for c in conns:
            by_node.setdefault(int(et.src_node_ind), set()).add(int(c.src_peak_ind))
            by_node.setdefault(int(et.dst_node_ind), set()).add(int(c.dst_peak_ind))
        by_node.setdefault(int(et.src_node_ind), set())
        by_node.setdefault(int(et.dst_node_ind), set())
    node_ids = sorted(by_node)
    node_slot = {n: i for i, n in enumerate(node_ids)}
    ranks = {n: {p: r for r, p in enumerate(sorted(by_node[n]))} for n in node_ids}
    chan, keys = [], []
    for n in node_ids:
        for p in sorted(by_node[n]):
            chan.append(node_slot[n])
            keys.append((n, p))
    me, ms, md, msc = [], [], [], []
    for k, (et, conns) in enumerate(connections.items()):
        for c in conns:
            me.append(k)
            ms.append(ranks[int(et.src_node_ind)][int(c.src_peak_ind)])
            md.append(ranks[int(et.dst_node_ind)][int(c.dst_peak_ind)])
            msc.append(float(c.score))
    edges = [(node_slot[int(et.src_node_ind)], node_slot[int(et.dst_node_ind)]) for et in edge_types]
    return edge_types, node_ids, keys, chan, edges, me, ms, md, msc


def assign_connections_to_instances(
    connections: Dict[EdgeType, List[EdgeConnection]],
    min_instance_peaks: Union[int, float] = 0,
    n_nodes: int = None,
) -> Dict[PeakID, int]:
    """Greedy partition of connections into instance ids; paf.py:705-820.

    Same dict-in / dict-out contract (insertion-ordered `{PeakID: instance_id}`); the
    partition itself is computed by the device assembly kernel.
    """
    edge_types, node_ids, keys, chan, edges, me, ms, md, msc = _flatten_connections(connections)
    if not keys:
        return {}
    if min_instance_peaks > 0 and isinstance(min_instance_peaks, float):
        if n_nodes is None:
            n_nodes = len(node_ids)
        min_instance_peaks = int(min_instance_peaks * n_nodes)
    dev = N.compute_device()
    with torch.cuda.device(dev):
        owner, order = _assemble_raw(chan, edges, me, ms, md, [0.0] * len(msc), len(node_ids),
                                     int(min_instance_peaks), dev)
    out: Dict[PeakID, int] = {}
    for i in order:
        if owner[i] >= 0:
            out[PeakID(keys[i][0], keys[i][1])] = int(owner[i])
    return out


def _assemble_raw(chan, edges, me, ms, md, msc, n_nodes, min_instance_peaks, dev):
    """Run the assembly kernel on one synthetic frame and read back (owner ids, insertion order)."""
    P = len(chan)
    i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=dev)
    chans = _Frames([i32(chan)], dev, torch.int32)
    xy = _Frames([torch.zeros((P, 2))], dev, torch.float32, tail=(2,))
    val = _Frames([torch.zeros((P,))], dev, torch.float32)
    edges_t = i32(edges).reshape(-1, 2)
    node_start, node_peaks, _, _ = _prepare(chans, edges_t[:0], n_nodes, dev)
    fr = lambda x, dt: _Frames([torch.tensor(x, dtype=dt)], dev, dt)
    m_e, m_s, m_d = fr(me, torch.int32), fr(ms, torch.int32), fr(md, torch.int32)
    m_sc = fr(msc, torch.float32)
    sorted_t = torch.arange(len(edges), dtype=torch.int32, device=dev)  # dict order IS the visiting order
    ws = torch.full((4 * max(P, 1),), -1, dtype=torch.int32, device=dev)
    inst_cap = max(P, 1)
    inst_xy = torch.empty((1, inst_cap, n_nodes, 2), dtype=torch.float32, device=dev)
    inst_val = torch.empty((1, inst_cap, n_nodes), dtype=torch.float32, device=dev)
    inst_score = torch.empty((1, inst_cap), dtype=torch.float32, device=dev)
    n_inst = torch.zeros((1,), dtype=torch.int32, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    N.check(
        N.lib.snb_assemble(N.ptr(xy.data), N.ptr(val.data), N.ptr(chans.data), N.ptr(chans.start), 0,
                           N.ptr(chans.count), 1, N.ptr(node_start), N.ptr(node_peaks), int(n_nodes), N.ptr(edges_t),
                           N.ptr(sorted_t), int(sorted_t.numel()), N.ptr(m_e.data), N.ptr(m_s.data), N.ptr(m_d.data),
                           N.ptr(m_sc.data), N.ptr(m_e.start), 0, N.ptr(m_e.count), int(min_instance_peaks),
                           float("-inf"), N.ptr(ws), max(P, 1), inst_cap, N.ptr(inst_xy), N.ptr(inst_val),
                           N.ptr(inst_score), N.ptr(n_inst), N.ptr(status), N.stream_ptr(dev)),
        "snb_assemble",
    )
    w = ws.cpu().numpy()
    # assign_connections_to_instances has no sanity check of its own: that one lives in make_predicted_instances
    status &= ~(N.STATUS_ASM_MISMATCH | N.STATUS_ASM_MISSING)
    _status_check(status, "assign_connections_to_instances")
    owner, order, id_count, id_rank = w[:P], w[P:2 * P], w[2 * P:3 * P], w[3 * P:4 * P]
    kept = np.array([o if (o >= 0 and id_rank[o] >= 0) else -1 for o in owner])
    n_ord = int(np.count_nonzero(owner >= 0))
    return kept, [int(i) for i in order[:n_ord]]


def make_predicted_instances(
    peaks: np.array,
    peak_scores: np.array,
    connections: List[EdgeConnection],
    instance_assignments: Dict[PeakID, int],
) -> Tuple[np.array, np.array, np.array]:
    """Scatter assigned peaks into NaN-filled arrays and sum edge scores; paf.py:823-887.

    `peaks` / `peak_scores` are node-grouped; `instance_assignments` as returned by
    `assign_connections_to_instances` (re-numbered in place to contiguous ids, as the reference).
    The dicts are flattened on the host (bookkeeping only); the NaN fill, the scatter and the
    fp32 running score sums run in `snb_scatter_instances`.
    """
    ids = sorted(set(instance_assignments.values()))
    rank = {inst: i for i, inst in enumerate(ids)}
    for pid in instance_assignments:
        instance_assignments[pid] = rank[instance_assignments[pid]]
    conn_inst, conn_score = [], []
    for et, conns in connections.items():
        for c in conns:
            src = PeakID(node_ind=et.src_node_ind, peak_ind=c.src_peak_ind)
            ind = instance_assignments.get(src, -1)
            if ind >= 0:
                assert ind == instance_assignments[PeakID(node_ind=et.dst_node_ind, peak_ind=c.dst_peak_ind)]
            conn_inst.append(ind)
            conn_score.append(float(c.score))
    n_nodes, n_inst = len(peaks), len(ids)
    a_xy = [np.asarray(peaks[pid.node_ind][pid.peak_ind], dtype=np.float32) for pid in instance_assignments]
    a_val = [np.float32(peak_scores[pid.node_ind][pid.peak_ind]) for pid in instance_assignments]
    a_inst = [int(i) for i in instance_assignments.values()]
    a_node = [int(pid.node_ind) for pid in instance_assignments]
    inst = np.full((n_inst, n_nodes, 2), np.nan, dtype="float32")
    pv = np.full((n_inst, n_nodes), np.nan, dtype="float32")
    scores = np.full((n_inst,), 0.0, dtype="float32")
    if n_inst == 0:
        return inst, pv, scores
    dev = N.compute_device()
    with torch.cuda.device(dev):
        t = lambda x, dt, shape: torch.tensor(np.asarray(x).reshape(shape), dtype=dt, device=dev)
        xy_t = t(a_xy, torch.float32, (-1, 2))
        val_t = t(a_val, torch.float32, (-1,))
        inst_t, node_t = t(a_inst, torch.int32, (-1,)), t(a_node, torch.int32, (-1,))
        ci_t, cs_t = t(conn_inst, torch.int32, (-1,)), t(conn_score, torch.float32, (-1,))
        o_xy = torch.empty((n_inst, n_nodes, 2), dtype=torch.float32, device=dev)
        o_val = torch.empty((n_inst, n_nodes), dtype=torch.float32, device=dev)
        o_sc = torch.empty((n_inst,), dtype=torch.float32, device=dev)
        N.check(
            N.lib.snb_scatter_instances(N.ptr(xy_t), N.ptr(val_t), N.ptr(inst_t), N.ptr(node_t), len(a_inst),
                                        N.ptr(ci_t), N.ptr(cs_t), len(conn_inst), n_inst, n_nodes, N.ptr(o_xy),
                                        N.ptr(o_val), N.ptr(o_sc), N.stream_ptr(dev)),
            "snb_scatter_instances",
        )
        return o_xy.cpu().numpy(), o_val.cpu().numpy(), o_sc.cpu().numpy()


# ------------------------------------------------------------------------------ PAFScorer
@attrs.define
class PAFScorer:
    """Scoring pipeline based on part affinity fields; paf.py:1152-1532.

    A mutable attrs holder of the grouping knobs and skeleton tables.  Callers overwrite
    fields after construction (tests, the legacy CLI), so every method reads the attributes
    at call time; nothing is cached on the device.
    """

    part_names: List[Text]
    edges: List[Tuple[Text, Text]]
    pafs_stride: int
    max_edge_length_ratio: float = 0.25
    dist_penalty_weight: float = 1.0
    n_points: int = 10
    min_instance_peaks: Union[int, float] = 0
    min_line_scores: float = 0.25
    edge_inds: List[Tuple[int, int]] = attr.ib(init=False)
    edge_types: List[EdgeType] = attr.ib(init=False)
    n_nodes: int = attr.ib(init=False)
    n_edges: int = attr.ib(init=False)
    sorted_edge_inds: Tuple[int] = attr.ib(init=False)

    def __attrs_post_init__(self):
        """Cache the computed skeleton attributes, as paf.py:1220-1232."""
        self.edge_inds = [(self.part_names.index(src), self.part_names.index(dst)) for (src, dst) in self.edges]
        self.edge_types = [EdgeType(src_node, dst_node) for src_node, dst_node in self.edge_inds]
        self.n_nodes = len(self.part_names)
        self.n_edges = len(self.edges)
        self.sorted_edge_inds = toposort_edges(self.edge_types)

    @classmethod
    def from_config(
        cls,
        config,
        max_edge_length_ratio: float = 0.25,
        dist_penalty_weight: float = 1.0,
        n_points: int = 10,
        min_instance_peaks: Union[int, float] = 0,
        min_line_scores: float = 0.25,
    ) -> "PAFScorer":
        """Build from a `MultiInstanceConfig`-style head config (attribute access); paf.py:1240-1281."""
        return cls(
            part_names=config.confmaps.part_names,
            edges=config.pafs.edges,
            pafs_stride=config.pafs.output_stride,
            max_edge_length_ratio=max_edge_length_ratio,
            dist_penalty_weight=dist_penalty_weight,
            n_points=n_points,
            min_instance_peaks=min_instance_peaks,
            min_line_scores=min_line_scores,
        )

    def score_paf_lines(self, pafs: torch.Tensor, peaks: torch.Tensor, peak_channel_inds: torch.Tensor):
        """Wrapper for `score_paf_lines_batch`; paf.py:1283-1330."""
        return score_paf_lines_batch(pafs, peaks, peak_channel_inds, self.edge_inds, self.n_points, self.pafs_stride,
                                     self.max_edge_length_ratio, self.dist_penalty_weight, self.n_nodes)

    def match_candidates(self, edge_inds: torch.Tensor, edge_peak_inds: torch.Tensor, line_scores: torch.Tensor):
        """Wrapper for `match_candidates_batch`; paf.py:1332-1384."""
        return match_candidates_batch(edge_inds, edge_peak_inds, line_scores, self.n_edges)

    def group_instances(self, peaks, peak_vals, peak_channel_inds, match_edge_inds, match_src_peak_inds,
                        match_dst_peak_inds, match_line_scores):
        """Wrapper for `group_instances_batch`; paf.py:1386-1467."""
        return group_instances_batch(peaks, peak_vals, peak_channel_inds, match_edge_inds, match_src_peak_inds,
                                     match_dst_peak_inds, match_line_scores, self.n_nodes, self.sorted_edge_inds,
                                     self.edge_types, self.min_instance_peaks, min_line_scores=self.min_line_scores)

    def predict(self, pafs: torch.Tensor, peaks: torch.Tensor, peak_vals: torch.Tensor,
                peak_channel_inds: torch.Tensor):
        """Score -> match -> group; returns the reference's 6-tuple; paf.py:1469-1532."""
        edge_inds, edge_peak_inds, line_scores = self.score_paf_lines(pafs, peaks, peak_channel_inds)
        m_e, m_s, m_d, m_sc = self.match_candidates(edge_inds, edge_peak_inds, line_scores)
        inst, pv, isc = self.group_instances(peaks, peak_vals, peak_channel_inds, m_e, m_s, m_d, m_sc)
        return inst, pv, isc, edge_inds, edge_peak_inds, line_scores
