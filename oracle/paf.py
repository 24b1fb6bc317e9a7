"""CPU oracle for PAF scoring, matching and instance assembly.  TEST INFRASTRUCTURE ONLY.

Restates sleap_nn/inference/ops/paf.py (cited below as `paf.py:NN`) and the two-knot
use of sleap_nn/inference/utils.py:interp1d (`utils.py:NN`) with torch CPU / numpy ops.
Never imported by the product path.  Parity is PINNED against the reference's golden
vectors (tests/golden) and, in the build container, against the reference itself.

Third-party arithmetic on this path (absent from /root/reference):
  * scipy.optimize.linear_sum_assignment (paf.py:589) - pyproject lists "scipy"
    unpinned; this image has scipy 1.18.1.  The oracle calls scipy itself;
    `lsap_jv` below restates its published algorithm (Crouse 2016, shortest
    augmenting path with dual variables, rows ascending) and is pinned against scipy.
  * networkx topological_sort / bfs_edges (paf.py:908-910), image has 3.6.1;
    `toposort_edge_order` restates it and is pinned against networkx in the tests.
"""

from __future__ import annotations

from collections import deque
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

F32_EPS = float(torch.finfo(torch.float32).eps)


# ------------------------------------------------------------------ candidates
def connection_candidates(channel_inds: torch.Tensor, edges, n_nodes: int):
    """All (src peak, dst peak) pairs per skeleton edge.  paf.py:84-130.

    Canonical order: edge-major, then source peak ascending, then destination peak
    ascending (a STABLE grouping by channel).  The reference uses torch.argsort, which
    is stable for n <= 16 on this build and implementation-defined above that
    (SURVEY section 7); downstream matching is invariant to the order inside an edge.
    Returns edge_inds (M,) i32 and edge_peak_inds (M,2) i64.
    """
    ch = torch.as_tensor(channel_inds).to(torch.int64)
    by_node = [torch.nonzero(ch == k)[:, 0] for k in range(n_nodes)]
    e_out, p_out = [], []
    for k, (s, d) in enumerate(_edge_list(edges)):
        src, dst = by_node[s], by_node[d]
        pairs = torch.stack(
            [src.repeat_interleave(dst.numel()), dst.repeat(src.numel())], dim=1
        )
        e_out.append(torch.full((pairs.shape[0],), k, dtype=torch.int32))
        p_out.append(pairs)
    if not e_out:
        return torch.zeros(0, dtype=torch.int32), torch.zeros((0, 2), dtype=torch.int64)
    return torch.cat(e_out), torch.cat(p_out)


def _edge_list(edges) -> List[Tuple[int, int]]:
    if isinstance(edges, torch.Tensor):
        return [(int(a), int(b)) for a, b in edges.tolist()]
    return [(int(a), int(b)) for a, b in edges]


# ------------------------------------------------------------------ line sampling
def line_subscripts(peaks, edge_peak_inds, edge_inds, n_points: int, stride, paf_hw):
    """Rounded, clipped [row, col, channel] subscripts of the line samples.

    paf.py:133-234 with utils.py:29-130 specialised to two knots at x = 0, 1:
      slope = (dst - src) / (eps + 1)      (eps added to an int64 1 -> fp32 1.00000012)
      val_k = src + slope * t_k,  t = linspace(0, 1, n)       (each op rounded to fp32)
      q_k   = round_half_even(val_k / stride) as int, then rows/cols clipped.
    Returns (M, n_points, 2, 3) int32; channel = 2*edge and 2*edge + 1.
    """
    peaks = torch.as_tensor(peaks, dtype=torch.float32)
    epi = torch.as_tensor(edge_peak_inds).to(torch.int64)
    src, dst = peaks[epi[:, 0]], peaks[epi[:, 1]]  # (M,2) x,y
    t = torch.linspace(0, 1, steps=n_points)
    one_eps = torch.tensor(F32_EPS, dtype=torch.float32) + 1
    slope = (dst - src) / one_eps
    val = src[:, :, None] + slope[:, :, None] * t[None, None, :]  # (M,2,n) x,y
    q = (val / stride).round().to(torch.int32)
    h, w = int(paf_hw[0]), int(paf_hw[1])
    rows = q[:, 1, :].clamp(0, h - 1)
    cols = q[:, 0, :].clamp(0, w - 1)
    e = torch.as_tensor(edge_inds).to(torch.int32).view(-1, 1).expand(-1, n_points)
    first = torch.stack([rows, cols, 2 * e], dim=-1)
    second = torch.stack([rows, cols, 2 * e + 1], dim=-1)
    return torch.stack([first, second], dim=2)


def paf_lines(pafs_hwc, peaks, edge_peak_inds, edge_inds, n_points: int, stride):
    """PAF vectors at the line samples, (M, n_points, 2).  paf.py:237-287."""
    subs = line_subscripts(peaks, edge_peak_inds, edge_inds, n_points, stride, pafs_hwc.shape[:2]).long()
    return pafs_hwc[subs[..., 0], subs[..., 1], subs[..., 2]]


def distance_penalty(lengths: torch.Tensor, max_edge_length: float, weight: float = 1.0):
    """min(max_len / len - 1, 0) * weight.  paf.py:290-332."""
    return ((max_edge_length / lengths) - 1).clamp(max=0) * weight


def score_lines(lines, peaks, edge_peak_inds, max_edge_length: float, weight: float = 1.0):
    """Mean over samples of paf . unit(dst - src), plus the distance penalty.  paf.py:335-410."""
    peaks = torch.as_tensor(peaks)
    epi = torch.as_tensor(edge_peak_inds).to(torch.int64)
    vec = peaks[epi[:, 1]] - peaks[epi[:, 0]]
    length = torch.linalg.vector_norm(vec, dim=1, keepdim=True)
    unit = vec / length
    dots = lines[..., 0] * unit[:, None, 0] + lines[..., 1] * unit[:, None, 1]
    return dots.mean(dim=1) + distance_penalty(length, max_edge_length, weight)[:, 0]


def max_edge_length_for(pafs_bhwc_shape, stride, ratio: float) -> float:
    """ratio * max(H, W, 2E) * stride - the max really includes the channel dim.  paf.py:457-461."""
    return ratio * max(pafs_bhwc_shape[-1], pafs_bhwc_shape[-2], pafs_bhwc_shape[-3]) * stride


def score_lines_batch(pafs_bhwc, peaks, channel_inds, edges, n_points, stride, ratio, weight, n_nodes):
    """Per-sample candidates -> line samples -> scores.  paf.py:413-497."""
    max_len = max_edge_length_for(pafs_bhwc.shape, stride, ratio)
    out_e, out_p, out_s = [], [], []
    for b in range(pafs_bhwc.shape[0]):
        e, p = connection_candidates(channel_inds[b], edges, n_nodes)
        ln = paf_lines(pafs_bhwc[b], peaks[b], p, e, n_points, stride)
        out_e.append(e)
        out_p.append(p)
        out_s.append(score_lines(ln, peaks[b], p, max_len, weight))
    return out_e, out_p, out_s


# ---------------------------------------------------------------------- matching
def lsap_jv(cost: np.ndarray):
    """Rectangular linear-sum assignment, float64, restating scipy's solver.

    Shortest-augmenting-path with duals (Crouse 2016) exactly as
    scipy.optimize.linear_sum_assignment runs it: a tall matrix is transposed, the
    free-column list is filled in reverse, among equal reduced costs the LAST free
    column in list order that is unassigned wins (else the first minimum), NaN or
    -inf entries are invalid and an all-inf row is infeasible (ValueError).  Returns
    (rows ascending, cols).
    """
    c = np.asarray(cost, dtype=np.float64)
    nr, nc = c.shape
    if nr == 0 or nc == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    transposed = nc < nr
    if transposed:
        c = c.T.copy()
        nr, nc = nc, nr
    if np.isnan(c).any() or np.isneginf(c).any():
        raise ValueError("matrix contains invalid numeric entries")
    u = np.zeros(nr)
    v = np.zeros(nc)
    col4row = -np.ones(nr, np.int64)
    row4col = -np.ones(nc, np.int64)
    for cur in range(nr):
        spc = np.full(nc, np.inf)
        path = -np.ones(nc, np.int64)
        in_sr = np.zeros(nr, bool)
        in_sc = np.zeros(nc, bool)
        free = list(range(nc - 1, -1, -1))
        i, sink, min_val = cur, -1, 0.0
        while sink == -1:
            in_sr[i] = True
            lowest, pick = np.inf, -1
            for it, j in enumerate(free):
                r = min_val + c[i, j] - u[i] - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest, pick = spc[j], it
            min_val = lowest
            if min_val == np.inf:
                raise ValueError("cost matrix is infeasible")
            j = free[pick]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            in_sc[j] = True
            free[pick] = free[-1]
            free.pop()
        u[cur] += min_val
        for r_ in range(nr):
            if in_sr[r_] and r_ != cur:
                u[r_] += min_val - spc[col4row[r_]]
        for j in range(nc):
            if in_sc[j]:
                v[j] -= min_val - spc[j]
        j = sink
        while True:
            i = path[j]
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur:
                break
    if transposed:
        order = np.argsort(col4row, kind="stable")
        return col4row[order].astype(np.int64), order.astype(np.int64)
    return np.arange(nr, dtype=np.int64), col4row


def match_sample(edge_inds, edge_peak_inds, scores, n_edges: int, solver=None):
    """Per-edge optimal assignment on cost = -score (NaN -> +inf).  paf.py:500-619.

    Returned src/dst indices are RANKS among the distinct peak ids of that edge's
    candidates, not peak ids.  `solver` defaults to scipy (what the reference calls).
    """
    if solver is None:
        from scipy.optimize import linear_sum_assignment as solver
    e = torch.as_tensor(edge_inds).cpu()
    p = torch.as_tensor(edge_peak_inds).cpu().to(torch.int64)
    s = torch.as_tensor(scores).cpu()
    me, ms, md, msc = [], [], [], []
    for k in range(n_edges):
        sel = torch.nonzero(e == k)[:, 0]
        pk, sk = p[sel], s[sel]
        src_ids, src_rank = torch.unique(pk[:, 0], return_inverse=True)
        dst_ids, dst_rank = torch.unique(pk[:, 1], return_inverse=True)
        cost = np.full((src_ids.numel(), dst_ids.numel()), np.inf, dtype=np.float32)
        cost[src_rank.numpy(), dst_rank.numpy()] = -sk.numpy().astype(np.float32)
        cost[np.isnan(cost)] = np.inf
        r, c = solver(cost)
        me.append(torch.full((len(r),), k, dtype=torch.int32))
        ms.append(torch.as_tensor(np.asarray(r), dtype=torch.int32))
        md.append(torch.as_tensor(np.asarray(c), dtype=torch.int32))
        msc.append(torch.as_tensor(-cost[r, c], dtype=torch.float32))
    if not me:
        z = torch.zeros(0, dtype=torch.int32)
        return z, z.clone(), z.clone(), torch.zeros(0, dtype=torch.float32)
    return torch.cat(me), torch.cat(ms), torch.cat(md), torch.cat(msc)


def match_batch(edge_inds, edge_peak_inds, scores, n_edges: int, solver=None):
    """paf.py:622-702."""
    cols = ([], [], [], [])
    for b in range(len(edge_inds)):
        for dst, item in zip(cols, match_sample(edge_inds[b], edge_peak_inds[b], scores[b], n_edges, solver)):
            dst.append(item)
    return cols


# ---------------------------------------------------------------------- assembly
def toposort_edge_order(edge_list: Sequence[Tuple[int, int]]) -> Tuple[int, ...]:
    """Edge visiting order for assembly.  paf.py:890-912.

    Root = first node (in order of appearance in the edge list) without an incoming
    edge; BFS tree edges from that root, children in insertion order; each tree edge
    is reported by its FIRST index in `edge_list`.  Edges outside the root's BFS tree
    (other components, cross edges) are absent from the result.
    """
    edges = [(int(a), int(b)) for a, b in edge_list]
    order: List[int] = []
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
        raise ValueError("skeleton graph has a cycle: no root node")
    seen = {roots[0]}
    todo = deque([roots[0]])
    while todo:
        a = todo.popleft()
        for b in succ[a]:
            if b not in seen:
                seen.add(b)
                order.append(edges.index((a, b)))
                todo.append(b)
    return tuple(order)


def assign_instances(conn_by_edge, min_instance_peaks=0, n_nodes=None):
    """Greedy union of matched connections into instance ids.  paf.py:705-820.

    `conn_by_edge`: ordered list of ((src_node, dst_node), [(src_peak, dst_peak, score), ...]).
    Returns an insertion-ordered dict {(node, peak): instance id}.  Faithful quirks:
    "source free / destination taken" does nothing; when both are taken the
    destination is moved first and the two instances merge only if their node sets
    (computed after the move) are disjoint.
    """
    owner: Dict[Tuple[int, int], int] = {}
    for (sn, dn), conns in conn_by_edge:
        for sp, dp, _score in conns:
            a, b = (sn, int(sp)), (dn, int(dp))
            ia, ib = owner.get(a), owner.get(b)
            if ia is None and ib is None:
                new_id = max(owner.values(), default=-1) + 1
                owner[a] = new_id
                owner[b] = new_id
            elif ia is not None and ib is None:
                owner[b] = ia
            elif ia is not None and ib is not None:
                owner[b] = ia
                nodes_a = {k[0] for k, inst in owner.items() if inst == ia}
                nodes_b = {k[0] for k, inst in owner.items() if inst == ib}
                if not (nodes_a & nodes_b):
                    for k in owner:
                        if owner[k] == ib:
                            owner[k] = ia
    if min_instance_peaks > 0:
        if isinstance(min_instance_peaks, float):
            if n_nodes is None:
                n_nodes = len({n for (sn, dn), _ in conn_by_edge for n in (sn, dn)})
            min_instance_peaks = int(min_instance_peaks * n_nodes)
        sizes: Dict[int, int] = {}
        for inst in owner.values():
            sizes[inst] = sizes.get(inst, 0) + 1
        owner = {k: inst for k, inst in owner.items() if sizes[inst] >= min_instance_peaks}
    return owner


def build_instances(peaks_by_node, vals_by_node, conn_by_edge, owner):
    """Scatter peaks into NaN-filled (I,N,2)/(I,N) arrays; score = fp32 running sum.  paf.py:823-887."""
    ids = sorted(set(owner.values()))
    rank = {inst: i for i, inst in enumerate(ids)}
    scores = np.zeros(len(ids), dtype=np.float32)
    for (sn, dn), conns in conn_by_edge:
        for sp, dp, sc in conns:
            if (sn, int(sp)) in owner:
                scores[rank[owner[(sn, int(sp))]]] += np.float32(sc)
                # the reference's sanity check (paf.py:866-873): KeyError if the destination is in no kept instance
                assert owner[(sn, int(sp))] == owner[(dn, int(dp))]
    n_nodes = len(peaks_by_node)
    pts = np.full((len(ids), n_nodes, 2), np.nan, dtype=np.float32)
    pv = np.full((len(ids), n_nodes), np.nan, dtype=np.float32)
    for (node, pk), inst in owner.items():  # insertion order: later entries overwrite
        pts[rank[inst], node] = peaks_by_node[node][pk]
        pv[rank[inst], node] = vals_by_node[node][pk]
    return pts, pv, scores


def group_sample(peaks, vals, channel_inds, m_edge, m_src, m_dst, m_score, n_nodes,
                 sorted_edge_inds, edge_list, min_instance_peaks, min_line_scores=0.25):
    """paf.py:915-1038.  `edge_list` is [(src_node, dst_node), ...]."""
    to_np = lambda t: t.cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    peaks, vals, ch = to_np(peaks), to_np(vals), to_np(channel_inds)
    m_edge, m_src, m_dst, m_score = map(to_np, (m_edge, m_src, m_dst, m_score))
    keep = m_score >= min_line_scores
    m_edge, m_src, m_dst, m_score = m_edge[keep], m_src[keep], m_dst[keep], m_score[keep]
    peaks_by_node = [peaks[ch == k] for k in range(n_nodes)]
    vals_by_node = [vals[ch == k] for k in range(n_nodes)]
    conn_by_edge = []
    seen_types = {}
    for e in sorted_edge_inds:
        sel = m_edge == e
        et = (int(edge_list[e][0]), int(edge_list[e][1]))
        entry = (et, list(zip(m_src[sel].tolist(), m_dst[sel].tolist(), m_score[sel])))
        if et in seen_types:  # dict semantics: a repeated edge type replaces the earlier list
            conn_by_edge[seen_types[et]] = entry
        else:
            seen_types[et] = len(conn_by_edge)
            conn_by_edge.append(entry)
    owner = assign_instances(conn_by_edge, min_instance_peaks, n_nodes)
    return build_instances(peaks_by_node, vals_by_node, conn_by_edge, owner)


def group_batch(peaks, vals, channel_inds, m_edge, m_src, m_dst, m_score, n_nodes,
                sorted_edge_inds, edge_list, min_instance_peaks, min_line_scores=0.25):
    """paf.py:1041-1149."""
    out = ([], [], [])
    for b in range(len(peaks)):
        res = group_sample(peaks[b], vals[b], channel_inds[b], m_edge[b], m_src[b], m_dst[b], m_score[b],
                           n_nodes, sorted_edge_inds, edge_list, min_instance_peaks, min_line_scores)
        for dst, item in zip(out, res):
            dst.append(torch.from_numpy(item))
    return out


def predict(pafs_bhwc, peaks, vals, channel_inds, edge_list, n_nodes, stride, ratio=0.25, weight=1.0,
            n_points=10, min_instance_peaks=0, min_line_scores=0.25, sorted_edge_inds=None):
    """PAFScorer.predict, paf.py:1469-1532: score -> match -> group; returns the 6-tuple."""
    if sorted_edge_inds is None:
        sorted_edge_inds = toposort_edge_order(edge_list)
    e, p, s = score_lines_batch(pafs_bhwc, peaks, channel_inds, edge_list, n_points, stride, ratio, weight, n_nodes)
    me, ms, md, msc = match_batch(e, p, s, len(edge_list))
    inst, pv, isc = group_batch(peaks, vals, channel_inds, me, ms, md, msc, n_nodes, sorted_edge_inds,
                                edge_list, min_instance_peaks, min_line_scores)
    return inst, pv, isc, e, p, s
