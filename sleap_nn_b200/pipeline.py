"""Fused bottom-up post-processing: confidence maps + PAFs -> grouped instances, one C call per batch.

`BottomUpPostproc` owns fixed-capacity device tables for one batch shape and enqueues the whole
kernel chain (peaks + refinement -> candidates + line scores -> per-edge assignment -> assembly)
on the current CUDA stream through `snb_bottomup_postproc`, with no host synchronisation and no
allocation, so several instances can pipeline batches over separate streams.  It is the
device-resident replacement for the reference call sequence in
`BottomUpLayer._score_pafs_on_gpu` + `group_scored_batch`
(sleap_nn/inference/layers/bottomup.py:95-236, sleap_nn/inference/streaming.py:147-255).
Knob names and defaults follow `PAFScorer` (sleap_nn/inference/ops/paf.py:1208-1218).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import torch

from sleap_nn_b200 import _native as N
from sleap_nn_b200.inference.ops.paf import EdgeType, _t_table, toposort_edges


@dataclass
class BottomUpResult:
    """Padded device tensors of one batch (valid prefix given by the count tensors)."""

    n_instances: torch.Tensor     # (B,) i32
    instances: torch.Tensor       # (B, inst_cap, N, 2) f32, NaN = missing node
    peak_scores: torch.Tensor     # (B, inst_cap, N) f32
    instance_scores: torch.Tensor  # (B, inst_cap) f32
    n_peaks: torch.Tensor         # (B,) i32 (true count; may exceed peak_cap -> status bit)
    peaks: torch.Tensor           # (B, peak_cap, 2) f32 image-scale (x, y), (y, x, channel)-ordered
    peak_vals: torch.Tensor       # (B, peak_cap)
    peak_channels: torch.Tensor   # (B, peak_cap) i32
    status: torch.Tensor          # (1,) i32, SNB_STATUS_* bits
    done: Optional[torch.cuda.Event] = None  # set in two-stream mode: the tail ran on another stream
    # group_scored_batch-shaped outputs (inference/streaming.py:196-243); filled in the same call when the
    # pipeline was built with max_instances, else by .outputs()
    pred_keypoints: Optional[torch.Tensor] = None        # (B, I, N, 2) NaN-padded, scale-undone
    pred_peak_values: Optional[torch.Tensor] = None      # (B, I, N)
    pred_instance_scores: Optional[torch.Tensor] = None  # (B, I)
    skip_flag: Optional[torch.Tensor] = None             # (1,) i32: the max_peaks_per_node guard tripped
    _pipe: Optional["BottomUpPostproc"] = None
    _scales: Tuple[float, Optional[torch.Tensor]] = (1.0, None)

    def wait(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Make `stream` (default: current) wait for this batch's tail kernel."""
        if self.done is not None:
            (stream or torch.cuda.current_stream(self.status.device)).wait_event(self.done)

    def outputs(self, max_instances: Optional[int] = None):
        """(pred_keypoints (B,I,N,2), pred_peak_values (B,I,N), instance_scores (B,I)) device tensors with the
        semantics of `group_scored_batch` (inference/streaming.py:147-255): NaN padding, top-N by score when
        truncating, input / effective scale undone, all-NaN when the max_peaks_per_node guard tripped.

        With a fixed `max_instances` (given here or at construction) nothing synchronises with the host; with
        None the per-batch maximum instance count is read back first (the reference's `_infer_max_instances`).
        """
        if max_instances is None and self.pred_keypoints is not None:
            return self.pred_keypoints, self.pred_peak_values, self.pred_instance_scores
        self.wait()
        pipe, dev = self._pipe, self.status.device
        B, cap, Nn = self.instances.shape[0], self.instances.shape[1], self.instances.shape[2]
        if max_instances is None:
            skipped = self.skip_flag is not None and bool(self.skip_flag.item())
            max_instances = 0 if skipped else int(self.n_instances.clamp(max=cap).max().item()) if B else 0
        I = max(int(max_instances), 1)
        with torch.cuda.device(dev):
            k = torch.empty((B, I, Nn, 2), dtype=torch.float32, device=dev)
            v = torch.empty((B, I, Nn), dtype=torch.float32, device=dev)
            s = torch.empty((B, I), dtype=torch.float32, device=dev)
            scale, eff = self._scales
            N.check(N.lib.snb_bottomup_outputs(N.ptr(self.n_instances), N.ptr(self.instances), N.ptr(self.peak_scores),
                                               N.ptr(self.instance_scores), B, cap, Nn, I, float(scale), N.ptr(eff),
                                               N.ptr(self.skip_flag), N.ptr(k), N.ptr(v), N.ptr(s), N.stream_ptr(dev)),
                    "snb_bottomup_outputs")
        return k, v, s

    def to_lists(self):
        """One host sync: per-sample CPU tensors shaped like `PAFScorer.predict`'s first three outputs."""
        self.wait()
        n = self.n_instances.cpu().tolist()
        status = int(self.status.item())
        if status:
            self.status.zero_()  # sticky bits are reported once
        if status & N.STATUS_ASM_MISMATCH:  # the reference's own sanity assert (ops/paf.py:866-873)
            raise AssertionError("both peaks of a connection should have been assigned to the same instance")
        if status & N.STATUS_ASM_MISSING:
            raise KeyError("destination peak of a scored connection is not assigned to an instance")
        if status & N.STATUS_LSAP_INFEASIBLE:
            raise ValueError("cost matrix is infeasible")
        if status:
            raise RuntimeError(f"bottom-up post-processing overflowed a fixed-capacity table (status 0x{status:x}); "
                               "raise peak_cap / cand_cap / match_cap / inst_cap")
        xy, pv, sc = self.instances.cpu(), self.peak_scores.cpu(), self.instance_scores.cpu()
        B = len(n)
        return ([xy[b, : n[b]] for b in range(B)], [pv[b, : n[b]] for b in range(B)], [sc[b, : n[b]] for b in range(B)])


class BottomUpPostproc:
    """Device-resident bottom-up post-processing for batches of a fixed shape.

    Args:
        n_nodes, edge_inds: skeleton (edge_inds = [(src_node, dst_node), ...]).
        batch, cms_hw: confidence maps are (batch, n_nodes, H, W); PAFs are (batch, 2E, Hp, Wp) or
            the permuted (batch, Hp, Wp, 2E) view - either is read in place through its strides.
        cms_stride / pafs_stride: output strides of the two heads.
        peak_threshold, refinement, integral_patch_size: as `find_local_peaks`.
        max_edge_length_ratio, dist_penalty_weight, n_points, min_instance_peaks, min_line_scores:
            as `PAFScorer`.
        peak_cap, cand_cap, match_cap, inst_cap: per-frame table capacities; overflow sets a status
            bit (reported by `BottomUpResult.to_lists`), it never corrupts memory.
        tail_stream: optional second (high-priority) CUDA stream: the streaming detect kernel runs on
            the caller's current stream and the per-frame tail on `tail_stream`, so the tail of one
            batch overlaps the detect pass of the next.
        max_instances: when given, every call also fills `(B, max_instances, N, ...)` NaN-padded outputs with
            `group_scored_batch` semantics (top-N by score, scales undone) in the same launch chain.
        max_peaks_per_node: the batch-wide guard of `BottomUpLayer` (layers/bottomup.py:128-148): if any node of
            any frame has more peaks, the outputs are all-NaN.
        fused_tail: run everything after the detect kernel as one CTA per frame with shared-memory
            tables (default; falls back automatically when the capacities do not fit).
        keep_tables: also write the intermediate tables (candidates, matches) to global memory.
        tail_cluster: batches of <= 16 frames run the tail as a 4-CTA thread-block cluster per frame (refinement,
            line scores and assignments split over the CTAs through distributed shared memory) - the latency of a lone
            call instead of the throughput of a pipelined one; False keeps one CTA per frame (A/B, tests).
    """

    def __init__(self, n_nodes: int, edge_inds: Sequence[Tuple[int, int]], batch: int, cms_hw: Tuple[int, int],
                 cms_stride: int = 2, pafs_stride: int = 2, peak_threshold: float = 0.2,
                 refinement: Optional[str] = "integral", integral_patch_size: int = 5,
                 max_edge_length_ratio: float = 0.25, dist_penalty_weight: float = 1.0, n_points: int = 10,
                 min_instance_peaks: Union[int, float] = 0, min_line_scores: float = 0.25, peak_cap: int = 256,
                 cand_cap: int = 4096, match_cap: int = 512, inst_cap: int = 64, lsap_max_dim: int = 32,
                 device: Optional[torch.device] = None, tail_stream: Optional[torch.cuda.Stream] = None,
                 fused_tail: bool = True, keep_tables: bool = True, max_instances: Optional[int] = None,
                 max_peaks_per_node: Optional[int] = None, tail_cluster: bool = True,
                 sorted_edge_inds: Optional[Sequence[int]] = None):
        self.device = torch.device(device) if device is not None else N.compute_device()
        self.n_nodes, self.batch = int(n_nodes), int(batch)
        self.edge_inds = [(int(a), int(b)) for a, b in edge_inds]
        self.n_edges = len(self.edge_inds)
        self.cms_hw = (int(cms_hw[0]), int(cms_hw[1]))
        self.cms_stride, self.pafs_stride = cms_stride, pafs_stride
        self.peak_threshold = float(peak_threshold)
        self.refine_size = int(integral_patch_size) if refinement == "integral" else 0
        self.max_edge_length_ratio = float(max_edge_length_ratio)
        self.dist_penalty_weight = float(dist_penalty_weight)
        self.n_points = int(n_points)
        if isinstance(min_instance_peaks, float):
            min_instance_peaks = int(min_instance_peaks * n_nodes) if min_instance_peaks > 0 else 0  # paf.py:791-802
        self.min_instance_peaks = int(min_instance_peaks)
        self.min_line_scores = float(min_line_scores)
        self.caps = dict(peak_cap=int(peak_cap), cand_cap=int(cand_cap), match_cap=int(match_cap),
                         inst_cap=int(inst_cap), lsap_max_dim=int(lsap_max_dim))
        # order in which the assembly visits the edges: the reference's PAFScorer.sorted_edge_inds (ops/paf.py:1230-1236,
        # toposort_edges); an explicit order is for tests of the assembly's non-forest path
        if sorted_edge_inds is not None:
            self.sorted_edge_inds = tuple(int(e) for e in sorted_edge_inds)
            if any(e < 0 or e >= self.n_edges for e in self.sorted_edge_inds):
                raise ValueError("sorted_edge_inds must index edge_inds")
        else:
            self.sorted_edge_inds = toposort_edges([EdgeType(a, b) for a, b in self.edge_inds]) if self.n_edges else ()
        dev, B, Nn, E = self.device, self.batch, self.n_nodes, self.n_edges
        i32 = lambda *shape: torch.empty(shape, dtype=torch.int32, device=dev)
        f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            self._edges = torch.tensor(self.edge_inds, dtype=torch.int32, device=dev).reshape(-1, 2)
            self._sorted = torch.tensor(list(self.sorted_edge_inds), dtype=torch.int32, device=dev)
            self._t = _t_table(self.n_points, dev)
            pc, cc, mc, ic = peak_cap, cand_cap, match_cap, inst_cap
            self.buf = dict(
                frame_count=i32(B), keys=i32(B * pc), peak_xy=f32(B, pc, 2), peak_val=f32(B, pc), peak_chan=i32(B, pc),
                node_start=i32(B, Nn + 1), node_peaks=i32(B * pc), edge_off=i32(B, E + 1), match_off=i32(B, E + 1),
                cand_edge=i32(B * cc), cand_epi=torch.empty((B * cc, 2), dtype=torch.int64, device=dev),
                cand_score=f32(B * cc), m_edge=i32(B * mc), m_src=i32(B * mc), m_dst=i32(B * mc), m_score=f32(B * mc),
                m_count=i32(B), asm_ws=i32(B * 4 * pc), inst_xy=f32(B, ic, Nn, 2), inst_val=f32(B, ic, Nn),
                inst_score=f32(B, ic), n_inst=i32(B), status=torch.zeros((1,), dtype=torch.int32, device=dev),
            )
            ws_bytes = N.lib.snb_lsap_workspace_bytes(int(lsap_max_dim)) * B * max(E, 1) if lsap_max_dim > 32 else 0
            self.buf["lsap_ws"] = torch.empty((max(ws_bytes // 8, 1),), dtype=torch.int64, device=dev)
        self._args = N.BottomUpArgs()
        a = self._args
        a.B, a.C, a.H, a.W = B, Nn, self.cms_hw[0], self.cms_hw[1]
        a.edges, a.n_edges = N.ptr(self._edges), E
        a.sorted_edges, a.n_sorted = N.ptr(self._sorted), len(self.sorted_edge_inds)
        a.t_table, a.n_points = N.ptr(self._t), self.n_points
        a.peak_threshold, a.refine_size = self.peak_threshold, self.refine_size
        a.cms_stride, a.pafs_stride = float(cms_stride), float(pafs_stride)
        a.dist_penalty_weight = self.dist_penalty_weight
        a.min_instance_peaks, a.min_line_scores = self.min_instance_peaks, self.min_line_scores
        for k, v in self.caps.items():
            setattr(a, k, v)
        for k, t in self.buf.items():
            setattr(a, k, N.ptr(t))
        if lsap_max_dim <= 32:
            a.lsap_ws = None
        a.ev_detect_begin = a.ev_detect_end = None
        a.flags = 0 if fused_tail else N.FLAG_UNFUSED_TAIL
        if not tail_cluster:
            a.flags |= N.FLAG_NO_TAIL_CLUSTER
        a.n_peaks = None
        self.fused = int(N.lib.snb_bottomup_launches_per_call(C.byref(a))) == 2
        if self.fused:
            # self-resetting peak counters: the detect kernel counts into a zero-initialised scratch, the tail copies
            # each frame's count to buf["frame_count"] (the result the caller reads) and zeroes the scratch again -
            # the chain is exactly two kernel launches, no memset node
            with torch.cuda.device(dev):
                self._peak_counter = torch.zeros((B,), dtype=torch.int32, device=dev)
            a.frame_count, a.n_peaks = N.ptr(self._peak_counter), N.ptr(self.buf["frame_count"])
            a.flags |= N.FLAG_SELF_RESET_COUNTERS
        if self.fused and not keep_tables:  # optional outputs of the fused tail
            for k in ("node_start", "node_peaks", "edge_off", "match_off", "cand_edge", "cand_epi", "cand_score",
                      "m_edge", "m_src", "m_dst", "m_score", "m_count"):
                setattr(a, k, None)
        self.max_instances = None if max_instances is None else max(int(max_instances), 1)
        self.max_peaks_per_node = max_peaks_per_node
        a.max_peaks_per_node, a.skip_flag = 0, None
        a.max_instances, a.input_scale, a.eff_scale = 1, 1.0, None
        a.out_kpts = a.out_vals = a.out_scores = None
        self._out = None
        with torch.cuda.device(dev):
            if max_peaks_per_node is not None:
                self._skip_flag = torch.zeros((1,), dtype=torch.int32, device=dev)
                a.max_peaks_per_node, a.skip_flag = int(max_peaks_per_node), N.ptr(self._skip_flag)
            else:
                self._skip_flag = None
            if self.max_instances is not None:
                I = self.max_instances
                self._out = (f32(B, I, Nn, 2), f32(B, I, Nn), f32(B, I))
                a.max_instances = I
                a.out_kpts, a.out_vals, a.out_scores = (N.ptr(t) for t in self._out)
        self._fast = {}
        self.tail_stream = tail_stream
        self._ev_handoff = self._ev_done = None
        a.tail_stream = a.ev_handoff = a.ev_tail_done = None
        if tail_stream is not None:
            with torch.cuda.device(dev):
                self._ev_handoff, self._ev_done = torch.cuda.Event(), torch.cuda.Event()
                for e in (self._ev_handoff, self._ev_done):  # torch creates the cudaEvent lazily on first record
                    e.record(torch.cuda.current_stream(dev))
                torch.cuda.synchronize(dev)
            a.tail_stream = tail_stream.cuda_stream
            a.ev_handoff, a.ev_tail_done = self._ev_handoff.cuda_event, self._ev_done.cuda_event

    # ------------------------------------------------------------------ device-resident path
    def __call__(self, cms: torch.Tensor, pafs: torch.Tensor, detect_events=None, input_scale: float = 1.0,
                 eff_scale: Optional[torch.Tensor] = None, stream: Optional[torch.cuda.Stream] = None) -> BottomUpResult:
        """Enqueue the chain on the current stream of `self.device`; returns padded device tensors.

        cms (B, N, H, W) CUDA; pafs (B, 2E, Hp, Wp) or (B, Hp, Wp, 2E) CUDA, any strides; fp32, or the fp16 / bf16
        heads of an autocast backbone, read in place: results are bit-identical to running on `cms.float()`,
        `pafs.float()` (what torch_backend.py:140-146 does before the reference's ops), without that copy.
        `detect_events` = (torch.cuda.Event, torch.cuda.Event) recorded around the streaming
        detect kernel (for the benchmark's roofline figure).  `input_scale` / `eff_scale` (B,) are
        `PreprocInfo`'s scale factors, undone in the `.outputs()` tensors (streaming.py:190-196).
        `stream`: enqueue on this stream instead of the current one (saves the ~10 us `with torch.cuda.stream(...)`
        costs a tight multi-stream launch loop per step).
        """
        # The C side launches on the CURRENT device: when the caller's current device is another GPU, switch for the
        # duration of the call (otherwise the chain would run there, reaching the tables over peer access, unordered
        # with this device's streams).
        if torch.cuda.current_device() != self.device.index:
            with torch.cuda.device(self.device):
                return self.__call__(cms, pafs, detect_events, input_scale, eff_scale, stream)
        # fast path: tensors this pipeline has already been launched on (a loop fed from a few fixed buffers, the
        # steady state of a streaming pipeline) - a filled copy of the argument block is kept per input, so only the
        # launch itself is left (~5 us of host time instead of ~25 us)
        # (the block is a pure function of the key - pointers, strides, shapes, scale, knobs - so nothing is kept alive)
        key = (cms.data_ptr(), pafs.data_ptr(), cms.stride(), pafs.stride(), tuple(cms.shape), tuple(pafs.shape),
               cms.dtype, pafs.dtype, input_scale, self._knobs())
        hit = self._fast.get(key) if (detect_events is None and eff_scale is None) else None
        if hit is not None:
            N.check(N.lib.snb_bottomup_postproc(C.byref(hit[0]), stream.cuda_stream if stream is not None
                                                else N.stream_ptr(self.device)), "snb_bottomup_postproc")
            return hit[1]
        if stream is not None:  # slow path (first call on these tensors): the ordinary current-stream route
            with torch.cuda.stream(stream):
                return self.__call__(cms, pafs, detect_events, input_scale, eff_scale, None)
        if not (cms.is_cuda and pafs.is_cuda) or cms.dtype not in N._DTYPES or pafs.dtype not in N._DTYPES:
            raise TypeError("BottomUpPostproc expects fp32 / fp16 / bf16 CUDA tensors; use .run_host() for host buffers")
        if cms.device != self.device or pafs.device != self.device:
            raise ValueError(f"inputs live on {cms.device} / {pafs.device}, the pipeline's tables on {self.device}")
        if tuple(cms.shape) != (self.batch, self.n_nodes) + self.cms_hw:
            raise ValueError(f"cms shape {tuple(cms.shape)} does not match the pipeline's batch shape")
        if pafs.dim() != 4 or pafs.shape[0] != self.batch:
            raise ValueError("pafs must be (B, 2E, H, W) or its (B, H, W, 2E) view")
        if pafs.shape[1] == 2 * self.n_edges and pafs.shape[-1] != 2 * self.n_edges:
            pafs = pafs.permute(0, 2, 3, 1)  # channels-first tensor -> the channels-last VIEW (no copy)
        elif pafs.shape[-1] != 2 * self.n_edges:
            raise ValueError("pafs channel count must be 2 * n_edges")
        res = self._launch(N.ptr(cms), cms.stride(), N.ptr(pafs), tuple(pafs.shape), pafs.stride(), detect_events,
                           input_scale, eff_scale, cms.dtype, pafs.dtype)
        if detect_events is None and eff_scale is None:
            if len(self._fast) >= 16:
                self._fast.clear()
            block = N.BottomUpArgs()
            C.memmove(C.byref(block), C.byref(self._args), C.sizeof(N.BottomUpArgs))
            self._fast[key] = (block, res)
        return res

    def _knobs(self):
        """The scalar knobs a caller may change on a live object (the reference's PAFScorer is mutable the same way,
        ops/paf.py:1208-1218); they are re-read on every launch and are part of the fast-path key."""
        return (self.peak_threshold, self.refine_size, self.max_edge_length_ratio, self.dist_penalty_weight,
                self.min_instance_peaks, self.min_line_scores, self.cms_stride, self.pafs_stride, self.max_peaks_per_node)

    def _launch(self, cms_ptr: int, cms_strides, pafs_ptr: int, pafs_shape, pafs_strides, detect_events=None,
                input_scale: float = 1.0, eff_scale: Optional[torch.Tensor] = None,
                cms_dtype: torch.dtype = torch.float32, pafs_dtype: torch.dtype = torch.float32) -> BottomUpResult:
        """Fill the argument block from raw device-visible pointers and enqueue the chain (current device must be
        `self.device`: every caller holds the guard)."""
        a = self._args
        a.peak_threshold, a.refine_size = float(self.peak_threshold), int(self.refine_size)
        a.cms_stride, a.pafs_stride = float(self.cms_stride), float(self.pafs_stride)
        a.dist_penalty_weight = float(self.dist_penalty_weight)
        a.min_instance_peaks, a.min_line_scores = int(self.min_instance_peaks), float(self.min_line_scores)
        if self._skip_flag is not None:
            a.max_peaks_per_node = int(self.max_peaks_per_node or 0)
        a.cms_dtype, a.pafs_dtype = N.dtype_code(cms_dtype), N.dtype_code(pafs_dtype)
        if eff_scale is not None:
            eff_scale = eff_scale.detach().to(device=self.device, dtype=torch.float32).reshape(-1).contiguous()
            if eff_scale.numel() == 1 and self.batch != 1:
                eff_scale = eff_scale.expand(self.batch).contiguous()
            if eff_scale.numel() != self.batch:
                raise ValueError("eff_scale must hold one factor per sample")
        a.input_scale, a.eff_scale = float(input_scale), N.ptr(eff_scale)
        a.cms = cms_ptr
        a.cms_sb, a.cms_sc, a.cms_sh, a.cms_sw = cms_strides
        a.pafs = pafs_ptr
        a.paf_H, a.paf_W = int(pafs_shape[1]), int(pafs_shape[2])
        a.paf_sb, a.paf_sy, a.paf_sx, a.paf_sc = pafs_strides
        # max(H, W, 2E) really includes the channel dim (paf.py:457-461)
        a.max_edge_length = self.max_edge_length_ratio * max(pafs_shape[-1], pafs_shape[-2], pafs_shape[-3]) * self.pafs_stride
        if detect_events is not None:
            a.ev_detect_begin, a.ev_detect_end = detect_events[0].cuda_event, detect_events[1].cuda_event
        else:
            a.ev_detect_begin = a.ev_detect_end = None
        N.check(N.lib.snb_bottomup_postproc(C.byref(a), N.stream_ptr(self.device)), "snb_bottomup_postproc")
        b = self.buf
        k, v, sc = self._out if self._out is not None else (None, None, None)
        return BottomUpResult(b["n_inst"], b["inst_xy"], b["inst_val"], b["inst_score"], b["frame_count"],
                              b["peak_xy"], b["peak_val"], b["peak_chan"], b["status"], self._ev_done,
                              k, v, sc, self._skip_flag, self, (float(input_scale), eff_scale))

    @property
    def launches_per_call(self) -> int:
        return int(N.lib.snb_bottomup_launches_per_call(C.byref(self._args)))

    # ------------------------------------------------------------------ host-buffer path
    def _launch_from_host(self, cms_host: torch.Tensor, pafs_host: torch.Tensor, zero_copy_pafs: bool = True,
                          zero_copy_cms: bool = False) -> BottomUpResult:
        """Enqueue copy + chain for one batch that lives in HOST memory, on the current stream of `self.device` (the
        caller holds the device guard).  Sets `last_h2d_bytes`, `last_zero_copy`, `last_zero_copy_cms`.

        The confidence maps are copied to a device staging buffer (every element has to be looked at) - or, with
        `zero_copy_cms` and a pinned buffer, streamed by the detect kernel itself straight out of host memory through
        the buffer's device alias (the kernel's 128-bit loads are the DMA: no staging write + re-read in HBM, same
        bytes over PCIe).  The PAF tensor is only SAMPLED (n_points x 2 taps per candidate), so a pinned one is read in
        place and only the sampled sectors cross the link; an unpinned buffer is staged like the confidence maps.
        """
        if cms_host.dtype not in N._DTYPES or pafs_host.dtype not in N._DTYPES:
            raise TypeError("host tensors must be float32, float16 or bfloat16")
        if pafs_host.dim() != 4 or pafs_host.shape[0] != self.batch:
            raise ValueError("pafs must be (B, 2E, H, W) or its (B, H, W, 2E) view")
        if tuple(cms_host.shape) != (self.batch, self.n_nodes) + self.cms_hw:
            raise ValueError(f"cms shape {tuple(cms_host.shape)} does not match the pipeline's batch shape")
        channels_first = pafs_host.shape[1] == 2 * self.n_edges and pafs_host.shape[-1] != 2 * self.n_edges
        if not channels_first and pafs_host.shape[-1] != 2 * self.n_edges:
            raise ValueError("pafs channel count must be 2 * n_edges")
        view = (lambda t: t.permute(0, 2, 3, 1)) if channels_first else (lambda t: t)

        def alias_of(t):
            p = C.c_void_p()
            ok = t.is_pinned() and N.lib.snb_host_device_pointer(t.data_ptr(), C.byref(p)) == N.OK and p.value
            return p.value if ok else None

        cms_alias = alias_of(cms_host) if zero_copy_cms else None
        if cms_alias is not None:
            cms_ptr, cms_strides = cms_alias, cms_host.stride()
        else:
            st = getattr(self, "_stage_cms", None)
            if st is None or st.shape != cms_host.shape or st.dtype != cms_host.dtype:
                st = self._stage_cms = torch.empty(cms_host.shape, dtype=cms_host.dtype, device=self.device)
            st.copy_(cms_host, non_blocking=True)
            cms_ptr, cms_strides = N.ptr(st), st.stride()
        paf_alias = alias_of(pafs_host) if zero_copy_pafs else None
        if paf_alias is not None:
            pv = view(pafs_host)
            paf_ptr = paf_alias
        else:  # stage in the host tensor's own memory order (a dense copy), then take the channels-last view
            sp = getattr(self, "_stage_pafs", None)
            if sp is None or sp.shape != pafs_host.shape or sp.dtype != pafs_host.dtype or sp.stride() != pafs_host.stride():
                sp = self._stage_pafs = torch.empty_strided(tuple(pafs_host.shape), pafs_host.stride(),
                                                            dtype=pafs_host.dtype, device=self.device)
            sp.copy_(pafs_host, non_blocking=True)
            pv = view(sp)
            paf_ptr = N.ptr(pv)
        res = self._launch(cms_ptr, cms_strides, paf_ptr, tuple(pv.shape), pv.stride(), None, 1.0, None,
                           cms_host.dtype, pafs_host.dtype)
        self.last_zero_copy, self.last_zero_copy_cms = paf_alias is not None, cms_alias is not None
        self.last_h2d_bytes = (cms_host.numel() * cms_host.element_size()
                               + (0 if paf_alias is not None else pafs_host.numel() * pafs_host.element_size()))
        return res

    def run_host(self, cms_host: torch.Tensor, pafs_host: torch.Tensor, zero_copy_pafs: bool = True,
                 zero_copy_cms: bool = False):
        """Host buffers in, per-sample CPU tensors out: H2D + chain + D2H (synchronous; `BottomUpHostStream` pipelines).

        `self.last_h2d_bytes` holds the bytes this call moved host -> device (the maps, plus the PAF tensor when it
        had to be staged).  Returns (instances, peak_scores, instance_scores) lists like `PAFScorer.predict`.
        """
        with torch.cuda.device(self.device):
            return self._launch_from_host(cms_host, pafs_host, zero_copy_pafs, zero_copy_cms).to_lists()


class BottomUpHostStream:
    """Software pipeline for batches that ARRIVE IN HOST MEMORY (a frame grabber, a decoder, another process).

    `BottomUpPostproc.run_host` is synchronous: copy, compute, read back, and only then the next batch's copy
    starts, so the PCIe link idles while the host unpacks results.  Here `depth` slots, each with its own CUDA
    stream, staging buffer, table set and pinned result buffers, keep the link busy: while slot k's results are
    read back and unpacked, slot k+1's confidence maps are already crossing PCIe.  The per-step work is unchanged
    (every step copies its inputs host -> device and its results device -> host); only the idle gaps go.

        stream = BottomUpHostStream(lambda: BottomUpPostproc(...), depth=2)
        for cms_host, pafs_host in batches:            # pinned fp32 tensors
            done = stream.submit(cms_host, pafs_host)  # results of the batch submitted `depth` calls ago, or None
        rest = stream.drain()                          # the remaining results, in submission order

    Results are `(instances, peak_scores, instance_scores)` per-sample lists like `PAFScorer.predict`.
    """

    def __init__(self, make_pipe, depth: int = 2, zero_copy_pafs: bool = True, zero_copy_cms: bool = False):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.pipes = [make_pipe() for _ in range(depth)]
        self.device = self.pipes[0].device
        self.zero_copy_pafs, self.zero_copy_cms = zero_copy_pafs, zero_copy_cms
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream(device=self.device) for _ in range(depth)]
            self.done = [torch.cuda.Event() for _ in range(depth)]
        self._host = [None] * depth   # pinned result buffers per slot
        self._busy = [False] * depth
        self._next = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _result_buffers(self, k: int):
        if self._host[k] is None:
            b = self.pipes[k].buf
            pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            self._host[k] = dict(n_inst=pin(b["n_inst"]), status=pin(b["status"]), inst_xy=pin(b["inst_xy"]),
                                 inst_val=pin(b["inst_val"]), inst_score=pin(b["inst_score"]))
        return self._host[k]

    def _collect(self, k: int):
        self.done[k].synchronize()
        self._busy[k] = False
        h = self._host[k]
        status = int(h["status"][0])
        if status & N.STATUS_ASM_MISMATCH:  # the reference's own sanity assert (ops/paf.py:866-873)
            raise AssertionError("both peaks of a connection should have been assigned to the same instance")
        if status & N.STATUS_ASM_MISSING:
            raise KeyError("destination peak of a scored connection is not assigned to an instance")
        if status & N.STATUS_LSAP_INFEASIBLE:
            raise ValueError("cost matrix is infeasible")
        if status:
            raise RuntimeError(f"bottom-up post-processing overflowed a fixed-capacity table (status 0x{status:x}); "
                               "raise peak_cap / cand_cap / match_cap / inst_cap")
        n = h["n_inst"].tolist()
        xy, pv, sc = h["inst_xy"], h["inst_val"], h["inst_score"]
        B = len(n)
        # clone: the pinned buffers are reused by the slot's next batch
        return ([xy[b, : n[b]].clone() for b in range(B)], [pv[b, : n[b]].clone() for b in range(B)],
                [sc[b, : n[b]].clone() for b in range(B)])

    def submit(self, cms_host: torch.Tensor, pafs_host: torch.Tensor):
        """Enqueue one batch; returns the results of the batch that previously occupied the slot (or None)."""
        k = self._next
        self._next = (k + 1) % len(self.pipes)
        out = self._collect(k) if self._busy[k] else None
        pipe, st = self.pipes[k], self.streams[k]
        h = self._result_buffers(k)
        with torch.cuda.device(self.device), torch.cuda.stream(st):
            res = pipe._launch_from_host(cms_host, pafs_host, self.zero_copy_pafs, self.zero_copy_cms)
            res.wait(st)
            b = pipe.buf
            for name in ("n_inst", "status", "inst_xy", "inst_val", "inst_score"):
                h[name].copy_(b[name], non_blocking=True)
            b["status"].zero_()  # sticky bits are reported once (after the copy above, in stream order)
            self.done[k].record(st)
        self._busy[k] = True
        self.h2d_bytes = pipe.last_h2d_bytes
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in h.values())
        self.last_zero_copy, self.last_zero_copy_cms = pipe.last_zero_copy, pipe.last_zero_copy_cms
        return out

    def drain(self):
        """Results of every batch still in flight, in submission order."""
        outs = []
        for i in range(len(self.pipes)):
            k = (self._next + i) % len(self.pipes)
            if self._busy[k]:
                outs.append(self._collect(k))
        return outs


class PipelineRing:
    """Round-robin of `BottomUpPostproc` instances over their own CUDA streams: the throughput configuration.

    With k instances the tail of one batch should hide under the detect pass of the next.  Left to itself it often does
    not: the next batch's detect kernel (20 480 CTAs, ready long before) floods every SM the moment the previous detect
    kernel drains, its CTAs keep refilling the register files, and the 64 tail CTAs (116 registers x 256 threads each)
    find no room until that kernel's LAST wave - so the chain that waits for this tail starts late and the detect passes
    end up strictly one after another with a launch bubble between them (52 us per cfg3 step; whether a process landed in
    that state or in the good one depended on how its streams happened to map onto hardware queues).  Giving the tails a
    HIGH-PRIORITY stream makes the block scheduler place them first: 44.4 us per step on every rank of every run,
    with 2, 3 or 4 instances (`profiles/r2_tail_priority_ab.txt`).

        ring = PipelineRing.build(3, n_nodes=N, edge_inds=edges, batch=B, cms_hw=(H, W), cms_stride=2, pafs_stride=2)
        for cms, pafs in batches: res = ring.submit(cms, pafs); ...; res.wait()   # res.wait(): the tail ran on another stream

    `PipelineRing(pipes, streams)` wraps pre-built instances instead; if those have no tail stream the ring staggers the
    chains' first detect kernels after every `reset()` (events around the detect launches), which is what
    `capture_rotation` does inside a CUDA graph, where cross-stream priorities are not available.
    """

    def __init__(self, pipes: Sequence["BottomUpPostproc"], streams: Optional[Sequence[torch.cuda.Stream]] = None,
                 stagger: Optional[bool] = None):
        if not pipes:
            raise ValueError("need at least one pipeline")
        self.pipes = list(pipes)
        self.device = self.pipes[0].device
        # stagger: after reset(), chain k's first detect kernel waits for chain k-1's (default: only when the pipes have
        # no tail stream; with one, it only shortens the fill of an idle ring, and measured slower over a short run: 52.8 vs 50.2 us per step over 20 steps, bench.py --fill-stagger)
        self._stagger = all(p.tail_stream is None for p in self.pipes) if stagger is None else bool(stagger)
        with torch.cuda.device(self.device):
            self.streams = list(streams) if streams is not None else [torch.cuda.Stream(device=self.device) for _ in self.pipes]
            self._ev = [(torch.cuda.Event(), torch.cuda.Event()) for _ in self.pipes]
            for pair in self._ev:  # torch creates the cudaEvent lazily on first record; the C side re-records them
                for e in pair:
                    e.record(torch.cuda.current_stream(self.device))
        self._i = 0
        self._fresh = len(self.pipes) if self._stagger else 0

    @classmethod
    def build(cls, n: int = 3, device: Optional[torch.device] = None, **pipe_kwargs) -> "PipelineRing":
        """n instances sharing ONE high-priority tail stream, each with its own detect stream."""
        dev = torch.device(device) if device is not None else N.compute_device()
        with torch.cuda.device(dev):
            tail = torch.cuda.Stream(device=dev, priority=-1)
        return cls([BottomUpPostproc(device=dev, tail_stream=tail, **pipe_kwargs) for _ in range(int(n))])

    def reset(self) -> None:
        """The ring is idle: (stagger mode only) offset the chains again on the next submissions."""
        self._fresh = len(self.pipes) if self._stagger else 0
        self._i = 0

    def submit(self, cms: torch.Tensor, pafs: torch.Tensor) -> "BottomUpResult":
        k = self._i
        self._i = (k + 1) % len(self.pipes)
        if self._fresh > 0:
            self._fresh -= 1
            with torch.cuda.device(self.device), torch.cuda.stream(self.streams[k]):
                if k > 0 and self.streams[k] is not self.streams[k - 1]:
                    self.streams[k].wait_event(self._ev[k - 1][1])
                return self.pipes[k](cms, pafs, detect_events=self._ev[k])
        return self.pipes[k](cms, pafs, stream=self.streams[k])


def capture_rotation(pipes: Sequence[BottomUpPostproc], inputs: Sequence[Tuple[torch.Tensor, torch.Tensor]],
                     streams: Optional[Sequence[torch.cuda.Stream]] = None, repeats: int = 1,
                     stagger_chains: bool = True) -> Tuple[torch.cuda.CUDAGraph, int]:
    """Capture one full rotation of a multi-stream pipeline into a CUDA graph.

    Step i runs `pipes[i % len(pipes)]` on `inputs[i % len(inputs)]` on `streams[i % len(pipes)]`, for
    i in range(repeats * lcm(len(pipes), len(inputs))).  Consecutive replays on one stream serialise, so the
    pipeline drains once per replay (the last tail is not overlapped): use enough `repeats` to amortise that (one
    un-overlapped tail costs ~30 us, i.e. 0.3 us per step over 100 steps).  The chain is two launches per batch and ~50 us of GPU time, so an eager
    loop needs the host to issue a ctypes call every ~45 us per GPU; with 8 ranks on 16 host cores that loop, not
    HBM, can set the pace.  Replaying the graph costs one `cudaGraphLaunch` per replay and keeps the fork/join
    structure inside the graph.  Measured on one B200 (cfg3): 54.4 us per step against 48.8 us for the eager loop -
    the graph's chains start together and stay in lockstep (all detect kernels, then all tails, HBM idle meanwhile),
    whereas the eager loop's launch cadence staggers them so a tail always hides under another batch's detect pass.
    Use it when the host, not the GPU, is the bottleneck.

    The inputs are the tensors whose ADDRESSES are baked into the graph: refill them in place between replays
    (e.g. the backbone writes its heads there).  Returns (graph, steps_per_replay); results are in each pipe's
    tables / `.outputs()` tensors after the replay, exactly as after eager calls.
    """
    import math

    if not pipes or not inputs:
        raise ValueError("need at least one pipeline and one input batch")
    if any(p.tail_stream is not None for p in pipes):
        raise ValueError("capture_rotation needs single-stream pipelines: the two-stream mode's hand-off events were last "
                         "recorded outside the capture, which CUDA does not allow a capturing stream to wait on")
    dev = pipes[0].device
    n = max(int(repeats), 1) * (len(pipes) * len(inputs) // math.gcd(len(pipes), len(inputs)))
    with torch.cuda.device(dev):
        if streams is None:
            streams = [torch.cuda.Stream(device=dev) for _ in pipes]
        for i in range(len(pipes)):  # warm-up outside capture: lazy module load / attribute setup must not be captured
            with torch.cuda.stream(streams[i]):
                pipes[i](*inputs[i % len(inputs)])
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=dev)
        # Stagger: chain k's FIRST detect kernel waits for chain k-1's first detect kernel (events recorded inside the
        # capture around the detect launches).  Without it the chains start together and stay in lockstep - all
        # detect kernels share HBM, then all tails run with HBM idle; offset by one detect pass, a tail always hides
        # under another chain's detect pass, as it does in the eager loop through the host's launch cadence.
        stagger = [(torch.cuda.Event(), torch.cuda.Event()) for _ in range(len(pipes))] if stagger_chains else []
        for pair in stagger:  # torch creates the cudaEvent lazily on first record; the C side re-records them in the capture
            for e in pair:
                e.record(torch.cuda.current_stream(dev))
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(graph, stream=cap):
            origin = torch.cuda.current_stream(dev)
            uniq = list({id(s): s for s in streams}.values())
            for s in uniq:
                s.wait_stream(origin)  # fork
            for i in range(n):
                with torch.cuda.stream(streams[i % len(pipes)]):
                    if stagger and i < len(pipes):
                        if i > 0 and streams[i] is not streams[i - 1]:
                            streams[i].wait_event(stagger[i - 1][1])
                        pipes[i](*inputs[i % len(inputs)], detect_events=stagger[i])
                    else:
                        pipes[i % len(pipes)](*inputs[i % len(inputs)])
            for s in uniq:
                origin.wait_stream(s)  # join
    return graph, n
