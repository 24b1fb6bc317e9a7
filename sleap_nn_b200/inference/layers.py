"""Device-resident `postprocess()` of the reference's inference layers.

Each class replaces the post-model half of one reference layer with a short launch chain and no host
synchronisation when the output shape is fixed:

    CentroidPostproc          CentroidLayer.postprocess        (inference/layers/centroid.py:196-258)
    CenteredInstancePostproc  CenteredInstanceLayer.postprocess (inference/layers/centered_instance.py:199-230)
                              + TopDownLayer._run_stage_2's un-crop / scatter (inference/layers/topdown.py:259-289)
    BottomUpPostproc          BottomUpLayer.postprocess         (sleap_nn_b200/pipeline.py)
    BottomUpMultiClassPostproc BottomUpMultiClassLayer.postprocess (inference/layers/bottomup_multiclass.py:75-190)

Knob names follow `PostprocessConfig` / `PreprocInfo`; returned tensors live on the device.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from sleap_nn_b200 import _native as N
from sleap_nn_b200.inference.ops.coord import make_ladder
from sleap_nn_b200.inference.ops.peaks import DEFAULT_PEAK_CAP, local_peaks_padded
from sleap_nn_b200.pipeline import BottomUpPostproc  # noqa: F401  (re-export: the third layer)


def _refine_size(refinement: Optional[str], integral_patch_size: int) -> int:
    return int(integral_patch_size) if refinement == "integral" else 0


class CentroidPostproc:
    """confmaps (B, 1, H, W) -> pred_centroids (B, max_instances, 2), pred_centroid_values (B, max_instances).

    find_local_peaks -> x stride -> / input_scale -> per-frame top-k by value when there are more peaks than
    `max_instances` -> NaN padding -> / eff_scale.  Two launches (K1 + the top-k epilogue reuse the padded peak
    table) plus the key sort; with `max_instances=None` the busiest frame's count is read back first
    (`_infer_max_instances`, centroid.py:264-270), which is the only host sync.
    """

    def __init__(self, peak_threshold: float = 0.2, refinement: Optional[str] = "integral", integral_patch_size: int = 5,
                 max_instances: Optional[int] = None, peak_cap: int = DEFAULT_PEAK_CAP):
        self.peak_threshold, self.refine_size = float(peak_threshold), _refine_size(refinement, integral_patch_size)
        self.max_instances, self.peak_cap = max_instances, int(peak_cap)

    def __call__(self, confmaps: torch.Tensor, output_stride: int = 1, input_scale: float = 1.0,
                 eff_scale: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if not confmaps.is_cuda or confmaps.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            raise TypeError("CentroidPostproc expects an fp32 / fp16 / bf16 CUDA tensor")
        dev, B = confmaps.device, int(confmaps.shape[0])
        with torch.cuda.device(dev):
            count, xy, val, _chan, status, cap = local_peaks_padded(confmaps.detach(), self.peak_threshold,
                                                                    self.refine_size, float(output_stride), self.peak_cap)
            max_inst = self.max_instances
            if not max_inst:
                # this path synchronises anyway (`_infer_max_instances`): a frame with more peaks than the table holds is
                # re-run with room for it, like ops.peaks._local_peaks, so the result never depends on which peaks the
                # overflowing table happened to keep
                worst = int(count.max().item()) if B else 0
                if worst > cap:
                    self.peak_cap = 1 << (worst - 1).bit_length()
                    count, xy, val, _chan, status, cap = local_peaks_padded(
                        confmaps.detach(), self.peak_threshold, self.refine_size, float(output_stride), self.peak_cap)
                max_inst = min(worst, cap)
            max_inst = max(int(max_inst), 1)  # always at least one slot (centroid.py:222-223)
            o_xy = torch.empty((B, max_inst, 2), dtype=torch.float32, device=dev)
            o_val = torch.empty((B, max_inst), dtype=torch.float32, device=dev)
            eff = None if eff_scale is None else eff_scale.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
            if eff is not None and eff.numel() == 1 and B != 1:
                eff = eff.expand(B).contiguous()
            N.check(N.lib.snb_peaks_topk(N.ptr(count), B, cap, N.ptr(xy), N.ptr(val), max_inst, float(input_scale),
                                         N.ptr(eff), N.ptr(o_xy), N.ptr(o_val), N.stream_ptr(dev)), "snb_peaks_topk")
        self.last_status = status  # SNB_STATUS_PEAK_OVERFLOW if a frame had more than peak_cap peaks
        return o_xy, o_val

    def check(self) -> None:
        """Synchronises: raises if a frame of the last call had more peaks than `peak_cap` (with a fixed `max_instances`
        nothing is read back during the call, so which peaks an overflowing table kept - and therefore the top-k - would
        depend on arrival order; the reference has no such limit).  Re-run with a larger `peak_cap`."""
        st = getattr(self, "last_status", None)
        if st is not None and int(st.item()) & N.STATUS_PEAK_OVERFLOW:
            st.zero_()
            raise RuntimeError(f"CentroidPostproc: a frame had more than peak_cap={self.peak_cap} peaks; raise peak_cap")


class CenteredInstancePostproc:
    """confmaps (n, N, h, w) -> keypoints + values in ONE launch, ladder and un-crop scatter included.

    Stand-alone (CenteredInstanceLayer): returns (n, 1, N, 2), (n, 1, N) with stride / input_scale / eff_scale
    undone.  Inside a top-down pipeline pass `crop_topleft` (n, 2), `per_crop_eff_scale` (n,) and
    `scatter_rows` (n,) = b * max_instances + i of each crop together with `out_shape=(B, max_instances)`:
    the kernel adds the crop offset, divides by the per-crop scale and writes straight into the NaN-filled
    (B, max_instances, N, 2) / (B, max_instances, N) tensors (topdown.py:259-289).
    """

    def __init__(self, peak_threshold: float = 0.2, refinement: Optional[str] = "integral", integral_patch_size: int = 5):
        self.peak_threshold, self.refine_size = float(peak_threshold), _refine_size(refinement, integral_patch_size)
        self._ws = {}  # (device, stream) -> ticket / partials workspace: two calls in flight on different streams
        #                must not share the self-resetting tickets of the chunked kernels

    def __call__(self, confmaps: torch.Tensor, output_stride: int = 1, input_scale: float = 1.0,
                 eff_scale: Optional[torch.Tensor] = None, crop_topleft: Optional[torch.Tensor] = None,
                 per_crop_eff_scale: Optional[torch.Tensor] = None, scatter_rows: Optional[torch.Tensor] = None,
                 out_shape: Optional[Tuple[int, int]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if not confmaps.is_cuda or confmaps.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            raise TypeError("CenteredInstancePostproc expects an fp32 / fp16 / bf16 CUDA tensor")
        dev = confmaps.device
        n, Cn, H, W = (int(v) for v in confmaps.shape)
        with torch.cuda.device(dev):
            if scatter_rows is not None:
                if out_shape is None:
                    raise ValueError("scatter_rows needs out_shape=(B, max_instances)")
                rows = int(out_shape[0]) * int(out_shape[1])
                kpts = torch.full((rows, Cn, 2), float("nan"), dtype=torch.float32, device=dev)
                vals = torch.full((rows, Cn), float("nan"), dtype=torch.float32, device=dev)
            else:
                kpts = torch.empty((n, Cn, 2), dtype=torch.float32, device=dev)
                vals = torch.empty((n, Cn), dtype=torch.float32, device=dev)
            if n * Cn:
                rpc, nch, nbytes = C.c_int(), C.c_int(), C.c_longlong()
                N.check(N.lib.snb_global_peaks_workspace(n, Cn, H, W, C.byref(rpc), C.byref(nch), C.byref(nbytes)), "workspace")
                need = (nbytes.value + 3) // 4
                ws_key = (dev.index, N.stream_ptr(dev))
                ws = self._ws.get(ws_key)
                if ws is None or ws.numel() < need:
                    if len(self._ws) >= 16:
                        self._ws.clear()
                    ws = self._ws[ws_key] = torch.zeros((need,), dtype=torch.int32, device=dev)  # tickets self-reset after each use
                lad, keep = make_ladder(dev, n, stride=float(output_stride), input_scale=float(input_scale),
                                        eff_scale=eff_scale, crop_offset=crop_topleft, eff_scale2=per_crop_eff_scale,
                                        scatter=scatter_rows)
                x = confmaps.detach()
                N.check(N.lib.snb_global_peaks_t(N.ptr(x), N.dtype_code(x.dtype), n, Cn, H, W, *x.stride(), self.peak_threshold,
                                                 self.refine_size, N.ptr(ws), C.byref(lad), N.ptr(kpts), N.ptr(vals),
                                                 N.stream_ptr(dev)), "snb_global_peaks_t")
                del keep
        if scatter_rows is not None:
            return kpts.view(out_shape[0], out_shape[1], Cn, 2), vals.view(out_shape[0], out_shape[1], Cn)
        return kpts.unsqueeze(1), vals.unsqueeze(1)


class BottomUpMultiClassPostproc:
    """confmaps (B, N, H, W) + class maps (B, K, Hc, Wc) -> per-class instances, all on the device.

    `BottomUpMultiClassLayer.postprocess` (inference/layers/bottomup_multiclass.py:75-146): find_local_peaks ->
    x cms_output_stride -> / class_maps_output_stride -> classify_peaks_from_maps (gather under the rounded peak,
    per-(frame, node) optimal assignment, arg-max filter) -> x class_maps_output_stride -> / input_scale ->
    / eff_scale -> nanmean scores -> `_cap_instances_by_score`.  Three launches (K1 + key sort/refine, the
    classification kernel, the epilogue), no host synchronisation; slot k of the outputs IS class k.

    Returns `(pred_keypoints (B, K, N, 2), pred_peak_values (B, K, N), instance_scores (B, K),
    instance_tracking_scores (B, K))`.  Status bits (a NaN class probability, a frame with more than `peak_cap`
    peaks, more than 128 peaks of one node in one frame) are raised by `.check()`, the only call that synchronises.
    """

    def __init__(self, peak_threshold: float = 0.2, refinement: Optional[str] = "integral", integral_patch_size: int = 5,
                 cms_output_stride: int = 2, class_maps_output_stride: int = 2, max_instances: Optional[int] = None,
                 peak_cap: int = DEFAULT_PEAK_CAP):
        self.peak_threshold, self.refine_size = float(peak_threshold), _refine_size(refinement, integral_patch_size)
        self.cms_output_stride, self.class_maps_output_stride = cms_output_stride, class_maps_output_stride
        self.max_instances, self.peak_cap = max_instances, int(peak_cap)
        self._status = []

    def __call__(self, confmaps: torch.Tensor, class_maps: torch.Tensor, input_scale: float = 1.0,
                 eff_scale: Optional[torch.Tensor] = None):
        if not (confmaps.is_cuda and class_maps.is_cuda) or confmaps.dtype != torch.float32 or class_maps.dtype != torch.float32:
            raise TypeError("BottomUpMultiClassPostproc expects fp32 CUDA tensors")
        dev = confmaps.device
        B, Nn = int(confmaps.shape[0]), int(confmaps.shape[1])
        if int(class_maps.shape[0]) != B:
            raise ValueError("confmaps and class maps must hold the same frames")
        K, Hc, Wc = (int(v) for v in class_maps.shape[1:])
        with torch.cuda.device(dev):
            count, xy, val, chan, status, cap = local_peaks_padded(confmaps.detach(), self.peak_threshold, self.refine_size,
                                                                   float(self.cms_output_stride), self.peak_cap)
            f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
            probs = f32(max(B * cap, 1), max(K, 1))
            c_xy, c_val, c_prob = f32(B, K, Nn, 2), f32(B, K, Nn), f32(B, K, Nn)
            cm = class_maps.detach()
            st = N.stream_ptr(dev)
            N.check(N.lib.snb_classify_peaks_padded(N.ptr(cm), B, K, Hc, Wc, *cm.stride(), N.ptr(count), cap, N.ptr(xy),
                                                    N.ptr(val), N.ptr(chan), float(self.class_maps_output_stride), Nn,
                                                    N.ptr(probs), N.ptr(c_xy), N.ptr(c_val), N.ptr(c_prob), N.ptr(status),
                                                    st), "snb_classify_peaks_padded")
            eff = None if eff_scale is None else eff_scale.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
            if eff is not None and eff.numel() == 1 and B != 1:
                eff = eff.expand(B).contiguous()
            if eff is not None and eff.numel() != B:
                raise ValueError("eff_scale must hold one factor per frame")
            kpts, vals, scores, tracking = f32(B, K, Nn, 2), f32(B, K, Nn), f32(B, K), f32(B, K)
            N.check(N.lib.snb_multiclass_outputs(N.ptr(c_xy), N.ptr(c_val), N.ptr(c_prob), B, K, Nn,
                                                 float(self.class_maps_output_stride), float(input_scale), N.ptr(eff),
                                                 -1 if self.max_instances is None else int(self.max_instances),
                                                 N.ptr(kpts), N.ptr(vals), N.ptr(scores), N.ptr(tracking), st),
                    "snb_multiclass_outputs")
        self._status.append(status)
        del self._status[:-8]
        return kpts, vals, scores, tracking

    def check(self) -> None:
        """Synchronise and raise if any recent call overflowed a table or met an invalid class probability."""
        bits = 0
        for s in self._status:
            bits |= int(s.item())
        self._status.clear()
        if bits & N.STATUS_LSAP_INVALID:
            raise ValueError("matrix contains invalid numeric entries")
        if bits & N.STATUS_LSAP_INFEASIBLE:
            raise ValueError("cost matrix is infeasible")
        if bits:
            raise RuntimeError(f"multi-class post-processing overflowed a fixed-capacity table (status 0x{bits:x})")


class SingleInstancePostproc:
    """confmaps (B, N, H, W) -> pred_keypoints (B, 1, N, 2), pred_peak_values (B, 1, N) in ONE launch.

    `SingleInstanceLayer.postprocess` (inference/layers/single_instance.py:71-106): find_global_peaks -> undo_stride ->
    undo_input_scale -> undo_eff_scale, the ladder running in the arg-max kernel's epilogue.
    """

    def __init__(self, peak_threshold: float = 0.2, refinement: Optional[str] = "integral", integral_patch_size: int = 5):
        self._inner = CenteredInstancePostproc(peak_threshold, refinement, integral_patch_size)

    def __call__(self, confmaps: torch.Tensor, output_stride: int = 1, input_scale: float = 1.0,
                 eff_scale: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._inner(confmaps, output_stride=output_stride, input_scale=input_scale, eff_scale=eff_scale)


def _aligned_ws(nbytes: int, dev: torch.device) -> Tuple[torch.Tensor, int]:
    ws = torch.empty((int(nbytes) + 15,), dtype=torch.uint8, device=dev)
    return ws, (ws.data_ptr() + 15) & ~15


def _raise_for_class_status(status: torch.Tensor) -> None:
    bits = int(status.item())
    if bits & N.STATUS_LSAP_INVALID:
        raise ValueError("matrix contains invalid numeric entries")
    if bits & N.STATUS_LSAP_INFEASIBLE:
        raise ValueError("cost matrix is infeasible")


class CenteredInstanceMultiClassPostproc(CenteredInstancePostproc):
    """Stand-alone `CenteredInstanceMultiClassLayer.postprocess` (inference/layers/topdown_multiclass.py:79-145).

    confmaps (n, N, h, w) + class vectors (n, K) -> dict with `pred_keypoints` (n, 1, N, 2), `pred_peak_values`
    (n, 1, N), `pred_class_inds` (n, 1, N) int64, `pred_class_probs` (n, 1, K) (the raw vectors) and
    `instance_tracking_scores` (n, 1): ONE assignment over all crops, as the stand-alone layer does.  Inside a
    top-down pipeline the assignment must run per frame: `TopDownPostproc` does that on the device.
    """

    def classify(self, confmaps: torch.Tensor, class_vectors: torch.Tensor, output_stride: int = 1,
                 input_scale: float = 1.0, eff_scale: Optional[torch.Tensor] = None) -> dict:
        kpts, vals = self(confmaps, output_stride=output_stride, input_scale=input_scale, eff_scale=eff_scale)
        dev = confmaps.device
        n, K = (int(v) for v in class_vectors.shape)
        Nn = int(confmaps.shape[1])
        probs = class_vectors.detach().to(device=dev, dtype=torch.float32).contiguous()
        inds = torch.full((n,), -1, dtype=torch.int64, device=dev)
        cp = torch.full((n,), float("nan"), dtype=torch.float32, device=dev)
        if n:
            with torch.cuda.device(dev):
                ws, ws_ptr = _aligned_ws(N.lib.snb_class_inds_workspace_bytes(n, K), dev)
                status = torch.zeros((1,), dtype=torch.int32, device=dev)
                N.check(N.lib.snb_class_inds_from_vectors(N.ptr(probs), n, K, ws_ptr, N.ptr(inds), N.ptr(cp), N.ptr(status),
                                                          N.stream_ptr(dev)), "snb_class_inds_from_vectors")
                _raise_for_class_status(status)
                del ws
        return dict(pred_keypoints=kpts, pred_peak_values=vals,
                    pred_class_inds=inds.view(n, 1, 1).expand(-1, -1, Nn), pred_class_probs=probs.unsqueeze(1),
                    instance_tracking_scores=cp.unsqueeze(1))


class TopDownPostproc:
    """`TopDownLayer.predict` after the centroid stage (inference/layers/topdown.py:98-289), on the device.

    `__call__(image, centroids, centroid_vals, model, ...)`:

    1. `snb_topdown_select` - NaN-centroid mask, optional greedy centroid NMS (`_centroid_nms_mask`, topdown.py:415-446),
       the valid (b, i) list in `nonzero` order, sized-space crop boxes, image-space centroids and boxes: one launch.
    2. the number of crops is read back (the crop tensor's batch dimension; the reference's `nonzero` syncs too).
    3. `snb_crop_bboxes` picks the crops out of the (uint8 or float) sized image.
    4. `model(crops)` - the caller's centred-instance network (PyTorch / cuDNN) -> confmaps (n, N, h, w), or
       `(confmaps, class_vectors (n, K))` for the multi-class variant.
    5. `snb_global_peaks_ex` (arg-max + refinement + the stage-2 ladder) and `snb_topdown_lift` (crop offset, per-crop
       eff_scale, scatter into the NaN-filled (B, max_inst, ...) outputs): two launches.
    6. multi-class only: `snb_class_inds_grouped` - one assignment PER FRAME (topdown.py:343-371), scattered.

    Returns a dict with the reference `Outputs` field names.
    """

    def __init__(self, crop_size: Tuple[int, int], peak_threshold: float = 0.2, refinement: Optional[str] = "integral",
                 integral_patch_size: int = 5, centroid_nms: bool = False, centroid_nms_threshold: float = 0.5,
                 return_crops: bool = False, return_class_vectors: bool = False, n_nodes: int = 1):
        self.crop_size = (int(crop_size[0]), int(crop_size[1]))
        self.stage2 = CenteredInstancePostproc(peak_threshold, refinement, integral_patch_size)
        self.centroid_nms, self.centroid_nms_threshold = bool(centroid_nms), float(centroid_nms_threshold)
        self.return_crops, self.return_class_vectors = bool(return_crops), bool(return_class_vectors)
        self.n_nodes = int(n_nodes)  # only used for the shape of the all-NaN result when no centroid is valid

    def select(self, centroids: torch.Tensor, centroid_vals: torch.Tensor, eff_scale: Optional[torch.Tensor] = None) -> dict:
        """Stage B + crop list (step 1); every tensor stays on the device."""
        if not centroids.is_cuda:
            raise TypeError("TopDownPostproc expects CUDA tensors")
        dev = centroids.device
        B, I = int(centroids.shape[0]), int(centroids.shape[1])
        cen = centroids.detach().to(torch.float32).contiguous()
        val = centroid_vals.detach().to(device=dev, dtype=torch.float32).contiguous()
        eff = None if eff_scale is None else eff_scale.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if eff is not None and eff.numel() != B:
            raise ValueError("eff_scale must hold one factor per frame")
        cap = max(B * I, 1)
        i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        out = dict(n_valid=i32(3), frame_off=i32(B + 1), sample_inds=torch.empty((cap,), dtype=torch.int64, device=dev),
                   rows=i32(cap), row_to_crop=i32(cap), crop_bboxes=f32(cap, 4, 2), crop_topleft=f32(cap, 2),
                   crop_eff=f32(cap), valid_mask=torch.empty((B, I), dtype=torch.uint8, device=dev),
                   centroids_img=f32(B, I, 2), full_bboxes=f32(B, I, 4, 2), centroid_vals=val)
        with torch.cuda.device(dev):
            N.check(N.lib.snb_topdown_select(N.ptr(cen), N.ptr(val), B, I, N.ptr(eff), self.crop_size[0], self.crop_size[1],
                                             int(self.centroid_nms), self.centroid_nms_threshold, N.ptr(out["n_valid"]),
                                             N.ptr(out["frame_off"]), N.ptr(out["sample_inds"]), N.ptr(out["rows"]),
                                             N.ptr(out["row_to_crop"]), N.ptr(out["crop_bboxes"]), N.ptr(out["crop_topleft"]),
                                             N.ptr(out["crop_eff"]), N.ptr(out["valid_mask"]), N.ptr(out["centroids_img"]),
                                             N.ptr(out["full_bboxes"]), N.stream_ptr(dev)), "snb_topdown_select")
        return out

    def __call__(self, image: torch.Tensor, centroids: torch.Tensor, centroid_vals: torch.Tensor, model,
                 eff_scale: Optional[torch.Tensor] = None, output_stride: int = 1, input_scale: float = 1.0) -> dict:
        if not image.is_cuda or image.dim() != 4:
            raise TypeError("TopDownPostproc expects a (B, C, H, W) CUDA image (the sized image of the centroid stage)")
        dev = image.device
        B, I = int(centroids.shape[0]), int(centroids.shape[1])
        sel = self.select(centroids, centroid_vals, eff_scale)
        # the one host sync: the crop tensor's batch dimension, and the crop size as crop_bboxes reads it off bbox 0
        # (ops/crops.py:66-67; normally == crop_size, one less when fp32 rounding of the first box says so)
        n, ch, cw = (int(v) for v in sel["n_valid"].tolist())
        st = lambda: N.stream_ptr(dev)
        res = dict(pred_centroids=sel["centroids_img"], pred_centroid_values=sel["centroid_vals"],
                   instance_scores=sel["centroid_vals"], valid_mask=sel["valid_mask"].bool())
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            kp = vals = class_vectors = None
            Nn = self.n_nodes
            if n:
                S, Cn, H, W = (int(v) for v in image.shape)
                crops = torch.empty((n, Cn, ch, cw), dtype=image.dtype, device=dev)
                status = torch.zeros((1,), dtype=torch.int32, device=dev)
                N.check(N.lib.snb_crop_bboxes(N.ptr(image), image.element_size(), S, Cn, H, W, *image.stride(),
                                              N.ptr(sel["crop_bboxes"]), N.ptr(sel["sample_inds"]), n, ch, cw, N.ptr(crops),
                                              N.ptr(status), st()), "snb_crop_bboxes")
                raw = model(crops)
                cms, class_vectors = raw if isinstance(raw, (tuple, list)) else (raw, None)
                if cms.dtype not in (torch.float32, torch.float16, torch.bfloat16):
                    cms = cms.float()  # fp16 / bf16 maps of an autocast network are read natively by the kernel
                Nn = int(cms.shape[1])
                k4, v3 = self.stage2(cms, output_stride=output_stride, input_scale=input_scale)
                kp, vals = k4.squeeze(1).contiguous(), v3.squeeze(1).contiguous()
                if self.return_crops:
                    if (ch, cw) != self.crop_size:  # the reference's scatter into (.., crop_h, crop_w) raises here too
                        raise RuntimeError(f"shape mismatch: value tensor of shape [{n}, {Cn}, {ch}, {cw}] cannot be broadcast "
                                           f"to indexing result of shape [{n}, {Cn}, {self.crop_size[0]}, {self.crop_size[1]}]")
                    full = torch.zeros((B * I, Cn, ch, cw), dtype=crops.dtype, device=dev)
                    full.index_copy_(0, sel["rows"][:n].long(), crops)  # topdown.py:293-303 (debug output)
                    res["crops"] = full.view(B, I, Cn, ch, cw)
            full_kpts, full_crop, full_vals = f32(B, I, Nn, 2), f32(B, I, Nn, 2), f32(B, I, Nn)
            N.check(N.lib.snb_topdown_lift(N.ptr(kp), N.ptr(vals), Nn, B * I, N.ptr(sel["row_to_crop"]),
                                           N.ptr(sel["crop_topleft"]), N.ptr(sel["crop_eff"]), N.ptr(full_kpts),
                                           N.ptr(full_crop), N.ptr(full_vals), st()), "snb_topdown_lift")
            res.update(pred_keypoints=full_kpts, pred_peak_values=full_vals)
            if n:  # the reference's empty early return carries no crop keypoints / boxes (topdown.py:218-233)
                res.update(pred_crop_keypoints=full_crop, instance_bboxes=sel["full_bboxes"])
            if class_vectors is not None:
                K = int(class_vectors.shape[1])
                probs = class_vectors.detach().to(device=dev, dtype=torch.float32).contiguous()
                ws, ws_ptr = _aligned_ws(N.lib.snb_class_inds_grouped_workspace_bytes(B, I, K), dev)
                cls = torch.empty((B, I, Nn), dtype=torch.int64, device=dev)
                trk = f32(B, I)
                vec = f32(B, I, K) if self.return_class_vectors else None
                status = torch.zeros((1,), dtype=torch.int32, device=dev)
                N.check(N.lib.snb_class_inds_grouped(N.ptr(probs), K, N.ptr(sel["frame_off"]), N.ptr(sel["rows"]), B, I, Nn,
                                                     ws_ptr, N.ptr(cls), N.ptr(trk), N.ptr(vec), N.ptr(status), st()),
                        "snb_class_inds_grouped")
                self.last_class_status = status  # LSAP_INVALID / INFEASIBLE bits; `check()` raises like scipy
                res.update(pred_class_inds=cls, instance_tracking_scores=trk)
                if vec is not None:
                    res["pred_class_vectors"] = vec
                del ws
        return res

    def check(self) -> None:
        """Synchronise and raise (ValueError, like scipy) if the last class assignment met an invalid matrix."""
        s = getattr(self, "last_class_status", None)
        if s is not None:
            _raise_for_class_status(s)
