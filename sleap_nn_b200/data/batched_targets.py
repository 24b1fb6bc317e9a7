"""Batched, device-resident target generation for the datasets' call sites.

The reference builds every training target on the CPU, one frame at a time, inside `Dataset.__getitem__`
(sleap_nn/data/custom_datasets.py:1305-1327 bottom-up, :1489-1511 multi-class bottom-up, :1788 centered
instance, :2835 centroid, :2986 single instance): `generate_multiconfmaps`, `generate_pafs`,
`generate_class_maps` and `generate_confmaps`, each a chain of full-frame tensor ops (0.84 s + 3.46 s per
cfg4 frame).  `BatchedTargets` produces the same tensors for a whole collated batch in one
kernel launch per target, straight into device memory where the training step consumes them: the instance
slice `[:, :num_instances]`, `filter_oob_points`, `get_edge_points` and `generate_pafs`' in-image instance
filter all happen inside the kernels (csrc/targets.cu, csrc/identity.cu).  Outputs have the COLLATED shapes of
the reference samples, so a training loop can swap `batch["confidence_maps"]` for these tensors.

Augmentation order is unchanged: these functions take the (already augmented, already resized) instance
coordinates a dataset would pass to `generate_*`.
"""

from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from sleap_nn_b200 import _native as N


class BatchedTargets:
    """Target maps for a batch of frames of one image size, generated on `device`.

    Args:
        img_hw: (height, width) of the (preprocessed) images the instances live in.
        device: CUDA device the targets are produced on.
        out_dtype: torch.float32 (the reference's dtype) or torch.bfloat16 for bf16 training pipelines.
    """

    def __init__(self, img_hw: Tuple[int, int], device: Optional[torch.device] = None,
                 out_dtype: torch.dtype = torch.float32):
        if out_dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("targets are produced in float32 or bfloat16")
        self.img_hw = (int(img_hw[0]), int(img_hw[1]))
        self.device = torch.device(device) if device is not None else N.compute_device()
        if self.device.type != "cuda":
            raise N.NativeLibraryError("BatchedTargets needs a CUDA device: there is no CPU fallback")
        self.out_dtype = out_dtype
        self._grids: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}

    # ------------------------------------------------------------------ helpers
    def _grid(self, stride: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """make_grid_vectors (data/utils.py:55-85) cached on the device; arange values are exact in fp32."""
        if stride not in self._grids:
            h, w = self.img_hw
            self._grids[stride] = (torch.arange(0, w, step=stride, dtype=torch.float32, device=self.device),
                                   torch.arange(0, h, step=stride, dtype=torch.float32, device=self.device))
        return self._grids[stride]

    def _f32(self, t: torch.Tensor) -> torch.Tensor:
        return t.detach().to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()

    def _counts(self, num_instances, B: int) -> Optional[torch.Tensor]:
        if num_instances is None:
            return None
        t = torch.as_tensor(num_instances).reshape(-1)
        if t.numel() != B:
            raise ValueError(f"num_instances must have one entry per frame ({B}), got {t.numel()}")
        return t.to(device=self.device, dtype=torch.int32, non_blocking=True).contiguous()

    def _confmaps(self, pts: torch.Tensor, G: int, I: int, Nn: int, strides, n_valid, sigma: float, stride: int,
                  oob: bool, out_dtype=None) -> torch.Tensor:
        xv, yv = self._grid(stride)
        h, w = int(yv.shape[0]), int(xv.shape[0])
        dt = out_dtype or self.out_dtype
        out = torch.empty((G, Nn, h, w), dtype=dt, device=self.device)
        H, W = self.img_hw
        with torch.cuda.device(self.device):
            N.check(N.lib.snb_confmaps_ex(N.ptr(pts), G, I, Nn, *strides, N.ptr(n_valid), float(W) if oob else 0.0,
                                          float(H) if oob else 0.0, N.ptr(xv), N.ptr(yv), h, w, float(2 * sigma**2),
                                          int(dt == torch.bfloat16), N.ptr(out), N.stream_ptr(self.device)),
                    "snb_confmaps_ex")
        return out

    # ------------------------------------------------------------------ targets
    def multi_confmaps(self, instances: torch.Tensor, num_instances=None, sigma: float = 1.5, output_stride: int = 2,
                       is_centroids: bool = False) -> torch.Tensor:
        """`generate_multiconfmaps` (data/confidence_maps.py:46-91) for every frame of the batch.

        instances (B, I, N, 2) - or (B, 1, I, N, 2) as collated - or, with `is_centroids`, (B, I, 2);
        num_instances (B,) or None.  Returns (B, 1, N, h, w) ((B, 1, 1, h, w) for centroids).
        """
        x = self._f32(instances)
        if is_centroids:
            x = x.reshape(x.shape[0], -1, 2)
            B, I = int(x.shape[0]), int(x.shape[1])
            out = self._confmaps(x, B, I, 1, (I * 2, 2, 0), self._counts(num_instances, B), sigma * output_stride,
                                 output_stride, False)
        else:
            x = x.reshape(x.shape[0], -1, x.shape[-2], 2)
            B, I, Nn = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
            out = self._confmaps(x, B, I, Nn, (I * Nn * 2, Nn * 2, 2), self._counts(num_instances, B),
                                 sigma * output_stride, output_stride, False)
        return out.unsqueeze(1)

    def confmaps(self, instance: torch.Tensor, sigma: float = 1.5, output_stride: int = 2,
                 filter_oob: bool = False) -> torch.Tensor:
        """`generate_confmaps` (data/confidence_maps.py:8-43) per frame / crop: instance (B, N, 2) or (B, 1, N, 2) ->
        (B, 1, N, h, w).  `filter_oob` applies `filter_oob_points` (data/providers.py:38-69) first, as the
        centered-instance and single-instance datasets do (custom_datasets.py:1784, :2980)."""
        x = self._f32(instance)
        x = x.reshape(x.shape[0], -1, 2)
        B, Nn = int(x.shape[0]), int(x.shape[1])
        out = self._confmaps(x, B, 1, Nn, (Nn * 2, 0, 2), None, sigma * output_stride, output_stride, filter_oob)
        return out.unsqueeze(1)

    def pafs(self, instances: torch.Tensor, edge_inds: Sequence[Sequence[int]], sigma: float = 1.5, output_stride: int = 2,
             flatten_channels: bool = True) -> torch.Tensor:
        """`generate_pafs` (data/edge_maps.py:250-323) per frame: instances (B, I, N, 2) (or (B, 1, I, N, 2)) ->
        (B, 2E, h, w), or (B, E, 2, h, w) without `flatten_channels`.  The PAF sigma is NOT scaled by the stride and
        ALL instance slots take part (the reference does not slice by num_instances here); instances without a
        node strictly inside the grid extent are dropped."""
        x = self._f32(instances)
        x = x.reshape(x.shape[0], -1, x.shape[-2], 2)
        B, I, Nn = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        e = torch.as_tensor(edge_inds).reshape(-1, 2).to(device=self.device, dtype=torch.int32).contiguous()
        E = int(e.shape[0])
        if E and (int(e.min()) < 0 or int(e.max()) >= Nn):
            raise IndexError("edge_inds refers to a node outside the skeleton")
        xv, yv = self._grid(output_stride)
        h, w = int(yv.shape[0]), int(xv.shape[0])
        out = torch.empty((B, E, 2, h, w), dtype=self.out_dtype, device=self.device)
        # bound = (xv[-1], yv[-1]) (edge_maps.py:293-296); arange values are exact, so computed on the host
        xmax = float(((self.img_hw[1] - 1) // output_stride) * output_stride)
        ymax = float(((self.img_hw[0] - 1) // output_stride) * output_stride)
        with torch.cuda.device(self.device):
            N.check(N.lib.snb_pafs_from_instances(N.ptr(x), B, I, Nn, N.ptr(e), E, xmax, ymax, N.ptr(xv), N.ptr(yv), h, w,
                                                  float(2 * sigma**2), int(self.out_dtype == torch.bfloat16), N.ptr(out),
                                                  N.stream_ptr(self.device)), "snb_pafs_from_instances")
        return out.reshape(B, 2 * E, h, w) if flatten_channels else out

    def class_maps(self, instances: torch.Tensor, num_instances, class_inds: torch.Tensor, num_tracks: int,
                   class_map_threshold: float = 0.2, sigma: float = 1.5, output_stride: int = 2,
                   is_centroids: bool = False) -> torch.Tensor:
        """`generate_class_maps` (data/identity.py:85-137) per frame -> (B, 1, num_tracks, h, w).

        instances (B, I, N, 2) or centroids (B, I, 2); class_inds (B, I) track index per instance slot (-1 = none);
        num_instances (B,): slots beyond it do not exist for that frame.
        """
        x = self._f32(instances)
        if is_centroids:
            x = x.reshape(x.shape[0], -1, 2)
            B, I = int(x.shape[0]), int(x.shape[1])
            strides, n_nodes = (I * 2, 0, 2), 1
        else:
            x = x.reshape(x.shape[0], -1, x.shape[-2], 2)
            B, I, n_nodes = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
            strides = (I * n_nodes * 2, 2, n_nodes * 2)  # instances <-> channels swapped (data/identity.py:122-129)
        # per-instance confidence maps: "instances" of the kernel = the nodes, "channels" = the instance slots
        cms = self._confmaps(x, B, n_nodes, I, strides, None, sigma * output_stride, output_stride, False,
                             out_dtype=torch.float32)
        h, w = int(cms.shape[-2]), int(cms.shape[-1])
        ci = torch.as_tensor(class_inds).reshape(B, I).to(self.device)
        if ci.dtype.is_floating_point:
            ci = torch.where(ci >= 0, ci, torch.full_like(ci, -1.0))
        ci = ci.to(torch.int32).contiguous()
        out = torch.empty((B, int(num_tracks), h, w), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(N.lib.snb_class_maps(N.ptr(cms), N.ptr(ci), N.ptr(self._counts(num_instances, B)), B, I, int(num_tracks),
                                         h, w, float(class_map_threshold), N.ptr(out), N.stream_ptr(self.device)),
                    "snb_class_maps")
        if self.out_dtype != torch.float32:
            out = out.to(self.out_dtype)
        return out.unsqueeze(1)

    # ------------------------------------------------------------------ dataset-shaped entry points
    def bottomup(self, instances, num_instances, edge_inds, confmap_sigma=1.5, confmap_stride=2, paf_sigma=1.5,
                 paf_stride=2, fused: bool = True) -> Dict[str, torch.Tensor]:
        """BottomUpDataset.__getitem__ targets (custom_datasets.py:1305-1327), collated: confidence maps
        (B, 1, N, h, w) and part-affinity fields (B, 2E, h', w').

        Both targets come out of ONE call (`snb_bottomup_targets`: the two kernels as a programmatic-dependent-launch
        pair that share the SMs), so a single frame - the reference's granularity - fills the GPU; values are
        bit-identical to `multi_confmaps` + `pafs` (`fused=False` issues those two calls instead)."""
        if fused:
            out = self._bottomup_fused(instances, num_instances, edge_inds, confmap_sigma, confmap_stride, paf_sigma, paf_stride)
            if out is not None:
                return out
        return {"confidence_maps": self.multi_confmaps(instances, num_instances, confmap_sigma, confmap_stride),
                "part_affinity_fields": self.pafs(instances, edge_inds, paf_sigma, paf_stride, flatten_channels=True)}

    def _bottomup_fused(self, instances, num_instances, edge_inds, confmap_sigma, confmap_stride, paf_sigma, paf_stride):
        x = self._f32(instances)
        x = x.reshape(x.shape[0], -1, x.shape[-2], 2)
        B, I, Nn = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        e = torch.as_tensor(edge_inds).reshape(-1, 2).to(device=self.device, dtype=torch.int32).contiguous()
        E = int(e.shape[0])
        if E and (int(e.min()) < 0 or int(e.max()) >= Nn):
            raise IndexError("edge_inds refers to a node outside the skeleton")
        xv7, yv7 = self._grid(confmap_stride)
        xv8, yv8 = self._grid(paf_stride)
        h7, w7, h8, w8 = int(yv7.shape[0]), int(xv7.shape[0]), int(yv8.shape[0]), int(xv8.shape[0])
        bf16 = self.out_dtype == torch.bfloat16
        if B == 0:
            return None
        n_valid = self._counts(num_instances, B)
        xmax = float(((self.img_hw[1] - 1) // paf_stride) * paf_stride)
        ymax = float(((self.img_hw[0] - 1) // paf_stride) * paf_stride)
        with torch.cuda.device(self.device):
            st = N.stream_ptr(self.device)
            cms = torch.empty((B, Nn, h7, w7), dtype=self.out_dtype, device=self.device)
            pafs = torch.empty((B, E, 2, h8, w8), dtype=self.out_dtype, device=self.device)
            sig7 = confmap_sigma * confmap_stride  # generate_multiconfmaps scales sigma by the stride, generate_pafs does not
            rc = N.lib.snb_bottomup_targets(N.ptr(x), B, I, Nn, N.ptr(n_valid), 0.0, 0.0, N.ptr(e), E, xmax, ymax,
                                            N.ptr(xv7), N.ptr(yv7), h7, w7, float(2 * sig7**2), N.ptr(xv8), N.ptr(yv8), h8, w8,
                                            float(2 * paf_sigma**2), int(bf16), N.ptr(cms), N.ptr(pafs), st)
        if rc == -2:  # SNB_ERR_UNSUPPORTED: shapes / smem the fused kernel does not take
            return None
        N.check(rc, "snb_bottomup_targets")
        return {"confidence_maps": cms.unsqueeze(1), "part_affinity_fields": pafs.reshape(B, 2 * E, h8, w8)}

    def bottomup_multiclass(self, instances, num_instances, class_inds, num_tracks, confmap_sigma=1.5, confmap_stride=2,
                            class_map_threshold=0.2, class_map_sigma=1.5, class_map_stride=2) -> Dict[str, torch.Tensor]:
        """BottomUpMultiClassDataset.__getitem__ targets (custom_datasets.py:1489-1511), collated."""
        return {"confidence_maps": self.multi_confmaps(instances, num_instances, confmap_sigma, confmap_stride),
                "class_maps": self.class_maps(instances, num_instances, class_inds, num_tracks, class_map_threshold,
                                              class_map_sigma, class_map_stride)}

    def centroid(self, centroids, num_instances, sigma=1.5, output_stride=2) -> Dict[str, torch.Tensor]:
        """CentroidDataset.__getitem__ targets (custom_datasets.py:2835-2842), collated."""
        return {"centroids_confidence_maps": self.multi_confmaps(centroids, num_instances, sigma, output_stride,
                                                                 is_centroids=True)}

    def centered_instance(self, instance, sigma=1.5, output_stride=2) -> Dict[str, torch.Tensor]:
        """CenteredInstanceDataset / SingleInstanceDataset targets (custom_datasets.py:1784-1793, :2980-2991): OOB
        keypoints are dropped, then one Gaussian per node."""
        return {"confidence_maps": self.confmaps(instance, sigma, output_stride, filter_oob=True)}
