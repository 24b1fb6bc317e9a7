// Top-down composition on device: stage B and the bookkeeping of stage 2 of
// sleap_nn/inference/layers/topdown.py (cited topdown.py:NN) - NaN-centroid mask (:98-104), optional greedy
// centroid NMS by centred-bbox IoU (:415-466), the valid (b, i) list in torch.nonzero order (:205-207), sized-space
// crop boxes (:231-236), and the lift of the centred-instance peaks back into the (B, max_inst, ...) outputs
// (:259-291).  Everything here is O(B * max_inst) on a few KB - latency-bound bookkeeping whose point is that the
// frames never visit the host between the centroid peaks and the centred-instance peaks, except for the one count
// the crop tensor's batch dimension needs.
//
// Arithmetic follows the reference op for op: every fp32 sum / product / quotient is rounded separately.
#include "common.cuh"

namespace snb {

// TopDownLayer._bbox_iou (topdown.py:448-460): two h x w boxes centred on c1, c2; all fp32 tensor ops.
__device__ __forceinline__ float centred_iou(float x1, float y1, float x2, float y2, float half_h, float half_w,
                                             float two_area) {
  const float a_y1 = __fsub_rn(y1, half_h), a_x1 = __fsub_rn(x1, half_w);
  const float a_y2 = __fadd_rn(y1, half_h), a_x2 = __fadd_rn(x1, half_w);
  const float b_y1 = __fsub_rn(y2, half_h), b_x1 = __fsub_rn(x2, half_w);
  const float b_y2 = __fadd_rn(y2, half_h), b_x2 = __fadd_rn(x2, half_w);
  const float ih = fmaxf(__fsub_rn(fminf(a_y2, b_y2), fmaxf(a_y1, b_y1)), 0.f);
  const float iw = fmaxf(__fsub_rn(fminf(a_x2, b_x2), fmaxf(a_x1, b_x1)), 0.f);
  const float inter = __fmul_rn(ih, iw);
  return __fdiv_rn(inter, __fsub_rn(two_area, inter));
}

// torch.argsort(descending=True) order: NaN first, then larger values; equal keys by ascending index.
__device__ __forceinline__ bool td_before(float a, int ia, float b, int ib) {
  const bool an = a != a, bn = b != b;
  if (an || bn) return (an && bn) ? (ia < ib) : an;
  if (a != b) return a > b;
  return ia < ib;
}

constexpr int TD_THREADS = 256;
constexpr int TD_WARPS = TD_THREADS / 32;
enum : unsigned char { TD_INVALID = 0, TD_CAND = 1, TD_KEPT = 2, TD_DROPPED = 3 };

// ONE CTA for the whole batch: warp w owns frames w, w + 8, ...  Phase 1 decides each slot's fate, phase 2 is an
// exclusive scan of the per-frame counts (thread 0; B is a few hundred at most), phase 3 writes the crop list in
// (b, i) order plus the per-slot outputs.  Dynamic shared memory: (B + 1) ints + B * I state bytes.
__global__ void __launch_bounds__(TD_THREADS)
topdown_select_kernel(const float* __restrict__ cen, const float* __restrict__ cen_val, int B, int I,
                      const float* __restrict__ eff, float half_h, float half_w, float two_area, int nms,
                      float nms_thr, int* __restrict__ n_valid, int* __restrict__ frame_off,
                      long long* __restrict__ sample_inds, int* __restrict__ rows, int* __restrict__ row_to_crop,
                      float* __restrict__ crop_bboxes, float* __restrict__ crop_topleft, float* __restrict__ crop_eff,
                      unsigned char* __restrict__ valid_mask, float* __restrict__ cen_img,
                      float* __restrict__ full_bboxes) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  int* s_off = reinterpret_cast<int*>(s_raw);                 // B + 1
  unsigned char* s_state = s_raw + sizeof(int) * (size_t)(B + 1);  // B * I
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  for (int b = warp; b < B; b += TD_WARPS) {
    unsigned char* st = s_state + (size_t)b * I;
    const float* c = cen + (size_t)b * I * 2;
    const float* v = cen_val + (size_t)b * I;
    int n_cand = 0;
    for (int i0 = 0; i0 < I; i0 += 32) {
      const int i = i0 + lane;
      bool ok = false;
      if (i < I) {
        const float x = c[2 * i], y = c[2 * i + 1];
        ok = !(x != x) && !(y != y);  // ~isnan(centroids).any(-1)   (topdown.py:101)
        st[i] = ok ? TD_CAND : TD_INVALID;
      }
      n_cand += __popc(__ballot_sync(FULL, ok));
    }
    __syncwarp();
    int kept = n_cand;
    if (nms && n_cand > 1) {  // frames with <= 1 valid centroid are left alone (topdown.py:428-429)
      kept = 0;
      for (int round = 0; round < n_cand; ++round) {
        // next candidate in argsort(descending) order
        float bv = 0.f;
        int bi = -1;
        for (int i = lane; i < I; i += 32)
          if (st[i] == TD_CAND && (bi < 0 || td_before(v[i], i, bv, bi))) { bv = v[i]; bi = i; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          const float ov = __shfl_xor_sync(FULL, bv, d);
          const int oi = __shfl_xor_sync(FULL, bi, d);
          if (oi >= 0 && (bi < 0 || td_before(ov, oi, bv, bi))) { bv = ov; bi = oi; }
        }
        const float cx = c[2 * bi], cy = c[2 * bi + 1];
        bool hit = false;
        for (int k = lane; k < I; k += 32)
          if (st[k] == TD_KEPT) hit = hit || (centred_iou(cx, cy, c[2 * k], c[2 * k + 1], half_h, half_w, two_area) > nms_thr);
        hit = __any_sync(FULL, hit);
        __syncwarp();
        if (lane == 0) st[bi] = hit ? TD_DROPPED : TD_KEPT;
        kept += hit ? 0 : 1;
        __syncwarp();
      }
    } else {
      for (int i = lane; i < I; i += 32)
        if (st[i] == TD_CAND) st[i] = TD_KEPT;
      __syncwarp();
    }
    if (lane == 0) s_off[b + 1] = kept;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    s_off[0] = 0;
    for (int b = 0; b < B; ++b) {
      acc += s_off[b + 1];
      s_off[b + 1] = acc;
    }
    n_valid[0] = acc;
    if (acc == 0) { n_valid[1] = (int)(2.f * half_h); n_valid[2] = (int)(2.f * half_w); }
  }
  __syncthreads();
  for (int b = threadIdx.x; b <= B; b += TD_THREADS) frame_off[b] = s_off[b];
  for (int b = warp; b < B; b += TD_WARPS) {
    const unsigned char* st = s_state + (size_t)b * I;
    const float e = eff ? eff[b] : 1.f;
    int pos = s_off[b];
    for (int i0 = 0; i0 < I; i0 += 32) {
      const int i = i0 + lane;
      const bool in = i < I;
      const bool keep = in && st[i] == TD_KEPT;
      const unsigned m = __ballot_sync(FULL, keep);
      if (in) {
        const size_t slot = (size_t)b * I + i;
        // predict(): sized = centroids * eff (topdown.py:147); _run_stage_2: image space = sized / eff (:216)
        const float sx = __fmul_rn(cen[2 * slot], e), sy = __fmul_rn(cen[2 * slot + 1], e);
        cen_img[2 * slot] = __fdiv_rn(sx, e);
        cen_img[2 * slot + 1] = __fdiv_rn(sy, e);
        valid_mask[slot] = keep ? 1 : 0;
        float* fb = full_bboxes + 8 * slot;
        if (keep) {
          const int r = pos + __popc(m & ((1u << lane) - 1u));
          // make_centered_bboxes (data/instance_cropping.py:129-171) on the sized-space centroid
          const float xl = __fadd_rn(__fsub_rn(sx, half_w), 0.5f), xr = __fadd_rn(__fadd_rn(sx, half_w), -0.5f);
          const float yt = __fadd_rn(__fsub_rn(sy, half_h), 0.5f), yb = __fadd_rn(__fadd_rn(sy, half_h), -0.5f);
          const float bx[8] = {xl, yt, xr, yt, xr, yb, xl, yb};
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            crop_bboxes[8 * (size_t)r + t] = bx[t];
            fb[t] = __fdiv_rn(bx[t], e);  // bboxes_img = bboxes / per_crop_eff_scale (topdown.py:272)
          }
          if (r == 0) {  // crop_bboxes reads the crop size off bbox 0 (ops/crops.py:66-67): int(|BL.y - TL.y|) + 1 - for a
            // centroid whose +/- half lands in another binade that is one less than the configured size, and the
            // reference then crops (and runs its network on) the smaller window; reproduced, not corrected
            n_valid[1] = (int)fabsf(__fsub_rn(yb, yt)) + 1;
            n_valid[2] = (int)fabsf(__fsub_rn(xr, xl)) + 1;
          }
          crop_topleft[2 * (size_t)r] = xl;
          crop_topleft[2 * (size_t)r + 1] = yt;
          crop_eff[r] = e;
          sample_inds[r] = b;
          rows[r] = (int)slot;
          row_to_crop[slot] = r;
        } else {
#pragma unroll
          for (int t = 0; t < 8; ++t) fb[t] = NAN;
          row_to_crop[slot] = -1;
        }
      }
      pos += __popc(m);
    }
  }
}

// The lift of stage 2 (topdown.py:259-291): for every (b, i) slot, either its crop's peaks
//   full_crop_kpts = stage-2 keypoints,  full_kpts = (stage-2 keypoints + crop top-left) / per-crop eff_scale,
//   full_vals = stage-2 peak values
// or NaN.  One thread per (slot, node).
__global__ void topdown_lift_kernel(const float* __restrict__ kp, const float* __restrict__ val, int n_nodes,
                                    long long n_slots, const int* __restrict__ row_to_crop,
                                    const float* __restrict__ crop_topleft, const float* __restrict__ crop_eff,
                                    float* __restrict__ full_kpts, float* __restrict__ full_crop_kpts,
                                    float* __restrict__ full_vals) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_slots * n_nodes) return;
  const long long slot = t / n_nodes;
  const int node = (int)(t - slot * n_nodes);
  const int r = row_to_crop[slot];
  float x = NAN, y = NAN, cx = NAN, cy = NAN, v = NAN;
  if (r >= 0) {
    const long long o = (long long)r * n_nodes + node;
    cx = kp[2 * o];
    cy = kp[2 * o + 1];
    v = val[o];
    const float e = crop_eff[r];
    x = __fdiv_rn(__fadd_rn(cx, crop_topleft[2 * (long long)r]), e);      // add_crop_offset, then / eff
    y = __fdiv_rn(__fadd_rn(cy, crop_topleft[2 * (long long)r + 1]), e);
  }
  full_kpts[2 * t] = x;
  full_kpts[2 * t + 1] = y;
  if (full_crop_kpts) {
    full_crop_kpts[2 * t] = cx;
    full_crop_kpts[2 * t + 1] = cy;
  }
  full_vals[t] = v;
}

}  // namespace snb

using namespace snb;

extern "C" long long snb_topdown_select_smem_bytes(int B, int I) {
  return (long long)sizeof(int) * (B + 1) + (long long)B * I;
}

extern "C" int snb_topdown_select(const float* centroids, const float* centroid_vals, int B, int I,
                                  const float* eff_scale, int crop_h, int crop_w, int centroid_nms,
                                  float nms_threshold, int* n_valid, int* frame_off, long long* sample_inds, int* rows,
                                  int* row_to_crop, float* crop_bboxes, float* crop_topleft, float* crop_eff,
                                  unsigned char* valid_mask, float* centroids_img, float* full_bboxes, void* stream_) {
  if (B < 0 || I < 0 || crop_h <= 0 || crop_w <= 0) return SNB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream_;
  if (B == 0 || I == 0) {
    cudaMemsetAsync(n_valid, 0, 3 * sizeof(int), st);
    if (frame_off) cudaMemsetAsync(frame_off, 0, sizeof(int) * (size_t)(B + 1), st);
    return SNB_OK;
  }
  const size_t smem = (size_t)snb_topdown_select_smem_bytes(B, I);
  if (smem > 200 * 1024) return SNB_ERR_UNSUPPORTED;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(topdown_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return SNB_ERR_CUDA_LAUNCH;
  // h / 2.0 and 2.0 * (h * w) are python floats in the reference; both are exact in fp32 for any real crop size
  topdown_select_kernel<<<1, TD_THREADS, smem, st>>>(centroids, centroid_vals, B, I, eff_scale, 0.5f * (float)crop_h,
                                                     0.5f * (float)crop_w, 2.0f * (float)(crop_h * crop_w),
                                                     centroid_nms, nms_threshold, n_valid, frame_off, sample_inds, rows,
                                                     row_to_crop, crop_bboxes, crop_topleft, crop_eff, valid_mask,
                                                     centroids_img, full_bboxes);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_topdown_lift(const float* kpts, const float* vals, int n_nodes, long long n_slots,
                                const int* row_to_crop, const float* crop_topleft, const float* crop_eff,
                                float* full_kpts, float* full_crop_kpts, float* full_vals, void* stream_) {
  if (n_nodes < 0 || n_slots < 0) return SNB_ERR_BAD_ARG;
  const long long n = n_slots * n_nodes;
  if (n == 0) return SNB_OK;
  topdown_lift_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      kpts, vals, n_nodes, n_slots, row_to_crop, crop_topleft, crop_eff, full_kpts, full_crop_kpts, full_vals);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
