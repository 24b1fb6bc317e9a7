// Multi-class identity grouping on device and class-map training targets.
// Replaces sleap_nn/inference/ops/identity.py (cited ops/identity.py:NN: group_class_peaks :13-71,
// classify_peaks_from_maps :74-149, get_class_inds_from_vectors :152-173) and sleap_nn/data/identity.py
// (data/identity.py:NN: make_class_vectors :10-32, make_class_maps :35-82).
//
// The grouping side is the second and third user of the device LSAP (paf_device.cuh, scipy semantics): one warp
// per (sample, channel) group gathers the group's peaks, reads their class probabilities under the rounded peak
// position, solves the assignment on cost = -(double)prob, keeps a match only when the assigned class is also
// the peak's arg-max class, and scatters straight into the NaN-filled (S, K, C) outputs.  O(#peaks) data:
// latency-bound, not bandwidth-bound.  The class-map kernel is a per-pixel stream (HBM-bound: reads I planes,
// writes K planes once).
#include "paf_device.cuh"

namespace snb {

constexpr int ID_MAX_DIM = 128;  // largest group / class count one warp solves out of shared memory
constexpr int ID_WS_BYTES = ((ID_MAX_DIM * (3 * 8 + 4 * 4 + 2) + 16) + 15) & ~15;  // == lsap_ws_bytes(ID_MAX_DIM)

// torch.round (half to even) in fp32 -> .to(int32) -> clamp(0, hi)   (ops/identity.py:106-114).  On the x86 hosts
// the reference runs on, NaN / inf / |r| >= 2^31 convert to INT_MIN, which the clamp sends to 0.
__device__ __forceinline__ int class_map_sub(float v, int hi) {
  const float r = rintf(v);
  if (!(fabsf(r) < 2147483648.f)) return 0;
  return min(max((int)r, 0), hi);
}

__device__ __forceinline__ float max_nan_propagating(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

// One warp per (sample, channel) group; grid = n_samples * n_channels.
//   class_maps != NULL: probs (P, K) is an OUTPUT (gathered here); else it is the input of group_class_peaks.
//   g_peak / g_class / g_count (optional): per-group matches after the arg-max filter, padded to K per group,
//     in assignment (row-ascending) order -> snb_pack_class_matches concatenates them in group order.
//   o_xy / o_val / o_prob (optional): the fixed-size outputs of classify_peaks_from_maps; the group owns the
//     slots [s, :, c] and NaN-fills them before writing its matches.
__global__ void __launch_bounds__(32)
classify_peaks_kernel(const float* __restrict__ class_maps, int K, int H, int W, long long ms, long long mk,
                      long long mh, long long mw, const float* __restrict__ peak_xy,
                      const float* __restrict__ peak_val, const int* __restrict__ sample, const int* __restrict__ chan,
                      long long P, const int* __restrict__ frame_count, int cap, float xy_div, int n_channels,
                      float* __restrict__ probs, long long* __restrict__ g_peak,
                      long long* __restrict__ g_class, int* __restrict__ g_count, float* __restrict__ o_xy,
                      float* __restrict__ o_val, float* __restrict__ o_prob, int* __restrict__ status) {
  __shared__ int members[ID_MAX_DIM];
  __shared__ int rows[ID_MAX_DIM], cols[ID_MAX_DIM];
  __shared__ __align__(16) unsigned char ws[ID_WS_BYTES];
  const int lane = threadIdx.x;
  const int s = blockIdx.x / n_channels, c = blockIdx.x - s * n_channels;
  if (o_xy) {
    for (int k = lane; k < K; k += 32) {
      const long long o = ((long long)s * K + k) * n_channels + c;
      o_xy[2 * o] = NAN;
      o_xy[2 * o + 1] = NAN;
      o_val[o] = NAN;
      o_prob[o] = NAN;
    }
  }
  if (g_count && lane == 0) g_count[blockIdx.x] = 0;
  // members of the group in ascending peak index (torch.nonzero(mask) order, ops/identity.py:42-46).  Padded
  // tables (frame_count != NULL: frame s owns slots [s*cap, s*cap + min(count, cap)), the layout K1 writes) only
  // scan their own frame.
  int n = 0;
  bool too_many = false;
  const long long scan_lo = frame_count ? (long long)s * cap : 0;
  const long long scan_hi = frame_count ? scan_lo + min(frame_count[s], cap) : P;
  for (long long base = scan_lo; base < scan_hi; base += 32) {
    const long long i = base + lane;
    const bool hit = i < scan_hi && (frame_count ? true : sample[i] == s) && chan[i] == c;
    const unsigned m = __ballot_sync(FULL, hit);
    if (hit) {
      const int slot = n + __popc(m & ((1u << lane) - 1));
      if (slot < ID_MAX_DIM) members[slot] = (int)i;
    }
    n += __popc(m);
  }
  if (n > ID_MAX_DIM || K > ID_MAX_DIM) too_many = true;
  if (n == 0 || K == 0) return;
  if (too_many) {
    if (lane == 0) atomicOr(status, SNB_STATUS_LSAP_TOO_LARGE);
    return;
  }
  __syncwarp();
  if (class_maps) {  // peak_class_probs = class_maps[sample, :, round(y), round(x)]   (ops/identity.py:103-116)
    for (int t = lane; t < n * K; t += 32) {
      const int r = t / K, k = t - r * K;
      const int p = members[r];
      // xy_div: `peaks / class_maps_output_stride` of the multi-class layer (layers/bottomup_multiclass.py:86-87)
      const int ry = class_map_sub(__fdiv_rn(peak_xy[2 * p + 1], xy_div), H - 1);
      const int rx = class_map_sub(__fdiv_rn(peak_xy[2 * p], xy_div), W - 1);
      probs[(long long)p * K + k] = __ldg(class_maps + (long long)s * ms + (long long)k * mk + (long long)ry * mh + (long long)rx * mw);
    }
    __syncwarp();
  }
  // scipy rejects NaN and -inf cost entries ("matrix contains invalid numeric entries")
  bool bad = false;
  for (int t = lane; t < n * K; t += 32) {
    const float v = probs[(long long)members[t / K] * K + (t % K)];
    bad = bad || (v != v) || (v == INFINITY);
  }
  if (__any_sync(FULL, bad)) {
    if (lane == 0) atomicOr(status, SNB_STATUS_LSAP_INVALID);
    return;
  }
  auto cost = [&](int i, int j) -> double { return -(double)probs[(long long)members[i] * K + j]; };
  bool ok;
  if (n <= 32 && K <= 32) {
    ok = lsap_solve_warp(n, K, cost, ws, rows, cols, lane);
  } else {
    int ok_i = 1;
    if (lane == 0) ok_i = lsap_solve(n, K, cost, ws, rows, cols) ? 1 : 0;
    ok = __shfl_sync(FULL, ok_i, 0) != 0;
    __syncwarp();
  }
  if (!ok) {
    if (lane == 0) atomicOr(status, SNB_STATUS_LSAP_INFEASIBLE);
    return;
  }
  // keep a match only where the assigned class is the peak's best class (ops/identity.py:66-71)
  const int n_match = min(n, K);
  int kept = 0;
  for (int base = 0; base < n_match; base += 32) {
    const int t = base + lane;
    bool keep = false;
    int p = 0, k = 0;
    float pr = 0.f;
    if (t < n_match) {
      p = members[rows[t]];
      k = cols[t];
      const float* row = probs + (long long)p * K;
      pr = row[k];
      float best = row[0];
      for (int j = 1; j < K; ++j) best = max_nan_propagating(best, row[j]);
      keep = (pr == best);
    }
    const unsigned m = __ballot_sync(FULL, keep);
    if (keep) {
      if (g_peak) {
        const long long slot = (long long)blockIdx.x * K + kept + __popc(m & ((1u << lane) - 1));
        g_peak[slot] = p;
        g_class[slot] = k;
      }
      if (o_xy) {
        const long long o = ((long long)s * K + k) * n_channels + c;
        o_xy[2 * o] = __fdiv_rn(peak_xy[2 * p], xy_div);
        o_xy[2 * o + 1] = __fdiv_rn(peak_xy[2 * p + 1], xy_div);
        o_val[o] = peak_val[p];
        o_prob[o] = pr;
      }
    }
    kept += __popc(m);
  }
  if (g_count && lane == 0) g_count[blockIdx.x] = kept;
}

// Per-group padded matches -> the concatenated (peak_inds, class_inds) of group_class_peaks, groups in
// (sample, channel) order (ops/identity.py:58-71).  total[0] receives the number of matches.
__global__ void pack_class_matches_kernel(const long long* __restrict__ g_peak, const long long* __restrict__ g_class,
                                          const int* __restrict__ g_count, int n_groups, int K,
                                          long long* __restrict__ o_peak, long long* __restrict__ o_class,
                                          int* __restrict__ total) {
  const int g = blockIdx.x;
  __shared__ int s_off;
  if (threadIdx.x < 32) {
    int acc = 0;
    for (int i = threadIdx.x; i < g; i += 32) acc += g_count[i];
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(FULL, acc, d);
    if (threadIdx.x == 0) {
      s_off = acc;
      if (g == n_groups - 1) total[0] = acc + g_count[g];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < g_count[g]; i += blockDim.x) {
    o_peak[s_off + i] = g_peak[(long long)g * K + i];
    o_class[s_off + i] = g_class[(long long)g * K + i];
  }
}

// get_class_inds_from_vectors (ops/identity.py:152-173) for ONE group of n rows x K classes, by one warp: the optimal
// assignment on cost = -(double)prob; emit(row, class, prob) is called once per matched row.  Returns false (and
// raises a status bit) on NaN / +inf probabilities or an infeasible matrix, like scipy.
template <typename Emit>
__device__ __forceinline__ bool class_inds_group(const float* __restrict__ probs, int n, int K, void* ws, int* rows,
                                                 int* cols, int lane, int* __restrict__ status, Emit emit) {
  if (n == 0 || K == 0) return true;
  bool bad = false;
  for (long long t = lane; t < (long long)n * K; t += 32) {
    const float v = probs[t];
    bad = bad || (v != v) || (v == INFINITY);
  }
  if (__any_sync(FULL, bad)) {
    if (lane == 0) atomicOr(status, SNB_STATUS_LSAP_INVALID);
    return false;
  }
  __syncwarp();
  auto cost = [&](int i, int j) -> double { return -(double)probs[(long long)i * K + j]; };
  bool ok;
  if (n <= 32 && K <= 32) {
    ok = lsap_solve_warp(n, K, cost, ws, rows, cols, lane);
  } else {
    int ok_i = 1;
    if (lane == 0) {
      ok_i = lsap_solve(n, K, cost, ws, rows, cols) ? 1 : 0;
      __threadfence_block();
    }
    ok = __shfl_sync(FULL, ok_i, 0) != 0;
    __syncwarp();
  }
  if (!ok) {
    if (lane == 0) atomicOr(status, SNB_STATUS_LSAP_INFEASIBLE);
    return false;
  }
  const int n_match = min(n, K);
  for (int t = lane; t < n_match; t += 32) {
    const int r = rows[t], k = cols[t];
    emit(r, k, probs[(long long)r * K + k]);
  }
  return true;
}

// ONE assignment over (n, K); one warp.
__global__ void __launch_bounds__(32)
class_inds_from_vectors_kernel(const float* __restrict__ probs, int n, int K, void* __restrict__ ws, int* __restrict__ rows,
                               int* __restrict__ cols, long long* __restrict__ o_inds, float* __restrict__ o_probs,
                               int* __restrict__ status) {
  const int lane = threadIdx.x;
  for (int i = lane; i < n; i += 32) {
    o_inds[i] = -1;
    o_probs[i] = NAN;
  }
  __syncwarp();
  class_inds_group(probs, n, K, ws, rows, cols, lane, status, [&](int r, int k, float p) {
    o_inds[r] = k;
    o_probs[r] = p;
  });
}

// TopDownLayer._run_stage_2's per-frame re-assignment (layers/topdown.py:343-371): the crops of a batch arrive
// flattened, but a class may be claimed once PER FRAME, so frame b's crops [frame_off[b], frame_off[b+1]) form one
// group.  One warp per frame: NaN / -1 fill of the frame's (max_inst, ...) slots, the assignment, and the scatter of
//   pred_class_inds[b, i, :] = class,  instance_tracking_scores[b, i] = prob,  pred_class_vectors[b, i, :] = vector
// through rows[r] = b * max_inst + i.
__global__ void __launch_bounds__(32)
class_inds_grouped_kernel(const float* __restrict__ probs, int K, const int* __restrict__ frame_off,
                          const int* __restrict__ rows_of_crop, int I, int n_nodes, unsigned char* __restrict__ ws_all,
                          long long ws_stride, int max_dim, long long* __restrict__ full_class_inds,
                          float* __restrict__ full_tracking, float* __restrict__ full_vectors, int* __restrict__ status) {
  const int b = blockIdx.x, lane = threadIdx.x;
  for (long long t = lane; t < (long long)I * n_nodes; t += 32) full_class_inds[(long long)b * I * n_nodes + t] = -1;
  for (int t = lane; t < I; t += 32) full_tracking[(long long)b * I + t] = NAN;
  if (full_vectors)
    for (long long t = lane; t < (long long)I * K; t += 32) full_vectors[(long long)b * I * K + t] = NAN;
  __syncwarp();
  const int r0 = frame_off[b], n = frame_off[b + 1] - r0;
  if (n <= 0) return;
  if (full_vectors)
    for (long long t = lane; t < (long long)n * K; t += 32)
      full_vectors[(long long)rows_of_crop[r0 + t / K] * K + t % K] = probs[(long long)r0 * K + t];
  unsigned char* ws = ws_all + (long long)b * ws_stride;
  int* rows = reinterpret_cast<int*>(ws + lsap_ws_bytes(max_dim));
  class_inds_group(probs + (long long)r0 * K, n, K, ws, rows, rows + max_dim, lane, status, [&](int r, int k, float p) {
    const long long slot = rows_of_crop[r0 + r];
    for (int c = 0; c < n_nodes; ++c) full_class_inds[slot * n_nodes + c] = k;
    full_tracking[slot] = p;
  });
}

// make_class_vectors (data/identity.py:10-32): one-hot int32 rows; index < 0 -> zeros.  Indices arrive as fp32 or
// int32 (the reference's own test passes a float tensor and relies on `.long()` truncation).
__global__ void class_vectors_kernel(const void* __restrict__ class_inds, int is_float, int n, int K, int* __restrict__ out,
                                     int* __restrict__ status) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * K) return;
  const int i = t / K, k = t - i * K;
  int idx;
  bool valid;
  if (is_float) {
    const float f = ((const float*)class_inds)[i];
    valid = f >= 0.f;
    idx = valid ? (int)f : 0;  // .long() truncates toward zero
  } else {
    idx = ((const int*)class_inds)[i];
    valid = idx >= 0;
  }
  if (valid && idx >= K) {  // F.one_hot raises "Class values must be smaller than num_classes"
    atomicOr(status, SNB_STATUS_BAD_INDEX);
    valid = false;
  }
  out[t] = (valid && idx == k) ? 1 : 0;
}

// make_class_maps (data/identity.py:35-82) for G frames.  cms (G, I, h, w) contiguous per-instance confidence maps,
// class_inds (G, I) int32 (-1 = no class), n_valid (G) or NULL = the datasets' per-frame num_instances (instances
// beyond it do not exist for that frame: the reference slices them away before building the maps).
//   total = sum_i cm_i (in instance order) ; share_i = cm_i > thr ? cm_i / total : 0 ;
//   out[c] = max_i (share_i * w[c][i]) with torch.max's NaN propagation,
// where w is the (Ig, K) one-hot matrix REINTERPRETED as (K, Ig) exactly like the reference's reshape: class c /
// instance i reads flat element f = c*Ig + i, i.e. row f / K, column f % K of the one-hot matrix.
// One thread per pixel, grid = (pixel blocks, G); the Ig shares of a pixel are staged in shared memory.
__global__ void __launch_bounds__(128)
class_maps_kernel(const float* __restrict__ cms, const int* __restrict__ class_inds, const int* __restrict__ n_valid,
                  int I, int K, long long hw, float thr, float* __restrict__ out) {
  extern __shared__ float s_dyn[];
  float* s_share = s_dyn;            // I x 128
  float* s_w = s_dyn + (size_t)I * 128;  // K x Ig
  const int tid = threadIdx.x, g = blockIdx.y;
  const int Ig = n_valid ? min(max(n_valid[g], 0), I) : I;
  const float* fcms = cms + (long long)g * I * hw;
  float* fout = out + (long long)g * K * hw;
  for (int f = tid; f < K * Ig; f += blockDim.x) {
    const int r = f / K, k = f - r * K;
    s_w[f] = (class_inds[(long long)g * I + r] == k) ? 1.f : 0.f;
  }
  __syncthreads();
  for (long long px = (long long)blockIdx.x * blockDim.x + tid; px < hw; px += (long long)gridDim.x * blockDim.x) {
    if (Ig == 0) {  // torch.max over an empty instance axis raises in the reference; emit zeros
      for (int c = 0; c < K; ++c) fout[(long long)c * hw + px] = 0.f;
      continue;
    }
    float total = __ldg(fcms + px);
    for (int i = 1; i < Ig; ++i) total = __fadd_rn(total, __ldg(fcms + (long long)i * hw + px));
    for (int i = 0; i < Ig; ++i) {
      const float v = __ldg(fcms + (long long)i * hw + px);
      s_share[i * 128 + tid] = (v > thr) ? __fdiv_rn(v, total) : 0.f;
    }
    for (int c = 0; c < K; ++c) {
      float acc = __fmul_rn(s_share[tid], s_w[c * Ig]);
      for (int i = 1; i < Ig; ++i) acc = max_nan_propagating(acc, __fmul_rn(s_share[i * 128 + tid], s_w[c * Ig + i]));
      fout[(long long)c * hw + px] = acc;
    }
  }
}

// BottomUpMultiClassLayer.postprocess after classify_peaks_from_maps (layers/bottomup_multiclass.py:99-146), one
// warp per frame: instances * class_maps_output_stride, / input_scale, / eff_scale[b] (each a separately rounded fp32
// op; multiplying / dividing by exactly 1 is the identity, so the reference's `!= 1.0` short-circuits need no
// branch), instance score = nanmean of the peak values over nodes, tracking score = nanmean of the class
// probabilities, then _cap_instances_by_score (:148-190): when more than max_instances classes are present, keep
// the first max_instances of np.argsort(scores)[::-1] - NaN scores FIRST (absent classes do take slots; a quirk
// kept from the reference), then descending, equal scores by descending index - and NaN-out the rest.
__global__ void __launch_bounds__(32)
multiclass_outputs_kernel(const float* __restrict__ xy, const float* __restrict__ val, const float* __restrict__ prob,
                          int K, int N, float class_stride, float input_scale, const float* __restrict__ eff_scale,
                          int max_instances, float* __restrict__ o_kpts, float* __restrict__ o_vals,
                          float* __restrict__ o_scores, float* __restrict__ o_tracking) {
  extern __shared__ float s_score[];                               // K
  unsigned char* s_drop = reinterpret_cast<unsigned char*>(s_score + K);  // K
  const int b = blockIdx.x, lane = threadIdx.x;
  const float eff = eff_scale ? eff_scale[b] : 1.f;
  for (int k = lane; k < K; k += 32) {
    const long long base = ((long long)b * K + k) * N;
    float sv = 0.f, sp = 0.f;
    int nv = 0, np_ = 0;
    for (int n = 0; n < N; ++n) {
      const float v = val[base + n], p = prob[base + n];
      if (v == v) { sv = __fadd_rn(sv, v); ++nv; }
      if (p == p) { sp = __fadd_rn(sp, p); ++np_; }
    }
    s_score[k] = __fdiv_rn(sv, (float)nv);  // 0/0 = NaN for a class without peaks, like torch.nanmean
    o_tracking[(long long)b * K + k] = __fdiv_rn(sp, (float)np_);
    s_drop[k] = 0;
  }
  __syncwarp();
  if (max_instances >= 0) {
    int present = 0;
    for (int k = lane; k < K; k += 32) present += (s_score[k] == s_score[k]) ? 1 : 0;
    for (int d = 16; d > 0; d >>= 1) present += __shfl_xor_sync(FULL, present, d);
    if (present > max_instances) {
      for (int k = lane; k < K; k += 32) {  // rank of class k in np.argsort(scores)[::-1]
        const float sk = s_score[k];
        const bool kn = sk != sk;
        int rank = 0;
        for (int j = 0; j < K; ++j) {
          if (j == k) continue;
          const float sj = s_score[j];
          const bool jn = sj != sj;
          bool before;
          if (kn || jn) before = (kn && jn) ? (j > k) : jn;
          else before = (sj > sk) || (sj == sk && j > k);
          rank += before ? 1 : 0;
        }
        s_drop[k] = rank >= max_instances ? 1 : 0;
      }
    }
    __syncwarp();
  }
  for (int t = lane; t < K * N; t += 32) {
    const int k = t / N;
    const long long src = (long long)b * K * N + t;
    const bool drop = s_drop[k] != 0;
    float x = __fdiv_rn(__fdiv_rn(__fmul_rn(xy[2 * src], class_stride), input_scale), eff);
    float y = __fdiv_rn(__fdiv_rn(__fmul_rn(xy[2 * src + 1], class_stride), input_scale), eff);
    o_kpts[2 * src] = drop ? NAN : x;
    o_kpts[2 * src + 1] = drop ? NAN : y;
    o_vals[src] = drop ? NAN : val[src];
  }
  for (int k = lane; k < K; k += 32) {
    const long long o = (long long)b * K + k;
    o_scores[o] = s_drop[k] ? NAN : s_score[k];
    if (s_drop[k]) o_tracking[o] = NAN;
  }
}

}  // namespace snb

using namespace snb;

extern "C" int snb_classify_peaks(const float* class_maps, int n_samples, int K, int H, int W, long long ms,
                                  long long mk, long long mh, long long mw, const float* peak_xy, const float* peak_val,
                                  const int* sample_inds, const int* channel_inds, long long P, int n_channels,
                                  float* probs, long long* g_peak, long long* g_class, int* g_count, float* o_xy,
                                  float* o_val, float* o_prob, int* status, void* stream) {
  if (n_samples < 0 || n_channels < 0 || K < 0 || P < 0 || !status) return SNB_ERR_BAD_ARG;
  if (P > 0 && (!probs || !sample_inds || !channel_inds)) return SNB_ERR_BAD_ARG;
  if (class_maps && P > 0 && !peak_xy) return SNB_ERR_BAD_ARG;
  if (o_xy && (!o_val || !o_prob || (P > 0 && (!peak_xy || !peak_val)))) return SNB_ERR_BAD_ARG;
  if (g_peak && (!g_class || !g_count)) return SNB_ERR_BAD_ARG;
  const long long groups = (long long)n_samples * n_channels;
  if (groups == 0) return SNB_OK;
  if (groups > 0x7fffffffLL) return SNB_ERR_UNSUPPORTED;
  classify_peaks_kernel<<<(unsigned)groups, 32, 0, (cudaStream_t)stream>>>(
      class_maps, K, H, W, ms, mk, mh, mw, peak_xy, peak_val, sample_inds, channel_inds, P, nullptr, 0, 1.0f, n_channels,
      probs, g_peak, g_class, g_count, o_xy, o_val, o_prob, status);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_classify_peaks_padded(const float* class_maps, int n_samples, int K, int H, int W, long long ms,
                                         long long mk, long long mh, long long mw, const int* frame_count, int cap,
                                         const float* peak_xy, const float* peak_val, const int* peak_chan,
                                         float xy_div, int n_channels, float* probs, float* o_xy, float* o_val,
                                         float* o_prob, int* status, void* stream) {
  if (n_samples < 0 || n_channels < 0 || K < 0 || cap <= 0 || !status || !class_maps || !frame_count || !peak_xy ||
      !peak_val || !peak_chan || !probs || !o_xy || !o_val || !o_prob)
    return SNB_ERR_BAD_ARG;
  const long long groups = (long long)n_samples * n_channels;
  if (groups == 0) return SNB_OK;
  if (groups > 0x7fffffffLL) return SNB_ERR_UNSUPPORTED;
  classify_peaks_kernel<<<(unsigned)groups, 32, 0, (cudaStream_t)stream>>>(
      class_maps, K, H, W, ms, mk, mh, mw, peak_xy, peak_val, nullptr, peak_chan, (long long)n_samples * cap, frame_count,
      cap, xy_div, n_channels, probs, nullptr, nullptr, nullptr, o_xy, o_val, o_prob, status);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_multiclass_outputs(const float* xy, const float* val, const float* prob, int B, int K, int N,
                                      float class_stride, float input_scale, const float* eff_scale, int max_instances,
                                      float* o_kpts, float* o_vals, float* o_scores, float* o_tracking, void* stream) {
  if (B < 0 || K < 0 || N < 0 || !xy || !val || !prob || !o_kpts || !o_vals || !o_scores || !o_tracking)
    return SNB_ERR_BAD_ARG;
  if (B == 0 || K == 0) return SNB_OK;
  const size_t smem = (size_t)K * (sizeof(float) + 1);
  if (smem > 48 * 1024) return SNB_ERR_UNSUPPORTED;
  multiclass_outputs_kernel<<<B, 32, smem, (cudaStream_t)stream>>>(xy, val, prob, K, N, class_stride, input_scale,
                                                                  eff_scale, max_instances, o_kpts, o_vals, o_scores,
                                                                  o_tracking);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_pack_class_matches(const long long* g_peak, const long long* g_class, const int* g_count, int n_groups,
                                      int K, long long* o_peak, long long* o_class, int* total, void* stream) {
  if (n_groups < 0 || !total) return SNB_ERR_BAD_ARG;
  if (n_groups == 0) return cudaMemsetAsync(total, 0, sizeof(int), (cudaStream_t)stream) == cudaSuccess ? SNB_OK : SNB_ERR_CUDA_LAUNCH;
  pack_class_matches_kernel<<<n_groups, 64, 0, (cudaStream_t)stream>>>(g_peak, g_class, g_count, n_groups, K, o_peak,
                                                                      o_class, total);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" long long snb_class_inds_workspace_bytes(int n, int K) {
  const int d = n > K ? n : K;
  return (long long)lsap_ws_bytes(d < 1 ? 1 : d) + 2LL * sizeof(int) * (d < 1 ? 1 : d);
}

extern "C" int snb_class_inds_from_vectors(const float* probs, int n, int K, void* workspace, long long* o_inds,
                                           float* o_probs, int* status, void* stream) {
  if (n < 0 || K < 0 || !status) return SNB_ERR_BAD_ARG;
  if (n == 0) return SNB_OK;
  if (!workspace || !o_inds || !o_probs || (K > 0 && !probs)) return SNB_ERR_BAD_ARG;
  const int d = n > K ? n : K;
  int* rows = (int*)((char*)workspace + lsap_ws_bytes(d));
  class_inds_from_vectors_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(probs, n, K, workspace, rows, rows + d, o_inds,
                                                                    o_probs, status);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" long long snb_class_inds_grouped_workspace_bytes(int B, int I, int K) {
  const long long per = (snb_class_inds_workspace_bytes(I, K) + 15) & ~15LL;
  return per * (B < 1 ? 1 : B);
}

extern "C" int snb_class_inds_grouped(const float* probs, int K, const int* frame_off, const int* rows_of_crop, int B,
                                     int I, int n_nodes, void* workspace, long long* full_class_inds,
                                     float* full_tracking, float* full_vectors, int* status, void* stream) {
  if (B < 0 || I < 0 || K < 0 || n_nodes < 0 || !status) return SNB_ERR_BAD_ARG;
  if (B == 0) return SNB_OK;
  if (!workspace || !frame_off || !full_class_inds || !full_tracking) return SNB_ERR_BAD_ARG;
  const int d = (I > K ? I : K) < 1 ? 1 : (I > K ? I : K);
  class_inds_grouped_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(probs, K, frame_off, rows_of_crop, I, n_nodes,
                                                               (unsigned char*)workspace,
                                                               (snb_class_inds_workspace_bytes(I, K) + 15) & ~15LL, d,
                                                               full_class_inds,
                                                               full_tracking, full_vectors, status);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_class_vectors(const void* class_inds, int is_float, int n, int K, int* out, int* status, void* stream) {
  if (n < 0 || K < 0 || !status) return SNB_ERR_BAD_ARG;
  const long long total = (long long)n * K;
  if (total == 0) return SNB_OK;
  if (total > 0x7fffffffLL) return SNB_ERR_UNSUPPORTED;
  class_vectors_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(class_inds, is_float, n, K, out,
                                                                                        status);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_class_maps(const float* confmaps, const int* class_inds, const int* n_valid, int G, int I, int K,
                              int h, int w, float threshold, float* out, void* stream) {
  if (G < 0 || I <= 0 || K < 0 || h < 0 || w < 0) return SNB_ERR_BAD_ARG;
  const long long hw = (long long)h * w;
  if (hw == 0 || K == 0 || G == 0) return SNB_OK;
  if (G > 65535) return SNB_ERR_UNSUPPORTED;
  const size_t smem = ((size_t)I * 128 + (size_t)K * I) * sizeof(float);
  if (smem > 200 * 1024) return SNB_ERR_UNSUPPORTED;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(class_maps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return SNB_ERR_CUDA_LAUNCH;
  long long blocks = (hw + 127) / 128;
  const long long cap = (148LL * 16 + G - 1) / G;
  if (blocks > cap) blocks = cap < 1 ? 1 : cap;
  class_maps_kernel<<<dim3((unsigned)blocks, G), 128, smem, (cudaStream_t)stream>>>(confmaps, class_inds, n_valid, I, K,
                                                                                   hw, threshold, out);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
