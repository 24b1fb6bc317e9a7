// Post-inference instance filters on device: the tensor-level FilterPipeline of
// sleap_nn/inference/filters.py (cited filters.py:NN) - min_peak_value (:165-176), node-count (:178-197), score
// filters (:199-243), greedy overlap NMS by bbox IoU or OKS (:245-344) and centroid-distance NMS (:375-412) - applied
// in the reference's fixed order to the padded (B, I, N, 2) / (B, I, N) / (B, I) outputs the grouping kernels
// leave in HBM, so the frames never visit the host between grouping and packaging.
//
// One warp per frame; everything is O(I * N) or O(I^2) on a few KB: latency-bound.  Arithmetic follows the
// reference op for op: fp32 tensor ops (min / max / sub / mul / div / exp) each rounded separately, `.item()`
// promotions to double where the reference compares python floats (IoU ratio, OKS mean, squared distance).
#include "common.cuh"

namespace snb {

struct FilterFrame {
  float* kpts;   // (I, N, 2) or NULL
  float* vals;   // (I, N) or NULL
  float* scores; // (I) or NULL
  float* cen;    // (I, 2) or NULL
  float* cenv;   // (I) or NULL
  int I, N;
};

// FilterPipeline._nan_out_where (filters.py:346-373) for one instance slot; called by all lanes.
__device__ __forceinline__ void nan_out(const FilterFrame& f, int i, int lane) {
  if (f.kpts)
    for (int t = lane; t < 2 * f.N; t += 32) f.kpts[(long long)i * f.N * 2 + t] = NAN;
  if (f.vals)
    for (int t = lane; t < f.N; t += 32) f.vals[(long long)i * f.N + t] = NAN;
  if (lane == 0) {
    if (f.scores) f.scores[i] = NAN;
    if (f.cen) { f.cen[2 * i] = NAN; f.cen[2 * i + 1] = NAN; }
    if (f.cenv) f.cenv[i] = NAN;
  }
}

// torch.argsort(descending=True) order: NaN first, then larger values, equal keys by ascending index.
__device__ __forceinline__ bool sorts_before(float a, int ia, float b, int ib) {
  const bool an = a != a, bn = b != b;
  if (an || bn) return (an && bn) ? (ia < ib) : an;
  if (a != b) return a > b;
  return ia < ib;
}

struct InstBox {  // bbox of the rows with both coordinates present (filters.py:297-305, :330-337)
  float x1, y1, x2, y2;
  int rows;     // rows with both coordinates non-NaN
  int any;      // any non-NaN coordinate at all (valid_b, filters.py:266)
};

__device__ __forceinline__ double bbox_iou(const InstBox& a, const InstBox& b) {
  if (a.rows == 0 || b.rows == 0) return 0.0;
  const float iw = fmaxf(__fsub_rn(fminf(a.x2, b.x2), fmaxf(a.x1, b.x1)), 0.f);
  const float ih = fmaxf(__fsub_rn(fminf(a.y2, b.y2), fmaxf(a.y1, b.y1)), 0.f);
  const double inter = (double)__fmul_rn(iw, ih);
  const double area_a = (double)__fmul_rn(__fsub_rn(a.x2, a.x1), __fsub_rn(a.y2, a.y1));
  const double area_b = (double)__fmul_rn(__fsub_rn(b.x2, b.x1), __fsub_rn(b.y2, b.y1));
  const double uni = area_a + area_b - inter;
  return uni > 0.0 ? inter / uni : 0.0;
}

// FilterPipeline._oks(a, b) (filters.py:311-344): scale = bbox AREA of a's own valid keypoints.
__device__ __forceinline__ double oks(const float* __restrict__ a, const InstBox& abox, const float* __restrict__ b, int N,
                                      float kappa_sq) {
  if (abox.rows < 2) {
    // (no keypoint visible in both) or (< 2 valid keypoints in a) -> 0.0 either way
    return 0.0;
  }
  const float scale_sq = __fmul_rn(__fsub_rn(abox.x2, abox.x1), __fsub_rn(abox.y2, abox.y1));
  if (!(scale_sq > 0.f)) return 0.0;  // `scale_sq.item() <= 0`; NaN cannot occur (valid rows only)
  const float den = __fmul_rn(__fmul_rn(2.f, scale_sq), kappa_sq);
  float sum = 0.f;
  int cnt = 0;
  for (int n = 0; n < N; ++n) {
    const float ax = a[2 * n], ay = a[2 * n + 1], bx = b[2 * n], by = b[2 * n + 1];
    if (ax != ax || ay != ay || bx != bx || by != by) continue;
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by);
    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    sum = __fadd_rn(sum, expf(__fdiv_rn(-d2, den)));
    ++cnt;
  }
  if (cnt == 0) return 0.0;
  return (double)__fdiv_rn(sum, (float)cnt);
}

struct FilterCfg {
  float min_peak_value, min_visible_node_fraction, min_instance_score, min_mean_node_score, oks_kappa_sq;
  int min_visible_nodes, overlapping;
  double overlapping_threshold, min_centroid_distance_sq;
};

__global__ void __launch_bounds__(32)
filter_instances_kernel(FilterCfg cfg, int I, int N, const float* __restrict__ in_kpts, const float* __restrict__ in_vals,
                        const float* __restrict__ in_scores, const float* __restrict__ in_cen,
                        const float* __restrict__ in_cenv, float* o_kpts, float* o_vals, float* o_scores, float* o_cen,
                        float* o_cenv) {
  extern __shared__ __align__(8) unsigned char s_raw[];
  InstBox* box = reinterpret_cast<InstBox*>(s_raw);          // I
  int* order = reinterpret_cast<int*>(box + I);              // I
  int* kept = order + I;                                     // I
  unsigned char* dropf = reinterpret_cast<unsigned char*>(kept + I);  // I
  const int b = blockIdx.x, lane = threadIdx.x;
  FilterFrame f;
  f.I = I; f.N = N;
  f.kpts = o_kpts ? o_kpts + (long long)b * I * N * 2 : nullptr;
  f.vals = o_vals ? o_vals + (long long)b * I * N : nullptr;
  f.scores = o_scores ? o_scores + (long long)b * I : nullptr;
  f.cen = o_cen ? o_cen + (long long)b * I * 2 : nullptr;
  f.cenv = o_cenv ? o_cenv + (long long)b * I : nullptr;
  // functional: copy the frame, then filter the copy in place
  if (f.kpts) for (int t = lane; t < I * N * 2; t += 32) f.kpts[t] = in_kpts[(long long)b * I * N * 2 + t];
  if (f.vals) for (int t = lane; t < I * N; t += 32) f.vals[t] = in_vals[(long long)b * I * N + t];
  if (f.scores) for (int t = lane; t < I; t += 32) f.scores[t] = in_scores[(long long)b * I + t];
  if (f.cen) for (int t = lane; t < I * 2; t += 32) f.cen[t] = in_cen[(long long)b * I * 2 + t];
  if (f.cenv) for (int t = lane; t < I; t += 32) f.cenv[t] = in_cenv[(long long)b * I + t];
  __syncwarp();

  // 1. min_peak_value (filters.py:165-176): NaN-out single keypoints
  if (cfg.min_peak_value > 0.f && f.kpts && f.vals) {
    for (int t = lane; t < I * N; t += 32) {
      if (f.vals[t] < cfg.min_peak_value) {
        f.kpts[2 * t] = NAN;
        f.kpts[2 * t + 1] = NAN;
        f.vals[t] = NAN;
      }
    }
    __syncwarp();
  }

  // 2. node count (filters.py:178-197)
  if ((cfg.min_visible_nodes > 0 || cfg.min_visible_node_fraction > 0.f) && f.kpts) {
    for (int i0 = 0; i0 < I; i0 += 32) {
      const int i = i0 + lane;
      bool drop = false;
      if (i < I) {
        int nv = 0;
        for (int n = 0; n < N; ++n) {
          const float x = f.kpts[((long long)i * N + n) * 2], y = f.kpts[((long long)i * N + n) * 2 + 1];
          nv += (x == x && y == y) ? 1 : 0;
        }
        bool keep = true;
        if (cfg.min_visible_nodes > 0) keep = keep && nv >= cfg.min_visible_nodes;
        if (cfg.min_visible_node_fraction > 0.f)
          keep = keep && __fdiv_rn((float)nv, (float)max(N, 1)) >= cfg.min_visible_node_fraction;
        drop = !keep;
      }
      unsigned m = __ballot_sync(FULL, drop);
      while (m) { nan_out(f, i0 + __ffs(m) - 1, lane); m &= m - 1; }
    }
    __syncwarp();
  }

  // 3. score filters (filters.py:199-243)
  if (cfg.min_instance_score > 0.f || cfg.min_mean_node_score > 0.f) {
    for (int i0 = 0; i0 < I; i0 += 32) {
      const int i = i0 + lane;
      bool drop = false;
      if (i < I) {
        if (!f.kpts) {  // centroid-only outputs: gate on instance_scores, else on the centroid value; NaN fails
          const float* sc = f.scores ? f.scores : f.cenv;
          if (f.cen && cfg.min_instance_score > 0.f && sc) drop = (sc[i] < cfg.min_instance_score) || (sc[i] != sc[i]);
        } else {
          bool keep = true;
          if (cfg.min_instance_score > 0.f && f.scores) keep = keep && f.scores[i] >= cfg.min_instance_score;
          if (cfg.min_mean_node_score > 0.f && f.vals) {
            float sum = 0.f;
            int cnt = 0;
            for (int n = 0; n < N; ++n) {
              const float v = f.vals[(long long)i * N + n];
              if (v == v) { sum = __fadd_rn(sum, v); ++cnt; }
            }
            float mean = __fdiv_rn(sum, (float)cnt);  // torch.nanmean: 0/0 -> NaN for all-NaN rows ...
            if (mean != mean) mean = 0.f;             // ... which count as failing (filters.py:236-241)
            keep = keep && mean >= cfg.min_mean_node_score;
          }
          drop = !keep;
        }
      }
      unsigned m = __ballot_sync(FULL, drop);
      while (m) { nan_out(f, i0 + __ffs(m) - 1, lane); m &= m - 1; }
    }
    __syncwarp();
  }

  // 4. greedy overlap NMS (filters.py:245-344), highest score first
  if (cfg.overlapping && f.kpts) {
    for (int i = lane; i < I; i += 32) {
      InstBox bx{INFINITY, INFINITY, -INFINITY, -INFINITY, 0, 0};
      for (int n = 0; n < N; ++n) {
        const float x = f.kpts[((long long)i * N + n) * 2], y = f.kpts[((long long)i * N + n) * 2 + 1];
        if (x == x || y == y) bx.any = 1;
        if (x == x && y == y) {
          bx.x1 = fminf(bx.x1, x); bx.y1 = fminf(bx.y1, y);
          bx.x2 = fmaxf(bx.x2, x); bx.y2 = fmaxf(bx.y2, y);
          ++bx.rows;
        }
      }
      box[i] = bx;
      dropf[i] = 0;
    }
    __syncwarp();
    int n_valid = 0;
    for (int i = lane; i < I; i += 32) n_valid += box[i].any;
    for (int d = 16; d > 0; d >>= 1) n_valid += __shfl_xor_sync(FULL, n_valid, d);
    if (n_valid > 1) {
      for (int i = lane; i < I; i += 32) {
        const float si = f.scores ? f.scores[i] : 0.f;
        int rank = 0;
        for (int j = 0; j < I; ++j) {
          const float sj = f.scores ? f.scores[j] : 0.f;
          rank += (j != i && sorts_before(sj, j, si, i)) ? 1 : 0;
        }
        order[rank] = i;
      }
      __syncwarp();
      int nk = 0;
      for (int r = 0; r < I; ++r) {
        const int idx = order[r];
        if (!box[idx].any) continue;
        bool hit = false;
        for (int k = lane; k < nk && !hit; k += 32) {
          const int other = kept[k];
          const double sim = (cfg.overlapping == 2)
                                 ? oks(f.kpts + (long long)idx * N * 2, box[idx], f.kpts + (long long)other * N * 2, N,
                                       cfg.oks_kappa_sq)
                                 : bbox_iou(box[idx], box[other]);
          hit = sim > cfg.overlapping_threshold;
        }
        if (__any_sync(FULL, hit)) {
          if (lane == 0) dropf[idx] = 1;
        } else {
          if (lane == 0) kept[nk] = idx;
          ++nk;
        }
        __syncwarp();
      }
      for (int i = 0; i < I; ++i)
        if (dropf[i]) nan_out(f, i, lane);
    }
    __syncwarp();
  }

  // 5. centroid-distance NMS (filters.py:375-412)
  if (cfg.min_centroid_distance_sq > 0.0 && f.cen) {
    const float* sc = f.cenv ? f.cenv : f.scores;
    int n_valid = 0;
    for (int i = lane; i < I; i += 32) {
      const float x = f.cen[2 * i], y = f.cen[2 * i + 1];
      const int v = (x == x && y == y) ? 1 : 0;
      box[i].any = v;
      dropf[i] = 0;
      n_valid += v;
    }
    for (int d = 16; d > 0; d >>= 1) n_valid += __shfl_xor_sync(FULL, n_valid, d);
    __syncwarp();
    if (n_valid > 1) {
      for (int i = lane; i < I; i += 32) {  // NaN scores sort LAST here (treated as -inf, filters.py:395-397)
        float si = sc ? sc[i] : 0.f;
        if (si != si) si = -INFINITY;
        int rank = 0;
        for (int j = 0; j < I; ++j) {
          float sj = sc ? sc[j] : 0.f;
          if (sj != sj) sj = -INFINITY;
          rank += (j != i && sorts_before(sj, j, si, i)) ? 1 : 0;
        }
        order[rank] = i;
      }
      __syncwarp();
      int nk = 0;
      for (int r = 0; r < I; ++r) {
        const int idx = order[r];
        if (!box[idx].any) continue;
        const float px = f.cen[2 * idx], py = f.cen[2 * idx + 1];
        bool hit = false;
        for (int k = lane; k < nk && !hit; k += 32) {
          const int other = kept[k];
          const float dx = __fsub_rn(px, f.cen[2 * other]), dy = __fsub_rn(py, f.cen[2 * other + 1]);
          hit = (double)__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < cfg.min_centroid_distance_sq;
        }
        if (__any_sync(FULL, hit)) {
          if (lane == 0) dropf[idx] = 1;
        } else {
          if (lane == 0) kept[nk] = idx;
          ++nk;
        }
        __syncwarp();
      }
      for (int i = 0; i < I; ++i)
        if (dropf[i]) nan_out(f, i, lane);
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// The Labels-level filters of sleap_nn/inference/ops/filters.py (cited ops/filters.py:NN) work on float64 numpy
// arrays taken from `instance.numpy()`; these are their numeric cores, for all frames of a Labels object in one
// launch (CSR over frames).  Arithmetic is float64 throughout, like numpy's.
// ------------------------------------------------------------------------------------------------------------------
struct BoxD {
  double x1, y1, x2, y2;
  int rows;
};

// _compute_iou_one_to_many (ops/filters.py:407-436) on _instance_bbox boxes (:300-316; no valid point -> [0,0,0,0])
__device__ __forceinline__ double iou_f64(const BoxD& a, const BoxD& b) {
  const double iw = fmax(0.0, fmin(a.x2, b.x2) - fmax(a.x1, b.x1));
  const double ih = fmax(0.0, fmin(a.y2, b.y2) - fmax(a.y1, b.y1));
  const double inter = iw * ih;
  const double uni = (a.x2 - a.x1) * (a.y2 - a.y1) + (b.x2 - b.x1) * (b.y2 - b.y1) - inter;
  return uni > 0.0 ? inter / uni : 0.0;
}

// _compute_oks(points_a, points_b) (ops/filters.py:439-495): a is the instance that was just KEPT, scale = its bbox area
__device__ __forceinline__ double oks_f64(const double* __restrict__ a, const BoxD& abox, const double* __restrict__ b,
                                          int N, double kappa) {
  double sum = 0.0;
  int cnt = 0;
  if (abox.rows < 2) return 0.0;
  const double scale_sq = (abox.x2 - abox.x1) * (abox.y2 - abox.y1);
  if (scale_sq <= 0.0) return 0.0;
  const double den = 2.0 * scale_sq * (kappa * kappa);
  for (int n = 0; n < N; ++n) {
    const double ax = a[2 * n], ay = a[2 * n + 1], bx = b[2 * n], by = b[2 * n + 1];
    if (ax != ax || ay != ay || bx != bx || by != by) continue;
    const double dx = ax - bx, dy = ay - by;
    sum += exp(-(dx * dx + dy * dy) / den);
    ++cnt;
  }
  return cnt ? sum / (double)cnt : 0.0;
}

// np.argsort(scores)[::-1] (ops/filters.py:349, :387): ascending with NaN last and equal keys in index order,
// reversed -> NaN first, then descending, equal keys by DESCENDING index.
__device__ __forceinline__ bool argsort_reversed_before(double a, int ia, double b, int ib) {
  const bool an = a != a, bn = b != b;
  if (an || bn) return (an && bn) ? (ia > ib) : an;
  if (a != b) return a > b;
  return ia > ib;
}

// One warp per frame: _nms_greedy_iou / _nms_greedy_oks (ops/filters.py:330-404).  keep[start .. start+count) = the kept
// LOCAL indices in keep order (decreasing score), the rest of the frame's slots -1.
__global__ void __launch_bounds__(32)
nms_greedy_f64_kernel(const double* __restrict__ pts, const double* __restrict__ scores, const int* __restrict__ frame_start,
                      int N, int method, double threshold, double kappa, int* __restrict__ keep,
                      int* __restrict__ keep_count) {
  extern __shared__ __align__(8) unsigned char s_raw[];
  const int f = blockIdx.x, lane = threadIdx.x;
  const int s0 = frame_start[f], n = frame_start[f + 1] - s0;
  BoxD* box = reinterpret_cast<BoxD*>(s_raw);  // n
  int* order = reinterpret_cast<int*>(box + n);
  int* kept = order + n;
  for (int i = lane; i < n; i += 32) {
    BoxD bx{0.0, 0.0, 0.0, 0.0, 0};
    const double* p = pts + (long long)(s0 + i) * N * 2;
    for (int k = 0; k < N; ++k) {
      const double x = p[2 * k], y = p[2 * k + 1];
      if (x == x && y == y) {
        if (bx.rows == 0) { bx.x1 = bx.x2 = x; bx.y1 = bx.y2 = y; }
        else { bx.x1 = fmin(bx.x1, x); bx.x2 = fmax(bx.x2, x); bx.y1 = fmin(bx.y1, y); bx.y2 = fmax(bx.y2, y); }
        ++bx.rows;
      }
    }
    box[i] = bx;
    keep[s0 + i] = -1;
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (j != i && argsort_reversed_before(scores[s0 + j], j, scores[s0 + i], i)) ? 1 : 0;
    order[rank] = i;
  }
  __syncwarp();
  int nk = 0;
  for (int r = 0; r < n; ++r) {
    const int idx = order[r];
    bool hit = false;
    for (int k = lane; k < nk && !hit; k += 32) {
      const int a = kept[k];
      const double sim = method == 1 ? oks_f64(pts + (long long)(s0 + a) * N * 2, box[a], pts + (long long)(s0 + idx) * N * 2, N, kappa)
                                     : iou_f64(box[a], box[idx]);
      hit = sim > threshold;  // survivors are the ones with similarity <= threshold
    }
    if (!__any_sync(FULL, hit)) {
      if (lane == 0) { kept[nk] = idx; keep[s0 + nk] = idx; }
      ++nk;
    }
    __syncwarp();
  }
  if (lane == 0) keep_count[f] = nk;
}

// _count_visible_nodes (ops/filters.py:178-190) and _mean_node_score (:193-226) for every instance: one thread each.
__global__ void instance_stats_f64_kernel(const double* __restrict__ pts, const double* __restrict__ point_scores,
                                          long long total, int N, int* __restrict__ n_visible,
                                          double* __restrict__ mean_score) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int nv = 0, ns = 0;
  double sum = 0.0;
  for (int k = 0; k < N; ++k) {
    const double x = pts[(i * N + k) * 2], y = pts[(i * N + k) * 2 + 1];
    if (x == x && y == y) {
      ++nv;
      if (point_scores) {
        const double v = point_scores[i * N + k];
        if (v == v) { sum += v; ++ns; }
      }
    }
  }
  n_visible[i] = nv;
  if (mean_score) mean_score[i] = ns ? sum / (double)ns : 0.0;  // no visible node / no finite score -> 0.0
}

}  // namespace snb

using namespace snb;

// FilterPipeline._bbox_iou / _oks (filters.py:290-338) for ONE pair of keypoint sets, in the dtype the caller holds
// (the reference computes in the tensors' dtype: fp32 inside the pipeline, float64 in its own unit tests).
// out[0] = IoU, out[1] = OKS(a, b) with the scale taken from a.  A single thread: N is a handful of nodes.
__global__ void pair_similarity_kernel(const void* __restrict__ a_, const void* __restrict__ b_, int N, int is_f64,
                                       double kappa, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (is_f64) {
    const double* a = (const double*)a_;
    const double* b = (const double*)b_;
    BoxD ba{INFINITY, INFINITY, -INFINITY, -INFINITY, 0}, bb = ba;
    for (int n = 0; n < N; ++n) {
      if (a[2 * n] == a[2 * n] && a[2 * n + 1] == a[2 * n + 1]) {
        ba.x1 = fmin(ba.x1, a[2 * n]); ba.y1 = fmin(ba.y1, a[2 * n + 1]);
        ba.x2 = fmax(ba.x2, a[2 * n]); ba.y2 = fmax(ba.y2, a[2 * n + 1]);
        ++ba.rows;
      }
      if (b[2 * n] == b[2 * n] && b[2 * n + 1] == b[2 * n + 1]) {
        bb.x1 = fmin(bb.x1, b[2 * n]); bb.y1 = fmin(bb.y1, b[2 * n + 1]);
        bb.x2 = fmax(bb.x2, b[2 * n]); bb.y2 = fmax(bb.y2, b[2 * n + 1]);
        ++bb.rows;
      }
    }
    out[0] = (ba.rows == 0 || bb.rows == 0) ? 0.0 : iou_f64(ba, bb);
    out[1] = oks_f64(a, ba, b, N, kappa);
  } else {
    const float* a = (const float*)a_;
    const float* b = (const float*)b_;
    InstBox ba{INFINITY, INFINITY, -INFINITY, -INFINITY, 0, 0}, bb = ba;
    for (int n = 0; n < N; ++n) {
      if (a[2 * n] == a[2 * n] && a[2 * n + 1] == a[2 * n + 1]) {
        ba.x1 = fminf(ba.x1, a[2 * n]); ba.y1 = fminf(ba.y1, a[2 * n + 1]);
        ba.x2 = fmaxf(ba.x2, a[2 * n]); ba.y2 = fmaxf(ba.y2, a[2 * n + 1]);
        ++ba.rows;
      }
      if (b[2 * n] == b[2 * n] && b[2 * n + 1] == b[2 * n + 1]) {
        bb.x1 = fminf(bb.x1, b[2 * n]); bb.y1 = fminf(bb.y1, b[2 * n + 1]);
        bb.x2 = fmaxf(bb.x2, b[2 * n]); bb.y2 = fmaxf(bb.y2, b[2 * n + 1]);
        ++bb.rows;
      }
    }
    out[0] = bbox_iou(ba, bb);
    out[1] = oks(a, ba, b, N, (float)(kappa * kappa));  // kappa**2 is a python float, cast once by the tensor op
  }
}

extern "C" int snb_pair_similarity(const void* a, const void* b, int N, int is_f64, double kappa, double* out,
                                   void* stream) {
  if (N < 0 || !out) return SNB_ERR_BAD_ARG;
  pair_similarity_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a, b, N, is_f64, kappa, out);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_filter_instances(const snb_filter_config* cfg, int B, int I, int N, const float* kpts,
                                    const float* vals, const float* scores, const float* centroids,
                                    const float* centroid_vals, float* o_kpts, float* o_vals, float* o_scores,
                                    float* o_centroids, float* o_centroid_vals, void* stream) {
  if (!cfg || B < 0 || I < 0 || N < 0) return SNB_ERR_BAD_ARG;
  if ((kpts == nullptr) != (o_kpts == nullptr) || (vals == nullptr) != (o_vals == nullptr) ||
      (scores == nullptr) != (o_scores == nullptr) || (centroids == nullptr) != (o_centroids == nullptr) ||
      (centroid_vals == nullptr) != (o_centroid_vals == nullptr))
    return SNB_ERR_BAD_ARG;
  if (B == 0 || I == 0) return SNB_OK;
  const size_t smem = (size_t)I * (sizeof(InstBox) + 2 * sizeof(int) + 1) + 16;
  if (smem > 200 * 1024) return SNB_ERR_UNSUPPORTED;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(filter_instances_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return SNB_ERR_CUDA_LAUNCH;
  const FilterCfg c{cfg->min_peak_value, cfg->min_visible_node_fraction, cfg->min_instance_score, cfg->min_mean_node_score,
                    cfg->oks_kappa_sq, cfg->min_visible_nodes, cfg->overlapping, cfg->overlapping_threshold,
                    cfg->min_centroid_distance_sq};
  filter_instances_kernel<<<B, 32, smem, (cudaStream_t)stream>>>(c, I, N, kpts, vals, scores, centroids, centroid_vals,
                                                               o_kpts, o_vals, o_scores, o_centroids, o_centroid_vals);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_nms_greedy_f64(const double* pts, const double* scores, const int* frame_start, int n_frames,
                                  int max_per_frame, int N, int method, double threshold, double kappa, int* keep,
                                  int* keep_count, void* stream) {
  if (n_frames < 0 || N < 0 || max_per_frame < 0 || (method != 0 && method != 1)) return SNB_ERR_BAD_ARG;
  if (n_frames == 0) return SNB_OK;
  if (!pts || !scores || !frame_start || !keep || !keep_count) return SNB_ERR_BAD_ARG;
  const size_t smem = (size_t)max_per_frame * (sizeof(BoxD) + 2 * sizeof(int)) + 16;
  if (smem > 200 * 1024) return SNB_ERR_UNSUPPORTED;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(nms_greedy_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return SNB_ERR_CUDA_LAUNCH;
  nms_greedy_f64_kernel<<<n_frames, 32, smem, (cudaStream_t)stream>>>(pts, scores, frame_start, N, method, threshold, kappa,
                                                                     keep, keep_count);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_instance_stats_f64(const double* pts, const double* point_scores, long long total, int N,
                                      int* n_visible, double* mean_score, void* stream) {
  if (total < 0 || N < 0) return SNB_ERR_BAD_ARG;
  if (total == 0) return SNB_OK;
  if (!pts || !n_visible) return SNB_ERR_BAD_ARG;
  instance_stats_f64_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(pts, point_scores, total, N,
                                                                                             n_visible, mean_score);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
