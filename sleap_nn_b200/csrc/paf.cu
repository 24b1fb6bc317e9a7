// PAF grouping on device: candidate enumeration, line-integral scoring, per-edge optimal
// assignment and greedy instance assembly (replaces sleap_nn/inference/ops/paf.py, cited as
// paf.py:NN, and the two-knot use of inference/utils.py:interp1d).
//
//   K4  paf_prepare + paf_score      candidates (paf.py:84-130), line subscripts (:133-234,
//                                    utils.py:29-130), gather (:237-287), score (:335-410)
//   K5  match_structured / generic   scipy.optimize.linear_sum_assignment semantics (:500-619)
//   K6  assemble                     assign_connections_to_instances + make_predicted_instances
//                                    + the min_line_scores filter (:705-887, :992-1038)
//
// All of these touch O(#peaks) data; they are latency-bound, not bandwidth-bound.  The PAF
// tensor is never streamed: each candidate reads n_points x 2 scalars through its strides
// (the caller's (B,H,W,2E) permuted VIEW of a (B,2E,H,W) tensor is read in place).
#include "paf_device.cuh"

namespace snb {

// ------------------------------------------------------------------------------------------
// K4a: per-frame preparation.  One warp per frame.
//   node_start[b][0..N]  exclusive prefix of peaks per node
//   node_peaks[...]      peak indices (local to the frame) grouped by node, ascending index
//                        inside a node = STABLE grouping; the reference's torch.argsort is
//                        stable only for n <= 16, this is the canonical order
//   edge_off[b][0..E]    exclusive prefix of candidates per edge (n_src * n_dst)
//   match_off[b][0..E]   exclusive prefix of matches per edge (min(n_src, n_dst))
// Peaks whose channel is outside [0, N) belong to no node (paf.py:110-112).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
paf_prepare_kernel(const int* __restrict__ peak_chan, const int* __restrict__ frame_start, int frame_stride,
                   const int* __restrict__ frame_count, const int* __restrict__ edges, int n_nodes, int n_edges,
                   int* __restrict__ node_start, int* __restrict__ node_peaks, int* __restrict__ edge_off,
                   int* __restrict__ match_off) {
  extern __shared__ int s_cnt[];  // n_nodes + 1 cursors
  const int b = blockIdx.x, lane = threadIdx.x;
  const long long base = tbl_start(frame_start, frame_stride, b);
  const int P = tbl_count(frame_count, frame_start, frame_stride, b);
  int* ns = node_start + (long long)b * (n_nodes + 1);
  group_by_node_warp(peak_chan + base, P, n_nodes, ns, s_cnt, node_peaks + base, lane);
  if (lane == 0)
    edge_offsets(edges, n_nodes, n_edges, ns, edge_off + (long long)b * (n_edges + 1),
                 match_off + (long long)b * (n_edges + 1));
}

// K4b: one thread per candidate.  Candidate m of frame b: edge k by search in edge_off, then
// (i_src, i_dst) = divmod(m - edge_off[k], n_dst)  -> edge-major, source-major order.
__global__ void __launch_bounds__(128)
paf_score_kernel(ScoreArgs a, const float* __restrict__ peak_xy, const int* __restrict__ frame_start,
                 int frame_stride, const int* __restrict__ edges, int n_nodes, int n_edges,
                 const int* __restrict__ node_start, const int* __restrict__ node_peaks,
                 const int* __restrict__ edge_off, const int* __restrict__ cand_start, int cand_stride,
                 int* __restrict__ cand_edge, long long* __restrict__ cand_epi, float* __restrict__ cand_score,
                 int* __restrict__ status) {
  const int b = blockIdx.y;
  const int* eo = edge_off + (long long)b * (n_edges + 1);
  const int M = eo[n_edges];
  const int limit = cand_start ? M : min(M, cand_stride);
  if (M > limit && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(status, SNB_STATUS_CAND_OVERFLOW);
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= limit) return;
  const int* ns = node_start + (long long)b * (n_nodes + 1);
  const long long base = tbl_start(frame_start, frame_stride, b);
  int k, ps, pd;
  decode_candidate(m, eo, n_edges, ns, edges, node_peaks + base, &k, &ps, &pd);
  const long long o = tbl_start(cand_start, cand_stride, b) + m;
  cand_edge[o] = k;
  cand_epi[2 * o] = ps;
  cand_epi[2 * o + 1] = pd;
  if (a.pafs) {
    const float* xy = peak_xy + 2 * base;
    cand_score[o] = score_candidate(a, b, k, xy[2 * ps], xy[2 * ps + 1], xy[2 * pd], xy[2 * pd + 1]);
  }
}

// make_line_subs (paf.py:133-234): (M, n_points, 2, 3) int32 [row, col, channel].
__global__ void line_subs_kernel(const float* __restrict__ peaks, long long n_peaks, const long long* __restrict__ epi,
                                 const int* __restrict__ edge_inds, long long M, const float* __restrict__ t,
                                 int n_points, float stride, int H, int W, int* __restrict__ out,
                                 int* __restrict__ status) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * n_points) return;
  const long long m = i / n_points;
  const int p = (int)(i % n_points);
  long long ps = epi[2 * m], pd = epi[2 * m + 1];
  if (ps < 0) ps += n_peaks;
  if (pd < 0) pd += n_peaks;
  if (ps < 0 || ps >= n_peaks || pd < 0 || pd >= n_peaks) {
    atomicOr(status, SNB_STATUS_BAD_INDEX);
    return;
  }
  const float tt = t[p];
  const int col = line_coord(peaks[2 * ps], peaks[2 * pd], tt, stride, W - 1);
  const int row = line_coord(peaks[2 * ps + 1], peaks[2 * pd + 1], tt, stride, H - 1);
  const int e = edge_inds[m];
  int* o = out + i * 6;
  o[0] = row; o[1] = col; o[2] = 2 * e;
  o[3] = row; o[4] = col; o[5] = 2 * e + 1;
}

// pafs_sample[line_subs] gather (paf.py:282-287): lines (M, n_points, 2) from a (H, W, Cn) strided view.
__global__ void paf_gather_kernel(const float* __restrict__ pafs, long long py, long long px, long long pc, int H,
                                  int W, int Cn, const int* __restrict__ subs, long long n_sub,
                                  float* __restrict__ out, int* __restrict__ status) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sub) return;
  int row = subs[3 * i], col = subs[3 * i + 1], ch = subs[3 * i + 2];
  if (row < 0) row += H;
  if (col < 0) col += W;
  if (ch < 0) ch += Cn;
  if (row < 0 || row >= H || col < 0 || col >= W || ch < 0 || ch >= Cn) {
    atomicOr(status, SNB_STATUS_BAD_INDEX);
    out[i] = NAN;
    return;
  }
  out[i] = pafs[(long long)row * py + (long long)col * px + (long long)ch * pc];
}

// score_paf_lines on pre-gathered lines (paf.py:335-410): one thread per candidate.
__global__ void score_lines_kernel(const float* __restrict__ lines, const float* __restrict__ peaks,
                                   long long n_peaks, const long long* __restrict__ epi, long long M, int n_points,
                                   float max_edge_length, float weight, float* __restrict__ out,
                                   int* __restrict__ status) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  long long ps = epi[2 * m], pd = epi[2 * m + 1];
  if (ps < 0) ps += n_peaks;
  if (pd < 0) pd += n_peaks;
  if (ps < 0 || ps >= n_peaks || pd < 0 || pd >= n_peaks) {
    atomicOr(status, SNB_STATUS_BAD_INDEX);
    out[m] = NAN;
    return;
  }
  const float vx = __fsub_rn(peaks[2 * pd], peaks[2 * ps]), vy = __fsub_rn(peaks[2 * pd + 1], peaks[2 * ps + 1]);
  const float len = sqrtf(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)));
  const float ux = __fdiv_rn(vx, len), uy = __fdiv_rn(vy, len);
  const float* l = lines + m * n_points * 2;
  double acc = 0.0;
  for (int p = 0; p < n_points; ++p) acc += (double)__fadd_rn(__fmul_rn(l[2 * p], ux), __fmul_rn(l[2 * p + 1], uy));
  const float mean = (float)(acc / (double)n_points);
  const float pen = __fmul_rn(fminf(__fsub_rn(__fdiv_rn(max_edge_length, len), 1.f), 0.f), weight);
  out[m] = __fadd_rn(mean, pen);
}

// compute_distance_penalty (paf.py:290-332), elementwise.
__global__ void distance_penalty_kernel(const float* __restrict__ len, long long n, float max_len, float weight,
                                        float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __fmul_rn(fminf(__fsub_rn(__fdiv_rn(max_len, len[i]), 1.f), 0.f), weight);
}

constexpr int LSAP_SMEM_DIM = 32;
constexpr int MATCH_WARPS = 4;

// K5a: structured matcher for the fused pipeline.  One warp per (frame, edge); candidates of an
// edge are the full n_src x n_dst cross product in source-major order, so cost(i, j) =
// -score[edge_off + i*n_dst + j] with NaN -> +inf (paf.py:582-586) is read in place.
__global__ void __launch_bounds__(32 * MATCH_WARPS)
match_structured_kernel(const float* __restrict__ cand_score, const int* __restrict__ cand_start, int cand_stride,
                        const int* __restrict__ edges, int n_nodes, int n_edges, const int* __restrict__ node_start,
                        const int* __restrict__ edge_off, const int* __restrict__ match_off,
                        const int* __restrict__ match_start, int match_stride, void* __restrict__ ws_global,
                        int ws_max_dim, int B, int* __restrict__ m_edge, int* __restrict__ m_src,
                        int* __restrict__ m_dst, float* __restrict__ m_score, int* __restrict__ m_count,
                        int* __restrict__ status) {
  __shared__ __align__(16) unsigned char s_ws[MATCH_WARPS][((LSAP_SMEM_DIM * 42 + 16) + 15) / 16 * 16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long prob = (long long)blockIdx.x * MATCH_WARPS + warp;
  if (prob >= (long long)B * n_edges) return;
  const int b = (int)(prob / n_edges), k = (int)(prob % n_edges);
  const int* ns = node_start + (long long)b * (n_nodes + 1);
  const int* eo = edge_off + (long long)b * (n_edges + 1);
  const int* mo = match_off + (long long)b * (n_edges + 1);
  const int M = eo[n_edges];
  const bool cand_ok = cand_start ? true : (M <= cand_stride);  // scores exist only if they fitted
  const int K = mo[n_edges];
  const int klimit = match_start ? K : min(K, match_stride);
  if (k == 0 && lane == 0) {
    m_count[b] = cand_ok ? klimit : 0;
    if (K > klimit) atomicOr(status, SNB_STATUS_MATCH_OVERFLOW);
  }
  if (!cand_ok) return;
  const int s = edges[2 * k], d = edges[2 * k + 1];
  const int n_src = ns[s + 1] - ns[s], n_dst = ns[d + 1] - ns[d];
  const int n_match = min(n_src, n_dst);
  if (n_match == 0 || mo[k] + n_match > klimit) return;
  const int dim = max(n_src, n_dst);
  const float* sc = cand_score + tbl_start(cand_start, cand_stride, b) + eo[k];
  auto cost = [&](int i, int j) -> double {
    const float x = sc[i * n_dst + j];
    return isnan(x) ? (double)INFINITY : -(double)x;
  };
  const long long o = tbl_start(match_start, match_stride, b) + mo[k];
  bool ok = true;
  if (dim <= LSAP_SMEM_DIM) {  // all 32 lanes: one free column per lane
    ok = lsap_solve_warp(n_src, n_dst, cost, s_ws[warp], m_src + o, m_dst + o, lane);
  } else if (dim <= ws_max_dim && ws_global) {
    if (lane == 0)
      ok = lsap_solve(n_src, n_dst, cost, (unsigned char*)ws_global + (size_t)prob * lsap_ws_bytes(ws_max_dim), m_src + o,
                      m_dst + o);
    ok = __shfl_sync(FULL, ok ? 1 : 0, 0) != 0;
    __syncwarp();
  } else {
    if (lane == 0) atomicOr(status, SNB_STATUS_LSAP_TOO_LARGE);
    return;
  }
  if (!ok) {
    if (lane == 0) atomicOr(status, SNB_STATUS_LSAP_INFEASIBLE);
    for (int r = lane; r < n_match; r += 32) { m_edge[o + r] = k; m_src[o + r] = -1; m_dst[o + r] = -1; m_score[o + r] = NAN; }
    return;
  }
  for (int r = lane; r < n_match; r += 32) {
    m_edge[o + r] = k;
    m_score[o + r] = sc[m_src[o + r] * n_dst + m_dst[o + r]];  // -cost, paf.py:592-594
  }
}

// K5b: generic matcher for arbitrary candidate lists (the public match_candidates_* API).
// One CTA per (edge, frame).  Distinct src / dst peak ids are ranked with shared-memory bitmaps
// (torch.unique + searchsorted, paf.py:564-581).  Phase 0 reports (n_src, n_dst); phase 1 builds
// the dense cost matrix (later duplicates win, like index_put) and solves it.
__global__ void __launch_bounds__(128)
match_generic_kernel(int phase, const int* __restrict__ cand_edge, const long long* __restrict__ cand_epi,
                     const float* __restrict__ cand_score, const int* __restrict__ cand_start,
                     const int* __restrict__ cand_count, int n_edges, int id_words, int* __restrict__ dims,
                     const long long* __restrict__ cost_off, double* __restrict__ cost, int* __restrict__ cell_src,
                     const int* __restrict__ match_start, void* __restrict__ ws_global, int ws_max_dim,
                     int* __restrict__ m_edge, int* __restrict__ m_src, int* __restrict__ m_dst,
                     float* __restrict__ m_score, int* __restrict__ status) {
  extern __shared__ unsigned s_bits[];  // [src bits | dst bits | src word prefix | dst word prefix]
  unsigned* sb = s_bits;
  unsigned* db = sb + id_words;
  unsigned* sp = db + id_words;
  unsigned* dp = sp + id_words;
  const int k = blockIdx.x, b = blockIdx.y;
  const long long c0 = cand_start[b];
  const int M = cand_count[b];
  for (int w = threadIdx.x; w < 2 * id_words; w += blockDim.x) s_bits[w] = 0;
  __syncthreads();
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    if (cand_edge[c0 + m] != k) continue;
    const long long ps = cand_epi[2 * (c0 + m)], pd = cand_epi[2 * (c0 + m) + 1];
    if (ps < 0 || pd < 0 || ps >= 32LL * id_words || pd >= 32LL * id_words) { atomicOr(status, SNB_STATUS_BAD_INDEX); continue; }
    atomicOr(&sb[ps >> 5], 1u << (ps & 31));
    atomicOr(&db[pd >> 5], 1u << (pd & 31));
  }
  __syncthreads();
  __shared__ int s_n[2];
  if (threadIdx.x == 0) {
    int a = 0, c = 0;
    for (int w = 0; w < id_words; ++w) { sp[w] = a; a += __popc(sb[w]); dp[w] = c; c += __popc(db[w]); }
    s_n[0] = a; s_n[1] = c;
  }
  __syncthreads();
  const int n_src = s_n[0], n_dst = s_n[1];
  const long long prob = (long long)b * n_edges + k;
  if (phase == 0) {
    if (threadIdx.x == 0) { dims[2 * prob] = n_src; dims[2 * prob + 1] = n_dst; }
    return;
  }
  const int n_match = min(n_src, n_dst);
  if (n_match == 0) return;
  double* cm = cost + cost_off[prob];
  int* cs = cell_src + cost_off[prob];
  const int cells = n_src * n_dst;
  for (int i = threadIdx.x; i < cells; i += blockDim.x) cs[i] = -1;
  __syncthreads();
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    if (cand_edge[c0 + m] != k) continue;
    const long long ps = cand_epi[2 * (c0 + m)], pd = cand_epi[2 * (c0 + m) + 1];
    if (ps < 0 || pd < 0 || ps >= 32LL * id_words || pd >= 32LL * id_words) continue;
    const int r = sp[ps >> 5] + __popc(sb[ps >> 5] & ((1u << (ps & 31)) - 1));
    const int c = dp[pd >> 5] + __popc(db[pd >> 5] & ((1u << (pd & 31)) - 1));
    atomicMax(&cs[r * n_dst + c], m);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cells; i += blockDim.x) {
    double c = INFINITY;
    if (cs[i] >= 0) {
      const float x = cand_score[c0 + cs[i]];
      c = isnan(x) ? (double)INFINITY : -(double)x;  // cost_matrix[np.isnan] = inf, paf.py:586
    }
    cm[i] = c;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int dim = max(n_src, n_dst);
  if (dim > ws_max_dim) { atomicOr(status, SNB_STATUS_LSAP_TOO_LARGE); return; }
  void* ws = (unsigned char*)ws_global + (size_t)prob * lsap_ws_bytes(ws_max_dim);
  auto costfn = [&](int i, int j) -> double { return cm[i * n_dst + j]; };
  const long long o = match_start[prob];
  if (!lsap_solve(n_src, n_dst, costfn, ws, m_src + o, m_dst + o)) {
    atomicOr(status, SNB_STATUS_LSAP_INFEASIBLE);
    return;
  }
  for (int r = 0; r < n_match; ++r) {
    m_edge[o + r] = k;
    // -cost_matrix_np[...] rounded to fp32 (paf.py:592-611)
    m_score[o + r] = (float)(-cm[m_src[o + r] * n_dst + m_dst[o + r]]);
  }
}

// K6 stand-alone kernel: one warp per frame, scratch in global memory (see assemble_frame_warp).
struct AsmArgs {
  const float* peak_xy;
  const float* peak_val;
  const int* peak_chan;
  const int* frame_start;
  int frame_stride;
  const int* frame_count;
  const int* node_start;
  const int* node_peaks;
  int n_nodes;
  const int* edges;
  const int* sorted_edges;
  int n_sorted;
  const int* m_edge;
  const int* m_src;
  const int* m_dst;
  const float* m_score;
  const int* match_start;
  int match_stride;
  const int* m_count;
  int min_instance_peaks;
  float min_line_scores;
  int* ws;        // B * 4 * ws_stride ints
  int ws_stride;  // >= max peaks per frame
  int inst_cap;
  float* inst_xy;     // (B, inst_cap, N, 2)
  float* inst_val;    // (B, inst_cap, N)
  float* inst_score;  // (B, inst_cap)
  int* n_inst;        // (B)
  int* status;
};

__global__ void __launch_bounds__(32) assemble_kernel(AsmArgs a) {
  extern __shared__ unsigned char s_flags[];  // 2 * n_nodes
  const int b = blockIdx.x, lane = threadIdx.x;
  const long long base = tbl_start(a.frame_start, a.frame_stride, b);
  const int P = tbl_count(a.frame_count, a.frame_start, a.frame_stride, b);
  const long long m0 = tbl_start(a.match_start, a.match_stride, b);
  if (P > a.ws_stride) {
    if (lane == 0) { atomicOr(a.status, SNB_STATUS_PEAK_OVERFLOW); a.n_inst[b] = 0; }
    return;
  }
  int* owner = a.ws + (long long)b * 4 * a.ws_stride;
  AsmFrame f;
  f.xy = a.peak_xy + 2 * base; f.val = a.peak_val + base; f.chan = a.peak_chan + base; f.P = P;
  f.ns = a.node_start + (long long)b * (a.n_nodes + 1); f.np_ = a.node_peaks + base; f.n_nodes = a.n_nodes;
  f.edges = a.edges; f.sorted = a.sorted_edges; f.n_sorted = a.n_sorted;
  f.m_edge = a.m_edge + m0; f.m_src = a.m_src + m0; f.m_dst = a.m_dst + m0; f.m_score = a.m_score + m0;
  f.K = a.match_start ? a.m_count[b] : min(a.m_count[b], a.match_stride);
  f.min_instance_peaks = a.min_instance_peaks; f.min_line_scores = a.min_line_scores;
  f.owner = owner; f.order = owner + a.ws_stride; f.id_count = owner + 2 * a.ws_stride; f.id_rank = owner + 3 * a.ws_stride;
  f.fa = s_flags; f.fb = s_flags + a.n_nodes;
  f.inst_cap = a.inst_cap;
  f.oxy = a.inst_xy + (long long)b * a.inst_cap * a.n_nodes * 2;
  f.oval = a.inst_val + (long long)b * a.inst_cap * a.n_nodes;
  f.osc = a.inst_score + (long long)b * a.inst_cap;
  f.n_inst_out = a.n_inst + b; f.status = a.status;
  assemble_frame_warp(f, lane);
}

// make_predicted_instances (paf.py:823-887) for the dict API: assignments arrive in insertion order
// with their compacted instance index; scatter (later entries overwrite) + fp32 running score sums.
__global__ void __launch_bounds__(32)
scatter_instances_kernel(const float* __restrict__ xy, const float* __restrict__ val, const int* __restrict__ inst,
                         const int* __restrict__ node, int n_assign, const int* __restrict__ conn_inst,
                         const float* __restrict__ conn_score, int n_conn, int n_inst, int n_nodes,
                         float* __restrict__ o_xy, float* __restrict__ o_val, float* __restrict__ o_score) {
  const int lane = threadIdx.x;
  for (int i = lane; i < n_inst * n_nodes; i += 32) { o_xy[2 * i] = NAN; o_xy[2 * i + 1] = NAN; o_val[i] = NAN; }
  for (int i = lane; i < n_inst; i += 32) o_score[i] = 0.f;
  __syncwarp();
  if (lane != 0) return;
  for (int c = 0; c < n_conn; ++c)
    if (conn_inst[c] >= 0) o_score[conn_inst[c]] = __fadd_rn(o_score[conn_inst[c]], conn_score[c]);
  for (int i = 0; i < n_assign; ++i) {
    const long long slot = (long long)inst[i] * n_nodes + node[i];
    o_xy[2 * slot] = xy[2 * i];
    o_xy[2 * slot + 1] = xy[2 * i + 1];
    o_val[slot] = val[i];
  }
}

// interp1d (inference/utils.py:29-130): searchsorted (left) - 1, clamped to [0, N-2];
// slope = (y1 - y0) / (eps + (x1 - x0)); out = y0 + slope * (xnew - x0), each op rounded in fp32.
__global__ void interp1d_kernel(const float* __restrict__ x, int x_rows, const float* __restrict__ y, int y_rows,
                                const float* __restrict__ xnew, int xn_rows, int n, int p, long long total,
                                float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long d = i / p;
  const int j = (int)(i % p);
  const float* xr = x + (x_rows == 1 ? 0 : d * n);
  const float* yr = y + (y_rows == 1 ? 0 : d * n);
  const float q = xnew[(xn_rows == 1 ? 0 : d * p) + j];
  int lo = 0, hi = n;  // first index with xr[idx] >= q  (NaN q -> n, as torch)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (xr[mid] < q) lo = mid + 1; else hi = mid;
  }
  if (isnan(q)) lo = n;
  const int k = min(max(lo - 1, 0), n - 2);
  // Reference quirk (utils.py:112-124): with ONE row of knots x the slope table is treated as flat and indexed by k
  // alone, i.e. every row takes ROW 0's slopes even when y has several rows (the intercept y[k] is still the row's own).
  const float* ys = (x_rows == 1) ? y : yr;
  const float slope = __fdiv_rn(__fsub_rn(ys[k + 1], ys[k]), __fadd_rn(1.1920928955078125e-07f, __fsub_rn(xr[k + 1], xr[k])));
  out[i] = __fadd_rn(yr[k], __fmul_rn(slope, __fsub_rn(q, xr[k])));
}

}  // namespace snb

using namespace snb;

extern "C" long long snb_lsap_workspace_bytes(int max_dim) { return (long long)lsap_ws_bytes(max_dim); }

extern "C" int snb_paf_prepare(const int* peak_chan, const int* frame_start, int frame_stride, const int* frame_count,
                               int B, const int* edges, int n_nodes, int n_edges, int* node_start, int* node_peaks,
                               int* edge_off, int* match_off, void* stream_) {
  if (B < 0 || n_nodes <= 0 || n_edges < 0) return SNB_ERR_BAD_ARG;
  if (B == 0) return SNB_OK;
  const size_t smem = sizeof(int) * (size_t)(n_nodes + 1);
  if (smem > 48 * 1024) return SNB_ERR_UNSUPPORTED;
  paf_prepare_kernel<<<B, 32, smem, (cudaStream_t)stream_>>>(peak_chan, frame_start, frame_stride, frame_count, edges,
                                                            n_nodes, n_edges, node_start, node_peaks, edge_off,
                                                            match_off);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_paf_score(const float* pafs, long long pb, long long py, long long px, long long pc, int H, int W,
                             const float* t_table, int n_points, float stride, float max_edge_length,
                             float penalty_weight, const float* peak_xy, const int* frame_start, int frame_stride,
                             int B, const int* edges, int n_nodes, int n_edges, const int* node_start,
                             const int* node_peaks, const int* edge_off, const int* cand_start, int cand_stride,
                             int max_cand_per_frame, int* cand_edge, long long* cand_epi, float* cand_score,
                             int* status, void* stream_) {
  return snb_paf_score_t(pafs, SNB_DTYPE_F32, pb, py, px, pc, H, W, t_table, n_points, stride, max_edge_length,
                         penalty_weight, peak_xy, frame_start, frame_stride, B, edges, n_nodes, n_edges, node_start,
                         node_peaks, edge_off, cand_start, cand_stride, max_cand_per_frame, cand_edge, cand_epi,
                         cand_score, status, stream_);
}

extern "C" int snb_paf_score_t(const void* pafs, int dtype, long long pb, long long py, long long px, long long pc,
                               int H, int W, const float* t_table, int n_points, float stride, float max_edge_length,
                               float penalty_weight, const float* peak_xy, const int* frame_start, int frame_stride,
                               int B, const int* edges, int n_nodes, int n_edges, const int* node_start,
                               const int* node_peaks, const int* edge_off, const int* cand_start, int cand_stride,
                               int max_cand_per_frame, int* cand_edge, long long* cand_epi, float* cand_score,
                               int* status, void* stream_) {
  if (B < 0 || n_edges < 0 || max_cand_per_frame < 0) return SNB_ERR_BAD_ARG;
  if (dtype != SNB_DTYPE_F32 && dtype != SNB_DTYPE_F16 && dtype != SNB_DTYPE_BF16) return SNB_ERR_BAD_ARG;
  if (B == 0 || n_edges == 0 || max_cand_per_frame == 0) return SNB_OK;
  if (B > 65535) return SNB_ERR_UNSUPPORTED;
  ScoreArgs a{pafs, dtype, pb, py, px, pc, H, W, t_table, n_points, stride, max_edge_length, penalty_weight};
  dim3 grid((max_cand_per_frame + 127) / 128, B);
  paf_score_kernel<<<grid, 128, 0, (cudaStream_t)stream_>>>(a, peak_xy, frame_start, frame_stride, edges, n_nodes,
                                                           n_edges, node_start, node_peaks, edge_off, cand_start,
                                                           cand_stride, cand_edge, cand_epi, cand_score, status);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_line_subs(const float* peaks, long long n_peaks, const long long* epi, const int* edge_inds,
                             long long M, const float* t_table, int n_points, float stride, int H, int W, int* out,
                             int* status, void* stream_) {
  const long long n = M * n_points;
  if (n <= 0) return SNB_OK;
  line_subs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(peaks, n_peaks, epi, edge_inds, M,
                                                                                   t_table, n_points, stride, H, W,
                                                                                   out, status);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_paf_gather(const float* pafs, long long py, long long px, long long pc, int H, int W, int Cn,
                              const int* subs, long long n_sub, float* out, int* status, void* stream_) {
  if (n_sub <= 0) return SNB_OK;
  paf_gather_kernel<<<(unsigned)((n_sub + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(pafs, py, px, pc, H, W, Cn,
                                                                                        subs, n_sub, out, status);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_score_lines(const float* lines, const float* peaks, long long n_peaks, const long long* epi,
                               long long M, int n_points, float max_edge_length, float weight, float* out,
                               int* status, void* stream_) {
  if (M <= 0) return SNB_OK;
  score_lines_kernel<<<(unsigned)((M + 127) / 128), 128, 0, (cudaStream_t)stream_>>>(lines, peaks, n_peaks, epi, M,
                                                                                     n_points, max_edge_length,
                                                                                     weight, out, status);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_distance_penalty(const float* lengths, long long n, float max_edge_length, float weight, float* out,
                                    void* stream_) {
  if (n <= 0) return SNB_OK;
  distance_penalty_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(lengths, n, max_edge_length,
                                                                                          weight, out);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_match_structured(const float* cand_score, const int* cand_start, int cand_stride, const int* edges,
                                    int n_nodes, int n_edges, const int* node_start, const int* edge_off,
                                    const int* match_off, const int* match_start, int match_stride, void* ws,
                                    int ws_max_dim, int B, int* m_edge, int* m_src, int* m_dst, float* m_score,
                                    int* m_count, int* status, void* stream_) {
  if (B < 0 || n_edges < 0) return SNB_ERR_BAD_ARG;
  if (B == 0) return SNB_OK;
  if (n_edges == 0) {
    if (cudaMemsetAsync(m_count, 0, sizeof(int) * B, (cudaStream_t)stream_) != cudaSuccess) return SNB_ERR_CUDA_LAUNCH;
    return SNB_OK;
  }
  const long long probs = (long long)B * n_edges;
  match_structured_kernel<<<(unsigned)((probs + MATCH_WARPS - 1) / MATCH_WARPS), 32 * MATCH_WARPS, 0,
                            (cudaStream_t)stream_>>>(cand_score, cand_start, cand_stride, edges, n_nodes, n_edges,
                                                     node_start, edge_off, match_off, match_start, match_stride, ws,
                                                     ws_max_dim, B, m_edge, m_src, m_dst, m_score, m_count, status);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_match_generic(int phase, const int* cand_edge, const long long* cand_epi, const float* cand_score,
                                 const int* cand_start, const int* cand_count, int B, int n_edges, int max_peak_id,
                                 int* dims, const long long* cost_off, double* cost, int* cell_src,
                                 const int* match_start, void* ws, int ws_max_dim, int* m_edge, int* m_src, int* m_dst,
                                 float* m_score, int* status, void* stream_) {
  if (B < 0 || n_edges < 0 || max_peak_id < 0) return SNB_ERR_BAD_ARG;
  if (B == 0 || n_edges == 0) return SNB_OK;
  if (B > 65535) return SNB_ERR_UNSUPPORTED;
  const int id_words = (max_peak_id + 32) / 32;
  const size_t smem = sizeof(unsigned) * 4 * (size_t)id_words;
  if (smem > 200 * 1024) return SNB_ERR_UNSUPPORTED;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(match_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return SNB_ERR_CUDA_LAUNCH;
  dim3 grid(n_edges, B);
  match_generic_kernel<<<grid, 128, smem, (cudaStream_t)stream_>>>(phase, cand_edge, cand_epi, cand_score, cand_start,
                                                                  cand_count, n_edges, id_words, dims, cost_off, cost,
                                                                  cell_src, match_start, ws, ws_max_dim, m_edge, m_src,
                                                                  m_dst, m_score, status);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_assemble(const float* peak_xy, const float* peak_val, const int* peak_chan, const int* frame_start,
                            int frame_stride, const int* frame_count, int B, const int* node_start,
                            const int* node_peaks, int n_nodes, const int* edges, const int* sorted_edges,
                            int n_sorted, const int* m_edge, const int* m_src, const int* m_dst, const float* m_score,
                            const int* match_start, int match_stride, const int* m_count, int min_instance_peaks,
                            float min_line_scores, int* ws, int ws_stride, int inst_cap, float* inst_xy,
                            float* inst_val, float* inst_score, int* n_inst, int* status, void* stream_) {
  if (B < 0 || n_nodes <= 0 || inst_cap < 0 || ws_stride < 0) return SNB_ERR_BAD_ARG;
  if (B == 0) return SNB_OK;
  if (2 * (size_t)n_nodes > 48 * 1024) return SNB_ERR_UNSUPPORTED;
  AsmArgs a{peak_xy, peak_val, peak_chan, frame_start, frame_stride, frame_count, node_start, node_peaks, n_nodes,
            edges, sorted_edges, n_sorted, m_edge, m_src, m_dst, m_score, match_start, match_stride, m_count,
            min_instance_peaks, min_line_scores, ws, ws_stride, inst_cap, inst_xy, inst_val, inst_score, n_inst,
            status};
  assemble_kernel<<<B, 32, 2 * (size_t)n_nodes, (cudaStream_t)stream_>>>(a);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_scatter_instances(const float* xy, const float* val, const int* inst, const int* node, int n_assign,
                                     const int* conn_inst, const float* conn_score, int n_conn, int n_inst,
                                     int n_nodes, float* o_xy, float* o_val, float* o_score, void* stream_) {
  if (n_inst <= 0) return SNB_OK;
  scatter_instances_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>(xy, val, inst, node, n_assign, conn_inst, conn_score,
                                                                n_conn, n_inst, n_nodes, o_xy, o_val, o_score);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_interp1d(const float* x, int x_rows, const float* y, int y_rows, const float* xnew, int xn_rows,
                            int n, int p, long long rows, float* out, void* stream_) {
  if (n < 2) return SNB_ERR_BAD_ARG;
  const long long total = rows * p;
  if (total <= 0) return SNB_OK;
  interp1d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(x, x_rows, y, y_rows, xnew,
                                                                                      xn_rows, n, p, total, out);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
