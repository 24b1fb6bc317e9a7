// Device-side building blocks of the PAF grouping path, shared by the stand-alone kernels
// (paf.cu) and the fused per-frame tail kernel (pipeline.cu).  Citations paf.py:NN are into
// sleap_nn/inference/ops/paf.py.
#pragma once

#include "common.cuh"

namespace snb {

// ------------------------------------------------------------------------------------------
// Frame addressing: a "table" is either padded (frame b starts at b*stride) or CSR (explicit
// start array).  Counts are clamped to the stride for padded tables.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ long long tbl_start(const int* start, int stride, int b) {
  return start ? (long long)start[b] : (long long)b * stride;
}
__device__ __forceinline__ int tbl_count(const int* count, const int* start, int stride, int b) {
  const int n = count[b];
  return start ? n : min(n, stride);
}

// ------------------------------------------------------------------------------------------
// Line sampling arithmetic shared by make_line_subs / get_paf_lines / the fused scorer.
// Bit-exact restatement: slope = (dst - src) / fl32(1 + eps); val = src + slope*t;
// q = rint(val / stride) (half-to-even); clip.  `t` comes from torch.linspace on the host.
// ------------------------------------------------------------------------------------------
#define SNB_ONE_PLUS_EPS 1.00000011920928955078125f

__device__ __forceinline__ int line_coord(float src, float dst, float t, float stride, int hi) {
  const float slope = __fdiv_rn(__fsub_rn(dst, src), SNB_ONE_PLUS_EPS);
  const float val = __fadd_rn(src, __fmul_rn(slope, t));
  const float q = rintf(__fdiv_rn(val, stride));
  // float -> int32 like ATen's CPU cast, then clip (paf.py:192-208); NaN / out-of-range end up clipped
  int qi;
  if (!(q >= -2147483648.f)) qi = INT_MIN;  // NaN or below range
  else if (q >= 2147483648.f) qi = INT_MIN;  // x86 cvttss2si overflow value, clipped to 0 like the reference
  else qi = (int)q;
  return min(max(qi, 0), hi);
}

struct ScoreArgs {
  const void* pafs;         // may be null: enumerate candidates only
  int dt;                   // SNB_DTYPE_* of the PAF tensor
  long long pb, py, px, pc; // element strides of the (B, H, W, 2E) view
  int H, W;
  const float* t;           // n_points linspace table
  int n_points;
  float stride;
  float max_edge_length;
  float penalty_weight;
};

__device__ __forceinline__ float score_candidate(const ScoreArgs& a, int b, int k, float sx, float sy, float dx,
                                                 float dy) {
  // spatial vector, its length and unit direction (paf.py:381-388)
  const float vx = __fsub_rn(dx, sx), vy = __fsub_rn(dy, sy);
  const float len = sqrtf(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)));
  const float ux = __fdiv_rn(vx, len), uy = __fdiv_rn(vy, len);
  const void* fb = elem_ptr(a.pafs, (long long)b * a.pb + (long long)(2 * k) * a.pc, a.dt);
  double acc = 0.0;
  for (int p0 = 0; p0 < a.n_points; p0 += 8) {  // 16 independent gathers in flight
    float fx[8], fy[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int p = p0 + u;
      if (p < a.n_points) {
        const float t = a.t[p];  // global or shared memory
        const int col = line_coord(sx, dx, t, a.stride, a.W - 1);
        const int row = line_coord(sy, dy, t, a.stride, a.H - 1);
        const long long q = (long long)row * a.py + (long long)col * a.px;
        fx[u] = ld_elem(fb, q, a.dt);
        fy[u] = ld_elem(fb, q + a.pc, a.dt);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (p0 + u < a.n_points) acc += (double)__fadd_rn(__fmul_rn(fx[u], ux), __fmul_rn(fy[u], uy));  // paf.py:392
  }
  const float mean = (float)(acc / (double)a.n_points);                                   // paf.py:407
  const float pen = __fmul_rn(fminf(__fsub_rn(__fdiv_rn(a.max_edge_length, len), 1.f), 0.f), a.penalty_weight);
  return __fadd_rn(mean, pen);                                                            // paf.py:408
}

// ------------------------------------------------------------------------------------------
// K5: rectangular linear-sum assignment with scipy's exact semantics (Crouse 2016 shortest
// augmenting path, float64 duals).  Sequential per problem so that tie handling is identical to
// scipy's: free columns are kept in a list filled in reverse, the scan prefers, among equal
// reduced costs, the LAST unassigned column met (else the first minimum); a tall matrix is
// solved transposed and reported rows-ascending.  `cost(i, j)` is an accessor in the ORIGINAL
// orientation.  Returns false when infeasible (scipy raises ValueError).
// Workspace (nr <= nc after the transpose, D = nc): u[nr] v[nc] spc[nc] doubles,
// path[nc] col4row[nr] row4col[nc] free_[nc] ints, in_sr[nr] in_sc[nc] bytes.
// ------------------------------------------------------------------------------------------
__host__ __device__ inline size_t lsap_ws_bytes(int max_dim) {
  return (((size_t)max_dim * (3 * 8 + 4 * 4 + 2) + 16) + 15) & ~(size_t)15;  // keeps every problem 16B-aligned
}

template <typename CostFn>
__device__ bool lsap_solve(int n_rows, int n_cols, CostFn cost, void* ws, int* out_row, int* out_col) {
  const bool transposed = n_cols < n_rows;
  const int nr = transposed ? n_cols : n_rows;
  const int nc = transposed ? n_rows : n_cols;
  if (nr == 0) return true;
  double* u = (double*)ws;
  double* v = u + nr;
  double* spc = v + nc;
  int* path = (int*)(spc + nc);
  int* col4row = path + nc;
  int* row4col = col4row + nr;
  int* free_ = row4col + nc;
  unsigned char* in_sr = (unsigned char*)(free_ + nc);
  unsigned char* in_sc = in_sr + nr;
  auto c_at = [&](int i, int j) -> double { return transposed ? cost(j, i) : cost(i, j); };
  for (int i = 0; i < nr; ++i) { u[i] = 0.0; col4row[i] = -1; }
  for (int j = 0; j < nc; ++j) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
  for (int cur = 0; cur < nr; ++cur) {
    for (int j = 0; j < nc; ++j) { spc[j] = INFINITY; in_sc[j] = 0; free_[j] = nc - 1 - j; }
    for (int i = 0; i < nr; ++i) in_sr[i] = 0;
    int n_free = nc, i = cur, sink = -1;
    double min_val = 0.0;
    while (sink == -1) {
      in_sr[i] = 1;
      double lowest = INFINITY;
      int pick = -1;
      const double ui = u[i];
      for (int it = 0; it < n_free; ++it) {
        const int j = free_[it];
        const double r = min_val + c_at(i, j) - ui - v[j];
        if (r < spc[j]) { path[j] = i; spc[j] = r; }
        if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; pick = it; }
      }
      min_val = lowest;
      if (min_val == INFINITY) return false;
      const int j = free_[pick];
      if (row4col[j] == -1) sink = j; else i = row4col[j];
      in_sc[j] = 1;
      free_[pick] = free_[--n_free];
    }
    u[cur] += min_val;
    for (int r_ = 0; r_ < nr; ++r_)
      if (in_sr[r_] && r_ != cur) u[r_] += min_val - spc[col4row[r_]];
    for (int j = 0; j < nc; ++j)
      if (in_sc[j]) v[j] -= min_val - spc[j];
    int j = sink;
    while (true) {
      const int ii = path[j];
      row4col[j] = ii;
      const int prev = col4row[ii];
      col4row[ii] = j;
      j = prev;
      if (ii == cur) break;
    }
  }
  if (!transposed) {
    for (int i = 0; i < nr; ++i) { out_row[i] = i; out_col[i] = col4row[i]; }
  } else {
    // original rows = our columns: report ascending original row, i.e. ascending col4row value
    int k = 0;
    for (int j = 0; j < nc; ++j)
      if (row4col[j] >= 0) { out_row[k] = j; out_col[k] = row4col[j]; ++k; }
  }
  return true;
}



// Warp-cooperative version for max(n_rows, n_cols) <= 32: bit-identical results (same fp64 expressions per
// column, same tie rules), but the scan over the free columns - the inner loop of every Dijkstra step - runs one
// column per lane.  Lane `it` holds free_[it]; scipy's scan order and tie rule (among equal reduced costs take
// the LAST unassigned column met, else the FIRST minimum) become two ballots.  in_sr / in_sc are 32-bit masks.
// Called by all 32 lanes convergently; `ws` as for lsap_solve (shared memory).  Frames with hundreds of
// peaks spend most of the serial solver's time in shared-memory latency chains; this removes them.
template <typename CostFn>
__device__ bool lsap_solve_warp(int n_rows, int n_cols, CostFn cost, void* ws, int* out_row, int* out_col, int lane) {
  const bool transposed = n_cols < n_rows;
  const int nr = transposed ? n_cols : n_rows;
  const int nc = transposed ? n_rows : n_cols;
  if (nr == 0) return true;
  double* u = (double*)ws;
  double* v = u + nr;
  double* spc = v + nc;
  int* path = (int*)(spc + nc);
  int* col4row = path + nc;
  int* row4col = col4row + nr;
  auto c_at = [&](int i, int j) -> double { return transposed ? cost(j, i) : cost(i, j); };
  if (lane < nr) { u[lane] = 0.0; col4row[lane] = -1; }
  if (lane < nc) { v[lane] = 0.0; row4col[lane] = -1; path[lane] = -1; }
  __syncwarp();
  for (int cur = 0; cur < nr; ++cur) {
    if (lane < nc) spc[lane] = INFINITY;
    int my_col = nc - 1 - lane;  // free_[it] = nc - 1 - it
    unsigned in_sr = 0, in_sc = 0;
    int n_free = nc, i = cur, sink = -1;
    double min_val = 0.0;
    __syncwarp();
    while (sink == -1) {
      in_sr |= 1u << i;
      const double ui = u[i];
      const bool active = lane < n_free;
      double val = INFINITY;
      bool unassigned = false;
      if (active) {
        const int j = my_col;
        const double r = min_val + c_at(i, j) - ui - v[j];
        if (r < spc[j]) { path[j] = i; spc[j] = r; }
        val = spc[j];
        unassigned = row4col[j] == -1;
      }
      double m = val;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) m = fmin(m, __shfl_xor_sync(FULL, m, d));
      if (m == INFINITY) return false;  // infeasible (uniform)
      const unsigned eq = __ballot_sync(FULL, active && val == m);
      const unsigned eq_un = __ballot_sync(FULL, active && val == m && unassigned);
      const int pick = eq_un ? (31 - __clz(eq_un)) : (__ffs(eq) - 1);
      min_val = m;
      const int j = __shfl_sync(FULL, my_col, pick);
      const int rc = row4col[j];
      if (rc == -1) sink = j; else i = rc;
      in_sc |= 1u << j;
      const int last_col = __shfl_sync(FULL, my_col, n_free - 1);
      if (lane == pick) my_col = last_col;  // free_[pick] = free_[--n_free]
      --n_free;
      __syncwarp();
    }
    if (lane == 0) u[cur] += min_val;
    if (lane < nr && lane != cur && ((in_sr >> lane) & 1u)) u[lane] += min_val - spc[col4row[lane]];
    if (lane < nc && ((in_sc >> lane) & 1u)) v[lane] -= min_val - spc[lane];
    __syncwarp();
    if (lane == 0) {
      int j = sink;
      while (true) {
        const int ii = path[j];
        row4col[j] = ii;
        const int prev = col4row[ii];
        col4row[ii] = j;
        j = prev;
        if (ii == cur) break;
      }
    }
    __syncwarp();
  }
  if (!transposed) {
    if (lane < nr) { out_row[lane] = lane; out_col[lane] = col4row[lane]; }
  } else {
    // original rows = our columns: report ascending original row
    const bool has = lane < nc && row4col[lane] >= 0;
    const unsigned mk = __ballot_sync(FULL, has);
    if (has) {
      const int k = __popc(mk & ((1u << lane) - 1));
      out_row[k] = lane;
      out_col[k] = row4col[lane];
    }
  }
  __syncwarp();
  return true;
}

// ------------------------------------------------------------------------------------------
// Per-frame grouping of peaks by node (one warp).  ns[0..N] = exclusive prefix of peaks per node,
// node_peaks = peak indices grouped by node, ascending inside a node (a STABLE grouping: the
// reference's torch.argsort is stable only for n <= 16; this is the canonical
// order).  Peaks whose channel is outside [0, N) belong to no node (paf.py:110-112).
// `cursor` is an (N+1)-int scratch; all pointers may be shared or global memory.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void group_by_node_warp(const int* chan, int P, int n_nodes, int* ns, int* cursor,
                                                   int* node_peaks, int lane) {
  for (int k = lane; k <= n_nodes; k += 32) cursor[k] = 0;
  __syncwarp();
  for (int i = lane; i < P; i += 32) {
    const int c = chan[i];
    if (c >= 0 && c < n_nodes) atomicAdd(&cursor[c], 1);
  }
  __syncwarp();
  if (lane == 0) {  // exclusive scan over nodes (N is small)
    int acc = 0;
    for (int k = 0; k < n_nodes; ++k) {
      const int c = cursor[k];
      cursor[k] = acc;
      acc += c;
    }
    cursor[n_nodes] = acc;
  }
  __syncwarp();
  for (int k = lane; k <= n_nodes; k += 32) ns[k] = cursor[k];
  __syncwarp();
  for (int i0 = 0; i0 < P; i0 += 32) {  // stable placement, 32 peaks at a time
    const int i = i0 + lane;
    const int c = (i < P) ? chan[i] : -1;
    const bool valid = (c >= 0 && c < n_nodes);
    const unsigned peers = __match_any_sync(FULL, valid ? c : -1 - lane);
    if (valid) node_peaks[cursor[c] + __popc(peers & ((1u << lane) - 1))] = i;
    __syncwarp();
    if (valid && (__ffs(peers) - 1) == lane) cursor[c] += __popc(peers);
    __syncwarp();
  }
}

// eo[0..E] = exclusive prefix of candidates per edge (n_src * n_dst), mo[0..E] = of matches
// per edge (min(n_src, n_dst)).  Single thread.
__device__ __forceinline__ void edge_offsets(const int* edges, int n_nodes, int n_edges, const int* ns, int* eo,
                                             int* mo) {
  int acc_c = 0, acc_m = 0;
  for (int k = 0; k < n_edges; ++k) {
    const int s = edges[2 * k], d = edges[2 * k + 1];
    const int cs = (s >= 0 && s < n_nodes) ? ns[s + 1] - ns[s] : 0;
    const int cd = (d >= 0 && d < n_nodes) ? ns[d + 1] - ns[d] : 0;
    eo[k] = acc_c;
    mo[k] = acc_m;
    acc_c += cs * cd;
    acc_m += min(cs, cd);
  }
  eo[n_edges] = acc_c;
  mo[n_edges] = acc_m;
}

// Candidate m of a frame -> (edge k, source peak, destination peak): edge-major, source-major.
__device__ __forceinline__ void decode_candidate(int m, const int* eo, int n_edges, const int* ns, const int* edges,
                                                 const int* node_peaks, int* k_out, int* ps, int* pd) {
  int lo = 0, hi = n_edges;  // largest k with eo[k] <= m
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (eo[mid] <= m) lo = mid; else hi = mid;
  }
  const int s = edges[2 * lo], d = edges[2 * lo + 1];
  const int nd = ns[d + 1] - ns[d];
  const int r = m - eo[lo];
  *k_out = lo;
  *ps = node_peaks[ns[s] + r / nd];
  *pd = node_peaks[ns[d] + r % nd];
}

// ------------------------------------------------------------------------------------------
// K6: greedy instance assembly of ONE frame by ONE warp, faithful to the reference's dict logic:
//   - edges visited in `sorted` order, connections in list order, only score >= min_line_scores
//   - neither peak owned -> new id = max(current ids) + 1;  src owned -> dst joins;  both owned ->
//     dst is MOVED first, then the two instances merge iff their node sets are disjoint;
//     "src free, dst owned" does nothing (paf.py:754-789)
//   - min_instance_peaks filter (:791-818), ids compacted in ascending order (:845-850)
//   - instance score: fp32 running sum in connection order (:853-865)
//   - scatter in first-assignment order, later entries overwrite (:879-885), NaN fill
// A peak is addressed as (node, rank within node) -> np_[ns[node] + rank].
// Scratch: owner / order / id_count / id_rank (P ints each), fa / fb (n_nodes bytes each).
// ------------------------------------------------------------------------------------------
struct AsmFrame {
  const float* xy; const float* val; const int* chan; int P;
  const int* ns; const int* np_; int n_nodes;
  const int* edges; const int* sorted; int n_sorted;
  const int* m_edge; const int* m_src; const int* m_dst; const float* m_score; int K;
  const int* mo = nullptr;  // optional (n_edges+1) per-edge offsets into the match list (matches grouped by edge):
                            // visits [mo[e], min(mo[e+1], K)) instead of scanning all K matches for every edge
  int min_instance_peaks; float min_line_scores;
  int* owner; int* order; int* id_count; int* id_rank; unsigned char* fa; unsigned char* fb;
  int inst_cap; float* oxy; float* oval; float* osc; int* n_inst_out; int* status;
  int* stamps = nullptr; long long t0 = 0;  // profiling build (-DSNB_TAIL_TIMING): clock64 stamps of the assembly's phases
  // optional scratch (the fused tail lends its candidate-score table, free once the assignments are solved): with at
  // least 2 * K + n_edges + n_sorted + inst_cap + 2 words the post-loop passes run flattened over the connections in
  // visiting order instead of edge by edge (see assemble_frame_warp)
  int* scratch = nullptr; int scratch_words = 0; int n_edges = 0;
  // the match list is known to be a proper matching per edge (every peak at most once as a source and once as a
  // destination of that edge) - true for the fused tail, whose assignments come from its own solver; the grouping
  // entry points take arbitrary lists and leave it false (the chunk path then checks for repeated peaks itself)
  bool proper = false;
};
#ifdef SNB_TAIL_TIMING
#define SNB_ASM_STAMP(k) do { if (lane == 0 && f.stamps) f.stamps[(k)] = (int)(clock64() - f.t0); } while (0)
#else
#define SNB_ASM_STAMP(k) do {} while (0)
#endif

// One connection, applied by the whole warp exactly as the reference's loop body does (paf.py:754-789).  Used when
// a chunk of an edge's connections cannot be applied together (a merge candidate, a repeated peak, a self edge).
__device__ __forceinline__ void assemble_one(const AsmFrame& f, int lane, int pa, int pb, int& n_order) {
  const int P = f.P;
  const int ia = f.owner[pa], ib = f.owner[pb];
  __syncwarp();  // every lane has read owner[] before lane 0 rewrites it below (WAR hazard found by racecheck)
  if (ia < 0 && ib < 0) {
    int mx = -1;
    for (int i = lane; i < P; i += 32) mx = max(mx, f.owner[i]);
    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, d));
    __syncwarp();
    if (lane == 0) {
      f.owner[pa] = mx + 1;
      f.owner[pb] = mx + 1;
      f.order[n_order] = pa;
      if (pb != pa) f.order[n_order + 1] = pb;
    }
    n_order += (pb != pa) ? 2 : 1;
  } else if (ia >= 0 && ib < 0) {
    if (lane == 0) { f.owner[pb] = ia; f.order[n_order] = pb; }
    n_order += 1;
  } else if (ia >= 0 && ib >= 0) {
    if (lane == 0) f.owner[pb] = ia;
    __syncwarp();
    if (ia != ib) {
      for (int k = lane; k < f.n_nodes; k += 32) { f.fa[k] = 0; f.fb[k] = 0; }
      __syncwarp();
      for (int i = lane; i < P; i += 32) {
        const int o = f.owner[i];
        const int c = f.chan[i];
        if (c >= 0 && c < f.n_nodes) {
          if (o == ia) f.fa[c] = 1;
          if (o == ib) f.fb[c] = 1;
        }
      }
      __syncwarp();
      int hit = 0;
      for (int k = lane; k < f.n_nodes; k += 32) hit |= (f.fa[k] & f.fb[k]);
      hit = __any_sync(FULL, hit);
      if (!hit)
        for (int i = lane; i < P; i += 32)
          if (f.owner[i] == ib) f.owner[i] = ia;
    }
  }
  __syncwarp();
}

// The greedy loop is sequential in the reference, but within ONE edge a proper matching touches every peak at most
// once, so as long as no connection of a 32-wide chunk joins two existing instances ("both owned, different ids": the
// only case that renames ids) the chunk's connections commute: each lane takes one connection, new ids are handed
// out by a prefix count over the "neither owned" lanes (= max id so far + 1 + rank, exactly the sequential values;
// the running maximum is kept in a register and re-scanned only after a fallback chunk), and the first-assignment
// list is appended at prefix offsets.  Chunks with a merge candidate, a repeated peak or a self edge fall back to
// assemble_one per connection.  ncu + clock64 stamps on busy frames (32 nodes, 8 animals): the one-connection-
// at-a-time version spent 265 us of the tail's 348 us here, half of it in a lane-0 loop that accumulated the
// instance scores with a global-memory read-modify-write per connection.
// Does the lent scratch hold the flat-mode tables (spos, voff, vrank, vscore, qa, qb, qpos, acc)?
__device__ __forceinline__ bool asm_flat_ok(const AsmFrame& f) {
  return f.scratch && f.mo && f.scratch_words >= 5 * f.K + f.n_edges + f.n_sorted + f.inst_cap + 2;
}

// Flat-mode tables of the assembly (see assemble_frame_warp), shared with the CTA-wide pre-pass.
// asm_scan_edges (one warp): spos[e] = position of edge e in `sorted` (-1 = never visited), voff[se] = first visiting
// position of sorted edge se; returns the number of visiting positions.
__device__ __forceinline__ int asm_scan_edges(const AsmFrame& f, int lane, int* spos, int* voff) {
  const int K = f.K;
  for (int e = lane; e < f.n_edges; e += 32) spos[e] = -1;
  __syncwarp();
  int run = 0;
  for (int s0 = 0; s0 < f.n_sorted; s0 += 32) {  // exclusive scan of the per-edge connection counts
    const int se = s0 + lane;
    int cnt = 0;
    if (se < f.n_sorted) {
      const int e = f.sorted[se];
      spos[e] = se;
      cnt = min(f.mo[e + 1], K) - min(f.mo[e], K);
    }
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(FULL, inc, d);
      if (lane >= d) inc += t;
    }
    if (se < f.n_sorted) voff[se] = run + inc - cnt;
    run += __shfl_sync(FULL, inc, 31);
  }
  if (lane == 0) voff[f.n_sorted] = run;
  __syncwarp();
  return run;
}

// asm_prepass: connections first, first + stride, ... -> (qa, qb, qpos).  Any group of threads; spos / voff complete.
__device__ __forceinline__ void asm_prepass(const AsmFrame& f, int first, int stride, const int* spos, const int* voff,
                                            int* qa, int* qb, int* qpos) {
  const int K = f.K;
  for (int m = first; m < K; m += stride) {
    int a = -1, b = -1, pos = -1;
    const int e = f.m_edge[m];
    if (e >= 0 && e < f.n_edges) {
      const int se = spos[e];
      const int m_lo = min(f.mo[e], K), m_hi = min(f.mo[e + 1], K);
      if (se >= 0 && m >= m_lo && m < m_hi && (f.m_score[m] >= f.min_line_scores)) {  // visited, paf.py:993
        const int sn = f.edges[2 * e], dn = f.edges[2 * e + 1];
        const bool sn_ok = sn >= 0 && sn < f.n_nodes, dn_ok = dn >= 0 && dn < f.n_nodes;
        const int s0 = sn_ok ? f.ns[sn] : 0, n_src = sn_ok ? f.ns[sn + 1] - s0 : 0;
        const int d0 = dn_ok ? f.ns[dn] : 0, n_dst = dn_ok ? f.ns[dn + 1] - d0 : 0;
        const int sp = f.m_src[m], dp = f.m_dst[m];
        const bool sp_ok = sp >= 0 && sp < n_src, dp_ok = dp >= 0 && dp < n_dst;
        // the loop applies a connection only when both nodes exist and both ranks are in range (anything else is
        // reported); the score sums count it as soon as its source resolves
        if (!(sn_ok && dn_ok && sp_ok && dp_ok)) atomicOr(f.status, SNB_STATUS_BAD_INDEX);
        if (sp_ok) { a = f.np_[s0 + sp]; pos = voff[se] + (m - m_lo); }
        if (dp_ok) b = f.np_[d0 + dp];
        if (!(sn_ok && dn_ok)) b = -1;  // never applied; still counted through `a` when the source node exists
      }
    }
    qa[m] = a;
    qb[m] = b;
    qpos[m] = pos;
  }
}

// Hand-over from the sequential part of the assembly (one warp) to its CTA-wide finish (assemble_finish_cta): lives in
// shared memory.  split == 1 means the warp stopped after the id compaction and everything after it is still to do.
struct AsmSplit {
  int pre;        // assemble_prepass_cta ran: ownership tables initialised, flat tables filled
  int forest;     // assemble_forest_cta produced owner[] / order[] / n_order: the sequential loop is skipped
  int split, n_inst, n_order, n_vis;
};

__device__ __forceinline__ void assemble_frame_warp(const AsmFrame& f, int lane, AsmSplit* hand = nullptr) {
  const int P = f.P, K = f.K;
  const unsigned lt = (1u << lane) - 1u;
  const bool pre = hand && hand->pre;  // owner / id_count initialised and the flat tables filled by the whole CTA
  if (!pre) {
    for (int i = lane; i < P; i += 32) { f.owner[i] = -1; f.id_count[i] = 0; }
    __syncwarp();
  }
  int n_order = 0, mx = -1;
  // ---- flat mode (scratch lent by the caller, matches grouped by edge).  A single warp issues in order, so every level
  // of a dependent shared-memory chain (sorted -> edges -> ns, m_src -> np_ -> owner) stalls it for a full latency, and
  // the edge-by-edge loop below spent ~750 cycles per edge mostly on such chains.  Here everything that does not depend
  // on the evolving ownership is resolved for ALL connections in one parallel pre-pass:
  //   qa[m]   source peak of connection m, or -1 when the loop never applies it / the sums never count it
  //   qb[m]   destination peak, or -1
  //   qpos[m] its position in visiting order (edges in `sorted` order, connections in list order), or -1
  // so that the sequential part per edge is: read qa / qb (one level), read owner[] (second level), ballots, writes.
  const bool flat = asm_flat_ok(f);
  int* spos = f.scratch;                        // n_edges : position of edge e in `sorted`, -1 = never visited
  int* voff = spos + f.n_edges;                 // n_sorted + 1 : first visiting position of each sorted edge
  int* vrank = voff + f.n_sorted + 1;           // K : instance rank at each visiting position
  float* vscore = reinterpret_cast<float*>(vrank + K);  // K
  int* qa = reinterpret_cast<int*>(vscore + K); // K
  int* qb = qa + K;                             // K
  int* qpos = qb + K;                           // K
  float* acc = reinterpret_cast<float*>(qpos + K);  // inst_cap
  int n_vis = 0;
  if (flat) {
    if (pre) {
      n_vis = hand->n_vis;
    } else {
      n_vis = asm_scan_edges(f, lane, spos, voff);
      asm_prepass(f, lane, 32, spos, voff, qa, qb, qpos);
      __syncwarp();
    }
    const bool forest = pre && hand->forest;  // owner[] / order[] already final (assemble_forest_cta)
    if (forest) n_order = hand->n_order;
    for (int c0 = 0; !forest && c0 < f.n_sorted; c0 += 32) {
      // headers of 32 sorted edges at a time, one per lane; the loop below fetches them by shuffle
      int h_e = -1, h_lo = 0, h_hi = 0, h_distinct = 0;
      if (c0 + lane < f.n_sorted) {
        h_e = f.sorted[c0 + lane];
        h_lo = min(f.mo[h_e], K);
        h_hi = min(f.mo[h_e + 1], K);
        h_distinct = f.edges[2 * h_e] != f.edges[2 * h_e + 1];
      }
      const int c1 = min(32, f.n_sorted - c0);
      for (int j = 0; j < c1; ++j) {
        const int e = __shfl_sync(FULL, h_e, j), m_lo = __shfl_sync(FULL, h_lo, j), m_hi = __shfl_sync(FULL, h_hi, j);
        const bool distinct = __shfl_sync(FULL, h_distinct, j) != 0;
        for (int mb = m_lo; mb < m_hi; mb += 32) {
          const int m = mb + lane;
          int pa = -1 - lane, pb = -1 - lane;  // distinct placeholders for __match_any_sync
          bool act = false;
          if (m < m_hi) {
            const int a = qa[m], b = qb[m];
            act = a >= 0 && b >= 0 && f.m_edge[m] == e;
            if (act) { pa = a; pb = b; }
          }
          const unsigned am = __ballot_sync(FULL, act);
          if (am == 0) continue;
          const int ia = act ? f.owner[pa] : -1, ib = act ? f.owner[pb] : -1;
          const bool k1 = act && ia < 0 && ib < 0, k2 = act && ia >= 0 && ib < 0;
          const bool merge = act && ia >= 0 && ib >= 0 && ia != ib;
          bool repeated = false;
          if (!f.proper) {
            const unsigned same_a = __match_any_sync(FULL, pa), same_b = __match_any_sync(FULL, pb);
            repeated = __popc(same_a) > 1 || __popc(same_b) > 1;
          }
          __syncwarp();  // owner[] reads above happen before any write below
          if (distinct && !__any_sync(FULL, merge || repeated)) {
            const unsigned b1 = __ballot_sync(FULL, k1), b2 = __ballot_sync(FULL, k2);
            const int off = n_order + 2 * __popc(b1 & lt) + __popc(b2 & lt);
            if (k1) {
              const int id = mx + 1 + __popc(b1 & lt);
              f.owner[pa] = id;
              f.owner[pb] = id;
              f.order[off] = pa;
              f.order[off + 1] = pb;
            } else if (k2) {
              f.owner[pb] = ia;
              f.order[off] = pb;
            }
            mx += __popc(b1);
            n_order += 2 * __popc(b1) + __popc(b2);
            __syncwarp();
          } else {
            unsigned todo = am;
            while (todo) {
              const int l = __ffs(todo) - 1;
              todo &= todo - 1;
              assemble_one(f, lane, __shfl_sync(FULL, pa, l), __shfl_sync(FULL, pb, l), n_order);
            }
            mx = -1;  // ids may have been renamed: re-scan the running maximum
            for (int i = lane; i < P; i += 32) mx = max(mx, f.owner[i]);
            for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, d));
          }
        }
      }
    }
  } else {
  // per-edge header (edge id, node offsets, match range): a chain of three dependent loads.  The NEXT edge's header is
  // fetched before the current edge's connections are applied, so the chain overlaps the body instead of heading it.
  struct EdgeHdr { int e, s0, d0, n_src, n_dst, m_lo, m_hi; bool distinct; };
  auto load_hdr = [&](int se) -> EdgeHdr {
    EdgeHdr hd;
    hd.e = f.sorted[se];
    const int sn = f.edges[2 * hd.e], dn = f.edges[2 * hd.e + 1];
    const bool nodes_ok = sn >= 0 && sn < f.n_nodes && dn >= 0 && dn < f.n_nodes;
    hd.s0 = nodes_ok ? f.ns[sn] : 0;
    hd.d0 = nodes_ok ? f.ns[dn] : 0;
    hd.n_src = nodes_ok ? f.ns[sn + 1] - hd.s0 : 0;
    hd.n_dst = nodes_ok ? f.ns[dn + 1] - hd.d0 : 0;
    hd.m_lo = f.mo ? min(f.mo[hd.e], K) : 0;
    hd.m_hi = f.mo ? min(f.mo[hd.e + 1], K) : K;
    hd.distinct = sn != dn;
    return hd;
  };
  EdgeHdr nxt = f.n_sorted > 0 ? load_hdr(0) : EdgeHdr{};
  for (int se = 0; se < f.n_sorted; ++se) {
    const EdgeHdr hd = nxt;
    if (se + 1 < f.n_sorted) nxt = load_hdr(se + 1);
    const int e = hd.e, s0 = hd.s0, d0 = hd.d0, n_src = hd.n_src, n_dst = hd.n_dst, m_lo = hd.m_lo, m_hi = hd.m_hi;
    for (int mb = m_lo; mb < m_hi; mb += 32) {
      const int m = mb + lane;
      bool act = m < m_hi && f.m_edge[m] == e && (f.m_score[m] >= f.min_line_scores);  // paf.py:993
      int pa = -1 - lane, pb = -1 - lane;  // distinct placeholders for __match_any_sync
      if (act) {
        const int sp = f.m_src[m], dp = f.m_dst[m];
        if (sp < 0 || dp < 0 || sp >= n_src || dp >= n_dst) {
          atomicOr(f.status, SNB_STATUS_BAD_INDEX);
          act = false;
        } else {
          pa = f.np_[s0 + sp];
          pb = f.np_[d0 + dp];
        }
      }
      const unsigned am = __ballot_sync(FULL, act);
      if (am == 0) continue;
      const int ia = act ? f.owner[pa] : -1, ib = act ? f.owner[pb] : -1;
      const bool c1 = act && ia < 0 && ib < 0, c2 = act && ia >= 0 && ib < 0;
      const bool merge = act && ia >= 0 && ib >= 0 && ia != ib;
      bool repeated = false;
      if (!f.proper) {  // (warp-uniform) two match.any per chunk are a large part of an edge's ~1 100 cycles
        const unsigned same_a = __match_any_sync(FULL, pa), same_b = __match_any_sync(FULL, pb);  // both by ALL lanes
        repeated = __popc(same_a) > 1 || __popc(same_b) > 1;
      }
      __syncwarp();  // owner[] reads above happen before any write below
      if (hd.distinct && !__any_sync(FULL, merge || repeated)) {
        const unsigned b1 = __ballot_sync(FULL, c1), b2 = __ballot_sync(FULL, c2);
        const int off = n_order + 2 * __popc(b1 & lt) + __popc(b2 & lt);
        if (c1) {
          const int id = mx + 1 + __popc(b1 & lt);
          f.owner[pa] = id;
          f.owner[pb] = id;
          f.order[off] = pa;
          f.order[off + 1] = pb;
        } else if (c2) {
          f.owner[pb] = ia;
          f.order[off] = pb;
        }  // "both owned, same id": owner[pb] = ia is a no-op; "src free, dst owned": nothing (paf.py:754-789)
        mx += __popc(b1);
        n_order += 2 * __popc(b1) + __popc(b2);
        __syncwarp();
      } else {
        unsigned todo = am;
        while (todo) {
          const int l = __ffs(todo) - 1;
          todo &= todo - 1;
          assemble_one(f, lane, __shfl_sync(FULL, pa, l), __shfl_sync(FULL, pb, l), n_order);
        }
        mx = -1;  // ids may have been renamed: re-scan the running maximum
        for (int i = lane; i < P; i += 32) mx = max(mx, f.owner[i]);
        for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, d));
      }
    }
  }
  }  // edge-by-edge greedy loop (no scratch)
  SNB_ASM_STAMP(6);
  // instance sizes, min_instance_peaks filter, ascending-id compaction (paf.py:791-818, :845-850)
  for (int i = lane; i < P; i += 32)
    if (f.owner[i] >= 0) atomicAdd(&f.id_count[f.owner[i]], 1);
  __syncwarp();
  int n_inst = 0;
  for (int i0 = 0; i0 < P; i0 += 32) {
    const int id = i0 + lane;
    const bool keep = id < P && f.id_count[id] > 0 && (f.min_instance_peaks <= 0 || f.id_count[id] >= f.min_instance_peaks);
    const unsigned kb = __ballot_sync(FULL, keep);
    if (id < P) f.id_rank[id] = keep ? n_inst + __popc(kb & lt) : -1;
    n_inst += __popc(kb);
  }
  __syncwarp();
  if (n_inst > f.inst_cap) {
    if (lane == 0) { atomicOr(f.status, SNB_STATUS_INSTANCE_OVERFLOW); *f.n_inst_out = n_inst; }
    return;
  }
  SNB_ASM_STAMP(7);
  // a caller with a whole CTA at hand finishes from here with all of its warps (flat mode only, and only when the
  // scratch also holds one word per output slot for the scatter)
  if (hand && flat && f.scratch_words >= 5 * K + f.n_edges + f.n_sorted + f.inst_cap + 2 + f.inst_cap * f.n_nodes) {
    if (lane == 0) { hand->n_inst = n_inst; hand->n_order = n_order; hand->n_vis = n_vis; hand->split = 1; }
    return;
  }
  for (int i = lane; i < n_inst * f.n_nodes; i += 32) { f.oxy[2 * i] = NAN; f.oxy[2 * i + 1] = NAN; f.oval[i] = NAN; }
  if (lane == 0) *f.n_inst_out = n_inst;
  // instance score = fp32 running sum of its connections' scores in visiting order (paf.py:853-865): lane r owns
  // instance r and walks the connection list (uniform, broadcast loads), adding the ones that belong to it.  The
  // connection -> instance map is precomputed in parallel into id_count[] (free by now) when it fits (K <= P);
  // otherwise each step resolves it on the fly (a chain of four dependent loads).
  // The same pass replays the reference's sanity check (ops/paf.py:866-873): a visited connection whose source is in a
  // kept instance must have its destination in the SAME instance - otherwise the reference raises (AssertionError, or
  // KeyError when the destination is in no kept instance).  Only improper matchings fed to the grouping API get there.
  if (flat) {
    // ---- flattened passes.  One parallel pass over ALL connections resolves each one's instance rank and writes
    // (rank, score) at its visiting position; the score sums then walk that list 32 positions at a time: lanes holding
    // the same instance form a group (__match_any_sync) whose lowest lane adds the group's scores in lane (= visiting)
    // order on top of the instance's running sum - the strict left-to-right fp32 sum of paf.py:853-865, without every
    // lane scanning every connection (was 18.6 + 12.4 us of a busy frame's 53 us).
    for (int r = lane; r < n_inst; r += 32) acc[r] = 0.f;
    for (int v = lane; v < n_vis; v += 32) vrank[v] = -1;
    __syncwarp();
    for (int m = lane; m < K; m += 32) {
      const int pos = qpos[m];
      if (pos < 0) continue;  // never visited, or its source does not resolve
      const int o = f.owner[qa[m]];
      const int rk = (o >= 0) ? f.id_rank[o] : -1;
      vrank[pos] = rk;
      vscore[pos] = f.m_score[m];
      const int b = qb[m];
      if (rk >= 0 && b >= 0) {  // the reference's sanity check (ops/paf.py:866-873)
        const int od = f.owner[b];
        const int rd = (od >= 0) ? f.id_rank[od] : -1;
        if (rd < 0) atomicOr(f.status, SNB_STATUS_ASM_MISSING);
        else if (rd != rk) atomicOr(f.status, SNB_STATUS_ASM_MISMATCH);
      }
    }
    __syncwarp();
    SNB_ASM_STAMP(8);
    for (int v0 = 0; v0 < n_vis; v0 += 32) {
      const int v = v0 + lane;
      const int rk = v < n_vis ? vrank[v] : -1;
      const float sc = v < n_vis ? vscore[v] : 0.f;
      const unsigned peers = __match_any_sync(FULL, rk >= 0 ? rk : -1 - lane);
      const bool leader = rk >= 0 && (__ffs(peers) - 1) == lane;
      const unsigned any_valid = __ballot_sync(FULL, rk >= 0);
      if (any_valid == 0) continue;
      float a = leader ? acc[rk] : 0.f;
      const int last = 31 - __clz(any_valid);
      for (int k = __ffs(any_valid) - 1; k <= last; ++k) {
        const float t = __shfl_sync(FULL, sc, k);
        if (leader && ((peers >> k) & 1u)) a = __fadd_rn(a, t);
      }
      if (leader) acc[rk] = a;
      __syncwarp();
    }
    for (int r = lane; r < n_inst; r += 32) f.osc[r] = acc[r];
  } else {
  int* m_rank = (K <= P) ? f.id_count : nullptr;
  __syncwarp();
  if (m_rank) {
    for (int m = lane; m < K; m += 32) m_rank[m] = -1;
    __syncwarp();
  }
  for (int se = 0; se < f.n_sorted; ++se) {
    const int e = f.sorted[se];
    const int sn = f.edges[2 * e], dn = f.edges[2 * e + 1];
    if (sn < 0 || sn >= f.n_nodes) continue;
    const int s0 = f.ns[sn], n_src = f.ns[sn + 1] - s0;
    const bool dn_ok = dn >= 0 && dn < f.n_nodes;
    const int d0 = dn_ok ? f.ns[dn] : 0, n_dst = dn_ok ? f.ns[dn + 1] - d0 : 0;
    const int m_lo = f.mo ? min(f.mo[e], K) : 0, m_hi = f.mo ? min(f.mo[e + 1], K) : K;
    for (int m = m_lo + lane; m < m_hi; m += 32) {
      if (f.m_edge[m] != e || !(f.m_score[m] >= f.min_line_scores)) continue;
      const int sp = f.m_src[m], dp = f.m_dst[m];
      if (sp < 0 || sp >= n_src) continue;
      const int o = f.owner[f.np_[s0 + sp]];
      const int rk = (o >= 0) ? f.id_rank[o] : -1;
      if (m_rank) m_rank[m] = rk;
      if (rk >= 0 && dp >= 0 && dp < n_dst) {
        const int od = f.owner[f.np_[d0 + dp]];
        const int rd = (od >= 0) ? f.id_rank[od] : -1;
        if (rd < 0) atomicOr(f.status, SNB_STATUS_ASM_MISSING);
        else if (rd != rk) atomicOr(f.status, SNB_STATUS_ASM_MISMATCH);
      }
    }
  }
  __syncwarp();
  SNB_ASM_STAMP(8);
  for (int r0 = 0; r0 < n_inst; r0 += 32) {
    const int r = r0 + lane;
    float acc = 0.f;
    for (int se = 0; se < f.n_sorted; ++se) {
      const int e = f.sorted[se];
      const int sn = f.edges[2 * e];
      if (sn < 0 || sn >= f.n_nodes) continue;
      const int s0 = f.ns[sn], n_src = f.ns[sn + 1] - s0;
      const int m_lo = f.mo ? min(f.mo[e], K) : 0, m_hi = f.mo ? min(f.mo[e + 1], K) : K;
      if (m_rank) {
        for (int m = m_lo; m < m_hi; ++m)
          if (f.m_edge[m] == e && m_rank[m] == r) acc = __fadd_rn(acc, f.m_score[m]);
      } else {
        for (int m = m_lo; m < m_hi; ++m) {
          const float sc = f.m_score[m];
          if (f.m_edge[m] != e || !(sc >= f.min_line_scores)) continue;
          const int sp = f.m_src[m];
          if (sp < 0 || sp >= n_src) continue;
          const int o = f.owner[f.np_[s0 + sp]];
          if (o >= 0 && f.id_rank[o] == r) acc = __fadd_rn(acc, sc);
        }
      }
    }
    if (r < n_inst) f.osc[r] = acc;
  }
  }  // edge-by-edge passes (no scratch)
  SNB_ASM_STAMP(9);
  __syncwarp();  // the NaN fill above is ordered before the scatter below
  // scatter in first-assignment order, later entries overwrite (paf.py:879-885): 32 entries per round, and inside a
  // round only the LAST lane aiming at a slot writes
  for (int t0 = 0; t0 < n_order; t0 += 32) {
    const int t = t0 + lane;
    long long slot = -1 - lane;
    int i = 0;
    if (t < n_order) {
      i = f.order[t];
      const int r = f.id_rank[f.owner[i]];
      if (r >= 0) slot = (long long)r * f.n_nodes + f.chan[i];
    }
    const unsigned same = __match_any_sync(FULL, slot);
    if (slot >= 0 && (same >> lane) == 1u) {  // no higher lane has the same slot
      f.oxy[2 * slot] = f.xy[2 * i];
      f.oxy[2 * slot + 1] = f.xy[2 * i + 1];
      f.oval[slot] = f.val[i];
    }
    __syncwarp();
  }
}

// CTA-wide start of the assembly: ownership tables initialised and the flat-mode tables filled by all threads (one
// connection per thread instead of eight rounds of one warp).  Every thread of the CTA calls it; the caller puts a
// barrier between this and assemble_frame_warp(f, lane, hand).
__device__ __forceinline__ void assemble_prepass_cta(const AsmFrame& f, AsmSplit* hand, int tid, int n_threads) {
  if (!asm_flat_ok(f)) return;  // (uniform) the warp does everything itself
  const int K = f.K;
  int* spos = f.scratch;
  int* voff = spos + f.n_edges;
  int* qa = voff + f.n_sorted + 1 + 2 * K;
  int* qb = qa + K;
  int* qpos = qb + K;
  if (tid < 32) {
    const int n_vis = asm_scan_edges(f, tid, spos, voff);
    if (tid == 0) hand->n_vis = n_vis;
  }
  for (int i = tid; i < f.P; i += n_threads) { f.owner[i] = -1; f.id_count[i] = 0; }
  __syncthreads();
  asm_prepass(f, tid, n_threads, spos, voff, qa, qb, qpos);
  if (tid == 0) hand->pre = 1;
}

constexpr int ASM_FOREST_MIN_EDGES = 12;  // visited edges from which assemble_forest_cta replaces the sequential loop

// CTA-wide replacement of the sequential edge loop for the common case.  On a skeleton visited as a forest, parents
// first (what toposort_edges' order gives), the loop only ever sees "neither owned" (a new instance) and "source owned,
// destination free" (the destination joins): every peak is a destination at most once, before any use as a source,
// and nothing is renamed.  The result is then a forest over the peaks (destination -> source links), an instance id =
// the number of roots created earlier in visiting order, and order[] = per connection in visiting order [source if it
// creates the instance], destination - scans over the visiting positions and pointer jumping instead of one dependent
// step per edge.  The conditions are CHECKED on the frame's actual connections; anything else (a node with two visited
// parents, children before parents, repeated edges, self edges, improper matchings) leaves hand->forest at 0 and the
// sequential loop runs.  (As ONE warp this cost what the loop costs - ~400 cycles per pass; it pays with one peak /
// connection per thread.)  Every thread of the CTA calls it, after assemble_prepass_cta and a barrier.
__device__ __forceinline__ void assemble_forest_cta(const AsmFrame& f, AsmSplit* hand, int tid, int n_threads) {
  const int K = f.K, P = f.P;
  if (!hand->pre) return;  // (uniform)
  // a dozen CTA barriers cost more than a short loop: 4 edges (cfg3) 2.4 K cycles sequentially, 5.8 K this way;
  // 31 edges (cfg4) 15 K against 8 K
  if (f.n_sorted < ASM_FOREST_MIN_EDGES) return;
  const int n_vis = hand->n_vis;
  if (f.scratch_words < 5 * K + f.n_edges + f.n_sorted + f.inst_cap + 2 + f.inst_cap * f.n_nodes + P) return;
  constexpr int INF = 0x7fffffff;
  int* spos = f.scratch;
  int* vcre = f.scratch + f.n_edges + f.n_sorted + 1;   // (= vrank) per visiting position: creates an instance -> its id
  int* vcnt = vcre + K;                                 // (= vscore) per visiting position: order[] entries -> offset
  int* qa = vcnt + K;
  int* qb = qa + K;
  int* qpos = qb + K;
  int* par2 = qpos + K + f.inst_cap + f.inst_cap * f.n_nodes;  // behind acc and the scatter's slot table
  int* par = f.owner;        // destination -> source; roots: -2 - id; untouched: -1 (as initialised by the pre-pass)
  int* dposv = f.id_rank;    // visiting position of the connection that has the peak as destination
  int* sfirst = f.id_count;  // first visiting position with the peak as source
  bool bad = n_vis > K;
  for (int i = tid; i < P; i += n_threads) { dposv[i] = INF; sfirst[i] = INF; }
  for (int v = tid; v < n_vis && v < K; v += n_threads) { vcre[v] = 0; vcnt[v] = 0; }
  for (int se = tid; se < f.n_sorted; se += n_threads) bad |= spos[f.sorted[se]] != se;  // an edge listed twice
  __syncthreads();
  for (int m = tid; m < K; m += n_threads) {
    const int a = qa[m], b = qb[m];
    if (a < 0 || b < 0) continue;
    const int pos = qpos[m], e = f.m_edge[m];
    bad |= f.edges[2 * e] == f.edges[2 * e + 1];             // self edge
    atomicMin(&sfirst[a], pos);
    if (atomicMin(&dposv[b], pos) != INF) bad = true;        // a destination twice
    else par[b] = a;                                         // (a single writer per peak, also when `bad`)
  }
  __syncthreads();
  for (int m = tid; m < K; m += n_threads) {
    const int a = qa[m], b = qb[m];
    if (a < 0 || b < 0) continue;
    const int pos = qpos[m], da = dposv[a];
    if (!(da == INF || da < pos)) bad = true;   // the source joins its own instance only later
    if (!(sfirst[b] > pos)) bad = true;         // the destination was already used as a source
    const int cre = (da == INF && sfirst[a] == pos) ? 1 : 0;
    vcre[pos] = cre;
    vcnt[pos] = 1 + cre;
  }
  if (__syncthreads_or(bad ? 1 : 0)) {  // not a forest: back to the initial state, the sequential loop takes over
    for (int i = tid; i < P; i += n_threads) { f.owner[i] = -1; f.id_count[i] = 0; }
    return;
  }
  if (tid < 32) {  // exclusive scans over the visiting positions: ids of the created instances, offsets into order[]
    const int lane = tid;
    int run_c = 0, run_o = 0;
    for (int v0 = 0; v0 < n_vis; v0 += 32) {
      const int v = v0 + lane;
      const int c = v < n_vis ? vcre[v] : 0, o = v < n_vis ? vcnt[v] : 0;
      int ic = c, io = o;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int tc = __shfl_up_sync(FULL, ic, d), to = __shfl_up_sync(FULL, io, d);
        if (lane >= d) { ic += tc; io += to; }
      }
      if (v < n_vis) { vcre[v] = run_c + ic - c; vcnt[v] = run_o + io - o; }
      run_c += __shfl_sync(FULL, ic, 31);
      run_o += __shfl_sync(FULL, io, 31);
    }
    if (lane == 0) hand->n_order = run_o;
  }
  __syncthreads();
  for (int m = tid; m < K; m += n_threads) {
    const int a = qa[m], b = qb[m];
    if (a < 0 || b < 0) continue;
    const int pos = qpos[m], off = vcnt[pos];
    if (dposv[a] == INF && sfirst[a] == pos) {
      par[a] = -2 - vcre[pos];
      f.order[off] = a;
      f.order[off + 1] = b;
    } else {
      f.order[off] = b;
    }
  }
  __syncthreads();
  // pointer jumping, double buffered (par -> par2 -> par ...): a touched peak's parent is itself touched, so a pointer
  // never lands on -1; <= ceil(log2(depth)) + 1 rounds
  int* src = par;
  int* dst = par2;
  for (int round = 0; round < 32; ++round) {
    int moved = 0;
    for (int i = tid; i < P; i += n_threads) {
      int q = src[i];
      if (q >= 0) { q = src[q]; moved = 1; }
      dst[i] = q;
    }
    const int any = __syncthreads_or(moved);
    int* t = src; src = dst; dst = t;
    if (!any) break;
  }
  for (int i = tid; i < P; i += n_threads) {
    const int q = src[i];  // (src may be f.owner itself: same thread, same index, read before write)
    f.owner[i] = q <= -2 ? -2 - q : -1;
    f.id_count[i] = 0;
  }
  if (tid == 0) hand->forest = 1;
}

// CTA-wide finish of assemble_frame_warp (after the id compaction): NaN fill, the rank pass over all connections, the
// score sums (still one warp: a strict left-to-right fp32 sum per instance) and the scatter.  Every thread of the CTA
// calls it; `hand` was written by the warp that ran the sequential part and is visible (the caller put a barrier in
// between).  The scatter's "later entries overwrite" (paf.py:879-885) becomes: per output slot the LAST position of the
// first-assignment list aiming at it (atomicMax), then only that entry writes.
__device__ __forceinline__ void assemble_finish_cta(const AsmFrame& f, const AsmSplit* hand, int tid, int n_threads) {
  if (!hand->split) return;
  const int K = f.K, n_inst = hand->n_inst, n_order = hand->n_order, n_vis = hand->n_vis;
  const int lane = tid & 31, warp = tid >> 5;
  int* vrank = f.scratch + f.n_edges + f.n_sorted + 1;
  float* vscore = reinterpret_cast<float*>(vrank + K);
  int* qa = reinterpret_cast<int*>(vscore + K);
  int* qb = qa + K;
  int* qpos = qb + K;
  float* acc = reinterpret_cast<float*>(qpos + K);
  int* last = reinterpret_cast<int*>(acc + f.inst_cap);  // inst_cap x n_nodes
  const int n_slots = n_inst * f.n_nodes;
  for (int i = tid; i < n_slots; i += n_threads) {
    f.oxy[2 * i] = NAN;
    f.oxy[2 * i + 1] = NAN;
    f.oval[i] = NAN;
    last[i] = -1;
  }
  if (tid == 0) *f.n_inst_out = n_inst;
  for (int r = tid; r < n_inst; r += n_threads) acc[r] = 0.f;
  for (int v = tid; v < n_vis; v += n_threads) vrank[v] = -1;
  __syncthreads();
  for (int m = tid; m < K; m += n_threads) {
    const int pos = qpos[m];
    if (pos < 0) continue;  // never visited, or its source does not resolve
    const int o = f.owner[qa[m]];
    const int rk = (o >= 0) ? f.id_rank[o] : -1;
    vrank[pos] = rk;
    vscore[pos] = f.m_score[m];
    const int b = qb[m];
    if (rk >= 0 && b >= 0) {  // the reference's sanity check (ops/paf.py:866-873)
      const int od = f.owner[b];
      const int rd = (od >= 0) ? f.id_rank[od] : -1;
      if (rd < 0) atomicOr(f.status, SNB_STATUS_ASM_MISSING);
      else if (rd != rk) atomicOr(f.status, SNB_STATUS_ASM_MISMATCH);
    }
  }
  for (int t = tid; t < n_order; t += n_threads) {
    const int i = f.order[t];
    const int r = f.id_rank[f.owner[i]];
    if (r >= 0) atomicMax(&last[r * f.n_nodes + f.chan[i]], t);
  }
  __syncthreads();
  if (warp == 0) {
    SNB_ASM_STAMP(8);
    for (int v0 = 0; v0 < n_vis; v0 += 32) {
      const int v = v0 + lane;
      const int rk = v < n_vis ? vrank[v] : -1;
      const float sc = v < n_vis ? vscore[v] : 0.f;
      const unsigned peers = __match_any_sync(FULL, rk >= 0 ? rk : -1 - lane);
      const bool leader = rk >= 0 && (__ffs(peers) - 1) == lane;
      const unsigned any_valid = __ballot_sync(FULL, rk >= 0);
      if (any_valid == 0) continue;
      float a = leader ? acc[rk] : 0.f;
      const int last_lane = 31 - __clz(any_valid);
      for (int k = __ffs(any_valid) - 1; k <= last_lane; ++k) {
        const float t = __shfl_sync(FULL, sc, k);
        if (leader && ((peers >> k) & 1u)) a = __fadd_rn(a, t);
      }
      if (leader) acc[rk] = a;
      __syncwarp();
    }
    for (int r = lane; r < n_inst; r += 32) f.osc[r] = acc[r];
    SNB_ASM_STAMP(9);
  }
  // the scatter runs on the other warps meanwhile (all of them when the CTA is a single warp)
  const int w0 = n_threads > 32 ? 32 : 0;
  if (tid >= w0) {
    for (int t = tid - w0; t < n_order; t += n_threads - w0) {
      const int i = f.order[t];
      const int r = f.id_rank[f.owner[i]];
      if (r < 0) continue;
      const int slot = r * f.n_nodes + f.chan[i];
      if (last[slot] == t) {
        f.oxy[2 * slot] = f.xy[2 * i];
        f.oxy[2 * slot + 1] = f.xy[2 * i + 1];
        f.oval[slot] = f.val[i];
      }
    }
  }
}

}  // namespace snb
