// Fused bottom-up post-processing: one C call enqueues the whole chain
//   K1 streaming detect  ->  per-frame TAIL (key sort + integral refinement + node grouping +
//   candidate enumeration + PAF line scores + per-edge assignment + instance assembly)
// with fixed-capacity tables, no host synchronisation and no allocation.  It replaces the call
// sequence find_local_peaks -> peaks * stride -> per-sample split -> PAFScorer.predict of
// BottomUpLayer (layers/bottomup.py:95-236) + group_scored_batch (inference/streaming.py:147-255).
//
// The tail touches O(#peaks) data per frame and is latency-bound, so it runs as ONE CTA per frame
// with every table in shared memory (a handful of dependent memory round trips instead of five
// kernel launches with global-memory tables).  When a second (high-priority) stream is given the
// tail of batch i overlaps the detect pass of batch i+1: the detect kernel is the only part that
// moves real bytes, so the step time tends to the HBM time of the confidence maps.
// If the capacities do not fit in shared memory the stand-alone kernels are chained instead.
#include <cooperative_groups.h>

#include "paf_device.cuh"

namespace cg = cooperative_groups;

namespace snb {

int detect_launch(const void* cms, int dtype, int B, int C, int H, int W, long long sb, long long sc, long long sh,
                  long long sw, float threshold, int cap, int* frame_count, uint32_t* keys, void* ev_begin, void* ev_end,
                  void* stream_, bool zero_counters);  // peaks.cu

struct TailLayout {
  int keys, xy, val, chan, ns, cursor, np, eo, mo, score, m_edge, m_src, m_dst, m_score, lsap, owner, order, idc, idr,
      flags, edges, sorted, ttab, total;
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline TailLayout tail_layout(int peak_cap, int n_nodes, int n_edges, int cand_cap, int match_cap,
                                                  int n_warps, int n_sorted, int n_points) {
  TailLayout L;
  int p2 = 1;
  while (p2 < peak_cap) p2 <<= 1;
  int o = 0;
  auto take = [&](int bytes) { const int at = o; o += align16(bytes); return at; };
  L.keys = take(4 * p2);
  L.xy = take(8 * peak_cap);
  L.val = take(4 * peak_cap);
  L.chan = take(4 * peak_cap);
  L.ns = take(4 * (n_nodes + 1));
  L.cursor = take(4 * (n_nodes + 1));
  L.np = take(4 * peak_cap);
  L.eo = take(4 * (n_edges + 1));
  L.mo = take(4 * (n_edges + 1));
  L.score = take(4 * cand_cap);
  L.m_edge = take(4 * match_cap);
  L.m_src = take(4 * match_cap);
  L.m_dst = take(4 * match_cap);
  L.m_score = take(4 * match_cap);
  L.lsap = take(n_warps * (int)lsap_ws_bytes(32));
  L.owner = take(4 * peak_cap);
  L.order = take(4 * peak_cap);
  L.idc = take(4 * peak_cap);
  L.idr = take(4 * peak_cap);
  L.flags = take(2 * n_nodes);
  L.edges = take(8 * (n_edges > 0 ? n_edges : 1));
  L.sorted = take(4 * (n_sorted > 0 ? n_sorted : 1));
  L.ttab = take(4 * (n_points > 0 ? n_points : 1));
  L.total = o;
  return L;
}

constexpr int TAIL_THREADS = 256;
constexpr int SNB_TAIL_CLUSTER = 4;              // CTAs per frame in the small-batch regime
constexpr int SNB_TAIL_CLUSTER_MAX_FRAMES = 16;  // ... used up to this batch size
constexpr int SNB_TAIL_CLUSTER_MIN_EDGES = 8;    // ... and from this many skeleton edges on

// Profiling build only (-DSNB_TAIL_TIMING, tools/tail_phases.py): thread 0 of every CTA stamps clock64() at
// the phase boundaries into asm_ws (unused by the fused tail), 16 ints per frame.
#ifdef SNB_TAIL_TIMING
#define SNB_STAMP(k)                                                                       \
  do {                                                                                     \
    if (threadIdx.x == 0 && R == 0 && a.asm_ws) a.asm_ws[b * 16 + (k)] = (int)(clock64() - t_start); \
  } while (0)
#else
#define SNB_STAMP(k) do {} while (0)
#endif

// CS = CTAs per frame.  CS = 1: one CTA owns a frame (the pipelined regime: with a batch of 64 the tail hides under the
// next batch's detect pass and more CTAs per frame would only take SMs away from it).  CS = 4 (small batches, where one
// CTA per frame leaves most of the GPU idle and the tail's LATENCY is what the caller waits for): a thread-block cluster
// per frame.  Every CTA of the cluster keeps the frame's full table set in its own shared memory; the cheap, sequential
// steps (key sort, grouping by node) are computed redundantly by each CTA, the expensive parallel ones are split -
// refinement over peaks (results broadcast to all CTAs through distributed shared memory), line scores over candidates
// (each score stored into the shared memory of the CTA that owns the candidate's edge), assignments over edges (matches
// stored into CTA 0) - and CTA 0 runs the assembly.  Three cluster barriers in total.
template <int CS>
__global__ void __launch_bounds__(TAIL_THREADS) bottomup_tail_kernel(snb_bottomup_args a) {
  extern __shared__ __align__(16) unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int R = CS > 1 ? (int)cluster.block_rank() : 0;  // this CTA's rank inside the frame's cluster
  // peer(p, r): the same shared-memory location in CTA r of the cluster (CS == 1: the local one)
  auto peer = [&](auto* p, int r) { return CS > 1 ? cluster.map_shared_rank(p, (unsigned)r) : p; };
  auto cluster_sync = [&]() { if (CS > 1) cluster.sync(); else __syncthreads(); };
  const int n_warps = TAIL_THREADS / 32;
  const TailLayout L = tail_layout(a.peak_cap, a.C, a.n_edges, a.cand_cap, a.match_cap, n_warps, a.n_sorted, a.n_points);
  uint32_t* s_keys = (uint32_t*)(smem + L.keys);
  float* s_xy = (float*)(smem + L.xy);
  float* s_val = (float*)(smem + L.val);
  int* s_chan = (int*)(smem + L.chan);
  int* s_ns = (int*)(smem + L.ns);
  int* s_cursor = (int*)(smem + L.cursor);
  int* s_np = (int*)(smem + L.np);
  int* s_eo = (int*)(smem + L.eo);
  int* s_mo = (int*)(smem + L.mo);
  float* s_score = (float*)(smem + L.score);
  int* s_m_edge = (int*)(smem + L.m_edge);
  int* s_m_src = (int*)(smem + L.m_src);
  int* s_m_dst = (int*)(smem + L.m_dst);
  float* s_m_score = (float*)(smem + L.m_score);

  const int b = blockIdx.x / CS, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_nodes = a.C, E = a.n_edges;
#ifdef SNB_TAIL_TIMING
  const long long t_start = clock64();
  unsigned long long g_start;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_start));
#endif
  // small read-only tables -> shared memory once (the sequential phases would otherwise pay an L2
  // round trip per dependent access)
  int* s_edges = (int*)(smem + L.edges);
  int* s_sorted = (int*)(smem + L.sorted);
  float* s_t = (float*)(smem + L.ttab);
  for (int i = tid; i < 2 * E; i += TAIL_THREADS) s_edges[i] = a.edges[i];
  for (int i = tid; i < a.n_sorted; i += TAIL_THREADS) s_sorted[i] = a.sorted_edges[i];
  for (int i = tid; i < a.n_points; i += TAIL_THREADS) s_t[i] = a.t_table[i];
  const int total = a.frame_count[b];
  if (CS > 1) cluster.sync();  // every CTA of the cluster is running before anyone touches a peer's shared memory
  if (total > a.peak_cap && tid == 0 && R == 0) atomicOr(a.status, SNB_STATUS_PEAK_OVERFLOW);
  const int n = min(total, a.peak_cap);

  // ---- 1. order the frame's keys: ascending key == (y, x, channel) == torch.where order
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  const uint32_t* gk = a.keys + (long long)b * a.peak_cap;
  for (int i = tid; i < n2; i += TAIL_THREADS) s_keys[i] = (i < n) ? gk[i] : 0xffffffffu;
  __syncthreads();
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n2; i += TAIL_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint32_t x = s_keys[i], y = s_keys[ixj];
          if ((x > y) == ((i & k) == 0)) { s_keys[i] = y; s_keys[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }

  SNB_STAMP(0);
  // ---- 2. value + integral refinement, one warp per peak (taps fetched in parallel)
  const int cdt = a.cms_dtype;
  const void* frame = elem_ptr(a.cms, (long long)b * a.cms_sb, cdt);
  constexpr int RQ = 4;  // peaks refined together by one warp (their taps are all in flight at once)
  for (int i0 = (R * n_warps + warp) * RQ; i0 < n; i0 += CS * n_warps * RQ) {
    const void* plane[RQ];
    float fx[RQ], fy[RQ], ox[RQ], oy[RQ];
    int cc[RQ], xi[RQ], yi[RQ];
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
      const int i = i0 + q;
      plane[q] = nullptr; fx[q] = fy[q] = 0.f; cc[q] = xi[q] = yi[q] = 0;
      if (i < n) {
        const uint32_t key = s_keys[i];
        cc[q] = (int)(key % (uint32_t)a.C);
        const uint32_t yx = key / (uint32_t)a.C;
        xi[q] = (int)(yx % (uint32_t)a.W); yi[q] = (int)(yx / (uint32_t)a.W);
        plane[q] = elem_ptr(frame, (long long)cc[q] * a.cms_sc, cdt);
        fx[q] = (float)xi[q]; fy[q] = (float)yi[q];
      }
    }
    if (a.refine_size > 0) integral_refine_warp_multi<RQ>(plane, cdt, a.H, a.W, a.cms_sh, a.cms_sw, fx, fy, a.refine_size, lane, ox, oy);
    if (lane < RQ) {  // lane q finishes peak i0 + q
#pragma unroll
      for (int q = 0; q < RQ; ++q) {
        if (lane != q || i0 + q >= n) continue;
        float x = fx[q], y = fy[q];
        if (a.refine_size > 0) { x = __fadd_rn(x, ox[q]); y = __fadd_rn(y, oy[q]); }
        if (a.cms_stride != 1.0f) { x = __fmul_rn(x, a.cms_stride); y = __fmul_rn(y, a.cms_stride); }
        const int i = i0 + q;
        const float v = ld_elem(plane[q], (long long)yi[q] * a.cms_sh + (long long)xi[q] * a.cms_sw, cdt);
#pragma unroll
        for (int r = 0; r < CS; ++r) {  // every CTA of the cluster gets the refined peak
          float* pxy = peer(s_xy, r);
          pxy[2 * i] = x; pxy[2 * i + 1] = y; peer(s_val, r)[i] = v; peer(s_chan, r)[i] = cc[q];
        }
        const long long o = (long long)b * a.peak_cap + i;
        a.peak_xy[2 * o] = x; a.peak_xy[2 * o + 1] = y; a.peak_val[o] = v; a.peak_chan[o] = cc[q];
      }
    }
  }
  cluster_sync();

  SNB_STAMP(1);
  // ---- 3. group peaks by node, candidate / match offsets
  if (warp == 0) {
    group_by_node_warp(s_chan, n, n_nodes, s_ns, s_cursor, s_np, lane);
    if (lane == 0) edge_offsets(s_edges, n_nodes, E, s_ns, s_eo, s_mo);
  }
  __syncthreads();
  if (a.max_peaks_per_node > 0 && a.skip_flag && R == 0) {  // layers/bottomup.py:128-148: batch-wide guard
    bool over = false;
    for (int k = tid; k < n_nodes; k += TAIL_THREADS) over = over || (s_ns[k + 1] - s_ns[k] > a.max_peaks_per_node);
    if (over) atomicOr(a.skip_flag, 1);
  }
  const int M = s_eo[E];
  const bool cand_ok = M <= a.cand_cap;
  if (!cand_ok && tid == 0 && R == 0) atomicOr(a.status, SNB_STATUS_CAND_OVERFLOW);
  const int K = s_mo[E];
  const int klimit = min(K, a.match_cap);
  if (K > klimit && tid == 0 && R == 0) atomicOr(a.status, SNB_STATUS_MATCH_OVERFLOW);
  if (a.node_start && R == 0) {  // optional copies of the intermediate tables (the API-level 6-tuple)
    for (int i = tid; i <= n_nodes; i += TAIL_THREADS) a.node_start[(long long)b * (n_nodes + 1) + i] = s_ns[i];
    for (int i = tid; i < n; i += TAIL_THREADS) a.node_peaks[(long long)b * a.peak_cap + i] = s_np[i];
    for (int i = tid; i <= E; i += TAIL_THREADS) {
      a.edge_off[(long long)b * (E + 1) + i] = s_eo[i];
      a.match_off[(long long)b * (E + 1) + i] = s_mo[i];
    }
  }

  SNB_STAMP(2);
  // ---- 4. PAF line scores, one thread per candidate
  if (cand_ok) {
    ScoreArgs sa{a.pafs, a.pafs_dtype, a.paf_sb, a.paf_sy, a.paf_sx, a.paf_sc, a.paf_H, a.paf_W, s_t, a.n_points,
                 a.pafs_stride, a.max_edge_length, a.dist_penalty_weight};
    for (int m = R * TAIL_THREADS + tid; m < M; m += CS * TAIL_THREADS) {
      int k, ps, pd;
      decode_candidate(m, s_eo, E, s_ns, s_edges, s_np, &k, &ps, &pd);
      const float sc = score_candidate(sa, b, k, s_xy[2 * ps], s_xy[2 * ps + 1], s_xy[2 * pd], s_xy[2 * pd + 1]);
      peer(s_score, k % CS)[m] = sc;  // into the CTA that solves edge k's assignment
      if (a.cand_edge) {
        const long long o = (long long)b * a.cand_cap + m;
        a.cand_edge[o] = k; a.cand_epi[2 * o] = ps; a.cand_epi[2 * o + 1] = pd; a.cand_score[o] = sc;
      }
    }
  }
  cluster_sync();

  SNB_STAMP(3);
  // ---- 5. per-edge optimal assignment, one warp per edge (scipy's algorithm; the scan over free columns runs
  //         one column per lane when the problem fits 32 x 32, else lane 0 solves it serially in global scratch)
  if (cand_ok) {
    int* r0_m_edge = peer(s_m_edge, 0);
    int* r0_m_src = peer(s_m_src, 0);
    int* r0_m_dst = peer(s_m_dst, 0);
    float* r0_m_score = peer(s_m_score, 0);
    for (int k = R + CS * warp; k < E; k += CS * n_warps) {  // edge k belongs to CTA k % CS
      const int s = s_edges[2 * k], d = s_edges[2 * k + 1];
      if (s < 0 || s >= n_nodes || d < 0 || d >= n_nodes) continue;
      const int n_src = s_ns[s + 1] - s_ns[s], n_dst = s_ns[d + 1] - s_ns[d];
      const int n_match = min(n_src, n_dst);
      if (n_match == 0 || s_mo[k] + n_match > klimit) continue;
      const int dim = max(n_src, n_dst);
      const float* sc = s_score + s_eo[k];
      auto cost = [&](int i, int j) -> double {
        const float x = sc[i * n_dst + j];
        return isnan(x) ? (double)INFINITY : -(double)x;
      };
      const int o = s_mo[k];
      bool ok = true;
      if (dim <= 32) {
        ok = lsap_solve_warp(n_src, n_dst, cost, smem + L.lsap + warp * (int)lsap_ws_bytes(32), s_m_src + o, s_m_dst + o, lane);
      } else if (dim <= a.lsap_max_dim && a.lsap_ws) {
        if (lane == 0)
          ok = lsap_solve(n_src, n_dst, cost, (unsigned char*)a.lsap_ws + ((size_t)b * E + k) * lsap_ws_bytes(a.lsap_max_dim),
                          s_m_src + o, s_m_dst + o);
        ok = __shfl_sync(FULL, ok ? 1 : 0, 0) != 0;
        __syncwarp();
      } else {
        if (lane == 0) atomicOr(a.status, SNB_STATUS_LSAP_TOO_LARGE);
        continue;
      }
      if (!ok) {
        if (lane == 0) atomicOr(a.status, SNB_STATUS_LSAP_INFEASIBLE);
        for (int r = lane; r < n_match; r += 32) { r0_m_edge[o + r] = k; r0_m_src[o + r] = -1; r0_m_dst[o + r] = -1; r0_m_score[o + r] = NAN; }
        continue;
      }
      __syncwarp();
      for (int r = lane; r < n_match; r += 32) {  // the solver wrote (src, dst) locally; CTA 0 gets the finished match rows
        const int ms = s_m_src[o + r], md = s_m_dst[o + r];
        r0_m_edge[o + r] = k;
        r0_m_score[o + r] = sc[ms * n_dst + md];
        if (CS > 1) { r0_m_src[o + r] = ms; r0_m_dst[o + r] = md; }
      }
    }
  }
  cluster_sync();
  if (R != 0) return;  // nobody reads this CTA's shared memory any more; CTA 0 finishes the frame alone
  const int n_matches = cand_ok ? klimit : 0;
  if (a.m_edge) {
    for (int i = tid; i < n_matches; i += TAIL_THREADS) {
      const long long o = (long long)b * a.match_cap + i;
      a.m_edge[o] = s_m_edge[i]; a.m_src[o] = s_m_src[i]; a.m_dst[o] = s_m_dst[i]; a.m_score[o] = s_m_score[i];
    }
    if (tid == 0) a.m_count[b] = n_matches;
  }

  SNB_STAMP(4);
  // ---- 6. assembly: the sequential part (greedy loop, id compaction) by warp 0, the rest by the whole CTA
  __shared__ AsmSplit s_hand;
  if (tid == 0) { s_hand.split = 0; s_hand.pre = 0; s_hand.forest = 0; }
  __syncthreads();
  {
    AsmFrame f;
    f.xy = s_xy; f.val = s_val; f.chan = s_chan; f.P = n;
    f.ns = s_ns; f.np_ = s_np; f.n_nodes = n_nodes;
    f.edges = s_edges; f.sorted = s_sorted; f.n_sorted = a.n_sorted;
    f.m_edge = s_m_edge; f.m_src = s_m_src; f.m_dst = s_m_dst; f.m_score = s_m_score; f.K = n_matches;
    f.mo = s_mo;  // the tail wrote edge k's matches at [s_mo[k], s_mo[k+1])
    f.min_instance_peaks = a.min_instance_peaks; f.min_line_scores = a.min_line_scores;
    f.owner = (int*)(smem + L.owner); f.order = (int*)(smem + L.order);
    f.id_count = (int*)(smem + L.idc); f.id_rank = (int*)(smem + L.idr);
    f.fa = smem + L.flags; f.fb = smem + L.flags + n_nodes;
    f.inst_cap = a.inst_cap;
    f.oxy = a.inst_xy + (long long)b * a.inst_cap * n_nodes * 2;
    f.oval = a.inst_val + (long long)b * a.inst_cap * n_nodes;
    f.osc = a.inst_score + (long long)b * a.inst_cap;
    f.n_inst_out = a.n_inst + b; f.status = a.status;
    f.scratch = reinterpret_cast<int*>(s_score);  // candidate scores are dead once the assignments are solved
    f.scratch_words = a.cand_cap; f.n_edges = E;
    f.proper = true;  // the matches come from this kernel's own assignment solver
#ifdef SNB_TAIL_TIMING
    f.stamps = a.asm_ws ? a.asm_ws + b * 16 : nullptr; f.t0 = t_start;
#endif
    assemble_prepass_cta(f, &s_hand, tid, TAIL_THREADS);
    __syncthreads();
    assemble_forest_cta(f, &s_hand, tid, TAIL_THREADS);
    __syncthreads();
    if (warp == 0) assemble_frame_warp(f, lane, &s_hand);
    __syncthreads();
    assemble_finish_cta(f, &s_hand, tid, TAIL_THREADS);
  }
  SNB_STAMP(5);
#ifdef SNB_TAIL_TIMING
  if (threadIdx.x == 0 && R == 0 && a.asm_ws) {  // wall time of the same span in ns (the SM clock is not fixed)
    unsigned long long g_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
    a.asm_ws[b * 16 + 10] = (int)(g_end - g_start);
  }
#endif
  // self-resetting peak counter: every thread read `total` before the first barrier above, so the counter can go back
  // to zero for the next call's detect kernel (no memset node in the chain)
  if ((a.flags & SNB_FLAG_SELF_RESET_COUNTERS) && tid == 0) {
    a.n_peaks[b] = total;
    a.frame_count[b] = 0;
  }
}

// Padded per-frame instance tables -> packed rows appended at a DEVICE-side running offset, so that a
// rank can accumulate the results of many batches with no host synchronisation and gather them once
// at the end of its frame shard (sleap_nn_b200/sharding.py).  One CTA per frame; the frame's offset
// is cursor[0] (rows already packed by earlier calls) + the prefix of the earlier frames of this call.
// The LAST CTA to finish advances cursor[0] by the call's total and cursor[1] by B (frames packed).
__global__ void __launch_bounds__(128)
pack_instances_kernel(const int* __restrict__ n_inst, int B, int inst_cap, int n_nodes, const float* __restrict__ xy,
                      const float* __restrict__ val, const float* __restrict__ score, int frame_base,
                      unsigned long long* __restrict__ cursor, unsigned* __restrict__ ticket, long long out_cap,
                      float* __restrict__ o_xy, float* __restrict__ o_val, float* __restrict__ o_score,
                      int* __restrict__ o_frame, int* __restrict__ o_count, int* __restrict__ status) {
  const int b = blockIdx.x;
  __shared__ long long s_off;
  __shared__ int s_total;
  if (threadIdx.x < 32) {
    int before = 0, total = 0;
    for (int i = threadIdx.x; i < B; i += 32) {
      const int c = min(n_inst[i], inst_cap);
      total += c;
      if (i < b) before += c;
    }
    for (int d = 16; d > 0; d >>= 1) {
      before += __shfl_xor_sync(FULL, before, d);
      total += __shfl_xor_sync(FULL, total, d);
    }
    if (threadIdx.x == 0) {
      s_off = (long long)cursor[0] + before;
      s_total = total;
    }
  }
  __syncthreads();
  const int n = min(n_inst[b], inst_cap);
  const long long off = s_off;
  if (threadIdx.x == 0 && o_count) o_count[(long long)cursor[1] + b] = n;
  if (off + n > out_cap) {
    if (threadIdx.x == 0) atomicOr(status, SNB_STATUS_INSTANCE_OVERFLOW);
  } else {
    const long long src = (long long)b * inst_cap;
    for (int i = threadIdx.x; i < n * n_nodes; i += blockDim.x) {
      o_xy[2 * (off * n_nodes + i)] = xy[2 * (src * n_nodes + i)];
      o_xy[2 * (off * n_nodes + i) + 1] = xy[2 * (src * n_nodes + i) + 1];
      o_val[off * n_nodes + i] = val[src * n_nodes + i];
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      o_score[off + i] = score[src + i];
      o_frame[off + i] = frame_base + b;
    }
  }
  // every CTA has read cursor[] before the last one updates it
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == (unsigned)(B - 1)) {
      *ticket = 0;
      cursor[0] += (unsigned long long)s_total;
      cursor[1] += (unsigned long long)B;
      __threadfence();
    }
  }
}

// Batch-wide max_peaks_per_node guard for the unfused chain (the fused tail checks it in place).
__global__ void node_guard_kernel(const int* __restrict__ node_start, int B, int n_nodes, int limit, int* skip_flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n_nodes) return;
  const int b = i / n_nodes, k = i % n_nodes;
  const int* ns = node_start + (long long)b * (n_nodes + 1);
  if (ns[k + 1] - ns[k] > limit) atomicOr(skip_flag, 1);
}

// group_scored_batch's epilogue (inference/streaming.py:196-243): top-N by score when truncating,
// undo input / effective scale, NaN-pad to (B, max_instances, N, ...).  One CTA per frame.
__global__ void __launch_bounds__(128)
bottomup_outputs_kernel(const int* __restrict__ n_inst, const float* __restrict__ inst_xy,
                        const float* __restrict__ inst_val, const float* __restrict__ inst_score, int inst_cap,
                        int n_nodes, int max_instances, float input_scale, const float* __restrict__ eff_scale,
                        const int* __restrict__ skip_flag, float* __restrict__ out_kpts, float* __restrict__ out_vals,
                        float* __restrict__ out_scores) {
  extern __shared__ int s_src[];  // s_src[r] = source row of output row r
  const int b = blockIdx.x;
  const bool skip = skip_flag && *skip_flag;
  const int n = skip ? 0 : min(n_inst[b], inst_cap);
  const int keep = min(n, max_instances);
  const float* sc = inst_score + (long long)b * inst_cap;
  if (n > max_instances) {
    // position of row i in np.argsort(scores)[::-1]: NaN first, then descending; equal -> higher index first
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float si = sc[i];
      const bool ni = isnan(si);
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        if (j == i) continue;
        const float sj = sc[j];
        const bool nj = isnan(sj);
        bool before;  // does j come before i?
        if (ni || nj) before = (nj && !ni) || (nj && ni && j > i);
        else before = (sj > si) || (sj == si && j > i);
        rank += before ? 1 : 0;
      }
      if (rank < max_instances) s_src[rank] = i;
    }
  } else {
    for (int i = threadIdx.x; i < keep; i += blockDim.x) s_src[i] = i;
  }
  __syncthreads();
  const float eff = eff_scale ? eff_scale[b] : 1.0f;
  const long long ob = (long long)b * max_instances;
  for (int t = threadIdx.x; t < max_instances * n_nodes; t += blockDim.x) {
    const int r = t / n_nodes, k = t - r * n_nodes;
    float x = NAN, y = NAN, v = NAN;
    if (r < keep) {
      const long long s = ((long long)b * inst_cap + s_src[r]) * n_nodes + k;
      x = __fdiv_rn(__fdiv_rn(inst_xy[2 * s], input_scale), eff);
      y = __fdiv_rn(__fdiv_rn(inst_xy[2 * s + 1], input_scale), eff);
      v = inst_val[s];
    }
    out_kpts[2 * (ob * n_nodes + t)] = x;
    out_kpts[2 * (ob * n_nodes + t) + 1] = y;
    out_vals[ob * n_nodes + t] = v;
  }
  for (int r = threadIdx.x; r < max_instances; r += blockDim.x)
    out_scores[ob + r] = (r < keep) ? sc[s_src[r]] : NAN;
}

}  // namespace snb

using namespace snb;


extern "C" long long snb_bottomup_tail_smem_bytes(int peak_cap, int n_nodes, int n_edges, int cand_cap, int match_cap,
                                                  int n_sorted, int n_points) {
  return tail_layout(peak_cap, n_nodes, n_edges, cand_cap, match_cap, TAIL_THREADS / 32, n_sorted, n_points).total;
}

static int unfused_tail(const snb_bottomup_args* a, void* stream);

extern "C" int snb_bottomup_postproc(const snb_bottomup_args* a, void* stream) {
  if (!a) return SNB_ERR_BAD_ARG;
  if (a->B < 0 || a->C <= 0 || a->peak_cap <= 0 || a->cand_cap <= 0 || a->match_cap <= 0 || a->inst_cap <= 0)
    return SNB_ERR_BAD_ARG;
  if (a->B == 0) return SNB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cudaStream_t tail_st = a->tail_stream ? (cudaStream_t)a->tail_stream : st;
  // buffers of this pipeline instance may still be read by its previous tail
  if (a->tail_stream && a->ev_tail_done) cudaStreamWaitEvent(st, (cudaEvent_t)a->ev_tail_done, 0);
  if (a->skip_flag && cudaMemsetAsync(a->skip_flag, 0, sizeof(int), st) != cudaSuccess) return SNB_ERR_CUDA_LAUNCH;
  const long long smem = snb_bottomup_tail_smem_bytes(a->peak_cap, a->C, a->n_edges, a->cand_cap, a->match_cap, a->n_sorted,
                                                      a->n_points);
  const bool fused = !(a->flags & SNB_FLAG_UNFUSED_TAIL) && smem <= 200 * 1024;
  const bool self_reset = (a->flags & SNB_FLAG_SELF_RESET_COUNTERS) != 0;
  if (self_reset && (!fused || !a->n_peaks)) return SNB_ERR_BAD_ARG;  // only the fused tail resets the counters
  int rc = detect_launch(a->cms, a->cms_dtype, a->B, a->C, a->H, a->W, a->cms_sb, a->cms_sc, a->cms_sh, a->cms_sw,
                         a->peak_threshold, a->peak_cap, a->frame_count, a->keys, a->ev_detect_begin, a->ev_detect_end,
                         stream, !self_reset);
  if (rc != SNB_OK) return rc;
  if (a->tail_stream) {
    if (!a->ev_handoff) return SNB_ERR_BAD_ARG;
    cudaEventRecord((cudaEvent_t)a->ev_handoff, st);
    cudaStreamWaitEvent(tail_st, (cudaEvent_t)a->ev_handoff, 0);
  }
  if (fused) {
    // Small batches: a 4-CTA cluster per frame (see bottomup_tail_kernel).  From 17 frames on, one CTA per frame - the
    // regime where tails hide under the next batch's detect pass and extra CTAs would only take SMs away from it.
    // ... and only for skeletons with enough edges to split: measured on B200 (tools/latency_small_batch.py), batch 8:
    // 32 nodes / 31 edges with 8 animals 143 -> 85 us, but 5 nodes / 4 edges with 2 animals 35 -> 43 us (three cluster
    // barriers cost more than four edges' worth of parallelism returns).
    const bool use_cluster = a->B <= SNB_TAIL_CLUSTER_MAX_FRAMES && a->n_edges >= SNB_TAIL_CLUSTER_MIN_EDGES &&
                             !(a->flags & SNB_FLAG_NO_TAIL_CLUSTER);
    if (use_cluster) {
      if (smem > 48 * 1024 && cudaFuncSetAttribute(bottomup_tail_kernel<SNB_TAIL_CLUSTER>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return SNB_ERR_CUDA_LAUNCH;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)a->B * SNB_TAIL_CLUSTER);
      cfg.blockDim = dim3(TAIL_THREADS);
      cfg.dynamicSmemBytes = (size_t)smem;
      cfg.stream = tail_st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = SNB_TAIL_CLUSTER;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (cudaLaunchKernelEx(&cfg, bottomup_tail_kernel<SNB_TAIL_CLUSTER>, *a) != cudaSuccess) return SNB_ERR_CUDA_LAUNCH;
    } else {
      if (smem > 48 * 1024 &&
          cudaFuncSetAttribute(bottomup_tail_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return SNB_ERR_CUDA_LAUNCH;
      bottomup_tail_kernel<1><<<a->B, TAIL_THREADS, (size_t)smem, tail_st>>>(*a);
    }
    SNB_LAUNCH_CHECK();
  } else {
    rc = unfused_tail(a, (void*)tail_st);
    if (rc != SNB_OK) return rc;
  }
  if (a->out_kpts) {
    rc = snb_bottomup_outputs(a->n_inst, a->inst_xy, a->inst_val, a->inst_score, a->B, a->inst_cap, a->C,
                              a->max_instances, a->input_scale, a->eff_scale, a->skip_flag, a->out_kpts, a->out_vals,
                              a->out_scores, (void*)tail_st);
    if (rc != SNB_OK) return rc;
  }
  if (a->tail_stream && a->ev_tail_done) cudaEventRecord((cudaEvent_t)a->ev_tail_done, tail_st);
  return SNB_OK;
}

extern "C" int snb_bottomup_outputs(const int* n_inst, const float* inst_xy, const float* inst_val,
                                    const float* inst_score, int B, int inst_cap, int n_nodes, int max_instances,
                                    float input_scale, const float* eff_scale, const int* skip_flag, float* out_kpts,
                                    float* out_vals, float* out_scores, void* stream) {
  if (B < 0 || inst_cap <= 0 || n_nodes <= 0 || max_instances <= 0) return SNB_ERR_BAD_ARG;
  if (B == 0) return SNB_OK;
  const size_t smem = sizeof(int) * (size_t)max_instances;
  if (smem > 48 * 1024) return SNB_ERR_UNSUPPORTED;
  bottomup_outputs_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(n_inst, inst_xy, inst_val, inst_score, inst_cap, n_nodes,
                                                                 max_instances, input_scale, eff_scale, skip_flag,
                                                                 out_kpts, out_vals, out_scores);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

// Stand-alone kernels chained on one stream (large capacities, or SNB_FLAG_UNFUSED_TAIL for tests).

static int unfused_tail(const snb_bottomup_args* a, void* stream) {
  const int n_nodes = a->C;
  if (!a->node_start || !a->cand_edge || !a->m_edge) return SNB_ERR_BAD_ARG;  // needs the global tables
  int rc = snb_local_peaks_finalize_t(a->cms, a->cms_dtype, a->B, a->C, a->H, a->W, a->cms_sb, a->cms_sc, a->cms_sh,
                                      a->cms_sw, a->refine_size, a->cms_stride, a->peak_cap, a->frame_count, a->keys,
                                      a->peak_xy, a->peak_val, a->peak_chan, a->status, stream);
  if (rc != SNB_OK) return rc;
  rc = snb_paf_prepare(a->peak_chan, nullptr, a->peak_cap, a->frame_count, a->B, a->edges, n_nodes, a->n_edges,
                       a->node_start, a->node_peaks, a->edge_off, a->match_off, stream);
  if (rc != SNB_OK) return rc;
  if (a->max_peaks_per_node > 0 && a->skip_flag) {
    const int n = a->B * n_nodes;
    node_guard_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a->node_start, a->B, n_nodes,
                                                                       a->max_peaks_per_node, a->skip_flag);
    SNB_LAUNCH_CHECK();
  }
  rc = snb_paf_score_t(a->pafs, a->pafs_dtype, a->paf_sb, a->paf_sy, a->paf_sx, a->paf_sc, a->paf_H, a->paf_W, a->t_table, a->n_points,
                     a->pafs_stride, a->max_edge_length, a->dist_penalty_weight, a->peak_xy, nullptr, a->peak_cap,
                     a->B, a->edges, n_nodes, a->n_edges, a->node_start, a->node_peaks, a->edge_off, nullptr,
                     a->cand_cap, a->cand_cap, a->cand_edge, a->cand_epi, a->cand_score, a->status, stream);
  if (rc != SNB_OK) return rc;
  rc = snb_match_structured(a->cand_score, nullptr, a->cand_cap, a->edges, n_nodes, a->n_edges, a->node_start,
                            a->edge_off, a->match_off, nullptr, a->match_cap, a->lsap_ws, a->lsap_max_dim, a->B,
                            a->m_edge, a->m_src, a->m_dst, a->m_score, a->m_count, a->status, stream);
  if (rc != SNB_OK) return rc;
  return snb_assemble(a->peak_xy, a->peak_val, a->peak_chan, nullptr, a->peak_cap, a->frame_count, a->B, a->node_start,
                      a->node_peaks, n_nodes, a->edges, a->sorted_edges, a->n_sorted, a->m_edge, a->m_src, a->m_dst,
                      a->m_score, nullptr, a->match_cap, a->m_count, a->min_instance_peaks, a->min_line_scores,
                      a->asm_ws, a->peak_cap, a->inst_cap, a->inst_xy, a->inst_val, a->inst_score, a->n_inst, a->status,
                      stream);
}

// Kernel launches per call: detect + fused tail (2), or detect + 5 stand-alone kernels (6).
extern "C" int snb_bottomup_launches_per_call(const snb_bottomup_args* a) {
  if (!a) return 2;
  const long long smem = snb_bottomup_tail_smem_bytes(a->peak_cap, a->C, a->n_edges, a->cand_cap, a->match_cap, a->n_sorted,
                                                      a->n_points);
  const bool fused = !(a->flags & SNB_FLAG_UNFUSED_TAIL) && smem <= 200 * 1024;
  return (fused ? 2 : 6 + ((a->max_peaks_per_node > 0 && a->skip_flag) ? 1 : 0)) + (a->out_kpts ? 1 : 0);
}

// Append one batch's instances to a packed per-rank result table (see pack_instances_kernel).
// cursor: 2 x u64 {rows packed, frames packed} + ticket u32, zeroed by the caller before the first call.
extern "C" int snb_pack_instances(const int* n_inst, int B, int inst_cap, int n_nodes, const float* inst_xy,
                                  const float* inst_val, const float* inst_score, int frame_base, void* cursor,
                                  long long out_cap, float* o_xy, float* o_val, float* o_score, int* o_frame,
                                  int* o_count, int* status, void* stream) {
  if (B < 0 || inst_cap <= 0 || n_nodes <= 0 || !cursor) return SNB_ERR_BAD_ARG;
  if (B == 0) return SNB_OK;
  unsigned long long* cur = (unsigned long long*)cursor;
  pack_instances_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(n_inst, B, inst_cap, n_nodes, inst_xy, inst_val, inst_score,
                                                             frame_base, cur, (unsigned*)(cur + 2), out_cap, o_xy, o_val,
                                                             o_score, o_frame, o_count, status);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
