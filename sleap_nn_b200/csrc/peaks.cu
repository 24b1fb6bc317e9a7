// Peak finding on confidence maps: sm_100a kernels + C-ABI entry points.
//
//   K1  snb_local_peaks        fused 3x3 NMS + threshold + peak emission (streaming read of the
//                              maps, 128-bit loads) -> per-frame key sort -> integral refinement
//                              (replaces ops/peaks.py:184-259 + ops/crops.py:31-124)
//   K2  snb_global_peaks       per-(sample,channel) arg-max with the reference's two independent
//                              arg-max semantics + threshold + refinement (ops/peaks.py:89-181)
//   K3  snb_crop_bboxes        integer-aligned zero-padded patch gather (ops/crops.py:31-124)
//       snb_integral_regression, snb_dilate8, snb_pack_peaks
//
// Bound: HBM read bandwidth.  K1 reads every map element exactly once (4*C*H*W bytes per
// frame); everything after the streaming pass touches O(#peaks) data.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace snb {

// ----------------------------------------------------------------------------------------
// K1a: streaming detect.  One warp owns one map row at a time (grid-stride over B*C*H rows);
// lanes stride over the row in float4s.  A lane only leaves the streaming path when one of
// its four values exceeds the threshold (a few pixels per thousand on real maps); it then
// re-reads the 8 neighbours through L1/L2 and applies the reference's test
//     v > threshold  &&  v > every in-image neighbour        (ops/peaks.py:47-63, 209)
// "v > each neighbour" equals "v > max(neighbours)" including NaN behaviour (a NaN neighbour
// makes torch's max NaN and the comparison false) and the -inf padding (out-of-image
// neighbours are skipped: v > -inf holds for every v that passed v > threshold).
// Peaks are appended to the frame's key list; key = (y*W + x)*C + c orders them the way
// torch.where over (B,H,W,C) does (ops/peaks.py:211-217).
// ----------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ bool is_strict_max(const T* __restrict__ plane, int H, int W, long long sh, long long sw,
                                              int y, int x, float v) {
  bool ok = true;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx;
      if ((dy == 0 && dx == 0) || xx < 0 || xx >= W) continue;
      const float nb = Elem<T>::load1(plane + (long long)yy * sh + (long long)xx * sw);
      ok = ok && (v > nb);
    }
  }
  return ok;
}

__device__ __forceinline__ void emit_peak(int* __restrict__ frame_count, uint32_t* __restrict__ keys, int cap,
                                          int b, int C, int W, int c, int y, int x) {
  const int pos = atomicAdd(frame_count + b, 1);
  // unsigned 32-bit math: the host guarantees H*W*C < 2^32, which a signed int would overflow above 2^31
  if (pos < cap) keys[(long long)b * cap + pos] = ((uint32_t)y * (uint32_t)W + (uint32_t)x) * (uint32_t)C + (uint32_t)c;
}

// Rare path of the streaming kernel: one 128-bit word of row y holds a value above the threshold.
template <typename T>
__device__ __noinline__ void detect_word(const T* __restrict__ plane, float thr, int H, int W, long long sh, int y, int x0,
                                         int b, int c, int C, int cap, int* __restrict__ frame_count,
                                         uint32_t* __restrict__ keys) {
  constexpr int PER = Elem<T>::PER16;
  float e[PER];
  Elem<T>::unpack(__ldg(reinterpret_cast<const uint4*>(plane + (long long)y * sh + x0)), e);  // an L2 hit
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    // Cheap exact pre-filter before the 8 neighbour loads: the horizontal neighbours that sit in the same 128-bit
    // word are already in registers, and a strict maximum must beat them too (same `v > nb` predicate, so NaN
    // neighbours reject as in the reference).  On a blob's row only the ridge pixel (and at most the word-boundary
    // pixels) goes on to is_strict_max - ~3x fewer L1/L2 neighbour reads on busy maps (cfg4: 256 blobs per frame).
    if (e[k] > thr && (k == 0 || e[k] > e[k - 1]) && (k == PER - 1 || e[k] > e[k + 1])) {
      if (is_strict_max<T>(plane, H, W, sh, 1, y, x0 + k, e[k])) emit_peak(frame_count, keys, cap, b, C, W, c, y, x0 + k);
    }
  }
}

template <typename T, int UNROLL, int ROWS, int MIN_BLOCKS, bool EXACT>
__global__ void __launch_bounds__(256, MIN_BLOCKS)
local_peaks_detect_vec(const T* __restrict__ cms, int C, int H, int W, long long sb, long long sc, long long sh,
                       float thr, int cap, int* __restrict__ frame_count, uint32_t* __restrict__ keys) {
  // grid = (row groups of a plane, C, B): a CTA's 8 warps own 8 * ROWS consecutive rows of ONE plane, so the plane
  // and the row come straight from the block / warp index - no integer division anywhere (with a flat row index the
  // div / mod by H and C was ~100 of a warp's ~165 instructions, and a warp only lives for one 2 KB step).  A warp
  // issues all of its 128-bit loads (ROWS * UNROLL per lane per step) before looking at any value: that is the
  // memory-level parallelism that keeps HBM busy.  A 128-bit load holds PER = 4 fp32 or 8 fp16 / bf16 elements;
  // half-precision maps are compared on their exact fp32 values.
  //
  // The streaming step only records WHICH of its words hold a value above the threshold (one bit each); the words
  // themselves are dead after that.  A lane with a non-zero mask (a few per thousand) hands its hot words to the
  // out-of-line neighbour test, which re-reads them (L2 hits).  With the rare path consuming the loaded registers
  // directly, the half-precision kernels (eight up-cast values per word) went past the 40-register budget of
  // 6 CTAs / SM and ptxas spilled inside the streaming loop (0.52 of the roofline).
  // EXACT: H is a multiple of the CTA's rows and the row a multiple of one step, so the loop carries no bounds test.
  constexpr int PER = Elem<T>::PER16;
  const int lane = lane_id();
  const int c = blockIdx.y, b = blockIdx.z;
  const int y0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS;
  if (!EXACT && y0 >= H) return;
  const T* plane = cms + (long long)b * sb + (long long)c * sc;
  const int WV = W / PER;
  const int n_rows = EXACT ? ROWS : min(ROWS, H - y0);
  const typename Elem<T>::Thr tv = Elem<T>::make_thr(thr);
  const T* rowp[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) rowp[r] = plane + (long long)(y0 + (EXACT ? r : min(r, n_rows - 1))) * sh + PER * lane;
  for (int xv = lane; xv < WV; xv += 32 * UNROLL) {
    uint4 v[ROWS][UNROLL];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (EXACT || xv + 32 * u < WV) v[r][u] = ldg_stream16(rowp[r] + PER * 32 * u);
        else v[r][u] = Elem<T>::neg_inf();
      }
    unsigned hot = 0;
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if ((EXACT || r < n_rows) && Elem<T>::any_gt(v[r][u], tv)) hot |= 1u << (r * UNROLL + u);
    while (hot) {  // rare
      const int k = __ffs(hot) - 1;
      hot &= hot - 1;
      const int r = k / UNROLL, u = k - r * UNROLL;
      detect_word<T>(plane, thr, H, W, sh, y0 + r, PER * (xv + 32 * u), b, c, C, cap, frame_count, keys);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) rowp[r] += PER * 32 * UNROLL;
  }
}

// Generic-stride scalar variant (non-contiguous views, W not a multiple of the vector width, unaligned base).
template <typename T>
__global__ void __launch_bounds__(256)
local_peaks_detect_scalar(const T* __restrict__ cms, int B, int C, int H, int W, long long sb, long long sc,
                          long long sh, long long sw, float thr, int cap, int* __restrict__ frame_count,
                          uint32_t* __restrict__ keys) {
  const long long n = (long long)B * C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    long long r = i / W;
    const int y = (int)(r % H);
    r /= H;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    const T* plane = cms + (long long)b * sb + (long long)c * sc;
    const float v = Elem<T>::load1(plane + (long long)y * sh + (long long)x * sw);
    if (v > thr && is_strict_max<T>(plane, H, W, sh, sw, y, x, v)) emit_peak(frame_count, keys, cap, b, C, W, c, y, x);
  }
}

#ifdef SNB_AB_VARIANTS
// ----------------------------------------------------------------------------------------
// A/B only (-DSNB_AB_VARIANTS, tools/): K1a for contiguous fp32 tensors as a bulk-async ring.  The maps are one flat
// array; each persistent CTA walks it in 16 KB chunks through a ring of shared-memory stages filled by
// cp.async.bulk (1-D TMA) and signalled by mbarriers.  Only a value above the threshold leaves the streaming path.
// The refill of a stage is issued by thread 0 right after the CTA-wide barrier that frees it.
// ----------------------------------------------------------------------------------------
constexpr int BULK_STAGE_FLOATS = 4096;  // 16 KB
constexpr int BULK_STAGES = 4;
constexpr int BULK_THREADS = 128;

__global__ void __launch_bounds__(BULK_THREADS)
local_peaks_detect_bulk(const float* __restrict__ cms, long long n_elems, int C, int H, int W, float thr, int cap,
                        int* __restrict__ frame_count, uint32_t* __restrict__ keys) {
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  float* stage_buf = reinterpret_cast<float*>(bulk_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(bulk_smem + (size_t)BULK_STAGES * BULK_STAGE_FLOATS * 4);
  const int tid = threadIdx.x;
  const long long n_chunks = (n_elems + BULK_STAGE_FLOATS - 1) / BULK_STAGE_FLOATS;
  if (tid == 0) {
    for (int s = 0; s < BULK_STAGES; ++s) mbar_init(full + s, 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](long long chunk, int s) {
    const long long e0 = chunk * BULK_STAGE_FLOATS;
    const uint32_t bytes = (uint32_t)(min((long long)BULK_STAGE_FLOATS, n_elems - e0) * 4);
    mbar_expect_tx(full + s, bytes);
    bulk_g2s(stage_buf + (size_t)s * BULK_STAGE_FLOATS, cms + e0, bytes, full + s);
  };
  if (tid == 0) {
    for (int s = 0; s < BULK_STAGES; ++s) {
      const long long chunk = (long long)blockIdx.x + (long long)s * gridDim.x;
      if (chunk < n_chunks) issue(chunk, s);
    }
  }
  const long long plane_elems = (long long)H * W;
  int it = 0;
  for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++it) {
    const int s = it % BULK_STAGES;
    mbar_wait(full + s, (uint32_t)((it / BULK_STAGES) & 1));
    const long long e0 = chunk * BULK_STAGE_FLOATS;
    const int n4 = (int)(min((long long)BULK_STAGE_FLOATS, n_elems - e0) >> 2);
    const float4* buf = reinterpret_cast<const float4*>(stage_buf + (size_t)s * BULK_STAGE_FLOATS);
#pragma unroll
    for (int u = 0; u < BULK_STAGE_FLOATS / 4 / BULK_THREADS; ++u) {
      const int i4 = tid + u * BULK_THREADS;
      if (i4 >= n4) break;
      const float4 v = buf[i4];
      if ((v.x > thr) || (v.y > thr) || (v.z > thr) || (v.w > thr)) {
        const float e[4] = {v.x, v.y, v.z, v.w};
        const long long ebase = e0 + 4LL * i4;  // W % 4 == 0: the four values share one row
        const long long pc = ebase / plane_elems;
        const int rem = (int)(ebase - pc * plane_elems);
        const int y = rem / W, x0 = rem - y * W;
        const int c = (int)(pc % C), b = (int)(pc / C);
        const float* plane = cms + pc * plane_elems;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (e[k] > thr && is_strict_max<float>(plane, H, W, W, 1, y, x0 + k, e[k]))
            emit_peak(frame_count, keys, cap, b, C, W, c, y, x0 + k);
      }
    }
    __syncthreads();  // every thread is done with stage s: refill it
    if (tid == 0) {
      const long long next = chunk + (long long)BULK_STAGES * gridDim.x;
      if (next < n_chunks) issue(next, s);
    }
  }
}

// A/B only: the same stream as a PRODUCER / CONSUMER bulk-async ring, the way TMA pipelines are meant to be built.
// One elected thread of a dedicated producer warp keeps up to TMA_STAGES 16 KB bulk copies in flight per CTA and
// refills a stage as soon as the four consumer warps have released it (an `empty` mbarrier per stage, one arrival per
// consumer warp) - no CTA-wide barrier anywhere, unlike local_peaks_detect_bulk above, whose single refill thread sits
// behind a __syncthreads() per chunk.  Any element type (the chunk is a flat byte range of a contiguous tensor).
#ifndef SNB_TMA_STAGE_BYTES
#define SNB_TMA_STAGE_BYTES 8192
#endif
#ifndef SNB_TMA_STAGES
#define SNB_TMA_STAGES 8
#endif
constexpr int TMA_STAGE_BYTES = SNB_TMA_STAGE_BYTES;
constexpr int TMA_STAGES = SNB_TMA_STAGES;
constexpr int TMA_CONSUMER_WARPS = 4;
constexpr int TMA_THREADS = 32 * (TMA_CONSUMER_WARPS + 1);

template <typename T>
__global__ void __launch_bounds__(TMA_THREADS)
local_peaks_detect_tma(const T* __restrict__ cms, long long n_elems, int C, int H, int W, float thr, int cap,
                       int* __restrict__ frame_count, uint32_t* __restrict__ keys) {
  constexpr int PER = Elem<T>::PER16;
  constexpr int STAGE_ELEMS = TMA_STAGE_BYTES / (int)sizeof(T);
  extern __shared__ __align__(128) unsigned char tma_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(tma_smem + (size_t)TMA_STAGES * TMA_STAGE_BYTES);
  uint64_t* empty = full + TMA_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n_chunks = (n_elems + STAGE_ELEMS - 1) / STAGE_ELEMS;
  if (tid == 0) {
    for (int s = 0; s < TMA_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, TMA_CONSUMER_WARPS); }
    mbar_fence_init();
  }
  __syncthreads();
  if (warp == TMA_CONSUMER_WARPS) {  // ===== producer warp: one elected thread issues every bulk copy of this CTA
    if (lane == 0) {
      int it = 0;
      for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++it) {
        const int s = it % TMA_STAGES;
        if (it >= TMA_STAGES) mbar_wait(empty + s, (uint32_t)(((it / TMA_STAGES) - 1) & 1));
        const long long e0 = chunk * STAGE_ELEMS;
        const uint32_t bytes = (uint32_t)(min((long long)STAGE_ELEMS, n_elems - e0) * (long long)sizeof(T));
        mbar_expect_tx(full + s, bytes);
        bulk_g2s(tma_smem + (size_t)s * TMA_STAGE_BYTES, cms + e0, bytes, full + s);
      }
    }
    return;
  }
  // ===== consumer warps: each owns a quarter of every stage
  const typename Elem<T>::Thr tv = Elem<T>::make_thr(thr);
  const long long plane_elems = (long long)H * W;
  constexpr int VEC_PER_WARP = TMA_STAGE_BYTES / 16 / TMA_CONSUMER_WARPS;  // 256 x 16 B
  int it = 0;
  for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++it) {
    const int s = it % TMA_STAGES;
    mbar_wait(full + s, (uint32_t)((it / TMA_STAGES) & 1));
    const long long e0 = chunk * STAGE_ELEMS;
    const int nv = (int)(min((long long)STAGE_ELEMS, n_elems - e0) / PER);
    const uint4* buf = reinterpret_cast<const uint4*>(tma_smem + (size_t)s * TMA_STAGE_BYTES);
#pragma unroll
    for (int u = 0; u < VEC_PER_WARP / 32; ++u) {
      const int iv = warp * VEC_PER_WARP + u * 32 + lane;
      if (iv >= nv) break;
      const uint4 v = buf[iv];
      if (Elem<T>::any_gt(v, tv)) {  // rare: locate the word in (b, c, y, x) and run the neighbour test
        const long long ebase = e0 + (long long)PER * iv;  // the row is a multiple of the vector: one row per word
        const long long pc = ebase / plane_elems;
        const int rem = (int)(ebase - pc * plane_elems);
        const int y = rem / W, x0 = rem - y * W;
        detect_word<T>(cms + pc * plane_elems, thr, H, W, W, y, x0, (int)(pc / C), (int)(pc % C), C, cap, frame_count, keys);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);  // this warp is done with the stage
  }
}
#endif  // SNB_AB_VARIANTS

// ----------------------------------------------------------------------------------------
// K1b: per-frame finalize.  One CTA per frame: bitonic-sort the frame's keys in shared
// memory (ascending key == (y, x, c) order), then one thread per peak decodes the key,
// re-reads the value, runs integral refinement on the size x size patch and writes the
// frame's slot of the padded peak table.  `n_sort` = power of two >= cap.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void bitonic_sort_smem(uint32_t* s, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint32_t a = s[i], b = s[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) {
            s[i] = b;
            s[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256)
local_peaks_finalize(const void* __restrict__ cms, int dt, int C, int H, int W, long long sb, long long sc, long long sh,
                     long long sw, int refine_size, float xy_scale, int cap, int keys_presorted,
                     const int* __restrict__ frame_count, uint32_t* __restrict__ keys, float* __restrict__ out_xy,
                     float* __restrict__ out_val, int* __restrict__ out_chan, int* __restrict__ status) {
  extern __shared__ uint32_t skeys[];
  const int b = blockIdx.x;
  const int total = frame_count[b];
  if (total > cap && threadIdx.x == 0) atomicOr(status, SNB_STATUS_PEAK_OVERFLOW);
  const int n = min(total, cap);
  if (n == 0) return;
  uint32_t* gk = keys + (long long)b * cap;
  const uint32_t* sorted = gk;
  if (!keys_presorted) {
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) skeys[i] = (i < n) ? gk[i] : 0xffffffffu;
    __syncthreads();
    bitonic_sort_smem(skeys, n2);
    for (int i = threadIdx.x; i < n; i += blockDim.x) gk[i] = skeys[i];  // keep sorted keys for callers
    sorted = skeys;
  }
  const void* frame = elem_ptr(cms, (long long)b * sb, dt);
  const int lane = lane_id(), warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  for (int i = warp; i < n; i += n_warps) {  // one warp per peak: the patch taps are fetched in parallel
    const uint32_t key = sorted[i];
    const int c = (int)(key % (uint32_t)C);
    const uint32_t yx = key / (uint32_t)C;
    const int x = (int)(yx % (uint32_t)W);
    const int y = (int)(yx / (uint32_t)W);
    const void* plane = elem_ptr(frame, (long long)c * sc, dt);
    float fx = (float)x, fy = (float)y;
    if (refine_size > 0) {
      float ox, oy;
      integral_refine_warp(plane, dt, H, W, sh, sw, fx, fy, refine_size, lane, &ox, &oy);
      fx = __fadd_rn(fx, ox);  // ops/peaks.py:258
      fy = __fadd_rn(fy, oy);
    }
    if (xy_scale != 1.0f) {  // layers/bottomup.py:111  peaks * cms_output_stride
      fx = __fmul_rn(fx, xy_scale);
      fy = __fmul_rn(fy, xy_scale);
    }
    if (lane == 0) {
      const long long o = (long long)b * cap + i;
      out_xy[2 * o] = fx;
      out_xy[2 * o + 1] = fy;
      out_val[o] = ld_elem(plane, (long long)y * sh + (long long)x * sw, dt);
      out_chan[o] = c;
    }
  }
}

// Global-memory bitonic step for frames whose key list does not fit in shared memory
// (pathological inputs such as pure noise).  Segments are `cap` (power of two) long.
__global__ void pad_keys_global(uint32_t* keys, const int* frame_count, int cap, int B) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * cap) return;
  const int b = (int)(i / cap);
  if ((int)(i % cap) >= min(frame_count[b], cap)) keys[i] = 0xffffffffu;
}
__global__ void bitonic_step_global(uint32_t* keys, int cap, long long total, int j, int k) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int li = (int)(i % cap);
  const int lx = li ^ j;
  if (lx > li) {
    uint32_t* seg = keys + (i - li);
    const uint32_t a = seg[li], b = seg[lx];
    const bool up = ((li & k) == 0);
    if ((a > b) == up) {
      seg[li] = b;
      seg[lx] = a;
    }
  }
}

// Padded per-frame peak table -> the reference's concatenated layout (points, vals,
// sample_inds, channel_inds), frame order preserved (ops/peaks.py:213-217).
__global__ void pack_peaks(const int* __restrict__ frame_count, int B, int cap, const float* __restrict__ xy,
                           const float* __restrict__ val, const int* __restrict__ chan, float* __restrict__ o_xy,
                           float* __restrict__ o_val, int* __restrict__ o_sample, int* __restrict__ o_chan) {
  const int b = blockIdx.x;
  __shared__ int s_off;
  if (threadIdx.x < 32) {
    int acc = 0;
    for (int i = threadIdx.x; i < b; i += 32) acc += min(frame_count[i], cap);
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(FULL, acc, d);
    if (threadIdx.x == 0) s_off = acc;
  }
  __syncthreads();
  const int n = min(frame_count[b], cap);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const long long s = (long long)b * cap + i, d = s_off + i;
    o_xy[2 * d] = xy[2 * s];
    o_xy[2 * d + 1] = xy[2 * s + 1];
    o_val[d] = val[s];
    o_sample[d] = b;
    o_chan[d] = chan[s];
  }
}

// ----------------------------------------------------------------------------------------
// K2: global peaks.  Reference semantics (ops/peaks.py:103-111): x = first column whose
// column-max equals the plane max, y = first row whose row-max equals it; NaN propagates as
// the greatest value.  That is the associative reduction of (v, x, y) with
//     a beats b  if a.v is NaN and b.v is not, or a.v > b.v;   on equality (or both NaN)
//     x = min(x), y = min(y)
// so any tree order gives the reference's answer.  Planes are split into row chunks; the
// last CTA to finish a plane combines the partials, applies the threshold (NaN coords, 0
// value when max < thr, ops/peaks.py:121-129) and the integral refinement.
// ----------------------------------------------------------------------------------------
struct Best {
  float v;
  int x, y;
};
__device__ __forceinline__ Best best_merge(Best a, Best b) {
  const bool an = isnan(a.v), bn = isnan(b.v);
  if ((an && bn) || a.v == b.v) return Best{a.v, min(a.x, b.x), min(a.y, b.y)};
  if (an) return a;
  if (bn) return b;
  return (a.v > b.v) ? a : b;
}
__device__ __forceinline__ Best best_warp(Best a) {
  for (int d = 16; d > 0; d >>= 1) {
    Best o{__shfl_xor_sync(FULL, a.v, d), __shfl_xor_sync(FULL, a.x, d), __shfl_xor_sync(FULL, a.y, d)};
    a = best_merge(a, o);
  }
  return a;
}

// Output of one (sample, channel): coordinate ladder, optional row scatter (TopDownLayer's valid_idx).
__device__ __forceinline__ void write_global_peak(const Ladder& lad, int plane_id, int C, float fx, float fy, float v,
                                                  float* __restrict__ out_xy, float* __restrict__ out_val) {
  const int b = plane_id / C, c = plane_id - b * C;
  ladder_apply(lad, b, fx, fy);  // NaN coordinates stay NaN
  const int row = lad.scatter ? lad.scatter[b] : b;
  if (row < 0) return;
  const long long o = (long long)row * C + c;
  out_xy[2 * o] = fx;
  out_xy[2 * o + 1] = fy;
  out_val[o] = v;
}

// Shared epilogue: thread 0 holds the CTA's Best `r`; publish it, let the plane's last CTA combine the
// chunk partials, then warp 0 of that CTA applies the threshold and the integral refinement.
__device__ __forceinline__ void global_peaks_finish(Best r, const void* __restrict__ plane, int dt, int plane_id,
                                                    int chunk, int n_chunks, int H, int W, long long sh, long long sw, float thr,
                                                    int refine_size, float* __restrict__ part_v,
                                                    int* __restrict__ part_xy, unsigned* __restrict__ tickets,
                                                    float* __restrict__ out_xy, float* __restrict__ out_val,
                                                    Best* s_best, bool* s_last, const Ladder& lad, int C) {
  if (threadIdx.x == 0) {
    if (n_chunks > 1) {
      const long long slot = (long long)plane_id * n_chunks + chunk;
      part_v[slot] = r.v;
      part_xy[2 * slot] = r.x;
      part_xy[2 * slot + 1] = r.y;
      __threadfence();
      const unsigned t = atomicAdd(tickets + plane_id, 1u);
      *s_last = (t == (unsigned)(n_chunks - 1));
      if (*s_last) {
        __threadfence();
        tickets[plane_id] = 0;  // self-reset so the workspace can be reused without a memset
        r = Best{__ldcg(part_v + (long long)plane_id * n_chunks), __ldcg(part_xy + 2LL * plane_id * n_chunks),
                 __ldcg(part_xy + 2LL * plane_id * n_chunks + 1)};
        for (int k = 1; k < n_chunks; ++k) {
          const long long s2 = (long long)plane_id * n_chunks + k;
          r = best_merge(r, Best{__ldcg(part_v + s2), __ldcg(part_xy + 2 * s2), __ldcg(part_xy + 2 * s2 + 1)});
        }
      }
    } else {
      *s_last = true;
    }
    if (*s_last) s_best[0] = r;
  }
  __syncthreads();
  if (*s_last && threadIdx.x < 32) {  // warp 0 of the plane's last CTA: threshold + refinement
    const Best b = s_best[0];
    // An empty plane (H*W == 0) cannot occur: the host rejects it.
    const bool low = b.v < thr;  // false for NaN, like torch (ops/peaks.py:121)
    float fx = low ? NAN : (float)b.x, fy = low ? NAN : (float)b.y;
    if (!low && refine_size > 0) {
      float ox, oy;
      integral_refine_warp(plane, dt, H, W, sh, sw, fx, fy, refine_size, threadIdx.x, &ox, &oy);
      fx = __fadd_rn(fx, ox);  // ops/peaks.py:179
      fy = __fadd_rn(fy, oy);
    }
    if (threadIdx.x == 0) write_global_peak(lad, plane_id, C, fx, fy, low ? 0.f : b.v, out_xy, out_val);
  }
}

// Generic variant: any strides, any chunk size, any element type; one associative (v, x, y) merge per element.
template <typename T>
__global__ void __launch_bounds__(256)
global_peaks_kernel(const T* __restrict__ cms, int C, int H, int W, long long sb, long long sc, long long sh,
                    long long sw, int vec_ok, int rows_per_chunk, int n_chunks, float thr, int refine_size,
                    float* __restrict__ part_v, int* __restrict__ part_xy, unsigned* __restrict__ tickets,
                    float* __restrict__ out_xy, float* __restrict__ out_val, Ladder lad) {
  constexpr int PER = Elem<T>::PER16;
  const int plane_id = blockIdx.x / n_chunks;
  const int chunk = blockIdx.x % n_chunks;
  const int b = plane_id / C, c = plane_id % C;
  const T* plane = cms + (long long)b * sb + (long long)c * sc;
  const int y0 = chunk * rows_per_chunk;
  const int y1 = min(H, y0 + rows_per_chunk);
  Best acc{-INFINITY, 0x7fffffff, 0x7fffffff};
  bool any = false;
  auto take = [&](float v, int x, int y) {
    Best o{v, x, y};
    acc = any ? best_merge(acc, o) : o;
    any = true;
  };
  if (vec_ok) {
    const int WV = W / PER;
    const int nv = (y1 - y0) * WV;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      const int y = y0 + i / WV, xv = i % WV;
      float e[PER];
      Elem<T>::unpack(ldg_stream16(plane + (long long)y * sh + PER * xv), e);
#pragma unroll
      for (int k = 0; k < PER; ++k) take(e[k], PER * xv + k, y);
    }
  } else {
    const int n = (y1 - y0) * W;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int y = y0 + i / W, x = i % W;
      take(Elem<T>::load1(plane + (long long)y * sh + (long long)x * sw), x, y);
    }
  }
  if (!any) acc = Best{-INFINITY, 0x7fffffff, 0x7fffffff};  // -inf loses to every real element
  __shared__ Best s_best[8];
  __shared__ bool s_last;
  acc = best_warp(acc);
  if (lane_id() == 0) s_best[threadIdx.x >> 5] = acc;
  __syncthreads();
  Best r = s_best[0];
  if (threadIdx.x == 0)
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = best_merge(r, s_best[w]);
  global_peaks_finish(r, plane, Elem<T>::DT, plane_id, chunk, n_chunks, H, W, sh, sw, thr, refine_size, part_v, part_xy,
                      tickets, out_xy, out_val, s_best, &s_last, lad, C);
}

// Register-resident variant (the product path for vectorisable planes whose chunk is <= V*1024 elements).
// The first version above spent ~40 issue slots per 16-byte load on the (v, x, y) merge and ran at 32 % of the
// HBM roofline.  Here a thread issues ALL of its 128-bit loads up front and keeps the values in registers:
// pass 1 is one FMNMX per element (+ NaN detection), the CTA agrees on the maximum m, pass 2 looks for
// elements equal to m (a rare branch) and reduces min(x), min(y) - the reference's two independent arg-maxes
// (ops/peaks.py:103-111).  A chunk that contains a NaN takes the exact generic merge over the same registers.
template <typename T, int V>
__global__ void __launch_bounds__(256)
global_peaks_regs_kernel(const T* __restrict__ cms, int C, int H, int W, long long sb, long long sc, long long sh,
                         int rows_per_chunk, int n_chunks, float thr, int refine_size, float* __restrict__ part_v,
                         int* __restrict__ part_xy, unsigned* __restrict__ tickets, float* __restrict__ out_xy,
                         float* __restrict__ out_val, Ladder lad) {
  constexpr int PER = Elem<T>::PER16;
  const int plane_id = blockIdx.x / n_chunks;
  const int chunk = blockIdx.x % n_chunks;
  const int b = plane_id / C, c = plane_id % C;
  const T* plane = cms + (long long)b * sb + (long long)c * sc;
  const int y0 = chunk * rows_per_chunk;
  const int y1 = min(H, y0 + rows_per_chunk);
  const int WV = W / PER;
  const int nv = (y1 - y0) * WV;
  const bool contig = (sh == W);
  const T* base = plane + (long long)y0 * sh;
  uint4 v[V];
#pragma unroll
  for (int u = 0; u < V; ++u) {
    const int i = threadIdx.x + u * 256;
    if (i < nv) {
      const T* p = contig ? base + (long long)PER * i : base + (long long)(i / WV) * sh + PER * (i % WV);
      v[u] = ldg_stream16(p);
    } else {
      v[u] = Elem<T>::neg_inf();  // loses to (or ties harmlessly with) real data
    }
  }
  // max.NaN: a NaN anywhere in the thread's values makes m NaN
  float m = -INFINITY;
#pragma unroll
  for (int u = 0; u < V; ++u) m = fmax_nan(m, Elem<T>::vmax_nan(v[u]));
  const bool has_nan = (m != m);
  if (has_nan) m = -INFINITY;
  __shared__ float s_m[8];
  __shared__ int s_x[8], s_y[8];
  __shared__ Best s_best[8];
  __shared__ bool s_last;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, d));
  if (lane == 0) s_m[warp] = m;
  const int any_nan = __syncthreads_or(has_nan ? 1 : 0);
  Best r;
  if (any_nan) {
    // exact generic merge over the registers (NaN ranks above everything; min x / min y among the NaNs)
    Best acc{-INFINITY, 0x7fffffff, 0x7fffffff};
    bool any = false;
#pragma unroll
    for (int u = 0; u < V; ++u) {
      const int i = threadIdx.x + u * 256;
      if (i < nv) {
        const int y = y0 + i / WV, x = PER * (i % WV);
        float e[PER];
        Elem<T>::unpack(v[u], e);
#pragma unroll
        for (int k = 0; k < PER; ++k) {
          const Best o{e[k], x + k, y};
          acc = any ? best_merge(acc, o) : o;
          any = true;
        }
      }
    }
    acc = best_warp(acc);
    if (lane == 0) s_best[warp] = acc;
    __syncthreads();
    r = s_best[0];
    if (threadIdx.x == 0)
      for (int w = 1; w < 8; ++w) r = best_merge(r, s_best[w]);
    __syncthreads();  // s_best[0] is rewritten by the epilogue
  } else {
    m = s_m[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_m[w]);
    int bx = 0x7fffffff, by = 0x7fffffff;
#pragma unroll
    for (int u = 0; u < V; ++u) {
      if (Elem<T>::vmax_nan(v[u]) == m) {  // rare
        const int i = threadIdx.x + u * 256;
        if (i < nv) {
          const int y = y0 + i / WV, x = PER * (i % WV);
          float e[PER];
          Elem<T>::unpack(v[u], e);
          int k = PER - 1;
#pragma unroll
          for (int q = PER - 2; q >= 0; --q) k = (e[q] == m) ? q : k;  // first match = min x
          bx = min(bx, x + k);
          by = min(by, y);
        }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      bx = min(bx, __shfl_xor_sync(FULL, bx, d));
      by = min(by, __shfl_xor_sync(FULL, by, d));
    }
    if (lane == 0) { s_x[warp] = bx; s_y[warp] = by; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int w = 1; w < 8; ++w) { bx = min(bx, s_x[w]); by = min(by, s_y[w]); }
    }
    r = Best{m, bx, by};
  }
  global_peaks_finish(r, plane, Elem<T>::DT, plane_id, chunk, n_chunks, H, W, sh, 1, thr, refine_size, part_v, part_xy,
                      tickets, out_xy, out_val, s_best, &s_last, lad, C);
}

#ifdef SNB_AB_VARIANTS
// A/B only (-DSNB_AB_VARIANTS): persistent ring variant for fp32 planes that fit one shared-memory stage (cfg2's 80x80
// crops).  With one CTA per small plane the loads are in flight for only about a third of a CTA's life (the
// rest is reductions + the refinement's dependent taps), which capped K2 at ~54 % of the HBM roofline.  Here a
// persistent CTA walks planes p, p + grid, ... through a 3-stage shared-memory ring filled by cp.async
// (LDGSTS, no registers tied up): the loads of the next two planes are always in flight while the current one
// is reduced out of shared memory, and the refinement reads its 5x5 taps from the same stage.
constexpr int GP_STAGES = 3;

__global__ void __launch_bounds__(256)
global_peaks_ring_kernel(const float* __restrict__ cms, int n_planes, int C, int H, int W, long long sb, long long sc,
                         long long sh, float thr, int refine_size, float* __restrict__ out_xy,
                         float* __restrict__ out_val, Ladder lad) {
  extern __shared__ __align__(16) float gp_ring[];
  __shared__ float s_m[8];
  __shared__ int s_x[8], s_y[8];
  __shared__ Best s_best[8];
  const int plane_elems = H * W, n4 = plane_elems >> 2, W4 = W >> 2;
  const bool contig = (sh == W);
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  auto issue = [&](int p, int stage) {
    if (p < n_planes) {
      const float* base = cms + (long long)(p / C) * sb + (long long)(p % C) * sc;
      float* dst = gp_ring + (size_t)stage * plane_elems;
      for (int i = tid; i < n4; i += 256) {
        const float* src = contig ? base + 4LL * i : base + (long long)(i / W4) * sh + 4 * (i % W4);
        cp_async16(dst + 4 * i, src);
      }
    }
    cp_async_commit();  // always commit (possibly empty) so the group count is uniform
  };
  const int p0 = blockIdx.x, step = gridDim.x;
  issue(p0, 0);
  issue(p0 + step, 1);
  int it = 0;
  for (int p = p0; p < n_planes; p += step, ++it) {
    issue(p + 2 * step, (it + 2) % GP_STAGES);  // refills the stage consumed in the previous iteration
    cp_async_wait<2>();                          // everything but the two newest groups has landed: plane p is in
    __syncthreads();
    const float* st = gp_ring + (size_t)(it % GP_STAGES) * plane_elems;
    const float4* st4 = reinterpret_cast<const float4*>(st);
    float m = -INFINITY;
    bool has_nan = false;
    for (int i = tid; i < n4; i += 256) {
      const float4 v = st4[i];
      m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
      has_nan = has_nan || (v.x != v.x) || (v.y != v.y) || (v.z != v.z) || (v.w != v.w);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, d));
    if (lane == 0) s_m[warp] = m;
    const int any_nan = __syncthreads_or(has_nan ? 1 : 0);
    if (any_nan) {  // exact generic merge (NaN ranks above everything; min x / min y among the NaNs)
      Best acc{-INFINITY, 0x7fffffff, 0x7fffffff};
      bool any = false;
      for (int i = tid; i < plane_elems; i += 256) {
        const Best o{st[i], i % W, i / W};
        acc = any ? best_merge(acc, o) : o;
        any = true;
      }
      acc = best_warp(acc);
      if (lane == 0) s_best[warp] = acc;
      __syncthreads();
      if (tid == 0) {
        Best r = s_best[0];
        for (int w = 1; w < 8; ++w) r = best_merge(r, s_best[w]);
        s_best[0] = r;
      }
    } else {
      m = s_m[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) m = fmaxf(m, s_m[w]);
      int bx = 0x7fffffff, by = 0x7fffffff;
      for (int i = tid; i < n4; i += 256) {
        const float4 v = st4[i];
        if (v.x == m || v.y == m || v.z == m || v.w == m) {  // rare
          const int k = (v.x == m) ? 0 : ((v.y == m) ? 1 : ((v.z == m) ? 2 : 3));  // first match = min x in the chunk
          bx = min(bx, 4 * (i % W4) + k);
          by = min(by, i / W4);
        }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        bx = min(bx, __shfl_xor_sync(FULL, bx, d));
        by = min(by, __shfl_xor_sync(FULL, by, d));
      }
      if (lane == 0) { s_x[warp] = bx; s_y[warp] = by; }
      __syncthreads();
      if (tid == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) { bx = min(bx, s_x[w]); by = min(by, s_y[w]); }
        s_best[0] = Best{m, bx, by};
      }
    }
    __syncthreads();
    if (tid < 32) {  // threshold (ops/peaks.py:121-129) + integral refinement out of the shared-memory stage
      const Best b = s_best[0];
      const bool low = b.v < thr;  // false for NaN, like torch
      float fx = low ? NAN : (float)b.x, fy = low ? NAN : (float)b.y;
      if (!low && refine_size > 0) {
        float ox, oy;
        integral_refine_warp<false>(st, SNB_DTYPE_F32, H, W, W, 1, fx, fy, refine_size, lane, &ox, &oy);
        fx = __fadd_rn(fx, ox);  // ops/peaks.py:179
        fy = __fadd_rn(fy, oy);
      }
      if (tid == 0) write_global_peak(lad, p, C, fx, fy, low ? 0.f : b.v, out_xy, out_val);
    }
    __syncthreads();  // everyone is done with this stage (and s_best) before the next iteration refills it
  }
  cp_async_wait<0>();
}
#endif  // SNB_AB_VARIANTS

// Warp-per-plane variant (the product path for small vectorisable planes, e.g. cfg2's 80x80 crops).  The ring
// kernel above still spends five CTA-wide barriers and a serial warp-0 refinement per plane, during which the
// other seven warps of the CTA idle: it measured 2.6 TB/s (40 % of the HBM roofline).  Here ONE WARP owns a
// plane end to end and never meets a barrier: a rolling window of U 128-bit streaming loads per lane stays in
// flight (U*512 B per warp) and the loop is branch-free (~12 issue slots per load: a first attempt that computed
// positions on a "new maximum" branch took that branch in almost every warp-step and spent 68).  The reference's
// two independent arg-maxes (ops/peaks.py:103-111) are min(x), min(y) over the elements equal to the maximum; a
// tie inside one lane's sequence or a NaN (max.NaN poisons the running maximum) sends the warp to the exact
// generic merge.  One 32-thread CTA per plane, so
// the hardware scheduler spreads the planes evenly over the 148 SMs (cfg2: 22.5 planes per SM, all resident).

#ifdef SNB_AB_VARIANTS
// A/B build only: globaltimer stamps of every plane's warp (start of streaming, end of streaming, end of epilogue),
// read back by tools/k2_probe.py through snb_ab_k2_stamps
constexpr int K2_STAMP_PLANES = 16384;
__device__ unsigned long long g_k2_stamps[3 * K2_STAMP_PLANES];
__device__ int g_k2_stamps_on = 0;
__device__ __forceinline__ void k2_stamp(int p, int which, int lane) {
  if (g_k2_stamps_on && lane == 0 && p < K2_STAMP_PLANES) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_k2_stamps[3 * p + which] = t;
  }
}
#define SNB_K2_STAMP(p, which, lane) k2_stamp(p, which, lane)
#else
#define SNB_K2_STAMP(p, which, lane) ((void)0)
#endif

constexpr int GP_WARPS_PER_CTA = 1;    // planes (= warps) per CTA of the warp kernel
// 128-bit loads in flight per lane.  Around one wave of warps (cfg2 batch 256: 3 328 planes on 148 SMs) a SHALLOW window
// is faster - 4 -> 15.9-16.9 us, 5 -> 15.8-16.1, 6 -> 15.9-16.5, 8 -> 16.5-16.9, 10 -> 16.6, 12 -> 17.0 on two B200s
// (fp16 maps: 5 -> 9.6, 8 -> 10.5) - while a sub-wave launch needs the deeper one to cover the latency (batch 64:
// 8 -> 8.4 us, 5 -> 10.3); many waves do not care (batch 1 024: 50-51 us at every depth).  profiles/r2_sweep_small_launch.jsonl
constexpr int GP_LOADS_IN_FLIGHT = 5;
constexpr int GP_LOADS_IN_FLIGHT_SUBWAVE = 8;  // fewer than 16 planes per SM

template <typename T, int U, bool CONTIG, int WPC>
__global__ void __launch_bounds__(32 * WPC)
global_peaks_warp_kernel(const T* __restrict__ cms, int planes, int C, int H, int W, long long sb, long long sc,
                         long long sh, float thr, int refine_size, float* __restrict__ out_xy,
                         float* __restrict__ out_val, Ladder lad) {
  constexpr int PER = Elem<T>::PER16;
  pdl_launch_dependents();
  // WPC independent warps per CTA (no barrier, no shared memory): fewer, fatter CTAs only shorten the launch ramp of
  // a one-wave grid
  const int p = blockIdx.x * WPC + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (p >= planes) return;
  const T* plane = cms + (long long)(p / C) * sb + (long long)(p % C) * sc;
  const int WV = W / PER, nv = H * WV;
  pdl_wait();  // the maps may be the previous kernel's output
  SNB_K2_STAMP(p, 0, lane);
  // Load cursor of this lane: 128-bit load number `li` = lane + 32 * (loads issued so far); for strided planes the
  // (row, column) pair is advanced incrementally (no integer division in the loop).
  int li = lane, ly = lane / WV, lxv = lane - ly * WV;
  const int dy32 = 32 / WV, dx32 = 32 - dy32 * WV;
  auto load_next = [&]() -> uint4 {
    uint4 v = Elem<T>::neg_inf();
    if (li < nv) v = ldg_stream16(CONTIG ? plane + (long long)PER * li : plane + (long long)ly * sh + PER * lxv);
    li += 32;
    if (!CONTIG) {
      ly += dy32;
      lxv += dx32;
      if (lxv >= WV) { lxv -= WV; ++ly; }
    }
    return v;
  };
  uint4 buf[U];
#pragma unroll
  for (int u = 0; u < U; ++u) buf[u] = load_next();
  // Per lane, branch-free: running maximum m (max.NaN: a NaN sticks), the number kbest of the FIRST load that
  // reached it, and whether a later load tied with it.
  // The 16-byte word that holds the running maximum stays in registers too (four predicated moves per load): the
  // epilogue then needs no re-read of that word, one dependent L2 round trip less at the end of a one-wave kernel.
  float m = -INFINITY;
  int kbest = 0, k = 0;
  bool tie = false;
  uint4 vbest = Elem<T>::neg_inf();
  for (int base = 0; base < nv; base += 32 * U) {
#pragma unroll
    for (int u = 0; u < U; ++u, ++k) {
      const uint4 v = buf[u];
      buf[u] = load_next();
      const float mx = Elem<T>::vmax_nan(v);
      const bool up = mx > m;
      tie = up ? false : (tie || mx == m);
      kbest = up ? k : kbest;
      vbest.x = up ? v.x : vbest.x; vbest.y = up ? v.y : vbest.y; vbest.z = up ? v.z : vbest.z; vbest.w = up ? v.w : vbest.w;
      m = fmax_nan(m, mx);
    }
  }
  // Warp: the plane maximum and the lanes that hold it.  Exactly one un-tied holder is the common case; several
  // holders still give min(x), min(y) exactly; a tie INSIDE a holder's own sequence (plateaus, saturated maps) or a
  // NaN sends the warp to the exact generic merge below.
  SNB_K2_STAMP(p, 1, lane);
  float gm = m;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) gm = fmax_nan(gm, __shfl_xor_sync(FULL, gm, d));
  Best r{gm, 0x7fffffff, 0x7fffffff};
  const bool low_all = gm < thr;  // below threshold the position is never reported (ops/peaks.py:121-129)
  if (!low_all) {
    const bool holder = (m == gm);
    const bool exact = (gm != gm) || __any_sync(FULL, holder && tie);
    if (exact) {
      Best acc{-INFINITY, 0x7fffffff, 0x7fffffff};
      bool any = false;
      for (int i = lane; i < H * W; i += 32) {
        const int y = i / W, x = i - y * W;
        const Best o{Elem<T>::load1(plane + (long long)y * sh + x), x, y};
        acc = any ? best_merge(acc, o) : o;
        any = true;
      }
      r = best_warp(acc);
    } else {
      int bx = 0x7fffffff, by = 0x7fffffff;
      if (holder) {  // the element inside the 16-byte word that held the maximum
        const int i = lane + 32 * kbest;
        const int y = i / WV, xv = i - y * WV;
        float e[PER];
        Elem<T>::unpack(vbest, e);
        int q = PER - 1;
#pragma unroll
        for (int j = PER - 2; j >= 0; --j) q = (e[j] == gm) ? j : q;  // first match = min x of the word
        bx = PER * xv + q;
        by = y;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        bx = min(bx, __shfl_xor_sync(FULL, bx, d));
        by = min(by, __shfl_xor_sync(FULL, by, d));
      }
      r.x = bx;
      r.y = by;
    }
  }
  const bool low = r.v < thr;  // false for NaN, like torch (ops/peaks.py:121)
  float fx = low ? NAN : (float)r.x, fy = low ? NAN : (float)r.y;
  if (!low && refine_size > 0) {
    float ox, oy;
    integral_refine_warp(plane, Elem<T>::DT, H, W, sh, 1, fx, fy, refine_size, lane, &ox, &oy);
    fx = __fadd_rn(fx, ox);  // ops/peaks.py:179
    fy = __fadd_rn(fy, oy);
  }
  if (lane == 0) write_global_peak(lad, p, C, fx, fy, low ? 0.f : r.v, out_xy, out_val);
  SNB_K2_STAMP(p, 2, lane);
}

// ----------------------------------------------------------------------------------------
// K3: crop_bboxes.  One thread per output element; top-left = trunc(tl + size//2) - size//2
// in fp32 exactly as ops/crops.py:85-90; taps outside the image are 0.
// ----------------------------------------------------------------------------------------
template <typename T>
__global__ void crop_bboxes_kernel(const T* __restrict__ img, int S, int C, int H, int W, long long sb, long long sc,
                                   long long sh, long long sw, const float* __restrict__ bboxes,
                                   const long long* __restrict__ sample_inds, long long n, int ch, int cw,
                                   T* __restrict__ out, int* __restrict__ status) {
  const long long total = n * C * ch * cw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xi = (int)(i % cw);
    long long r = i / cw;
    const int yi = (int)(r % ch);
    r /= ch;
    const int c = (int)(r % C);
    const long long k = r / C;
    const float tlx = __fadd_rn(__ldg(bboxes + k * 8 + 0), (float)(cw / 2));
    const float tly = __fadd_rn(__ldg(bboxes + k * 8 + 1), (float)(ch / 2));
    // float -> int64 truncation; far-out-of-range values saturate, which still lands outside the image
    const long long x = (long long)truncf(fminf(fmaxf(tlx, -1e15f), 1e15f)) - (cw / 2) + xi;
    const long long y = (long long)truncf(fminf(fmaxf(tly, -1e15f), 1e15f)) - (ch / 2) + yi;
    long long s = sample_inds[k];
    if (s < 0) s += S;  // python-style negative index
    T v = T(0);
    if (s < 0 || s >= S) {
      atomicOr(status, SNB_STATUS_BAD_INDEX);
    } else if (x >= 0 && x < W && y >= 0 && y < H) {
      v = img[s * sb + (long long)c * sc + y * sh + x * sw];
    }
    out[i] = v;
  }
}

// integral_regression(cms, xv, yv) on arbitrary patches (ops/peaks.py:66-86): one warp per (n, c).
__global__ void integral_regression_kernel(const float* __restrict__ p, long long n_planes, int h, int w,
                                           const float* __restrict__ xv, const float* __restrict__ yv,
                                           float* __restrict__ ox, float* __restrict__ oy) {
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= n_planes) return;
  const float* q = p + wid * h * w;
  double z = 0, sx = 0, sy = 0;
  for (int i = lane_id(); i < h * w; i += 32) {
    const float v = q[i];
    z += v;
    sx += (double)__fmul_rn(__ldg(xv + i % w), v);
    sy += (double)__fmul_rn(__ldg(yv + i / w), v);
  }
  for (int d = 16; d > 0; d >>= 1) {
    z += __shfl_xor_sync(FULL, z, d);
    sx += __shfl_xor_sync(FULL, sx, d);
    sy += __shfl_xor_sync(FULL, sy, d);
  }
  if (lane_id() == 0) {
    ox[wid] = __fdiv_rn((float)sx, (float)z);
    oy[wid] = __fdiv_rn((float)sy, (float)z);
  }
}

// morphological_dilation (ops/peaks.py:26-63): max of the 8 neighbours, -inf outside, NaN propagates.
__global__ void dilate8_kernel(const float* __restrict__ img, long long P, int H, int W, float* __restrict__ out) {
  const long long total = P * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long r = i / W;
    const int y = (int)(r % H);
    const float* plane = img + (r / H) * H * W;
    float m = -INFINITY;
    bool nan = false;
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = x + dx;
        if ((dy == 0 && dx == 0) || xx < 0 || xx >= W) continue;
        const float v = __ldg(plane + (long long)yy * W + xx);
        nan = nan || isnan(v);
        m = fmaxf(m, v);
      }
    }
    out[i] = nan ? NAN : m;
  }
}

// make_centered_bboxes (data/instance_cropping.py:129-171): corners TL,TR,BR,BL = centre -/+ half,
// then +/- 0.5 inset; two separately rounded fp32 ops per coordinate.
__global__ void centered_bboxes_kernel(const float* __restrict__ c, long long n, float half_h, float half_w,
                                       float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = c[2 * i], y = c[2 * i + 1];
  const float xl = __fadd_rn(__fsub_rn(x, half_w), 0.5f), xr = __fadd_rn(__fadd_rn(x, half_w), -0.5f);
  const float yt = __fadd_rn(__fsub_rn(y, half_h), 0.5f), yb = __fadd_rn(__fadd_rn(y, half_h), -0.5f);
  float* o = out + 8 * i;
  o[0] = xl; o[1] = yt;
  o[2] = xr; o[3] = yt;
  o[4] = xr; o[5] = yb;
  o[6] = xl; o[7] = yb;
}

// CentroidLayer.postprocess after find_local_peaks (inference/layers/centroid.py:196-258): per frame, when
// there are more peaks than max_instances keep the top max_instances by VALUE (torch.topk: descending; equal
// values: lower index first), else keep (y, x, channel) order; coordinates / input_scale, NaN padding, then
// / eff_scale[b].  The table's coordinates already carry the stride (snb_local_peaks' xy_scale).  One CTA / frame.
__global__ void __launch_bounds__(128)
peaks_topk_kernel(const int* __restrict__ frame_count, int cap, const float* __restrict__ xy,
                  const float* __restrict__ val, int max_instances, float input_scale, const float* __restrict__ eff,
                  float* __restrict__ o_xy, float* __restrict__ o_val) {
  extern __shared__ int s_src[];
  const int b = blockIdx.x;
  const int n = min(frame_count[b], cap);
  const int keep = min(n, max_instances);
  const float* v = val + (long long)b * cap;
  if (n > max_instances) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float vi = v[i];
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const float vj = v[j];
        rank += (vj > vi || (vj == vi && j < i)) ? 1 : 0;
      }
      if (rank < max_instances) s_src[rank] = i;
    }
  } else {
    for (int i = threadIdx.x; i < keep; i += blockDim.x) s_src[i] = i;
  }
  __syncthreads();
  const float e = eff ? eff[b] : 1.0f;
  for (int r = threadIdx.x; r < max_instances; r += blockDim.x) {
    float x = NAN, y = NAN, pv = NAN;
    if (r < keep) {
      const long long s = (long long)b * cap + s_src[r];
      x = __fdiv_rn(__fdiv_rn(xy[2 * s], input_scale), e);
      y = __fdiv_rn(__fdiv_rn(xy[2 * s + 1], input_scale), e);
      pv = val[s];
    }
    const long long o = (long long)b * max_instances + r;
    o_xy[2 * o] = x;
    o_xy[2 * o + 1] = y;
    o_val[o] = pv;
  }
}

// The coordinate ladder as a stand-alone elementwise op (inference/ops/coord.py:27-90): coords is
// (n_samples, pairs_per_sample, 2) contiguous.
__global__ void coord_ladder_kernel(const float* __restrict__ xy, long long n_pairs, long long pairs_per_sample,
                                    Ladder lad, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  float x = xy[2 * i], y = xy[2 * i + 1];
  ladder_apply(lad, (int)(i / pairs_per_sample), x, y);
  out[2 * i] = x;
  out[2 * i + 1] = y;
}

static inline int grid_for(long long work_items, int per_block, int max_blocks) {
  long long g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

}  // namespace snb

using namespace snb;

// SM count of the CURRENT device (cached per device: one process may drive several, possibly different, GPUs).
static int sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

static inline bool dtype_ok(int dt) { return dt == SNB_DTYPE_F32 || dt == SNB_DTYPE_F16 || dt == SNB_DTYPE_BF16; }

// 128-bit vector path: unit column stride, row length / strides multiples of the vector width, 16-byte aligned base.
static inline bool vec_ok_for(const void* p, int dt, int W, long long sb, long long sc, long long sh, long long sw) {
  const int per = 16 / dtype_size(dt);
  return (sw == 1) && (W % per == 0) && (sh % per == 0) && (sc % per == 0) && (sb % per == 0) && (((uintptr_t)p) % 16 == 0);
}

template <typename T>
static int launch_detect(const T* cms, int B, int C, int H, int W, long long sb, long long sc, long long sh,
                         long long sw, float threshold, int cap, int* frame_count, uint32_t* keys, cudaStream_t st) {
  constexpr int PER = Elem<T>::PER16;
  const bool vec = vec_ok_for(cms, Elem<T>::DT, W, sb, sc, sh, sw);
  const long long rows = (long long)B * C * H;
#ifdef SNB_AB_VARIANTS
  {
    // A/B: producer / consumer bulk-async (TMA) ring, any element type, contiguous tensors (tools/detect_variants.py)
    static const bool use_tma = getenv("SNB_DETECT_TMA") != nullptr;
    static const int tma_ctas = getenv("SNB_DETECT_TMA_CTAS") ? atoi(getenv("SNB_DETECT_TMA_CTAS")) : 3;
    const bool contiguous = vec && sh == W && sc == (long long)H * W && sb == (long long)C * H * W;
    if (contiguous && use_tma) {
      const size_t smem = (size_t)TMA_STAGES * TMA_STAGE_BYTES + 2 * TMA_STAGES * sizeof(uint64_t);
      if (cudaFuncSetAttribute(local_peaks_detect_tma<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
          cudaSuccess)
        return SNB_ERR_CUDA_LAUNCH;
      const long long n_elems = rows * W;
      const long long n_chunks = (n_elems * (long long)sizeof(T) + TMA_STAGE_BYTES - 1) / TMA_STAGE_BYTES;
      const int grid = (int)std::min<long long>(n_chunks, (long long)sm_count() * tma_ctas);
      local_peaks_detect_tma<T><<<grid, TMA_THREADS, smem, st>>>(cms, n_elems, C, H, W, threshold, cap, frame_count, keys);
      return SNB_OK;
    }
  }
  if constexpr (Elem<T>::DT == SNB_DTYPE_F32) {
    // A/B: the cp.async.bulk ring (tools/detect_variants.py); the LDG.128 kernel is the product
    const bool contiguous = vec && sh == W && sc == (long long)H * W && sb == (long long)C * H * W;
    static const bool use_bulk = getenv("SNB_DETECT_BULK") != nullptr;
    if (contiguous && use_bulk) {
      const size_t smem = (size_t)BULK_STAGES * BULK_STAGE_FLOATS * 4 + BULK_STAGES * sizeof(uint64_t);
      if (cudaFuncSetAttribute(local_peaks_detect_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
          cudaSuccess)
        return SNB_ERR_CUDA_LAUNCH;
      const long long n_elems = rows * W;
      const long long n_chunks = (n_elems + BULK_STAGE_FLOATS - 1) / BULK_STAGE_FLOATS;
      const int grid = (int)std::min<long long>(n_chunks, (long long)sm_count() * 3);  // 3 CTAs x 64 KB per SM
      local_peaks_detect_bulk<<<grid, BULK_THREADS, smem, st>>>(cms, n_elems, C, H, W, threshold, cap, frame_count, keys);
      return SNB_OK;
    }
  }
#endif
  if (vec && B <= 65535 && C <= 65535) {
    // 8 warps per CTA; the variant (128-bit loads in flight per lane, rows per warp, CTAs per SM) is picked by the
    // number of 128-bit vectors per row.  One row per warp with 4 loads per lane for >= 128 vectors (fp32 W >= 512,
    // half W >= 1024), 40 registers -> 6 CTAs = 48 warps per SM, and a NON-persistent grid so the hardware CTA
    // scheduler balances the SMs.  Rows of 64..127 vectors (fp32 W = 256; half-precision cfg3 maps: 512 x 2 B = 1 KB):
    // one row of two loads per lane at 32 registers -> 8 CTAs = 64 warps per SM.  A/B on B200 for f16 cfg3 maps
    // (profiles/r2_detect_ab.jsonl): (2 loads, 1 row, 8 CTAs) 28.9 us, (2, 2, 6) 30.8 us, (2, 4, 4) 33.0 us, (4, 1, 6) 34.9 us.
    const int vecs = W / PER;
    int variant = (vecs >= 128) ? 0 : (vecs >= 64 ? 1 : 2);
#ifdef SNB_AB_VARIANTS
    static const int forced = getenv("SNB_DETECT_VARIANT") ? atoi(getenv("SNB_DETECT_VARIANT")) : -1;
    if (forced >= 0) variant = forced;
#endif
#define SNB_DETECT(U, R, MB)                                                                                     \
  do {                                                                                                           \
    const dim3 grid((unsigned)((H + 8 * R - 1) / (8 * R)), (unsigned)C, (unsigned)B);                            \
    if (H % (8 * R) == 0 && vecs % (32 * U) == 0)                                                                \
      local_peaks_detect_vec<T, U, R, MB, true><<<grid, 256, 0, st>>>(cms, C, H, W, sb, sc, sh, threshold, cap,  \
                                                                      frame_count, keys);                        \
    else                                                                                                         \
      local_peaks_detect_vec<T, U, R, MB, false><<<grid, 256, 0, st>>>(cms, C, H, W, sb, sc, sh, threshold, cap, \
                                                                       frame_count, keys);                       \
  } while (0)
    switch (variant) {
      case 0: SNB_DETECT(4, 1, 6); break;  // >= 128 vectors per row
      case 1: SNB_DETECT(2, 1, 8); break;  // 64..127 vectors per row
      case 2: SNB_DETECT(1, 4, 6); break;  // narrow maps: four rows of one load each
#ifdef SNB_AB_VARIANTS
      case 6: SNB_DETECT(2, 2, 6); break;  // A/B: two rows of two loads
      case 4: SNB_DETECT(4, 2, 4); break;  // A/B: 8 loads per lane
      case 7: SNB_DETECT(2, 4, 4); break;  // A/B: four rows of two loads
      case 8: SNB_DETECT(2, 1, 8); break;  // A/B: one row of two loads, 8 CTAs / SM
#endif
      default: return SNB_ERR_BAD_ARG;
    }
#undef SNB_DETECT
  } else {
    const int grid = grid_for(rows * W, 256 * 4, sm_count() * 16);
    local_peaks_detect_scalar<T><<<grid, 256, 0, st>>>(cms, B, C, H, W, sb, sc, sh, sw, threshold, cap, frame_count, keys);
  }
  return SNB_OK;
}

namespace snb {
// K1a launcher shared with the fused pipeline (pipeline.cu): zero_counters = false when the caller guarantees that
// frame_count is already zero (the fused tail resets it, SNB_FLAG_SELF_RESET_COUNTERS).
int detect_launch(const void* cms, int dtype, int B, int C, int H, int W, long long sb, long long sc, long long sh,
                  long long sw, float threshold, int cap, int* frame_count, uint32_t* keys, void* ev_begin, void* ev_end,
                  void* stream_, bool zero_counters) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (B < 0 || C <= 0 || H <= 0 || W <= 0 || cap <= 0 || !dtype_ok(dtype)) return SNB_ERR_BAD_ARG;
  if ((double)H * W * C >= 4294967295.0) return SNB_ERR_UNSUPPORTED;
  if (B == 0) return SNB_OK;
  if (zero_counters && cudaMemsetAsync(frame_count, 0, sizeof(int) * B, st) != cudaSuccess) return SNB_ERR_CUDA_LAUNCH;
  if (ev_begin) cudaEventRecord((cudaEvent_t)ev_begin, st);
  int rc;
  switch (dtype) {
    case SNB_DTYPE_F16:
      rc = launch_detect((const __half*)cms, B, C, H, W, sb, sc, sh, sw, threshold, cap, frame_count, keys, st);
      break;
    case SNB_DTYPE_BF16:
      rc = launch_detect((const __nv_bfloat16*)cms, B, C, H, W, sb, sc, sh, sw, threshold, cap, frame_count, keys, st);
      break;
    default:
      rc = launch_detect((const float*)cms, B, C, H, W, sb, sc, sh, sw, threshold, cap, frame_count, keys, st);
  }
  if (rc != SNB_OK) return rc;
  if (ev_end) cudaEventRecord((cudaEvent_t)ev_end, st);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
}  // namespace snb

// K1a: zero the per-frame counters and run the streaming detect kernel.  ev_begin / ev_end are
// optional cudaEvent_t handles recorded right around the kernel (in-situ timing for benchmarks).
extern "C" int snb_local_peaks_detect_t(const void* cms, int dtype, int B, int C, int H, int W, long long sb,
                                        long long sc, long long sh, long long sw, float threshold, int cap,
                                        int* frame_count, uint32_t* keys, void* ev_begin, void* ev_end, void* stream_) {
  return detect_launch(cms, dtype, B, C, H, W, sb, sc, sh, sw, threshold, cap, frame_count, keys, ev_begin, ev_end,
                       stream_, true);
}

extern "C" int snb_local_peaks_detect(const float* cms, int B, int C, int H, int W, long long sb, long long sc,
                                      long long sh, long long sw, float threshold, int cap, int* frame_count,
                                      uint32_t* keys, void* ev_begin, void* ev_end, void* stream_) {
  return snb_local_peaks_detect_t(cms, SNB_DTYPE_F32, B, C, H, W, sb, sc, sh, sw, threshold, cap, frame_count, keys,
                                  ev_begin, ev_end, stream_);
}

// K1b: per-frame key sort + value + integral refinement -> padded peak table.
extern "C" int snb_local_peaks_finalize_t(const void* cms, int dtype, int B, int C, int H, int W, long long sb,
                                          long long sc, long long sh, long long sw, int refine_size, float xy_scale,
                                          int cap, const int* frame_count, uint32_t* keys, float* out_xy, float* out_val,
                                          int* out_chan, int* status, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (B < 0 || C <= 0 || H <= 0 || W <= 0 || cap <= 0 || refine_size < 0 || !dtype_ok(dtype)) return SNB_ERR_BAD_ARG;
  if (B == 0) return SNB_OK;
  int n2 = 1;
  while (n2 < cap) n2 <<= 1;
  const size_t smem = sizeof(uint32_t) * (size_t)n2;
  int presorted = 0;
  if (smem > 200 * 1024) {
    // Pathological key counts: sort in global memory. Requires cap to be a power of two.
    if (n2 != cap) return SNB_ERR_BAD_ARG;
    const long long total = (long long)B * cap;
    const int g = (int)((total + 255) / 256);
    pad_keys_global<<<g, 256, 0, st>>>(keys, frame_count, cap, B);
    for (int k = 2; k <= cap; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) bitonic_step_global<<<g, 256, 0, st>>>(keys, cap, total, j, k);
    SNB_LAUNCH_CHECK();
    presorted = 1;
  } else if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(local_peaks_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return SNB_ERR_CUDA_LAUNCH;
  }
  local_peaks_finalize<<<B, 256, presorted ? 0 : smem, st>>>(cms, dtype, C, H, W, sb, sc, sh, sw, refine_size, xy_scale,
                                                            cap, presorted, frame_count, keys, out_xy, out_val,
                                                            out_chan, status);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_local_peaks_finalize(const float* cms, int B, int C, int H, int W, long long sb, long long sc,
                                        long long sh, long long sw, int refine_size, float xy_scale, int cap,
                                        const int* frame_count, uint32_t* keys, float* out_xy, float* out_val,
                                        int* out_chan, int* status, void* stream_) {
  return snb_local_peaks_finalize_t(cms, SNB_DTYPE_F32, B, C, H, W, sb, sc, sh, sw, refine_size, xy_scale, cap,
                                    frame_count, keys, out_xy, out_val, out_chan, status, stream_);
}

extern "C" int snb_local_peaks_t(const void* cms, int dtype, int B, int C, int H, int W, long long sb, long long sc,
                                 long long sh, long long sw, float threshold, int refine_size, float xy_scale, int cap,
                                 int* frame_count, uint32_t* keys, float* out_xy, float* out_val, int* out_chan,
                                 int* status, void* stream_) {
  if (refine_size < 0) return SNB_ERR_BAD_ARG;
  const int rc = snb_local_peaks_detect_t(cms, dtype, B, C, H, W, sb, sc, sh, sw, threshold, cap, frame_count, keys,
                                          nullptr, nullptr, stream_);
  if (rc != SNB_OK) return rc;
  return snb_local_peaks_finalize_t(cms, dtype, B, C, H, W, sb, sc, sh, sw, refine_size, xy_scale, cap, frame_count,
                                    keys, out_xy, out_val, out_chan, status, stream_);
}

extern "C" int snb_local_peaks(const float* cms, int B, int C, int H, int W, long long sb, long long sc, long long sh,
                               long long sw, float threshold, int refine_size, float xy_scale, int cap,
                               int* frame_count, uint32_t* keys, float* out_xy, float* out_val, int* out_chan,
                               int* status, void* stream_) {
  return snb_local_peaks_t(cms, SNB_DTYPE_F32, B, C, H, W, sb, sc, sh, sw, threshold, refine_size, xy_scale, cap,
                           frame_count, keys, out_xy, out_val, out_chan, status, stream_);
}

extern "C" int snb_pack_peaks(const int* frame_count, int B, int cap, const float* xy, const float* val,
                              const int* chan, float* o_xy, float* o_val, int* o_sample, int* o_chan, void* stream_) {
  if (B <= 0) return SNB_OK;
  pack_peaks<<<B, 128, 0, (cudaStream_t)stream_>>>(frame_count, B, cap, xy, val, chan, o_xy, o_val, o_sample, o_chan);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_global_peaks_workspace(int B, int C, int H, int W, int* rows_per_chunk, int* n_chunks,
                                          long long* n_bytes) {
  if (B < 0 || C <= 0 || H <= 0 || W <= 0) return SNB_ERR_BAD_ARG;
  int rpc = (int)((8192 + (long long)W - 1) / W);  // ~8K elements (32 KB) per CTA
  if (rpc < 1) rpc = 1;
  if (rpc > H) rpc = H;
  const int nc = (H + rpc - 1) / rpc;
  *rows_per_chunk = rpc;
  *n_chunks = nc;
  const long long planes = (long long)B * C;
  // part_v (f32) + part_xy (2 x i32) per (plane, chunk) + one ticket per plane
  *n_bytes = planes * nc * 12 + planes * 4 + 16;
  return SNB_OK;
}

template <typename T>
static int launch_global_peaks(const T* cms, int B, int C, int H, int W, long long sb, long long sc, long long sh,
                               long long sw, float threshold, int refine_size, int rpc, int nc, unsigned* tickets,
                               float* part_v, int* part_xy, const Ladder& lad, float* out_xy, float* out_val,
                               cudaStream_t st) {
  constexpr int PER = Elem<T>::PER16;
  const long long planes = (long long)B * C;
  const int vec = vec_ok_for(cms, Elem<T>::DT, W, sb, sc, sh, sw) ? 1 : 0;
  const long long chunkv = ((long long)rpc * W) / PER;  // 128-bit loads per chunk
  const unsigned grid = (unsigned)(planes * nc);
  bool force_generic = false, no_warp = false;
  int pad_smem = 0;
#ifdef SNB_AB_VARIANTS
  // A/B (tools/k2_probe.py): SNB_GLOBAL_GENERIC = the first, merge-per-element kernel; SNB_GLOBAL_NO_RING = one CTA per
  // plane with the values in registers; SNB_GLOBAL_RING = the persistent cp.async ring; SNB_GLOBAL_WARP_SMEM = bytes of
  // unused dynamic shared memory per warp CTA (caps residency: measured WORSE at every cap, 20.5-28.2 us vs 16.7 us).
  static const bool e_generic = getenv("SNB_GLOBAL_GENERIC") != nullptr;
  static const bool e_no_ring = getenv("SNB_GLOBAL_NO_RING") != nullptr;
  static const bool e_ring = getenv("SNB_GLOBAL_RING") != nullptr;
  static const int e_pad = getenv("SNB_GLOBAL_WARP_SMEM") ? atoi(getenv("SNB_GLOBAL_WARP_SMEM")) : 0;
  force_generic = e_generic;
  no_warp = e_no_ring || e_ring;
  pad_smem = e_pad;
  if constexpr (Elem<T>::DT == SNB_DTYPE_F32) {
    const size_t ring_smem = (size_t)GP_STAGES * H * W * sizeof(float);
    if (vec && e_ring && !force_generic && nc == 1 && ring_smem <= 96 * 1024 && planes < 0x7fffffffLL) {
      if (cudaFuncSetAttribute(global_peaks_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) !=
          cudaSuccess)
        return SNB_ERR_CUDA_LAUNCH;
      int per_sm = (int)((220 * 1024) / (ring_smem + 1024));
      per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);
      const long long want = (long long)sm_count() * per_sm;
      const unsigned rgrid = (unsigned)(planes < want ? planes : want);
      global_peaks_ring_kernel<<<rgrid, 256, ring_smem, st>>>(cms, (int)planes, C, H, W, sb, sc, sh, threshold,
                                                             refine_size, out_xy, out_val, lad);
      return SNB_OK;
    }
  }
#endif
  if (vec && !force_generic && !no_warp && (long long)H * W <= 16384 && planes < 0x7fffffffLL) {
    // small planes (cfg2's 80x80 crops): one warp per plane, no barrier anywhere
    cudaError_t err = cudaErrorInvalidValue;
#define SNB_GPW(U_, WPC_)                                                                                              \
  do {                                                                                                                 \
    const dim3 g((unsigned)((planes + (WPC_) - 1) / (WPC_)));                                                          \
    if (sh == W)                                                                                                       \
      err = launch_pdl(global_peaks_warp_kernel<T, U_, true, WPC_>, g, dim3(32 * (WPC_)), (size_t)pad_smem, st, cms,   \
                       (int)planes, C, H, W, sb, sc, sh, threshold, refine_size, out_xy, out_val, lad);                \
    else                                                                                                               \
      err = launch_pdl(global_peaks_warp_kernel<T, U_, false, WPC_>, g, dim3(32 * (WPC_)), (size_t)pad_smem, st, cms,  \
                       (int)planes, C, H, W, sb, sc, sh, threshold, refine_size, out_xy, out_val, lad);                \
  } while (0)
#ifdef SNB_AB_VARIANTS
    // A/B: SNB_K2_WPC in {1,2,4,8} warps per CTA, SNB_K2_U in {4,5,6,8,12} loads in flight per lane
    const int e_wpc = getenv("SNB_K2_WPC") ? atoi(getenv("SNB_K2_WPC")) : GP_WARPS_PER_CTA;
    const int e_u = getenv("SNB_K2_U") ? atoi(getenv("SNB_K2_U"))
                                       : (planes < 16LL * sm_count() ? GP_LOADS_IN_FLIGHT_SUBWAVE : GP_LOADS_IN_FLIGHT);
#define SNB_GPW_U(WPC_)                                                                                                \
  do {                                                                                                                 \
    if (e_u == 4) SNB_GPW(4, WPC_); else if (e_u == 6) SNB_GPW(6, WPC_); else if (e_u == 8) SNB_GPW(8, WPC_);          \
    else if (e_u == 10) SNB_GPW(10, WPC_); else if (e_u == 12) SNB_GPW(12, WPC_); else SNB_GPW(5, WPC_);               \
  } while (0)
    if (e_wpc == 1) SNB_GPW_U(1); else if (e_wpc == 2) SNB_GPW_U(2); else if (e_wpc == 8) SNB_GPW_U(8); else SNB_GPW_U(4);
#undef SNB_GPW_U
#else
    if (planes < 16LL * sm_count()) SNB_GPW(GP_LOADS_IN_FLIGHT_SUBWAVE, GP_WARPS_PER_CTA);
    else SNB_GPW(GP_LOADS_IN_FLIGHT, GP_WARPS_PER_CTA);
#endif
#undef SNB_GPW
    if (err != cudaSuccess) return SNB_ERR_CUDA_LAUNCH;
  } else if (vec && !force_generic && chunkv <= 8 * 256) {
#define SNB_GP(V)                                                                                                      \
  global_peaks_regs_kernel<T, V><<<grid, 256, 0, st>>>(cms, C, H, W, sb, sc, sh, rpc, nc, threshold, refine_size, part_v, \
                                                       part_xy, tickets, out_xy, out_val, lad)
    if (chunkv <= 2 * 256) SNB_GP(2);
    else if (chunkv <= 4 * 256) SNB_GP(4);
    else SNB_GP(8);
#undef SNB_GP
  } else {
    global_peaks_kernel<T><<<grid, 256, 0, st>>>(cms, C, H, W, sb, sc, sh, sw, vec, rpc, nc, threshold, refine_size,
                                                 part_v, part_xy, tickets, out_xy, out_val, lad);
  }
  return SNB_OK;
}

extern "C" int snb_global_peaks_t(const void* cms, int dtype, int B, int C, int H, int W, long long sb, long long sc,
                                  long long sh, long long sw, float threshold, int refine_size, void* workspace,
                                  const snb_coord_ladder* ladder, float* out_xy, float* out_val, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!dtype_ok(dtype)) return SNB_ERR_BAD_ARG;
  Ladder lad{1.f, 1.f, nullptr, nullptr, nullptr, nullptr};
  if (ladder) lad = Ladder{ladder->stride, ladder->input_scale, ladder->eff_scale, ladder->crop_offset, ladder->eff_scale2,
                           ladder->scatter};
  int rpc, nc;
  long long nbytes;
  int rc = snb_global_peaks_workspace(B, C, H, W, &rpc, &nc, &nbytes);
  if (rc != SNB_OK) return rc;
  if (B == 0) return SNB_OK;
  const long long planes = (long long)B * C;
  if (planes * nc > 0x7fffffffLL) return SNB_ERR_UNSUPPORTED;
  // workspace layout: tickets (zero on first use, self-resetting) | part_v | part_xy
  unsigned* tickets = (unsigned*)workspace;
  float* part_v = (float*)(tickets + planes);
  int* part_xy = (int*)(part_v + planes * nc);
  switch (dtype) {
    case SNB_DTYPE_F16:
      rc = launch_global_peaks((const __half*)cms, B, C, H, W, sb, sc, sh, sw, threshold, refine_size, rpc, nc, tickets,
                               part_v, part_xy, lad, out_xy, out_val, st);
      break;
    case SNB_DTYPE_BF16:
      rc = launch_global_peaks((const __nv_bfloat16*)cms, B, C, H, W, sb, sc, sh, sw, threshold, refine_size, rpc, nc,
                               tickets, part_v, part_xy, lad, out_xy, out_val, st);
      break;
    default:
      rc = launch_global_peaks((const float*)cms, B, C, H, W, sb, sc, sh, sw, threshold, refine_size, rpc, nc, tickets,
                               part_v, part_xy, lad, out_xy, out_val, st);
  }
  if (rc != SNB_OK) return rc;
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_global_peaks_ex(const float* cms, int B, int C, int H, int W, long long sb, long long sc, long long sh,
                                   long long sw, float threshold, int refine_size, void* workspace,
                                   const snb_coord_ladder* ladder, float* out_xy, float* out_val, void* stream_) {
  return snb_global_peaks_t(cms, SNB_DTYPE_F32, B, C, H, W, sb, sc, sh, sw, threshold, refine_size, workspace, ladder,
                            out_xy, out_val, stream_);
}

extern "C" int snb_global_peaks(const float* cms, int B, int C, int H, int W, long long sb, long long sc, long long sh,
                                long long sw, float threshold, int refine_size, void* workspace, float* out_xy,
                                float* out_val, void* stream_) {
  return snb_global_peaks_t(cms, SNB_DTYPE_F32, B, C, H, W, sb, sc, sh, sw, threshold, refine_size, workspace, nullptr,
                            out_xy, out_val, stream_);
}

extern "C" int snb_crop_bboxes(const void* images, int elem_size, int S, int C, int H, int W, long long sb,
                               long long sc, long long sh, long long sw, const float* bboxes,
                               const long long* sample_inds, long long n, int crop_h, int crop_w, void* out,
                               int* status, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n < 0 || C < 0 || crop_h < 0 || crop_w < 0) return SNB_ERR_BAD_ARG;
  const long long total = n * C * crop_h * crop_w;
  if (total == 0) return SNB_OK;
  const int grid = grid_for(total, 256, sm_count() * 16);
#define SNB_CROP(T)                                                                                              \
  crop_bboxes_kernel<T><<<grid, 256, 0, st>>>((const T*)images, S, C, H, W, sb, sc, sh, sw, bboxes, sample_inds, \
                                              n, crop_h, crop_w, (T*)out, status)
  switch (elem_size) {
    case 1: SNB_CROP(uint8_t); break;
    case 2: SNB_CROP(uint16_t); break;
    case 4: SNB_CROP(uint32_t); break;
    case 8: SNB_CROP(unsigned long long); break;
    default: return SNB_ERR_UNSUPPORTED;
  }
#undef SNB_CROP
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_integral_regression(const float* patches, long long n_planes, int h, int w, const float* xv,
                                       const float* yv, float* out_x, float* out_y, void* stream_) {
  if (n_planes <= 0) return SNB_OK;
  const long long threads = n_planes * 32;
  integral_regression_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      patches, n_planes, h, w, xv, yv, out_x, out_y);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_dilate8(const float* image, long long n_planes, int H, int W, float* out, void* stream_) {
  const long long total = n_planes * H * W;
  if (total <= 0) return SNB_OK;
  dilate8_kernel<<<grid_for(total, 256, sm_count() * 16), 256, 0, (cudaStream_t)stream_>>>(image, n_planes, H, W, out);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_centered_bboxes(const float* centers, long long n, float half_h, float half_w, float* out,
                                   void* stream_) {
  if (n <= 0) return SNB_OK;
  centered_bboxes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(centers, n, half_h, half_w,
                                                                                         out);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_peaks_topk(const int* frame_count, int B, int cap, const float* xy, const float* val,
                              int max_instances, float input_scale, const float* eff_scale, float* out_xy,
                              float* out_val, void* stream_) {
  if (B < 0 || cap <= 0 || max_instances <= 0) return SNB_ERR_BAD_ARG;
  if (B == 0) return SNB_OK;
  const size_t smem = sizeof(int) * (size_t)max_instances;
  if (smem > 48 * 1024) return SNB_ERR_UNSUPPORTED;
  peaks_topk_kernel<<<B, 128, smem, (cudaStream_t)stream_>>>(frame_count, cap, xy, val, max_instances, input_scale,
                                                            eff_scale, out_xy, out_val);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_coord_ladder_apply(const float* xy, long long n_samples, long long pairs_per_sample,
                                      const snb_coord_ladder* ladder, float* out, void* stream_) {
  if (!ladder || n_samples < 0 || pairs_per_sample < 0) return SNB_ERR_BAD_ARG;
  const long long n = n_samples * pairs_per_sample;
  if (n == 0) return SNB_OK;
  const Ladder lad{ladder->stride, ladder->input_scale, ladder->eff_scale, ladder->crop_offset, ladder->eff_scale2,
                   nullptr};
  coord_ladder_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(xy, n, pairs_per_sample, lad, out);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

#ifdef SNB_AB_VARIANTS
// A/B build only (not in include/sleapnn_b200.h): switch the K2 stamps on / off and copy them out.
extern "C" int snb_ab_k2_stamps(int on, unsigned long long* host_out, int n_planes) {
  if (cudaDeviceSynchronize() != cudaSuccess) return SNB_ERR_CUDA_LAUNCH;
  if (host_out && n_planes > 0) {
    if (n_planes > snb::K2_STAMP_PLANES) n_planes = snb::K2_STAMP_PLANES;
    if (cudaMemcpyFromSymbol(host_out, snb::g_k2_stamps, sizeof(unsigned long long) * 3 * (size_t)n_planes) != cudaSuccess)
      return SNB_ERR_CUDA_LAUNCH;
  }
  if (cudaMemcpyToSymbol(snb::g_k2_stamps_on, &on, sizeof(int)) != cudaSuccess) return SNB_ERR_CUDA_LAUNCH;
  return SNB_OK;
}
#endif
