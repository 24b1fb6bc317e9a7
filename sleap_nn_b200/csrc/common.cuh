// Shared device helpers for the sm_100a heat-map path kernels.
//
// Arithmetic rule for everything that feeds an integer decision (line subscripts, crop
// corners, thresholds): every fp32 product / sum / quotient is a SEPARATE IEEE rounding
// (__fmul_rn / __fadd_rn / __fdiv_rn), never contracted into FMA, because the reference
// is a chain of individually rounded ATen CPU ops.  The library is
// built WITHOUT --use_fast_math and with -ftz=false -prec-div=true -prec-sqrt=true.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/sleapnn_b200.h"

#define SNB_LAUNCH_CHECK()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return SNB_ERR_CUDA_LAUNCH;      \
  } while (0)

namespace snb {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  // 128-bit read-only streaming load; the confidence maps are read exactly once.
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t smem_u32_early(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ int div_up(int a, int b) { return (a + b - 1) / b; }

// Integral-regression offsets on a size x size patch whose centre tap is the integer
// peak (px, py) of plane `plane` (row stride sh, col stride sw, in elements).
// Patch top-left follows crop_bboxes (ops/crops.py:85-90): trunc((p - s/2 + 0.5) + s//2) - s//2;
// taps outside the image read 0 (ops/crops.py:97-99).  Grid = arange(s) - (s-1)/2
// (ops/peaks.py:174).  Sums are accumulated in fp64 and rounded once to fp32 before the
// fp32 division that the reference performs (ops/peaks.py:84-85); 0/0 -> NaN as there.
__device__ __forceinline__ void integral_refine(const float* __restrict__ plane, int H, int W, long long sh,
                                                long long sw, float px, float py, int size, float* ox,
                                                float* oy) {
  const float half_f = 0.5f * (float)size;           // box / 2 (python float, exact in fp32 for int size)
  const int half_i = size / 2;                       // box // 2
  // make_centered_bboxes: (x - half) + 0.5, each rounded in fp32 (instance_cropping.py:151-171)
  const float tlx_f = __fadd_rn(__fadd_rn(__fsub_rn(px, half_f), 0.5f), (float)half_i);
  const float tly_f = __fadd_rn(__fadd_rn(__fsub_rn(py, half_f), 0.5f), (float)half_i);
  const int x0 = (int)truncf(tlx_f) - half_i;
  const int y0 = (int)truncf(tly_f) - half_i;
  const float g0 = -0.5f * (float)(size - 1);        // (size-1)/2 exact in fp32
  double z = 0.0, sx = 0.0, sy = 0.0;
  for (int j = 0; j < size; ++j) {
    const int yy = y0 + j;
    const bool yin = (yy >= 0) && (yy < H);
    const float gy = g0 + (float)j;
    for (int i = 0; i < size; ++i) {
      const int xx = x0 + i;
      float p = 0.f;
      if (yin && xx >= 0 && xx < W) p = __ldg(plane + (long long)yy * sh + (long long)xx * sw);
      const float gx = g0 + (float)i;
      z += (double)p;
      sx += (double)__fmul_rn(gx, p);
      sy += (double)__fmul_rn(gy, p);
    }
  }
  const float zf = (float)z;
  *ox = __fdiv_rn((float)sx, zf);
  *oy = __fdiv_rn((float)sy, zf);
}

// Warp-cooperative variant: the size x size taps are spread over the 32 lanes (one DRAM/L2 round
// trip instead of size^2 serial ones); fp64 partial sums are combined by shuffles.  All lanes
// return the same offsets.
template <bool GLOBAL_MEM = true>  // false: `plane` points into shared memory (plain loads, not ld.global.nc)
__device__ __forceinline__ void integral_refine_warp(const float* __restrict__ plane, int H, int W, long long sh,
                                                     long long sw, float px, float py, int size, int lane, float* ox,
                                                     float* oy) {
  const float half_f = 0.5f * (float)size;
  const int half_i = size / 2;
  const float tlx_f = __fadd_rn(__fadd_rn(__fsub_rn(px, half_f), 0.5f), (float)half_i);
  const float tly_f = __fadd_rn(__fadd_rn(__fsub_rn(py, half_f), 0.5f), (float)half_i);
  const int x0 = (int)truncf(tlx_f) - half_i;
  const int y0 = (int)truncf(tly_f) - half_i;
  const float g0 = -0.5f * (float)(size - 1);
  double z = 0.0, sx = 0.0, sy = 0.0;
  for (int t = lane; t < size * size; t += 32) {
    const int j = t / size, i = t - j * size;
    const int yy = y0 + j, xx = x0 + i;
    float p = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const float* q = plane + (long long)yy * sh + (long long)xx * sw;
      p = GLOBAL_MEM ? __ldg(q) : *q;
    }
    z += (double)p;
    sx += (double)__fmul_rn(g0 + (float)i, p);
    sy += (double)__fmul_rn(g0 + (float)j, p);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    z += __shfl_xor_sync(FULL, z, d);
    sx += __shfl_xor_sync(FULL, sx, d);
    sy += __shfl_xor_sync(FULL, sy, d);
  }
  const float zf = (float)z;
  *ox = __fdiv_rn((float)sx, zf);
  *oy = __fdiv_rn((float)sy, zf);
}

// Q peaks at once: the Q x size^2 taps are all in flight before anything is reduced, so a warp that owns many
// peaks (busy frames: hundreds of peaks per frame) pays one memory round trip per Q peaks instead of one per peak.
// Same arithmetic as integral_refine_warp.  plane[q] == nullptr marks an unused slot (offsets returned as 0).
template <int Q>
__device__ __forceinline__ void integral_refine_warp_multi(const float* const (&plane)[Q], int H, int W, long long sh,
                                                           long long sw, const float (&px)[Q], const float (&py)[Q],
                                                           int size, int lane, float (&ox)[Q], float (&oy)[Q]) {
  const float half_f = 0.5f * (float)size;
  const int half_i = size / 2;
  const float g0 = -0.5f * (float)(size - 1);
  int x0[Q], y0[Q];
  double z[Q], sx[Q], sy[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    x0[q] = (int)truncf(__fadd_rn(__fadd_rn(__fsub_rn(px[q], half_f), 0.5f), (float)half_i)) - half_i;
    y0[q] = (int)truncf(__fadd_rn(__fadd_rn(__fsub_rn(py[q], half_f), 0.5f), (float)half_i)) - half_i;
    z[q] = sx[q] = sy[q] = 0.0;
  }
  for (int t = lane; t < size * size; t += 32) {
    const int j = t / size, i = t - j * size;
    float p[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int yy = y0[q] + j, xx = x0[q] + i;
      p[q] = 0.f;
      if (plane[q] && yy >= 0 && yy < H && xx >= 0 && xx < W) p[q] = __ldg(plane[q] + (long long)yy * sh + (long long)xx * sw);
    }
    const float gx = g0 + (float)i, gy = g0 + (float)j;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      z[q] += (double)p[q];
      sx[q] += (double)__fmul_rn(gx, p[q]);
      sy[q] += (double)__fmul_rn(gy, p[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < Q; ++q) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      z[q] += __shfl_xor_sync(FULL, z[q], d);
      sx[q] += __shfl_xor_sync(FULL, sx[q], d);
      sy[q] += __shfl_xor_sync(FULL, sy[q], d);
    }
    const float zf = (float)z[q];
    ox[q] = __fdiv_rn((float)sx[q], zf);
    oy[q] = __fdiv_rn((float)sy[q], zf);
  }
}

// ---- the coordinate ladder of the inference layers (inference/ops/coord.py:27-90), fused into the peak
// kernels' epilogues.  Each step is one separately rounded fp32 op, exactly like the tensor op it replaces;
// x * 1.0f and x / 1.0f are identities, so the reference's `== 1` short-circuits need no branch.
struct Ladder {
  float stride;          // undo_stride:      xy * stride
  float input_scale;     // undo_input_scale: xy / input_scale
  const float* eff;      // undo_eff_scale:   xy / eff[b]        (NULL = skip)
  const float* off;      // add_crop_offset:  xy + off[b]        (NULL = skip)
  const float* eff2;     // TopDownLayer:     xy / eff2[b] after the crop offset (NULL = skip)
  const int* scatter;    // output row of sample b (NULL = b; negative = drop)
};
__device__ __forceinline__ Ladder ladder_identity() { return Ladder{1.f, 1.f, nullptr, nullptr, nullptr, nullptr}; }
__device__ __forceinline__ void ladder_apply(const Ladder& L, int b, float& x, float& y) {
  x = __fdiv_rn(__fmul_rn(x, L.stride), L.input_scale);
  y = __fdiv_rn(__fmul_rn(y, L.stride), L.input_scale);
  if (L.eff) { const float e = L.eff[b]; x = __fdiv_rn(x, e); y = __fdiv_rn(y, e); }
  if (L.off) { x = __fadd_rn(x, L.off[2 * b]); y = __fadd_rn(y, L.off[2 * b + 1]); }
  if (L.eff2) { const float e = L.eff2[b]; x = __fdiv_rn(x, e); y = __fdiv_rn(y, e); }
}

// ---- cp.async (LDGSTS): 16-byte global -> shared copies that tie up no registers ----------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32_early(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- mbarrier + bulk async copy (cp.async.bulk, the 1-D TMA path; SASS: UBLKCP) -------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

}  // namespace snb
