// Shared device helpers for the sm_100a heat-map path kernels.
//
// Arithmetic rule for everything that feeds an integer decision (line subscripts, crop
// corners, thresholds): every fp32 product / sum / quotient is a SEPARATE IEEE rounding
// (__fmul_rn / __fadd_rn / __fdiv_rn), never contracted into FMA, because the reference
// is a chain of individually rounded ATen CPU ops.  The library is
// built WITHOUT --use_fast_math and with -ftz=false -prec-div=true -prec-sqrt=true.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
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

// ---- programmatic dependent launch (PDL, sm_90+).  Short one-wave kernels that are issued back to back on one stream
// (per-frame target launches from a data loader, global-peak launches per crop batch) otherwise pay the full
// launch + ramp latency after the previous kernel has drained.  A kernel launched through launch_pdl() may become
// resident while the previous kernel's last wave is still running; it must call pdl_wait() before its first global
// memory access (the wait returns when the previous grid has completed and its writes are visible, so there is no
// hazard on inputs or outputs) and calls pdl_launch_dependents() first thing so that its own successor can do the same.
// A predecessor that knows nothing about PDL simply releases its dependents when it exits.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  // 128-bit read-only streaming load; the confidence maps are read exactly once.
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// ---- element types of the maps (ABI v5: SNB_DTYPE_*).  Under autocast the backbone emits fp16 / bf16 heads
// (layers/backends/torch_backend.py:125-146 casts them back with .float()); every value of those types is exactly
// representable in fp32, so reading them natively and up-casting in registers gives bit-identical results to the
// reference's .float() + fp32 ops while moving half the bytes.
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float fmax_nan(float a, float b) {  // maximum that PROPAGATES NaN (torch.max semantics)
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float bf16_bits_to_float(unsigned short u) { return __uint_as_float((unsigned)u << 16); }
__device__ __forceinline__ float f16_bits_to_float(unsigned short u) { return __half2float(__ushort_as_half(u)); }

// One element of a tensor whose type is only known at run time (the latency-bound kernels: O(#peaks) taps).
// `dt` is warp-uniform, so the branch does not diverge.
__device__ __forceinline__ float ld_elem(const void* base, long long idx, int dt) {
  if (dt == SNB_DTYPE_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
  const unsigned short u = __ldg(reinterpret_cast<const unsigned short*>(base) + idx);
  return dt == SNB_DTYPE_F16 ? f16_bits_to_float(u) : bf16_bits_to_float(u);
}
__host__ __device__ __forceinline__ int dtype_size(int dt) { return dt == SNB_DTYPE_F32 ? 4 : 2; }
// address of element `off` of a run-time typed tensor
__device__ __forceinline__ const void* elem_ptr(const void* base, long long off, int dt) {
  return reinterpret_cast<const char*>(base) + off * dtype_size(dt);
}

// Compile-time element traits for the streaming kernels: PER16 elements per 128-bit load.
template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr int PER16 = 4, DT = SNB_DTYPE_F32;
  struct Thr { float f; };
  static __device__ __forceinline__ Thr make_thr(float thr) { return Thr{thr}; }
  static __device__ __forceinline__ void unpack(const uint4& v, float (&e)[4]) {
    e[0] = __uint_as_float(v.x); e[1] = __uint_as_float(v.y); e[2] = __uint_as_float(v.z); e[3] = __uint_as_float(v.w);
  }
  static __device__ __forceinline__ bool any_gt(const uint4& v, const Thr& t) {
    return (__uint_as_float(v.x) > t.f) || (__uint_as_float(v.y) > t.f) || (__uint_as_float(v.z) > t.f) ||
           (__uint_as_float(v.w) > t.f);
  }
  // maximum of the vector with NaN propagation (max.NaN: a NaN sticks)
  static __device__ __forceinline__ float vmax_nan(const uint4& v) {
    float a, b, r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(a) : "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)));
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(b) : "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)));
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
  }
  static __device__ __forceinline__ float load1(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ uint4 neg_inf() {
    const unsigned u = 0xff800000u;
    return make_uint4(u, u, u, u);
  }
};
template <> struct Elem<__half> {
  static constexpr int PER16 = 8, DT = SNB_DTYPE_F16;
  // `v > thr` for a half v and an fp32 thr  <=>  v > (thr rounded DOWN to half): no half lies in (rd(thr), thr].
  struct Thr { float f; __half2 h2; };
  static __device__ __forceinline__ Thr make_thr(float thr) {
    const __half h = __float2half_rd(thr);
    return Thr{thr, __halves2half2(h, h)};
  }
  static __device__ __forceinline__ __half2 h2(unsigned u) { return *reinterpret_cast<const __half2*>(&u); }
  static __device__ __forceinline__ void unpack(const uint4& v, float (&e)[8]) {
    const float2 a = __half22float2(h2(v.x)), b = __half22float2(h2(v.y)), c = __half22float2(h2(v.z)),
                 d = __half22float2(h2(v.w));
    e[0] = a.x; e[1] = a.y; e[2] = b.x; e[3] = b.y; e[4] = c.x; e[5] = c.y; e[6] = d.x; e[7] = d.y;
  }
  static __device__ __forceinline__ bool any_gt(const uint4& v, const Thr& t) {
    // __hmax2 returns the non-NaN operand: a NaN element never passes `v > thr`, so ignoring it is exact
    const __half2 m = __hmax2(__hmax2(h2(v.x), h2(v.y)), __hmax2(h2(v.z), h2(v.w)));
    return __hgt2_mask(m, t.h2) != 0u;
  }
  static __device__ __forceinline__ float vmax_nan(const uint4& v) {
    const __half2 m = __hmax2_nan(__hmax2_nan(h2(v.x), h2(v.y)), __hmax2_nan(h2(v.z), h2(v.w)));
    const float2 f = __half22float2(m);
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(f.x), "f"(f.y));
    return r;
  }
  static __device__ __forceinline__ float load1(const __half* p) {
    return f16_bits_to_float(__ldg(reinterpret_cast<const unsigned short*>(p)));
  }
  static __device__ __forceinline__ uint4 neg_inf() {
    const unsigned u = 0xfc00fc00u;
    return make_uint4(u, u, u, u);
  }
};
template <> struct Elem<__nv_bfloat16> {
  static constexpr int PER16 = 8, DT = SNB_DTYPE_BF16;
  struct Thr { float f; __nv_bfloat162 h2; };
  static __device__ __forceinline__ Thr make_thr(float thr) {
    const __nv_bfloat16 h = __float2bfloat16_rd(thr);
    return Thr{thr, __halves2bfloat162(h, h)};
  }
  static __device__ __forceinline__ __nv_bfloat162 h2(unsigned u) { return *reinterpret_cast<const __nv_bfloat162*>(&u); }
  static __device__ __forceinline__ void unpack(const uint4& v, float (&e)[8]) {
    // bf16 -> fp32 is a 16-bit shift
    e[0] = __uint_as_float(v.x << 16); e[1] = __uint_as_float(v.x & 0xffff0000u);
    e[2] = __uint_as_float(v.y << 16); e[3] = __uint_as_float(v.y & 0xffff0000u);
    e[4] = __uint_as_float(v.z << 16); e[5] = __uint_as_float(v.z & 0xffff0000u);
    e[6] = __uint_as_float(v.w << 16); e[7] = __uint_as_float(v.w & 0xffff0000u);
  }
  static __device__ __forceinline__ bool any_gt(const uint4& v, const Thr& t) {
    const __nv_bfloat162 m = __hmax2(__hmax2(h2(v.x), h2(v.y)), __hmax2(h2(v.z), h2(v.w)));
    return __hgt2_mask(m, t.h2) != 0u;
  }
  static __device__ __forceinline__ float vmax_nan(const uint4& v) {
    const __nv_bfloat162 m = __hmax2_nan(__hmax2_nan(h2(v.x), h2(v.y)), __hmax2_nan(h2(v.z), h2(v.w)));
    const unsigned u = *reinterpret_cast<const unsigned*>(&m);
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(__uint_as_float(u << 16)), "f"(__uint_as_float(u & 0xffff0000u)));
    return r;
  }
  static __device__ __forceinline__ float load1(const __nv_bfloat16* p) {
    return bf16_bits_to_float(__ldg(reinterpret_cast<const unsigned short*>(p)));
  }
  static __device__ __forceinline__ uint4 neg_inf() {
    const unsigned u = 0xff80ff80u;
    return make_uint4(u, u, u, u);
  }
};

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
__device__ __forceinline__ void integral_refine(const void* __restrict__ plane, int dt, int H, int W, long long sh,
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
      if (yin && xx >= 0 && xx < W) p = ld_elem(plane, (long long)yy * sh + (long long)xx * sw, dt);
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
template <bool GLOBAL_MEM = true>  // false: `plane` points into shared memory (fp32, plain loads, not ld.global.nc)
__device__ __forceinline__ void integral_refine_warp(const void* __restrict__ plane, int dt, int H, int W, long long sh,
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
      const long long q = (long long)yy * sh + (long long)xx * sw;
      p = GLOBAL_MEM ? ld_elem(plane, q, dt) : reinterpret_cast<const float*>(plane)[q];
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
__device__ __forceinline__ void integral_refine_warp_multi(const void* const (&plane)[Q], int dt, int H, int W,
                                                           long long sh, long long sw, const float (&px)[Q],
                                                           const float (&py)[Q], int size, int lane, float (&ox)[Q],
                                                           float (&oy)[Q]) {
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
      if (plane[q] && yy >= 0 && yy < H && xx >= 0 && xx < W) p[q] = ld_elem(plane[q], (long long)yy * sh + (long long)xx * sw, dt);
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
