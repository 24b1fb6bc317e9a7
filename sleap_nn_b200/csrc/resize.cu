// apply_input_scale (sleap_nn/inference/ops/coord.py:93-109): bilinear resize of an image batch to
// (int(H * s), int(W * s)) - what F.interpolate(mode="bilinear", align_corners=False) computes, without antialiasing.
// Source index of output d along an axis of in -> out samples (ATen area_pixel_compute_source_index):
//   src = max((in / out) * (d + 0.5) - 0.5, 0),  i0 = floor(src),  i1 = i0 + (i0 < in - 1),  l1 = src - i0,  l0 = 1 - l1
//   out = l0y * (l0x * p00 + l1x * p01) + l1y * (l0x * p10 + l1x * p11)
// Pure HBM stream (4 gathered taps per output, neighbours share sectors): one thread per output pixel, x fastest.
// ATen's CPU kernel is vectorised with FMA, so agreement with it is to ~1 ulp of the result, not bit for bit.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace snb {

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256)
bilinear_resize_kernel(const T* __restrict__ in, long long planes, int H, int W, long long sp, long long sh,
                       long long sw, int oh, int ow, float scale_y, float scale_x, T* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = planes * oh * ow;
  if (t >= total) return;
  const int ox = (int)(t % ow);
  const long long r = t / ow;
  const int oy = (int)(r % oh);
  const long long p = r / oh;
  const float sy = fmaxf(scale_y * ((float)oy + 0.5f) - 0.5f, 0.f);
  const float sx = fmaxf(scale_x * ((float)ox + 0.5f) - 0.5f, 0.f);
  const int y0 = min((int)sy, H - 1), x0 = min((int)sx, W - 1);
  const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
  const float ly1 = sy - (float)y0, lx1 = sx - (float)x0;
  const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const T* pl = in + p * sp;
  const float p00 = to_f32<T>(pl[y0 * sh + x0 * sw]), p01 = to_f32<T>(pl[y0 * sh + x1 * sw]);
  const float p10 = to_f32<T>(pl[y1 * sh + x0 * sw]), p11 = to_f32<T>(pl[y1 * sh + x1 * sw]);
  out[t] = from_f32<T>(ly0 * (lx0 * p00 + lx1 * p01) + ly1 * (lx0 * p10 + lx1 * p11));
}

}  // namespace snb

using namespace snb;

// dtype: 0 = fp32, 1 = fp16, 2 = bf16.  in: `planes` planes of H x W with element strides (sp, sh, sw); out contiguous.
extern "C" int snb_bilinear_resize(const void* in, int dtype, long long planes, int H, int W, long long sp, long long sh,
                                   long long sw, int oh, int ow, void* out, void* stream_) {
  if (planes < 0 || H <= 0 || W <= 0 || oh < 0 || ow < 0 || dtype < 0 || dtype > 2) return SNB_ERR_BAD_ARG;
  const long long total = planes * oh * ow;
  if (total == 0) return SNB_OK;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream_;
  const float scale_y = (float)H / (float)oh, scale_x = (float)W / (float)ow;
  if (dtype == 0)
    bilinear_resize_kernel<float><<<blocks, 256, 0, st>>>((const float*)in, planes, H, W, sp, sh, sw, oh, ow, scale_y,
                                                         scale_x, (float*)out);
  else if (dtype == 1)
    bilinear_resize_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)in, planes, H, W, sp, sh, sw, oh, ow, scale_y,
                                                          scale_x, (__half*)out);
  else
    bilinear_resize_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)in, planes, H, W, sp, sh, sw, oh,
                                                                 ow, scale_y, scale_x, (__nv_bfloat16*)out);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
