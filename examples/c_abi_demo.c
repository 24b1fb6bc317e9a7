/* The C ABI without Python or torch: a plain C program (gcc + libcudart) that builds a two-channel confidence map on
 * the host, copies it to the device and calls libsleapnn_b200.so the way a non-Python host would:
 *   snb_local_peaks  (find_local_peaks: 3x3 NMS + threshold + ordered emission + integral refinement)
 *   snb_global_peaks (find_global_peaks)
 *   snb_confmaps     (make_multi_confmaps) -> fed back into snb_local_peaks: the planted points come back.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/c_abi_demo.c -o /tmp/c_abi_demo \
 *       -L sleap_nn_b200/lib -lsleapnn_b200 -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/sleap_nn_b200/lib
 *
 * Prints one line per result and "C ABI demo: OK"; exit status 1 on any mismatch.  tests/test_abi.py compiles and links it
 * on every box and runs it where a GPU is present.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sleapnn_b200.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define CHECK_SNB(x) do { int r_ = (x); if (r_ != SNB_OK) { fprintf(stderr, "%s -> %d\n", #x, r_); return 1; } } while (0)

int main(void) {
  if (snb_abi_version() != SNB_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }
  enum { B = 1, C = 2, H = 48, W = 64, CAP = 16 };
  /* three planted points: (x, y, channel); make_multi_confmaps takes (G, I, N, 2) with NaN = absent */
  const float planted[3][3] = {{10.f, 12.f, 0.f}, {40.f, 30.f, 0.f}, {50.f, 8.f, 1.f}};
  float pts[2 /*I*/][C][2];
  for (int i = 0; i < 2; ++i) for (int c = 0; c < C; ++c) pts[i][c][0] = pts[i][c][1] = NAN;
  pts[0][0][0] = planted[0][0]; pts[0][0][1] = planted[0][1];
  pts[1][0][0] = planted[1][0]; pts[1][0][1] = planted[1][1];
  pts[0][1][0] = planted[2][0]; pts[0][1][1] = planted[2][1];
  float xv[W], yv[H];
  for (int x = 0; x < W; ++x) xv[x] = (float)x;
  for (int y = 0; y < H; ++y) yv[y] = (float)y;

  float *d_pts, *d_xv, *d_yv, *d_cms, *d_xy, *d_val, *d_gxy, *d_gval;
  int *d_count, *d_chan, *d_status;
  uint32_t* d_keys;
  void* d_ws;
  CHECK_CUDA(cudaMalloc((void**)&d_pts, sizeof(pts)));
  CHECK_CUDA(cudaMalloc((void**)&d_xv, sizeof(xv)));
  CHECK_CUDA(cudaMalloc((void**)&d_yv, sizeof(yv)));
  CHECK_CUDA(cudaMalloc((void**)&d_cms, sizeof(float) * B * C * H * W));
  CHECK_CUDA(cudaMalloc((void**)&d_xy, sizeof(float) * B * CAP * 2));
  CHECK_CUDA(cudaMalloc((void**)&d_val, sizeof(float) * B * CAP));
  CHECK_CUDA(cudaMalloc((void**)&d_chan, sizeof(int) * B * CAP));
  CHECK_CUDA(cudaMalloc((void**)&d_keys, sizeof(uint32_t) * B * CAP));
  CHECK_CUDA(cudaMalloc((void**)&d_count, sizeof(int) * B));
  CHECK_CUDA(cudaMalloc((void**)&d_status, sizeof(int)));
  CHECK_CUDA(cudaMalloc((void**)&d_gxy, sizeof(float) * B * C * 2));
  CHECK_CUDA(cudaMalloc((void**)&d_gval, sizeof(float) * B * C));
  CHECK_CUDA(cudaMemset(d_status, 0, sizeof(int)));
  CHECK_CUDA(cudaMemcpy(d_pts, pts, sizeof(pts), cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_xv, xv, sizeof(xv), cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_yv, yv, sizeof(yv), cudaMemcpyHostToDevice));

  /* make_multi_confmaps(points (1, 2, 2, 2), xv, yv, sigma = 2): den = 2 sigma^2 */
  CHECK_SNB(snb_confmaps(d_pts, 1, 2, C, d_xv, d_yv, H, W, 8.0f, 0, d_cms, NULL));
  /* find_local_peaks(cms, threshold = 0.2, refinement = "integral", integral_patch_size = 5) */
  CHECK_SNB(snb_local_peaks(d_cms, B, C, H, W, (long long)C * H * W, (long long)H * W, W, 1, 0.2f, 5, 1.0f, CAP, d_count,
                            d_keys, d_xy, d_val, d_chan, d_status, NULL));
  /* find_global_peaks(cms, threshold = 0.2, refinement = "integral") */
  int rpc, nch;
  long long ws_bytes;
  CHECK_SNB(snb_global_peaks_workspace(B, C, H, W, &rpc, &nch, &ws_bytes));
  CHECK_CUDA(cudaMalloc(&d_ws, (size_t)(ws_bytes > 0 ? ws_bytes : 4)));
  CHECK_CUDA(cudaMemset(d_ws, 0, (size_t)(ws_bytes > 0 ? ws_bytes : 4)));
  CHECK_SNB(snb_global_peaks(d_cms, B, C, H, W, (long long)C * H * W, (long long)H * W, W, 1, 0.2f, 5, d_ws, d_gxy, d_gval,
                             NULL));
  CHECK_CUDA(cudaDeviceSynchronize());

  int count, status, chan[CAP];
  float xy[CAP][2], val[CAP], gxy[C][2], gval[C];
  CHECK_CUDA(cudaMemcpy(&count, d_count, sizeof(int), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(xy, d_xy, sizeof(xy), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(val, d_val, sizeof(val), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(chan, d_chan, sizeof(chan), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(gxy, d_gxy, sizeof(gxy), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(gval, d_gval, sizeof(gval), cudaMemcpyDeviceToHost));

  int bad = (count != 3) || (status != 0);
  printf("local peaks: %d (status %d)\n", count, status);
  /* reference order: ascending (y, x, channel) -> (50, 8, ch 1), (10, 12, ch 0), (40, 30, ch 0) */
  const int order[3] = {2, 0, 1};
  for (int k = 0; k < count && k < 3; ++k) {
    const float* p = planted[order[k]];
    printf("  peak %d: x %.4f y %.4f value %.6f channel %d\n", k, xy[k][0], xy[k][1], val[k], chan[k]);
    bad |= fabsf(xy[k][0] - p[0]) > 1e-3f || fabsf(xy[k][1] - p[1]) > 1e-3f || chan[k] != (int)p[2] || val[k] != 1.0f;
  }
  for (int c = 0; c < C; ++c) printf("global peak, channel %d: x %.4f y %.4f value %.6f\n", c, gxy[c][0], gxy[c][1], gval[c]);
  /* channel 0 holds two equal maxima: the reference's two independent arg-maxes give min x and min y separately */
  bad |= gval[0] != 1.0f || gval[1] != 1.0f || fabsf(gxy[1][0] - 50.f) > 1e-3f || fabsf(gxy[1][1] - 8.f) > 1e-3f;
  puts(bad ? "C ABI demo: MISMATCH" : "C ABI demo: OK");
  return bad;
}
