// esb200 — input-side kernels.
//  (1) image normalisation = Det3DDataPreprocessor.preprocess_img + multiview_img_stack_batch
//      (embodiedscan/models/data_preprocessors/data_preprocessor.py:249-264, utils.py:9-63): BGR->RGB,
//      (x-mean)/std in fp32, right/bottom zero pad to the /32 shape, written NCHW or channels-last in one pass
//      (reads 1 B/px/channel, writes e B): pure HBM streaming.
//  (2) depth -> point unprojection = ConvertRGBDToPoints + points_img2cam + AggregateMultiViewPoints
//      (embodiedscan/datasets/transforms/points.py:30-81, structures/bbox_3d/utils.py:335-368,
//      datasets/transforms/multiview.py:139-169): [u*d, v*d, d, 1] through one host-composed 4x4
//      (E^-1 . K^-1), zero-depth pixels dropped, row-major order kept via flag + exclusive scan.
#include "common.cuh"
#include <cub/cub.cuh>

namespace {

template <typename T>
__global__ void img_normalize_kernel(const unsigned char* __restrict__ src, int n_img, int H, int W, int Hp, int Wp,
                                     float m0, float m1, float m2, float s0, float s1, float s2, int bgr_to_rgb,
                                     int channels_last, T* __restrict__ dst) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)n_img * Hp * Wp;
  if (t >= total) return;
  int x = (int)(t % Wp);
  int y = (int)((t / Wp) % Hp);
  long long n = t / ((long long)Wp * Hp);
  float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
  float v[3] = {0.f, 0.f, 0.f};
  if (x < W && y < H) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int sc = bgr_to_rgb ? 2 - c : c;
      float px = (float)src[((n * 3 + sc) * H + y) * W + x];
      v[c] = __fdiv_rn(__fsub_rn(px, mean[c]), stdv[c]);
    }
  }
  if (channels_last) {
    T* d = dst + ((n * Hp + y) * Wp + x) * 3;
    d[0] = esb_from_float<T>(v[0]); d[1] = esb_from_float<T>(v[1]); d[2] = esb_from_float<T>(v[2]);
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[((n * 3 + c) * Hp + y) * Wp + x] = esb_from_float<T>(v[c]);
  }
}

__global__ void depth_flag_kernel(const unsigned short* __restrict__ depth, long long n, int* __restrict__ flag) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < n) flag[t] = depth[t] != 0 ? 1 : 0;
}

// mats: (V,16) row-major cam-pixel -> world ; out: compacted (count,3) ; view_of: optional (count) view index
__global__ void unproject_kernel(const unsigned short* __restrict__ depth, int V, int H, int W, float depth_shift,
                                 const float* __restrict__ mats, const int* __restrict__ flag,
                                 const int* __restrict__ rank, float* __restrict__ out, int* __restrict__ view_of,
                                 int* __restrict__ count) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)V * H * W;
  if (t >= total) return;
  if (t == total - 1) *count = rank[t] + flag[t];
  if (!flag[t]) return;
  int u = (int)(t % W);
  int vv = (int)((t / W) % H);
  int view = (int)(t / ((long long)W * H));
  float d = __fdiv_rn((float)depth[t], depth_shift);
  float a = __fmul_rn((float)u, d), b = __fmul_rn((float)vv, d);
  const float* M = mats + view * 16;
  float x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, M[0]), __fmul_rn(b, M[1])), __fmul_rn(d, M[2])), M[3]);
  float y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, M[4]), __fmul_rn(b, M[5])), __fmul_rn(d, M[6])), M[7]);
  float z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, M[8]), __fmul_rn(b, M[9])), __fmul_rn(d, M[10])), M[11]);
  int r = rank[t];
  out[3 * (long long)r] = x;
  out[3 * (long long)r + 1] = y;
  out[3 * (long long)r + 2] = z;
  if (view_of) view_of[r] = view;
}

static size_t scan_bytes(long long n) {
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, (const int*)nullptr, (int*)nullptr, (int)n);
  return tb;
}

}  // namespace

// src: (n_img,3,H,W) uint8 ; dst: (n_img,3,Hp,Wp) NCHW or (n_img,Hp,Wp,3) channels-last, dtype f32/bf16
extern "C" int esb_img_normalize(const unsigned char* src, int n_img, int H, int W, int Hp, int Wp, const float* mean3,
                                 const float* std3, int bgr_to_rgb, int channels_last, void* dst, int dtype,
                                 void* stream) {
  ESB_CHECK_ARG(Hp >= H && Wp >= W, "esb_img_normalize: padded shape smaller than the image");
  long long total = (long long)n_img * Hp * Wp;
  if (total == 0) return ESB_OK;
  int grid = esb_div_up(total, 256);
  if (dtype == ESB_F32)
    img_normalize_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(src, n_img, H, W, Hp, Wp, mean3[0], mean3[1], mean3[2],
                                                                         std3[0], std3[1], std3[2], bgr_to_rgb,
                                                                         channels_last, (float*)dst);
  else
    img_normalize_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        src, n_img, H, W, Hp, Wp, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], bgr_to_rgb, channels_last,
        (__nv_bfloat16*)dst);
  ESB_CUDA_LAUNCH_CHECK("img_normalize_kernel");
  return ESB_OK;
}

extern "C" size_t esb_unproject_depth_workspace_bytes(int V, int H, int W) {
  long long n = (long long)V * H * W;
  if (n < 1) n = 1;
  return 2 * esb_align((size_t)n * 4) + esb_align(scan_bytes(n));
}

// depth (V,H,W) uint16 ; mats (V,16) fp32 device ; out (V*H*W,3) capacity ; count_dev device int32.
extern "C" int esb_unproject_depth(const unsigned short* depth, int V, int H, int W, float depth_shift,
                                   const float* mats, float* out, int* view_of, int* count_dev, void* ws,
                                   size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  long long n = (long long)V * H * W;
  ESB_CHECK_ARG(n < 2147483647LL, "esb_unproject_depth: too many pixels");
  if (ws_bytes < esb_unproject_depth_workspace_bytes(V, H, W)) {
    esb_set_error("esb_unproject_depth: workspace too small");
    return ESB_ENOMEM;
  }
  if (n == 0) { ESB_CUDA_CALL(cudaMemsetAsync(count_dev, 0, 4, stream)); return ESB_OK; }
  char* p = (char*)ws;
  int* flag = (int*)p; p += esb_align((size_t)n * 4);
  int* rank = (int*)p; p += esb_align((size_t)n * 4);
  size_t cb = scan_bytes(n);
  depth_flag_kernel<<<esb_div_up(n, 256), 256, 0, stream>>>(depth, n, flag);
  cub::DeviceScan::ExclusiveSum((void*)p, cb, flag, rank, (int)n, stream);
  unproject_kernel<<<esb_div_up(n, 256), 256, 0, stream>>>(depth, V, H, W, depth_shift, mats, flag, rank, out, view_of,
                                                            count_dev);
  ESB_CUDA_LAUNCH_CHECK("esb_unproject_depth");
  return ESB_OK;
}
