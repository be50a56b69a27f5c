// Depthwise 3x3 convolution for MobileNetV2 inference (models/mobilenet.py:38-76): HBM-bound element-wise work -
// 9 multiply-adds per output element, NHWC so that 8 channels travel in one 16-byte access; BatchNorm (running
// statistics) and ReLU6 are applied before the single bf16 store.
#include "common.h"
#include "ptx.cuh"
#include <cuda_bf16.h>

namespace sseg {

__global__ void __launch_bounds__(256) dwconv_affine_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            __nv_bfloat16* __restrict__ out, int N, int H, int W, int C,
                                                            int Ho, int Wo, int stride, int dil, int relu6) {
  pdl_sync();
  const int cg = C >> 3;
  const long total = (long)N * Ho * Wo * cg;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c0 = (idx % cg) << 3;
    long r = idx / cg;
    const int wo = r % Wo;
    r /= Wo;
    const int ho = r % Ho;
    const int n = r / Ho;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kr = 0; kr < 3; ++kr) {
      const int h = ho * stride + (kr - 1) * dil;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        const int ww = wo * stride + (ks - 1) * dil;
        if (ww < 0 || ww >= W) continue;
        const uint4 q = *reinterpret_cast<const uint4*>(x + (((long)n * H + h) * W + ww) * C + c0);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h2[e]);
          acc[2 * e] = fmaf(f.x, __ldg(w + (c0 + 2 * e) * 9 + kr * 3 + ks), acc[2 * e]);
          acc[2 * e + 1] = fmaf(f.y, __ldg(w + (c0 + 2 * e + 1) * 9 + kr * 3 + ks), acc[2 * e + 1]);
        }
      }
    }
    uint4 o;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float a = acc[2 * e], b = acc[2 * e + 1];
      if (scale != nullptr) {
        a = fmaf(a, scale[c0 + 2 * e], shift[c0 + 2 * e]);
        b = fmaf(b, scale[c0 + 2 * e + 1], shift[c0 + 2 * e + 1]);
      }
      if (relu6) a = fminf(fmaxf(a, 0.f), 6.f), b = fminf(fmaxf(b, 0.f), 6.f);
      o2[e] = __floats2bfloat162_rn(a, b);
    }
    *reinterpret_cast<uint4*>(out + (((long)n * Ho + ho) * Wo + wo) * C + c0) = o;
  }
}

// First layer of MobileNetV2 (3 -> Cout <= 64, 3x3, stride 2, pad 1; models/mobilenet.py:102) straight from the fp32 NCHW
// image, BatchNorm (running statistics) + ReLU6 folded in, bf16 NHWC out. One thread per output pixel.
__global__ void __launch_bounds__(128) stem_conv_affine_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                               const float* __restrict__ scale, const float* __restrict__ shift,
                                                               __nv_bfloat16* __restrict__ out, int N, int H, int W, int Ho,
                                                               int Wo, int Cout, int relu6) {
  pdl_sync();
  __shared__ float sw[27][64];
  __shared__ float ssc[64], ssh[64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) {
    const int co = i % 64, k = i / 64;
    sw[k][co] = co < Cout ? w[co * 27 + k] : 0.f;
  }
  if (threadIdx.x < 64) {
    ssc[threadIdx.x] = (scale != nullptr && threadIdx.x < Cout) ? scale[threadIdx.x] : 1.f;
    ssh[threadIdx.x] = (shift != nullptr && threadIdx.x < Cout) ? shift[threadIdx.x] : 0.f;
  }
  __syncthreads();
  const long P = (long)N * Ho * Wo;
  const long pidx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (pidx >= P) return;
  const int wo = pidx % Wo;
  const long r = pidx / Wo;
  const int ho = r % Ho;
  const int n = r / Ho;
  float acc[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) acc[c] = 0.f;
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int kr = 0; kr < 3; ++kr) {
      const int h = 2 * ho + kr - 1;
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        const int ww = 2 * wo + ks - 1;
        float x = 0.f;
        if (h >= 0 && h < H && ww >= 0 && ww < W) x = __ldg(img + (((long)n * 3 + ci) * H + h) * W + ww);
        const int k = ci * 9 + kr * 3 + ks;
#pragma unroll
        for (int c = 0; c < 64; ++c) acc[c] = fmaf(x, sw[k][c], acc[c]);
      }
    }
  __nv_bfloat16* op = out + pidx * Cout;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    if (g * 8 < Cout) {
      uint4 q;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float a = fmaf(acc[g * 8 + 2 * e], ssc[g * 8 + 2 * e], ssh[g * 8 + 2 * e]);
        float b = fmaf(acc[g * 8 + 2 * e + 1], ssc[g * 8 + 2 * e + 1], ssh[g * 8 + 2 * e + 1]);
        if (relu6) a = fminf(fmaxf(a, 0.f), 6.f), b = fminf(fmaxf(b, 0.f), 6.f);
        h2[e] = __floats2bfloat162_rn(a, b);
      }
      *reinterpret_cast<uint4*>(op + g * 8) = q;
    }
  }
}

}  // namespace sseg

using namespace sseg;

extern "C" int sseg_stem_conv_affine(const float* img, int N, int H, int W, const float* w, int cout, const float* scale,
                                     const float* shift, int relu6, void* out, sseg_stream_t st) {
  SSEG_REQUIRE(img && w && out && cout >= 8 && cout <= 64 && cout % 8 == 0, "sseg_stem_conv_affine: cout must be 8..64, x8");
  SSEG_REQUIRE((scale == nullptr) == (shift == nullptr), "sseg_stem_conv_affine: scale/shift must pair");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long P = (long)N * Ho * Wo;
  launch_k(stem_conv_affine_kernel, dim3((int)((P + 127) / 128)), dim3(128), 0, (cudaStream_t)st, img, w, scale, shift,
           (__nv_bfloat16*)out, N, H, W, Ho, Wo, cout, relu6);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "stem_conv_affine_kernel");
}

extern "C" int sseg_dwconv_affine(const void* x, int N, int H, int W, int C, const float* w, int stride, int dilation,
                                  const float* scale, const float* shift, int relu6, void* out, sseg_stream_t st) {
  SSEG_REQUIRE(x && w && out && C % 8 == 0 && N >= 1 && H >= 1 && W >= 1, "sseg_dwconv_affine: bad argument");
  SSEG_REQUIRE((stride == 1 || stride == 2) && dilation >= 1, "sseg_dwconv_affine: stride %d dilation %d", stride, dilation);
  SSEG_REQUIRE((scale == nullptr) == (shift == nullptr), "sseg_dwconv_affine: scale/shift must pair");
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;   // 'same' padding: pad = dilation
  long work = (long)N * Ho * Wo * (C / 8);
  long grid = (work + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  launch_k(dwconv_affine_kernel, dim3((int)grid), dim3(256), 0, (cudaStream_t)st, (const __nv_bfloat16*)x, w, scale, shift,
           (__nv_bfloat16*)out, N, H, W, C, Ho, Wo, stride, dilation, relu6);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "dwconv_affine_kernel");
}
