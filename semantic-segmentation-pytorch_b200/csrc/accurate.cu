// fp32-accurate inference on bf16 tensor cores: the element-wise half.
//
// north_star / BASELINE config 2 asks for forward logits within 1e-3 of the reference's fp32 path with a bit-exact arg-max
// map; one bf16 rounding per layer gives 6e-3 (measured with the oracle). In the accurate mode every activation is
// a PAIR of bf16 tensors (hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits), every weight likewise, and a convolution is
//      conv(x, w) ~= conv(x_hi, w_hi) + conv(x_lo, w_hi) + conv(x_hi, w_lo)          (error 1e-5 on the logits)
// which is ONE launch of the existing tcgen05 implicit-GEMM kernel: the three products are a virtual channel
// concatenation [x_hi | x_lo | x_hi] against the K-concatenated weight [w_hi | w_hi | w_lo], accumulated in fp32 in
// tensor memory and written as fp32 (sseg_conv_igemm, out_f32 = 1). This file holds what surrounds those launches:
// the fp32 -> (affine, shortcut, ReLU) -> pair split, pooling / resizing of pairs, the fp32 stem, the weight split.
#include "common.h"
#include "ptx.cuh"
#include <cuda_bf16.h>

namespace sseg {
namespace acc {

__device__ __forceinline__ void load8f(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x, f[2 * i + 1] = t.y;
  }
}
// pair value = hi + lo
__device__ __forceinline__ void load_pair(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float (&f)[8]) {
  float a[8], b[8];
  load8f(hi, a);
  load8f(lo, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = a[i] + b[i];
}
// hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void store_pair(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&f)[8]) {
  uint4 qh, ql;
  __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(&qh);
  __nv_bfloat16* l = reinterpret_cast<__nv_bfloat16*>(&ql);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = __float2bfloat16_rn(f[i]);
    l[i] = __float2bfloat16_rn(f[i] - __bfloat162float(h[i]));
  }
  *reinterpret_cast<uint4*>(hi) = qh;
  *reinterpret_cast<uint4*>(lo) = ql;
}
__device__ __forceinline__ void coeff(int dst, int in, int out, int& i0, int& i1, float& lam) {
  const float scale = (float)in / (float)out;
  float src = ((float)dst + 0.5f) * scale - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  i1 = i0 < in - 1 ? i0 + 1 : i0;
  lam = src - (float)i0;
}
static inline int grid_for(long work, int block, int max_blocks = 148 * 16) {
  long g = (work + block - 1) / block;
  return (int)(g > max_blocks ? max_blocks : (g < 1 ? 1 : g));
}

// out pair = relu?( z * scale + shift (+ res pair) )   (res_after_relu: relu first, then the add)
struct SplitAffineParams {
  const float* z;
  long z_ld, z_row, z_img;  // fp32 NHWC view (a strided view subsamples a stride-1 convolution output)
  const float *scale, *shift;
  const __nv_bfloat16 *res_hi, *res_lo;
  long res_ld;
  __nv_bfloat16 *out_hi, *out_lo;
  long out_ld;
  int N, H, W, C, relu, res_after_relu;
};
__global__ void __launch_bounds__(256) split_affine_kernel(const SplitAffineParams p) {
  pdl_sync();
  const int cg = p.C >> 3;
  const long total = (long)p.N * p.H * p.W * cg;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c0 = (idx % cg) << 3;
    long r = idx / cg;
    const long pix = r;
    const int w = r % p.W;
    r /= p.W;
    const int h = r % p.H;
    const int n = r / p.H;
    const float* zp = p.z + n * p.z_img + h * p.z_row + (long)w * p.z_ld + c0;
    const float4 a = *reinterpret_cast<const float4*>(zp), b = *reinterpret_cast<const float4*>(zp + 4);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (p.scale != nullptr) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], p.scale[c0 + e], p.shift[c0 + e]);
    }
    float rr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (p.res_hi != nullptr) load_pair(p.res_hi + pix * p.res_ld + c0, p.res_lo + pix * p.res_ld + c0, rr);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = v[e];
      if (!p.res_after_relu) t += rr[e];
      if (p.relu) t = fmaxf(t, 0.f);
      if (p.res_after_relu) t += rr[e];
      v[e] = t;
    }
    store_pair(p.out_hi + pix * p.out_ld + c0, p.out_lo + pix * p.out_ld + c0, v);
  }
}

// conv1 of the deep stem (3 -> 64, 3x3, stride 2, pad 1) in fp32 from the fp32 image, fp32 NHWC output
__global__ void __launch_bounds__(128) stem_conv_f32_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                            float* __restrict__ out, int N, int H, int W, int Ho, int Wo) {
  pdl_sync();
  __shared__ float sw[27][64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) sw[i / 64][i % 64] = w[(i % 64) * 27 + i / 64];
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
  float* op = out + pidx * 64;
#pragma unroll
  for (int g = 0; g < 16; ++g) *reinterpret_cast<float4*>(op + g * 4) = make_float4(acc[g * 4], acc[g * 4 + 1], acc[g * 4 + 2], acc[g * 4 + 3]);
}

// MaxPool2d(3, 2, 1) on pairs: the window's maximum of hi + lo, and that element's own pair
__global__ void maxpool_pair_kernel(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl,
                                    __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol, int N, int H, int W, int C,
                                    int Ho, int Wo) {
  pdl_sync();
  const int cg = C >> 3;
  const long total = (long)N * Ho * Wo * cg;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c0 = (i % cg) << 3;
    long r = i / cg;
    const int wo = r % Wo;
    r /= Wo;
    const int ho = r % Ho;
    const int n = r / Ho;
    float best[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) best[e] = -INFINITY;
    for (int kr = 0; kr < 3; ++kr) {
      const int h = 2 * ho + kr - 1;
      if (h < 0 || h >= H) continue;
      for (int ks = 0; ks < 3; ++ks) {
        const int w = 2 * wo + ks - 1;
        if (w < 0 || w >= W) continue;
        const long off = (((long)n * H + h) * W + w) * C + c0;
        float v[8];
        load_pair(xh + off, xl + off, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) best[e] = fmaxf(best[e], v[e]);
      }
    }
    const long o = (((long)n * Ho + ho) * Wo + wo) * C + c0;
    store_pair(oh + o, ol + o, best);  // hi + lo of a pair is exactly representable again as a pair
  }
}

// AdaptiveAvgPool2d(S) on pairs (ATen's overlapping bins), one thread per output element group
__global__ void avgpool_pair_kernel(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl, long x_ld,
                                    __nv_bfloat16* __restrict__ oh, __nv_bfloat16* __restrict__ ol, int N, int H, int W, int C,
                                    int S) {
  pdl_sync();
  const int cg = C >> 3;
  const long total = (long)N * S * S * cg;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c0 = (idx % cg) << 3;
    long r = idx / cg;
    const int j = r % S;
    r /= S;
    const int i = r % S;
    const int n = r / S;
    const int h0 = (i * H) / S, h1 = ((i + 1) * H + S - 1) / S;
    const int w0 = (j * W) / S, w1 = ((j + 1) * W + S - 1) / S;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int h = h0; h < h1; ++h)
      for (int w = w0; w < w1; ++w) {
        const long off = (((long)n * H + h) * W + w) * x_ld + c0;
        float v[8];
        load_pair(xh + off, xl + off, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += v[e];
      }
    const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= inv;
    const long o = (((long)n * S + i) * S + j) * C + c0;
    store_pair(oh + o, ol + o, acc);
  }
}

// bilinear resize (align_corners = False) on pairs
__global__ void bilinear_pair_kernel(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl, long x_ld,
                                     int N, int Hi, int Wi, int C, __nv_bfloat16* __restrict__ oh,
                                     __nv_bfloat16* __restrict__ ol, long out_ld, int Ho, int Wo) {
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
    int h0, h1, w0, w1;
    float lh, lw;
    coeff(ho, Hi, Ho, h0, h1, lh);
    coeff(wo, Wi, Wo, w0, w1, lw);
    float a[8], b[8], c[8], d[8], o[8];
    const long o00 = (((long)n * Hi + h0) * Wi + w0) * x_ld + c0, o01 = (((long)n * Hi + h0) * Wi + w1) * x_ld + c0;
    const long o10 = (((long)n * Hi + h1) * Wi + w0) * x_ld + c0, o11 = (((long)n * Hi + h1) * Wi + w1) * x_ld + c0;
    load_pair(xh + o00, xl + o00, a);
    load_pair(xh + o01, xl + o01, b);
    load_pair(xh + o10, xl + o10, c);
    load_pair(xh + o11, xl + o11, d);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      o[e] = (1.f - lh) * ((1.f - lw) * a[e] + lw * b[e]) + lh * ((1.f - lw) * c[e] + lw * d[e]);
    const long oo = (((long)n * Ho + ho) * Wo + wo) * out_ld + c0;
    store_pair(oh + oo, ol + oo, o);
  }
}

// fp32 OIHW weight -> bf16 [O][ld] with K = per tap [w_hi (I) | w_hi (I) | w_lo (I)], matching sources [x_hi | x_lo | x_hi]
__global__ void prep_weight_split_kernel(const float* __restrict__ w, int O, int I, int T, __nv_bfloat16* __restrict__ out,
                                         long ld) {
  pdl_sync();
  const long total = (long)O * I * T;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int t = idx % T;
    long r = idx / T;
    const int i = r % I;
    const int o = r / I;
    const float v = w[idx];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* row = out + (long)o * ld + (long)t * 3 * I;
    row[i] = hi, row[I + i] = hi, row[2 * I + i] = lo;
  }
}

}  // namespace acc
}  // namespace sseg

using namespace sseg;
using namespace sseg::acc;
#define ACC_LAUNCH_CHECK(name) \
  count_launch(1);             \
  return check_cuda(cudaGetLastError(), name)

extern "C" {

int sseg_split_affine(const sseg_act_t* z, const float* scale, const float* shift, const void* res_hi, const void* res_lo,
                      long res_ld, void* out_hi, void* out_lo, long out_ld, int relu, int res_after_relu, sseg_stream_t st) {
  SSEG_REQUIRE(z && z->ptr && out_hi && out_lo && z->c % 8 == 0 && z->ld % 4 == 0 && out_ld % 8 == 0,
               "sseg_split_affine: bad argument");
  SSEG_REQUIRE((scale == nullptr) == (shift == nullptr) && (res_hi == nullptr) == (res_lo == nullptr) &&
                   (!res_hi || res_ld % 8 == 0),
               "sseg_split_affine: scale/shift and res_hi/res_lo must pair");
  SSEG_REQUIRE((reinterpret_cast<uintptr_t>(z->ptr) & 15) == 0 && z->row_stride % 4 == 0 && z->img_stride % 4 == 0,
               "sseg_split_affine: z not 16B aligned");
  SplitAffineParams p;
  memset(&p, 0, sizeof(p));
  p.z = static_cast<const float*>(z->ptr), p.z_ld = z->ld, p.z_row = z->row_stride, p.z_img = z->img_stride;
  p.scale = scale, p.shift = shift;
  p.res_hi = static_cast<const __nv_bfloat16*>(res_hi), p.res_lo = static_cast<const __nv_bfloat16*>(res_lo), p.res_ld = res_ld;
  p.out_hi = static_cast<__nv_bfloat16*>(out_hi), p.out_lo = static_cast<__nv_bfloat16*>(out_lo), p.out_ld = out_ld;
  p.N = z->n, p.H = z->h, p.W = z->w, p.C = z->c, p.relu = relu, p.res_after_relu = res_after_relu;
  launch_k(split_affine_kernel, dim3(grid_for((long)p.N * p.H * p.W * (p.C / 8), 256)), dim3(256), 0, (cudaStream_t)st, p);
  ACC_LAUNCH_CHECK("split_affine_kernel");
}

int sseg_stem_conv_fwd_f32(const float* img, int N, int H, int W, const float* w, float* out, sseg_stream_t st) {
  SSEG_REQUIRE(img && w && out, "sseg_stem_conv_fwd_f32: null argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long P = (long)N * Ho * Wo;
  launch_k(stem_conv_f32_kernel, dim3((int)((P + 127) / 128)), dim3(128), 0, (cudaStream_t)st, img, w, out, N, H, W, Ho, Wo);
  ACC_LAUNCH_CHECK("stem_conv_f32_kernel");
}

int sseg_maxpool_pair_fwd(const void* x_hi, const void* x_lo, int N, int H, int W, int C, void* out_hi, void* out_lo,
                          sseg_stream_t st) {
  SSEG_REQUIRE(x_hi && x_lo && out_hi && out_lo && C % 8 == 0, "sseg_maxpool_pair_fwd: bad argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  launch_k(maxpool_pair_kernel, dim3(grid_for((long)N * Ho * Wo * (C / 8), 256)), dim3(256), 0, (cudaStream_t)st,
           (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, N, H, W, C,
           Ho, Wo);
  ACC_LAUNCH_CHECK("maxpool_pair_kernel");
}

int sseg_avgpool_pair_fwd(const void* x_hi, const void* x_lo, long x_ld, int N, int H, int W, int C, int S, void* out_hi,
                          void* out_lo, sseg_stream_t st) {
  SSEG_REQUIRE(x_hi && x_lo && out_hi && out_lo && C % 8 == 0 && x_ld % 8 == 0 && S >= 1, "sseg_avgpool_pair_fwd: bad argument");
  launch_k(avgpool_pair_kernel, dim3(grid_for((long)N * S * S * (C / 8), 128)), dim3(128), 0, (cudaStream_t)st,
           (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, x_ld, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, N, H, W,
           C, S);
  ACC_LAUNCH_CHECK("avgpool_pair_kernel");
}

int sseg_bilinear_pair_fwd(const void* x_hi, const void* x_lo, long x_ld, int N, int Hi, int Wi, int C, void* out_hi,
                           void* out_lo, long out_ld, int Ho, int Wo, sseg_stream_t st) {
  SSEG_REQUIRE(x_hi && x_lo && out_hi && out_lo && C % 8 == 0 && x_ld % 8 == 0 && out_ld % 8 == 0,
               "sseg_bilinear_pair_fwd: bad argument");
  launch_k(bilinear_pair_kernel, dim3(grid_for((long)N * Ho * Wo * (C / 8), 256)), dim3(256), 0, (cudaStream_t)st,
           (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, x_ld, N, Hi, Wi, C, (__nv_bfloat16*)out_hi,
           (__nv_bfloat16*)out_lo, out_ld, Ho, Wo);
  ACC_LAUNCH_CHECK("bilinear_pair_kernel");
}

int sseg_prep_conv_weight_split(const float* w_oihw, int O, int I, int T, void* out, long ld, sseg_stream_t st) {
  SSEG_REQUIRE(w_oihw && out && ld >= (long)3 * T * I && ld % 8 == 0, "sseg_prep_conv_weight_split: bad argument");
  launch_k(prep_weight_split_kernel, dim3(grid_for((long)O * I * T, 256)), dim3(256), 0, (cudaStream_t)st, w_oihw, O, I, T,
           (__nv_bfloat16*)out, ld);
  ACC_LAUNCH_CHECK("prep_weight_split_kernel");
}

}  // extern "C"
