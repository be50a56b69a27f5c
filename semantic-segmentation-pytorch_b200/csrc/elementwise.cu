// HBM-bound kernels of the path: batch-norm (finalize / apply / backward), max-pool, adaptive average pool,
// bilinear resize, the stem convolution (Cin = 3), weight re-layout, log-softmax + NLL loss, column sums.
// All activations are NHWC bf16 (C contiguous) accessed with 16-byte vectors (8 channels per thread).
#include "common.h"
#include "ptx.cuh"
#include "peer.cuh"
#include <cuda_bf16.h>

namespace sseg {

struct alignas(16) BF8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x, f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 ldq(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x, f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 q;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = q;
}

static inline int grid_for(long work, int block, int max_blocks = 148 * 16) {
  long g = (work + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------------
// weight re-layout: fp32 OIHW master weights -> bf16 [O][T*I] (forward) and bf16 [I][T*Opad] (data gradient)
__global__ void prep_weight_kernel(const float* __restrict__ w, int O, int I, int T, __nv_bfloat16* __restrict__ wf,
                                   long wf_ld, __nv_bfloat16* __restrict__ wd, long wd_ld, int o_pad) {
  pdl_sync();
  const long total = (long)O * I * T;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    // idx enumerates the forward layout (o, t, i) so the bf16 stores are coalesced
    const int i = idx % I;
    const long r = idx / I;
    const int t = r % T;
    const int o = r / T;
    const float v = w[((long)o * I + i) * T + t];
    const __nv_bfloat16 b = __float2bfloat16(v);
    if (wf) wf[(long)o * wf_ld + (long)t * I + i] = b;
    if (wd) wd[(long)i * wd_ld + (long)t * o_pad + o] = b;
  }
}

// gradient re-layout: fp32 [O][T*I] (what wgrad produces) -> fp32 OIHW (+= or =), times scale
__global__ void grad_to_oihw_kernel(const float* __restrict__ g, long g_ld, int O, int I, int T, float* __restrict__ out,
                                    float scale, int accumulate) {
  pdl_sync();
  const long total = (long)O * I * T;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    // idx enumerates OIHW so the stores are coalesced
    const int t = idx % T;
    const long r = idx / T;
    const int i = r % I;
    const int o = r / I;
    const float v = g[(long)o * g_ld + (long)t * I + i] * scale;
    out[idx] = accumulate ? out[idx] + v : v;
  }
}

// Flat index over [n][h][w][channel group] -> coordinates. 32-bit divisions whenever the index fits: a 64-bit integer
// division costs ~100 instructions on the GPU, and four of them per 32 bytes of traffic made these memory-bound kernels
// instruction-bound (maxpool backward: 40 us for a 50 MB pass).
__device__ __forceinline__ void split_idx(long idx, int cg, int W, int H, int& c0, int& w, int& h, int& n) {
  if (idx < (1L << 31)) {
    unsigned r = (unsigned)idx;
    unsigned q = r / (unsigned)cg;
    c0 = (int)(r - q * (unsigned)cg) << 3, r = q;
    q = r / (unsigned)W;
    w = (int)(r - q * (unsigned)W), r = q;
    q = r / (unsigned)H;
    h = (int)(r - q * (unsigned)H), n = (int)q;
  } else {
    long r = idx;
    c0 = (int)(r % cg) << 3, r /= cg;
    w = (int)(r % W), r /= W;
    h = (int)(r % H), n = (int)(r / H);
  }
}

// ------------------------------------------------------------------------------------------------
// stem conv: fp32 NCHW image [N,3,H,W] -> bf16 NHWC [N,Ho,Wo,64], 3x3 stride 2 pad 1 (models/resnet.py:100)
// + batch-norm statistics. (0.06 % of the step's FLOPs, on CUDA cores - but on the step's critical path at both ends.)
// One block per output row (n, ho), walking it in tiles of 64 pixels. The three input rows x three colour planes the tile
// needs are staged in shared memory with coalesced loads; thread = 4 pixels x 8 output channels (register tile: 6 shared
// loads per 32 FMAs; the first version's one-pixel-per-thread layout was bound by shared-memory loads and by 640 shuffles
// per thread for the statistics: 60 us, now ~15). Every output is one FMA chain over the 27 taps in the same order as before.
constexpr int kStemTW = 64;                 // output pixels per tile
constexpr int kStemInW = 2 * kStemTW + 4;   // staged input columns (2*64 + 1 used), padded
__global__ void __launch_bounds__(128) stem_conv_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                            __nv_bfloat16* __restrict__ out, float* __restrict__ ssum,
                                                            float* __restrict__ ssq, int N, int H, int W, int Ho,
                                                            int Wo) {
  pdl_sync();
  __shared__ __align__(16) float sw[27][64];
  __shared__ float s_in[9][kStemInW];
  __shared__ float bsum[64], bsq[64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) {
    const int co = i % 64, k = i / 64;  // k = ci*9 + r*3 + s, OIHW source index = co*27 + k
    sw[k][co] = w[co * 27 + k];
  }
  if (threadIdx.x < 64) bsum[threadIdx.x] = 0.f, bsq[threadIdx.x] = 0.f;
  const int ho = blockIdx.x % Ho, n = blockIdx.x / Ho;
  const int cog = threadIdx.x & 7, pxg = threadIdx.x >> 3;   // 8 channel groups x 16 pixel groups
  float st_s[8], st_q[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) st_s[c] = 0.f, st_q[c] = 0.f;
  for (int wo0 = 0; wo0 < Wo; wo0 += kStemTW) {
    __syncthreads();  // the previous tile's readers are done (first pass: sw / bsum are complete)
    for (int v = threadIdx.x; v < 9 * (2 * kStemTW + 1); v += blockDim.x) {
      const int row = v / (2 * kStemTW + 1), col = v - row * (2 * kStemTW + 1);
      const int ci = row / 3, kr = row - ci * 3;
      const int h = 2 * ho + kr - 1, ww = 2 * wo0 - 1 + col;
      float x = 0.f;
      if (h >= 0 && h < H && ww >= 0 && ww < W) x = __ldg(img + (((long)n * 3 + ci) * H + h) * W + ww);
      s_in[row][col] = x;
    }
    __syncthreads();
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
#pragma unroll
    for (int row = 0; row < 9; ++row) {
      float xv[9];  // input columns 2*px0 .. 2*px0 + 8 of this (ci, kr) row: the three taps of four neighbouring pixels
#pragma unroll
      for (int q = 0; q < 9; ++q) xv[q] = s_in[row][8 * pxg + q];
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        const float4 wa = *reinterpret_cast<const float4*>(&sw[row * 3 + ks][cog * 8]);
        const float4 wb = *reinterpret_cast<const float4*>(&sw[row * 3 + ks][cog * 8 + 4]);
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(xv[2 * i + ks], wv[c], acc[i][c]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int wo = wo0 + pxg * 4 + i;
      if (wo < Wo) {
        store8(out + (((long)n * Ho + ho) * Wo + wo) * 64 + cog * 8, acc[i]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float v = __bfloat162float(__float2bfloat16(acc[i][c]));  // statistics of the stored values
          st_s[c] += v, st_q[c] = fmaf(v, v, st_q[c]);
        }
      }
    }
  }
  if (ssum != nullptr) {
    // lanes of a warp: 8 channel groups x 4 pixel groups -> fold the pixel groups (lane bits 3, 4), then the 4 warps
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
      for (int off = 8; off <= 16; off <<= 1) {
        st_s[c] += __shfl_xor_sync(0xffffffffu, st_s[c], off);
        st_q[c] += __shfl_xor_sync(0xffffffffu, st_q[c], off);
      }
    }
    if ((threadIdx.x & 31) < 8) {
#pragma unroll
      for (int c = 0; c < 8; ++c) atomicAdd(&bsum[cog * 8 + c], st_s[c]), atomicAdd(&bsq[cog * 8 + c], st_q[c]);
    }
    __syncthreads();
    if (threadIdx.x < 64) atomicAdd(ssum + threadIdx.x, bsum[threadIdx.x]), atomicAdd(ssq + threadIdx.x, bsq[threadIdx.x]);
  }
}

// stem conv weight gradient: dW[co][ci][r][s] += sum_pixels dy[p, co] * x[n, ci, 2ho+r-1, 2wo+s-1]
// One block per output row, tiles of 64 pixels: the input rows are staged as in the forward kernel, the dy tile (64 px x
// 64 co, bf16) next to them. 256 threads = 4 pixel lanes x 16 channel groups (4 co) x 4 tap groups (8 of the 27 taps, padded
// to 32): a 4 x 8 register tile per thread, 9 shared loads per 32 FMAs (the first version: 8 per 7, and 64-bit index
// arithmetic per staged element: 110 us at the very end of the step). Pixel lanes are folded with shuffles at the end.
__global__ void __launch_bounds__(256, 3) stem_conv_wgrad_kernel(const float* __restrict__ img,
                                                              const __nv_bfloat16* __restrict__ dy,
                                                              float* __restrict__ dw, int N, int H, int W, int Ho, int Wo) {
  pdl_sync();
  __shared__ __align__(16) __nv_bfloat16 s_dy[kStemTW][96];   // rows padded to 192 B: the 4 pixel lanes hit disjoint banks
  __shared__ float s_in[9 * kStemInW + 8];
  const int ho = blockIdx.x % Ho, n = blockIdx.x / Ho;
  const int pxl = threadIdx.x & 3, cog = (threadIdx.x >> 2) & 15, tg = threadIdx.x >> 6;
  int off[8];   // shared-memory offset of tap k = tg*8 + j for pixel 0: row (ci, kr) * pitch + ks
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = tg * 8 + j;
    off[j] = k < 27 ? (k / 3) * kStemInW + (k % 3) : 0;
  }
  float acc[4][8];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
  for (int wo0 = 0; wo0 < Wo; wo0 += kStemTW) {
    const int npx = min(kStemTW, Wo - wo0);
    __syncthreads();
    for (int v = threadIdx.x; v < 9 * (2 * kStemTW + 1); v += blockDim.x) {
      const int row = v / (2 * kStemTW + 1), col = v - row * (2 * kStemTW + 1);
      const int ci = row / 3, kr = row - ci * 3;
      const int h = 2 * ho + kr - 1, ww = 2 * wo0 - 1 + col;
      float x = 0.f;
      if (h >= 0 && h < H && ww >= 0 && ww < W) x = __ldg(img + (((long)n * 3 + ci) * H + h) * W + ww);
      s_in[row * kStemInW + col] = x;
    }
    for (int v = threadIdx.x; v < kStemTW * 8; v += blockDim.x) {
      const int px = v >> 3, part = v & 7;
      uint4 q = make_uint4(0, 0, 0, 0);   // pixels beyond the row contribute zeros
      if (px < npx) q = *reinterpret_cast<const uint4*>(dy + (((long)n * Ho + ho) * Wo + wo0 + px) * 64 + part * 8);
      *reinterpret_cast<uint4*>(&s_dy[px][part * 8]) = q;
    }
    __syncthreads();
#pragma unroll 2
    for (int px = pxl; px < kStemTW; px += 4) {
      const uint2 gq = *reinterpret_cast<const uint2*>(&s_dy[px][cog * 4]);
      const float2 g01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gq.x));
      const float2 g23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gq.y));
      const float g[4] = {g01.x, g01.y, g23.x, g23.y};
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = s_in[off[j] + 2 * px];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[c][j] = fmaf(g[c], x[j], acc[c][j]);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[c][j];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      const int k = tg * 8 + j;
      if (pxl == 0 && k < 27) atomicAdd(dw + (cog * 4 + c) * 27 + k, v);
    }
}

// ------------------------------------------------------------------------------------------------
// batch norm
//   finalize: (sum, sqsum, count) -> mean, inv_std, scale = gamma*inv_std, shift = beta - mean*scale; running stats.
//   mode 0: F.batch_norm training  (inv_std = 1/sqrt(var+eps); running = (1-m)*running + m*{mean, unbiased var})
//           [lib/nn/modules/batchnorm.py:58-61 -> torch native batch_norm]
//   mode 1: SynchronizedBatchNorm parallel branch (inv_std = clamp(var,eps)^-0.5; accumulator-style running stats)
//           [lib/nn/modules/batchnorm.py:123-139]
//   mode 2: evaluation: statistics = running_mean / running_var (inv_std = 1/sqrt(var+eps))
__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sqsum,
                                   const float* __restrict__ count_dev, float count_host, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, int mode, int update_running,
                                   float* running_mean, float* running_var, float* tmp_mean, float* tmp_var,
                                   float* running_iter, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                   float* __restrict__ scale, float* __restrict__ shift, int C) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, inv_std;
  if (mode == 2) {
    mean = running_mean[c];
    inv_std = rsqrtf(running_var[c] + eps);
  } else {
    const float cnt = count_dev ? *count_dev : count_host;
    mean = sum[c] / cnt;
    const float sumvar = sqsum[c] - sum[c] * mean;
    const float bias_var = sumvar / cnt;
    const float unbias_var = sumvar / (cnt - 1.f);
    if (mode == 0) {
      inv_std = rsqrtf(fmaxf(bias_var, 0.f) + eps);
      if (update_running) {
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbias_var;
      }
    } else {
      inv_std = rsqrtf(fmaxf(bias_var, eps));
      if (update_running) {
        const float frac = 1.f - momentum;
        const float it = running_iter[0] * frac + 1.f;  // every thread computes the same value; thread 0 stores it
        const float tm = tmp_mean[c] * frac + mean;
        const float tv = tmp_var[c] * frac + unbias_var;
        tmp_mean[c] = tm, tmp_var[c] = tv;
        running_mean[c] = tm / it, running_var[c] = tv / it;
      }
    }
  }
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  mean_out[c] = mean, invstd_out[c] = inv_std;
  scale[c] = g * inv_std;
  shift[c] = b - mean * g * inv_std;
}
__global__ void bn_iter_update_kernel(float* running_iter, float momentum) {
  pdl_sync();
  running_iter[0] = running_iter[0] * (1.f - momentum) + 1.f;
}

// The three streaming BN kernels share one decomposition: a block of 256 threads = CGB channel groups (8 channels
// each) x ROWS pixel lanes; a thread keeps its per-channel coefficients in registers and walks pixels
// (pix = first + k * stride), UNROLL pixels per iteration with all loads issued before any use.
constexpr int kBnUnroll = 4;

struct BnTiling {
  int cgb, rows;    // channel groups per block, pixel lanes per block
  int gx, gy;       // grid
};
static inline BnTiling bn_tiling(long P, int C, bool reduce = false) {
  BnTiling t;
  const int cg = C >> 3;
  t.cgb = cg < 32 ? cg : 32;  // <= 32 groups (256 channels) per block keeps >= 8 pixel lanes
  while (256 % t.cgb != 0) --t.cgb;
  t.rows = 256 / t.cgb;
  t.gy = (cg + t.cgb - 1) / t.cgb;
  long want = (P + (long)t.rows * kBnUnroll * 2 - 1) / ((long)t.rows * kBnUnroll * 2);  // ~8 pixels per thread
  // the reduce kernel pays a block-level tail (smem reduction + atomics): keep its grid at one wave of 2 blocks/SM
  long cap = ((reduce ? 148L * 2 : 148L * 8) + t.gy - 1) / t.gy;
  t.gx = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  return t;
}

// apply: out = [relu]( y*scale + shift + residual' ) * chanmul[n][c];  residual' = r (* rscale + rshift)
struct BnApplyParams {
  const __nv_bfloat16* y;
  long y_ld;
  const float *scale, *shift;
  const __nv_bfloat16* res;
  long res_ld;
  const float *rscale, *rshift;
  const float* chanmul;  // [N][C] or null (Dropout2d keep-mask / (1-p))
  __nv_bfloat16* out;
  long out_ld;
  long P;          // pixels
  long pix_per_img;
  int C, relu, cgb, rows;
  int res_after_relu;  // out = relu(y*scale+shift) + res  (FPN lateral + top-down add) instead of relu(... + res)
  // fused finalize (F.batch_norm training branch, SSEG_BN_TRAIN): when fin_sum != null every thread derives scale / shift
  // for its 8 channels from the statistics instead of reading them, and block column 0 publishes mean / invstd / scale /
  // shift (for the backward pass) and updates the running statistics - one kernel boundary less per layer.
  const float *fin_sum, *fin_sqsum, *fin_gamma, *fin_beta;
  float fin_count, fin_eps, fin_momentum;
  float *fin_mean, *fin_invstd, *fin_scale, *fin_shift, *fin_running_mean, *fin_running_var;
};
__global__ void __launch_bounds__(256, 2) bn_apply_kernel(const BnApplyParams p) {
  pdl_sync();
  const int tcol = threadIdx.x % p.cgb, trow = threadIdx.x / p.cgb;
  const int cgrp = blockIdx.y * p.cgb + tcol;
  if (cgrp >= (p.C >> 3)) return;
  const int c0 = cgrp << 3;
  float sc[8], sh[8], rs[8], rb[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (p.fin_sum != nullptr) {
      const int c = c0 + e;
      const float s = p.fin_sum[c], mean = s / p.fin_count;
      const float sumvar = p.fin_sqsum[c] - s * mean;
      const float inv_std = rsqrtf(fmaxf(sumvar / p.fin_count, 0.f) + p.fin_eps);
      const float g = p.fin_gamma ? p.fin_gamma[c] : 1.f, b = p.fin_beta ? p.fin_beta[c] : 0.f;
      sc[e] = g * inv_std, sh[e] = b - mean * g * inv_std;
      if (blockIdx.x == 0 && trow == 0) {
        p.fin_mean[c] = mean, p.fin_invstd[c] = inv_std, p.fin_scale[c] = sc[e], p.fin_shift[c] = sh[e];
        if (p.fin_running_mean != nullptr) {
          p.fin_running_mean[c] = (1.f - p.fin_momentum) * p.fin_running_mean[c] + p.fin_momentum * mean;
          p.fin_running_var[c] = (1.f - p.fin_momentum) * p.fin_running_var[c] + p.fin_momentum * sumvar / (p.fin_count - 1.f);
        }
      }
    } else {
      sc[e] = p.scale[c0 + e], sh[e] = p.shift[c0 + e];
    }
    rs[e] = p.rscale ? p.rscale[c0 + e] : 1.f, rb[e] = p.rshift ? p.rshift[c0 + e] : 0.f;
  }
  const long stride = (long)gridDim.x * p.rows;
  for (long pix0 = (long)blockIdx.x * p.rows + trow; pix0 < p.P; pix0 += stride * kBnUnroll) {
    uint4 qv[kBnUnroll], qr[kBnUnroll];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const long pix = pix0 + u * stride;
      if (pix < p.P) {
        qv[u] = ldq(p.y + pix * p.y_ld + c0);
        if (p.res) qr[u] = ldq(p.res + pix * p.res_ld + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const long pix = pix0 + u * stride;
      if (pix >= p.P) continue;
      float v[8], r[8];
      unpack8(qv[u], v);
      if (p.res) unpack8(qr[u], r);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float t = fmaf(v[e], sc[e], sh[e]);
        const float rr = p.res ? fmaf(r[e], rs[e], rb[e]) : 0.f;
        if (!p.res_after_relu) t += rr;
        if (p.relu) t = fmaxf(t, 0.f);
        if (p.res_after_relu) t += rr;
        v[e] = t;
      }
      if (p.chanmul) {
        const float* m = p.chanmul + (pix / p.pix_per_img) * p.C + c0;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= m[e];
      }
      store8(p.out + pix * p.out_ld + c0, v);
    }
  }
}

// backward.  g' = g * chanmul * [relu active]; relu active = (a > 0) from the saved output, or, for layers without a
// shortcut, recomputed from y as (y*fscale + fshift > 0) so the saved output need not be read at all.
//   pass 1 (reduce): s1 += sum g', s2 += sum g' * xhat            (xhat = (y - mean) * invstd)
//   pass 2 (apply) : dy = scale * (g' - s1/M - xhat * s2/M)   (eval_mode: dy = scale * g');  dres (optional) = g'
struct BnBwdParams {
  const __nv_bfloat16* g;
  long g_ld;
  const __nv_bfloat16* a;  // saved block output (post-ReLU) or null
  long a_ld;
  const __nv_bfloat16* y;  // saved conv output (pre-BN)
  long y_ld;
  const float *mean, *invstd, *scale;  // scale = gamma * invstd (forward scale)
  const float* fshift;                 // forward shift; non-null (with a == null) => ReLU mask recomputed from y
  const float* chanmul;
  float* s1;
  float* s2;
  const float* count_dev;
  float count_host;
  __nv_bfloat16* dy;
  long dy_ld;
  __nv_bfloat16* dres;
  long dres_ld;
  long P, pix_per_img;
  int C;
  int eval_mode;  // statistics were constants (running stats): dy = scale * g'
  int cgb, rows;
  // s2 holds RAW sums (sum g'*y, from the fused dgrad epilogue) instead of sum g'*xhat: the kernel converts
  // s2 = invstd * (s2_raw - mean * s1) on the fly and block column 0 stores the converted value to dgamma_out
  int s2_raw;
  float* dgamma_out;
  // SyncBN across GPUs folded into the apply pass (sseg_bn_bwd_apply_peer): s1 / s2 are then THIS rank's partial sums at
  // part_off / part_off + C inside its peer arena; the kernel runs the flag handshake of csrc/peer.cu itself, pools the
  // partials of all ranks straight out of peer memory (one NVLink round trip) and block column 0 stores dbeta / dgamma
  // divided by world (the gradient bucket all-reduce sums them again)
  PeerTable pt;   // pt.world <= 1: single-GPU behaviour
  long part_off, flag_off;
  const int* step;
  float* dbeta_out;
};

template <bool kApply>
__global__ void __launch_bounds__(256, 2) bn_bwd_kernel(const BnBwdParams p) {
  pdl_sync();
  __shared__ float red[kApply ? 1 : 2][kApply ? 1 : 256][8];
  const int tcol = threadIdx.x % p.cgb, trow = threadIdx.x / p.cgb;
  const int cgrp = blockIdx.y * p.cgb + tcol;
  const bool active = cgrp < (p.C >> 3);
  const int c0 = (active ? cgrp : 0) << 3;
  const bool need_y = !p.eval_mode || (p.a == nullptr && p.fshift != nullptr);
  const bool mask_y = p.a == nullptr && p.fshift != nullptr;
  // per-channel constants:  reduce: mu, inv ;  apply: dy = ka*g' + kb*y + kc
  float mu[8], inv[8], fs[8], fb[8], ka[8], kb[8], kc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mu[e] = p.mean ? p.mean[c0 + e] : 0.f, inv[e] = p.invstd ? p.invstd[c0 + e] : 1.f;
    fs[e] = p.scale ? p.scale[c0 + e] : 1.f, fb[e] = p.fshift ? p.fshift[c0 + e] : 0.f;
  }
  if (kApply) {
    float s1p[8], s2p[8];
    if (p.pt.world > 1) {
      // pooled sums of this block's channels: thread (tcol, trow < 8) fetches channel 8 * group + trow from every rank (all
      // loads of the thread in flight together), the block shares them through shared memory
      // (dynamic shared memory, 2 KB, passed by sseg_bn_bwd_apply_peer only: the CPU simulator models static __shared__
      // arrays as one static object, which two "GPUs" of one test process would share)
      SSEG_DYN_SMEM(pooled_raw);
      float (*pooled)[256] = reinterpret_cast<float (*)[256]>(pooled_raw);
      peer_handshake(p.pt, p.flag_off, *p.step, blockIdx.x == 0 && blockIdx.y == 0);
      if (trow < 8 && active) {
        float a, b;
        peer_sum2(p.pt, p.part_off + c0 + trow, p.part_off + p.C + c0 + trow, &a, &b);
        pooled[0][tcol * 8 + trow] = a, pooled[1][tcol * 8 + trow] = b;
      }
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 8; ++e) s1p[e] = pooled[0][tcol * 8 + e], s2p[e] = pooled[1][tcol * 8 + e];
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s1p[e] = (p.s1 != nullptr && active) ? p.s1[c0 + e] : 0.f;
        s2p[e] = (p.s2 != nullptr && active) ? p.s2[c0 + e] : 0.f;
      }
    }
    const float inv_m = 1.f / (p.count_dev ? *p.count_dev : p.count_host);
    const float inv_w = p.pt.world > 1 ? 1.f / (float)p.pt.world : 1.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float s2v = s2p[e];
      if (p.s2_raw) s2v = inv[e] * (s2v - mu[e] * s1p[e]);
      if (active && blockIdx.x == 0 && trow == 0) {
        if ((p.s2_raw || p.pt.world > 1) && p.dgamma_out != nullptr) p.dgamma_out[c0 + e] = s2v * inv_w;
        if (p.pt.world > 1 && p.dbeta_out != nullptr) p.dbeta_out[c0 + e] = s1p[e] * inv_w;
      }
      if (p.eval_mode) {
        ka[e] = fs[e], kb[e] = 0.f, kc[e] = 0.f;
      } else {
        const float t = fs[e] * inv[e] * s2v * inv_m;
        ka[e] = fs[e], kb[e] = -t, kc[e] = t * mu[e] - fs[e] * s1p[e] * inv_m;
      }
    }
  }
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s1[e] = 0.f, s2[e] = 0.f;
  if (active) {
    const long stride = (long)gridDim.x * p.rows;
    for (long pix0 = (long)blockIdx.x * p.rows + trow; pix0 < p.P; pix0 += stride * kBnUnroll) {
      uint4 qg[kBnUnroll], qy[kBnUnroll], qa[kBnUnroll];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const long pix = pix0 + u * stride;
        if (pix < p.P) {
          qg[u] = ldq(p.g + pix * p.g_ld + c0);
          if (need_y) qy[u] = ldq(p.y + pix * p.y_ld + c0);
          if (p.a) qa[u] = ldq(p.a + pix * p.a_ld + c0);
        }
      }
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const long pix = pix0 + u * stride;
        if (pix >= p.P) continue;
        float g[8], y[8];
        unpack8(qg[u], g);
        if (need_y) unpack8(qy[u], y);
        if (p.chanmul) {
          const float* m = p.chanmul + (pix / p.pix_per_img) * p.C + c0;
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] *= m[e];
        }
        if (p.a) {
          float a[8];
          unpack8(qa[u], a);
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] = a[e] > 0.f ? g[e] : 0.f;
        } else if (mask_y) {
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] = fmaf(y[e], fs[e], fb[e]) > 0.f ? g[e] : 0.f;
        }
        if (kApply) {
          if (p.dres) store8(p.dres + pix * p.dres_ld + c0, g);
          float d[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) d[e] = p.eval_mode ? ka[e] * g[e] : fmaf(ka[e], g[e], fmaf(kb[e], y[e], kc[e]));
          store8(p.dy + pix * p.dy_ld + c0, d);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            s1[e] += g[e];
            s2[e] = fmaf(g[e], (y[e] - mu[e]) * inv[e], s2[e]);
          }
        }
      }
    }
  }
  if (!kApply) {
#pragma unroll
    for (int e = 0; e < 8; ++e) red[0][threadIdx.x][e] = s1[e], red[1][threadIdx.x][e] = s2[e];
    __syncthreads();
    if (trow == 0 && active) {
      for (int r = 1; r < p.rows; ++r) {
#pragma unroll
        for (int e = 0; e < 8; ++e) s1[e] += red[0][r * p.cgb + tcol][e], s2[e] += red[1][r * p.cgb + tcol][e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(p.s1 + c0 + e, s1[e]), atomicAdd(p.s2 + c0 + e, s2[e]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// max pool 3x3 stride 2 pad 1 (models/resnet.py:109); the argmax tap (0..8, first maximum in scan order like torch)
// is kept in one byte per output element for the backward gather.
__global__ void maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                   uint8_t* __restrict__ idx, int N, int H, int W, int C, int Ho, int Wo) {
  pdl_sync();
  const int cg = C >> 3;
  const long total = (long)N * Ho * Wo * cg;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c0, wo, ho, n;
    split_idx(i, cg, Wo, Ho, c0, wo, ho, n);
    float best[8];
    int bidx[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) best[e] = -INFINITY, bidx[e] = 0;
#pragma unroll
    for (int kr = 0; kr < 3; ++kr) {
      const int h = 2 * ho + kr - 1;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        const int w = 2 * wo + ks - 1;
        if (w < 0 || w >= W) continue;
        float v[8];
        load8(x + (((long)n * H + h) * W + w) * C + c0, v);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (v[e] > best[e]) best[e] = v[e], bidx[e] = kr * 3 + ks;
      }
    }
    const long o = (((long)n * Ho + ho) * Wo + wo) * C + c0;
    store8(out + o, best);
    if (idx) {
      uint2 packed;
      packed.x = bidx[0] | (bidx[1] << 8) | (bidx[2] << 16) | (bidx[3] << 24);
      packed.y = bidx[4] | (bidx[5] << 8) | (bidx[6] << 16) | (bidx[7] << 24);
      *reinterpret_cast<uint2*>(idx + o) = packed;
    }
  }
}

__global__ void maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const uint8_t* __restrict__ idx,
                                   __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int Ho, int Wo) {
  pdl_sync();
  const int cg = C >> 3;
  const long total = (long)N * H * W * cg;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c0, w, h, n;
    split_idx(i, cg, W, H, c0, w, h, n);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    // output windows containing (h, w): ho with 2ho-1 <= h <= 2ho+1
    for (int ho = (h) / 2; ho <= (h + 1) / 2; ++ho) {
      if (ho < 0 || ho >= Ho) continue;
      const int kr = h - 2 * ho + 1;
      for (int wo = (w) / 2; wo <= (w + 1) / 2; ++wo) {
        if (wo < 0 || wo >= Wo) continue;
        const int ks = w - 2 * wo + 1;
        const int tap = kr * 3 + ks;
        const long o = (((long)n * Ho + ho) * Wo + wo) * C + c0;
        const uint2 packed = *reinterpret_cast<const uint2*>(idx + o);
        float g[8];
        load8(dout + o, g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int b = ((e < 4 ? packed.x : packed.y) >> ((e & 3) * 8)) & 0xff;
          if (b == tap) acc[e] += g[e];
        }
      }
    }
    store8(dx + i * 8, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// adaptive average pool to s x s (models/models.py:447): bin i covers [floor(i*H/s), ceil((i+1)*H/s)).
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, long x_ld,
                                                          __nv_bfloat16* __restrict__ out, int N, int H, int W, int C,
                                                          int S) {
  pdl_sync();
  // grid: (N*S*S bins, ceil(C/64)); block = 8 channel groups (64 channels) x 32 pixel lanes striding over the bin
  __shared__ float red[32][8][8];
  const int tcg = threadIdx.x & 7, lane = threadIdx.x >> 3;
  const int c0 = (blockIdx.y * 8 + tcg) << 3;
  int b = blockIdx.x;
  const int j = b % S;
  b /= S;
  const int i = b % S;
  const int n = b / S;
  const int h0 = (i * H) / S, h1 = ((i + 1) * H + S - 1) / S;
  const int w0 = (j * W) / S, w1 = ((j + 1) * W + S - 1) / S;
  const int bw = w1 - w0, npix = (h1 - h0) * bw;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (c0 < C) {
    for (int p = lane; p < npix; p += 32) {
      const int h = h0 + p / bw, w = w0 + p % bw;
      float v[8];
      load8(x + (((long)n * H + h) * W + w) * x_ld + c0, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[lane][tcg][e] = acc[e];
  __syncthreads();
  if (lane == 0 && c0 < C) {
    for (int l = 1; l < 32; ++l) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += red[l][tcg][e];
    }
    const float inv = 1.f / (float)npix;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= inv;
    store8(out + (((long)n * S + i) * S + j) * C + c0, acc);
  }
}

// backward of up to 4 pooling scales at once, accumulated onto a base gradient:
//   dx[n,h,w,c] = base[n,h,w,c] + sum_k sum_{bins of scale k containing (h,w)} dpool_k[n,i,j,c] / binsize
struct AvgPoolBwdParams {
  const __nv_bfloat16* base;
  long base_ld;
  const __nv_bfloat16* dpool[4];
  int scales[4];
  int nscales;
  __nv_bfloat16* dx;
  long dx_ld;
  int N, H, W, C;
};
// One block = one image row segment (n, h, kAvgWT consecutive w): the bins of every scale that contain row h are found once
// per block, and all index arithmetic is 32-bit (the first version spent its time in 64-bit divisions: 143 us for a 67 MB
// pass). Threads: channel groups fastest (16-byte vectors, coalesced), the remaining lanes walk the pixels of the segment.
constexpr int kAvgWT = 8;
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const AvgPoolBwdParams p) {
  pdl_sync();
  const int cg = p.C >> 3;
  const int wtiles = (p.W + kAvgWT - 1) / kAvgWT;
  int b = blockIdx.x;
  const int w_base = (b % wtiles) * kAvgWT;
  b /= wtiles;
  const int h = b % p.H, n = b / p.H;
  // rows of bins containing h: candidates ic-1 .. ic+1 around ic = floor(h*S/H) (ATen's adaptive bins
  // [floor(i*H/S), ceil((i+1)*H/S)) overlap their neighbours by at most one row)
  __shared__ int bi[4][3], bh[4][3];   // bin row index, bin height (0 = not a bin of this row)
  if (threadIdx.x >= 224 && threadIdx.x < 236) {
    const int k = (threadIdx.x - 224) / 3, d = (threadIdx.x - 224) % 3;
    int i = 0, hh = 0;
    if (k < p.nscales) {
      const int S = p.scales[k];
      i = (h * S) / p.H - 1 + d;
      if (i >= 0 && i < S) {
        const int h0 = (i * p.H) / S, h1 = ((i + 1) * p.H + S - 1) / S;
        if (h >= h0 && h < h1) hh = h1 - h0;
      }
    }
    bi[k][d] = i, bh[k][d] = hh;
  }
  // ... and the columns of bins containing each of the segment's pixels, once per block (table in shared memory: the
  // per-pixel loop below is then free of integer divisions - with them it still ran at 90 us for a 67 MB pass)
  __shared__ int s_j[kAvgWT][4][3];    // bin column index
  __shared__ int s_bw[kAvgWT][4][3];   // bin width, 0 = not a bin of this column
  if (threadIdx.x < kAvgWT * 12) {
    const int wl = threadIdx.x / 12, k = (threadIdx.x % 12) / 3, dj = threadIdx.x % 3;
    const int w = w_base + wl;
    int j = 0, bwid = 0;
    if (k < p.nscales && w < p.W) {
      const int S = p.scales[k];
      j = (w * S) / p.W - 1 + dj;
      if (j >= 0 && j < S) {
        const int w0 = (j * p.W) / S, w1 = ((j + 1) * p.W + S - 1) / S;
        if (w >= w0 && w < w1) bwid = w1 - w0;
      }
    }
    s_j[wl][k][dj] = j, s_bw[wl][k][dj] = bwid;
  }
  __syncthreads();
  int cgb = cg < 256 ? cg : 256;
  while (256 % cgb != 0) --cgb;
  const int tcol = threadIdx.x % cgb, trow = threadIdx.x / cgb, lanes_w = 256 / cgb;
  for (int cgi = tcol; cgi < cg; cgi += cgb) {
    const int c0 = cgi << 3;
    for (int wl = trow; wl < kAvgWT; wl += lanes_w) {
      const int w = w_base + wl;
      if (w >= p.W) break;
      const long pix = ((long)n * p.H + h) * p.W + w;
      float acc[8];
      if (p.base)
        load8(p.base + pix * p.base_ld + c0, acc);
      else {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k >= p.nscales) break;
        const int S = p.scales[k];
#pragma unroll
        for (int dj = 0; dj < 3; ++dj) {
          const int bwid = s_bw[wl][k][dj];
          if (bwid == 0) continue;
          const int j = s_j[wl][k][dj];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            if (bh[k][d] == 0) continue;
            float g[8];
            load8(p.dpool[k] + (((long)n * S + bi[k][d]) * S + j) * p.C + c0, g);
            const float inv = 1.f / (bh[k][d] * bwid);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += g[e] * inv;
          }
        }
      }
      store8(p.dx + pix * p.dx_ld + c0, acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// bilinear resize, align_corners=False (models/models.py:472-475): src = max((dst+0.5)*in/out - 0.5, 0)
__device__ __forceinline__ void bilinear_coeff(int dst, int in, int out, int& i0, int& i1, float& lam) {
  const float scale = (float)in / (float)out;
  float src = ((float)dst + 0.5f) * scale - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  i1 = i0 < in - 1 ? i0 + 1 : i0;
  lam = src - (float)i0;
}

__global__ void bilinear_fwd_kernel(const __nv_bfloat16* __restrict__ x, long x_ld, int N, int Hi, int Wi, int C,
                                    __nv_bfloat16* __restrict__ out, long out_ld, int Ho, int Wo) {
  pdl_sync();
  const int cg = C >> 3;
  const long total = (long)N * Ho * Wo * cg;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int c0, wo, ho, n;
    split_idx(idx, cg, Wo, Ho, c0, wo, ho, n);
    int h0, h1, w0, w1;
    float lh, lw;
    bilinear_coeff(ho, Hi, Ho, h0, h1, lh);
    bilinear_coeff(wo, Wi, Wo, w0, w1, lw);
    float a[8], b[8], c[8], d[8], o[8];
    load8(x + (((long)n * Hi + h0) * Wi + w0) * x_ld + c0, a);
    load8(x + (((long)n * Hi + h0) * Wi + w1) * x_ld + c0, b);
    load8(x + (((long)n * Hi + h1) * Wi + w0) * x_ld + c0, c);
    load8(x + (((long)n * Hi + h1) * Wi + w1) * x_ld + c0, d);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      o[e] = (1.f - lh) * ((1.f - lw) * a[e] + lw * b[e]) + lh * ((1.f - lw) * c[e] + lw * d[e]);
    store8(out + (((long)n * Ho + ho) * Wo + wo) * out_ld + c0, o);
  }
}

// ------------------------------------------------------------------------------------------------
// HRNet exchange unit (models/hrnet.py:225-250): out = ReLU( sum_k affine_k( resample_k( x_k ) ) ) in ONE pass.
//   term k = a branch tensor at the output resolution (identity),
//          | scale*y + shift of a conv output y (the term's batch norm) at the output resolution (stride-2 chains),
//          | the same affine of a LOWER-resolution conv output, bilinearly sampled (align_corners=False) on the fly:
//            interpolation weights sum to one, so BN-then-upsample == upsample-then-BN and the up-sampled tensor
//            never exists in memory.
struct SumTermsParams {
  const __nv_bfloat16* x[SSEG_MAX_SUM_TERMS];
  const float* scale[SSEG_MAX_SUM_TERMS];
  const float* shift[SSEG_MAX_SUM_TERMS];
  int h[SSEG_MAX_SUM_TERMS], w[SSEG_MAX_SUM_TERMS];
  long ld[SSEG_MAX_SUM_TERMS];
  int nterms, N, Ho, Wo, C, relu;
  __nv_bfloat16* out;
  long out_ld;
};

__global__ void __launch_bounds__(256) sum_terms_kernel(const SumTermsParams p) {
  pdl_sync();
  const int cg = p.C >> 3;
  const long total = (long)p.N * p.Ho * p.Wo * cg;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int c0, wo, ho, n;
    split_idx(idx, cg, p.Wo, p.Ho, c0, wo, ho, n);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < SSEG_MAX_SUM_TERMS; ++k) {
      if (k >= p.nterms) break;
      const __nv_bfloat16* x = p.x[k];
      const int Hi = p.h[k], Wi = p.w[k];
      const long ld = p.ld[k];
      float v[8];
      if (Hi == p.Ho && Wi == p.Wo) {
        load8(x + (((long)n * Hi + ho) * Wi + wo) * ld + c0, v);
      } else {
        int h0, h1, w0, w1;
        float lh, lw;
        bilinear_coeff(ho, Hi, p.Ho, h0, h1, lh);
        bilinear_coeff(wo, Wi, p.Wo, w0, w1, lw);
        float a[8], b[8], c[8], d[8];
        load8(x + (((long)n * Hi + h0) * Wi + w0) * ld + c0, a);
        load8(x + (((long)n * Hi + h0) * Wi + w1) * ld + c0, b);
        load8(x + (((long)n * Hi + h1) * Wi + w0) * ld + c0, c);
        load8(x + (((long)n * Hi + h1) * Wi + w1) * ld + c0, d);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          v[e] = (1.f - lh) * ((1.f - lw) * a[e] + lw * b[e]) + lh * ((1.f - lw) * c[e] + lw * d[e]);
      }
      if (p.scale[k] != nullptr) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.scale[k] + c0));
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.scale[k] + c0 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.shift[k] + c0));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.shift[k] + c0 + 4));
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float sh[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += fmaf(v[e], sc[e], sh[e]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += v[e];
      }
    }
    if (p.relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaxf(acc[e], 0.f);
    }
    store8(p.out + (((long)n * p.Ho + ho) * p.Wo + wo) * p.out_ld + c0, acc);
  }
}

// backward of the ReLU of an exchange output: ds = g * [out > 0] (the gradient every term of the sum receives), and the
// identity term's share in the same pass: acc_out (+)= ds.
__global__ void __launch_bounds__(256) relu_mask_bwd_kernel(const __nv_bfloat16* __restrict__ g, long g_ld,
                                                            const __nv_bfloat16* __restrict__ out, long out_ld,
                                                            __nv_bfloat16* __restrict__ ds, long ds_ld,
                                                            __nv_bfloat16* __restrict__ acc_out, long acc_ld, int accumulate,
                                                            long P, int C) {
  pdl_sync();
  const int cg = C >> 3;
  const long total = P * cg;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c0 = (idx % cg) << 3;
    const long pix = idx / cg;
    float gv[8], ov[8];
    load8(g + pix * g_ld + c0, gv);
    load8(out + pix * out_ld + c0, ov);
#pragma unroll
    for (int e = 0; e < 8; ++e) gv[e] = ov[e] > 0.f ? gv[e] : 0.f;
    store8(ds + pix * ds_ld + c0, gv);
    if (acc_out != nullptr) {
      if (accumulate) {
        float av[8];
        load8(acc_out + pix * acc_ld + c0, av);
#pragma unroll
        for (int e = 0; e < 8; ++e) av[e] += gv[e];
        store8(acc_out + pix * acc_ld + c0, av);
      } else {
        store8(acc_out + pix * acc_ld + c0, gv);
      }
    }
  }
}

// backward = adjoint of the separable interpolation, in two gather passes (no atomics):
//   pass 1 (along W): tmp[n,ho,wi,c] = sum_wo  ww(wo -> wi) * dout[n,ho,wo,c]        (fp32 scratch [N,Ho,Wi,C])
//   pass 2 (along H): dx [n,hi,wi,c] (+)= sum_ho wh(ho -> hi) * tmp[n,ho,wi,c]
// Each thread owns 8 channels of one destination element and walks only the source positions that reference it.
__device__ __forceinline__ void bilinear_src_window(int i, int in, int out, int& lo, int& hi) {
  const float s = (float)out / (float)in;
  lo = max(0, (int)floorf(((float)i - 1.f + 0.5f) * s - 0.5f) - 1);
  hi = min(out - 1, (int)ceilf(((float)i + 1.f + 0.5f) * s - 0.5f) + 1);
}

__global__ void bilinear_bwd_w_kernel(const __nv_bfloat16* __restrict__ dout, long dout_ld, int N, int Ho, int Wo, int C,
                                      float* __restrict__ tmp, int Wi) {
  pdl_sync();
  const int cg = C >> 3;
  const long total = (long)N * Ho * Wi * cg;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int c0, wi, ho, n;
    split_idx(idx, cg, Wi, Ho, c0, wi, ho, n);
    int lo, hi;
    bilinear_src_window(wi, Wi, Wo, lo, hi);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int wo = lo; wo <= hi; ++wo) {
      int w0, w1;
      float lw;
      bilinear_coeff(wo, Wi, Wo, w0, w1, lw);
      float ww = 0.f;
      if (w0 == wi) ww += 1.f - lw;
      if (w1 == wi) ww += lw;
      if (ww == 0.f) continue;
      float g[8];
      load8(dout + (((long)n * Ho + ho) * Wo + wo) * dout_ld + c0, g);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(g[e], ww, acc[e]);
    }
    float4* tp = reinterpret_cast<float4*>(tmp + idx * 8);
    tp[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    tp[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

__global__ void bilinear_bwd_h_kernel(const float* __restrict__ tmp, int N, int Ho, int C, __nv_bfloat16* __restrict__ dx,
                                      long dx_ld, int Hi, int Wi, int accumulate) {
  pdl_sync();
  const int cg = C >> 3;
  const long total = (long)N * Hi * Wi * cg;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int c0, wi, hi, n;
    split_idx(idx, cg, Wi, Hi, c0, wi, hi, n);
    int lo, hi_;
    bilinear_src_window(hi, Hi, Ho, lo, hi_);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int ho = lo; ho <= hi_; ++ho) {
      int h0, h1;
      float lh;
      bilinear_coeff(ho, Hi, Ho, h0, h1, lh);
      float wh = 0.f;
      if (h0 == hi) wh += 1.f - lh;
      if (h1 == hi) wh += lh;
      if (wh == 0.f) continue;
      const float4* tp = reinterpret_cast<const float4*>(tmp + ((((long)n * Ho + ho) * Wi + wi) * cg + (c0 >> 3)) * 8);
      const float4 a = tp[0], b = tp[1];
      acc[0] = fmaf(a.x, wh, acc[0]), acc[1] = fmaf(a.y, wh, acc[1]), acc[2] = fmaf(a.z, wh, acc[2]);
      acc[3] = fmaf(a.w, wh, acc[3]), acc[4] = fmaf(b.x, wh, acc[4]), acc[5] = fmaf(b.y, wh, acc[5]);
      acc[6] = fmaf(b.z, wh, acc[6]), acc[7] = fmaf(b.w, wh, acc[7]);
    }
    __nv_bfloat16* dp = dx + (((long)n * Hi + hi) * Wi + wi) * dx_ld + c0;
    if (accumulate) {
      float o[8];
      load8(dp, o);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += o[e];
    }
    store8(dp, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// log-softmax + NLLLoss(ignore_index=-1) + pixel accuracy (models/models.py:12-18,37-42,492-493; train.py:154)
//   accum[0] += sum over valid pixels of -log p[label];  accum[1] += #valid;  accum[2] += #(valid & argmax==label)
// One warp per pixel; logits fp32 [P][ld]. lse[p] is kept for the backward.
__global__ void __launch_bounds__(256) softmax_nll_fwd_kernel(const float* __restrict__ logits, long ld, int C,
                                                              const long long* __restrict__ label, long P,
                                                              float* __restrict__ lse, float* __restrict__ accum) {
  pdl_sync();
  __shared__ float bl[8], bc[8], ba[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float loss = 0.f, cnt = 0.f, correct = 0.f;
  for (long p = (long)blockIdx.x * 8 + warp; p < P; p += (long)gridDim.x * 8) {
    const float* row = logits + p * ld;
    float m = -INFINITY;
    int am = 0;
    for (int c = lane; c < C; c += 32) {
      const float v = row[c];
      if (v > m) m = v, am = c;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, off);
      const int oa = __shfl_xor_sync(0xffffffffu, am, off);
      if (om > m || (om == m && oa < am)) m = om, am = oa;  // first maximum wins, like torch.max
    }
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += __expf(row[c] - m);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const float l = m + __logf(s);
    if (lane == 0) {
      lse[p] = l;
      const long long lab = label[p];
      if (lab >= 0 && lab < C) {  // labels >= C would be an error in the reference; never index out of bounds
        loss += l - row[lab];
        cnt += 1.f;
        if (am == (int)lab) correct += 1.f;
      }
    }
  }
  if (lane == 0) bl[warp] = loss, bc[warp] = cnt, ba[warp] = correct;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int i = 0; i < 8; ++i) a += bl[i], b += bc[i], c += ba[i];
    atomicAdd(accum + 0, a), atomicAdd(accum + 1, b), atomicAdd(accum + 2, c);
  }
}

// loss = main/cnt + ds_scale * ds/cnt ; acc = correct/(cnt + 1e-10)   (models/models.py:37-42, :12-18)
__global__ void nll_finalize_kernel(const float* accum_main, const float* accum_ds, float ds_scale, float* out) {
  pdl_sync();
  float loss = accum_main[0] / accum_main[1];
  if (accum_ds) loss += ds_scale * accum_ds[0] / accum_ds[1];
  out[0] = loss;
  out[1] = accum_main[2] / (accum_main[1] + 1e-10f);
}

// dlogits[p][c] = weight/cnt * (softmax - onehot) for valid pixels, 0 otherwise; bf16 [P][ld_out], columns >= C zeroed.
__global__ void __launch_bounds__(256) softmax_nll_bwd_kernel(const float* __restrict__ logits, long ld, int C,
                                                              const long long* __restrict__ label,
                                                              const float* __restrict__ lse, const float* __restrict__ accum,
                                                              float weight, long P, __nv_bfloat16* __restrict__ dlogits,
                                                              long ld_out, int c_store) {
  pdl_sync();
  const float coef = weight / accum[1];
  const long total = P * c_store;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = idx % c_store;
    const long p = idx / c_store;
    float v = 0.f;
    const long long lab = label[p];
    if (c < C && lab >= 0 && lab < C) {
      v = __expf(logits[p * ld + c] - lse[p]);
      if (c == (int)lab) v -= 1.f;
      v *= coef;
    }
    dlogits[p * ld_out + c] = __float2bfloat16(v);
  }
}

// column sums of a bf16 [P][ld] matrix into fp32[C] (+=): bias gradients.
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, long ld, long P, int C,
                                                     float* __restrict__ out) {
  pdl_sync();
  // block handles a slab of rows; thread t handles columns t, t+256, ...
  const long rows_per_block = (P + gridDim.x - 1) / gridDim.x;
  const long r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, P);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (long r = r0; r < r1; ++r) s += __bfloat162float(x[r * ld + c]);
    atomicAdd(out + c, s);
  }
}

// inference head: bilinear upsample of fp32 NHWC logits to segSize + softmax -> fp32 NCHW probabilities,
// optionally accumulated (scores += probs * weight): models/models.py:480-484, eval.py:71-72.
__global__ void __launch_bounds__(256) upsample_softmax_kernel(const float* __restrict__ logits, long ld, int N, int Hi,
                                                               int Wi, int C, float* __restrict__ probs, int Ho, int Wo,
                                                               float weight, int accumulate, int log_output) {
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long P = (long)N * Ho * Wo;
  for (long p = (long)blockIdx.x * 8 + warp; p < P; p += (long)gridDim.x * 8) {
    const int wo = p % Wo;
    long r = p / Wo;
    const int ho = r % Ho;
    const int n = r / Ho;
    int h0, h1, w0, w1;
    float lh, lw;
    bilinear_coeff(ho, Hi, Ho, h0, h1, lh);
    bilinear_coeff(wo, Wi, Wo, w0, w1, lw);
    const float* a = logits + (((long)n * Hi + h0) * Wi + w0) * ld;
    const float* b = logits + (((long)n * Hi + h0) * Wi + w1) * ld;
    const float* c = logits + (((long)n * Hi + h1) * Wi + w0) * ld;
    const float* d = logits + (((long)n * Hi + h1) * Wi + w1) * ld;
    float v[8];  // up to 256 classes
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = lane + 32 * k;
      v[k] = -INFINITY;
      if (ch < C) {
        v[k] = (1.f - lh) * ((1.f - lw) * a[ch] + lw * b[ch]) + lh * ((1.f - lw) * c[ch] + lw * d[ch]);
        m = fmaxf(m, v[k]);
      }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    float s = 0.f, dm[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      dm[k] = v[k] - m;
      v[k] = (lane + 32 * k < C) ? __expf(dm[k]) : 0.f;
      s += v[k];
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const float inv = weight / s, logs = __logf(s);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = lane + 32 * k;
      if (ch < C) {
        float* o = probs + (((long)n * C + ch) * Ho + ho) * Wo + wo;
        const float r = log_output ? (dm[k] - logs) * weight : v[k] * inv;  // log-softmax = (x - m) - log(sum)
        *o = accumulate ? *o + r : r;
      }
    }
  }
}

// layout conversion NHWC bf16 -> NCHW fp32 (feature maps handed back through the module-level API)
__global__ void nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ x, long ld, int N, int H, int W, int C,
                                             float* __restrict__ out) {
  pdl_sync();
  __shared__ float tile[32][33];
  // grid: (ceil(HW/32), ceil(C/32), N); block (32, 8)
  const int n = blockIdx.z;
  const long HW = (long)H * W;
  const long p0 = (long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const long p = p0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < HW && c < C) ? __bfloat162float(x[((long)n * HW + p) * ld + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i;
    const long p = p0 + threadIdx.x;
    if (p < HW && c < C) out[((long)n * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ x, int N, int H, int W, int C,
                                             __nv_bfloat16* __restrict__ out, long ld) {
  pdl_sync();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long HW = (long)H * W;
  const long p0 = (long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i;
    const long p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < HW && c < C) ? x[((long)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const long p = p0 + i;
    const int c = c0 + threadIdx.x;
    if (p < HW && c < C) out[((long)n * HW + p) * ld + c] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

}  // namespace sseg

using namespace sseg;
#define LAUNCH_CHECK(name)   \
  count_launch(1);           \
  return check_cuda(cudaGetLastError(), name)

extern "C" {

int sseg_prep_conv_weight(const float* w_oihw, int O, int I, int T, void* w_fwd, long fwd_ld, void* w_dgrad,
                          long dgrad_ld, int o_pad, sseg_stream_t st) {
  SSEG_REQUIRE(w_oihw && (w_fwd || w_dgrad), "sseg_prep_conv_weight: null argument");
  SSEG_REQUIRE(!w_fwd || fwd_ld >= (long)T * I, "sseg_prep_conv_weight: fwd_ld too small");
  SSEG_REQUIRE(!w_dgrad || (o_pad >= O && dgrad_ld >= (long)T * o_pad), "sseg_prep_conv_weight: dgrad_ld too small");
  const long total = (long)O * I * T;
  launch_k(prep_weight_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)st, 
      w_oihw, O, I, T, (__nv_bfloat16*)w_fwd, fwd_ld, (__nv_bfloat16*)w_dgrad, dgrad_ld, o_pad);
  LAUNCH_CHECK("prep_weight_kernel");
}

int sseg_grad_to_oihw(const float* g, long g_ld, int O, int I, int T, float* out, float scale, int accumulate,
                      sseg_stream_t st) {
  SSEG_REQUIRE(g && out, "sseg_grad_to_oihw: null argument");
  const long total = (long)O * I * T;
  launch_k(grad_to_oihw_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)st, g, g_ld, O, I, T, out, scale, accumulate);
  LAUNCH_CHECK("grad_to_oihw_kernel");
}

int sseg_stem_conv_fwd(const float* img, int N, int H, int W, const float* w, void* out, float* stat_sum,
                       float* stat_sqsum, sseg_stream_t st) {
  SSEG_REQUIRE(img && w && out, "sseg_stem_conv_fwd: null argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long P = (long)N * Ho * Wo;
  launch_k(stem_conv_fwd_kernel, dim3(N * Ho), dim3(128), 0, (cudaStream_t)st, img, w, (__nv_bfloat16*)out, stat_sum, stat_sqsum, N, H,
           W, Ho, Wo);
  LAUNCH_CHECK("stem_conv_fwd_kernel");
}

int sseg_stem_conv_wgrad(const float* img, int N, int H, int W, const void* dy, float* dw, sseg_stream_t st) {
  SSEG_REQUIRE(img && dy && dw, "sseg_stem_conv_wgrad: null argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long P = (long)N * Ho * Wo;
  launch_k(stem_conv_wgrad_kernel, dim3(N * Ho), dim3(256), 0, (cudaStream_t)st, img, (const __nv_bfloat16*)dy, dw, N, H, W, Ho, Wo);
  LAUNCH_CHECK("stem_conv_wgrad_kernel");
}

int sseg_bn_finalize(const float* sum, const float* sqsum, const float* count_dev, float count_host, const float* gamma,
                     const float* beta, float eps, float momentum, int mode, int update_running, float* running_mean,
                     float* running_var, float* tmp_mean, float* tmp_var, float* running_iter, float* mean_out,
                     float* invstd_out, float* scale, float* shift, int C, sseg_stream_t st) {
  SSEG_REQUIRE(mean_out && invstd_out && scale && shift && C > 0, "sseg_bn_finalize: null argument");
  SSEG_REQUIRE(mode == 2 || (sum && sqsum), "sseg_bn_finalize: statistics required in training modes");
  SSEG_REQUIRE(mode != 2 || (running_mean && running_var), "sseg_bn_finalize: running statistics required in eval mode");
  SSEG_REQUIRE(!(update_running && mode == 1) || (tmp_mean && tmp_var && running_iter && running_mean && running_var),
               "sseg_bn_finalize: sync-mode running buffers required");
  launch_k(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, (cudaStream_t)st, 
      sum, sqsum, count_dev, count_host, gamma, beta, eps, momentum, mode, update_running, running_mean, running_var,
      tmp_mean, tmp_var, running_iter, mean_out, invstd_out, scale, shift, C);
  count_launch(1);
  if (update_running && mode == 1) {
    launch_k(bn_iter_update_kernel, dim3(1), dim3(1), 0, (cudaStream_t)st, running_iter, momentum);
    count_launch(1);
  }
  return check_cuda(cudaGetLastError(), "bn_finalize_kernel");
}

static int bn_apply_impl(const void* y, long y_ld, const float* scale, const float* shift, const void* res, long res_ld,
                         const float* rscale, const float* rshift, const float* chanmul, void* out, long out_ld, long P,
                         long pix_per_img, int C, int relu, int res_after_relu, const float* fin_sum,
                         const float* fin_sqsum, const float* gamma, const float* beta, float count, float eps,
                         float momentum, float* mean_out, float* invstd_out, float* scale_out, float* shift_out,
                         float* running_mean, float* running_var, sseg_stream_t st) {
  SSEG_REQUIRE(y && out && C % 8 == 0 && y_ld % 8 == 0 && out_ld % 8 == 0 && (!res || res_ld % 8 == 0),
               "sseg_bn_apply: bad argument (channels and strides must be multiples of 8)");
  SSEG_REQUIRE(fin_sum != nullptr || (scale && shift), "sseg_bn_apply: scale/shift required");
  const BnTiling t = bn_tiling(P, C);
  BnApplyParams p{(const __nv_bfloat16*)y, y_ld, scale, shift, (const __nv_bfloat16*)res, res_ld, rscale, rshift,
                  chanmul, (__nv_bfloat16*)out, out_ld, P, pix_per_img, C, relu, t.cgb, t.rows, res_after_relu,
                  fin_sum, fin_sqsum, gamma, beta, count, eps, momentum, mean_out, invstd_out, scale_out, shift_out,
                  running_mean, running_var};
  launch_k(bn_apply_kernel, dim3(dim3(t.gx, t.gy)), dim3(256), 0, (cudaStream_t)st, p);
  LAUNCH_CHECK("bn_apply_kernel");
}

int sseg_bn_apply(const void* y, long y_ld, const float* scale, const float* shift, const void* res, long res_ld,
                  const float* rscale, const float* rshift, const float* chanmul, void* out, long out_ld, long P,
                  long pix_per_img, int C, int relu, int res_after_relu, sseg_stream_t st) {
  return bn_apply_impl(y, y_ld, scale, shift, res, res_ld, rscale, rshift, chanmul, out, out_ld, P, pix_per_img, C, relu,
                       res_after_relu, nullptr, nullptr, nullptr, nullptr, 1.f, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr,
                       nullptr, nullptr, st);
}

int sseg_bn_finalize_apply(const float* sum, const float* sqsum, float count, const float* gamma, const float* beta,
                           float eps, float momentum, float* running_mean, float* running_var, float* mean_out,
                           float* invstd_out, float* scale_out, float* shift_out, const void* y, long y_ld, const void* res,
                           long res_ld, const float* rscale, const float* rshift, const float* chanmul, void* out,
                           long out_ld, long P, long pix_per_img, int C, int relu, int res_after_relu, sseg_stream_t st) {
  SSEG_REQUIRE(sum && sqsum && mean_out && invstd_out && scale_out && shift_out && count > 1.f,
               "sseg_bn_finalize_apply: null argument");
  SSEG_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "sseg_bn_finalize_apply: running stats must pair");
  return bn_apply_impl(y, y_ld, nullptr, nullptr, res, res_ld, rscale, rshift, chanmul, out, out_ld, P, pix_per_img, C,
                       relu, res_after_relu, sum, sqsum, gamma, beta, count, eps, momentum, mean_out, invstd_out, scale_out,
                       shift_out, running_mean, running_var, st);
}

static int fill_bwd(BnBwdParams& p, const void* g, long g_ld, const void* a, long a_ld, const void* y, long y_ld,
                    const float* mean, const float* invstd, const float* scale, const float* fshift, const float* chanmul,
                    float* s1, float* s2, const float* count_dev, float count_host, void* dy, long dy_ld, void* dres,
                    long dres_ld, long P, long pix_per_img, int C, int eval_mode, BnTiling* t, bool reduce = false,
                    int s2_raw = 0, float* dgamma_out = nullptr) {
  SSEG_REQUIRE(g && C % 8 == 0 && g_ld % 8 == 0 && (!a || a_ld % 8 == 0) && (!y || y_ld % 8 == 0),
               "sseg_bn_bwd: bad argument (channels and strides must be multiples of 8)");
  SSEG_REQUIRE(a == nullptr || fshift == nullptr, "sseg_bn_bwd: pass either the saved output `a` or fshift, not both");
  SSEG_REQUIRE(fshift == nullptr || (scale != nullptr && y != nullptr), "sseg_bn_bwd: fshift needs scale and y");
  *t = bn_tiling(P, C, reduce);
  p = BnBwdParams{(const __nv_bfloat16*)g, g_ld, (const __nv_bfloat16*)a, a_ld, (const __nv_bfloat16*)y, y_ld, mean,
                  invstd, scale, fshift, chanmul, s1, s2, count_dev, count_host, (__nv_bfloat16*)dy, dy_ld,
                  (__nv_bfloat16*)dres, dres_ld, P, pix_per_img, C, eval_mode, t->cgb, t->rows, s2_raw, dgamma_out,
                  PeerTable{}, 0, 0, nullptr, nullptr};
  return 0;
}

int sseg_bn_bwd_reduce(const void* g, long g_ld, const void* a, long a_ld, const void* y, long y_ld, const float* mean,
                       const float* invstd, const float* scale, const float* fshift, const float* chanmul, float* s1,
                       float* s2, long P, long pix_per_img, int C, sseg_stream_t st) {
  BnBwdParams p;
  BnTiling t;
  int rc = fill_bwd(p, g, g_ld, a, a_ld, y, y_ld, mean, invstd, scale, fshift, chanmul, s1, s2, nullptr, 1.f, nullptr, 0,
                    nullptr, 0, P, pix_per_img, C, 0, &t, true);
  if (rc) return rc;
  SSEG_REQUIRE(y && mean && invstd && s1 && s2, "sseg_bn_bwd_reduce: null argument");
  launch_k(bn_bwd_kernel<false>, dim3(dim3(t.gx, t.gy)), dim3(256), 0, (cudaStream_t)st, p);
  LAUNCH_CHECK("bn_bwd_reduce_kernel");
}

int sseg_bn_bwd_apply(const void* g, long g_ld, const void* a, long a_ld, const void* y, long y_ld, const float* mean,
                      const float* invstd, const float* scale, const float* fshift, const float* chanmul, const float* s1,
                      const float* s2, const float* count_dev, float count_host, void* dy, long dy_ld, void* dres,
                      long dres_ld, long P, long pix_per_img, int C, int eval_mode, int s2_raw, float* dgamma_out,
                      sseg_stream_t st) {
  BnBwdParams p;
  BnTiling t;
  int rc = fill_bwd(p, g, g_ld, a, a_ld, y, y_ld, mean, invstd, scale, fshift, chanmul, const_cast<float*>(s1),
                    const_cast<float*>(s2), count_dev, count_host, dy, dy_ld, dres, dres_ld, P, pix_per_img, C, eval_mode,
                    &t, false, s2_raw, dgamma_out);
  if (rc) return rc;
  SSEG_REQUIRE(!s2_raw || (s1 && s2 && mean && invstd), "sseg_bn_bwd_apply: raw s2 needs s1, mean and invstd");
  SSEG_REQUIRE(dy && scale && dy_ld % 8 == 0 && (!dres || dres_ld % 8 == 0), "sseg_bn_bwd_apply: bad argument");
  SSEG_REQUIRE(eval_mode || (y && mean && invstd && s1 && s2), "sseg_bn_bwd_apply: training mode needs y/mean/invstd/s1/s2");
  launch_k(bn_bwd_kernel<true>, dim3(dim3(t.gx, t.gy)), dim3(256), 0, (cudaStream_t)st, p);
  LAUNCH_CHECK("bn_bwd_apply_kernel");
}

int sseg_bn_bwd_apply_peer(void* const* bases, int world, int rank, long part_off, long flag_off, const int* step,
                           const void* g, long g_ld, const void* a, long a_ld, const void* y, long y_ld, const float* mean,
                           const float* invstd, const float* scale, const float* fshift, const float* chanmul,
                           const float* count_dev, void* dy, long dy_ld, void* dres, long dres_ld, long P, long pix_per_img,
                           int C, int s2_raw, float* dbeta_out, float* dgamma_out, sseg_stream_t st) {
  BnBwdParams p;
  BnTiling t;
  int rc = fill_bwd(p, g, g_ld, a, a_ld, y, y_ld, mean, invstd, scale, fshift, chanmul, nullptr, nullptr, count_dev, 1.f, dy,
                    dy_ld, dres, dres_ld, P, pix_per_img, C, 0, &t, false, s2_raw, dgamma_out);
  if (rc) return rc;
  SSEG_REQUIRE(step && count_dev && y && mean && invstd && scale && dy && dbeta_out && dgamma_out && dy_ld % 8 == 0 &&
                   (!dres || dres_ld % 8 == 0),
               "sseg_bn_bwd_apply_peer: bad argument");
  rc = make_peer_table(&p.pt, bases, world, rank, "sseg_bn_bwd_apply_peer");
  if (rc) return rc;
  p.part_off = part_off, p.flag_off = flag_off, p.step = step, p.dbeta_out = dbeta_out;
  // every block pays the handshake + one NVLink round trip before it streams: ONE wave of blocks (the pixel loop is
  // grid-strided), or each further wave pays it again (measured: 4 waves cost +20 us per layer)
  {
    static const int waves = getenv("SSEG_PEER_APPLY_BLOCKS") ? atoi(getenv("SSEG_PEER_APPLY_BLOCKS")) : 2;
    const int cap = (148 * waves + t.gy - 1) / t.gy;
    if (t.gx > cap) t.gx = cap < 1 ? 1 : cap;
  }
  launch_k(bn_bwd_kernel<true>, dim3(dim3(t.gx, t.gy)), dim3(256), 2 * 256 * sizeof(float), (cudaStream_t)st, p);
  LAUNCH_CHECK("bn_bwd_apply_peer_kernel");
}

int sseg_maxpool_fwd(const void* x, int N, int H, int W, int C, void* out, void* idx, sseg_stream_t st) {
  SSEG_REQUIRE(x && out && C % 8 == 0, "sseg_maxpool_fwd: bad argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  launch_k(maxpool_fwd_kernel, dim3(grid_for((long)N * Ho * Wo * (C / 8), 256)), dim3(256), 0, (cudaStream_t)st, 
      (const __nv_bfloat16*)x, (__nv_bfloat16*)out, (uint8_t*)idx, N, H, W, C, Ho, Wo);
  LAUNCH_CHECK("maxpool_fwd_kernel");
}

int sseg_maxpool_bwd(const void* dout, const void* idx, void* dx, int N, int H, int W, int C, sseg_stream_t st) {
  SSEG_REQUIRE(dout && idx && dx && C % 8 == 0, "sseg_maxpool_bwd: bad argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  launch_k(maxpool_bwd_kernel, dim3(grid_for((long)N * H * W * (C / 8), 256)), dim3(256), 0, (cudaStream_t)st, 
      (const __nv_bfloat16*)dout, (const uint8_t*)idx, (__nv_bfloat16*)dx, N, H, W, C, Ho, Wo);
  LAUNCH_CHECK("maxpool_bwd_kernel");
}

int sseg_avgpool_fwd(const void* x, long x_ld, int N, int H, int W, int C, int S, void* out, sseg_stream_t st) {
  SSEG_REQUIRE(x && out && C % 8 == 0 && x_ld % 8 == 0 && S >= 1, "sseg_avgpool_fwd: bad argument");
  dim3 grid(N * S * S, (C / 8 + 7) / 8);
  launch_k(avgpool_fwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)st, (const __nv_bfloat16*)x, x_ld, (__nv_bfloat16*)out, N, H, W, C, S);
  LAUNCH_CHECK("avgpool_fwd_kernel");
}

int sseg_avgpool_bwd(const void* base, long base_ld, const void* const* dpool, const int* scales, int nscales, void* dx,
                     long dx_ld, int N, int H, int W, int C, sseg_stream_t st) {
  SSEG_REQUIRE(dx && nscales >= 0 && nscales <= 4 && C % 8 == 0 && dx_ld % 8 == 0 && (!base || base_ld % 8 == 0),
               "sseg_avgpool_bwd: bad argument");
  AvgPoolBwdParams p;
  memset(&p, 0, sizeof(p));
  p.base = (const __nv_bfloat16*)base, p.base_ld = base_ld;
  for (int k = 0; k < nscales; ++k) p.dpool[k] = (const __nv_bfloat16*)dpool[k], p.scales[k] = scales[k];
  p.nscales = nscales, p.dx = (__nv_bfloat16*)dx, p.dx_ld = dx_ld, p.N = N, p.H = H, p.W = W, p.C = C;
  launch_k(avgpool_bwd_kernel, dim3(N * H * ((W + kAvgWT - 1) / kAvgWT)), dim3(256), 0, (cudaStream_t)st, p);
  LAUNCH_CHECK("avgpool_bwd_kernel");
}

int sseg_bilinear_fwd(const void* x, long x_ld, int N, int Hi, int Wi, int C, void* out, long out_ld, int Ho, int Wo,
                      sseg_stream_t st) {
  SSEG_REQUIRE(x && out && C % 8 == 0 && x_ld % 8 == 0 && out_ld % 8 == 0, "sseg_bilinear_fwd: bad argument");
  launch_k(bilinear_fwd_kernel, dim3(grid_for((long)N * Ho * Wo * (C / 8), 256)), dim3(256), 0, (cudaStream_t)st, 
      (const __nv_bfloat16*)x, x_ld, N, Hi, Wi, C, (__nv_bfloat16*)out, out_ld, Ho, Wo);
  LAUNCH_CHECK("bilinear_fwd_kernel");
}

int sseg_sum_terms(const sseg_sum_term_t* terms, int nterms, int N, int Ho, int Wo, int C, void* out, long out_ld, int relu,
                   sseg_stream_t st) {
  SSEG_REQUIRE(terms && out && nterms >= 1 && nterms <= SSEG_MAX_SUM_TERMS, "sseg_sum_terms: 1..%d terms", SSEG_MAX_SUM_TERMS);
  SSEG_REQUIRE(C % 8 == 0 && out_ld % 8 == 0 && N >= 1 && Ho >= 1 && Wo >= 1, "sseg_sum_terms: bad shape");
  SumTermsParams p;
  memset(&p, 0, sizeof(p));
  for (int k = 0; k < nterms; ++k) {
    const sseg_sum_term_t& t = terms[k];
    SSEG_REQUIRE(t.x && t.h >= 1 && t.w >= 1 && t.ld % 8 == 0 && t.ld >= C, "sseg_sum_terms: term %d invalid", k);
    SSEG_REQUIRE((t.scale == nullptr) == (t.shift == nullptr), "sseg_sum_terms: term %d scale/shift must pair", k);
    SSEG_REQUIRE(!t.scale || ((reinterpret_cast<uintptr_t>(t.scale) | reinterpret_cast<uintptr_t>(t.shift)) & 15) == 0,
                 "sseg_sum_terms: term %d scale/shift not 16B aligned", k);
    p.x[k] = static_cast<const __nv_bfloat16*>(t.x), p.scale[k] = t.scale, p.shift[k] = t.shift;
    p.h[k] = t.h, p.w[k] = t.w, p.ld[k] = t.ld;
  }
  p.nterms = nterms, p.N = N, p.Ho = Ho, p.Wo = Wo, p.C = C, p.relu = relu;
  p.out = static_cast<__nv_bfloat16*>(out), p.out_ld = out_ld;
  launch_k(sum_terms_kernel, dim3(grid_for((long)N * Ho * Wo * (C / 8), 256)), dim3(256), 0, (cudaStream_t)st, p);
  LAUNCH_CHECK("sum_terms_kernel");
}

int sseg_relu_mask_bwd(const void* g, long g_ld, const void* out, long out_ld, void* ds, long ds_ld, void* acc_out,
                       long acc_ld, int accumulate, long P, int C, sseg_stream_t st) {
  SSEG_REQUIRE(g && out && ds && C % 8 == 0 && g_ld % 8 == 0 && out_ld % 8 == 0 && ds_ld % 8 == 0 &&
                   (!acc_out || acc_ld % 8 == 0) && P >= 1,
               "sseg_relu_mask_bwd: bad argument");
  launch_k(relu_mask_bwd_kernel, dim3(grid_for(P * (C / 8), 256)), dim3(256), 0, (cudaStream_t)st,
           (const __nv_bfloat16*)g, g_ld, (const __nv_bfloat16*)out, out_ld, (__nv_bfloat16*)ds, ds_ld,
           (__nv_bfloat16*)acc_out, acc_ld, accumulate, P, C);
  LAUNCH_CHECK("relu_mask_bwd_kernel");
}

int sseg_bilinear_bwd(const void* dout, long dout_ld, int N, int Ho, int Wo, int C, void* dx, long dx_ld, int Hi, int Wi,
                      int accumulate, float* scratch, sseg_stream_t st) {
  SSEG_REQUIRE(dout && dx && scratch && C % 8 == 0 && dout_ld % 8 == 0 && dx_ld % 8 == 0,
               "sseg_bilinear_bwd: bad argument (scratch = float[N*Ho*Wi*C] required)");
  SSEG_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 15) == 0, "sseg_bilinear_bwd: scratch not 16B aligned");
  launch_k(bilinear_bwd_w_kernel, dim3(grid_for((long)N * Ho * Wi * (C / 8), 128)), dim3(128), 0, (cudaStream_t)st, 
      (const __nv_bfloat16*)dout, dout_ld, N, Ho, Wo, C, scratch, Wi);
  launch_k(bilinear_bwd_h_kernel, dim3(grid_for((long)N * Hi * Wi * (C / 8), 128)), dim3(128), 0, (cudaStream_t)st, 
      scratch, N, Ho, C, (__nv_bfloat16*)dx, dx_ld, Hi, Wi, accumulate);
  count_launch(2);
  return check_cuda(cudaGetLastError(), "bilinear_bwd kernels");
}

int sseg_softmax_nll_fwd(const float* logits, long ld, int C, const long long* label, long P, float* lse, float* accum,
                         sseg_stream_t st) {
  SSEG_REQUIRE(logits && label && lse && accum && C >= 1, "sseg_softmax_nll_fwd: bad argument");
  launch_k(softmax_nll_fwd_kernel, dim3(grid_for(P, 8, 148 * 8)), dim3(256), 0, (cudaStream_t)st, logits, ld, C, label, P, lse, accum);
  LAUNCH_CHECK("softmax_nll_fwd_kernel");
}

int sseg_nll_finalize(const float* accum_main, const float* accum_ds, float ds_scale, float* out, sseg_stream_t st) {
  SSEG_REQUIRE(accum_main && out, "sseg_nll_finalize: bad argument");
  launch_k(nll_finalize_kernel, dim3(1), dim3(1), 0, (cudaStream_t)st, accum_main, accum_ds, ds_scale, out);
  LAUNCH_CHECK("nll_finalize_kernel");
}

int sseg_softmax_nll_bwd(const float* logits, long ld, int C, const long long* label, const float* lse, const float* accum,
                         float weight, long P, void* dlogits, long ld_out, int c_store, sseg_stream_t st) {
  SSEG_REQUIRE(logits && label && lse && accum && dlogits && c_store >= C && c_store <= ld_out,
               "sseg_softmax_nll_bwd: bad argument");
  launch_k(softmax_nll_bwd_kernel, dim3(grid_for(P * c_store, 256)), dim3(256), 0, (cudaStream_t)st, 
      logits, ld, C, label, lse, accum, weight, P, (__nv_bfloat16*)dlogits, ld_out, c_store);
  LAUNCH_CHECK("softmax_nll_bwd_kernel");
}

int sseg_colsum(const void* x, long ld, long P, int C, float* out, sseg_stream_t st) {
  SSEG_REQUIRE(x && out, "sseg_colsum: bad argument");
  launch_k(colsum_kernel, dim3(grid_for(P, 64, 148 * 2)), dim3(256), 0, (cudaStream_t)st, (const __nv_bfloat16*)x, ld, P, C, out);
  LAUNCH_CHECK("colsum_kernel");
}

int sseg_upsample_softmax(const float* logits, long ld, int N, int Hi, int Wi, int C, float* probs, int Ho, int Wo,
                          float weight, int accumulate, int log_output, sseg_stream_t st) {
  SSEG_REQUIRE(logits && probs && C >= 1 && C <= 256, "sseg_upsample_softmax: bad argument (C <= 256)");
  launch_k(upsample_softmax_kernel, dim3(grid_for((long)N * Ho * Wo, 8, 148 * 16)), dim3(256), 0, (cudaStream_t)st, 
      logits, ld, N, Hi, Wi, C, probs, Ho, Wo, weight, accumulate, log_output);
  LAUNCH_CHECK("upsample_softmax_kernel");
}

int sseg_nhwc_bf16_to_nchw_f32(const void* x, long ld, int N, int H, int W, int C, float* out, sseg_stream_t st) {
  SSEG_REQUIRE(x && out, "sseg_nhwc_bf16_to_nchw_f32: bad argument");
  dim3 grid((unsigned)(((long)H * W + 31) / 32), (C + 31) / 32, N), block(32, 8);
  launch_k(nhwc_bf16_to_nchw_f32_kernel, dim3(grid), dim3(block), 0, (cudaStream_t)st, (const __nv_bfloat16*)x, ld, N, H, W, C, out);
  LAUNCH_CHECK("nhwc_bf16_to_nchw_f32_kernel");
}

int sseg_nchw_f32_to_nhwc_bf16(const float* x, int N, int H, int W, int C, void* out, long ld, sseg_stream_t st) {
  SSEG_REQUIRE(x && out, "sseg_nchw_f32_to_nhwc_bf16: bad argument");
  dim3 grid((unsigned)(((long)H * W + 31) / 32), (C + 31) / 32, N), block(32, 8);
  launch_k(nchw_f32_to_nhwc_bf16_kernel, dim3(grid), dim3(block), 0, (cudaStream_t)st, x, N, H, W, C, (__nv_bfloat16*)out, ld);
  LAUNCH_CHECK("nchw_f32_to_nhwc_bf16_kernel");
}

}  // extern "C"
