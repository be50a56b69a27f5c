// Batched weight re-layout: ONE launch converts every conv of the model.
//   prep : fp32 master weights -> bf16 [O][T*I] (forward operand) and bf16 [I][T*Opad] (data-gradient operand). The master
//          is OIHW-contiguous, or (descriptor flag `reserved` = 1) channels-last = [O][T][I] in memory: then the forward
//          operand is a plain cast and the weight gradient the GEMM produces IS the parameter's gradient layout - the
//          engine keeps its convolution parameters that way, which removes the gradient re-layout below from the step
//   grad : fp32 [O][T*I] weight gradients (sseg_conv_wgrad layout) -> fp32 OIHW, times scale (OIHW masters only)
// Each CTA owns a 32(o) x 32(i) x T tile, staged through shared memory so that global reads AND writes are coalesced
// in both layouts. The per-conv descriptors live in a device table built once by the caller.
#include "common.h"
#include "ptx.cuh"
#include <cuda_bf16.h>

namespace sseg {

constexpr int kTile = 32;
constexpr int kMaxT = 9;

__device__ __forceinline__ const sseg_weight_desc_t* find_desc(const sseg_weight_desc_t* table, int n, int tile,
                                                               int& local) {
  int lo = 0, hi = n - 1;  // last descriptor whose first_tile <= tile (binary search: ~6 dependent loads, not n)
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(&table[mid].first_tile) <= tile) lo = mid; else hi = mid - 1;
  }
  local = tile - __ldg(&table[lo].first_tile);
  return table + lo;
}

// i-tiles handled by one CTA: pointwise convs (T = 1) have tiny 32x32 tiles, so a CTA walks 8 of them
__host__ __device__ inline int i_tiles_per_cta(int T) { return T == 1 ? 8 : 1; }

// The grid may be smaller than the number of tiles (blocks then walk the tiles with a grid stride): a thin grid on a side
// stream leaves the SMs to the main stream's kernels - a launch with thousands of pending blocks would make every later
// kernel of the other stream wait until its last block has been dispatched.
__global__ void __launch_bounds__(256) weights_batched_kernel(const sseg_weight_desc_t* __restrict__ table, int n,
                                                              int total_tiles, int mode, float scale) {
  pdl_sync();
  __shared__ float tile[kTile][kTile * kMaxT + 1];
  for (int tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x) {
  int local;
  const sseg_weight_desc_t* d = find_desc(table, n, tile_id, local);
  const int O = d->O, I = d->I, T = d->T;
  const int tiles_i = (I + kTile - 1) / kTile;
  const int rep = i_tiles_per_cta(T);
  const int ctas_i = (tiles_i + rep - 1) / rep;
  const int o0 = (local / ctas_i) * kTile;
  const int no = min(kTile, O - o0);
  if (mode == 0 && T == 1) {
    // ---- pointwise convolutions ([O][I] master): the CTA's `rep` consecutive 32-wide i-tiles are staged AT ONCE, as if
    //      they were taps - one load phase with 8 loads in flight per row, one barrier, one store phase. (Walking them one
    //      after the other - load, barrier, store, barrier, eight times - made these CTAs live 24 us for 64 KB of traffic.)
    const int it0 = (local % ctas_i) * rep;
    const int K = min(rep, tiles_i - it0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ibase = it0 * kTile + lane;
    __syncthreads();  // the shared tile is reused across the grid-stride loop
    const float* w = d->w;
    for (int ol = warp; ol < no; ol += 8) {
      const float* src = w + (long)(o0 + ol) * I + ibase;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < K && ibase + k * kTile < I) tile[ol][k * kTile + lane] = src[k * kTile];
    }
    __syncthreads();
    __nv_bfloat16* wf = static_cast<__nv_bfloat16*>(d->wf);
    if (wf) {
      for (int ol = warp; ol < no; ol += 8) {
        __nv_bfloat16* dst = wf + (long)(o0 + ol) * d->fwd_ld + ibase;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < K && ibase + k * kTile < I) dst[k * kTile] = __float2bfloat16(tile[ol][k * kTile + lane]);
      }
    }
    __nv_bfloat16* wd = static_cast<__nv_bfloat16*>(d->wd);
    if (wd && lane < no) {
      for (int il = warp; il < K * kTile; il += 8) {
        const int i = it0 * kTile + il;
        if (i < I) wd[(long)i * d->dgrad_ld + o0 + lane] = __float2bfloat16(tile[lane][il]);
      }
    }
    continue;
  }
  for (int it = (local % ctas_i) * rep; it < min(tiles_i, (local % ctas_i) * rep + rep); ++it) {
  const int i0 = it * kTile;
  const int ni = min(kTile, I - i0);
  const int row = ni * T;
  __syncthreads();  // the shared tile is reused across iterations
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (mode == 0 && d->wd == nullptr && d->wf != nullptr && (d->reserved == 1 || T == 1) && (ni & 7) == 0 && (I & 7) == 0 &&
      (d->fwd_ld & 7) == 0) {
    // ---- forward operand only, rows of i contiguous in the master ([O][T][I], or OIHW with T = 1): a plain cast, 8
    //      elements per thread and step straight from registers (two 16-byte loads -> one 16-byte store), no staging
    const int nv = ni >> 3;
    const float* w = d->w;
    __nv_bfloat16* wf = static_cast<__nv_bfloat16*>(d->wf);
    for (int v = threadIdx.x; v < no * T * nv; v += 256) {
      const int iv = v % nv, r = v / nv;
      const int t = r % T, ol = r / T;
      const float* src = w + ((long)(o0 + ol) * T + t) * I + i0 + iv * 8;
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      uint4 q;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
      h[0] = __floats2bfloat162_rn(a.x, a.y), h[1] = __floats2bfloat162_rn(a.z, a.w);
      h[2] = __floats2bfloat162_rn(b.x, b.y), h[3] = __floats2bfloat162_rn(b.z, b.w);
      *reinterpret_cast<uint4*>(wf + (long)(o0 + ol) * d->fwd_ld + (long)t * I + i0 + iv * 8) = q;
    }
  } else if (mode == 0 && d->reserved == 1) {
    // ---- channels-last master [O][T][I]: rows of 32 consecutive i (128 B) per (o, t); tile[ol][t * 32 + il]
    const float* w = d->w;
    for (int ol = warp; ol < no; ol += 8) {
      const float* src = w + (long)(o0 + ol) * T * I + i0 + lane;
#pragma unroll
      for (int t = 0; t < kMaxT; ++t)
        if (t < T && lane < ni) tile[ol][t * kTile + lane] = src[(long)t * I];
    }
    __syncthreads();
    __nv_bfloat16* wf = static_cast<__nv_bfloat16*>(d->wf);
    if (wf && lane < ni) {
      for (int ol = warp; ol < no; ol += 8) {
        __nv_bfloat16* dst = wf + (long)(o0 + ol) * d->fwd_ld + i0 + lane;
#pragma unroll
        for (int t = 0; t < kMaxT; ++t)
          if (t < T) dst[(long)t * I] = __float2bfloat16(tile[ol][t * kTile + lane]);
      }
    }
    __nv_bfloat16* wd = static_cast<__nv_bfloat16*>(d->wd);
    if (wd && lane < no) {
      for (int il = warp; il < ni; il += 8) {
        __nv_bfloat16* dst = wd + (long)(i0 + il) * d->dgrad_ld + o0 + lane;
#pragma unroll
        for (int t = 0; t < kMaxT; ++t)
          if (t < T) dst[(long)t * d->o_pad] = __float2bfloat16(tile[lane][t * kTile + il]);
      }
    }
  } else if (mode == 0) {
    // ---- load OIHW: for a fixed o the (i, t) range is contiguous
    const float* w = d->w;
    for (int ol = warp; ol < no; ol += 8) {
      const float* src = w + ((long)(o0 + ol) * I + i0) * T;
#pragma unroll
      for (int k = 0; k < kMaxT; ++k) {  // row <= 32 * kMaxT: all loads of a row are issued back to back
        const int r = lane + 32 * k;
        if (r < row) tile[ol][r] = src[r];
      }
    }
    __syncthreads();
    __nv_bfloat16* wf = static_cast<__nv_bfloat16*>(d->wf);
    if (wf && lane < ni) {
      for (int ol = warp; ol < no; ol += 8) {
        __nv_bfloat16* dst = wf + (long)(o0 + ol) * d->fwd_ld + i0 + lane;
#pragma unroll
        for (int t = 0; t < kMaxT; ++t)
          if (t < T) dst[(long)t * I] = __float2bfloat16(tile[ol][lane * T + t]);
      }
    }
    __nv_bfloat16* wd = static_cast<__nv_bfloat16*>(d->wd);
    if (wd && lane < no) {
      for (int il = warp; il < ni; il += 8) {
        __nv_bfloat16* dst = wd + (long)(i0 + il) * d->dgrad_ld + o0 + lane;
#pragma unroll
        for (int t = 0; t < kMaxT; ++t)
          if (t < T) dst[(long)t * d->o_pad] = __float2bfloat16(tile[lane][il * T + t]);
      }
    }
  } else {
    // ---- gradients: load [O][T*I] (i fastest), store OIHW ((i,t) contiguous per o)
    const float* g = d->g_src;
    if (lane < ni) {
      for (int ol = warp; ol < no; ol += 8) {
        const float* src = g + (long)(o0 + ol) * d->g_ld + i0 + lane;
#pragma unroll
        for (int t = 0; t < kMaxT; ++t)
          if (t < T) tile[ol][lane * T + t] = src[(long)t * I];
      }
    }
    __syncthreads();
    float* out = d->g_dst;
    for (int ol = warp; ol < no; ol += 8) {
      float* dst = out + ((long)(o0 + ol) * I + i0) * T;
#pragma unroll
      for (int k = 0; k < kMaxT; ++k) {
        const int r = lane + 32 * k;
        if (r < row) dst[r] = tile[ol][r] * scale;
      }
    }
  }
  }  // i-tile loop
  }  // grid-stride loop over the tiles
}

}  // namespace sseg

using namespace sseg;

static int launch_batched(const sseg_weight_desc_t* table_dev, int n, int total_tiles, int max_blocks, int mode,
                          float scale, cudaStream_t st, const char* who) {
  SSEG_REQUIRE(table_dev != nullptr && n >= 1 && total_tiles >= 1 && max_blocks >= 0, "%s: bad argument", who);
  const int grid = (max_blocks > 0 && max_blocks < total_tiles) ? max_blocks : total_tiles;
  launch_k(weights_batched_kernel, dim3(grid), dim3(256), 0, st, table_dev, n, total_tiles, mode, scale);
  count_launch(1);
  return check_cuda(cudaGetLastError(), who);
}

extern "C" int sseg_prep_conv_weights_batched(const sseg_weight_desc_t* table_dev, int n, int total_tiles,
                                              sseg_stream_t st) {
  return launch_batched(table_dev, n, total_tiles, 0, 0, 1.f, (cudaStream_t)st, "sseg_prep_conv_weights_batched");
}

extern "C" int sseg_prep_conv_weights_batched_ex(const sseg_weight_desc_t* table_dev, int n, int total_tiles,
                                                 int max_blocks, sseg_stream_t st) {
  return launch_batched(table_dev, n, total_tiles, max_blocks, 0, 1.f, (cudaStream_t)st, "sseg_prep_conv_weights_batched_ex");
}

extern "C" int sseg_grads_to_oihw_batched(const sseg_weight_desc_t* table_dev, int n, int total_tiles, float scale,
                                          sseg_stream_t st) {
  return launch_batched(table_dev, n, total_tiles, 0, 1, scale, (cudaStream_t)st, "sseg_grads_to_oihw_batched");
}
