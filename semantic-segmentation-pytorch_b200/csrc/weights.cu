// Batched weight re-layout: ONE launch converts every conv of the model.
//   prep : fp32 OIHW master weights -> bf16 [O][T*I] (forward operand) and bf16 [I][T*Opad] (data-gradient operand)
//   grad : fp32 [O][T*I] weight gradients (sseg_conv_wgrad layout) -> fp32 OIHW, times scale
// Each CTA owns a 32(o) x 32(i) x T tile, staged through shared memory so that global reads AND writes are coalesced
// in both layouts. The per-conv descriptors live in a device table built once by the caller.
#include "common.h"
#include <cuda_bf16.h>

namespace sseg {

constexpr int kTile = 32;
constexpr int kMaxT = 9;

__device__ __forceinline__ const sseg_weight_desc_t* find_desc(const sseg_weight_desc_t* table, int n, int tile,
                                                               int& local) {
  int lo = 0;
  for (int k = 1; k < n; ++k)
    if (table[k].first_tile <= tile) lo = k;
  local = tile - table[lo].first_tile;
  return table + lo;
}

__global__ void __launch_bounds__(256) weights_batched_kernel(const sseg_weight_desc_t* __restrict__ table, int n,
                                                              int mode, float scale) {
  __shared__ float tile[kTile][kTile * kMaxT + 1];
  int local;
  const sseg_weight_desc_t* d = find_desc(table, n, blockIdx.x, local);
  const int O = d->O, I = d->I, T = d->T;
  const int tiles_i = (I + kTile - 1) / kTile;
  const int o0 = (local / tiles_i) * kTile, i0 = (local % tiles_i) * kTile;
  const int no = min(kTile, O - o0), ni = min(kTile, I - i0);
  const int row = ni * T;
  if (mode == 0) {
    // ---- load OIHW: for a fixed o the (i, t) range is contiguous
    const float* w = d->w;
    for (int idx = threadIdx.x; idx < no * row; idx += 256) {
      const int ol = idx / row, r = idx % row;
      tile[ol][r] = w[((long)(o0 + ol) * I + i0) * T + r];
    }
    __syncthreads();
    __nv_bfloat16* wf = static_cast<__nv_bfloat16*>(d->wf);
    if (wf) {
      for (int idx = threadIdx.x; idx < no * T * ni; idx += 256) {
        const int il = idx % ni, t = (idx / ni) % T, ol = idx / (ni * T);
        wf[(long)(o0 + ol) * d->fwd_ld + (long)t * I + i0 + il] = __float2bfloat16(tile[ol][il * T + t]);
      }
    }
    __nv_bfloat16* wd = static_cast<__nv_bfloat16*>(d->wd);
    if (wd) {
      for (int idx = threadIdx.x; idx < ni * T * no; idx += 256) {
        const int ol = idx % no, t = (idx / no) % T, il = idx / (no * T);
        wd[(long)(i0 + il) * d->dgrad_ld + (long)t * d->o_pad + o0 + ol] = __float2bfloat16(tile[ol][il * T + t]);
      }
    }
  } else {
    // ---- gradients: load [O][T*I] (i fastest), store OIHW ((i,t) contiguous per o)
    const float* g = d->g_src;
    for (int idx = threadIdx.x; idx < no * T * ni; idx += 256) {
      const int il = idx % ni, t = (idx / ni) % T, ol = idx / (ni * T);
      tile[ol][il * T + t] = g[(long)(o0 + ol) * d->g_ld + (long)t * I + i0 + il];
    }
    __syncthreads();
    float* out = d->g_dst;
    for (int idx = threadIdx.x; idx < no * row; idx += 256) {
      const int ol = idx / row, r = idx % row;
      out[((long)(o0 + ol) * I + i0) * T + r] = tile[ol][r] * scale;
    }
  }
}

}  // namespace sseg

using namespace sseg;

static int launch_batched(const sseg_weight_desc_t* table_dev, int n, int total_tiles, int mode, float scale,
                          cudaStream_t st, const char* who) {
  SSEG_REQUIRE(table_dev != nullptr && n >= 1 && total_tiles >= 1, "%s: bad argument", who);
  weights_batched_kernel<<<total_tiles, 256, 0, st>>>(table_dev, n, mode, scale);
  count_launch(1);
  return check_cuda(cudaGetLastError(), who);
}

extern "C" int sseg_prep_conv_weights_batched(const sseg_weight_desc_t* table_dev, int n, int total_tiles,
                                              sseg_stream_t st) {
  return launch_batched(table_dev, n, total_tiles, 0, 1.f, (cudaStream_t)st, "sseg_prep_conv_weights_batched");
}

extern "C" int sseg_grads_to_oihw_batched(const sseg_weight_desc_t* table_dev, int n, int total_tiles, float scale,
                                          sseg_stream_t st) {
  return launch_batched(table_dev, n, total_tiles, 1, scale, (cudaStream_t)st, "sseg_grads_to_oihw_batched");
}
