// Multi-tensor SGD with momentum and weight decay: ONE launch updates every parameter of the model.
// torch.optim.SGD semantics (the optimizer train.py builds, train.py:115-127):
//     d = g + wd * p ;  buf = momentum * buf + d   (buf = d on the first step) ;  p -= lr * buf
// The per-tensor work is described by a device table of chunks (<= 65536 elements each) built once by the caller.
#include "common.h"
#include "ptx.cuh"

namespace sseg {

__global__ void __launch_bounds__(256) sgd_chunks_kernel(const sseg_sgd_chunk_t* __restrict__ chunks, float lr,
                                                         float momentum, int first_step) {
  pdl_sync();
  const sseg_sgd_chunk_t c = chunks[blockIdx.x];
  float* __restrict__ p = c.param;
  const float* __restrict__ g = c.grad;
  float* __restrict__ b = c.momentum_buf;
  const float wd = c.weight_decay;
  const int n = c.n;
  if (c.vec4) {
    // 4 independent float4 triples per thread in flight (all loads of an iteration are issued before the first use)
    const int n4 = n >> 2;
    constexpr int U = 4;
    for (int i0 = threadIdx.x; i0 < n4; i0 += 256 * U) {
      float4 pv[U], gv[U], bv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 256;
        if (i < n4) {
          pv[u] = reinterpret_cast<float4*>(p)[i];
          gv[u] = reinterpret_cast<const float4*>(g)[i];
          bv[u] = first_step ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(b)[i];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 256;
        if (i < n4) {
          float4 q = pv[u], m = bv[u];
          const float4 d = gv[u];
          m.x = momentum * m.x + fmaf(wd, q.x, d.x), m.y = momentum * m.y + fmaf(wd, q.y, d.y);
          m.z = momentum * m.z + fmaf(wd, q.z, d.z), m.w = momentum * m.w + fmaf(wd, q.w, d.w);
          q.x -= lr * m.x, q.y -= lr * m.y, q.z -= lr * m.z, q.w -= lr * m.w;
          reinterpret_cast<float4*>(b)[i] = m;
          reinterpret_cast<float4*>(p)[i] = q;
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < n; i += 256) {
      const float pv = p[i];
      const float bv = (first_step ? 0.f : momentum * b[i]) + fmaf(wd, pv, g[i]);
      b[i] = bv;
      p[i] = pv - lr * bv;
    }
  }
}

// x *= *scalar, skipped altogether when the scalar is exactly 1 (the gradient `loss.backward()` seeds): the engine's
// gradients are d loss / d param already, autograd's incoming grad_output only rescales them (engine/functional.py).
__global__ void __launch_bounds__(256) scale_by_scalar_kernel(float* __restrict__ x, long n, const float* __restrict__ s) {
  pdl_sync();
  const float v = *s;
  if (v == 1.f) return;
  const long n4 = n >> 2;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long)gridDim.x * 256) {
    float4 q = reinterpret_cast<float4*>(x)[i];
    q.x *= v, q.y *= v, q.z *= v, q.w *= v;
    reinterpret_cast<float4*>(x)[i] = q;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) x[(n4 << 2) + threadIdx.x] *= v;
}

}  // namespace sseg

using namespace sseg;

extern "C" int sseg_scale_by_scalar(float* x, long n, const float* scalar_dev, sseg_stream_t st) {
  SSEG_REQUIRE(x != nullptr && scalar_dev != nullptr && n >= 0, "sseg_scale_by_scalar: bad argument");
  SSEG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "sseg_scale_by_scalar: x must be 16-byte aligned");
  if (n == 0) return 0;
  count_launch(1);
  const long blocks = (n / 4 + 255) / 256;
  return check_cuda(launch_k(scale_by_scalar_kernel, dim3((unsigned)(blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks))),
                             dim3(256), 0, (cudaStream_t)st, x, n, scalar_dev),
                    "scale_by_scalar_kernel");
}

extern "C" int sseg_sgd_step(const sseg_sgd_chunk_t* chunks_dev, int nchunks, float lr, float momentum, int first_step,
                             sseg_stream_t st) {
  SSEG_REQUIRE(chunks_dev != nullptr && nchunks >= 1, "sseg_sgd_step: bad argument");
  count_launch(1);
  return check_cuda(launch_k(sgd_chunks_kernel, dim3(nchunks), dim3(256), 0, (cudaStream_t)st, chunks_dev, lr, momentum,
                             first_step),
                    "sgd_chunks_kernel");
}
