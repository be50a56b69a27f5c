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
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 pv = reinterpret_cast<float4*>(p)[i];
      const float4 gv = reinterpret_cast<const float4*>(g)[i];
      float4 bv = first_step ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(b)[i];
      bv.x = momentum * bv.x + fmaf(wd, pv.x, gv.x), bv.y = momentum * bv.y + fmaf(wd, pv.y, gv.y);
      bv.z = momentum * bv.z + fmaf(wd, pv.z, gv.z), bv.w = momentum * bv.w + fmaf(wd, pv.w, gv.w);
      pv.x -= lr * bv.x, pv.y -= lr * bv.y, pv.z -= lr * bv.z, pv.w -= lr * bv.w;
      reinterpret_cast<float4*>(b)[i] = bv;
      reinterpret_cast<float4*>(p)[i] = pv;
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

}  // namespace sseg

using namespace sseg;

extern "C" int sseg_sgd_step(const sseg_sgd_chunk_t* chunks_dev, int nchunks, float lr, float momentum, int first_step,
                             sseg_stream_t st) {
  SSEG_REQUIRE(chunks_dev != nullptr && nchunks >= 1, "sseg_sgd_step: bad argument");
  count_launch(1);
  return check_cuda(launch_k(sgd_chunks_kernel, dim3(nchunks), dim3(256), 0, (cudaStream_t)st, chunks_dev, lr, momentum,
                             first_step),
                    "sgd_chunks_kernel");
}
