// Input pipeline kernels: the reference's per-sample CPU transforms (mit_semseg/dataset.py:53-63) as
// two streaming kernels, so that uint8 bytes - not fp32 tensors - cross PCIe and the loader workers only decode / resize.
// Pure HBM-bound byte work: 3 B in / 12 B out per pixel, every thread moves 4 pixels with 32-bit loads and float4 stores.
#include "common.h"
#include "ptx.cuh"

namespace sseg {

struct ImageTransformParams {
  const uint8_t* img;  // [N][H][W][3]
  float* out;          // [N][3][H][W]
  const int* valid;    // [N][2] rows, cols
  int N, H, W;
  float mean[3], stdv[3];
};

__global__ void __launch_bounds__(256) image_transform_kernel(const __grid_constant__ ImageTransformParams p) {
  pdl_sync();
  const int groups_per_row = p.W >> 2;
  const long total = static_cast<long>(p.N) * p.H * groups_per_row;
  for (long g = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; g < total; g += static_cast<long>(gridDim.x) * blockDim.x) {
    const int gx = static_cast<int>(g % groups_per_row);
    const long rowi = g / groups_per_row;
    const int y = static_cast<int>(rowi % p.H), n = static_cast<int>(rowi / p.H);
    const int vh = p.valid[2 * n], vw = p.valid[2 * n + 1];
    const int x0 = gx * 4;
    // 4 pixels = 12 bytes = three aligned 32-bit words (W % 4 == 0 keeps every row 4-byte aligned)
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.img + ((static_cast<long>(n) * p.H + y) * p.W + x0) * 3);
    const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
    const uint8_t b[12] = {static_cast<uint8_t>(w0),       static_cast<uint8_t>(w0 >> 8),  static_cast<uint8_t>(w0 >> 16),
                           static_cast<uint8_t>(w0 >> 24), static_cast<uint8_t>(w1),       static_cast<uint8_t>(w1 >> 8),
                           static_cast<uint8_t>(w1 >> 16), static_cast<uint8_t>(w1 >> 24), static_cast<uint8_t>(w2),
                           static_cast<uint8_t>(w2 >> 8),  static_cast<uint8_t>(w2 >> 16), static_cast<uint8_t>(w2 >> 24)};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // float32(x) / 255 ; (x - mean) / std : the reference's operations, correctly rounded one by one (no contraction)
        const float x = __fdiv_rn(static_cast<float>(b[3 * i + c]), 255.f);
        v[i] = (y < vh && x0 + i < vw) ? __fdiv_rn(__fsub_rn(x, p.mean[c]), p.stdv[c]) : 0.f;
      }
      *reinterpret_cast<float4*>(p.out + ((static_cast<long>(n) * 3 + c) * p.H + y) * p.W + x0) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

__global__ void __launch_bounds__(256) label_transform_kernel(const uint8_t* __restrict__ seg, long long* __restrict__ out,
                                                              const int* __restrict__ valid, int N, int Hs, int Ws, int rate) {
  pdl_sync();
  const long total = static_cast<long>(N) * Hs * Ws;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % Ws);
    const long r = i / Ws;
    const int y = static_cast<int>(r % Hs), n = static_cast<int>(r / Hs);
    const int vh = (valid[2 * n] + rate - 1) / rate, vw = (valid[2 * n + 1] + rate - 1) / rate;
    out[i] = (y < vh && x < vw) ? static_cast<long long>(seg[i]) - 1 : 0;
  }
}

}  // namespace sseg

extern "C" int sseg_image_transform(const void* img_u8, int N, int H, int W, const int* valid_hw, const float* mean_std,
                                    float* out, sseg_stream_t stream) {
  using namespace sseg;
  SSEG_REQUIRE(img_u8 && valid_hw && mean_std && out && N > 0 && H > 0 && W > 0, "sseg_image_transform: null / empty argument");
  SSEG_REQUIRE(W % 4 == 0, "sseg_image_transform: W = %d must be a multiple of 4 (batches are padded to the network's stride)", W);
  SSEG_REQUIRE((reinterpret_cast<uintptr_t>(img_u8) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "sseg_image_transform: image must be 4-byte, output 16-byte aligned");
  ImageTransformParams p;
  p.img = static_cast<const uint8_t*>(img_u8), p.out = out, p.valid = valid_hw, p.N = N, p.H = H, p.W = W;
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = mean_std[c], p.stdv[c] = mean_std[3 + c];
    SSEG_REQUIRE(p.stdv[c] != 0.f, "sseg_image_transform: std[%d] = 0", c);
  }
  const long groups = static_cast<long>(N) * H * (W / 4);
  const int grid = static_cast<int>(groups / 256 + 1 < 148L * 8 ? groups / 256 + 1 : 148L * 8);
  SSEG_CUDA(launch_k(image_transform_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), p));
  count_launch(1);
  return 0;
}

extern "C" int sseg_label_transform(const void* seg_u8, int N, int Hs, int Ws, const int* valid_hw, int rate, long long* out,
                                    sseg_stream_t stream) {
  using namespace sseg;
  SSEG_REQUIRE(seg_u8 && valid_hw && out && N > 0 && Hs > 0 && Ws > 0 && rate > 0, "sseg_label_transform: null / empty argument");
  const long total = static_cast<long>(N) * Hs * Ws;
  const int grid = static_cast<int>(total / 256 + 1 < 148L * 4 ? total / 256 + 1 : 148L * 4);
  SSEG_CUDA(launch_k(label_transform_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream),
                     static_cast<const uint8_t*>(seg_u8), out, valid_hw, N, Hs, Ws, rate));
  count_launch(1);
  return 0;
}
