// Convolution forward / data-gradient as implicit GEMM on tcgen05 tensor cores (sm_100a).
//
//   D[pixel, co] = sum over K-steps (tap t, 64-channel block b) of  A_tb[pixel, 64] . B[co, t, b*64 : b*64+64]
//
//   M = 128 output pixels (a BH x BW spatial box of one image), N = BLOCK_N output channels, K-step = 64 channels.
//   A tile: ONE 4-D TMA box load (64 ch, BW, BH, 1) from the NHWC activation at the tap-shifted coordinate
//           (w0 + dw*?, h0 + dh) -- hardware out-of-bounds zero fill is the convolution's zero padding, and the
//           128B-swizzled box lands in shared memory already in the K-major UMMA canonical layout.
//   B tile: 2-D TMA box (64 k, BLOCK_N rows) from the [cout][taps*cin] bf16 weight matrix.
//   Accumulator: TMEM (128 lanes x BLOCK_N fp32 columns), read back with tcgen05.ld by 4 epilogue warps.
//   Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue.
//   Epilogue fusions: +bias, +addend (residual / gradient accumulation), per-channel sum & sum-of-squares
//   (batch-norm statistics, lib/nn/modules/batchnorm.py:68-70 of the reference) and bf16 / fp32 store.
#include "common.h"
#include "ptx.cuh"

namespace sseg {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // bf16 elements = 128 bytes = one swizzle span
constexpr int kNumThreads = 192;
constexpr int kABytes = kBlockM * kBlockK * 2;

struct IgemmParams {
  CUtensorMap tmA[SSEG_MAX_SRCS];
  CUtensorMap tmB;
  int nsrc;
  int src_blk_end[SSEG_MAX_SRCS];  // cumulative count of 64-channel blocks
  int blocks_per_tap;
  int ntaps;
  int tap_dh[SSEG_MAX_TAPS], tap_dw[SSEG_MAX_TAPS];
  int N, H, W;
  int BH, BW, bw_shift;
  int tiles_h, tiles_w, n_tiles;
  void* out;
  int out_f32, ld_out, n_store, cout;
  const float* bias;
  const __nv_bfloat16* addend;
  int ld_addend;
  float* stat_sum;
  float* stat_sqsum;
};

template <int BLOCK_N, int STAGES>
struct IgemmSmem {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOff = STAGES * kStageBytes;
  static constexpr int kStatOff = kBarOff + 256;  // barriers + tmem ptr live in the first 256 bytes
  static constexpr int kTotal = kStatOff + 2 * BLOCK_N * 4;
  static constexpr int kDynBytes = kTotal + 1024;  // slack for manual 1024B alignment
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kNumThreads) igemm_kernel(const __grid_constant__ IgemmParams p) {
  using L = IgemmSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_sum = reinterpret_cast<float*>(smem + L::kStatOff);
  float* s_sq = s_sum + BLOCK_N;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile decode: N-tile fastest so CTAs sharing an activation tile run together (L2 reuse)
  const int n_tile = blockIdx.x % p.n_tiles;
  int m_tile = blockIdx.x / p.n_tiles;
  const int tw = m_tile % p.tiles_w;
  m_tile /= p.tiles_w;
  const int th = m_tile % p.tiles_h;
  const int img = m_tile / p.tiles_h;
  const int h0 = th * p.BH, w0 = tw * p.BW, n0 = n_tile * BLOCK_N;
  const int num_k_steps = p.ntaps * p.blocks_per_tap;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * BLOCK_N; i += kNumThreads) s_sum[i] = 0.f;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsrc; ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1) tmem_alloc<BLOCK_N>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int t = 0; t < p.ntaps; ++t) {
        const int hh = h0 + p.tap_dh[t], ww = w0 + p.tap_dw[t];
        int src = 0, blk_begin = 0;
        for (int b = 0; b < p.blocks_per_tap; ++b) {
          while (b >= p.src_blk_end[src]) {
            blk_begin = p.src_blk_end[src];
            ++src;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          tma_load_4d(sa, &p.tmA[src], &full_bar[stage], (b - blk_begin) * kBlockK, ww, hh, img);
          tma_load_2d(sa + kABytes, &p.tmB, &full_bar[stage], (t * p.blocks_per_tap + b) * kBlockK, n0);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(/*bf16*/ 1, 0, 0, kBlockM, BLOCK_N);
      int stage = 0, phase = 0;
      for (int ks = 0; ks < num_k_steps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
        const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
          umma_bf16(tmem_base, da, db, idesc, (ks | k) != 0);
        }
        umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
    __syncwarp();
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int quarter = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int row = quarter * 32 + lane;
    const int hh = h0 + (row >> p.bw_shift), ww = w0 + (row & (p.BW - 1));
    const bool valid = (hh < p.H) && (ww < p.W);
    const size_t pix = (static_cast<size_t>(img) * p.H + hh) * p.W + ww;
    const bool do_stats = p.stat_sum != nullptr;

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
      uint32_t raw[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + chunk * 32, raw);
      tmem_ld_wait();
      float v[32];
      const int col0 = n0 + chunk * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
      if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.cout) v[j] += __ldg(p.bias + col0 + j);
      }
      if (p.addend != nullptr && valid) {
        const __nv_bfloat16* ap = p.addend + pix * p.ld_addend + col0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (col0 + g * 8 < p.n_store) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(ap + g * 8));
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(h2[e]);
              v[g * 8 + 2 * e] += f.x;
              v[g * 8 + 2 * e + 1] += f.y;
            }
          }
        }
      }
      if (valid) {
        if (p.out_f32) {
          float* op = reinterpret_cast<float*>(p.out) + pix * p.ld_out + col0;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            if (col0 + g * 4 < p.n_store)
              *reinterpret_cast<float4*>(op + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        } else {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.ld_out + col0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (col0 + g * 8 < p.n_store) {
              uint4 q;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
              for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
              *reinterpret_cast<uint4*>(op + g * 8) = q;
            }
          }
        }
      }
      if (do_stats) {
        // Column sums over the warp's 32 rows by a transposing butterfly: after the loop lane l holds column l.
        float s[32], q[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          s[j] = valid ? v[j] : 0.f;
          q[j] = s[j] * s[j];
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int j = 0; j < off; ++j) {
            const float send_s = upper ? s[j] : s[j + off];
            const float keep_s = upper ? s[j + off] : s[j];
            s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, off);
            const float send_q = upper ? q[j] : q[j + off];
            const float keep_q = upper ? q[j + off] : q[j];
            q[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, off);
          }
        }
        atomicAdd(&s_sum[chunk * 32 + lane], s[0]);
        atomicAdd(&s_sq[chunk * 32 + lane], q[0]);
      }
    }
    if (do_stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the 4 epilogue warps only
      for (int c = threadIdx.x - 64; c < BLOCK_N; c += 128) {
        if (n0 + c < p.cout) {
          atomicAdd(p.stat_sum + n0 + c, s_sum[c]);
          atomicAdd(p.stat_sqsum + n0 + c, s_sq[c]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BLOCK_N>(tmem_base);
  }
}

template <int BLOCK_N, int STAGES>
static int launch(const IgemmParams& p, int grid, cudaStream_t stream) {
  using L = IgemmSmem<BLOCK_N, STAGES>;
  static bool configured[64] = {};
  int dev = 0;
  SSEG_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !configured[dev]) {
    SSEG_CUDA(cudaFuncSetAttribute(igemm_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   L::kDynBytes));
    configured[dev] = true;
  }
  igemm_kernel<BLOCK_N, STAGES><<<grid, kNumThreads, L::kDynBytes, stream>>>(p);
  count_launch(1);
  return check_cuda(cudaGetLastError(), "igemm_kernel launch");
}

}  // namespace sseg

using namespace sseg;

extern "C" int sseg_conv_igemm(const sseg_act_t* srcs, int nsrc, const void* w_bf16, int cout, int ntaps,
                               const int* tap_dh, const int* tap_dw, void* out, int out_f32, int ld_out, int n_store,
                               const float* bias, const void* addend, int ld_addend, float* stat_sum,
                               float* stat_sqsum, sseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SSEG_REQUIRE(nsrc >= 1 && nsrc <= SSEG_MAX_SRCS, "sseg_conv_igemm: nsrc=%d out of range", nsrc);
  SSEG_REQUIRE(ntaps >= 1 && ntaps <= SSEG_MAX_TAPS, "sseg_conv_igemm: ntaps=%d out of range", ntaps);
  SSEG_REQUIRE(cout >= 1 && n_store >= cout && n_store % 8 == 0 && n_store <= ld_out,
               "sseg_conv_igemm: need cout <= n_store (mult of 8) <= ld_out, got %d %d %d", cout, n_store, ld_out);
  SSEG_REQUIRE((stat_sum == nullptr) == (stat_sqsum == nullptr), "sseg_conv_igemm: stat_sum/stat_sqsum must pair");
  const int N = srcs[0].n, H = srcs[0].h, W = srcs[0].w;
  SSEG_REQUIRE(N >= 1 && H >= 1 && W >= 1, "sseg_conv_igemm: empty activation");

  IgemmParams p;
  memset(&p, 0, sizeof(p));
  // M tile = BH x BW pixels of one image. 1x1 convs see the batch as one long row of pixels.
  int vn = N, vh = H, vw = W;
  const bool pointwise = (ntaps == 1 && tap_dh[0] == 0 && tap_dw[0] == 0);
  if (pointwise) {
    bool dense = true;  // all sources must be plain [pixels][ld] views (always true for NHWC)
    if (dense) vn = 1, vh = 1, vw = N * H * W;
  }
  int BW = 128;
  while (BW > 8 && BW / 2 >= vw) BW /= 2;  // smallest power of two >= W, capped to [8,128]
  if (vw >= 128) BW = 128;
  int BH = kBlockM / BW;
  p.BH = BH, p.BW = BW;
  p.bw_shift = 0;
  while ((1 << p.bw_shift) < BW) ++p.bw_shift;
  p.N = vn, p.H = vh, p.W = vw;
  p.tiles_h = ceil_div(vh, BH), p.tiles_w = ceil_div(vw, BW);

  int cin_total = 0;
  for (int s = 0; s < nsrc; ++s) {
    SSEG_REQUIRE(srcs[s].n == N && srcs[s].h == H && srcs[s].w == W, "sseg_conv_igemm: source %d shape mismatch", s);
    SSEG_REQUIRE(srcs[s].c % kBlockK == 0 && srcs[s].c > 0, "sseg_conv_igemm: source %d channels %d not a multiple of 64",
                 s, srcs[s].c);
    SSEG_REQUIRE(srcs[s].ld % 8 == 0 && srcs[s].ld >= srcs[s].c, "sseg_conv_igemm: source %d ld %d invalid", s,
                 srcs[s].ld);
    int rc = get_tmap_act(&p.tmA[s], srcs[s].ptr, 2, vn, vh, vw, srcs[s].c, srcs[s].ld, kBlockK, BW, BH);
    if (rc) return rc;
    cin_total += srcs[s].c;
    p.src_blk_end[s] = cin_total / kBlockK;
  }
  p.nsrc = nsrc;
  p.blocks_per_tap = cin_total / kBlockK;
  p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) p.tap_dh[t] = tap_dh[t], p.tap_dw[t] = tap_dw[t];

  const int block_n = cout <= 64 ? 64 : 128;
  p.n_tiles = ceil_div(n_store, block_n);
  {
    const long K = static_cast<long>(ntaps) * cin_total;
    int rc = get_tmap_2d(&p.tmB, w_bf16, 2, cout, K, K, kBlockK, block_n);
    if (rc) return rc;
  }
  p.out = out, p.out_f32 = out_f32, p.ld_out = ld_out, p.n_store = n_store, p.cout = cout;
  p.bias = bias;
  p.addend = static_cast<const __nv_bfloat16*>(addend);
  p.ld_addend = ld_addend;
  p.stat_sum = stat_sum, p.stat_sqsum = stat_sqsum;
  SSEG_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (ld_out * (out_f32 ? 4 : 2)) % 16 == 0,
               "sseg_conv_igemm: output not 16B aligned");
  SSEG_REQUIRE(addend == nullptr || ((reinterpret_cast<uintptr_t>(addend) & 15) == 0 && ld_addend % 8 == 0),
               "sseg_conv_igemm: addend not 16B aligned");

  const int grid = vn * p.tiles_h * p.tiles_w * p.n_tiles;
  if (block_n == 64) return launch<64, 4>(p, grid, stream);
  return launch<128, 3>(p, grid, stream);
}
