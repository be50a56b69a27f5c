// Convolution forward / data-gradient as implicit GEMM on tcgen05 tensor cores (sm_100a).
//
//   D[pixel, co] = sum over K-steps (tap t, 64-channel block b) of  A_tb[pixel, 64] . B[co, t, b*64 : b*64+64]
//   (source channel counts need not be multiples of 64: a partial block is zero-filled by TMA)
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
#include <stdlib.h>

#define SSEG_SIM_PARALLEL_CTAS 6  // CPU simulator only: no static __shared__ variables in this file
#include "common.h"
#include "ptx.cuh"

namespace sseg {

// Cycle stamps of the plain kernel's phases (development builds only: `make trace` -> libsseg_b200_trace.so, read back with
// sseg_debug_read_trace; tools/trace_igemm.py). One row of 16 clock64() values per CTA.
#ifdef SSEG_TRACE
__device__ long long g_trace[4096 * 16];
#define SSEG_STAMP(k)                                                        \
  do {                                                                       \
    if (blockIdx.x < 4096) g_trace[blockIdx.x * 16 + (k)] = clock64();       \
  } while (0)
#else
#define SSEG_STAMP(k) \
  do {                \
  } while (0)
#endif

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // bf16 elements = 128 bytes = one swizzle span
constexpr int kNumThreads = 192;
// the plain conv / weight-gradient kernels carry a SECOND producer warp: one thread issuing both operands' TMA boxes
// back to back is the bottleneck of the main loop (measured with cycle stamps, profiles/r2_summary.md: ~230 cycles per
// box issue = 470-550 cycles per k-step against 256 cycles of MMA), so the A and the B boxes are issued by different warps
constexpr int kNumThreads2 = 224;
constexpr int kABytes = kBlockM * kBlockK * 2;

struct IgemmParams {
  CUtensorMap tmA[SSEG_MAX_SRCS];
  CUtensorMap tmB;
  int nsrc;
  int src_blk_end[SSEG_MAX_SRCS];  // cumulative count of 64-channel blocks (a source's last block may be partial)
  int src_choff[SSEG_MAX_SRCS];    // channel offset of each source inside the virtual concat (= its weight K offset)
  int blocks_per_tap;
  int ntaps;
  int tap_dh[SSEG_MAX_TAPS], tap_dw[SSEG_MAX_TAPS], tap_src[SSEG_MAX_TAPS], tap_koff[SSEG_MAX_TAPS];
  int num_k_steps;
  int N, H, W;
  int BH, BW, bw_shift;
  int tiles_h, tiles_w, n_tiles;
  void* out;
  int out_f32, ld_out, n_store, cout;
  long out_row_stride, out_img_stride;
  const float* bias;
  const __nv_bfloat16* addend;
  int ld_addend;
  long add_row_stride, add_img_stride;
  float* stat_sum;
  float* stat_sqsum;
  // fused inference epilogue (BatchNorm with running statistics is a per-channel affine): v = v*ep_scale + ep_shift,
  // ReLU before ((ep_relu & 3) == 2) or after (== 1) the addend (= shortcut / top-down tensor) is added; bit 2: clip at 6
  const float* ep_scale;
  const float* ep_shift;
  int ep_relu;
  // fused BN-backward reduction of the layer that PRODUCED the tensor whose gradient this launch writes (dgrad):
  //   g' = out * [bw_y * bw_fscale + bw_fshift > 0] ;  bw_s1[c] += sum g' ;  bw_s2[c] += sum g' * bw_y
  const __nv_bfloat16* bw_y;
  int bw_ld;
  long bw_row_stride, bw_img_stride;
  const float* bw_fscale;
  const float* bw_fshift;
  float* bw_s1;
  float* bw_s2;
  // ... or, for a producer with a shortcut (a = relu(bn(y) + shortcut): the mask is not a function of y alone), the ReLU
  // mask comes from the producer's saved OUTPUT a (same geometry as y): g' = out * [a > 0]   (bw_fscale / bw_fshift unused)
  const __nv_bfloat16* bw_a;
  int bw_a_ld;
  long bw_a_row_stride, bw_a_img_stride;
  // BatchNorm finalize by the LAST CTA of the launch (train-mode, single GPU; sseg_conv_igemm_bnfin): once every CTA has
  // added its tile's statistics, the CTA that takes the last ticket turns (sum, sum of squares) into mean / inv_std /
  // scale / shift and updates the running statistics - the separate bn_finalize launch of every layer disappears
  unsigned int* fin_counter;
  const float* fin_gamma;
  const float* fin_beta;
  float fin_eps, fin_momentum, fin_count;
  float *fin_mean, *fin_invstd, *fin_scale, *fin_shift, *fin_rmean, *fin_rvar;
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

// Epilogue of one 128 x BLOCK_N accumulator tile (the 4 epilogue warps = threads 64..191 of the CTA): TMEM -> registers ->
// (+bias, affine, addend, ReLU) -> bf16 tile staged in the idle pipeline buffers -> per-channel statistics / fused BN-backward
// sums as column sums of the staged tile -> coalesced stores. Shared by the one-CTA kernel and the CTA-pair kernel (there
// each CTA drains the 128 accumulator rows that live in its own tensor memory).
template <int BLOCK_N>
__device__ __forceinline__ void igemm_tile_epilogue(const IgemmParams& p, uint8_t* smem, uint32_t tmem_base,
                                                    uint64_t* tmem_full_bar, int warp, int lane, int h0, int w0, int n0,
                                                    int img) {
    const int quarter = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int row = quarter * 32 + lane;
    const int hh = h0 + (row >> p.bw_shift), ww = w0 + (row & (p.BW - 1));
    const bool valid = (hh < p.H) && (ww < p.W);
    const size_t out_off = img * p.out_img_stride + hh * p.out_row_stride + static_cast<size_t>(ww) * p.ld_out;
    const size_t add_off = img * p.add_img_stride + hh * p.add_row_stride + static_cast<size_t>(ww) * p.ld_addend;
    const bool do_stats = p.stat_sum != nullptr;

    mbar_wait(tmem_full_bar, 0);
    if (threadIdx.x == 64) SSEG_STAMP(7);
    tc_fence_after();
    // bf16 outputs are staged through shared memory (the pipeline stages are idle once the accumulator is complete):
    // rows are then stored with full 16-byte-per-lane coalescing and the per-channel statistics are column sums of the
    // staged (bf16-rounded = as stored) tile. fp32 outputs (classifier logits) are written straight from registers.
    constexpr int kPitch = BLOCK_N * 2 + 16;  // +16 B: consecutive rows start in different 16-byte bank groups
    uint8_t* stg = smem;
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
      if (p.ep_scale != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.cout) v[j] = fmaf(v[j], __ldg(p.ep_scale + col0 + j), __ldg(p.ep_shift + col0 + j));
      }
      if ((p.ep_relu & 3) == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (p.addend != nullptr && valid) {
        const __nv_bfloat16* ap = p.addend + add_off + col0;
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
      if ((p.ep_relu & 3) == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (p.ep_relu & 4) {  // ReLU6 (MobileNetV2): the ReLU above, clipped at 6
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fminf(v[j], 6.f);
      }
      if (p.out_f32) {
        if (valid) {
          float* op = reinterpret_cast<float*>(p.out) + out_off + col0;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            if (col0 + g * 4 < p.n_store)
              *reinterpret_cast<float4*>(op + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        }
      } else {
        uint8_t* sp = stg + row * kPitch + chunk * 64;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 q = make_uint4(0u, 0u, 0u, 0u);  // rows outside the image contribute zeros to the statistics
          if (valid) {
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
            for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
          }
          *reinterpret_cast<uint4*>(sp + g * 16) = q;
        }
      }
    }
    if (threadIdx.x == 64) SSEG_STAMP(8);
    if (!p.out_f32) {
      bar_sync_epilogue();  // the 4 epilogue warps only: the staged tile is complete
      const int t = threadIdx.x - 64;
      if (threadIdx.x == 64) SSEG_STAMP(9);
      // coalesced store first: BLOCK_N/8 lanes cover one row (16 B each), several rows per pass. The stores are
      // fire-and-forget, so the column sums below (shared-memory reads only) run while they drain.
      {
      constexpr int kLanesPerRow = BLOCK_N / 8, kRowsPerPass = 128 / kLanesPerRow;
      const int seg = t % kLanesPerRow, r0 = t / kLanesPerRow;
            if (n0 + seg * 8 < p.n_store) {
#pragma unroll 4
        for (int pass = 0; pass < 128 / kRowsPerPass; ++pass) {
          const int r = pass * kRowsPerPass + r0;
          const int rh = h0 + (r >> p.bw_shift), rw = w0 + (r & (p.BW - 1));
          if (rh < p.H && rw < p.W) {
            const uint4 q = *reinterpret_cast<const uint4*>(stg + r * kPitch + seg * 16);
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + img * p.out_img_stride + rh * p.out_row_stride +
                                static_cast<size_t>(rw) * p.ld_out + n0 + seg * 8;
            *reinterpret_cast<uint4*>(op) = q;
          }
        }
      }
      }
      if (threadIdx.x == 64) SSEG_STAMP(11);
      if (do_stats) {
        // thread = one pair of adjacent columns x one slab of rows; fp32 sums of the bf16 values as stored
        constexpr int kPairs = BLOCK_N / 2, kSlabs = 128 / kPairs, kRowsPerSlab = 128 / kSlabs;
        const int cp = t % kPairs, slab = t / kPairs;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        const uint8_t* base = stg + (slab * kRowsPerSlab) * kPitch + cp * 4;
#pragma unroll 8
        for (int r = 0; r < kRowsPerSlab; ++r) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(base + r * kPitch));
          s0 += f.x, s1 += f.y;
          q0 = fmaf(f.x, f.x, q0), q1 = fmaf(f.y, f.y, q1);
        }
        const int col = n0 + cp * 2;
        if (col < p.cout) atomicAdd(p.stat_sum + col, s0), atomicAdd(p.stat_sqsum + col, q0);
        if (col + 1 < p.cout) atomicAdd(p.stat_sum + col + 1, s1), atomicAdd(p.stat_sqsum + col + 1, q1);
      }
      if (threadIdx.x == 64) SSEG_STAMP(10);
      if (p.bw_s1 != nullptr) {
        // BN-backward partial sums of the producer layer. A tile of its saved tensors (same geometry as this output tile) is
        // copied into shared memory with fully coalesced 16-byte loads (batches of loads in flight per thread), then the
        // per-channel sums run from shared memory next to the staged gradient tile.
        uint8_t* ytile = stg + 128 * kPitch;
        auto copy_tile = [&](const __nv_bfloat16* src, int ld, long row_stride, long img_stride) {
          constexpr int kLanesPerRow = BLOCK_N / 8, kRowsPerPass = 128 / kLanesPerRow, kPasses = 128 / kRowsPerPass;
          constexpr int kBatch = kPasses < 16 ? kPasses : 16;   // loads in flight per thread (registers: 4 per load)
          const int seg = t % kLanesPerRow, r0 = t / kLanesPerRow;
#pragma unroll 1
          for (int pb = 0; pb < kPasses; pb += kBatch) {
            uint4 q[kBatch];
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
              const int r = (pb + j) * kRowsPerPass + r0;
              const int rh = h0 + (r >> p.bw_shift), rw = w0 + (r & (p.BW - 1));
              q[j] = make_uint4(0u, 0u, 0u, 0u);
              if (rh < p.H && rw < p.W && n0 + seg * 8 < p.cout)
                q[j] = __ldg(reinterpret_cast<const uint4*>(src + img * img_stride + rh * row_stride +
                                                            static_cast<size_t>(rw) * ld + n0 + seg * 8));
            }
#pragma unroll
            for (int j = 0; j < kBatch; ++j)
              *reinterpret_cast<uint4*>(ytile + ((pb + j) * kRowsPerPass + r0) * kPitch + seg * 16) = q[j];
          }
        };
        constexpr int kPairs = BLOCK_N / 2, kSlabs = 128 / kPairs, kRowsPerSlab = 128 / kSlabs;
        const int cp = t % kPairs, slab = t / kPairs;
        const int col = n0 + cp * 2;
        const uint8_t* gbase = stg + (slab * kRowsPerSlab) * kPitch + cp * 4;
        const uint8_t* ybase = ytile + (slab * kRowsPerSlab) * kPitch + cp * 4;
        if (p.bw_a != nullptr) {
          // mask from the producer's saved output a (shortcut layers): pass 1 over the a tile builds one mask bit per
          // (row, column) of this thread's slab in registers and the sums of g' = g * [a > 0]; pass 2 brings the y tile
          // into the same buffer for the sums of g' * y
          copy_tile(p.bw_a, p.bw_a_ld, p.bw_a_row_stride, p.bw_a_img_stride);
          bar_sync_epilogue();
          uint32_t m0[kRowsPerSlab / 32], m1[kRowsPerSlab / 32];
          float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
          for (int w = 0; w < kRowsPerSlab / 32; ++w) {
            uint32_t bits0 = 0u, bits1 = 0u;
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
              const int r = w * 32 + rr;
              const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(gbase + r * kPitch));
              const float2 av = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ybase + r * kPitch));
              const bool on0 = av.x > 0.f, on1 = av.y > 0.f;
              bits0 |= (on0 ? 1u : 0u) << rr, bits1 |= (on1 ? 1u : 0u) << rr;
              a0 += on0 ? g.x : 0.f, a1 += on1 ? g.y : 0.f;
            }
            m0[w] = bits0, m1[w] = bits1;
          }
          bar_sync_epilogue();  // every thread has read the a tile: the buffer may be overwritten
          copy_tile(p.bw_y, p.bw_ld, p.bw_row_stride, p.bw_img_stride);
          bar_sync_epilogue();
#pragma unroll
          for (int w = 0; w < kRowsPerSlab / 32; ++w) {
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
              const int r = w * 32 + rr;
              const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(gbase + r * kPitch));
              const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ybase + r * kPitch));
              b0 = fmaf(((m0[w] >> rr) & 1u) ? g.x : 0.f, y.x, b0);
              b1 = fmaf(((m1[w] >> rr) & 1u) ? g.y : 0.f, y.y, b1);
            }
          }
          if (col < p.cout) {
            atomicAdd(p.bw_s1 + col, a0), atomicAdd(p.bw_s2 + col, b0);
            if (col + 1 < p.cout) atomicAdd(p.bw_s1 + col + 1, a1), atomicAdd(p.bw_s2 + col + 1, b1);
          }
        } else {
          copy_tile(p.bw_y, p.bw_ld, p.bw_row_stride, p.bw_img_stride);
          bar_sync_epilogue();
          if (col < p.cout) {
            const float fs0 = p.bw_fscale[col], fb0 = p.bw_fshift[col];
            const float fs1 = col + 1 < p.cout ? p.bw_fscale[col + 1] : 0.f, fb1 = col + 1 < p.cout ? p.bw_fshift[col + 1] : -1.f;
            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 8
            for (int r = 0; r < kRowsPerSlab; ++r) {
              // rows outside the image hold zeros in the staged gradient tile: they contribute nothing
              const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(gbase + r * kPitch));
              const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ybase + r * kPitch));
              const float g0 = fmaf(y.x, fs0, fb0) > 0.f ? g.x : 0.f;
              const float g1 = fmaf(y.y, fs1, fb1) > 0.f ? g.y : 0.f;
              a0 += g0, a1 += g1;
              b0 = fmaf(g0, y.x, b0), b1 = fmaf(g1, y.y, b1);
            }
            atomicAdd(p.bw_s1 + col, a0), atomicAdd(p.bw_s2 + col, b0);
            if (col + 1 < p.cout) atomicAdd(p.bw_s1 + col + 1, a1), atomicAdd(p.bw_s2 + col + 1, b1);
          }
        }
      }
    }
}

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kNumThreads2) igemm_kernel(const __grid_constant__ IgemmParams p) {
  using L = IgemmSmem<BLOCK_N, STAGES>;
  static_assert(L::kDynBytes <= 232448, "shared memory budget");
  static_assert(2 * 128 * (BLOCK_N * 2 + 16) <= STAGES * L::kStageBytes, "the epilogue stages two tiles in the pipeline buffers");
  SSEG_DYN_SMEM(smem_raw);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  uint32_t* tmem_ptr_smem2 = tmem_ptr_smem + 1;  // "this CTA took the last ticket" flag of the fused BN finalize

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) SSEG_STAMP(0);

  // tile decode: N-tile fastest so CTAs sharing an activation tile run together (L2 reuse)
  const int n_tile = blockIdx.x % p.n_tiles;
  int m_tile = blockIdx.x / p.n_tiles;
  const int tw = m_tile % p.tiles_w;
  m_tile /= p.tiles_w;
  const int th = m_tile % p.tiles_h;
  const int img = m_tile / p.tiles_h;
  const int h0 = th * p.BH, w0 = tw * p.BW, n0 = n_tile * BLOCK_N;
  const int num_k_steps = p.num_k_steps;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);  // one arrive.expect_tx per producer warp (A boxes, B boxes)
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsrc; ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1) tmem_alloc<BLOCK_N>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) SSEG_STAMP(1);
  pdl_sync();  // everything above overlapped the previous kernel's tail; global memory is touched only below
  if (threadIdx.x == 0) SSEG_STAMP(2);

  if (warp == 0 || warp == 6) {
    // ===================== TMA producers: warp 0 the activation boxes (A), warp 6 the weight boxes (B) =====================
    // The WHOLE warp walks the loop and one elected lane issues: with warp-uniform control flow the coordinates and
    // barrier addresses stay in uniform registers; issuing from a divergent `if (lane == 0)` region made the compiler wrap
    // every uniform-datapath instruction in an election loop (measured: ~400 cycles per box issue).
    {
      const bool is_a = warp == 0;
      int stage = 0, phase = 0;
      for (int t = 0; t < p.ntaps; ++t) {
        const int hh = h0 + p.tap_dh[t], ww = w0 + p.tap_dw[t];
        const int fixed_src = p.tap_src[t];
        const int nblk = fixed_src >= 0 ? p.src_blk_end[0] : p.blocks_per_tap;
        int src = fixed_src >= 0 ? fixed_src : 0, blk_begin = 0;
        for (int b = 0; b < nblk; ++b) {
          if (fixed_src < 0) {
            while (b >= p.src_blk_end[src]) {
              blk_begin = p.src_blk_end[src];
              ++src;
            }
          }
          mbar_wait_warp(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          // a partial last block of a source (channels % 64 != 0) reads zeros beyond the source's channels (TMA
          // out-of-bounds fill), so whatever weight columns sit under them contribute nothing
          const int kcol = p.tap_koff[t] + (fixed_src >= 0 ? 0 : p.src_choff[src]) + (b - blk_begin) * kBlockK;
          if (elect_one()) {
            if (is_a) {
              mbar_expect_tx(&full_bar[stage], kABytes);
              tma_load_4d(sa, &p.tmA[src], &full_bar[stage], (b - blk_begin) * kBlockK, ww, hh, img);
              if (t == 0 && b == 0) SSEG_STAMP(3);
            } else {
              mbar_expect_tx(&full_bar[stage], L::kBBytes);
              tma_load_2d(sa + kABytes, &p.tmB, &full_bar[stage], kcol, n0);
            }
          }
          __syncwarp();
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
      }
      if (is_a && lane == 0) SSEG_STAMP(4);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the whole warp walks the loop, one elected lane issues =====================
    // (same reason as above; the shared-memory descriptors are stepped by adding to their low word instead of being
    // re-encoded for every instruction: the issue loop, not the tensor core, was the limit - ~140 cycles per MMA
    // instruction against 64 needed)
    constexpr uint32_t idesc = make_idesc(/*bf16*/ 1, 0, 0, kBlockM, BLOCK_N);
    constexpr uint32_t dhi = smem_desc_hi_sw128(1024);
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem), 16);
    uint32_t d_tmem = tmem_base;
#ifndef __CUSIM__
    asm volatile("" : "+r"(d_tmem));  // a register of this branch's own (the compiler kept the kernel-wide value on the stack)
#endif
    int stage = 0, phase = 0;
    for (int ks = 0; ks < num_k_steps; ++ks) {
      mbar_wait_warp(&full_bar[stage], phase);
      if (ks == 0 && lane == 0) SSEG_STAMP(5);
      tc_fence_after();
      const uint32_t a_lo = a_lo0 + stage * (L::kStageBytes >> 4);
      const uint32_t b_lo = a_lo + (kABytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)
          umma_bf16(d_tmem, smem_desc_join(a_lo + 2 * k, dhi), smem_desc_join(b_lo + 2 * k, dhi), idesc, (ks | k) != 0);
        umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
      }
      __syncwarp();
      if (++stage == STAGES) stage = 0, phase ^= 1;
    }
    if (elect_one()) {
      umma_commit(tmem_full_bar);  // accumulator complete
      SSEG_STAMP(6);
    }
    __syncwarp();
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    igemm_tile_epilogue<BLOCK_N>(p, smem, tmem_base, tmem_full_bar, warp, lane, h0, w0, n0, img);
  }

  if (threadIdx.x == 64) SSEG_STAMP(12);
  if (p.fin_counter != nullptr && warp >= 2 && warp < 6) {
    // every epilogue thread's statistics atomics are ordered before its arrival at the barrier; one thread then takes the
    // launch-wide ticket (the pattern of the threadFenceReduction sample); the flag travels through the tmem-pointer word
    __threadfence();
    bar_sync_epilogue();
    if (threadIdx.x == 64) {
      const unsigned int ticket = atomicAdd(p.fin_counter, 1u);
      *tmem_ptr_smem2 = (ticket == gridDim.x - 1) ? 1u : 0u;
    }
    bar_sync_epilogue();
    if (*tmem_ptr_smem2 != 0u) {
      __threadfence();
      const float cnt = p.fin_count;
      for (int c = threadIdx.x - 64; c < p.cout; c += 128) {
        const float s = __ldcg(p.stat_sum + c), q = __ldcg(p.stat_sqsum + c);
        const float mean = s / cnt;
        const float sumvar = q - s * mean;
        const float bias_var = sumvar / cnt, unbias_var = sumvar / (cnt - 1.f);
        const float inv_std = rsqrtf(fmaxf(bias_var, 0.f) + p.fin_eps);
        if (p.fin_rmean != nullptr) {
          p.fin_rmean[c] = (1.f - p.fin_momentum) * p.fin_rmean[c] + p.fin_momentum * mean;
          p.fin_rvar[c] = (1.f - p.fin_momentum) * p.fin_rvar[c] + p.fin_momentum * unbias_var;
        }
        const float g = p.fin_gamma ? p.fin_gamma[c] : 1.f, b = p.fin_beta ? p.fin_beta[c] : 0.f;
        p.fin_mean[c] = mean, p.fin_invstd[c] = inv_std;
        p.fin_scale[c] = g * inv_std;
        p.fin_shift[c] = b - mean * g * inv_std;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) SSEG_STAMP(13);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BLOCK_N>(tmem_base);
  }
  if (threadIdx.x == 32) SSEG_STAMP(14);
}

// Persistent variant: one CTA per SM walks tiles (tile = blockIdx.x + i * gridDim.x). Two TMEM accumulator buffers let the
// MMA warp start tile i+1 while the epilogue warps still drain / reduce / store tile i, and the TMA producer runs ahead
// across tile boundaries, so the pipeline-fill and epilogue latencies are paid once per CTA instead of once per tile.
// The main loop is bound by TMA round trips (~1.5 us per box pair: throughput = bytes in flight / latency), so shared
// memory goes to pipeline stages (192 KB) and the epilogue works on 32-column slices of the accumulator through a small
// staging buffer (20 KB) instead of staging the whole output tile.
constexpr int kEpiN = 32;                      // accumulator columns per epilogue slice
constexpr int kEpiPitch = kEpiN * 2 + 16;      // bytes per staged row (+16: rows start in different 16-byte bank groups)
template <int BLOCK_N, int STAGES>
struct IgemmPersistSmem {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStgOff = STAGES * kStageBytes;                    // staged output slice + staged y slice
  static constexpr int kBarOff = kStgOff + ((2 * 128 * kEpiPitch + 1023) / 1024) * 1024;
  static constexpr int kTotal = kBarOff + 256;
  static constexpr int kDynBytes = kTotal + 1024;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kNumThreads, 1) igemm_persistent_kernel(const __grid_constant__ IgemmParams p,
                                                                          const int num_tiles) {
  using L = IgemmPersistSmem<BLOCK_N, STAGES>;
  static_assert(L::kDynBytes <= 232448, "shared memory budget");
  SSEG_DYN_SMEM(smem_raw);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]: accumulator buffer b holds a finished tile
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]: the epilogue has drained accumulator buffer b
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_k_steps = p.num_k_steps;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsrc; ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1) tmem_alloc<2 * BLOCK_N>(tmem_ptr_smem);  // two accumulator buffers
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();  // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      // tile decode: N-tile fastest so CTAs sharing an activation tile run together (L2 reuse)
      const int n_tile = tile % p.n_tiles;
      int m_tile = tile / p.n_tiles;
      const int tw = m_tile % p.tiles_w;
      m_tile /= p.tiles_w;
      const int th = m_tile % p.tiles_h;
      const int img = m_tile / p.tiles_h;
      const int h0 = th * p.BH, w0 = tw * p.BW, n0 = n_tile * BLOCK_N;
      (void)n0;
      for (int t = 0; t < p.ntaps; ++t) {
        const int hh = h0 + p.tap_dh[t], ww = w0 + p.tap_dw[t];
        const int fixed_src = p.tap_src[t];
        const int nblk = fixed_src >= 0 ? p.src_blk_end[0] : p.blocks_per_tap;
        int src = fixed_src >= 0 ? fixed_src : 0, blk_begin = 0;
        for (int b = 0; b < nblk; ++b) {
          if (fixed_src < 0) {
            while (b >= p.src_blk_end[src]) {
              blk_begin = p.src_blk_end[src];
              ++src;
            }
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          tma_load_4d(sa, &p.tmA[src], &full_bar[stage], (b - blk_begin) * kBlockK, ww, hh, img);
          // a partial last block of a source (channels % 64 != 0) reads zeros beyond the source's channels (TMA
          // out-of-bounds fill), so whatever weight columns sit under them contribute nothing
          const int kcol = p.tap_koff[t] + (fixed_src >= 0 ? 0 : p.src_choff[src]) + (b - blk_begin) * kBlockK;
          tma_load_2d(sa + kABytes, &p.tmB, &full_bar[stage], kcol, n0);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
      }
      }  // tile loop
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(/*bf16*/ 1, 0, 0, kBlockM, BLOCK_N);
      int stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1, use = it >> 1;
      mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);  // the epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
      for (int ks = 0; ks < num_k_steps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
        const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
          umma_bf16(d_tmem, da, db, idesc, (ks | k) != 0);
        }
        umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
      umma_commit(&tmem_full_bar[buf]);  // accumulator of this tile complete
      }  // tile loop
    }
    __syncwarp();
  } else {
    // ===================== epilogue: TMEM -> registers -> (staging slice) -> global =====================
    const int quarter = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int row = quarter * 32 + lane;
    const int t = threadIdx.x - 64;
    const bool do_stats = p.stat_sum != nullptr;
    uint8_t* stg = smem + L::kStgOff;  // dedicated staging: the pipeline stages already carry the next tile
    uint8_t* ytile = stg + 128 * kEpiPitch;
    // thread mappings of the slice passes: column sums (thread = one pair of adjacent columns x one slab of rows) and
    // row copies (kLanesPerRow lanes cover the 64 bytes of one row, several rows per pass)
    constexpr int kPairs = kEpiN / 2, kSlabs = 128 / kPairs, kRowsPerSlab = 128 / kSlabs;
    constexpr int kLanesPerRow = kEpiN / 8, kRowsPerPass = 128 / kLanesPerRow, kPasses = 128 / kRowsPerPass;
    const int cp = t % kPairs, slab = t / kPairs;
    const int seg = t % kLanesPerRow, r0 = t / kLanesPerRow;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    // tile decode: N-tile fastest so CTAs sharing an activation tile run together (L2 reuse)
    const int n_tile = tile % p.n_tiles;
    int m_tile = tile / p.n_tiles;
    const int tw = m_tile % p.tiles_w;
    m_tile /= p.tiles_w;
    const int th = m_tile % p.tiles_h;
    const int img = m_tile / p.tiles_h;
    const int h0 = th * p.BH, w0 = tw * p.BW, n0 = n_tile * BLOCK_N;
    const int buf = it & 1, use = it >> 1;
    const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
    const int hh = h0 + (row >> p.bw_shift), ww = w0 + (row & (p.BW - 1));
    const bool valid = (hh < p.H) && (ww < p.W);
    const size_t out_off = img * p.out_img_stride + hh * p.out_row_stride + static_cast<size_t>(ww) * p.ld_out;
    const size_t add_off = img * p.add_img_stride + hh * p.add_row_stride + static_cast<size_t>(ww) * p.ld_addend;
    (void)out_off;

    uint4 yq[kPasses];
    auto load_y = [&](int c0) {
#pragma unroll
      for (int pass = 0; pass < kPasses; ++pass) {
        const int r = pass * kRowsPerPass + r0;
        const int rh = h0 + (r >> p.bw_shift), rw = w0 + (r & (p.BW - 1));
        yq[pass] = make_uint4(0u, 0u, 0u, 0u);
        if (rh < p.H && rw < p.W && c0 + seg * 8 < p.cout)
          yq[pass] = __ldg(reinterpret_cast<const uint4*>(p.bw_y + img * p.bw_img_stride + rh * p.bw_row_stride +
                                                          static_cast<size_t>(rw) * p.bw_ld + c0 + seg * 8));
      }
    };
    if (p.bw_s1 != nullptr) load_y(n0);  // independent of the accumulator: in flight while the main loop finishes

    mbar_wait(&tmem_full_bar[buf], use & 1);
    tc_fence_after();
#pragma unroll 1
    for (int chunk = 0; chunk < BLOCK_N / kEpiN; ++chunk) {
      uint32_t raw[32];
      tmem_ld_32x32(d_tmem + (static_cast<uint32_t>(quarter * 32) << 16) + chunk * 32, raw);
      tmem_ld_wait();
      if (chunk == BLOCK_N / kEpiN - 1) {
        // every tcgen05.ld of this tile has completed in this thread; once all four warps are here (the barrier below)
        // the MMA warp may refill this accumulator buffer
        tc_fence_before();
      }
      float v[32];
      const int col0 = n0 + chunk * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
      if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.cout) v[j] += __ldg(p.bias + col0 + j);
      }
      if (p.ep_scale != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.cout) v[j] = fmaf(v[j], __ldg(p.ep_scale + col0 + j), __ldg(p.ep_shift + col0 + j));
      }
      if ((p.ep_relu & 3) == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (p.addend != nullptr && valid) {
        const __nv_bfloat16* ap = p.addend + add_off + col0;
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
      if ((p.ep_relu & 3) == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (p.ep_relu & 4) {  // ReLU6 (MobileNetV2): the ReLU above, clipped at 6
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fminf(v[j], 6.f);
      }
      if (p.out_f32) {
        // fp32 outputs (classifier logits) are written straight from registers
        if (valid) {
          float* op = reinterpret_cast<float*>(p.out) + out_off + col0;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            if (col0 + g * 4 < p.n_store)
              *reinterpret_cast<float4*>(op + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        }
        if (chunk == BLOCK_N / kEpiN - 1) {
          bar_sync_epilogue();
          if (threadIdx.x == 64) mbar_arrive(&tmem_empty_bar[buf]);
        }
        continue;
      }
      // bf16 outputs: the 128 x 32 slice is staged in shared memory, so that rows are stored with 16 bytes per lane over
      // whole 64-byte runs and the per-channel statistics are column sums of the staged (bf16-rounded = as stored) values
      {
        uint8_t* sp = stg + row * kEpiPitch;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 q = make_uint4(0u, 0u, 0u, 0u);  // rows outside the image contribute zeros to the statistics
          if (valid) {
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
            for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
          }
          *reinterpret_cast<uint4*>(sp + g * 16) = q;
        }
      }
      if (p.bw_s1 != nullptr) {
        // BN-backward partial sums of the producer layer: the matching slice of its saved conv output y goes to shared
        // memory next to the staged gradient slice; the loads of the NEXT slice are issued right away, so their latency
        // is covered by this slice's reductions and stores
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass)
          *reinterpret_cast<uint4*>(ytile + (pass * kRowsPerPass + r0) * kEpiPitch + seg * 16) = yq[pass];
        if (chunk + 1 < BLOCK_N / kEpiN) load_y(col0 + kEpiN);
      }
      bar_sync_epilogue();  // the staged slice (and y slice) is complete
      if (chunk == BLOCK_N / kEpiN - 1 && threadIdx.x == 64) mbar_arrive(&tmem_empty_bar[buf]);
      const int col = col0 + cp * 2;
      if (do_stats) {
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        const uint8_t* base = stg + (slab * kRowsPerSlab) * kEpiPitch + cp * 4;
#pragma unroll
        for (int r = 0; r < kRowsPerSlab; ++r) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(base + r * kEpiPitch));
          s0 += f.x, s1 += f.y;
          q0 = fmaf(f.x, f.x, q0), q1 = fmaf(f.y, f.y, q1);
        }
        if (col < p.cout) atomicAdd(p.stat_sum + col, s0), atomicAdd(p.stat_sqsum + col, q0);
        if (col + 1 < p.cout) atomicAdd(p.stat_sum + col + 1, s1), atomicAdd(p.stat_sqsum + col + 1, q1);
      }
      if (p.bw_s1 != nullptr && col < p.cout) {
        const float fs0 = p.bw_fscale[col], fb0 = p.bw_fshift[col];
        const float fs1 = col + 1 < p.cout ? p.bw_fscale[col + 1] : 0.f, fb1 = col + 1 < p.cout ? p.bw_fshift[col + 1] : -1.f;
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
        const uint8_t* gbase = stg + (slab * kRowsPerSlab) * kEpiPitch + cp * 4;
        const uint8_t* ybase = ytile + (slab * kRowsPerSlab) * kEpiPitch + cp * 4;
#pragma unroll
        for (int r = 0; r < kRowsPerSlab; ++r) {
          // rows outside the image hold zeros in the staged gradient slice: they contribute nothing
          const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(gbase + r * kEpiPitch));
          const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ybase + r * kEpiPitch));
          const float g0 = fmaf(y.x, fs0, fb0) > 0.f ? g.x : 0.f;
          const float g1 = fmaf(y.y, fs1, fb1) > 0.f ? g.y : 0.f;
          a0 += g0, a1 += g1;
          b0 = fmaf(g0, y.x, b0), b1 = fmaf(g1, y.y, b1);
        }
        atomicAdd(p.bw_s1 + col, a0), atomicAdd(p.bw_s2 + col, b0);
        if (col + 1 < p.cout) atomicAdd(p.bw_s1 + col + 1, a1), atomicAdd(p.bw_s2 + col + 1, b1);
      }
      // coalesced store of the slice
      if (col0 + seg * 8 < p.n_store) {
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass) {
          const int r = pass * kRowsPerPass + r0;
          const int rh = h0 + (r >> p.bw_shift), rw = w0 + (r & (p.BW - 1));
          if (rh < p.H && rw < p.W) {
            const uint4 q = *reinterpret_cast<const uint4*>(stg + r * kEpiPitch + seg * 16);
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + img * p.out_img_stride + rh * p.out_row_stride +
                                static_cast<size_t>(rw) * p.ld_out + col0 + seg * 8;
            *reinterpret_cast<uint4*>(op) = q;
          }
        }
      }
      bar_sync_epilogue();  // staging slice fully consumed before the next slice overwrites it
    }
    }  // tile loop
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * BLOCK_N>(tmem_base);
  }
}

template <int BLOCK_N, int STAGES>
static int launch_persistent(const IgemmParams& p, int num_tiles, int grid, cudaStream_t stream) {
  using L = IgemmPersistSmem<BLOCK_N, STAGES>;
  static bool configured[64] = {};
  int dev = 0;
  SSEG_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !configured[dev]) {
    SSEG_CUDA(cudaFuncSetAttribute(igemm_persistent_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   L::kDynBytes));
    configured[dev] = true;
  }
  count_launch(1);
  return check_cuda(launch_k(igemm_persistent_kernel<BLOCK_N, STAGES>, dim3(grid), dim3(kNumThreads), L::kDynBytes, stream,
                             p, num_tiles),
                    "igemm_persistent_kernel launch");
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
  count_launch(1);
  return check_cuda(launch_k(igemm_kernel<BLOCK_N, STAGES>, dim3(grid), dim3(kNumThreads2), L::kDynBytes, stream, p),
                    "igemm_kernel launch");
}

#ifndef __CUSIM__
// =====================================================================================================
// CTA-pair variant (tcgen05 cta_group::2): clusters of two CTAs, each on one SM of a TPC, compute a 256 x 256 output tile.
//   CTA r of a pair owns M tile (2j + r): it streams its own 128-pixel A box and HALF of the 256-channel weight tile (B rows
//   n0 + r*128 .. +128) into its own shared memory - 32 KB per k-step for 128 x 256 x 64 of MMA work per SM (64 B/cycle
//   where the one-CTA 128 x 256 tile needs 96 and the 128 x 128 tile 128; the SM takes in ~80-100 B/cycle through TMA).
//   The leader (even rank) issues ONE tcgen05.mma.cta_group::2 per 16-wide K slice: A rows 0-127 / 128-255 come from the two
//   CTAs' shared memories, B columns 0-127 / 128-255 likewise, and each CTA's tensor memory receives its own 128 rows.
//   Barriers: every TMA of the pair signals the LEADER's full barrier (its producer posts the byte count of both CTAs, the
//   peer's producer arrives remotely); tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs.
//   Epilogue: each CTA drains its own tensor memory with the ordinary tile epilogue.
// Used for the long-K, wide-N layers (conv_last, cbr_deepsup, layer4's 3x3 convs and their data gradients).
// =====================================================================================================
template <int kPairN, int STAGES>
struct IgemmPairSmem {
  static constexpr int kBBytes = (kPairN / 2) * kBlockK * 2;   // this CTA's half of the weight tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOff = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOff + 256;
  static constexpr int kDynBytes = kTotal + 1024;
};

template <int kPairN, int STAGES>
__global__ void __launch_bounds__(kNumThreads, 1) igemm_pair_kernel(const __grid_constant__ IgemmParams p) {
  using L = IgemmPairSmem<kPairN, STAGES>;
  static_assert(L::kDynBytes <= 232448, "shared memory budget");
  static_assert(2 * 128 * (kPairN * 2 + 16) <= STAGES * L::kStageBytes, "the epilogue stages two tiles in the pipeline buffers");
  SSEG_DYN_SMEM(smem_raw);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();   // 0 = leader
  const bool leader = rank == 0;

  // tile decode: pairs walk N tiles fastest; the two CTAs of a pair take consecutive M tiles
  const int pair = blockIdx.x >> 1;
  const int n_tile = pair % p.n_tiles;
  int m_tile = (pair / p.n_tiles) * 2 + static_cast<int>(rank);
  const int tw = m_tile % p.tiles_w;
  m_tile /= p.tiles_w;
  const int th = m_tile % p.tiles_h;
  const int img = m_tile / p.tiles_h;
  const int h0 = th * p.BH, w0 = tw * p.BW, n0 = n_tile * kPairN;
  const int num_k_steps = p.num_k_steps;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);   // leader's arrive.expect_tx (bytes of both CTAs) + the peer producer's remote arrive
      mbar_init(&empty_bar[s], 1);  // one multicast tcgen05.commit
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsrc; ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1) tmem_alloc_2sm<kPairN>(tmem_ptr_smem);   // one warp of EACH CTA of the pair
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before anything signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();

  if (warp == 0) {
    // ===================== TMA producer (one per CTA; whole warp walks the loop, one elected lane issues) =====================
    int stage = 0, phase = 0;
    for (int t = 0; t < p.ntaps; ++t) {
      const int hh = h0 + p.tap_dh[t], ww = w0 + p.tap_dw[t];
      const int fixed_src = p.tap_src[t];
      const int nblk = fixed_src >= 0 ? p.src_blk_end[0] : p.blocks_per_tap;
      int src = fixed_src >= 0 ? fixed_src : 0, blk_begin = 0;
      for (int b = 0; b < nblk; ++b) {
        if (fixed_src < 0) {
          while (b >= p.src_blk_end[src]) {
            blk_begin = p.src_blk_end[src];
            ++src;
          }
        }
        mbar_wait_bounded(&empty_bar[stage], phase ^ 1);   // (own barrier: the leader's commit is multicast to both CTAs)
        uint8_t* sa = smem + stage * L::kStageBytes;
        const int kcol = p.tap_koff[t] + (fixed_src >= 0 ? 0 : p.src_choff[src]) + (b - blk_begin) * kBlockK;
        if (elect_one()) {
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
          else mbar_arrive_cluster(full_leader);
          tma_load_4d_2sm(sa, &p.tmA[src], full_leader, (b - blk_begin) * kBlockK, ww, hh, img);
          tma_load_2d_2sm(sa + kABytes, &p.tmB, full_leader, kcol, n0 + static_cast<int>(rank) * (kPairN / 2));
        }
        __syncwarp();
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the leader CTA's warp 1 (whole warp walks the loop, one elected lane issues) =====================
    if (leader) {
      constexpr uint32_t idesc = make_idesc(/*bf16*/ 1, 0, 0, 2 * kBlockM, kPairN);
      constexpr uint32_t dhi = smem_desc_hi_sw128(1024);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem), 16);
      int stage = 0, phase = 0;
      for (int ks = 0; ks < num_k_steps; ++ks) {
        mbar_wait_bounded(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + stage * (L::kStageBytes >> 4);
        const uint32_t b_lo = a_lo + (kABytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16_2sm(tmem_base, smem_desc_join(a_lo + 2 * k, dhi), smem_desc_join(b_lo + 2 * k, dhi), idesc, (ks | k) != 0);
          umma_commit_2sm(&empty_bar[stage], 3);   // this stage is free again in BOTH CTAs
        }
        __syncwarp();
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
      if (elect_one()) umma_commit_2sm(tmem_full_bar, 3);   // both CTAs' accumulators are complete
      __syncwarp();
    }
  } else {
    // ===================== epilogue: each CTA drains the 128 accumulator rows in its own tensor memory =====================
    igemm_tile_epilogue<kPairN>(p, smem, tmem_base, tmem_full_bar, warp, lane, h0, w0, n0, img);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's shared memory / barriers are no longer referenced by anyone
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<kPairN>(tmem_base);
  }
}

template <int kPairN, int STAGES>
static int launch_pair(const IgemmParams& p, int grid, cudaStream_t stream) {
  using L = IgemmPairSmem<kPairN, STAGES>;
  static bool configured[64] = {};
  int dev = 0;
  SSEG_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !configured[dev]) {
    SSEG_CUDA(cudaFuncSetAttribute(igemm_pair_kernel<kPairN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynBytes));
    configured[dev] = true;
  }
  count_launch(1);
  return check_cuda(launch_k_pair(igemm_pair_kernel<kPairN, STAGES>, dim3(grid), dim3(kNumThreads), L::kDynBytes, stream, p),
                    "igemm_pair_kernel launch");
}
#endif  // __CUSIM__

// =====================================================================================================
// Convolution + train-mode BatchNorm (+ shortcut, ReLU, Dropout2d mask) in ONE kernel.
//
// BatchNorm needs the statistics of the WHOLE layer between "accumulate" and "normalise". One persistent CTA per SM
// keeps the fp32 accumulators of ALL its tiles in tensor memory (512 columns = 4 tiles of 128 or 8 tiles of 64):
//   phase 1  per tile: main loop -> tcgen05.ld -> bf16 rounding -> (store y, which the backward pass reads) ->
//            per-channel sum / sum of squares of the rounded values -> atomics into the layer's statistics
//   barrier  every CTA arrives on one global counter (the grid is at most one CTA per SM, so all CTAs are resident)
//   phase 2  per tile: scale / shift from the now complete statistics (the CTA owning the first M tile of a channel
//            block publishes mean / inv_std / scale / shift and updates the running statistics) -> second tcgen05.ld
//            of the still-resident accumulator -> affine (+ shortcut, ReLU, dropout mask) -> store of the activation.
// Replaces sseg_conv_igemm(stats) + sseg_bn_finalize(SSEG_BN_TRAIN) + sseg_bn_apply: two launches and one full read
// of y per layer (reference: conv -> F.batch_norm(training) -> (+residual) -> ReLU, models/resnet.py:37-53,72-92;
// lib/nn/modules/batchnorm.py:58-61).
// =====================================================================================================
// Cross-GPU half of SynchronizedBatchNorm inside the cooperative kernels (world > 1): after the local grid barrier CTA 0
// publishes "this rank's partial sums of the layer are complete" into every peer's arena, every CTA waits for all peers'
// flags, and phase 2 pools the partial sums straight out of peer memory over NVLink - the protocol of csrc/peer.cu
// (step-number flags, never reset), without a kernel boundary.
struct CoopPeer {
  float* base[SSEG_MAX_PEERS];
  int world, rank;       // world <= 1: single-GPU behaviour
  long data_off;         // [sum C | sqsum C | count] (forward) or [s1 | s2raw] (backward) inside every rank's arena
  long data_stride;      // distance between the two vectors (C forward, the 8-padded C backward)
  long flag_off;
  const int* step;
};
// Bound on every cross-CTA / cross-GPU poll of the cooperative kernels (~10-30 s of polling; a legitimate wait is
// micro- to milliseconds): a protocol bug then aborts the grid with a launch failure instead of spinning forever.
constexpr unsigned int kCoopSpinLimit = 1u << 24;
// called by the 128 epilogue threads of every CTA right after the local grid barrier; t = epilogue thread index
__device__ __forceinline__ void coop_peer_handshake(const CoopPeer& pr, int t) {
  if (pr.world <= 1) return;
  const int step = *pr.step;
  if (blockIdx.x == 0 && t < pr.world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<int*>(pr.base[t]) + pr.flag_off + pr.rank, step);
  }
  if (t < pr.world) {
    const int* mine = reinterpret_cast<const int*>(pr.base[pr.rank]) + pr.flag_off + t;
    unsigned int polls = 0;
    while (ld_acquire_sys(mine) < step) {
      __nanosleep(32);
      if (++polls > kCoopSpinLimit) __trap();  // a peer that never arrives ends in a launch failure, not in a hung GPU
    }
  }
  bar_sync_epilogue();
}

struct IgemmBnParams {
  IgemmParams g;  // operands / geometry; g.out = y (bf16) or null; g.stat_sum / g.stat_sqsum = the layer's statistics
  __nv_bfloat16* a_out;  // activation, same geometry as y
  int ld_a;
  long a_row_stride, a_img_stride;
  const float* gamma;
  const float* beta;
  float eps, momentum, count;
  float *mean_out, *invstd_out, *scale_out, *shift_out;
  float *running_mean, *running_var;  // null: no running-statistics update
  const __nv_bfloat16* res;           // shortcut / top-down tensor or null
  int ld_res;
  long res_row_stride, res_img_stride;
  const float *rscale, *rshift;  // the shortcut's own BN affine (projection shortcut) or null
  const float* chanmul;          // [N][cout] Dropout2d keep-mask / (1-p) or null
  int relu, res_after_relu;
  long pix_per_img;  // to find the image of a pixel when the batch is viewed as one row of pixels (1x1 convs)
  unsigned int* counter;  // zeroed by the caller before every launch
  int num_tiles, tiles_per_cta;
  // synchronised branch (lib/nn/modules/batchnorm.py:63-81,123-139): pooled statistics, clamp(var, eps), accumulator
  // running statistics (tmp_*; running_iter and running = tmp / iter are finished by sseg_bn_running_from_tmp)
  CoopPeer peer;
  float *tmp_mean, *tmp_var, *running_iter, *count_out;
};

template <int BLOCK_N, int STAGES>
struct IgemmBnSmem {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kPitch = BLOCK_N * 2 + 16;
  static constexpr int kStgOff = STAGES * kStageBytes;
  static constexpr int kCoefOff = kStgOff + ((128 * kPitch + 1023) / 1024) * 1024;  // scale | shift | rscale | rshift
  static constexpr int kBarOff = kCoefOff + 4 * BLOCK_N * 4;
  static constexpr int kTotal = kBarOff + 256;
  static constexpr int kDynBytes = kTotal + 1024;
  static constexpr int kMaxTiles = 512 / BLOCK_N;  // accumulators that fit the SM's tensor memory
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kNumThreads, 1) igemm_bn_kernel(const __grid_constant__ IgemmBnParams q) {
  using L = IgemmBnSmem<BLOCK_N, STAGES>;
  const IgemmParams& p = q.g;
  SSEG_DYN_SMEM(smem_raw);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;  // [kMaxTiles]: accumulator i is complete
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + L::kMaxTiles);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_k_steps = p.num_k_steps;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < L::kMaxTiles; ++b) mbar_init(&tmem_full_bar[b], 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsrc; ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);  // every accumulator of this CTA stays resident until phase 2
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();

  if (warp == 0) {
    // ===================== TMA producer: runs ahead across tile boundaries =====================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < q.num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        int m_tile = tile / p.n_tiles;
        const int tw = m_tile % p.tiles_w;
        m_tile /= p.tiles_w;
        const int th = m_tile % p.tiles_h;
        const int img = m_tile / p.tiles_h;
        const int h0 = th * p.BH, w0 = tw * p.BW, n0 = n_tile * BLOCK_N;
        for (int t = 0; t < p.ntaps; ++t) {
          const int hh = h0 + p.tap_dh[t], ww = w0 + p.tap_dw[t];
          const int fixed_src = p.tap_src[t];
          const int nblk = fixed_src >= 0 ? p.src_blk_end[0] : p.blocks_per_tap;
          int src = fixed_src >= 0 ? fixed_src : 0, blk_begin = 0;
          for (int b = 0; b < nblk; ++b) {
            if (fixed_src < 0) {
              while (b >= p.src_blk_end[src]) {
                blk_begin = p.src_blk_end[src];
                ++src;
              }
            }
            mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * L::kStageBytes;
            mbar_expect_tx(&full_bar[stage], L::kStageBytes);
            tma_load_4d(sa, &p.tmA[src], &full_bar[stage], (b - blk_begin) * kBlockK, ww, hh, img);
            const int kcol = p.tap_koff[t] + (fixed_src >= 0 ? 0 : p.src_choff[src]) + (b - blk_begin) * kBlockK;
            tma_load_2d(sa + kABytes, &p.tmB, &full_bar[stage], kcol, n0);
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: accumulator i lives at TMEM columns [i*BLOCK_N, +BLOCK_N) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(/*bf16*/ 1, 0, 0, kBlockM, BLOCK_N);
      int stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < q.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t d_tmem = tmem_base + it * BLOCK_N;
        for (int ks = 0; ks < num_k_steps; ++ks) {
          mbar_wait_bounded(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16(d_tmem, da, db, idesc, (ks | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        umma_commit(&tmem_full_bar[it]);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps: phase 1, grid barrier, phase 2 =====================
    const int quarter = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int row = quarter * 32 + lane;
    const int t = threadIdx.x - 64;  // 0..127
    constexpr int kPitch = L::kPitch;
    uint8_t* stg = smem + L::kStgOff;
    float* coef = reinterpret_cast<float*>(smem + L::kCoefOff);  // [scale | shift | rscale | rshift][BLOCK_N]
    constexpr int kLanesPerRow = BLOCK_N / 8, kRowsPerPass = 128 / kLanesPerRow;
    const int seg = t % kLanesPerRow, r0 = t / kLanesPerRow;

    for (int phase2 = 0; phase2 < 2; ++phase2) {
      int it = 0;
      for (int tile = blockIdx.x; tile < q.num_tiles; tile += gridDim.x, ++it) {
        const int n_tile = tile % p.n_tiles;
        int m_tile = tile / p.n_tiles;
        const bool first_m_tile = m_tile == 0;
        const int tw = m_tile % p.tiles_w;
        m_tile /= p.tiles_w;
        const int th = m_tile % p.tiles_h;
        const int img = m_tile / p.tiles_h;
        const int h0 = th * p.BH, w0 = tw * p.BW, n0 = n_tile * BLOCK_N;
        const uint32_t d_tmem = tmem_base + it * BLOCK_N;
        const int hh = h0 + (row >> p.bw_shift), ww = w0 + (row & (p.BW - 1));
        const bool valid = (hh < p.H) && (ww < p.W);

        if (phase2) {
          // per-channel coefficients of this tile's channel block from the complete statistics
          if (t < BLOCK_N) {
            const int c = n0 + t;
            float sc = 0.f, sh = 0.f, rs = 1.f, rb = 0.f;
            if (c < p.cout) {
              float sum, sq, cnt, inv_std;
              if (q.peer.world > 1) {  // pooled over the ranks, in rank order (bit-identical on every rank)
                sum = 0.f, sq = 0.f, cnt = 0.f;
                for (int r = 0; r < q.peer.world; ++r) {
                  const float* st = q.peer.base[r] + q.peer.data_off;
                  sum += __ldcv(st + c), sq += __ldcv(st + q.peer.data_stride + c), cnt += __ldcv(st + 2 * q.peer.data_stride);
                }
              } else {
                sum = __ldcg(p.stat_sum + c), sq = __ldcg(p.stat_sqsum + c), cnt = q.count;
              }
              const float mean = sum / cnt;
              const float sumvar = sq - sum * mean;
              if (q.peer.world > 1) inv_std = rsqrtf(fmaxf(sumvar / cnt, q.eps));   // clamp(var, eps)^-1/2, batchnorm.py:139
              else inv_std = rsqrtf(fmaxf(sumvar / cnt, 0.f) + q.eps);
              const float gm = q.gamma ? q.gamma[c] : 1.f, bt = q.beta ? q.beta[c] : 0.f;
              sc = gm * inv_std, sh = bt - mean * gm * inv_std;
              if (q.rscale != nullptr) rs = q.rscale[c], rb = q.rshift[c];
              if (first_m_tile) {  // exactly one tile per channel block publishes (the backward pass reads these)
                q.mean_out[c] = mean, q.invstd_out[c] = inv_std, q.scale_out[c] = sc, q.shift_out[c] = sh;
                if (q.peer.world > 1) {
                  if (c == 0 && q.count_out != nullptr) *q.count_out = cnt;
                  if (q.tmp_mean != nullptr) {  // accumulator running statistics (batchnorm.py:131-137)
                    const float frac = 1.f - q.momentum;
                    q.tmp_mean[c] = q.tmp_mean[c] * frac + mean;
                    q.tmp_var[c] = q.tmp_var[c] * frac + sumvar / (cnt - 1.f);
                    if (c == 0) q.running_iter[0] = q.running_iter[0] * frac + 1.f;
                  }
                } else if (q.running_mean != nullptr) {
                  q.running_mean[c] = (1.f - q.momentum) * q.running_mean[c] + q.momentum * mean;
                  q.running_var[c] = (1.f - q.momentum) * q.running_var[c] + q.momentum * sumvar / (cnt - 1.f);
                }
              }
            }
            coef[t] = sc, coef[BLOCK_N + t] = sh, coef[2 * BLOCK_N + t] = rs, coef[3 * BLOCK_N + t] = rb;
          }
          bar_sync_epilogue();
          tc_fence_after();
        } else {
          mbar_wait_bounded(&tmem_full_bar[it], 0);
          tc_fence_after();
        }

        const long pix_off_a = img * q.a_img_stride + hh * q.a_row_stride + static_cast<long>(ww) * q.ld_a;
        const long pix_off_r = img * q.res_img_stride + hh * q.res_row_stride + static_cast<long>(ww) * q.ld_res;
        // image of this pixel (Dropout2d draws one value per image and channel)
        const long img_of_pix = q.pix_per_img > 0 ? (static_cast<long>(hh) * p.W + ww) / q.pix_per_img : img;
        (void)pix_off_a;
#pragma unroll 1
        for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
          uint32_t raw[32];
          tmem_ld_32x32(d_tmem + (static_cast<uint32_t>(quarter * 32) << 16) + chunk * 32, raw);
          tmem_ld_wait();
          float v[32];
          const int col0 = n0 + chunk * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j)  // y as stored: one bf16 rounding of the fp32 accumulator
            v[j] = __bfloat162float(__float2bfloat16_rn(__uint_as_float(raw[j])));
          if (phase2) {
            const float* sc = coef + chunk * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], sc[j], sc[BLOCK_N + j]);
            float rr[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) rr[j] = 0.f;
            if (q.res != nullptr && valid) {
              const __nv_bfloat16* rp = q.res + pix_off_r + col0;
#pragma unroll
              for (int g8 = 0; g8 < 4; ++g8) {
                if (col0 + g8 * 8 < p.n_store) {
                  const uint4 u = __ldg(reinterpret_cast<const uint4*>(rp + g8 * 8));
                  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h2[e]);
                    rr[g8 * 8 + 2 * e] = fmaf(f.x, sc[2 * BLOCK_N + g8 * 8 + 2 * e], sc[3 * BLOCK_N + g8 * 8 + 2 * e]);
                    rr[g8 * 8 + 2 * e + 1] =
                        fmaf(f.y, sc[2 * BLOCK_N + g8 * 8 + 2 * e + 1], sc[3 * BLOCK_N + g8 * 8 + 2 * e + 1]);
                  }
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = v[j];
              if (!q.res_after_relu) x += rr[j];
              if (q.relu) x = fmaxf(x, 0.f);
              if (q.res_after_relu) x += rr[j];
              v[j] = x;
            }
            if (q.chanmul != nullptr) {
              const float* m = q.chanmul + img_of_pix * p.cout + col0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.cout) v[j] *= __ldg(m + j);
            }
          }
          uint8_t* sp = stg + row * kPitch + chunk * 64;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            uint4 u = make_uint4(0u, 0u, 0u, 0u);  // rows outside the image contribute zeros to the statistics
            if (valid) {
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[g8 * 8 + 2 * e], v[g8 * 8 + 2 * e + 1]);
            }
            *reinterpret_cast<uint4*>(sp + g8 * 16) = u;
          }
        }
        tc_fence_before();
        bar_sync_epilogue();  // the staged tile is complete

        if (!phase2) {
          // statistics of the values as stored: thread = one pair of adjacent columns x one slab of rows
          constexpr int kPairs = BLOCK_N / 2, kSlabs = 128 / kPairs, kRowsPerSlab = 128 / kSlabs;
          const int cp = t % kPairs, slab = t / kPairs;
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          const uint8_t* base = stg + (slab * kRowsPerSlab) * kPitch + cp * 4;
#pragma unroll 8
          for (int r = 0; r < kRowsPerSlab; ++r) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(base + r * kPitch));
            s0 += f.x, s1 += f.y;
            q0 = fmaf(f.x, f.x, q0), q1 = fmaf(f.y, f.y, q1);
          }
          const int col = n0 + cp * 2;
          if (col < p.cout) atomicAdd(p.stat_sum + col, s0), atomicAdd(p.stat_sqsum + col, q0);
          if (col + 1 < p.cout) atomicAdd(p.stat_sum + col + 1, s1), atomicAdd(p.stat_sqsum + col + 1, q1);
        }
        // coalesced store of the staged tile: y in phase 1 (if requested), the activation in phase 2
        __nv_bfloat16* dst = phase2 ? q.a_out : reinterpret_cast<__nv_bfloat16*>(p.out);
        if (dst != nullptr && n0 + seg * 8 < p.n_store) {
          const long istr = phase2 ? q.a_img_stride : p.out_img_stride, rstr = phase2 ? q.a_row_stride : p.out_row_stride;
          const long ld = phase2 ? q.ld_a : p.ld_out;
#pragma unroll 4
          for (int pass = 0; pass < 128 / kRowsPerPass; ++pass) {
            const int r = pass * kRowsPerPass + r0;
            const int rh = h0 + (r >> p.bw_shift), rw = w0 + (r & (p.BW - 1));
            if (rh < p.H && rw < p.W) {
              const uint4 u = *reinterpret_cast<const uint4*>(stg + r * kPitch + seg * 16);
              *reinterpret_cast<uint4*>(dst + img * istr + rh * rstr + static_cast<long>(rw) * ld + n0 + seg * 8) = u;
            }
          }
        }
        bar_sync_epilogue();  // staging tile (and coefficients) free for the next tile
      }
      if (!phase2) {
        // ---- grid barrier: the layer's statistics are complete once every CTA has arrived
        __threadfence();
        bar_sync_epilogue();
        if (t == 0) {
          __threadfence();  // release: everything this CTA's threads published before the barrier above, then the arrival
          atomicAdd(q.counter, 1u);
          unsigned int polls = 0;
          while (ld_acquire_gpu(q.counter) < gridDim.x) {
            if (++polls > kCoopSpinLimit) __trap();  // see kCoopSpinLimit
          }
        }
        bar_sync_epilogue();
        coop_peer_handshake(q.peer, t);  // world > 1: every rank's partial sums are complete and visible
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int BLOCK_N, int STAGES>
static int launch_bn(const IgemmBnParams& q, int grid, cudaStream_t stream) {
  using L = IgemmBnSmem<BLOCK_N, STAGES>;
  static bool configured[64] = {};
  int dev = 0;
  SSEG_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !configured[dev]) {
    SSEG_CUDA(cudaFuncSetAttribute(igemm_bn_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   L::kDynBytes));
    configured[dev] = true;
  }
  count_launch(1);
  return check_cuda(launch_coop(igemm_bn_kernel<BLOCK_N, STAGES>, dim3(grid), dim3(kNumThreads), L::kDynBytes, stream, q),
                    "igemm_bn_kernel launch");
}

// Backward twin: the data-gradient GEMM of the CONSUMER layer + the whole BatchNorm backward of the PRODUCER layer.
//   phase 1  g = conv^T(dy_next) per tile (kept in TMEM) -> partial sums  s1 += g', s2raw += g'*y  (g' = g * ReLU mask
//            recomputed from the producer's saved conv output y)
//   barrier
//   phase 2  dy = ka*g' + kb*y + kc per channel from the complete sums (bn_bwd_apply's formula) -> only dy is written.
// Replaces sseg_conv_igemm_bnbwd + sseg_bn_bwd_apply for producers with a single consumer and no shortcut.
struct IgemmDgradBnParams {
  IgemmParams g;  // operands / geometry of the data gradient; g.bw_* = producer's y, forward scale / shift, sum slots
  __nv_bfloat16* dy_out;
  int ld_dy;
  long dy_row_stride, dy_img_stride;
  const float* mean;
  const float* invstd;
  float count;
  float* dgamma_out;
  unsigned int* counter;
  int num_tiles, tiles_per_cta;
  // synchronised branch: s1 / s2raw partial sums live in the rank's arena, totals are pooled over the ranks in phase 2;
  // dbeta / dgamma are stored divided by world (the gradient-bucket all-reduce sums them over the ranks again)
  CoopPeer peer;
  const float* count_dev;  // pooled pixel count written by the forward kernel
  float* dbeta_out;
};

template <int BLOCK_N, int STAGES>
struct IgemmDgradBnSmem {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kPitch = BLOCK_N * 2 + 16;
  static constexpr int kTileBytes = ((128 * kPitch + 1023) / 1024) * 1024;
  static constexpr int kStgOff = STAGES * kStageBytes;
  static constexpr int kYOff = kStgOff + kTileBytes;
  static constexpr int kCoefOff = kYOff + kTileBytes;
  static constexpr int kBarOff = kCoefOff + 4 * BLOCK_N * 4;
  static constexpr int kTotal = kBarOff + 256;
  static constexpr int kDynBytes = kTotal + 1024;
  static constexpr int kMaxTiles = 512 / BLOCK_N;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kNumThreads, 1) igemm_dgrad_bn_kernel(const __grid_constant__ IgemmDgradBnParams q) {
  using L = IgemmDgradBnSmem<BLOCK_N, STAGES>;
  const IgemmParams& p = q.g;
  SSEG_DYN_SMEM(smem_raw);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;  // [kMaxTiles]: accumulator i is complete
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + L::kMaxTiles);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_k_steps = p.num_k_steps;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < L::kMaxTiles; ++b) mbar_init(&tmem_full_bar[b], 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsrc; ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);  // every accumulator of this CTA stays resident until phase 2
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();

  if (warp == 0) {
    // ===================== TMA producer: runs ahead across tile boundaries =====================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < q.num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        int m_tile = tile / p.n_tiles;
        const int tw = m_tile % p.tiles_w;
        m_tile /= p.tiles_w;
        const int th = m_tile % p.tiles_h;
        const int img = m_tile / p.tiles_h;
        const int h0 = th * p.BH, w0 = tw * p.BW, n0 = n_tile * BLOCK_N;
        for (int t = 0; t < p.ntaps; ++t) {
          const int hh = h0 + p.tap_dh[t], ww = w0 + p.tap_dw[t];
          const int fixed_src = p.tap_src[t];
          const int nblk = fixed_src >= 0 ? p.src_blk_end[0] : p.blocks_per_tap;
          int src = fixed_src >= 0 ? fixed_src : 0, blk_begin = 0;
          for (int b = 0; b < nblk; ++b) {
            if (fixed_src < 0) {
              while (b >= p.src_blk_end[src]) {
                blk_begin = p.src_blk_end[src];
                ++src;
              }
            }
            mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * L::kStageBytes;
            mbar_expect_tx(&full_bar[stage], L::kStageBytes);
            tma_load_4d(sa, &p.tmA[src], &full_bar[stage], (b - blk_begin) * kBlockK, ww, hh, img);
            const int kcol = p.tap_koff[t] + (fixed_src >= 0 ? 0 : p.src_choff[src]) + (b - blk_begin) * kBlockK;
            tma_load_2d(sa + kABytes, &p.tmB, &full_bar[stage], kcol, n0);
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: accumulator i lives at TMEM columns [i*BLOCK_N, +BLOCK_N) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(/*bf16*/ 1, 0, 0, kBlockM, BLOCK_N);
      int stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < q.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t d_tmem = tmem_base + it * BLOCK_N;
        for (int ks = 0; ks < num_k_steps; ++ks) {
          mbar_wait_bounded(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16(d_tmem, da, db, idesc, (ks | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        umma_commit(&tmem_full_bar[it]);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps: phase 1, grid barrier, phase 2 =====================
    const int quarter = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int row = quarter * 32 + lane;
    const int t = threadIdx.x - 64;  // 0..127
    constexpr int kPitch = L::kPitch;
    uint8_t* stg = smem + L::kStgOff;
    uint8_t* ytile = smem + L::kYOff;                              // the producer layer's saved conv output, same tile
    float* coef = reinterpret_cast<float*>(smem + L::kCoefOff);  // [ka | kb | kc | fshift][BLOCK_N]: dy = ka*g' + kb*y + kc
    constexpr int kLanesPerRow = BLOCK_N / 8, kRowsPerPass = 128 / kLanesPerRow;
    const int seg = t % kLanesPerRow, r0 = t / kLanesPerRow;

    for (int phase2 = 0; phase2 < 2; ++phase2) {
      int it = 0;
      for (int tile = blockIdx.x; tile < q.num_tiles; tile += gridDim.x, ++it) {
        const int n_tile = tile % p.n_tiles;
        int m_tile = tile / p.n_tiles;
        const bool first_m_tile = m_tile == 0;
        const int tw = m_tile % p.tiles_w;
        m_tile /= p.tiles_w;
        const int th = m_tile % p.tiles_h;
        const int img = m_tile / p.tiles_h;
        const int h0 = th * p.BH, w0 = tw * p.BW, n0 = n_tile * BLOCK_N;
        const uint32_t d_tmem = tmem_base + it * BLOCK_N;
        const int hh = h0 + (row >> p.bw_shift), ww = w0 + (row & (p.BW - 1));
        const bool valid = (hh < p.H) && (ww < p.W);

        if (phase2) {
          // per-channel coefficients of this tile's channel block from the complete statistics
          if (t < BLOCK_N) {
            const int c = n0 + t;
            float ka = 0.f, kb = 0.f, kc = 0.f, fb = -1.f;
            if (c < p.cout) {
              // totals of the whole layer: s1 = sum g', s2 = inv_std * (sum g'*y - mean * s1) = sum g'*xhat
              float s1, s2raw, cnt = q.count, inv_w = 1.f;
              if (q.peer.world > 1) {
                s1 = 0.f, s2raw = 0.f;
                for (int r = 0; r < q.peer.world; ++r) {
                  const float* pt = q.peer.base[r] + q.peer.data_off;
                  s1 += __ldcv(pt + c), s2raw += __ldcv(pt + q.peer.data_stride + c);
                }
                cnt = *q.count_dev, inv_w = 1.f / (float)q.peer.world;
              } else {
                s1 = __ldcg(p.bw_s1 + c), s2raw = __ldcg(p.bw_s2 + c);
              }
              const float mu = q.mean[c], inv = q.invstd[c], fs = p.bw_fscale[c];
              const float s2 = inv * (s2raw - mu * s1);
              const float inv_m = 1.f / cnt;
              const float tt = fs * inv * s2 * inv_m;
              ka = fs, kb = -tt, kc = tt * mu - fs * s1 * inv_m, fb = p.bw_fshift[c];
              if (first_m_tile) {   // single GPU: dbeta = s1 is already in place (the partial sums went straight into it)
                if (q.dgamma_out != nullptr) q.dgamma_out[c] = s2 * inv_w;
                if (q.dbeta_out != nullptr) q.dbeta_out[c] = s1 * inv_w;
              }
            }
            coef[t] = ka, coef[BLOCK_N + t] = kb, coef[2 * BLOCK_N + t] = kc, coef[3 * BLOCK_N + t] = fb;
          }
          bar_sync_epilogue();
          tc_fence_after();
        } else {
          mbar_wait_bounded(&tmem_full_bar[it], 0);
          tc_fence_after();
        }

        const long pix_off_y = img * p.bw_img_stride + hh * p.bw_row_stride + static_cast<long>(ww) * p.bw_ld;
#pragma unroll 1
        for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
          uint32_t raw[32];
          tmem_ld_32x32(d_tmem + (static_cast<uint32_t>(quarter * 32) << 16) + chunk * 32, raw);
          tmem_ld_wait();
          float v[32];
          const int col0 = n0 + chunk * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j)  // y as stored: one bf16 rounding of the fp32 accumulator
            v[j] = __bfloat162float(__float2bfloat16_rn(__uint_as_float(raw[j])));
          if (phase2) {
            // dy = ka*g' + kb*y + kc with g' = g * [y*fscale + fshift > 0]; y is read row-wise (this thread's pixel)
            const float* kf = coef + chunk * 32;
            float yv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) yv[j] = 0.f;
            if (valid) {
              const __nv_bfloat16* yp = p.bw_y + pix_off_y + col0;
#pragma unroll
              for (int g8 = 0; g8 < 4; ++g8) {
                if (col0 + g8 * 8 < p.cout) {
                  const uint4 u = __ldg(reinterpret_cast<const uint4*>(yp + g8 * 8));
                  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h2[e]);
                    yv[g8 * 8 + 2 * e] = f.x, yv[g8 * 8 + 2 * e + 1] = f.y;
                  }
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float gm = fmaf(yv[j], kf[j], kf[3 * BLOCK_N + j]) > 0.f ? v[j] : 0.f;
              v[j] = fmaf(kf[j], gm, fmaf(kf[BLOCK_N + j], yv[j], kf[2 * BLOCK_N + j]));
            }
          }
          uint8_t* sp = stg + row * kPitch + chunk * 64;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            uint4 u = make_uint4(0u, 0u, 0u, 0u);  // rows outside the image contribute zeros to the statistics
            if (valid) {
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[g8 * 8 + 2 * e], v[g8 * 8 + 2 * e + 1]);
            }
            *reinterpret_cast<uint4*>(sp + g8 * 16) = u;
          }
        }
        tc_fence_before();
        bar_sync_epilogue();  // the staged tile is complete

        if (!phase2) {
          // the producer's saved conv output y, same tile geometry, copied with coalesced 16-byte loads
          {
            constexpr int kPasses = 128 / kRowsPerPass;
            uint4 u[kPasses];
#pragma unroll
            for (int pass = 0; pass < kPasses; ++pass) {
              const int r = pass * kRowsPerPass + r0;
              const int rh = h0 + (r >> p.bw_shift), rw = w0 + (r & (p.BW - 1));
              u[pass] = make_uint4(0u, 0u, 0u, 0u);
              if (rh < p.H && rw < p.W && n0 + seg * 8 < p.cout)
                u[pass] = __ldg(reinterpret_cast<const uint4*>(p.bw_y + img * p.bw_img_stride + rh * p.bw_row_stride +
                                                               static_cast<size_t>(rw) * p.bw_ld + n0 + seg * 8));
            }
#pragma unroll
            for (int pass = 0; pass < kPasses; ++pass)
              *reinterpret_cast<uint4*>(ytile + (pass * kRowsPerPass + r0) * kPitch + seg * 16) = u[pass];
          }
          bar_sync_epilogue();
          constexpr int kPairs = BLOCK_N / 2, kSlabs = 128 / kPairs, kRowsPerSlab = 128 / kSlabs;
          const int cp = t % kPairs, slab = t / kPairs;
          const int col = n0 + cp * 2;
          if (col < p.cout) {
            const float fs0 = p.bw_fscale[col], fb0 = p.bw_fshift[col];
            const float fs1 = col + 1 < p.cout ? p.bw_fscale[col + 1] : 0.f, fb1 = col + 1 < p.cout ? p.bw_fshift[col + 1] : -1.f;
            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
            const uint8_t* gbase = stg + (slab * kRowsPerSlab) * kPitch + cp * 4;
            const uint8_t* ybase = ytile + (slab * kRowsPerSlab) * kPitch + cp * 4;
#pragma unroll 8
            for (int r = 0; r < kRowsPerSlab; ++r) {
              const float2 gg = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(gbase + r * kPitch));
              const float2 yy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ybase + r * kPitch));
              const float g0 = fmaf(yy.x, fs0, fb0) > 0.f ? gg.x : 0.f;
              const float g1 = fmaf(yy.y, fs1, fb1) > 0.f ? gg.y : 0.f;
              a0 += g0, a1 += g1;
              b0 = fmaf(g0, yy.x, b0), b1 = fmaf(g1, yy.y, b1);
            }
            atomicAdd(p.bw_s1 + col, a0), atomicAdd(p.bw_s2 + col, b0);
            if (col + 1 < p.cout) atomicAdd(p.bw_s1 + col + 1, a1), atomicAdd(p.bw_s2 + col + 1, b1);
          }
        }
        // coalesced store of the staged tile: nothing in phase 1 (the gradient g itself never reaches HBM), dy in phase 2
        __nv_bfloat16* dst = phase2 ? q.dy_out : nullptr;
        if (dst != nullptr && n0 + seg * 8 < p.n_store) {
          const long istr = q.dy_img_stride, rstr = q.dy_row_stride;
          const long ld = q.ld_dy;
#pragma unroll 4
          for (int pass = 0; pass < 128 / kRowsPerPass; ++pass) {
            const int r = pass * kRowsPerPass + r0;
            const int rh = h0 + (r >> p.bw_shift), rw = w0 + (r & (p.BW - 1));
            if (rh < p.H && rw < p.W) {
              const uint4 u = *reinterpret_cast<const uint4*>(stg + r * kPitch + seg * 16);
              *reinterpret_cast<uint4*>(dst + img * istr + rh * rstr + static_cast<long>(rw) * ld + n0 + seg * 8) = u;
            }
          }
        }
        bar_sync_epilogue();  // staging tile (and coefficients) free for the next tile
      }
      if (!phase2) {
        // ---- grid barrier: the layer's statistics are complete once every CTA has arrived
        __threadfence();
        bar_sync_epilogue();
        if (t == 0) {
          __threadfence();  // release: everything this CTA's threads published before the barrier above, then the arrival
          atomicAdd(q.counter, 1u);
          unsigned int polls = 0;
          while (ld_acquire_gpu(q.counter) < gridDim.x) {
            if (++polls > kCoopSpinLimit) __trap();  // see kCoopSpinLimit
          }
        }
        bar_sync_epilogue();
        coop_peer_handshake(q.peer, t);  // world > 1: every rank's partial sums are complete and visible
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int BLOCK_N, int STAGES>
static int launch_dgrad_bn(const IgemmDgradBnParams& q, int grid, cudaStream_t stream) {
  using L = IgemmDgradBnSmem<BLOCK_N, STAGES>;
  static bool configured[64] = {};
  int dev = 0;
  SSEG_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !configured[dev]) {
    SSEG_CUDA(cudaFuncSetAttribute(igemm_dgrad_bn_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   L::kDynBytes));
    configured[dev] = true;
  }
  count_launch(1);
  return check_cuda(launch_coop(igemm_dgrad_bn_kernel<BLOCK_N, STAGES>, dim3(grid), dim3(kNumThreads), L::kDynBytes, stream,
                                q),
                    "igemm_dgrad_bn_kernel launch");
}

static int sm_count() {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0)
      num_sms = 148;
  }
  return num_sms;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e != nullptr && e[0] != 0) ? atoi(e) : dflt;
}

// Fills the geometry-derived part of the parameters shared by the forward and weight-gradient kernels.
// `box_pixels` = pixels per TMA box (128 for igemm M tiles, 64 for wgrad K steps).
struct GeomHost {
  int vn, vh, vw, BH, BW, bw_shift, tiles_h, tiles_w;
  int cin_total, blocks_per_tap, chan_per_src;
  bool flat;
  bool ragged;  // some source's channel count is not a multiple of 64
  int src_choff[SSEG_MAX_SRCS], src_c[SSEG_MAX_SRCS];
};

static int setup_geom(const sseg_conv_geom_t* g, int box_pixels, bool others_dense, GeomHost* gh, CUtensorMap* tmA,
                      int* src_blk_end, const char* who) {
  SSEG_REQUIRE(g != nullptr, "%s: null geometry", who);
  SSEG_REQUIRE(g->nsrc >= 1 && g->nsrc <= SSEG_MAX_SRCS, "%s: nsrc=%d out of range", who, g->nsrc);
  SSEG_REQUIRE(g->ntaps >= 1 && g->ntaps <= SSEG_MAX_TAPS, "%s: ntaps=%d out of range", who, g->ntaps);
  const int N = g->srcs[0].n, H = g->srcs[0].h, W = g->srcs[0].w;
  SSEG_REQUIRE(N >= 1 && H >= 1 && W >= 1, "%s: empty activation", who);
  bool pointwise = true, any_fixed = false;
  for (int t = 0; t < g->ntaps; ++t) {
    if (g->tap_dh[t] != 0 || g->tap_dw[t] != 0) pointwise = false;
    if (g->tap_src[t] >= 0) any_fixed = true;
    SSEG_REQUIRE(g->tap_src[t] >= -1 && g->tap_src[t] < g->nsrc, "%s: tap_src[%d]=%d out of range", who, t,
                 g->tap_src[t]);
    SSEG_REQUIRE(g->tap_koff[t] >= 0 && g->tap_koff[t] % 8 == 0, "%s: tap_koff[%d]=%d invalid", who, t,
                 g->tap_koff[t]);
  }
  // 1x1 convs see the whole batch as one long row of pixels (no halo, no per-image tiling waste)
  gh->vn = N, gh->vh = H, gh->vw = W;
  bool all_dense = others_dense;
  for (int s = 0; s < g->nsrc; ++s) all_dense = all_dense && act_is_dense(g->srcs[s]);
  gh->flat = pointwise && all_dense;
  if (gh->flat) gh->vn = 1, gh->vh = 1, gh->vw = N * H * W;
  int BW = box_pixels;
  while (BW > 8 && BW / 2 >= gh->vw) BW /= 2;  // smallest power of two >= W, in [8, box_pixels]
  gh->BW = BW, gh->BH = box_pixels / BW;
  gh->bw_shift = 0;
  while ((1 << gh->bw_shift) < BW) ++gh->bw_shift;
  gh->tiles_h = ceil_div(gh->vh, gh->BH), gh->tiles_w = ceil_div(gh->vw, BW);
  int cin_total = 0, blocks = 0;
  gh->ragged = false;
  for (int s = 0; s < g->nsrc; ++s) {
    const sseg_act_t& a = g->srcs[s];
    SSEG_REQUIRE(a.n == N && a.h == H && a.w == W, "%s: source %d shape mismatch", who, s);
    SSEG_REQUIRE(a.c % 4 == 0 && a.c > 0, "%s: source %d channels %d not a multiple of 4", who, s, a.c);
    SSEG_REQUIRE(a.ld % 8 == 0 && a.ld >= a.c, "%s: source %d ld %d invalid", who, s, a.ld);
    SSEG_REQUIRE(!any_fixed || a.c == g->srcs[0].c, "%s: per-tap sources must have equal channels", who);
    int rc = gh->flat ? get_tmap_act(&tmA[s], a.ptr, 2, 1, 1, gh->vw, a.c, a.ld, (long)gh->vw * a.ld,
                                     (long)gh->vw * a.ld, kBlockK, BW, gh->BH)
                      : get_tmap_act(&tmA[s], a.ptr, 2, N, H, W, a.c, a.ld, a.row_stride, a.img_stride, kBlockK, BW,
                                     gh->BH);
    if (rc) return rc;
    gh->src_choff[s] = cin_total, gh->src_c[s] = a.c;
    if (a.c % kBlockK != 0) gh->ragged = true;
    cin_total += a.c;
    blocks += ceil_div(a.c, kBlockK);
    src_blk_end[s] = blocks;
  }
  gh->cin_total = cin_total;
  gh->blocks_per_tap = blocks;
  gh->chan_per_src = g->srcs[0].c;
  return 0;
}

}  // namespace sseg

using namespace sseg;

struct EpilogueAffine {
  const float* scale;
  const float* shift;
  int relu;
};

static int conv_igemm_impl(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout, const sseg_act_t* out,
                           int out_f32, const float* bias, const sseg_act_t* addend, float* stat_sum, float* stat_sqsum,
                           const sseg_act_t* bw_y, const float* bw_fscale, const float* bw_fshift, float* bw_s1,
                           float* bw_s2, sseg_stream_t stream_, const EpilogueAffine* ep = nullptr,
                           IgemmParams* params_out = nullptr, int* block_n_out = nullptr,
                           const sseg_bn_fused_t* fin = nullptr, const sseg_act_t* bw_a = nullptr) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SSEG_REQUIRE(g != nullptr && out != nullptr && w_bf16 != nullptr, "sseg_conv_igemm: null argument");
  const int n_store = out->c;
  SSEG_REQUIRE(cout >= 1 && n_store >= cout && n_store % 8 == 0 && n_store <= out->ld,
               "sseg_conv_igemm: need cout <= out->c (mult of 8) <= out->ld, got %d %d %d", cout, n_store, out->ld);
  SSEG_REQUIRE((stat_sum == nullptr) == (stat_sqsum == nullptr), "sseg_conv_igemm: stat_sum/stat_sqsum must pair");
  SSEG_REQUIRE(!(out_f32 && stat_sum != nullptr), "sseg_conv_igemm: BN statistics are produced for bf16 outputs only");
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  GeomHost gh;
  const bool others_dense = act_is_dense(*out) && (addend == nullptr || act_is_dense(*addend));
  int rc = setup_geom(g, kBlockM, others_dense, &gh, p.tmA, p.src_blk_end, "sseg_conv_igemm");
  if (rc) return rc;
  SSEG_REQUIRE(out->n == g->srcs[0].n && out->h == g->srcs[0].h && out->w == g->srcs[0].w,
               "sseg_conv_igemm: output shape mismatch");
  p.BH = gh.BH, p.BW = gh.BW, p.bw_shift = gh.bw_shift;
  p.N = gh.vn, p.H = gh.vh, p.W = gh.vw;
  p.tiles_h = gh.tiles_h, p.tiles_w = gh.tiles_w;
  p.nsrc = g->nsrc;
  p.blocks_per_tap = gh.blocks_per_tap;
  p.ntaps = g->ntaps;
  p.num_k_steps = 0;
  for (int t = 0; t < g->ntaps; ++t) {
    p.tap_dh[t] = g->tap_dh[t], p.tap_dw[t] = g->tap_dw[t], p.tap_src[t] = g->tap_src[t];
    p.tap_koff[t] = g->tap_koff[t];
    const int span = g->tap_src[t] >= 0 ? gh.chan_per_src : gh.cin_total;
    SSEG_REQUIRE(g->tap_koff[t] + span <= w_ld, "sseg_conv_igemm: tap %d K range exceeds w_ld", t);
    p.num_k_steps += g->tap_src[t] >= 0 ? ceil_div(gh.chan_per_src, kBlockK) : gh.blocks_per_tap;
  }
  for (int s = 0; s < g->nsrc; ++s) p.src_choff[s] = gh.src_choff[s];
  // N tile: 128 unless that leaves most SMs without a CTA: then halve it to double the grid. Threshold swept on B200
  // (whole training step, CUDA-graph replay): 0 -> 6.58 ms, 80 -> 6.56, 160 -> 6.80, 300 -> 7.17, 600 -> 7.36.
  const int m_tiles = gh.vn * gh.tiles_h * gh.tiles_w;
  int block_n = cout <= 64 ? 64 : 128;
  static const int ntile_thresh = env_int("SSEG_NTILE_THRESH", 80);
  if (block_n == 128 && m_tiles * ceil_div(n_store, 128) <= ntile_thresh) block_n = 64;
  // N tile 256 for long-K layers with enough tiles (conv_last, cbr_deepsup, layer4's 3x3 convs and their data gradients):
  // the main loop is bound by what one SM can take in through TMA (~100 B/cycle measured); a 128 x 256 tile needs
  // 48 KB per 512 MMA cycles = 96 B/cycle where 128 x 128 needs 128 B/cycle. One CTA per SM then (192 KB of stages), so
  // only where the un-overlapped epilogue is small against the main loop.
  static const int pair_on = env_int("SSEG_IGEMM_2CTA", 0);
  static const int n256 = env_int("SSEG_IGEMM_N256", 1);
  static const int n256_min_ksteps = env_int("SSEG_IGEMM_N256_KSTEPS", 48);
  static const int n256_min_tiles = env_int("SSEG_IGEMM_N256_TILES", 96);   // (both knobs: test hooks for small shapes)
  if (n256 && block_n == 128 && n_store >= 256 && p.num_k_steps >= n256_min_ksteps && params_out == nullptr &&
      m_tiles * ceil_div(n_store, 256) >= n256_min_tiles)
    block_n = 256;
  // ... and as a CTA pair (256 x 256 tile over two SMs, tcgen05 cta_group::2) when the M tiles pair up
  bool use_pair = false;
#ifndef __CUSIM__
  use_pair = pair_on && block_n == 256 && m_tiles % 2 == 0 && fin == nullptr;
  // SSEG_IGEMM_2CTA=2: also the 128-wide tiles of mid-size layers as 256 x 128 pair tiles (24 KB per CTA and k-step)
  if (pair_on >= 2 && block_n == 128 && m_tiles % 2 == 0 && fin == nullptr && p.num_k_steps >= 8) use_pair = true;
#endif
  p.n_tiles = ceil_div(n_store, block_n);
  rc = get_tmap_2d(&p.tmB, w_bf16, 2, cout, w_ld, w_ld, kBlockK, use_pair ? block_n / 2 : block_n);
  if (rc) return rc;
  const int esz = out_f32 ? 4 : 2;
  p.out = out->ptr, p.out_f32 = out_f32, p.ld_out = out->ld, p.n_store = n_store, p.cout = cout;
  p.out_row_stride = out->row_stride, p.out_img_stride = out->img_stride;
  p.bias = bias;
  SSEG_REQUIRE((reinterpret_cast<uintptr_t>(out->ptr) & 15) == 0 && (out->ld * esz) % 16 == 0 &&
                   (out->row_stride * esz) % 16 == 0 && (out->img_stride * esz) % 16 == 0,
               "sseg_conv_igemm: output not 16B aligned");
  if (addend != nullptr) {
    SSEG_REQUIRE(addend->n == out->n && addend->h == out->h && addend->w == out->w && addend->c >= n_store,
                 "sseg_conv_igemm: addend shape mismatch");
    SSEG_REQUIRE((reinterpret_cast<uintptr_t>(addend->ptr) & 15) == 0 && addend->ld % 8 == 0 &&
                     addend->row_stride % 8 == 0 && addend->img_stride % 8 == 0,
                 "sseg_conv_igemm: addend not 16B aligned");
    p.addend = static_cast<const __nv_bfloat16*>(addend->ptr);
    p.ld_addend = addend->ld, p.add_row_stride = addend->row_stride, p.add_img_stride = addend->img_stride;
  }
  p.stat_sum = stat_sum, p.stat_sqsum = stat_sqsum;
  if (ep != nullptr) {
    SSEG_REQUIRE((ep->scale == nullptr) == (ep->shift == nullptr), "sseg_conv_igemm_affine: scale/shift must pair");
    SSEG_REQUIRE(ep->relu >= 0 && (ep->relu & 3) <= 2 && ep->relu <= 6, "sseg_conv_igemm_affine: relu mode %d", ep->relu);
    p.ep_scale = ep->scale, p.ep_shift = ep->shift, p.ep_relu = ep->relu;
  }
  if (bw_y != nullptr) {
    SSEG_REQUIRE(!out_f32 && bw_s1 && bw_s2 && ((bw_fscale && bw_fshift) || bw_a), "sseg_conv_igemm_bnbwd: null argument");
    if (bw_a != nullptr) {
      SSEG_REQUIRE(bw_a->n == out->n && bw_a->h == out->h && bw_a->w == out->w && bw_a->c >= cout && bw_a->ld % 8 == 0 &&
                       (reinterpret_cast<uintptr_t>(bw_a->ptr) & 15) == 0 && bw_a->row_stride % 8 == 0 &&
                       bw_a->img_stride % 8 == 0,
                   "sseg_conv_igemm_bnbwd_res: saved output shape / alignment mismatch");
      SSEG_REQUIRE(gh.flat == act_is_dense(*bw_a) || !gh.flat, "sseg_conv_igemm_bnbwd_res: a must be dense for 1x1 launches");
      p.bw_a = static_cast<const __nv_bfloat16*>(bw_a->ptr);
      p.bw_a_ld = bw_a->ld, p.bw_a_row_stride = bw_a->row_stride, p.bw_a_img_stride = bw_a->img_stride;
    }
    SSEG_REQUIRE(bw_y->n == out->n && bw_y->h == out->h && bw_y->w == out->w && bw_y->c >= cout && cout % 8 == 0 &&
                     bw_y->ld % 8 == 0 && (reinterpret_cast<uintptr_t>(bw_y->ptr) & 15) == 0,
                 "sseg_conv_igemm_bnbwd: y shape / alignment mismatch");
    SSEG_REQUIRE(gh.flat == act_is_dense(*bw_y) || !gh.flat, "sseg_conv_igemm_bnbwd: y must be dense for 1x1 launches");
    p.bw_y = static_cast<const __nv_bfloat16*>(bw_y->ptr);
    p.bw_ld = bw_y->ld, p.bw_row_stride = bw_y->row_stride, p.bw_img_stride = bw_y->img_stride;
    p.bw_fscale = bw_fscale, p.bw_fshift = bw_fshift, p.bw_s1 = bw_s1, p.bw_s2 = bw_s2;
  }
  if (fin != nullptr) {
    SSEG_REQUIRE(stat_sum != nullptr && fin->counter != nullptr && fin->mean_out && fin->invstd_out && fin->scale_out &&
                     fin->shift_out && fin->count > 1.f,
                 "sseg_conv_igemm_bnfin: statistics, counter and the four output vectors are required");
    SSEG_REQUIRE((fin->running_mean == nullptr) == (fin->running_var == nullptr), "sseg_conv_igemm_bnfin: running stats must pair");
    p.fin_counter = fin->counter;
    p.fin_gamma = fin->gamma, p.fin_beta = fin->beta;
    p.fin_eps = fin->eps, p.fin_momentum = fin->momentum, p.fin_count = fin->count;
    p.fin_mean = fin->mean_out, p.fin_invstd = fin->invstd_out, p.fin_scale = fin->scale_out, p.fin_shift = fin->shift_out;
    p.fin_rmean = fin->running_mean, p.fin_rvar = fin->running_var;
  }
  const int grid = gh.vn * p.tiles_h * p.tiles_w * p.n_tiles;
  if (params_out != nullptr) {  // the caller launches a different kernel over the same operands / geometry
    *params_out = p;
    *block_n_out = block_n;
    return grid;
  }
  // persistent CTAs (one per SM, double-buffered accumulators) once there are clearly more tiles than SMs
  static const int persistent_min_tiles = env_int("SSEG_IGEMM_PERSISTENT", 0);  // 0 = off; e.g. 200 = on for >= 200 tiles
  // (its sliced epilogue knows neither the fused finalize nor the mask-from-saved-output reduction: those launches keep the plain kernel)
  if (persistent_min_tiles > 0 && grid >= persistent_min_tiles && block_n != 256 && fin == nullptr && bw_a == nullptr) {
    static int num_sms = 0;
    if (num_sms == 0) {
      int dev = 0;
      SSEG_CUDA(cudaGetDevice(&dev));
      SSEG_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    static const int max_ctas = env_int("SSEG_IGEMM_PERSISTENT_CTAS", 0);  // test knob: force many tiles per CTA
    const int cap = max_ctas > 0 ? max_ctas : num_sms;
    const int pgrid = grid < cap ? grid : cap;
    if (block_n == 64) return launch_persistent<64, 8>(p, grid, pgrid, stream);
    return launch_persistent<128, 6>(p, grid, pgrid, stream);
  }
  // At most one CTA per SM (grid <= #SMs): the kernel is bound by TMA round trips (measured: ~35 B/cycle/SM with 96 KB in
  // flight, profiles/r2_summary.md), so the whole shared memory of the SM goes to pipeline stages (192 KB in flight).
#ifndef __CUSIM__
  if (use_pair && block_n == 256) return launch_pair<256, 5>(p, grid, stream);
  if (use_pair) return launch_pair<128, 6>(p, grid, stream);
#endif
  static const int deep = env_int("SSEG_IGEMM_DEEP", 1);
  if (deep && grid <= sm_count() && block_n != 256) {
    if (block_n == 64) return launch<64, 8>(p, grid, stream);
    return launch<128, 6>(p, grid, stream);
  }
  if (block_n == 256) return launch<256, 4>(p, grid, stream);
  if (block_n == 64) return launch<64, 4>(p, grid, stream);
  return launch<128, 3>(p, grid, stream);
}

extern "C" int sseg_conv_igemm(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout,
                               const sseg_act_t* out, int out_f32, const float* bias, const sseg_act_t* addend,
                               float* stat_sum, float* stat_sqsum, sseg_stream_t stream) {
  return conv_igemm_impl(g, w_bf16, w_ld, cout, out, out_f32, bias, addend, stat_sum, stat_sqsum, nullptr, nullptr, nullptr,
                         nullptr, nullptr, stream);
}

extern "C" int sseg_conv_igemm_bnfin(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout,
                                     const sseg_act_t* out, const sseg_bn_fused_t* bn, sseg_stream_t stream) {
  SSEG_REQUIRE(bn != nullptr, "sseg_conv_igemm_bnfin: null BatchNorm description");
  return conv_igemm_impl(g, w_bf16, w_ld, cout, out, 0, nullptr, nullptr, bn->stat_sum, bn->stat_sqsum, nullptr, nullptr,
                         nullptr, nullptr, nullptr, stream, nullptr, nullptr, nullptr, bn);
}

extern "C" int sseg_conv_igemm_affine(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout,
                                      const sseg_act_t* out, const float* scale, const float* shift, int relu,
                                      const sseg_act_t* addend, sseg_stream_t stream) {
  const EpilogueAffine ep = {scale, shift, relu};
  return conv_igemm_impl(g, w_bf16, w_ld, cout, out, 0, nullptr, addend, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                         nullptr, stream, &ep);
}

extern "C" int sseg_conv_igemm_bnbwd(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout,
                                     const sseg_act_t* out, const sseg_act_t* addend, const sseg_act_t* y,
                                     const float* fscale, const float* fshift, float* s1, float* s2_raw,
                                     sseg_stream_t stream) {
  SSEG_REQUIRE(y != nullptr, "sseg_conv_igemm_bnbwd: y required");
  return conv_igemm_impl(g, w_bf16, w_ld, cout, out, 0, nullptr, addend, nullptr, nullptr, y, fscale, fshift, s1, s2_raw,
                         stream);
}

extern "C" int sseg_conv_igemm_bnbwd_res(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout,
                                         const sseg_act_t* out, const sseg_act_t* addend, const sseg_act_t* y,
                                         const sseg_act_t* a, float* s1, float* s2_raw, sseg_stream_t stream) {
  SSEG_REQUIRE(y != nullptr && a != nullptr, "sseg_conv_igemm_bnbwd_res: y and a required");
  return conv_igemm_impl(g, w_bf16, w_ld, cout, out, 0, nullptr, addend, nullptr, nullptr, y, nullptr, nullptr, s1, s2_raw,
                         stream, nullptr, nullptr, nullptr, nullptr, a);
}

static int num_sms_of_current_device(int* out) {
  static int cached[64] = {};
  int dev = 0;
  SSEG_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || cached[dev] == 0) {
    int n = 0;
    SSEG_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    if (dev < 64) cached[dev] = n;
    *out = n;
    return 0;
  }
  *out = cached[dev];
  return 0;
}

static int fill_coop_peer(CoopPeer* out, const sseg_coop_peer_t* in, const char* who) {
  memset(out, 0, sizeof(*out));
  if (in == nullptr || in->world <= 1) return 0;
  SSEG_REQUIRE(in->bases != nullptr && in->world <= SSEG_MAX_PEERS && in->rank >= 0 && in->rank < in->world && in->step != nullptr,
               "%s: bad peer table (world %d rank %d)", who, in->world, in->rank);
  for (int r = 0; r < in->world; ++r) {
    SSEG_REQUIRE(in->bases[r] != nullptr, "%s: peer %d not mapped", who, r);
    out->base[r] = static_cast<float*>(in->bases[r]);
  }
  out->world = in->world, out->rank = in->rank;
  out->data_off = in->data_off, out->data_stride = in->data_stride, out->flag_off = in->flag_off, out->step = in->step;
  return 0;
}

static int conv_bn_train_impl(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout, const sseg_act_t* y,
                              const sseg_act_t* a_out, const sseg_bn_fused_t* bn, int query_only, sseg_stream_t stream_) {
  SSEG_REQUIRE(a_out != nullptr && bn != nullptr, "sseg_conv_bn_train: null argument");
  SSEG_REQUIRE(bn->stat_sum && bn->stat_sqsum && bn->counter && bn->mean_out && bn->invstd_out && bn->scale_out &&
                   bn->shift_out && bn->count > 1.f,
               "sseg_conv_bn_train: statistics / output vectors required");
  SSEG_REQUIRE((bn->running_mean == nullptr) == (bn->running_var == nullptr), "sseg_conv_bn_train: running stats must pair");
  SSEG_REQUIRE((bn->rscale == nullptr) == (bn->rshift == nullptr), "sseg_conv_bn_train: rscale/rshift must pair");
  // geometry / operand set-up is the plain convolution's; `y` may be omitted (inference-like use), then the activation
  // tensor stands in for the shape checks
  IgemmBnParams q;
  memset(&q, 0, sizeof(q));
  int block_n = 0;
  const sseg_act_t* shape_ref = y != nullptr ? y : a_out;
  const int tiles = conv_igemm_impl(g, w_bf16, w_ld, cout, shape_ref, 0, nullptr, nullptr, bn->stat_sum, bn->stat_sqsum, nullptr,
                                    nullptr, nullptr, nullptr, nullptr, stream_, nullptr, &q.g, &block_n);
  if (tiles < 0) return tiles;
  if (y == nullptr) q.g.out = nullptr;
  int sms = 0;
  int rc = num_sms_of_current_device(&sms);
  if (rc) return rc;
  const int max_tiles = 512 / block_n;
  const int per_cta = ceil_div(tiles, sms);
  if (query_only) return per_cta <= max_tiles ? 1 : 0;
  SSEG_REQUIRE(per_cta <= max_tiles, "sseg_conv_bn_train: %d tiles do not fit the tensor memory of %d SMs", tiles, sms);
  SSEG_REQUIRE(a_out->n == shape_ref->n && a_out->h == shape_ref->h && a_out->w == shape_ref->w && a_out->c == shape_ref->c &&
                   a_out->ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a_out->ptr) & 15) == 0 &&
                   a_out->row_stride % 8 == 0 && a_out->img_stride % 8 == 0,
               "sseg_conv_bn_train: activation output shape / alignment");
  // 1x1 launches view the batch as one row of pixels (conv_igemm_impl decides; it needs every operand dense)
  const bool flat = q.g.N == 1 && q.g.H == 1 && (shape_ref->n != 1 || shape_ref->h != 1);
  SSEG_REQUIRE(!flat || (act_is_dense(*a_out) && (bn->res == nullptr || act_is_dense(*bn->res))),
               "sseg_conv_bn_train: 1x1 launches need dense activation / shortcut tensors");
  q.a_out = static_cast<__nv_bfloat16*>(a_out->ptr);
  q.ld_a = a_out->ld, q.a_row_stride = a_out->row_stride, q.a_img_stride = a_out->img_stride;
  q.gamma = bn->gamma, q.beta = bn->beta, q.eps = bn->eps, q.momentum = bn->momentum, q.count = bn->count;
  q.mean_out = bn->mean_out, q.invstd_out = bn->invstd_out, q.scale_out = bn->scale_out, q.shift_out = bn->shift_out;
  q.running_mean = bn->running_mean, q.running_var = bn->running_var;
  if (bn->res != nullptr) {
    const sseg_act_t* r = bn->res;
    SSEG_REQUIRE(r->n == a_out->n && r->h == a_out->h && r->w == a_out->w && r->c >= a_out->c && r->ld % 8 == 0 &&
                     (reinterpret_cast<uintptr_t>(r->ptr) & 15) == 0 && r->row_stride % 8 == 0 && r->img_stride % 8 == 0,
                 "sseg_conv_bn_train: shortcut shape / alignment");
    q.res = static_cast<const __nv_bfloat16*>(r->ptr);
    q.ld_res = r->ld, q.res_row_stride = r->row_stride, q.res_img_stride = r->img_stride;
    q.rscale = bn->rscale, q.rshift = bn->rshift;
  }
  q.chanmul = bn->chanmul;
  q.relu = bn->relu, q.res_after_relu = bn->res_after_relu;
  q.pix_per_img = flat ? (long)shape_ref->h * shape_ref->w : 0;
  q.counter = bn->counter;
  q.num_tiles = tiles;
  q.tiles_per_cta = per_cta;
  rc = fill_coop_peer(&q.peer, bn->peer, "sseg_conv_bn_train");
  if (rc) return rc;
  if (q.peer.world > 1) {
    SSEG_REQUIRE(bn->count_out != nullptr, "sseg_conv_bn_train: count_out required with peers");
    SSEG_REQUIRE((bn->tmp_running_mean == nullptr) == (bn->tmp_running_var == nullptr) &&
                     (bn->tmp_running_mean == nullptr) == (bn->running_iter == nullptr),
                 "sseg_conv_bn_train: tmp_running_mean / tmp_running_var / running_iter must come together");
    q.tmp_mean = bn->tmp_running_mean, q.tmp_var = bn->tmp_running_var, q.running_iter = bn->running_iter;
    q.count_out = bn->count_out;
  }
  const int grid = ceil_div(tiles, per_cta);  // <= number of SMs: every CTA is resident, the in-kernel barrier is safe
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (block_n == 64) return launch_bn<64, 6>(q, grid, stream);
  return launch_bn<128, 4>(q, grid, stream);
}

extern "C" int sseg_conv_bn_train(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout, const sseg_act_t* y,
                                  const sseg_act_t* a_out, const sseg_bn_fused_t* bn, sseg_stream_t stream) {
  return conv_bn_train_impl(g, w_bf16, w_ld, cout, y, a_out, bn, 0, stream);
}

extern "C" int sseg_conv_bn_train_fits(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout,
                                       const sseg_act_t* y, const sseg_act_t* a_out, const sseg_bn_fused_t* bn) {
  return conv_bn_train_impl(g, w_bf16, w_ld, cout, y, a_out, bn, 1, nullptr);
}

static int conv_dgrad_bn_impl(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout, const sseg_act_t* y,
                              const sseg_act_t* dy_out, const float* fscale, const float* fshift, const float* mean,
                              const float* invstd, float count, float* s1, float* s2_raw, float* dgamma_out,
                              unsigned int* counter, const sseg_coop_peer_t* peer, const float* count_dev, float* dbeta_out,
                              int query_only, sseg_stream_t stream_) {
  SSEG_REQUIRE(y != nullptr && dy_out != nullptr && fscale && fshift && mean && invstd && s1 && s2_raw && counter &&
                   count > 1.f,
               "sseg_conv_dgrad_bn: null argument");
  IgemmDgradBnParams q;
  memset(&q, 0, sizeof(q));
  int block_n = 0;
  // geometry / operand / fused-reduction set-up is sseg_conv_igemm_bnbwd's, with dy_out standing in for the (never
  // written) gradient tensor in the shape checks
  const int tiles = conv_igemm_impl(g, w_bf16, w_ld, cout, dy_out, 0, nullptr, nullptr, nullptr, nullptr, y, fscale, fshift, s1,
                                    s2_raw, stream_, nullptr, &q.g, &block_n);
  if (tiles < 0) return tiles;
  q.g.out = nullptr;
  int sms = 0;
  int rc = num_sms_of_current_device(&sms);
  if (rc) return rc;
  const int per_cta = ceil_div(tiles, sms);
  if (query_only) return per_cta <= 512 / block_n ? 1 : 0;
  SSEG_REQUIRE(per_cta <= 512 / block_n, "sseg_conv_dgrad_bn: %d tiles do not fit the tensor memory of %d SMs", tiles, sms);
  q.dy_out = static_cast<__nv_bfloat16*>(dy_out->ptr);
  q.ld_dy = dy_out->ld, q.dy_row_stride = dy_out->row_stride, q.dy_img_stride = dy_out->img_stride;
  q.mean = mean, q.invstd = invstd, q.count = count, q.dgamma_out = dgamma_out, q.counter = counter;
  q.num_tiles = tiles, q.tiles_per_cta = per_cta;
  rc = fill_coop_peer(&q.peer, peer, "sseg_conv_dgrad_bn");
  if (rc) return rc;
  if (q.peer.world > 1) {
    SSEG_REQUIRE(count_dev != nullptr && dbeta_out != nullptr && dgamma_out != nullptr,
                 "sseg_conv_dgrad_bn: count_dev / dbeta_out / dgamma_out required with peers");
    q.count_dev = count_dev, q.dbeta_out = dbeta_out;
  }
  const int grid = ceil_div(tiles, per_cta);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (block_n == 64) return launch_dgrad_bn<64, 6>(q, grid, stream);
  return launch_dgrad_bn<128, 4>(q, grid, stream);
}

extern "C" int sseg_conv_dgrad_bn(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout, const sseg_act_t* y,
                                  const sseg_act_t* dy_out, const float* fscale, const float* fshift, const float* mean,
                                  const float* invstd, float count, float* s1, float* s2_raw, float* dgamma_out,
                                  unsigned int* counter, const sseg_coop_peer_t* peer, const float* count_dev,
                                  float* dbeta_out, sseg_stream_t stream) {
  return conv_dgrad_bn_impl(g, w_bf16, w_ld, cout, y, dy_out, fscale, fshift, mean, invstd, count, s1, s2_raw, dgamma_out,
                            counter, peer, count_dev, dbeta_out, 0, stream);
}

extern "C" int sseg_conv_dgrad_bn_fits(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout,
                                       const sseg_act_t* y, const sseg_act_t* dy_out, const float* fscale,
                                       const float* fshift, const float* mean, const float* invstd, float count, float* s1,
                                       float* s2_raw, float* dgamma_out, unsigned int* counter) {
  return conv_dgrad_bn_impl(g, w_bf16, w_ld, cout, y, dy_out, fscale, fshift, mean, invstd, count, s1, s2_raw, dgamma_out,
                            counter, nullptr, nullptr, nullptr, 1, nullptr);
}

// =====================================================================================================
// Weight gradient:  dW[co][koff_t + ci] += sum_pixels dY[pixel, co] * X_t[pixel shifted by tap t, ci]
//
//   GEMM with M = 128 output channels (co), N = BLOCK_N input channels (ci) of ONE tap, K = pixels.
//   Both operands come straight from the NHWC tensors as TMA boxes of (64 channels x 64 pixels): the channel
//   (M / N) index is the contiguous one, i.e. both are "MN-major" UMMA operands:
//       canonical SWIZZLE_128B MN-major layout ((8,m),(8,k)) : ((1,LBO),(8,SBO))   [units of 16 bytes]
//       -> 8 pixels x 128 B swizzle atoms, SBO = 1024 B between 8-pixel groups, LBO = 8 KB between 64-channel boxes.
//   One K step = 64 pixels = 4 UMMA instructions (K=16 pixels each, +2 KB start-address advance).
//   K is split across CTAs (blockIdx / num_tiles = split index); partial sums land with fp32 vector atomics.
// =====================================================================================================
namespace sseg {

constexpr int kWgKPix = 64;                       // pixels per K step (one TMA box)
constexpr int kWgBoxBytes = kWgKPix * 64 * 2;     // 8 KB: 64 pixels x 64 channels bf16

struct WgradParams {
  CUtensorMap tmX[SSEG_MAX_SRCS];
  CUtensorMap tmDY;
  int nsrc;
  int src_blk_end[SSEG_MAX_SRCS];
  int ntaps;
  int tap_dh[SSEG_MAX_TAPS], tap_dw[SSEG_MAX_TAPS], tap_src[SSEG_MAX_TAPS], tap_koff[SSEG_MAX_TAPS];
  int ci_tiles_per_tap;  // number of BLOCK_N-wide ci tiles per tap
  int m_tiles;           // ceil(cout / 128)
  int num_tiles;         // m_tiles * ntaps * ci_tiles_per_tap
  int BH, BW, tiles_h, tiles_w;
  int total_boxes, boxes_per_split;
  int cout, ci_span;     // valid co rows; valid ci per tap
  int ragged;            // some source has channels % 64 != 0: ci tiles are single 64-blocks of ONE source
  int src_choff[SSEG_MAX_SRCS], src_c[SSEG_MAX_SRCS];
  float* dw;
  long dw_ld;
};

template <int BLOCK_N, int STAGES>
struct WgradSmem {
  static constexpr int kABytes_ = 2 * kWgBoxBytes;                 // 128 co
  static constexpr int kBBytes_ = (BLOCK_N / 64) * kWgBoxBytes;    // BLOCK_N ci
  static constexpr int kStageBytes = kABytes_ + kBBytes_;
  static constexpr int kBarOff = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOff + 256;
  static constexpr int kDynBytes = kTotal + 1024;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kNumThreads2) wgrad_kernel(const __grid_constant__ WgradParams p) {
  using L = WgradSmem<BLOCK_N, STAGES>;
  SSEG_DYN_SMEM(smem_raw);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int split = blockIdx.x / p.num_tiles;
  int tile = blockIdx.x % p.num_tiles;
  const int ci_tile = tile % p.ci_tiles_per_tap;
  tile /= p.ci_tiles_per_tap;
  const int tap = tile % p.ntaps;
  const int m_tile = tile / p.ntaps;
  const int co0 = m_tile * 128, ci0 = ci_tile * BLOCK_N;
  const int box_begin = split * p.boxes_per_split;
  const int box_end = min(box_begin + p.boxes_per_split, p.total_boxes);
  const int num_k_steps = box_end - box_begin;  // host guarantees >= 1

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);  // one arrive.expect_tx per producer warp (dy boxes, x boxes)
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsrc; ++s) tma_prefetch_desc(&p.tmX[s]);
    tma_prefetch_desc(&p.tmDY);
  }
  if (warp == 1) tmem_alloc<BLOCK_N>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();

  if (warp == 0 || warp == 6) {
    // warp 0 issues the two dy boxes (A, 128 output channels), warp 6 the x boxes (B, BLOCK_N input channels) of every k-step;
    // the whole warp walks the loop (warp-uniform control flow), one elected lane issues
    {
      const bool is_a = warp == 0;
      // resolve, once, which source / channel offset each 64-channel B box comes from
      int bsrc[BLOCK_N / 64], bchan[BLOCK_N / 64];
#pragma unroll
      for (int j = 0; j < BLOCK_N / 64; ++j) {
        const int blk = ci0 / 64 + j;
        if (p.tap_src[tap] >= 0) {
          bsrc[j] = p.tap_src[tap];
          bchan[j] = blk * 64;
        } else {
          int s = 0, begin = 0;
          while (s + 1 < p.nsrc && blk >= p.src_blk_end[s]) begin = p.src_blk_end[s], ++s;
          bsrc[j] = s;
          bchan[j] = (blk - begin) * 64;  // beyond the last source -> out of bounds -> zeros
        }
      }
      const int dh = p.tap_dh[tap], dw = p.tap_dw[tap];
      int stage = 0, phase = 0;
      for (int box = box_begin; box < box_end; ++box) {
        int r = box;
        const int tw = r % p.tiles_w;
        r /= p.tiles_w;
        const int th = r % p.tiles_h;
        const int img = r / p.tiles_h;
        const int h0 = th * p.BH, w0 = tw * p.BW;
        mbar_wait_warp(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::kStageBytes;
        uint8_t* sb = sa + L::kABytes_;
        if (elect_one()) {
          if (is_a) {
            mbar_expect_tx(&full_bar[stage], L::kABytes_);
            tma_load_4d(sa, &p.tmDY, &full_bar[stage], co0, w0, h0, img);
            tma_load_4d(sa + kWgBoxBytes, &p.tmDY, &full_bar[stage], co0 + 64, w0, h0, img);
          } else {
            mbar_expect_tx(&full_bar[stage], L::kBBytes_);
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_4d(sb + j * kWgBoxBytes, &p.tmX[bsrc[j]], &full_bar[stage], bchan[j], w0 + dw, h0 + dh, img);
          }
        }
        __syncwarp();
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the whole warp walks the loop, one elected lane issues; descriptors stepped through their low words
    constexpr uint32_t idesc = make_idesc(/*bf16*/ 1, /*A MN-major*/ 1, /*B MN-major*/ 1, 128, BLOCK_N);
    constexpr uint32_t dhi = smem_desc_hi_sw128(1024);
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem), kWgBoxBytes);
    int stage = 0, phase = 0;
    for (int ks = 0; ks < num_k_steps; ++ks) {
      mbar_wait_warp(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t a_lo = a_lo0 + stage * (L::kStageBytes >> 4);
      const uint32_t b_lo = a_lo + (L::kABytes_ >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kWgKPix / 16; ++k)
          umma_bf16(tmem_base, smem_desc_join(a_lo + k * (2048 >> 4), dhi), smem_desc_join(b_lo + k * (2048 >> 4), dhi), idesc,
                    (ks | k) != 0);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == STAGES) stage = 0, phase ^= 1;
    }
    if (elect_one()) umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int co = co0 + quarter * 32 + lane;
    const bool valid = co < p.cout;
    // columns of dW this tile owns: [colbase, colbase + width) inside the tap (width may exceed BLOCK_N: no clipping)
    int colbase = ci0, width = p.ci_span - ci0;
    if (p.tap_src[tap] < 0) {
      const int blk0 = ci0 / 64;
      int s = 0, begin = 0;
      while (s + 1 < p.nsrc && blk0 >= p.src_blk_end[s]) begin = p.src_blk_end[s], ++s;
      const int chan = (blk0 - begin) * 64;
      colbase = p.src_choff[s] + chan;
      width = p.ragged ? p.src_c[s] - chan : p.ci_span - colbase;
    }
    float* drow = p.dw + static_cast<size_t>(co) * p.dw_ld + p.tap_koff[tap] + colbase;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
      uint32_t raw[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + chunk * 32, raw);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (chunk * 32 + g * 4 < width) {
            float4 f = make_float4(__uint_as_float(raw[g * 4]), __uint_as_float(raw[g * 4 + 1]),
                                   __uint_as_float(raw[g * 4 + 2]), __uint_as_float(raw[g * 4 + 3]));
            atomicAdd(reinterpret_cast<float4*>(drow + chunk * 32 + g * 4), f);
          }
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
static int launch_wgrad(const WgradParams& p, int grid, cudaStream_t stream) {
  using L = WgradSmem<BLOCK_N, STAGES>;
  static bool configured[64] = {};
  int dev = 0;
  SSEG_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !configured[dev]) {
    SSEG_CUDA(cudaFuncSetAttribute(wgrad_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   L::kDynBytes));
    configured[dev] = true;
  }
  count_launch(1);
  return check_cuda(launch_k(wgrad_kernel<BLOCK_N, STAGES>, dim3(grid), dim3(kNumThreads2), L::kDynBytes, stream, p),
                    "wgrad_kernel launch");
}

}  // namespace sseg

extern "C" int sseg_conv_wgrad(const sseg_conv_geom_t* g, const sseg_act_t* dy, int cout, float* dw, long dw_ld,
                               sseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SSEG_REQUIRE(dy != nullptr && dw != nullptr, "sseg_conv_wgrad: null argument");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  GeomHost gh;
  int rc = setup_geom(g, kWgKPix, act_is_dense(*dy), &gh, p.tmX, p.src_blk_end, "sseg_conv_wgrad");
  if (rc) return rc;
  SSEG_REQUIRE(dy->n == g->srcs[0].n && dy->h == g->srcs[0].h && dy->w == g->srcs[0].w,
               "sseg_conv_wgrad: dy shape mismatch");
  SSEG_REQUIRE(dy->c % 8 == 0 && dy->ld % 8 == 0 && dy->ld >= dy->c, "sseg_conv_wgrad: dy channels/ld must be x8");
  SSEG_REQUIRE((reinterpret_cast<uintptr_t>(dw) & 15) == 0 && dw_ld % 4 == 0, "sseg_conv_wgrad: dw not 16B aligned");
  rc = gh.flat ? get_tmap_act(&p.tmDY, dy->ptr, 2, 1, 1, gh.vw, dy->c, dy->ld, (long)gh.vw * dy->ld,
                              (long)gh.vw * dy->ld, 64, gh.BW, gh.BH)
               : get_tmap_act(&p.tmDY, dy->ptr, 2, dy->n, dy->h, dy->w, dy->c, dy->ld, dy->row_stride, dy->img_stride,
                              64, gh.BW, gh.BH);
  if (rc) return rc;
  p.nsrc = g->nsrc;
  p.ntaps = g->ntaps;
  bool fixed = g->tap_src[0] >= 0;
  p.ci_span = fixed ? gh.chan_per_src : gh.cin_total;
  for (int t = 0; t < g->ntaps; ++t) {
    p.tap_dh[t] = g->tap_dh[t], p.tap_dw[t] = g->tap_dw[t], p.tap_src[t] = g->tap_src[t];
    p.tap_koff[t] = g->tap_koff[t];
    SSEG_REQUIRE((g->tap_src[t] >= 0) == fixed, "sseg_conv_wgrad: taps must be all concat or all per-plane");
    SSEG_REQUIRE(g->tap_koff[t] % 4 == 0 && g->tap_koff[t] + p.ci_span <= dw_ld, "sseg_conv_wgrad: tap %d K range", t);
  }
  p.ragged = gh.ragged ? 1 : 0;
  for (int s = 0; s < g->nsrc; ++s) p.src_choff[s] = gh.src_choff[s], p.src_c[s] = gh.src_c[s];
  int block_n = (!gh.ragged && p.ci_span % 128 == 0) ? 128 : 64;
  // 256 input channels per tile where that still leaves enough tiles: 48 KB per 512 MMA cycles instead of 32 KB per 256
  // (the main loop is bound by what the SM takes in through TMA, see conv_igemm_impl)
  static const int wg256 = env_int("SSEG_WGRAD_N256", 1);
  static const int wg256_min_tiles = env_int("SSEG_WGRAD_N256_TILES", 74);
  if (wg256 && block_n == 128 && p.ci_span % 256 == 0 && !fixed &&
      ceil_div(cout, 128) * g->ntaps * (p.ci_span / 256) >= wg256_min_tiles)
    block_n = 256;
  p.ci_tiles_per_tap = fixed ? ceil_div(gh.chan_per_src, block_n) : (gh.ragged ? gh.blocks_per_tap : p.ci_span / block_n);
  SSEG_REQUIRE(cout >= 1 && cout <= dy->c, "sseg_conv_wgrad: cout %d vs dy channels %d", cout, dy->c);
  p.cout = cout;
  p.m_tiles = ceil_div(p.cout, 128);
  p.num_tiles = p.m_tiles * p.ntaps * p.ci_tiles_per_tap;
  p.BH = gh.BH, p.BW = gh.BW, p.tiles_h = gh.tiles_h, p.tiles_w = gh.tiles_w;
  p.total_boxes = gh.vn * gh.tiles_h * gh.tiles_w;
  // split K (pixels) only until the grid covers the 148 SMs once, and keep >= 4 K steps per CTA: every extra split
  // costs a full fp32 atomic pass over the tile. Swept on B200 (step time): 74 -> 6.50 ms, 148 -> 6.44, 200 -> 6.46,
  // 296 -> 6.56, 592 -> 6.80, 1184 -> 7.11.
  static const int wgrad_ctas = env_int("SSEG_WGRAD_CTAS", 148);
  int splits = ceil_div(wgrad_ctas, p.num_tiles);
  splits = max(1, min(splits, ceil_div(p.total_boxes, 4)));
  p.boxes_per_split = ceil_div(p.total_boxes, splits);
  splits = ceil_div(p.total_boxes, p.boxes_per_split);
  p.dw = dw, p.dw_ld = dw_ld;
  const int grid = p.num_tiles * splits;
  if (block_n == 256) return launch_wgrad<256, 4>(p, grid, stream);
  if (block_n == 64) return launch_wgrad<64, 4>(p, grid, stream);
  return launch_wgrad<128, 3>(p, grid, stream);
}

#ifdef SSEG_TRACE
extern "C" int sseg_debug_read_trace(long long* host_out, long count) {
  return sseg::check_cuda(cudaMemcpyFromSymbol(host_out, sseg::g_trace, count * sizeof(long long)), "read trace");
}
extern "C" int sseg_debug_clear_trace() {
  void* p = nullptr;
  if (cudaGetSymbolAddress(&p, sseg::g_trace) != cudaSuccess) return -1;
  return sseg::check_cuda(cudaMemset(p, 0, sizeof(long long) * 4096 * 16), "clear trace");
}
#endif
