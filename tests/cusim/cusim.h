// cusim: a functional CPU model of the CUDA execution model + the sm_100a instructions the kernels of this repo use.
//
// TEST INFRASTRUCTURE ONLY. The product is semantic-segmentation-pytorch_b200/csrc/*.cu compiled by nvcc for sm_100a; this header lets
// the SAME .cu sources compile with g++ (`-x c++ -include cusim.h`) into tests/cusim/_build/libsseg_sim.so, where
//   * every CUDA thread is an OS thread (CTAs of an ordinary launch run one after another, CTAs of a cooperative launch
//     run concurrently), __syncthreads / named barriers / warp shuffles are real rendezvous,
//   * mbarrier, TMA (cp.async.bulk.tensor with 128B swizzle and out-of-bounds zero fill), tensor memory and
//     tcgen05.mma / .commit / .ld are modelled functionally (cusim_ptx.h),
//   * a wait that does not complete within CUSIM_TIMEOUT seconds aborts the launch with a report of who waits on what
//     (a deadlock on the GPU box costs a strike; here it costs a test failure).
// It checks control flow, indexing, barrier phases and arithmetic -- not timing, not the memory model.
#pragma once
#include <time.h>
#define __CUSIM__ 1

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <sched.h>
#include <stdint.h>
#include <string.h>

#include <functional>

// ---- keywords: the toolkit's host_defines.h makes __global__/__device__/... empty for a host compiler -------------------
#undef __shared__
#define __shared__ static /* CTAs that declare static shared memory run one at a time (see cusim::launch) */
#ifndef __grid_constant__
#define __grid_constant__
#endif
#undef __launch_bounds__
#define __launch_bounds__(...)

namespace cusim {

struct Cta;
struct ThreadCtx {
  uint3 tid, bid;
  dim3 bdim, gdim;
  Cta* cta;
  int linear_tid;
};
extern thread_local ThreadCtx tl;

void syncthreads();
void named_barrier(int id, int count);
void syncwarp();
uint64_t warp_exchange(uint64_t v, int src_lane);  // every lane of the warp calls; returns lane src_lane's v
uint8_t* dyn_smem();
void atomic_lock();
void atomic_unlock();
void trap();
void spin_pause();  // called from polling loops: yields and throws when the launch has been aborted
int launch(dim3 grid, dim3 block, size_t smem, bool cooperative, const std::function<void()>& body, int parallel_ctas = 1);

}  // namespace cusim

#define threadIdx (::cusim::tl.tid)
#define blockIdx (::cusim::tl.bid)
#define blockDim (::cusim::tl.bdim)
#define gridDim (::cusim::tl.gdim)
#define warpSize 32

// ---- execution-model builtins ---------------------------------------------------------------------------------------------
static inline void __syncthreads() { ::cusim::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { ::cusim::syncwarp(); }
static inline void __threadfence() { __sync_synchronize(); }
static inline void __threadfence_system() { __sync_synchronize(); }
static inline void __nanosleep(unsigned) { ::cusim::spin_pause(); }
static inline void __trap() { ::cusim::trap(); }
static inline long long clock64() {   // "SM cycles": host nanoseconds x 2 (a 2 GHz clock), monotonic
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (static_cast<long long>(ts.tv_sec) * 1000000000ll + ts.tv_nsec) * 2;
}

template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int = 32) {
  static_assert(sizeof(T) <= 8, "shuffle of a type wider than 8 bytes");
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  raw = ::cusim::warp_exchange(raw, src & 31);
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}
template <typename T>
static inline T __shfl_xor_sync(unsigned m, T v, int lanemask, int w = 32) {
  return __shfl_sync(m, v, (::cusim::tl.linear_tid & 31) ^ lanemask, w);
}
template <typename T>
static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int w = 32) {
  const int lane = ::cusim::tl.linear_tid & 31;
  return __shfl_sync(m, v, lane + (int)d < 32 ? lane + (int)d : lane, w);
}

// ---- memory builtins ------------------------------------------------------------------------------------------------------
template <typename T>
static inline T __ldg(const T* p) { return *p; }
template <typename T>
static inline T __ldcv(const T* p) { return *(const volatile T*)p; }
template <typename T>
static inline T __ldcg(const T* p) { return *(const volatile T*)p; }

static inline float atomicAdd(float* p, float v) {
  ::cusim::atomic_lock();
  const float old = *p;
  *p = old + v;
  ::cusim::atomic_unlock();
  return old;
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
static inline float4 atomicAdd(float4* p, float4 v) {  // red.global.add.v4.f32
  ::cusim::atomic_lock();
  const float4 old = *p;
  p->x += v.x, p->y += v.y, p->z += v.z, p->w += v.w;
  ::cusim::atomic_unlock();
  return old;
}
static inline float2 atomicAdd(float2* p, float2 v) {
  ::cusim::atomic_lock();
  const float2 old = *p;
  p->x += v.x, p->y += v.y;
  ::cusim::atomic_unlock();
  return old;
}
static inline int atomicMax(int* p, int v) {
  ::cusim::atomic_lock();
  const int old = *p;
  if (v > old) *p = v;
  ::cusim::atomic_unlock();
  return old;
}
static inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }

template <typename T>
static inline cudaError_t cudaFuncSetAttribute(T*, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- math -------------------------------------------------------------------------------------------------------------------
static inline float __uint_as_float(unsigned u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline unsigned __float_as_uint(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float __int_as_float(int u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline int __float_as_int(float f) {
  int u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
#define __expf(x) expf(x)
#define __logf(x) logf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __saturatef(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
template <typename T>
static inline T min(T a, T b) { return b < a ? b : a; }
template <typename T>
static inline T max(T a, T b) { return a < b ? b : a; }
