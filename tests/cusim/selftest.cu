// Self-test kernels for the simulator's detectors (tests/test_cusim.py): each has one deliberate defect.
#include <stdint.h>

#include <chrono>

#include "../../semantic-segmentation-pytorch_b200/csrc/common.h"
#include "../../semantic-segmentation-pytorch_b200/csrc/ptx.cuh"

namespace sseg {

// (1) shared-memory race: the reader does not wait for the writer (missing __syncthreads). Repeated, so that the conflicting
// accesses overlap in time however the host schedules the threads (ThreadSanitizer is a dynamic detector).
static int g_started = 0;  // relaxed counter: carries no happens-before edge, it only makes the threads overlap in time

__global__ void race_kernel(float* out, int with_barrier, int iters) {
  __shared__ float slot[64];
  float acc = 0.f;
  if (threadIdx.x == 0) __atomic_store_n(&g_started, 0, __ATOMIC_RELAXED);
  __syncthreads();
  __atomic_fetch_add(&g_started, 1, __ATOMIC_RELAXED);
  while (__atomic_load_n(&g_started, __ATOMIC_RELAXED) < 64) sched_yield();  // every thread is inside the body from here on
  const auto t0 = std::chrono::steady_clock::now();
  // without the barrier: keep going for a quarter of a second of wall-clock, so that every pair of conflicting threads is
  // alive and running at the same time at some point (the detector needs both accesses in its recent history)
  for (int it = 0; with_barrier ? it < iters
                                : std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 0.25; ++it) {
    slot[threadIdx.x] = static_cast<float>(threadIdx.x + it);
    if (with_barrier) __syncthreads();
    acc += slot[63 - threadIdx.x] - static_cast<float>(it);
    if (with_barrier) __syncthreads();
  }
  out[threadIdx.x] = with_barrier ? acc / static_cast<float>(iters) : acc;
}

// (2) pipeline bug: the consumer waits on the parity the producer never completes
__global__ void stuck_kernel(int wrong_parity) {
  SSEG_DYN_SMEM(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) mbar_arrive(bar);                       // completes phase 0
  __syncthreads();                                              // ... before anybody waits
  if (threadIdx.x == 32) mbar_wait(bar, wrong_parity ? 1 : 0);  // parity 1 = phase 1: never completes
}

// (3) out-of-bounds global store one element past the buffer
__global__ void oob_kernel(float* out, int n) { out[threadIdx.x < n ? threadIdx.x : n] = 1.f; }

}  // namespace sseg

extern "C" {
int cusim_selftest_race(float* out, int with_barrier) {
  return (int)sseg::launch_k(sseg::race_kernel, dim3(1), dim3(64), 0, nullptr, out, with_barrier, with_barrier ? 200 : 50000);
}
int cusim_selftest_stuck(int wrong_parity) {
  return (int)sseg::launch_k(sseg::stuck_kernel, dim3(1), dim3(64), 64, nullptr, wrong_parity);
}
int cusim_selftest_oob(float* out, int n) { return (int)sseg::launch_k(sseg::oob_kernel, dim3(1), dim3(64), 0, nullptr, out, n); }
}
