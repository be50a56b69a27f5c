// Self-test kernels for the simulator's detectors (tests/test_cusim.py): each has one deliberate defect.
#include <stdint.h>

#include "../../semantic-segmentation-pytorch_b200/csrc/common.h"
#include "../../semantic-segmentation-pytorch_b200/csrc/ptx.cuh"

namespace sseg {

// (1) shared-memory race: the reader does not wait for the writer (missing __syncthreads)
__global__ void race_kernel(float* out, int with_barrier) {
  __shared__ float slot[64];
  slot[threadIdx.x] = static_cast<float>(threadIdx.x);
  if (with_barrier) __syncthreads();
  out[threadIdx.x] = slot[63 - threadIdx.x];
}

// (2) pipeline bug: the consumer waits on the parity the producer never completes
__global__ void stuck_kernel(int wrong_parity) {
  SSEG_DYN_SMEM(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) mbar_arrive(bar);                       // completes phase 0
  if (threadIdx.x == 32) mbar_wait(bar, wrong_parity ? 1 : 0);  // parity 1 = phase 1: never completes
}

// (3) out-of-bounds global store one element past the buffer
__global__ void oob_kernel(float* out, int n) { out[threadIdx.x < n ? threadIdx.x : n] = 1.f; }

}  // namespace sseg

extern "C" {
int cusim_selftest_race(float* out, int with_barrier) {
  return (int)sseg::launch_k(sseg::race_kernel, dim3(1), dim3(64), 0, nullptr, out, with_barrier);
}
int cusim_selftest_stuck(int wrong_parity) {
  return (int)sseg::launch_k(sseg::stuck_kernel, dim3(1), dim3(64), 64, nullptr, wrong_parity);
}
int cusim_selftest_oob(float* out, int n) { return (int)sseg::launch_k(sseg::oob_kernel, dim3(1), dim3(64), 0, nullptr, out, n); }
}
