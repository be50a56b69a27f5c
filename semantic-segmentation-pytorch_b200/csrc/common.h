// Host-side helpers shared by every translation unit of libsseg_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sseg_b200.h"

namespace sseg {

// Error plumbing of the C ABI: every entry point returns 0 or a negative code and
// leaves a message retrievable with sseg_last_error(). No exceptions cross the boundary.
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
void count_launch(int n);

#define SSEG_CUDA(expr)                                  \
  do {                                                   \
    int _rc = ::sseg::check_cuda((expr), #expr);         \
    if (_rc) return _rc;                                 \
  } while (0)

#define SSEG_REQUIRE(cond, ...)                          \
  do {                                                   \
    if (!(cond)) {                                       \
      ::sseg::set_error(__VA_ARGS__);                    \
      return SSEG_ERR_ARG;                               \
    }                                                    \
  } while (0)

// TMA tensor-map construction (cuTensorMapEncodeTiled through the runtime's driver entry point,
// so the library does not link libcuda). Maps are cached by their full geometry.
//   rank-4 bf16/f32 activation view (C, W, H, N) with 128B swizzle, box (box_c, box_w, box_h, 1)
int get_tmap_act(CUtensorMap* out, const void* ptr, int elem_bytes, int n, int h, int w, int c, long ld,
                 long row_stride, long img_stride, int box_c, int box_w, int box_h);
inline bool act_is_dense(const sseg_act_t& a) {
  return a.row_stride == (long)a.w * a.ld && a.img_stride == (long)a.h * a.row_stride;
}
//   rank-2 matrix [rows][cols] (cols contiguous) with 128B swizzle, box (box_cols, box_rows)
int get_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, long rows, long cols, long ld, int box_cols,
                int box_rows);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Kernel launch with the programmatic-stream-serialization attribute (PDL) unless disabled (SSEG_PDL=0 in the
// environment or sseg_set_pdl(0)); then a plain launch.
bool pdl_enabled();
#ifdef __CUSIM__
// CPU simulator build (tests/cusim): the same kernels run as OS threads; a cooperative launch runs all CTAs concurrently.
// A translation unit without static __shared__ variables may let ordinary launches run several CTAs at a time.
#ifndef SSEG_SIM_PARALLEL_CTAS
#define SSEG_SIM_PARALLEL_CTAS 1
#endif
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t, Args&&... args) {
  return static_cast<cudaError_t>(::cusim::launch(grid, block, smem, false, [=]() { kernel(args...); }, SSEG_SIM_PARALLEL_CTAS));
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_coop(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t, Args&&... args) {
  return static_cast<cudaError_t>(::cusim::launch(grid, block, smem, true, [=]() { kernel(args...); }));
}
#else
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Launch as clusters of two CTAs (CTA pairs on the two SMs of a TPC: tcgen05 cta_group::2 kernels), with the PDL attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_pair(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Cooperative launch: the driver starts the grid only when ALL its CTAs can be resident at once, which is what makes an
// in-kernel grid barrier safe even when other streams compete for the SMs (used by the fused conv+BN kernel). No PDL
// attribute here: the two are not combined.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_coop(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                               Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif  // __CUSIM__

}  // namespace sseg
