// SyncBN statistics exchange over NVLink peer memory (no NCCL call on the BN critical path).
//
// Every rank owns one "arena" (cudaMalloc'ed, exported with CUDA IPC and mapped by all peers). The conv epilogue /
// bn_bwd_reduce accumulate a layer's partial sums into the LOCAL arena; the finalize kernels below then
//   1. publish "my partials of layer L are complete" by storing the step number into a flag slot in EVERY peer's arena
//      (st.release.sys over NVLink),
//   2. spin on their own flag slots until every peer has published (ld.acquire.sys on local memory),
//   3. read the peers' partials straight out of peer memory (cache-volatile loads), sum them in a fixed rank order
//      (bit-identical on every rank) and finish the BN arithmetic.
// Flags are monotonically increasing step numbers, so nothing is ever reset; the per-step gradient-bucket all-reduce is
// the barrier that makes re-zeroing the partials for the next step safe.
// Replaces, for world_size > 1: lib/nn/modules/batchnorm.py:98-117 (_data_parallel_master: ReduceAddCoalesced +
// Broadcast) and the host-side SyncMaster/SlavePipe rendezvous (comm.py:18-131).
#include "common.h"
#include "ptx.cuh"
#include "peer.cuh"

namespace sseg {

__global__ void peer_step_kernel(int* step) {
  pdl_sync();
  if (threadIdx.x == 0) *step += 1;
}

// forward: pooled (sum, sqsum, count) over ranks -> mean, inv_std, scale, shift (+ accumulator running statistics):
// the SynchronizedBatchNorm parallel branch, lib/nn/modules/batchnorm.py:123-139.
__global__ void __launch_bounds__(256) bn_finalize_peer_kernel(const PeerTable pt, long stats_off, long flag_off,
                                                               const int* step_ptr, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float eps, float momentum,
                                                               int update_running, float* running_mean, float* running_var,
                                                               float* tmp_mean, float* tmp_var, float* running_iter,
                                                               float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                                               float* __restrict__ scale, float* __restrict__ shift,
                                                               float* __restrict__ count_out, int C, long inbox_off) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int cc = c < C ? c : 0;
  float s, q, cnt;
  if (inbox_off >= 0) {
    // push protocol: no block-level handshake; every thread sends its channel's two sums (thread 0 also the pixel count) to
    // all peers, then polls the peers' messages for the same slots in local memory
    const int step = *step_ptr, n = 2 * C + 1;
    const float* mine = pt.base[pt.rank] + stats_off;
    const float ms = mine[cc], mq = mine[C + cc], mc = mine[2 * C];
    if (c < C) ll_push(pt, inbox_off, n, c, ms, step), ll_push(pt, inbox_off, n, C + c, mq, step);
    if (c == 0) ll_push(pt, inbox_off, n, 2 * C, mc, step);
    s = ll_pool(pt, inbox_off, n, cc, ms, step);
    q = ll_pool(pt, inbox_off, n, C + cc, mq, step);
    cnt = ll_pool(pt, inbox_off, n, 2 * C, mc, step);
  } else {
    peer_handshake(pt, flag_off, *step_ptr, blockIdx.x == 0);
    peer_sum2(pt, stats_off + cc, stats_off + C + cc, &s, &q);   // in flight together with the count loads below
    cnt = peer_sum(pt, stats_off + 2 * C);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *count_out = cnt;
    if (update_running) running_iter[0] = running_iter[0] * (1.f - momentum) + 1.f;
  }
  if (c >= C) return;
  const float mean = s / cnt;
  const float sumvar = q - s * mean;
  const float bias_var = sumvar / cnt, unbias_var = sumvar / (cnt - 1.f);
  const float inv_std = rsqrtf(fmaxf(bias_var, eps));
  if (update_running) {
    const float frac = 1.f - momentum;  // running_mean/var are refreshed from tmp/iter by the follow-up kernel
    const float tm = tmp_mean[c] * frac + mean;
    const float tv = tmp_var[c] * frac + unbias_var;
    tmp_mean[c] = tm, tmp_var[c] = tv;
  }
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  mean_out[c] = mean, invstd_out[c] = inv_std;
  scale[c] = g * inv_std;
  shift[c] = b - mean * g * inv_std;
}
// running_mean/var = tmp / running_iter (batchnorm.py:136-137) once running_iter has been advanced
__global__ void bn_running_from_tmp_kernel(const float* tmp_mean, const float* tmp_var, const float* running_iter,
                                           float* running_mean, float* running_var, int C) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float it = running_iter[0];
  running_mean[c] = tmp_mean[c] / it;
  running_var[c] = tmp_var[c] / it;
}

// backward: totals of (s1 = sum g', s2 = sum g' * xhat) over ranks; also stores dbeta = s1/world, dgamma = s2/world into
// the gradient buffer (the gradient-bucket all-reduce sums those over the ranks again).
__global__ void __launch_bounds__(256) bn_bwd_peer_sum_kernel(const PeerTable pt, long part_off, long flag_off,
                                                              const int* step_ptr, float* __restrict__ s1_tot,
                                                              float* __restrict__ s2_tot, float* __restrict__ dbeta,
                                                              float* __restrict__ dgamma, const float* __restrict__ mean,
                                                              const float* __restrict__ invstd, int s2_raw, int C,
                                                              long inbox_off) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  float a, b;
  if (inbox_off >= 0) {
    if (c >= C) return;
    const int step = *step_ptr, n = 2 * C;
    const float* mine = pt.base[pt.rank] + part_off;
    const float ma = mine[c], mb = mine[C + c];
    ll_push(pt, inbox_off, n, c, ma, step), ll_push(pt, inbox_off, n, C + c, mb, step);
    a = ll_pool(pt, inbox_off, n, c, ma, step);
    b = ll_pool(pt, inbox_off, n, C + c, mb, step);
  } else {
    peer_handshake(pt, flag_off, *step_ptr, blockIdx.x == 0);
    if (c >= C) return;
    peer_sum2(pt, part_off + c, part_off + C + c, &a, &b);
  }
  if (s2_raw) b = invstd[c] * (b - mean[c] * a);  // partials were sum g'*y (fused dgrad epilogue): convert to sum g'*xhat
  s1_tot[c] = a, s2_tot[c] = b;
  const float inv_w = 1.f / (float)pt.world;
  dbeta[c] = a * inv_w, dgamma[c] = b * inv_w;
}

int make_peer_table(PeerTable* t, void* const* bases, int world, int rank, const char* who) {
  SSEG_REQUIRE(bases != nullptr && world >= 1 && world <= SSEG_MAX_PEERS && rank >= 0 && rank < world,
               "%s: bad peer table (world %d rank %d)", who, world, rank);
  memset(t, 0, sizeof(*t));
  for (int r = 0; r < world; ++r) {
    SSEG_REQUIRE(bases[r] != nullptr, "%s: peer %d not mapped", who, r);
    t->base[r] = static_cast<float*>(bases[r]);
  }
  t->world = world, t->rank = rank;
  return 0;
}

}  // namespace sseg

using namespace sseg;

extern "C" {

int sseg_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  SSEG_REQUIRE(ptr && handle64 && bytes > 0, "sseg_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  SSEG_CUDA(cudaMalloc(ptr, bytes));
  SSEG_CUDA(cudaMemset(*ptr, 0, bytes));
  cudaIpcMemHandle_t h;
  SSEG_CUDA(cudaIpcGetMemHandle(&h, *ptr));
  memcpy(handle64, &h, 64);
  return 0;
}

int sseg_peer_open(const unsigned char* handle64, void** ptr) {
  SSEG_REQUIRE(ptr && handle64, "sseg_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  SSEG_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int sseg_peer_close(void* ptr) {
  SSEG_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int sseg_peer_free(void* ptr) {
  SSEG_CUDA(cudaFree(ptr));
  return 0;
}

int sseg_peer_step(int* step, sseg_stream_t st) {
  SSEG_REQUIRE(step, "sseg_peer_step: null");
  count_launch(1);
  return check_cuda(launch_k(peer_step_kernel, dim3(1), dim3(32), 0, (cudaStream_t)st, step), "peer_step_kernel");
}

static int bn_finalize_peer_impl(void* const* bases, int world, int rank, long stats_off, long flag_off, long inbox_off,
                                 const int* step, const float* gamma, const float* beta, float eps, float momentum,
                                 int update_running, float* running_mean, float* running_var, float* tmp_mean,
                                 float* tmp_var, float* running_iter, float* mean_out, float* invstd_out, float* scale,
                                 float* shift, float* count_out, int C, sseg_stream_t st) {
  PeerTable t;
  int rc = make_peer_table(&t, bases, world, rank, "sseg_bn_finalize_peer");
  if (rc) return rc;
  SSEG_REQUIRE(step && mean_out && invstd_out && scale && shift && count_out && C > 0, "sseg_bn_finalize_peer: null");
  SSEG_REQUIRE(!update_running || (running_mean && running_var && tmp_mean && tmp_var && running_iter),
               "sseg_bn_finalize_peer: running buffers required");
  const int grid = (C + 255) / 256;
  rc = check_cuda(launch_k(bn_finalize_peer_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)st, t, stats_off, flag_off,
                           step, gamma, beta, eps, momentum, update_running, running_mean, running_var, tmp_mean, tmp_var,
                           running_iter, mean_out, invstd_out, scale, shift, count_out, C, inbox_off),
                  "bn_finalize_peer_kernel");
  count_launch(1);
  // update_running == 2: the accumulators and running_iter are advanced here, running_mean / running_var are refreshed by
  // the caller (sseg_bn_running_from_tmp, off the forward pass's dependency chain)
  if (rc == 0 && update_running == 1) {
    rc = check_cuda(launch_k(bn_running_from_tmp_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)st,
                             (const float*)tmp_mean, (const float*)tmp_var, (const float*)running_iter, running_mean,
                             running_var, C),
                    "bn_running_from_tmp_kernel");
    count_launch(1);
  }
  return rc;
}

int sseg_bn_finalize_peer(void* const* bases, int world, int rank, long stats_off, long flag_off, const int* step,
                          const float* gamma, const float* beta, float eps, float momentum, int update_running,
                          float* running_mean, float* running_var, float* tmp_mean, float* tmp_var, float* running_iter,
                          float* mean_out, float* invstd_out, float* scale, float* shift, float* count_out, int C,
                          sseg_stream_t st) {
  return bn_finalize_peer_impl(bases, world, rank, stats_off, flag_off, -1, step, gamma, beta, eps, momentum, update_running,
                               running_mean, running_var, tmp_mean, tmp_var, running_iter, mean_out, invstd_out, scale,
                               shift, count_out, C, st);
}

int sseg_bn_finalize_peer_ll(void* const* bases, int world, int rank, long stats_off, long inbox_off, const int* step,
                             const float* gamma, const float* beta, float eps, float momentum, int update_running,
                             float* running_mean, float* running_var, float* tmp_mean, float* tmp_var,
                             float* running_iter, float* mean_out, float* invstd_out, float* scale, float* shift,
                             float* count_out, int C, sseg_stream_t st) {
  SSEG_REQUIRE(inbox_off >= 0 && inbox_off % 2 == 0, "sseg_bn_finalize_peer_ll: the inbox must be 8-byte aligned");
  return bn_finalize_peer_impl(bases, world, rank, stats_off, 0, inbox_off, step, gamma, beta, eps, momentum, update_running,
                               running_mean, running_var, tmp_mean, tmp_var, running_iter, mean_out, invstd_out, scale,
                               shift, count_out, C, st);
}

int sseg_bn_running_from_tmp(const float* tmp_mean, const float* tmp_var, const float* running_iter, float* running_mean,
                             float* running_var, int C, sseg_stream_t st) {
  SSEG_REQUIRE(tmp_mean && tmp_var && running_iter && running_mean && running_var && C > 0, "sseg_bn_running_from_tmp: null");
  count_launch(1);
  return check_cuda(launch_k(bn_running_from_tmp_kernel, dim3((C + 255) / 256), dim3(256), 0, (cudaStream_t)st, tmp_mean,
                             tmp_var, running_iter, running_mean, running_var, C),
                    "bn_running_from_tmp_kernel");
}

static int bn_bwd_peer_sum_impl(void* const* bases, int world, int rank, long part_off, long flag_off, long inbox_off,
                                const int* step, float* s1_tot, float* s2_tot, float* dbeta, float* dgamma,
                                const float* mean, const float* invstd, int s2_raw, int C, sseg_stream_t st) {
  PeerTable t;
  int rc = make_peer_table(&t, bases, world, rank, "sseg_bn_bwd_peer_sum");
  if (rc) return rc;
  SSEG_REQUIRE(step && s1_tot && s2_tot && dbeta && dgamma && C > 0, "sseg_bn_bwd_peer_sum: null");
  SSEG_REQUIRE(!s2_raw || (mean && invstd), "sseg_bn_bwd_peer_sum: raw partials need mean and invstd");
  count_launch(1);
  return check_cuda(launch_k(bn_bwd_peer_sum_kernel, dim3((C + 255) / 256), dim3(256), 0, (cudaStream_t)st, t, part_off,
                             flag_off, step, s1_tot, s2_tot, dbeta, dgamma, mean, invstd, s2_raw, C, inbox_off),
                    "bn_bwd_peer_sum_kernel");
}

int sseg_bn_bwd_peer_sum(void* const* bases, int world, int rank, long part_off, long flag_off, const int* step,
                         float* s1_tot, float* s2_tot, float* dbeta, float* dgamma, const float* mean, const float* invstd,
                         int s2_raw, int C, sseg_stream_t st) {
  return bn_bwd_peer_sum_impl(bases, world, rank, part_off, flag_off, -1, step, s1_tot, s2_tot, dbeta, dgamma, mean, invstd,
                              s2_raw, C, st);
}

int sseg_bn_bwd_peer_sum_ll(void* const* bases, int world, int rank, long part_off, long inbox_off, const int* step,
                            float* s1_tot, float* s2_tot, float* dbeta, float* dgamma, const float* mean,
                            const float* invstd, int s2_raw, int C, sseg_stream_t st) {
  SSEG_REQUIRE(inbox_off >= 0 && inbox_off % 2 == 0, "sseg_bn_bwd_peer_sum_ll: the inbox must be 8-byte aligned");
  return bn_bwd_peer_sum_impl(bases, world, rank, part_off, 0, inbox_off, step, s1_tot, s2_tot, dbeta, dgamma, mean, invstd,
                              s2_raw, C, st);
}

}  // extern "C"
