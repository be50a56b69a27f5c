// Device-side helpers of the peer-memory SyncBN exchange (csrc/peer.cu), shared with the kernels that fold the exchange into
// their own prologue (csrc/elementwise.cu: sseg_bn_bwd_apply_peer).
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace sseg {

struct PeerTable {
  float* base[SSEG_MAX_PEERS];
  int world, rank;
};

// All threads of every block call this. Block 0 publishes; every block waits for all peers. The wait is bounded: a peer
// that never publishes (a dead or diverged rank) ends the launch with a trap after ~10 s of SM clocks instead of hanging
// the job for ever (the reference's host-side rendezvous, comm.py:113, blocks without a timeout).
__device__ __forceinline__ void peer_handshake(const PeerTable& pt, long flag_off, int step, bool publisher) {
  if (publisher && threadIdx.x < pt.world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<int*>(pt.base[threadIdx.x]) + flag_off + pt.rank, step);
  }
  if (threadIdx.x < pt.world) {
    const int* mine = reinterpret_cast<const int*>(pt.base[pt.rank]) + flag_off + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < step) {
      __nanosleep(32);
      if (clock64() - t0 > 20000000000ll) __trap();
    }
  }
  __syncthreads();
}

// Sum of one value per rank, read straight out of the peers' arenas. All loads are issued before the first use (one NVLink
// round trip, not `world` of them: with a data-dependent loop the eight loads of an 8-GPU job serialise, ~1 us each) and
// added in rank order, so every rank computes bit-identical totals.
__device__ __forceinline__ float peer_sum(const PeerTable& pt, long off) {
  float v[SSEG_MAX_PEERS];
#pragma unroll
  for (int r = 0; r < SSEG_MAX_PEERS; ++r) v[r] = r < pt.world ? __ldcv(pt.base[r] + off) : 0.f;
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < SSEG_MAX_PEERS; ++r) s += v[r];
  return s;
}
__device__ __forceinline__ void peer_sum2(const PeerTable& pt, long off_a, long off_b, float* a, float* b) {
  float va[SSEG_MAX_PEERS], vb[SSEG_MAX_PEERS];
#pragma unroll
  for (int r = 0; r < SSEG_MAX_PEERS; ++r) {
    va[r] = r < pt.world ? __ldcv(pt.base[r] + off_a) : 0.f;
    vb[r] = r < pt.world ? __ldcv(pt.base[r] + off_b) : 0.f;
  }
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int r = 0; r < SSEG_MAX_PEERS; ++r) sa += va[r], sb += vb[r];
  *a = sa, *b = sb;
}

int make_peer_table(PeerTable* t, void* const* bases, int world, int rank, const char* who);

}  // namespace sseg
