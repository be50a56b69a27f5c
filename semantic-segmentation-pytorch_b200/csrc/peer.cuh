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

// ---- push protocol ("LL": every 8-byte message carries its own flag). Rank r owns, in EVERY rank's arena, the slots
// inbox[r * n + i], i < n, of a layer's inbox; a message is {fp32 value, int32 step}. The sender writes its n values into
// all peers' arenas with plain 64-bit stores (fire and forget: no fence, no separate flag); the receiver polls its LOCAL
// copies until their tag equals the step. One NVLink one-way latency per exchange instead of flag + fence + a remote
// load round trip. Tags are step numbers (monotonic, never reset): a slot is rewritten one step later, after the reader
// has long moved on (it has itself pushed every later layer of the step, which the writer had to receive first).
__device__ __forceinline__ void ll_push(const PeerTable& pt, long inbox_off, int n, int i, float v, int step) {
  const unsigned long long msg =
      (static_cast<unsigned long long>(static_cast<unsigned int>(step)) << 32) | static_cast<unsigned long long>(__float_as_uint(v));
#pragma unroll
  for (int r = 0; r < SSEG_MAX_PEERS; ++r)
    if (r < pt.world && r != pt.rank)
      st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(pt.base[r] + inbox_off) + (long)pt.rank * n + i, msg);
}
// Sum over ranks of slot i (this rank's own value passed in `mine`), added in rank order: bit-identical on all ranks.
// Bounded wait: a peer that never sends (dead / diverged rank) ends the launch with a trap after ~10 s of SM clocks.
__device__ __forceinline__ float ll_pool(const PeerTable& pt, long inbox_off, int n, int i, float mine, int step) {
  const unsigned long long* box = reinterpret_cast<const unsigned long long*>(pt.base[pt.rank] + inbox_off);
  float v[SSEG_MAX_PEERS];
  const long long t0 = clock64();
#pragma unroll
  for (int r = 0; r < SSEG_MAX_PEERS; ++r) {
    v[r] = 0.f;
    if (r < pt.world) {
      if (r == pt.rank) {
        v[r] = mine;
      } else {
        unsigned long long m = ld_relaxed_sys_u64(box + (long)r * n + i);
        while (static_cast<int>(m >> 32) != step) {
          if (clock64() - t0 > 20000000000ll) __trap();
          m = ld_relaxed_sys_u64(box + (long)r * n + i);
        }
        v[r] = __uint_as_float(static_cast<unsigned int>(m));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < SSEG_MAX_PEERS; ++r) s += v[r];
  return s;
}

int make_peer_table(PeerTable* t, void* const* bases, int world, int rank, const char* who);

}  // namespace sseg
