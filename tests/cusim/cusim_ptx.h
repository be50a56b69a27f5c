// cusim model of csrc/ptx.cuh: the same wrapper names, implemented on the CPU (see cusim.h). TEST INFRASTRUCTURE ONLY.
//
//   mbarrier        64-bit word in (simulated) shared memory = {phase, pending arrivals, expected arrivals, tx bytes};
//                   try_wait.parity P succeeds once the phase with parity P has completed.
//   TMA             cp.async.bulk.tensor.{2,4}d tile mode: box copy global -> shared, coordinates outside the tensor read
//                   zeros, CU_TENSOR_MAP_SWIZZLE_128B applied to the shared-memory address (16-byte chunk index XOR
//                   128-byte row index mod 8), then complete_tx(box bytes) on the mbarrier. Executed at issue.
//   tensor memory   128 lanes x 512 fp32 columns per CTA, bump-allocated by tcgen05.alloc.
//   tcgen05.mma     kind::f16 (bf16 x bf16 -> fp32) and kind::tf32, M = 128, cta_group::1, SWIZZLE_128B descriptors.
//                   K-major operand: row r, byte k is read from swz(start + (r % 8) * 128 + (r / 8) * SBO + k);
//                   MN-major operand (weight gradient): element (mn, k) from
//                   swz(start + (mn / 64) * LBO + (k / 8) * SBO + (k % 8) * 128 + (mn % 64) * 2).
//                   Issued MMAs are QUEUED and only executed by the
//                   tcgen05.commit that covers them (as on the device, results are not visible before the commit's
//                   mbarrier completes) -- a missing commit / wait shows up as a wrong result.
//   tcgen05.ld      32x32b.x32: thread i of warp w reads lane 32*(w%4)+i; the lane field of the address must name the
//                   warp's own lane quarter.
#pragma once
#include <stdint.h>

namespace cusim {
uint32_t smem_handle(const void* p);
void mbar_init(uint64_t* bar, uint32_t count);
void mbar_arrive(uint64_t* bar, uint32_t tx_expect);
void mbar_complete_tx(uint64_t* bar, uint32_t bytes);
void mbar_wait(uint64_t* bar, uint32_t parity);
bool mbar_test(uint64_t* bar, uint32_t parity);
void tma_load(void* smem_dst, const void* tmap, uint64_t* bar, int rank, const int* coords);
void tmem_alloc(uint32_t* smem_result, uint32_t ncols);
void tmem_dealloc(uint32_t taddr, uint32_t ncols);
void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate, int kind);
void umma_commit(uint64_t* bar);
void tmem_ld_32x32(uint32_t taddr, uint32_t* v);
}  // namespace cusim

namespace sseg {

inline uint32_t smem_u32(const void* p) { return ::cusim::smem_handle(p); }
inline bool elect_one() { return (::cusim::tl.linear_tid & 31) == 0; }

inline void pdl_wait() {}
inline void pdl_launch_dependents() {}
inline void pdl_sync() {}

#define SSEG_DYN_SMEM(name) uint8_t* name = ::cusim::dyn_smem()

inline void bar_sync_epilogue() { ::cusim::named_barrier(1, 128); }

inline void st_release_sys(int* p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline int ld_acquire_sys(const int* p) {
  ::cusim::spin_pause();
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
}
inline void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
inline unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  ::cusim::spin_pause();
  return __atomic_load_n(p, __ATOMIC_RELAXED);
}
inline unsigned int ld_acquire_gpu(const unsigned int* p) {
  ::cusim::spin_pause();
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
}

inline void mbar_init(uint64_t* bar, uint32_t count) { ::cusim::mbar_init(bar, count); }
inline void fence_barrier_init() {}
inline void fence_proxy_async() {}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { ::cusim::mbar_arrive(bar, bytes); }
inline void mbar_arrive(uint64_t* bar) { ::cusim::mbar_arrive(bar, 0); }
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) { return ::cusim::mbar_test(bar, parity); }
inline void mbar_wait(uint64_t* bar, uint32_t parity) { ::cusim::mbar_wait(bar, parity); }
// warp-collective wait: on hardware the lanes of a converged warp execute ONE try_wait; here they are independent threads, so
// nobody leaves before every lane has seen the phase complete (otherwise a slow lane could look two phases later)
inline void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  ::cusim::mbar_wait(bar, parity);
  ::cusim::syncwarp();
}
inline void mbar_wait_bounded(uint64_t* bar, uint32_t parity) { ::cusim::mbar_wait(bar, parity); }

inline void tma_prefetch_desc(const CUtensorMap*) {}
inline void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  const int c[2] = {c0, c1};
  ::cusim::tma_load(smem_dst, m, bar, 2, c);
}
inline void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  const int c[4] = {c0, c1, c2, c3};
  ::cusim::tma_load(smem_dst, m, bar, 4, c);
}

template <uint32_t NCOLS>
inline void tmem_alloc(uint32_t* smem_result) {
  static_assert(NCOLS >= 32 && NCOLS <= 512 && (NCOLS & (NCOLS - 1)) == 0, "tcgen05.alloc: 32..512 columns, power of 2");
  ::cusim::tmem_alloc(smem_result, NCOLS);
}
template <uint32_t NCOLS>
inline void tmem_dealloc(uint32_t taddr) {
  ::cusim::tmem_dealloc(taddr, NCOLS);
}
inline void tc_fence_before() {}
inline void tc_fence_after() {}

inline void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  ::cusim::umma(tmem_d, da, db, idesc, accumulate, /*kind f16*/ 0);
}
inline void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  ::cusim::umma(tmem_d, da, db, idesc, accumulate, /*kind tf32*/ 1);
}
inline void umma_commit(uint64_t* bar) { ::cusim::umma_commit(bar); }
inline void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) { ::cusim::tmem_ld_32x32(taddr, v); }
inline void tmem_ld_wait() {}

}  // namespace sseg
