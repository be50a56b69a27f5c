// cusim runtime: the executor (OS thread per CUDA thread), barriers, mbarrier / TMA / tensor-memory / tcgen05 models and
// the handful of CUDA runtime entry points the library's host code calls. TEST INFRASTRUCTURE ONLY (see cusim.h).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <ucontext.h>

#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "cusim.h"
#include "cusim_ptx.h"

namespace cusim {

thread_local ThreadCtx tl;

struct Abort {};

struct Bar {
  int count = 0;
  unsigned gen = 0;
};

struct PendingMma {
  uint32_t d;
  uint64_t da, db;
  uint32_t idesc, acc;
  int kind;
};

struct Cta {
  std::mutex mu;
  std::condition_variable cv;
  int nthreads = 0, alive = 0;
  Bar sync;
  Bar named[16];
  struct WarpSync {
    std::mutex mu;
    std::condition_variable cv;
    Bar bar;
    uint64_t slot[32];
  };
  std::vector<WarpSync> warps;
  uint8_t* smem = nullptr;
  size_t smem_bytes = 0;
  float* tmem = nullptr;  // [128][512]
  uint32_t tmem_next = 0;
  std::vector<PendingMma> pending;
  int index = 0;
  bool fresh = true;
  std::chrono::steady_clock::time_point launch_t0;
};

static std::atomic<bool> g_abort{false};
static std::mutex g_msg_mu;
static std::string g_abort_msg;
static std::mutex g_atomic_mu;
static cudaError_t g_last_error = cudaSuccess;

static double timeout_seconds() {
  const char* e = getenv("CUSIM_TIMEOUT");
  return e ? atof(e) : 60.0;
}

static void raise_abort(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  {
    std::lock_guard<std::mutex> lk(g_msg_mu);
    if (!g_abort.load()) g_abort_msg = buf;
  }
  g_abort.store(true);
  throw Abort();
}

#define CUSIM_CHECK(cond, ...)                 \
  do {                                         \
    if (!(cond)) raise_abort(__VA_ARGS__);     \
  } while (0)

// ---------------------------------------------------------------------------------------------------------- fibers
// Default executor: the CUDA threads of a CTA are ucontext fibers on ONE OS thread (one OS thread per resident CTA). A
// blocked fiber parks a predicate with the scheduler, which re-evaluates it without switching; when every fiber of a CTA is
// parked on a false predicate the CTA is deadlocked, which is reported at once (no timeout involved). Polling loops on
// global memory (grid barrier, peer flags) yield on every poll. CUSIM_EXECUTOR=threads selects the OS-thread-per-CUDA-thread
// executor below instead (needed under ThreadSanitizer / AddressSanitizer, which do not follow ucontext switches).
struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  ThreadCtx tctx;
  bool done = false;
  const std::function<bool()>* pred = nullptr;
  const char* what = nullptr;
  long detail = 0;
};
static thread_local Fiber* t_fiber = nullptr;       // the running fiber (null on plain OS threads)
static thread_local ucontext_t t_sched;             // the scheduler of this OS thread
static const size_t kFiberStack = 512 * 1024;

static bool use_fibers() {
  const char* e = getenv("CUSIM_EXECUTOR");
  return !(e != nullptr && strcmp(e, "threads") == 0);
}

static void fiber_switch_out() {
  Fiber* f = t_fiber;
  swapcontext(&f->ctx, &t_sched);
  if (g_abort.load()) throw Abort();
}

static void fiber_block(const std::function<bool()>& pred, const char* what, long detail) {
  Fiber* f = t_fiber;
  f->pred = &pred, f->what = what, f->detail = detail;
  fiber_switch_out();
}

template <typename Pred>
static void wait_until(Cta* c, std::unique_lock<std::mutex>& lk, Pred pred, const char* what, long detail,
                       std::condition_variable* cv = nullptr) {
  if (t_fiber != nullptr) {
    if (pred()) return;
    lk.unlock();  // never switch away with a lock held
    const std::function<bool()> fn = [&] { return pred(); };
    fiber_block(fn, what, detail);
    lk.lock();
    return;
  }
  const auto t0 = std::chrono::steady_clock::now();
  const double limit = timeout_seconds();
  if (cv == nullptr) cv = &c->cv;
  while (!pred()) {
    if (g_abort.load()) throw Abort();
    cv->wait_for(lk, std::chrono::milliseconds(20));
    if (pred()) break;
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (dt > limit) {
      lk.unlock();
      raise_abort("deadlock: CTA %d thread %d waited %.0f s on %s (%ld)", c->index, tl.linear_tid, dt, what, detail);
    }
  }
}

static void bar_arrive_wait(Cta* c, Bar& b, int expected, const char* what, long detail) {
  std::unique_lock<std::mutex> lk(c->mu);
  const unsigned g = b.gen;
  if (++b.count >= expected) {
    b.count = 0;
    b.gen++;
    c->cv.notify_all();
    return;
  }
  wait_until(c, lk, [&] { return b.gen != g; }, what, detail);
}

void syncthreads() {
  Cta* c = tl.cta;
  std::unique_lock<std::mutex> lk(c->mu);
  Bar& b = c->sync;
  const unsigned g = b.gen;
  if (++b.count >= c->alive) {
    b.count = 0;
    b.gen++;
    c->cv.notify_all();
    return;
  }
  wait_until(c, lk, [&] { return b.gen != g; }, "__syncthreads", b.count);
}

void named_barrier(int id, int count) { bar_arrive_wait(tl.cta, tl.cta->named[id & 15], count, "bar.sync id", id); }

void syncwarp() {  // own mutex / condition variable per warp: a shuffle wakes 31 threads, not the whole CTA
  Cta* c = tl.cta;
  const int w = tl.linear_tid >> 5;
  const int in_warp = std::min(32, c->nthreads - w * 32);
  Cta::WarpSync& ws = c->warps[w];
  std::unique_lock<std::mutex> lk(ws.mu);
  const unsigned g = ws.bar.gen;
  if (++ws.bar.count >= in_warp) {
    ws.bar.count = 0;
    ws.bar.gen++;
    ws.cv.notify_all();
    return;
  }
  wait_until(c, lk, [&] { return ws.bar.gen != g; }, "__syncwarp / shuffle of warp", w, &ws.cv);
}

uint64_t warp_exchange(uint64_t v, int src_lane) {
  Cta* c = tl.cta;
  const int w = tl.linear_tid >> 5, lane = tl.linear_tid & 31;
  c->warps[w].slot[lane] = v;
  syncwarp();
  const uint64_t r = c->warps[w].slot[src_lane];
  syncwarp();
  return r;
}

uint8_t* dyn_smem() { return tl.cta->smem; }
void atomic_lock() { g_atomic_mu.lock(); }
void atomic_unlock() { g_atomic_mu.unlock(); }

void trap() { raise_abort("__trap() executed by CTA %d thread %d", tl.cta ? tl.cta->index : -1, tl.linear_tid); }

void spin_pause() {
  if (g_abort.load()) throw Abort();
  static thread_local long spins = 0;
  if ((++spins & 1023) == 0 && tl.cta != nullptr) {  // watchdog of polling loops: measured from the start of the launch
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - tl.cta->launch_t0).count();
    if (dt > timeout_seconds())
      raise_abort("deadlock: CTA %d thread %d still polls a global flag %.0f s after the launch started", tl.cta->index, tl.linear_tid, dt);
  }
  if (t_fiber != nullptr) fiber_switch_out();  // let the other fibers of this CTA run, then poll again
  else sched_yield();
}

uint32_t smem_handle(const void* p) {
  Cta* c = tl.cta;
  const uint8_t* q = static_cast<const uint8_t*>(p);
  if (q >= c->smem && q < c->smem + c->smem_bytes) return static_cast<uint32_t>(q - c->smem);
  return 0x40000u | (static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p)) & 0xFFF0u);  // static shared: not addressable by descriptors
}

static uint8_t* smem_ptr(Cta* c, uint32_t handle, uint32_t bytes, const char* who) {
  CUSIM_CHECK((size_t)handle + bytes <= c->smem_bytes, "%s: shared-memory address 0x%x (+%u) outside the CTA's %zu bytes", who, handle,
              bytes, c->smem_bytes);
  return c->smem + handle;
}

// ---------------------------------------------------------------------------------------------------------- mbarrier
struct MBar {
  int32_t tx;
  int16_t pending;
  uint8_t expected;
  uint8_t phase;
};
static_assert(sizeof(MBar) == 8, "mbarrier word");

static void mbar_settle(Cta* c, MBar* b) {
  if (b->pending <= 0 && b->tx == 0) {
    b->phase ^= 1;
    b->pending = b->expected;
    c->cv.notify_all();
  }
}

void mbar_init(uint64_t* bar, uint32_t count) {
  Cta* c = tl.cta;
  std::lock_guard<std::mutex> lk(c->mu);
  MBar* b = reinterpret_cast<MBar*>(bar);
  b->tx = 0, b->pending = (int16_t)count, b->expected = (uint8_t)count, b->phase = 0;
}

static void mbar_arrive_on(Cta* c, uint64_t* bar, uint32_t tx_expect) {
  std::lock_guard<std::mutex> lk(c->mu);
  MBar* b = reinterpret_cast<MBar*>(bar);
  b->tx += (int32_t)tx_expect;
  b->pending -= 1;
  mbar_settle(c, b);
}
void mbar_arrive(uint64_t* bar, uint32_t tx_expect) { mbar_arrive_on(tl.cta, bar, tx_expect); }

void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
  Cta* c = tl.cta;
  std::lock_guard<std::mutex> lk(c->mu);
  MBar* b = reinterpret_cast<MBar*>(bar);
  b->tx -= (int32_t)bytes;
  mbar_settle(c, b);
}

bool mbar_test(uint64_t* bar, uint32_t parity) {
  Cta* c = tl.cta;
  bool ok;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    ok = reinterpret_cast<MBar*>(bar)->phase != (parity & 1);
  }
  if (!ok) spin_pause();  // a try_wait loop in kernel code must let the other threads run
  return ok;
}

void mbar_wait(uint64_t* bar, uint32_t parity) {
  Cta* c = tl.cta;
  std::unique_lock<std::mutex> lk(c->mu);
  MBar* b = reinterpret_cast<MBar*>(bar);
  wait_until(c, lk, [&] { return b->phase != (parity & 1); }, "mbarrier (smem offset) parity",
             (long)(reinterpret_cast<uint8_t*>(bar) - c->smem) * 10 + (parity & 1));
}

// ---------------------------------------------------------------------------------------------------------- TMA
struct SimTmap {
  uint64_t magic;
  uint8_t* ptr;
  uint32_t elem, rank;
  uint64_t dims[4];
  uint64_t strides[3];  // bytes, dims 1..3
  uint32_t box[4];
  uint32_t swizzle;
};
static_assert(sizeof(SimTmap) <= sizeof(CUtensorMap), "SimTmap must fit the opaque tensor map");
static const uint64_t kTmapMagic = 0x50414D5453554Cull;

static inline uintptr_t swz128(uintptr_t a) { return a ^ (((a >> 7) & 7) << 4); }

void tma_load(void* smem_dst, const void* tmap, uint64_t* bar, int rank, const int* coords) {
  Cta* c = tl.cta;
  SimTmap m;
  memcpy(&m, tmap, sizeof(m));
  CUSIM_CHECK(m.magic == kTmapMagic, "TMA: tensor map was not produced by cuTensorMapEncodeTiled");
  CUSIM_CHECK((int)m.rank == rank, "TMA: %dd load through a rank-%u tensor map", rank, m.rank);
  uint8_t* dst = static_cast<uint8_t*>(smem_dst);
  CUSIM_CHECK(dst >= c->smem && dst < c->smem + c->smem_bytes, "TMA: destination is not in dynamic shared memory");
  CUSIM_CHECK(((dst - c->smem) & 1023) == 0, "TMA: SWIZZLE_128B destination offset 0x%lx is not 1024-byte aligned", (long)(dst - c->smem));
  uint32_t box[4] = {1, 1, 1, 1};
  for (int d = 0; d < rank; ++d) box[d] = m.box[d];
  const size_t bytes = (size_t)box[0] * box[1] * box[2] * box[3] * m.elem;
  CUSIM_CHECK((size_t)(dst - c->smem) + bytes <= c->smem_bytes, "TMA: box of %zu bytes overruns shared memory", bytes);
  const size_t row_bytes = (size_t)box[0] * m.elem;
  size_t off = 0;
  for (uint32_t i3 = 0; i3 < box[3]; ++i3)
    for (uint32_t i2 = 0; i2 < box[2]; ++i2)
      for (uint32_t i1 = 0; i1 < box[1]; ++i1) {
        long g[4] = {0, 0, 0, 0};
        const uint32_t idx[4] = {0, i1, i2, i3};
        bool outer_ok = true;
        for (int d = 1; d < rank; ++d) {
          g[d] = (long)coords[d] + idx[d];
          if (g[d] < 0 || g[d] >= (long)m.dims[d]) outer_ok = false;
        }
        const uint8_t* src = m.ptr;
        if (outer_ok)
          for (int d = 1; d < rank; ++d) src += (size_t)g[d] * m.strides[d - 1];
        for (uint32_t i0 = 0; i0 < box[0]; ++i0, off += m.elem) {
          const long g0 = (long)coords[0] + i0;
          uint8_t* out = c->smem + swz128((uintptr_t)(dst - c->smem) + off);
          if (outer_ok && g0 >= 0 && g0 < (long)m.dims[0]) memcpy(out, src + (size_t)g0 * m.elem, m.elem);
          else memset(out, 0, m.elem);
        }
        (void)row_bytes;
      }
  mbar_complete_tx(bar, (uint32_t)bytes);
}

// ---------------------------------------------------------------------------------------------------------- tensor memory
void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  Cta* c = tl.cta;
  if ((tl.linear_tid & 31) != 0) return;  // .sync.aligned: the whole warp executes it, one lane does the bookkeeping
  std::unique_lock<std::mutex> lk(c->mu);
  if (c->tmem_next + ncols > 512) {
    lk.unlock();
    raise_abort("tcgen05.alloc: %u columns requested with %u already allocated (512 per SM)", ncols, c->tmem_next);
  }
  *smem_result = c->tmem_next;  // lane 0, column offset
  c->tmem_next += ncols;
}
void tmem_dealloc(uint32_t, uint32_t) {}

static inline float bf16_bits_to_float(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline float tf32_of(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;  // 10-bit mantissa (truncation)
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// Gather one operand into a dense [rows][K] float matrix.
static void gather_operand(Cta* c, uint64_t desc, int rows, int kelems, int esize, int mn_major, int kind, std::vector<float>& out) {
  const uint32_t start = (uint32_t)(desc & 0x3FFF) << 4;
  const uint32_t lbo = (uint32_t)((desc >> 16) & 0x3FFF) << 4;
  const uint32_t sbo = (uint32_t)((desc >> 32) & 0x3FFF) << 4;
  CUSIM_CHECK(((desc >> 61) & 7) == 2, "tcgen05.mma: smem descriptor layout %d, only SWIZZLE_128B (2) is modelled", (int)((desc >> 61) & 7));
  CUSIM_CHECK(((desc >> 46) & 3) == 1, "tcgen05.mma: smem descriptor version %d, Blackwell needs 1", (int)((desc >> 46) & 3));
  out.resize((size_t)rows * kelems);
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < kelems; ++k) {
      uint32_t addr;
      if (!mn_major) {
        // K-major: 8-row x 128-byte atoms, SBO between 8-row groups (LBO unused for a 32-byte K extent)
        addr = start + (uint32_t)(r & 7) * 128 + (uint32_t)(r >> 3) * sbo + (uint32_t)k * esize;
      } else {
        // MN-major: (64 MN elements = 128 B) x 8 K atoms; LBO between 64-element MN groups, SBO between 8-K groups
        CUSIM_CHECK(esize == 2, "tcgen05.mma: MN-major operands are modelled for 16-bit types only");
        addr = start + (uint32_t)(r >> 6) * lbo + (uint32_t)(k >> 3) * sbo + (uint32_t)(k & 7) * 128 + (uint32_t)(r & 63) * 2;
      }
      const uint8_t* p = smem_ptr(c, (uint32_t)swz128(addr), esize, "tcgen05.mma operand");
      float v;
      if (esize == 2) {
        uint16_t h;
        memcpy(&h, p, 2);
        v = bf16_bits_to_float(h);
      } else {
        memcpy(&v, p, 4);
        if (kind == 1) v = tf32_of(v);
      }
      out[(size_t)r * kelems + k] = v;
    }
}

static void execute_mma(Cta* c, const PendingMma& q) {
  const uint32_t i = q.idesc;
  const int c_fmt = (i >> 4) & 3, a_fmt = (i >> 7) & 7, b_fmt = (i >> 10) & 7;
  const int a_mn = (i >> 15) & 1, b_mn = (i >> 16) & 1;
  const int N = (int)((i >> 17) & 0x3F) << 3, M = (int)((i >> 24) & 0x1F) << 4;
  CUSIM_CHECK(c_fmt == 1, "tcgen05.mma: accumulator format %d, expected F32 (1)", c_fmt);
  CUSIM_CHECK(a_fmt == b_fmt, "tcgen05.mma: A format %d != B format %d", a_fmt, b_fmt);
  CUSIM_CHECK(q.kind == 0 ? a_fmt == 1 : a_fmt == 2, "tcgen05.mma kind %d with operand format %d", q.kind, a_fmt);
  CUSIM_CHECK(M == 128, "tcgen05.mma: M = %d, only the M = 128 cta_group::1 layout is modelled", M);
  CUSIM_CHECK(N >= 16 && N <= 256 && N % 16 == 0, "tcgen05.mma: N = %d is not legal for M = 128", N);
  const int esize = q.kind == 0 ? 2 : 4, K = 32 / esize;
  const uint32_t lane = q.d >> 16, col = q.d & 0xFFFF;
  CUSIM_CHECK(lane == 0, "tcgen05.mma: accumulator lane base %u (M = 128 uses all lanes, base 0)", lane);
  CUSIM_CHECK(col + N <= 512 && col + N <= c->tmem_next, "tcgen05.mma: accumulator columns [%u, %u) outside the allocation (%u columns)",
              col, col + N, c->tmem_next);
  static thread_local std::vector<float> A, B;
  gather_operand(c, q.da, M, K, esize, a_mn, q.kind, A);
  gather_operand(c, q.db, N, K, esize, b_mn, q.kind, B);
  for (int m = 0; m < M; ++m) {
    float* drow = c->tmem + (size_t)m * 512 + col;
    const float* a = &A[(size_t)m * K];
    for (int n = 0; n < N; ++n) {
      const float* b = &B[(size_t)n * K];
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc += a[k] * b[k];
      drow[n] = q.acc ? drow[n] + acc : acc;
    }
  }
}

void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate, int kind) {
  tl.cta->pending.push_back(PendingMma{tmem_d, da, db, idesc, accumulate, kind});  // only the single MMA thread issues
}

void umma_commit(uint64_t* bar) {
  Cta* c = tl.cta;
  for (const PendingMma& q : c->pending) execute_mma(c, q);
  c->pending.clear();
  mbar_arrive_on(c, bar, 0);
}

void tmem_ld_32x32(uint32_t taddr, uint32_t* v) {
  Cta* c = tl.cta;
  const uint32_t lane_base = taddr >> 16, col = taddr & 0xFFFF;
  const int warp = tl.linear_tid >> 5, lane = tl.linear_tid & 31;
  CUSIM_CHECK(lane_base == (uint32_t)(warp & 3) * 32, "tcgen05.ld: warp %d may only read lanes [%d, +32), address names lane %u", warp,
              (warp & 3) * 32, lane_base);
  CUSIM_CHECK(col + 32 <= c->tmem_next, "tcgen05.ld: columns [%u, %u) outside the allocation (%u columns)", col, col + 32, c->tmem_next);
  memcpy(v, c->tmem + (size_t)(lane_base + lane) * 512 + col, 32 * sizeof(float));
}

// ---------------------------------------------------------------------------------------------------------- executor
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static void cta_thread_exit(Cta* c) {
  std::lock_guard<std::mutex> lk(c->mu);
  c->alive -= 1;
  Bar& b = c->sync;
  if (b.count > 0 && b.count >= c->alive) {  // exited threads no longer take part in __syncthreads
    b.count = 0;
    b.gen++;
  }
  c->cv.notify_all();
}

static Cta* new_cta(int nthreads, size_t smem) {
  Cta* c = new Cta();
  c->nthreads = nthreads;
  c->warps = std::vector<Cta::WarpSync>((nthreads + 31) / 32);
  c->smem_bytes = smem;
  if (posix_memalign(reinterpret_cast<void**>(&c->smem), 1024, smem + 1024) != 0) abort();
  c->tmem = static_cast<float*>(malloc(128 * 512 * sizeof(float)));
  return c;
}

static void reset_cta(Cta* c, long index) {  // a fresh CTA: garbage in shared and tensor memory, no barrier state
  c->index = (int)index;
  c->alive = c->nthreads;
  c->sync = Bar();
  for (Bar& b : c->named) b = Bar();
  for (auto& w : c->warps) w.bar = Bar();
  memset(c->smem, 0xCD, c->smem_bytes + 1024);
  if (c->tmem_next != 0 || c->fresh)
    for (int i = 0; i < 128 * 512; ++i) c->tmem[i] = NAN;
  c->fresh = false;
  c->tmem_next = 0;
  c->pending.clear();
}

static void free_cta(Cta* c) {
  free(c->smem);
  free(c->tmem);
  delete c;
}

// A reusable rendezvous of the worker threads of one CTA slot (between consecutive CTAs of an ordinary launch).
struct PoolBarrier {
  std::mutex mu;
  std::condition_variable cv;
  int count = 0, n = 0;
  unsigned gen = 0;
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    const unsigned g = gen;
    if (++count == n) {
      count = 0;
      gen++;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
  }
};

static void fiber_entry() {
  Fiber* f = t_fiber;
  const std::function<void()>* body = reinterpret_cast<const std::function<void()>*>(f->detail);
  try {
    (*body)();
  } catch (const Abort&) {
  }
  f->done = true;
  swapcontext(&f->ctx, &t_sched);
}

// All CTAs j = first, first + stride, ... of one launch, on the calling OS thread.
static void run_ctas_as_fibers(Cta* c, long first, long stride, long nctas, dim3 grid, dim3 block, const std::function<void()>& body) {
  const int n = c->nthreads;
  std::vector<Fiber> fibers(n);
  for (Fiber& f : fibers) {
    f.stack = static_cast<char*>(mmap(nullptr, kFiberStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
    if (f.stack == MAP_FAILED) abort();
  }
  for (long j = first; j < nctas && !g_abort.load(); j += stride) {
    reset_cta(c, j);
    for (int t = 0; t < n; ++t) {
      Fiber& f = fibers[t];
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack, f.ctx.uc_stack.ss_size = kFiberStack, f.ctx.uc_link = nullptr;
      makecontext(&f.ctx, fiber_entry, 0);
      f.done = false, f.pred = nullptr;
      f.detail = reinterpret_cast<long>(&body);  // read once by fiber_entry
      f.tctx.cta = c, f.tctx.linear_tid = t, f.tctx.bdim = block, f.tctx.gdim = grid;
      f.tctx.tid = make_uint3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
      f.tctx.bid = make_uint3((unsigned)(j % grid.x), (unsigned)((j / grid.x) % grid.y), (unsigned)(j / ((long)grid.x * grid.y)));
    }
    int remaining = n;
    while (remaining > 0) {
      bool ran = false;
      for (int t = 0; t < n; ++t) {
        Fiber& f = fibers[t];
        if (f.done) continue;
        if (f.pred != nullptr) {
          if (!g_abort.load() && !(*f.pred)()) continue;
          f.pred = nullptr;
        }
        ran = true;
        tl = f.tctx;
        t_fiber = &f;
        swapcontext(&t_sched, &f.ctx);
        t_fiber = nullptr;
        if (f.done) {
          cta_thread_exit(c);
          --remaining;
        }
      }
      if (!ran) {  // every live fiber is parked on a false predicate: nothing inside this CTA can ever release them
        std::string msg = "deadlock in CTA " + std::to_string(c->index) + ": every live thread is blocked;";
        int shown = 0;
        for (int t = 0; t < n && shown < 6; ++t)
          if (!fibers[t].done && (t == 0 || fibers[t].what != fibers[t - 1].what || fibers[t].detail != fibers[t - 1].detail)) {
            msg += " thread " + std::to_string(t) + " waits on " + fibers[t].what + " (" + std::to_string(fibers[t].detail) + ");";
            ++shown;
          }
        {
          std::lock_guard<std::mutex> lk(g_msg_mu);
          if (!g_abort.load()) g_abort_msg = msg;
        }
        g_abort.store(true);  // the parked fibers are resumed once more and unwind with Abort
      }
    }
    tl.cta = nullptr;
  }
  for (Fiber& f : fibers) munmap(f.stack, kFiberStack);
}

// Ordinary launch: `parallel_ctas` CTA slots (1 for translation units with static __shared__ variables, which are
// process-wide statics here), each slot = `nthreads` worker threads walking its share of the CTAs. Cooperative launch:
// every CTA gets its own slot, all run concurrently.
int launch(dim3 grid, dim3 block, size_t smem, bool cooperative, const std::function<void()>& body, int parallel_ctas) {
  const int nthreads = (int)(block.x * block.y * block.z);
  const long nctas = (long)grid.x * grid.y * grid.z;
  if (nthreads <= 0 || nthreads > 1024 || nctas <= 0 || smem > 232448) {
    g_last_error = cudaErrorInvalidConfiguration;
    return (int)g_last_error;
  }
  if (cooperative && nctas > env_int("CUSIM_SMS", 4)) {
    g_last_error = cudaErrorCooperativeLaunchTooLarge;
    return (int)g_last_error;
  }
  g_abort.store(false);
  const long slots = cooperative ? nctas : std::max(1L, std::min<long>(nctas, parallel_ctas));  // CTAs resident at the same time
  std::vector<Cta*> ctas;
  std::vector<PoolBarrier> gates(slots);
  for (long s = 0; s < slots; ++s) {
    ctas.push_back(new_cta(nthreads, smem));
    ctas.back()->launch_t0 = std::chrono::steady_clock::now();
    gates[s].n = nthreads;
  }
  std::vector<std::thread> threads;
  if (use_fibers()) {
    if (slots == 1) {
      run_ctas_as_fibers(ctas[0], 0, 1, nctas, grid, block, body);
    } else {
      for (long s = 0; s < slots; ++s)
        threads.emplace_back([=, &body, &ctas] { run_ctas_as_fibers(ctas[s], s, slots, nctas, grid, block, body); });
    }
  } else {
  threads.reserve(slots * nthreads);
  for (long s = 0; s < slots; ++s)
    for (int t = 0; t < nthreads; ++t)
      threads.emplace_back([=, &body, &ctas, &gates] {
        Cta* c = ctas[s];
        tl.linear_tid = t;
        tl.bdim = block, tl.gdim = grid;
        tl.tid = make_uint3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
        for (long j = s; j < nctas; j += slots) {
          if (t == 0) reset_cta(c, j);
          gates[s].wait();
          if (!g_abort.load()) {
            tl.cta = c;
            tl.bid = make_uint3((unsigned)(j % grid.x), (unsigned)((j / grid.x) % grid.y), (unsigned)(j / ((long)grid.x * grid.y)));
            try {
              body();
            } catch (const Abort&) {
            }
            cta_thread_exit(c);
            tl.cta = nullptr;
          }
          gates[s].wait();
        }
      });
  }
  for (auto& th : threads) th.join();
  for (Cta* c : ctas) free_cta(c);
  if (g_abort.load()) {
    fprintf(stderr, "[cusim] launch aborted: %s\n", g_abort_msg.c_str());
    g_last_error = cudaErrorLaunchFailure;
    return (int)g_last_error;
  }
  return (int)cudaSuccess;
}

const char* abort_message() { return g_abort_msg.c_str(); }

}  // namespace cusim

// ============================================================================================================ CUDA API stubs
static CUresult sim_encode_tiled(CUtensorMap* out, CUtensorMapDataType dt, cuuint32_t rank, void* ptr, const cuuint64_t* gdim,
                                 const cuuint64_t* gstr, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il,
                                 CUtensorMapSwizzle sw, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  using namespace cusim;
  const uint32_t elem = dt == CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 ? 2 : (dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 0);
  if (elem == 0 || rank < 1 || rank > 4 || il != CU_TENSOR_MAP_INTERLEAVE_NONE || sw != CU_TENSOR_MAP_SWIZZLE_128B)
    return CUDA_ERROR_INVALID_VALUE;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return CUDA_ERROR_INVALID_VALUE;
  SimTmap m;
  memset(&m, 0, sizeof(m));
  m.magic = kTmapMagic, m.ptr = static_cast<uint8_t*>(ptr), m.elem = elem, m.rank = rank, m.swizzle = 128;
  for (uint32_t d = 0; d < rank; ++d) {
    if (gdim[d] == 0 || gdim[d] > (1ull << 32) || box[d] == 0 || box[d] > 256 || estr[d] != 1) return CUDA_ERROR_INVALID_VALUE;
    m.dims[d] = gdim[d], m.box[d] = box[d];
  }
  for (uint32_t d = 0; d + 1 < rank; ++d) {
    if (gstr[d] % 16 != 0 || gstr[d] >= (1ull << 40)) return CUDA_ERROR_INVALID_VALUE;
    m.strides[d] = gstr[d];
  }
  if ((uint64_t)box[0] * elem > 128) return CUDA_ERROR_INVALID_VALUE;  // SWIZZLE_128B: inner box extent <= 128 bytes
  memset(out, 0, sizeof(*out));
  memcpy(out, &m, sizeof(m));
  return CUDA_SUCCESS;
}

extern "C" {

cudaError_t cudaGetLastError(void) {
  const cudaError_t e = cusim::g_last_error;
  cusim::g_last_error = cudaSuccess;
  return e;
}
cudaError_t cudaPeekAtLastError(void) { return cusim::g_last_error; }
const char* cudaGetErrorString(cudaError_t e) {
  if (e == cudaSuccess) return "no error";
  if (e == cudaErrorLaunchFailure) return cusim::g_abort_msg.c_str();
  if (e == cudaErrorCooperativeLaunchTooLarge) return "too many blocks in cooperative launch";
  return "cusim error";
}
cudaError_t cudaGetDevice(int* dev) {
  *dev = 0;
  return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int* value, enum cudaDeviceAttr attr, int) {
  if (attr == cudaDevAttrMultiProcessorCount) *value = cusim::env_int("CUSIM_SMS", 4);
  else *value = 0;
  return cudaSuccess;
}
cudaError_t cudaFuncSetAttribute(const void*, enum cudaFuncAttribute, int) { return cudaSuccess; }
// "Device" allocations made through cudaMalloc (only the peer arenas of csrc/peer.cu) live in POSIX shared memory, so that
// cudaIpcGetMemHandle / cudaIpcOpenMemHandle really map one rank's arena into another PROCESS (two-rank runs under
// torch.distributed.run poll each other's flags through it). The 64-byte handle carries the object's name.
struct ShmAlloc {
  void* ptr;
  size_t bytes;
  std::string name;
  bool owner;
};
static std::mutex g_shm_mu;
static std::vector<ShmAlloc> g_shm;

static void shm_cleanup_at_exit() {
  for (const ShmAlloc& a : g_shm)
    if (a.owner) shm_unlink(a.name.c_str());
}

cudaError_t cudaMalloc(void** p, size_t bytes) {
  static std::atomic<int> counter{0};
  static std::once_flag once;
  std::call_once(once, [] { atexit(shm_cleanup_at_exit); });
  if (bytes == 0) bytes = 256;
  char name[64];
  snprintf(name, sizeof(name), "/cusim_%d_%d", (int)getpid(), counter.fetch_add(1));
  const int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) return cudaErrorMemoryAllocation;
  if (ftruncate(fd, (off_t)bytes) != 0) {
    close(fd);
    shm_unlink(name);
    return cudaErrorMemoryAllocation;
  }
  void* q = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (q == MAP_FAILED) {
    shm_unlink(name);
    return cudaErrorMemoryAllocation;
  }
  std::lock_guard<std::mutex> lk(g_shm_mu);
  g_shm.push_back(ShmAlloc{q, bytes, name, true});
  *p = q;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
  std::lock_guard<std::mutex> lk(g_shm_mu);
  for (size_t i = 0; i < g_shm.size(); ++i)
    if (g_shm[i].ptr == p && g_shm[i].owner) {
      munmap(p, g_shm[i].bytes);
      shm_unlink(g_shm[i].name.c_str());
      g_shm.erase(g_shm.begin() + i);
      return cudaSuccess;
    }
  return cudaErrorInvalidValue;
}
cudaError_t cudaMemset(void* p, int v, size_t n) {
  memset(p, v, n);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
  memset(p, v, n);
  return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  std::lock_guard<std::mutex> lk(g_shm_mu);
  for (const ShmAlloc& a : g_shm)
    if (a.ptr == p && a.owner) {
      memset(h, 0, sizeof(*h));
      memcpy(h, a.name.c_str(), a.name.size() + 1);
      return cudaSuccess;
    }
  return cudaErrorInvalidValue;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned int) {
  char name[64];
  memcpy(name, &h, sizeof(name));
  name[63] = 0;
  std::lock_guard<std::mutex> lk(g_shm_mu);
  for (const ShmAlloc& a : g_shm)
    if (a.name == name) {  // same process (or opened before): the existing mapping
      *p = a.ptr;
      return cudaSuccess;
    }
  const int fd = shm_open(name, O_RDWR, 0600);
  if (fd < 0) return cudaErrorInvalidValue;
  struct stat st;
  if (fstat(fd, &st) != 0) {
    close(fd);
    return cudaErrorInvalidValue;
  }
  void* q = mmap(nullptr, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (q == MAP_FAILED) return cudaErrorInvalidValue;
  g_shm.push_back(ShmAlloc{q, (size_t)st.st_size, name, false});
  *p = q;
  return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) {
  std::lock_guard<std::mutex> lk(g_shm_mu);
  for (size_t i = 0; i < g_shm.size(); ++i)
    if (g_shm[i].ptr == p && !g_shm[i].owner) {
      munmap(p, g_shm[i].bytes);
      g_shm.erase(g_shm.begin() + i);
      break;
    }
  return cudaSuccess;
}

cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long, enum cudaDriverEntryPointQueryResult* status) {
  *fn = nullptr;
  if (strcmp(symbol, "cuTensorMapEncodeTiled") == 0) *fn = reinterpret_cast<void*>(&sim_encode_tiled);
  if (status) *status = *fn ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
  return cudaSuccess;
}

const char* cusim_abort_message(void) { return cusim::abort_message(); }

}  // extern "C"
