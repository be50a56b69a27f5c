// Library runtime: error reporting, launch accounting, TMA tensor-map factory + cache.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>

#include "common.h"

namespace sseg {

static thread_local char g_err[512] = "";
static thread_local long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %s (%d) at %s", cudaGetErrorString(e), (int)e, what);
  // the runtime also keeps a failed launch as its "last error": consume it, or the next entry point that checks
  // cudaGetLastError() after its own (successful) launch would report this failure a second time under its own name
  (void)cudaGetLastError();
  return SSEG_ERR_CUDA;
}

void count_launch(int n) { g_launches += n; }

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("SSEG_PDL");
    g_pdl = (e != nullptr && e[0] == '0') ? 0 : 1;  // on by default; SSEG_PDL=0 turns it off
  }
  return g_pdl == 1;
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  int elem_bytes, rank;
  uint64_t dims[4];
  uint64_t strides[3];
  uint32_t box[4];
  bool operator<(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) < 0; }
};

static std::mutex g_tmap_mu;
static std::map<TmapKey, CUtensorMap>& tmap_cache() {
  static std::map<TmapKey, CUtensorMap> c;
  return c;
}

static int encode(CUtensorMap* out, const void* ptr, int elem_bytes, int rank, const uint64_t* dims,
                  const uint64_t* strides, const uint32_t* box) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr;
  key.elem_bytes = elem_bytes;
  key.rank = rank;
  for (int i = 0; i < rank; ++i) key.dims[i] = dims[i], key.box[i] = box[i];
  for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides[i];
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = tmap_cache().find(key);
    if (it != tmap_cache().end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return SSEG_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) {
    set_error("TMA base pointer %p not 16-byte aligned", ptr);
    return SSEG_ERR_ARG;
  }
  cuuint64_t gdim[4];
  cuuint64_t gstr[3];
  cuuint32_t bdim[4], estr[4];
  for (int i = 0; i < rank; ++i) gdim[i] = dims[i], bdim[i] = box[i], estr[i] = 1;
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides[i];
    if (strides[i] % 16 != 0) {
      set_error("TMA global stride %llu (dim %d) not a multiple of 16 bytes", (unsigned long long)strides[i], i + 1);
      return SSEG_ERR_ARG;
    }
  }
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u", (int)r, rank,
              (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
              (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0), bdim[0],
              rank > 1 ? bdim[1] : 0, rank > 2 ? bdim[2] : 0, rank > 3 ? bdim[3] : 0);
    return SSEG_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  tmap_cache()[key] = *out;
  return 0;
}

int get_tmap_act(CUtensorMap* out, const void* ptr, int elem_bytes, int n, int h, int w, int c, long ld,
                 long row_stride, long img_stride, int box_c, int box_w, int box_h) {
  uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
  uint64_t strides[3] = {(uint64_t)ld * elem_bytes, (uint64_t)row_stride * elem_bytes, (uint64_t)img_stride * elem_bytes};
  uint32_t box[4] = {(uint32_t)box_c, (uint32_t)box_w, (uint32_t)box_h, 1u};
  return encode(out, ptr, elem_bytes, 4, dims, strides, box);
}

int get_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, long rows, long cols, long ld, int box_cols,
                int box_rows) {
  uint64_t dims[4] = {(uint64_t)cols, (uint64_t)rows, 0, 0};
  uint64_t strides[3] = {(uint64_t)ld * elem_bytes, 0, 0};
  uint32_t box[4] = {(uint32_t)box_cols, (uint32_t)box_rows, 0, 0};
  return encode(out, ptr, elem_bytes, 2, dims, strides, box);
}

}  // namespace sseg

extern "C" {
const char* sseg_last_error(void) { return sseg::g_err; }
int sseg_version(void) { return 100; }
long sseg_launch_count(void) { return sseg::g_launches; }
void sseg_launch_count_reset(void) { sseg::g_launches = 0; }
void sseg_set_pdl(int enable) { sseg::g_pdl = enable ? 1 : 0; }
int sseg_get_pdl(void) { return sseg::pdl_enabled() ? 1 : 0; }
}
