/*
 * sseg_b200.h — C ABI of libsseg_b200.so, the B200 (sm_100a) kernels underneath the
 * mit_semseg model API (ModelBuilder / SegmentationModule / SynchronizedBatchNorm2d).
 *
 * The reference (CSAILVision/semantic-segmentation-pytorch) has no FFI: its "operators" are
 * torch.nn library calls.  Each entry point below names the reference call site it replaces
 * (path:line relative to the reference repo).  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - the caller owns all memory (outputs and workspaces are allocated by the caller);
 *   - every call is asynchronous on `stream` and performs no hidden synchronisation;
 *   - return 0 on success, negative SSEG_ERR_* otherwise; sseg_last_error() gives the message
 *     (thread-local); no exception crosses this boundary;
 *   - activations are NHWC ("channels-last"), bf16 unless stated; `ld` is the element stride
 *     between consecutive pixels (>= c), so a tensor may be a channel slice of a wider buffer;
 *   - the library is re-entrant and takes the device from the current CUDA context of the
 *     calling thread (one process per GPU under torch.distributed; replica threads under
 *     nn.DataParallel each set their own device — see SURVEY.md §8(b) "threading").
 */
#ifndef SSEG_B200_H_
#define SSEG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSEG_OK 0
#define SSEG_ERR_ARG (-1)
#define SSEG_ERR_CUDA (-2)
#define SSEG_ERR_UNSUPPORTED (-3)

#define SSEG_MAX_SRCS 5
#define SSEG_MAX_TAPS 9

typedef void* sseg_stream_t; /* cudaStream_t */

/* NHWC activation view. */
typedef struct {
  void* ptr; /* device pointer to element (n=0,h=0,w=0,c=0) */
  int n, h, w, c;
  int ld; /* elements between consecutive pixels */
} sseg_act_t;

/* ---- library ---------------------------------------------------------------------------- */
const char* sseg_last_error(void);
int sseg_version(void);
/* Number of kernel launches issued by this library on the calling thread since the last reset
 * (bench.py's "gpu_launches"). */
long sseg_launch_count(void);
void sseg_launch_count_reset(void);

/* ---- convolution as implicit GEMM on tcgen05 tensor cores ------------------------------- */
/*
 * Geometry of the input side of a stride-1 convolution.
 *   tap t reads X[n, h + tap_dh[t], w + tap_dw[t], :] ; out-of-range pixels read zeros (TMA out-of-bounds
 *   fill = the zero padding of nn.Conv2d).
 *   tap_src[t] == -1 : X is the virtual channel concatenation of all `nsrc` sources (same n,h,w; every c a
 *                      multiple of 64), so torch.cat (models/models.py:476) is never materialised.
 *   tap_src[t] == k  : tap t reads source k only (all sources then have equal c). This is how a stride-2
 *                      convolution is expressed over the 4 space-to-depth parity planes of its input.
 *   tap_koff[t]      : element offset of tap t's K segment inside one weight row.
 */
typedef struct {
  int nsrc;
  sseg_act_t srcs[SSEG_MAX_SRCS];
  int ntaps;
  int tap_dh[SSEG_MAX_TAPS];
  int tap_dw[SSEG_MAX_TAPS];
  int tap_src[SSEG_MAX_TAPS];
  int tap_koff[SSEG_MAX_TAPS];
} sseg_conv_geom_t;

/*
 * out[n,h,w,co] = sum_t sum_ci X_t[n, h+dh_t, w+dw_t, ci] * W[co][tap_koff[t] + ci]   (+bias[co]) (+addend)
 *
 * Replaces: nn.Conv2d forward for every 1x1 / 3x3 (dilation 1,2,4) convolution of the path
 *   (models/resnet.py:18-21,61-66,130-131; models/models.py:160-167,449,454-462) and, with the
 *   transposed/flipped weight, the data-gradient of the same convolutions (autograd of those sites).
 *
 * w_bf16   : [cout][w_ld] bf16 rows (K-major)
 * out      : NHWC, out_f32 ? float : bf16, pixel stride ld_out; columns [0, n_store) are written
 *            (n_store multiple of 8, cout <= n_store <= ld_out; columns >= cout receive 0).
 * bias     : optional float[cout]
 * addend   : optional bf16 NHWC tensor with pixel stride ld_addend added before the store
 * stat_sum / stat_sqsum : optional float[cout]; per-channel sum and sum of squares of the fp32
 *            results are ATOMICALLY ADDED (caller zeroes them) - the first half of
 *            SynchronizedBatchNorm2d.forward (lib/nn/modules/batchnorm.py:68-70).
 */
int sseg_conv_igemm(const sseg_conv_geom_t* geom, const void* w_bf16, long w_ld, int cout, void* out, int out_f32,
                    int ld_out, int n_store, const float* bias, const void* addend, int ld_addend, float* stat_sum,
                    float* stat_sqsum, sseg_stream_t stream);

/*
 * Weight gradient of the same convolution (autograd of nn.Conv2d w.r.t. weight):
 *   dw[co][tap_koff[t] + ci] += sum_{n,h,w} dy[n,h,w,co] * X_t[n, h+dh_t, w+dw_t, ci]
 * GEMM with K = pixels (both operands MN-major in shared memory), split over pixels across CTAs and
 * accumulated with fp32 atomics: the caller zeroes dw (float [cout][dw_ld]). dy may carry zero padding channels
 * beyond cout (dy->c >= cout, multiple of 8); rows >= cout are not written.
 */
int sseg_conv_wgrad(const sseg_conv_geom_t* geom, const sseg_act_t* dy, int cout, float* dw, long dw_ld,
                    sseg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SSEG_B200_H_ */
