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
#define SSEG_MAX_PEERS 8
#define SSEG_MAX_SUM_TERMS 4

typedef void* sseg_stream_t; /* cudaStream_t */

/* NHWC activation view: element (n,h,w,c) lives at ptr + n*img_stride + h*row_stride + w*ld + c (in elements).
 * Dense tensors have row_stride = w*ld and img_stride = h*row_stride; a strided view (e.g. the parity plane
 * x[:, 1::2, 0::2, :] of a stride-2 convolution's input, or a channel slice of a wider buffer) is legal wherever
 * a view is accepted. */
typedef struct {
  void* ptr;
  int n, h, w, c;
  int ld;          /* elements between consecutive pixels of a row */
  long row_stride; /* elements between consecutive rows */
  long img_stride; /* elements between consecutive images */
} sseg_act_t;

/* ---- library ---------------------------------------------------------------------------- */
const char* sseg_last_error(void);
int sseg_version(void);
/* Number of kernel launches issued by this library on the calling thread since the last reset
 * (bench.py's "gpu_launches"). */
long sseg_launch_count(void);
void sseg_launch_count_reset(void);
/* Programmatic dependent launch: when enabled every kernel is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization, so the prologue of kernel N+1 (barrier init, TMEM allocation,
 * descriptor prefetch, coefficient loads) overlaps the tail of kernel N. Default: on (SSEG_PDL=0 disables). */
void sseg_set_pdl(int enable);
int sseg_get_pdl(void);

/* ---- convolution as implicit GEMM on tcgen05 tensor cores ------------------------------- */
/*
 * Geometry of the input side of a stride-1 convolution.
 *   tap t reads X[n, h + tap_dh[t], w + tap_dw[t], :] ; out-of-range pixels read zeros (TMA out-of-bounds
 *   fill = the zero padding of nn.Conv2d).
 *   tap_src[t] == -1 : X is the virtual channel concatenation of all `nsrc` sources (same n,h,w; every c a
 *                      multiple of 4 with a pixel stride `ld` multiple of 8 - HRNet's 48/96-channel branches and C1's 180-channel
 *                      hidden layer run as zero-filled partial 64-blocks), so
 *                      torch.cat (models/models.py:476, models/hrnet.py:434) is never materialised.
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
 * out      : NHWC view (same n,h,w as the sources), out_f32 ? float : bf16; channels [0, out->c) are written
 *            (out->c multiple of 8, cout <= out->c <= out->ld; channels >= cout receive 0).
 * bias     : optional float[cout]
 * addend   : optional bf16 NHWC view (c >= out->c) added before the store (may alias out)
 * stat_sum / stat_sqsum : optional float[cout] (bf16 outputs only); per-channel sum and sum of squares of the
 *            results AS STORED (bf16-rounded, accumulated in fp32) are ATOMICALLY ADDED (caller zeroes them) - the
 *            first half of SynchronizedBatchNorm2d.forward (lib/nn/modules/batchnorm.py:68-70).
 */
int sseg_conv_igemm(const sseg_conv_geom_t* geom, const void* w_bf16, long w_ld, int cout, const sseg_act_t* out,
                    int out_f32, const float* bias, const sseg_act_t* addend, float* stat_sum, float* stat_sqsum,
                    sseg_stream_t stream);

/*
 * Convolution + TRAIN-mode BatchNorm (+ shortcut, ReLU, Dropout2d mask) in one kernel (single-GPU F.batch_norm branch,
 * lib/nn/modules/batchnorm.py:58-61): conv -> batch statistics -> normalise -> (+residual) -> ReLU of
 * models/resnet.py:37-53,72-92 / models/models.py:160-167 without sseg_bn_finalize / sseg_bn_apply launches.
 * One persistent CTA per SM keeps all its accumulators in tensor memory across an in-kernel grid barrier, so the layer
 * must fit: ceil(tiles / #SMs) * tile_columns <= 512 (sseg_conv_bn_train_fits returns 1 / 0 for the same arguments).
 *   y        : optional bf16 tensor receiving the raw convolution output (the backward pass reads it)
 *   a_out    : bf16 activation, same shape as y
 *   stat_sum / stat_sqsum : float[cout], zeroed by the caller; counter: one zeroed uint32 (the grid barrier)
 *   mean_out .. shift_out : float[cout] written for the backward pass; running_* updated with `momentum` if given
 *   res (+ rscale/rshift) : shortcut tensor (and, for projection shortcuts, its own BN affine); res_after_relu as in
 *                           sseg_bn_apply; chanmul: float[N][cout] Dropout2d keep-mask/(1-p) or NULL
 */
/* world > 1 (SynchronizedBatchNorm's data-parallel branch, batchnorm.py:63-81,123-139): the cooperative kernels pool the
 * per-rank partial sums over NVLink peer memory themselves. bases: every rank's arena (sseg_peer_alloc / sseg_peer_open);
 * data_off: [sum C | sqsum C | count] (forward, data_stride = C) or [s1 | s2raw] (backward) inside each arena; flag_off:
 * `world` int slots of the step-number handshake (never reset); step: device counter advanced by sseg_peer_step. */
typedef struct {
  void* const* bases;
  int world, rank;
  long data_off, data_stride, flag_off;
  const int* step;
} sseg_coop_peer_t;

typedef struct {
  const float* gamma;
  const float* beta;
  float eps, momentum, count;
  float* stat_sum;
  float* stat_sqsum;
  unsigned int* counter;
  float* mean_out;
  float* invstd_out;
  float* scale_out;
  float* shift_out;
  float* running_mean;
  float* running_var;
  const sseg_act_t* res;
  const float* rscale;
  const float* rshift;
  const float* chanmul;
  int relu, res_after_relu;
  /* synchronised branch (all NULL / 0 on a single GPU): stat_sum / stat_sqsum must then point at data_off inside this rank's
   * arena; clamp(var, eps)^-1/2; tmp_running_* accumulators + running_iter are advanced here, running_mean / running_var
   * are refreshed from them by sseg_bn_running_from_tmp; count_out receives the pooled pixel count. */
  const sseg_coop_peer_t* peer;
  float* tmp_running_mean;
  float* tmp_running_var;
  float* running_iter;
  float* count_out;
} sseg_bn_fused_t;
/* running_mean = tmp_running_mean / running_iter, running_var likewise (batchnorm.py:136-137). */
int sseg_bn_running_from_tmp(const float* tmp_running_mean, const float* tmp_running_var, const float* running_iter,
                             float* running_mean, float* running_var, int C, sseg_stream_t stream);
int sseg_conv_bn_train(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout, const sseg_act_t* y,
                       const sseg_act_t* a_out, const sseg_bn_fused_t* bn, sseg_stream_t stream);
/* sseg_conv_igemm with the layer's BatchNorm statistics AND their finalisation (train mode, one GPU: F.batch_norm's
 * training branch, lib/nn/modules/batchnorm.py:58-61) in the same launch: every CTA adds its tile's per-channel sum / sum
 * of squares to bn->stat_sum / stat_sqsum and takes a ticket from bn->counter (one zeroed uint32); the CTA that takes the
 * last ticket computes mean / inv_std / scale / shift for all channels and updates running_mean / running_var (both may be
 * NULL). Replaces sseg_conv_igemm(stats) + sseg_bn_finalize(SSEG_BN_TRAIN); the res / relu / peer fields of *bn are not
 * used (the normalisation itself stays in sseg_bn_apply). out: bf16, the raw convolution output. */
int sseg_conv_igemm_bnfin(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout, const sseg_act_t* out,
                          const sseg_bn_fused_t* bn, sseg_stream_t stream);
int sseg_conv_bn_train_fits(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout, const sseg_act_t* y,
                            const sseg_act_t* a_out, const sseg_bn_fused_t* bn);

/*
 * Backward twin of sseg_conv_bn_train: the data gradient of the CONSUMER convolution and the complete BatchNorm backward
 * of the PRODUCER layer (train-mode statistics, single GPU) in one persistent kernel - the gradient with respect to the
 * producer's activation stays in tensor memory across the grid barrier and only dy (the gradient with respect to the
 * producer's convolution output) is written. Replaces sseg_conv_igemm_bnbwd + sseg_bn_bwd_apply; same restrictions as
 * sseg_conv_igemm_bnbwd (producer has one consumer, ReLU, no shortcut / dropout mask), same fit rule as
 * sseg_conv_bn_train (sseg_conv_dgrad_bn_fits). autograd of models/resnet.py:37-43,72-82 + lib/nn/modules/batchnorm.py:58-61.
 *   g, w_bf16, cout : as for sseg_conv_igemm with the transposed / tap-mirrored weight (cout = the producer's channels)
 *   y        : the producer's saved convolution output; fscale / fshift its forward scale / shift (ReLU mask)
 *   mean, invstd, count : the producer's batch statistics;  s1 (= dbeta), s2_raw : float[cout], zeroed by the caller
 *   dgamma_out : float[cout] or NULL;  counter : one zeroed uint32;  dy_out : bf16, shape of y
 *   peer (NULL on a single GPU): s1 / s2_raw then are this rank's partial-sum slots at peer->data_off inside its arena,
 *   count_dev the pooled pixel count, and dbeta_out / dgamma_out receive the pooled sums divided by world.
 */
int sseg_conv_dgrad_bn(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout, const sseg_act_t* y,
                       const sseg_act_t* dy_out, const float* fscale, const float* fshift, const float* mean,
                       const float* invstd, float count, float* s1, float* s2_raw, float* dgamma_out,
                       unsigned int* counter, const sseg_coop_peer_t* peer, const float* count_dev, float* dbeta_out,
                       sseg_stream_t stream);
int sseg_conv_dgrad_bn_fits(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout, const sseg_act_t* y,
                            const sseg_act_t* dy_out, const float* fscale, const float* fshift, const float* mean,
                            const float* invstd, float count, float* s1, float* s2_raw, float* dgamma_out,
                            unsigned int* counter);

/*
 * Convolution with the inference-time epilogue fused: out = relu?( conv(x) * scale[co] + shift[co] (+ addend) ).
 * BatchNorm with running statistics is a per-channel affine (F.batch_norm eval branch, lib/nn/modules/batchnorm.py:
 * 58-61), so conv -> BN -> (+shortcut) -> ReLU (models/resnet.py:37-53,72-92; models/models.py:160-167) is ONE kernel
 * and the raw convolution output never reaches HBM. relu: 0 none | 1 after the addend (residual blocks) | 2 before the
 * addend (FPN lateral + top-down add, models/models.py:561-563); + 4: ReLU6 (nn.ReLU6, models/mobilenet.py:26,34).
 * bf16 output; scale/shift float[cout] or both NULL.
 */
int sseg_conv_igemm_affine(const sseg_conv_geom_t* g, const void* w_bf16, long w_ld, int cout, const sseg_act_t* out,
                           const float* scale, const float* shift, int relu, const sseg_act_t* addend,
                           sseg_stream_t stream);

/*
 * Data gradient with the BN-backward reduction of the PRODUCER layer fused into the epilogue.  `out` is the gradient
 * w.r.t. the producer's output a = relu(y*fscale + fshift); while the tile is still on chip the epilogue also accumulates
 *     s1[c] += sum g',   s2_raw[c] += sum g' * y,     g' = out * [y*fscale + fshift > 0]
 * (caller zeroes s1/s2_raw; y: the producer's saved conv output, same n,h,w as out, c >= cout). This replaces
 * sseg_bn_bwd_reduce for layers with a single consumer and no shortcut; s2_raw is converted to sum g'*xhat by
 * sseg_bn_bwd_apply(..., s2_raw = 1) / sseg_bn_bwd_peer_sum(..., s2_raw = 1).
 */
int sseg_conv_igemm_bnbwd(const sseg_conv_geom_t* geom, const void* w_bf16, long w_ld, int cout, const sseg_act_t* out,
                          const sseg_act_t* addend, const sseg_act_t* y, const float* fscale, const float* fshift,
                          float* s1, float* s2_raw, sseg_stream_t stream);
/* The same for a producer WITH a shortcut (Bottleneck / BasicBlock output a = relu(bn(y) + shortcut), models/resnet.py:37-53,
 * 72-92): the ReLU mask is not a function of y alone, so it is taken from the producer's saved output,
 *     g' = out * [a > 0],   s1[c] += sum g',   s2_raw[c] += sum g' * y.
 * `out` must be the COMPLETE gradient of a: this is the launch that adds the last contribution (addend = the gradient
 * accumulated so far). Replaces sseg_bn_bwd_reduce for the block outputs of the residual stages. */
int sseg_conv_igemm_bnbwd_res(const sseg_conv_geom_t* geom, const void* w_bf16, long w_ld, int cout, const sseg_act_t* out,
                              const sseg_act_t* addend, const sseg_act_t* y, const sseg_act_t* a, float* s1, float* s2_raw,
                              sseg_stream_t stream);

/*
 * Weight gradient of the same convolution (autograd of nn.Conv2d w.r.t. weight):
 *   dw[co][tap_koff[t] + ci] += sum_{n,h,w} dy[n,h,w,co] * X_t[n, h+dh_t, w+dw_t, ci]
 * GEMM with K = pixels (both operands MN-major in shared memory), split over pixels across CTAs and
 * accumulated with fp32 atomics: the caller zeroes dw (float [cout][dw_ld]). dy may carry zero padding channels
 * beyond cout (dy->c >= cout, multiple of 8); rows >= cout are not written.
 */
int sseg_conv_wgrad(const sseg_conv_geom_t* geom, const sseg_act_t* dy, int cout, float* dw, long dw_ld,
                    sseg_stream_t stream);


/* ---- weights ---------------------------------------------------------------------------- */
/* fp32 OIHW master weight (nn.Conv2d.weight, T = kh*kw) -> bf16 GEMM operands:
 *   w_fwd   [O][fwd_ld]    : w_fwd[o][t*I + i]          (forward; K ordered tap-major like sseg_conv_geom_t)
 *   w_dgrad [I][dgrad_ld]  : w_dgrad[i][t*o_pad + o]    (data gradient; o_pad = O rounded up to 64, the caller
 *                                                         zero-fills the buffer once so padding stays 0)
 * Either output may be NULL. */
int sseg_prep_conv_weight(const float* w_oihw, int O, int I, int T, void* w_fwd, long fwd_ld, void* w_dgrad,
                          long dgrad_ld, int o_pad, sseg_stream_t stream);
/* fp32 [O][g_ld] tap-major weight gradient (sseg_conv_wgrad output) -> fp32 OIHW: out (=|+=) scale * g. */
int sseg_grad_to_oihw(const float* g, long g_ld, int O, int I, int T, float* out, float scale, int accumulate,
                      sseg_stream_t stream);

/* Batched variants: one launch for every convolution of a model. `table_dev` is a DEVICE array of n descriptors
 * (built once by the caller); descriptor k owns CTAs [first_tile, first_tile + ceil(O/32)*ceil(ceil(I/32)/r)) with
 * r = 8 for T == 1 (a CTA walks 8 consecutive 32-wide I tiles of a pointwise conv) else 1; T <= 9. */
typedef struct {
  const float* w; /* fp32 OIHW master weight (prep) */
  void* wf;       /* bf16 [O][fwd_ld] or NULL */
  void* wd;       /* bf16 [I][dgrad_ld] or NULL */
  const float* g_src; /* fp32 [O][g_ld] tap-major gradient (grads) */
  float* g_dst;       /* fp32 OIHW gradient */
  long fwd_ld, dgrad_ld, g_ld;
  int O, I, T, o_pad;
  int first_tile;
  int reserved;
} sseg_weight_desc_t;
int sseg_prep_conv_weights_batched(const sseg_weight_desc_t* table_dev, int n, int total_tiles, sseg_stream_t stream);
/* The same pass on at most max_blocks thread blocks (0 = one block per tile): a thin grid for a side stream, so that the
 * pass does not keep the main stream's kernels waiting behind thousands of pending blocks. */
int sseg_prep_conv_weights_batched_ex(const sseg_weight_desc_t* table_dev, int n, int total_tiles, int max_blocks,
                                      sseg_stream_t stream);
int sseg_grads_to_oihw_batched(const sseg_weight_desc_t* table_dev, int n, int total_tiles, float scale,
                               sseg_stream_t stream);

/* ---- optimizer -------------------------------------------------------------------------- */
/* torch.optim.SGD(momentum, weight_decay) step for every parameter in ONE launch (train.py:115-127, :47-48).
 * chunks_dev: DEVICE array; each chunk is a contiguous piece (n <= 65536 elements) of one parameter tensor with its
 * gradient and momentum buffer; vec4 = 1 when the three pointers are 16-byte aligned and n % 4 == 0. */
typedef struct {
  float* param;
  const float* grad;
  float* momentum_buf;
  float weight_decay;
  int n;
  int vec4;
  int reserved;
} sseg_sgd_chunk_t;
int sseg_sgd_step(const sseg_sgd_chunk_t* chunks_dev, int nchunks, float lr, float momentum, int first_step,
                  sseg_stream_t stream);
/* x[0..n) *= *scalar_dev (a DEVICE scalar); nothing is read or written when the scalar is exactly 1. Replaces the
 * multiplication by grad_output that autograd's chain rule applies to the step's parameter gradients when
 * `loss.backward()` is called (train.py:43): the seed gradient is 1 there, so the pass is free. x 16-byte aligned. */
int sseg_scale_by_scalar(float* x, long n, const float* scalar_dev, sseg_stream_t stream);

/* ---- stem convolution (Cin = 3, 3x3, stride 2, pad 1, Cout = 64): models/resnet.py:100 -------------- */
/* img: fp32 NCHW [N,3,H,W]; w: fp32 OIHW [64,3,3,3]; out: bf16 NHWC [N,Ho,Wo,64] dense; optional BN statistics. */
int sseg_stem_conv_fwd(const float* img, int N, int H, int W, const float* w, void* out, float* stat_sum,
                       float* stat_sqsum, sseg_stream_t stream);
/* dw (fp32 OIHW [64,3,3,3], caller-zeroed) += weight gradient; dy bf16 NHWC [N,Ho,Wo,64] dense. */
int sseg_stem_conv_wgrad(const float* img, int N, int H, int W, const void* dy, float* dw, sseg_stream_t stream);

/* ---- batch normalisation: lib/nn/modules/batchnorm.py:56-139 ---------------------------- */
#define SSEG_BN_TRAIN 0      /* F.batch_norm training branch (:58-61): var + eps, torch running-stat update      */
#define SSEG_BN_TRAIN_SYNC 1 /* data-parallel branch (:63-81, :123-139): clamp(var, eps), accumulator running stats */
#define SSEG_BN_EVAL 2       /* running statistics (:58-61 with training=False)                                  */
/* (sum, sqsum, count) -> mean, inv_std, scale = gamma*inv_std, shift = beta - mean*scale, + running-stat update.
 * count: *count_dev if non-NULL (after a cross-rank all-reduce), else count_host. */
int sseg_bn_finalize(const float* sum, const float* sqsum, const float* count_dev, float count_host, const float* gamma,
                     const float* beta, float eps, float momentum, int mode, int update_running, float* running_mean,
                     float* running_var, float* tmp_running_mean, float* tmp_running_var, float* running_iter,
                     float* mean_out, float* invstd_out, float* scale, float* shift, int C, sseg_stream_t stream);
/* out = [relu](y*scale + shift + res') * chanmul[n][c];  res' = res (*rscale + rshift if given);
 * res_after_relu = 1: out = relu(y*scale + shift) + res' (UPerNet lateral + top-down add, models/models.py:557-563).
 * Fuses BN-apply, the residual add of Bottleneck/BasicBlock (models/resnet.py:45-53,84-92), ReLU and the
 * Dropout2d channel mask (models/models.py:460). y/res/out: bf16 [P][ld]. chanmul: float [N][C] or NULL. */
int sseg_bn_apply(const void* y, long y_ld, const float* scale, const float* shift, const void* res, long res_ld,
                  const float* rscale, const float* rshift, const float* chanmul, void* out, long out_ld, long P,
                  long pix_per_img, int C, int relu, int res_after_relu, sseg_stream_t stream);
/* sseg_bn_finalize(SSEG_BN_TRAIN) + sseg_bn_apply in ONE launch (single-GPU training: F.batch_norm semantics,
 * lib/nn/modules/batchnorm.py:58-61): every thread derives scale/shift for its channels from (sum, sqsum, count);
 * mean/invstd/scale/shift are also stored for the backward pass and the running statistics (optional) are updated. */
int sseg_bn_finalize_apply(const float* sum, const float* sqsum, float count, const float* gamma, const float* beta,
                           float eps, float momentum, float* running_mean, float* running_var, float* mean_out,
                           float* invstd_out, float* scale_out, float* shift_out, const void* y, long y_ld, const void* res,
                           long res_ld, const float* rscale, const float* rshift, const float* chanmul, void* out,
                           long out_ld, long P, long pix_per_img, int C, int relu, int res_after_relu,
                           sseg_stream_t stream);
/* backward.  g' = g * chanmul * [ReLU active].  The ReLU mask comes from the saved layer output (a > 0) or, for layers
 * without a shortcut, is recomputed as (y*scale + fshift > 0) when `a` is NULL and `fshift` is given (one tensor less to
 * read); both NULL = the layer has no ReLU.
 * pass 1: s1[c] += sum g', s2[c] += sum g' * xhat          (= dbeta, dgamma; caller zeroes them) */
int sseg_bn_bwd_reduce(const void* g, long g_ld, const void* a, long a_ld, const void* y, long y_ld, const float* mean,
                       const float* invstd, const float* scale, const float* fshift, const float* chanmul, float* s1,
                       float* s2, long P, long pix_per_img, int C, sseg_stream_t stream);
/* pass 2: dy = scale*(g' - s1/M - xhat*s2/M) (eval_mode: dy = scale*g'); dres (optional) = g'.
 * s2_raw = 1: s2 holds sum g'*y (sseg_conv_igemm_bnbwd); it is converted as invstd*(s2 - mean*s1) and the converted
 * value (= dgamma) is stored to dgamma_out (optional). */
int sseg_bn_bwd_apply(const void* g, long g_ld, const void* a, long a_ld, const void* y, long y_ld, const float* mean,
                      const float* invstd, const float* scale, const float* fshift, const float* chanmul, const float* s1,
                      const float* s2, const float* count_dev, float count_host, void* dy, long dy_ld, void* dres,
                      long dres_ld, long P, long pix_per_img, int C, int eval_mode, int s2_raw, float* dgamma_out,
                      sseg_stream_t stream);

/* ---- SyncBN over NVLink peer memory (world_size > 1) ------------------------------------- */
/* One cudaMalloc'ed, zero-initialised arena per rank, exported with CUDA IPC (64-byte handle) and mapped by the peers.
 * Offsets below are in 4-byte units from the arena base. */
int sseg_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int sseg_peer_open(const unsigned char* handle64, void** ptr);
int sseg_peer_close(void* ptr);
int sseg_peer_free(void* ptr);
/* *step += 1 (once per training step; the handshake flags carry step numbers and are never reset) */
int sseg_peer_step(int* step, sseg_stream_t stream);
/* Forward of the synchronised branch (lib/nn/modules/batchnorm.py:98-139) without a collective call: handshake with all
 * peers on flag slots [flag_off, flag_off + world), then pool [sum C | sqsum C | count] found at stats_off in EVERY
 * rank's arena and finish like sseg_bn_finalize(SSEG_BN_TRAIN_SYNC). count_out receives the pooled pixel count.
 * update_running: 0 = leave the running statistics alone; 1 = advance the accumulators / running_iter and refresh
 * running_mean / running_var (a second tiny launch); 2 = advance only - the caller refreshes them later with
 * sseg_bn_running_from_tmp (e.g. on another stream, off the forward pass's dependency chain). */
int sseg_bn_finalize_peer(void* const* bases, int world, int rank, long stats_off, long flag_off, const int* step,
                          const float* gamma, const float* beta, float eps, float momentum, int update_running,
                          float* running_mean, float* running_var, float* tmp_running_mean, float* tmp_running_var,
                          float* running_iter, float* mean_out, float* invstd_out, float* scale, float* shift,
                          float* count_out, int C, sseg_stream_t stream);
/* Backward: pool [s1 C | s2 C] partials at part_off over the ranks -> s1_tot, s2_tot; dbeta = s1_tot/world,
 * dgamma = s2_tot/world (the gradient-bucket all-reduce sums them over ranks again). */
int sseg_bn_bwd_peer_sum(void* const* bases, int world, int rank, long part_off, long flag_off, const int* step,
                         float* s1_tot, float* s2_tot, float* dbeta, float* dgamma, const float* mean, const float* invstd,
                         int s2_raw, int C, sseg_stream_t stream);
/* The same two exchanges with a PUSH protocol (the default of the step programs): rank r owns, in every rank's arena, the
 * 8-byte slots [inbox_off/2 + r*n, +n) of the layer's inbox (n = 2C+1 forward, 2C backward; inbox_off in 4-byte units,
 * even). Each thread sends its channel's partial sums to all peers as {fp32 value, int32 step} messages (plain 64-bit stores
 * over NVLink, no fence, no separate flag) and polls the peers' messages in LOCAL memory until their tag is the current
 * step: one NVLink one-way latency per exchange instead of flag + fence + a remote-load round trip. Results are
 * bit-identical to the flag protocol's (same rank-ordered sums). Inbox space per layer: world * n * 8 bytes, zero at start. */
int sseg_bn_finalize_peer_ll(void* const* bases, int world, int rank, long stats_off, long inbox_off, const int* step,
                             const float* gamma, const float* beta, float eps, float momentum, int update_running,
                             float* running_mean, float* running_var, float* tmp_running_mean, float* tmp_running_var,
                             float* running_iter, float* mean_out, float* invstd_out, float* scale, float* shift,
                             float* count_out, int C, sseg_stream_t stream);
int sseg_bn_bwd_peer_sum_ll(void* const* bases, int world, int rank, long part_off, long inbox_off, const int* step,
                            float* s1_tot, float* s2_tot, float* dbeta, float* dgamma, const float* mean,
                            const float* invstd, int s2_raw, int C, sseg_stream_t stream);
/* sseg_bn_bwd_peer_sum + sseg_bn_bwd_apply in ONE launch (one dependent kernel less per BatchNorm layer on the backward
 * chain of a multi-GPU step): every block runs the flag handshake, pools this layer's partial sums [s1 | s2] (at part_off /
 * part_off + C of every rank's arena; s2 raw when s2_raw) straight out of peer memory, applies
 * dy = scale * (g' - s1/M - xhat * s2/M) with M = *count_dev (the pooled pixel count), and block column 0 stores
 * dbeta_out = s1 / world, dgamma_out = s2 / world. Arguments as for the two functions it replaces. */
int sseg_bn_bwd_apply_peer(void* const* bases, int world, int rank, long part_off, long flag_off, const int* step,
                           const void* g, long g_ld, const void* a, long a_ld, const void* y, long y_ld, const float* mean,
                           const float* invstd, const float* scale, const float* fshift, const float* chanmul,
                           const float* count_dev, void* dy, long dy_ld, void* dres, long dres_ld, long P, long pix_per_img,
                           int C, int s2_raw, float* dbeta_out, float* dgamma_out, sseg_stream_t stream);

/* ---- pooling / resize ------------------------------------------------------------------- */
/* nn.MaxPool2d(3, 2, 1) (models/resnet.py:109). Dense bf16 NHWC; idx (1 byte / output element) feeds the backward. */
int sseg_maxpool_fwd(const void* x, int N, int H, int W, int C, void* out, void* idx, sseg_stream_t stream);
int sseg_maxpool_bwd(const void* dout, const void* idx, void* dx, int N, int H, int W, int C, sseg_stream_t stream);
/* nn.AdaptiveAvgPool2d(S) (models/models.py:447): x bf16 [N,H,W,C] (pixel stride x_ld) -> out dense [N,S,S,C]. */
int sseg_avgpool_fwd(const void* x, long x_ld, int N, int H, int W, int C, int S, void* out, sseg_stream_t stream);
/* dx = base + sum_k avgpool_backward_k(dpool[k]) for up to 4 scales in one pass (base may be NULL). */
int sseg_avgpool_bwd(const void* base, long base_ld, const void* const* dpool, const int* scales, int nscales, void* dx,
                     long dx_ld, int N, int H, int W, int C, sseg_stream_t stream);
/* F.interpolate(mode='bilinear', align_corners=False) (models/models.py:472-475) and its adjoint (gather form). */
int sseg_bilinear_fwd(const void* x, long x_ld, int N, int Hi, int Wi, int C, void* out, long out_ld, int Ho, int Wo,
                      sseg_stream_t stream);
/* scratch: float[N*Ho*Wi*C] (the W-pass intermediate of the separable adjoint) */
int sseg_bilinear_bwd(const void* dout, long dout_ld, int N, int Ho, int Wo, int C, void* dx, long dx_ld, int Hi, int Wi,
                      int accumulate, float* scratch, sseg_stream_t stream);

/* HRNet exchange unit (models/hrnet.py:225-250: `y = y + x[j]` / `+ fuse_layers[i][j](x[j])` /
 * `+ F.interpolate(fuse_layers[i][j](x[j]), ...)`, then ReLU) as ONE pass:
 *   out[n,ho,wo,c] = relu?( sum_k ( scale_k[c] * sample_k(x_k)[n,ho,wo,c] + shift_k[c] ) )
 * x: bf16 NHWC [N,h,w,C] dense with pixel stride ld. (h,w) == (Ho,Wo): read in place; otherwise sampled bilinearly
 * (align_corners=False) on the fly. scale/shift: float[C] (the term's batch-norm affine, 16B aligned) or both NULL. */
typedef struct {
  const void* x;
  int h, w;
  long ld;
  const float* scale;
  const float* shift;
} sseg_sum_term_t;
int sseg_sum_terms(const sseg_sum_term_t* terms, int nterms, int N, int Ho, int Wo, int C, void* out, long out_ld, int relu,
                   sseg_stream_t stream);
/* Backward of that ReLU: ds = g * [out > 0] (bf16 [P][*_ld], C channels); optionally acc_out (+)= ds in the same pass
 * (the identity term's gradient; accumulate = 0 overwrites). autograd of models/hrnet.py:248. */
int sseg_relu_mask_bwd(const void* g, long g_ld, const void* out, long out_ld, void* ds, long ds_ld, void* acc_out,
                       long acc_ld, int accumulate, long P, int C, sseg_stream_t stream);

/* ---- fp32-accurate inference on bf16 tensor cores (BASELINE config 2: logits within 1e-3 of the fp32 reference) ------
 * Every activation / weight is a PAIR of bf16 tensors (hi = bf16(x), lo = bf16(x - hi)); conv(x, w) ~= x_hi*w_hi +
 * x_lo*w_hi + x_hi*w_lo is ONE sseg_conv_igemm launch (out_f32 = 1) over the virtual concat [x_hi | x_lo | x_hi] and the
 * K-concatenated weight [w_hi | w_hi | w_lo] per tap (sseg_prep_conv_weight_split). The functions below surround it. */
/* out pair = relu?( z*scale + shift (+ res pair) ); z: fp32 NHWC view (a strided view subsamples a stride-1 output, which
 * is how stride-2 convolutions run in this mode); scale/shift NULL = identity; out/res: pixel-dense bf16 [N*H*W][ld]. */
int sseg_split_affine(const sseg_act_t* z, const float* scale, const float* shift, const void* res_hi, const void* res_lo,
                      long res_ld, void* out_hi, void* out_lo, long out_ld, int relu, int res_after_relu,
                      sseg_stream_t stream);
/* conv1 of the deep stem in fp32 (models/resnet.py:100): img fp32 NCHW -> out fp32 NHWC [N,Ho,Wo,64]. */
int sseg_stem_conv_fwd_f32(const float* img, int N, int H, int W, const float* w, float* out, sseg_stream_t stream);
/* nn.MaxPool2d(3,2,1), nn.AdaptiveAvgPool2d(S), F.interpolate(bilinear) on pairs (fp32 arithmetic on hi + lo). */
int sseg_maxpool_pair_fwd(const void* x_hi, const void* x_lo, int N, int H, int W, int C, void* out_hi, void* out_lo,
                          sseg_stream_t stream);
int sseg_avgpool_pair_fwd(const void* x_hi, const void* x_lo, long x_ld, int N, int H, int W, int C, int S, void* out_hi,
                          void* out_lo, sseg_stream_t stream);
int sseg_bilinear_pair_fwd(const void* x_hi, const void* x_lo, long x_ld, int N, int Hi, int Wi, int C, void* out_hi,
                           void* out_lo, long out_ld, int Ho, int Wo, sseg_stream_t stream);
/* fp32 OIHW weight -> bf16 [O][ld], per tap t: out[o][t*3I + i] = out[o][t*3I + I + i] = w_hi, out[o][t*3I + 2I + i] = w_lo. */
int sseg_prep_conv_weight_split(const float* w_oihw, int O, int I, int T, void* out, long ld, sseg_stream_t stream);

/* First layer of MobileNetV2 (nn.Conv2d(3, cout, 3, 2, 1), cout <= 64, multiple of 8; models/mobilenet.py:102) from the
 * fp32 NCHW image with the inference epilogue: out = relu6?(conv * scale + shift), bf16 NHWC [N,Ho,Wo,cout] dense. */
int sseg_stem_conv_affine(const float* img, int N, int H, int W, const float* w, int cout, const float* scale,
                          const float* shift, int relu6, void* out, sseg_stream_t stream);
/* Depthwise 3x3 convolution (nn.Conv2d(C, C, 3, stride, dilation, groups=C), models/mobilenet.py:52,62) with the
 * inference epilogue: out = relu6?( dwconv(x) * scale + shift ). x: bf16 NHWC [N,H,W,C] dense; w: fp32 [C][3][3] (the
 * module's weight [C,1,3,3]); 'same' padding = dilation; out: bf16 [N,Ho,Wo,C] dense, Ho = ceil(H/stride). */
int sseg_dwconv_affine(const void* x, int N, int H, int W, int C, const float* w, int stride, int dilation,
                       const float* scale, const float* shift, int relu6, void* out, sseg_stream_t stream);

/* ---- loss ------------------------------------------------------------------------------- */
/* F.log_softmax + nn.NLLLoss(ignore_index=-1) + pixel_acc (models/models.py:12-18,37-42,492-493; train.py:154).
 * logits fp32 [P][ld]; label int64 [P]; lse float [P] out; accum float[3] (caller-zeroed):
 *   accum[0] += sum_valid(lse - logit[label]); accum[1] += #valid; accum[2] += #(valid and argmax == label). */
int sseg_softmax_nll_fwd(const float* logits, long ld, int C, const long long* label, long P, float* lse, float* accum,
                         sseg_stream_t stream);
/* out[0] = main[0]/main[1] + ds_scale * ds[0]/ds[1] (ds may be NULL); out[1] = main[2]/(main[1] + 1e-10). */
int sseg_nll_finalize(const float* accum_main, const float* accum_ds, float ds_scale, float* out, sseg_stream_t stream);
/* dlogits bf16 [P][ld_out]: weight/accum[1] * (softmax - onehot) on valid pixels, else 0; columns [C, c_store) = 0. */
int sseg_softmax_nll_bwd(const float* logits, long ld, int C, const long long* label, const float* lse, const float* accum,
                         float weight, long P, void* dlogits, long ld_out, int c_store, sseg_stream_t stream);
/* out[c] += sum_p x[p][c]  (bias gradients); x bf16 [P][ld]. */
int sseg_colsum(const void* x, long ld, long P, int C, float* out, sseg_stream_t stream);
/* Inference head (models/models.py:480-484; eval.py:71-72): bilinear-upsample fp32 NHWC logits [N,Hi,Wi,ld] to
 * (Ho,Wo), softmax over C, write fp32 NCHW probs (= or +=) weight * softmax.  log_output = 1 writes log-softmax
 * instead (F.log_softmax, models/models.py:492-493; with Ho,Wo = Hi,Wi the resize is the identity). */
int sseg_upsample_softmax(const float* logits, long ld, int N, int Hi, int Wi, int C, float* probs, int Ho, int Wo,
                          float weight, int accumulate, int log_output, sseg_stream_t stream);

/* ---- layout ----------------------------------------------------------------------------- */
int sseg_nhwc_bf16_to_nchw_f32(const void* x, long ld, int N, int H, int W, int C, float* out, sseg_stream_t stream);
int sseg_nchw_f32_to_nhwc_bf16(const float* x, int N, int H, int W, int C, void* out, long ld, sseg_stream_t stream);

/* ---- input pipeline --------------------------------------------------- */
/* The reference's img_transform (mit_semseg/dataset.py:53-58: float32(x)/255, then (x - mean)/std per channel, HWC -> CHW)
 * on the device, so that the bytes - a quarter of the fp32 tensor - are what crosses PCIe.
 * img_u8 uint8 [N][H][W][3] (W % 4 == 0); valid_hw DEVICE int32 [N][2] = rows, columns of image n that are real: outside
 * them the output is 0.0f, which is what the reference's pre-zeroed batch tensor holds there (dataset.py:150-151,179);
 * mean_std HOST float[6] = mean r,g,b then std r,g,b; out fp32 [N][3][H][W]. Same fp32 operations in the same order as
 * the reference (IEEE division): bit-identical. */
int sseg_image_transform(const void* img_u8, int N, int H, int W, const int* valid_hw, const float* mean_std, float* out,
                         sseg_stream_t stream);
/* The reference's segm_transform (dataset.py:60-63: stored id - 1 as int64) on the device. seg_u8 uint8 [N][Hs][Ws] = the
 * label map already strided by the segm_downsampling_rate `rate`; rows / columns beyond ceil(valid_hw / rate) are padding
 * and come out as 0 (NOT -1: the reference pads its batch label tensor with zeros, dataset.py:152-155,180). */
int sseg_label_transform(const void* seg_u8, int N, int Hs, int Ws, const int* valid_hw, int rate, long long* out,
                         sseg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SSEG_B200_H_ */
