"""ctypes binding of libsseg_b200.so (the C ABI declared in include/sseg_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception is
raised.  The product path never routes through the oracle or a CPU implementation.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_long, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.abspath(os.path.join(_HERE, "..", "..", os.environ.get("SSEG_LIB", "libsseg_b200.so")))

MAX_SRCS = 5
MAX_TAPS = 9
MAX_SUM_TERMS = 4


class SsegError(RuntimeError):
    pass


class Act(Structure):
    """sseg_act_t: NHWC activation view."""
    _fields_ = [("ptr", c_void_p), ("n", c_int), ("h", c_int), ("w", c_int), ("c", c_int), ("ld", c_int),
                ("row_stride", c_long), ("img_stride", c_long)]


class Geom(Structure):
    """sseg_conv_geom_t: input-side geometry of a stride-1 convolution (sources, taps)."""
    _fields_ = [("nsrc", c_int), ("srcs", Act * MAX_SRCS), ("ntaps", c_int), ("tap_dh", c_int * MAX_TAPS),
                ("tap_dw", c_int * MAX_TAPS), ("tap_src", c_int * MAX_TAPS), ("tap_koff", c_int * MAX_TAPS)]


class WeightDesc(Structure):
    """sseg_weight_desc_t"""
    _fields_ = [("w", c_void_p), ("wf", c_void_p), ("wd", c_void_p), ("g_src", c_void_p), ("g_dst", c_void_p),
                ("fwd_ld", c_long), ("dgrad_ld", c_long), ("g_ld", c_long), ("O", c_int), ("I", c_int), ("T", c_int),
                ("o_pad", c_int), ("first_tile", c_int), ("reserved", c_int)]


class SumTerm(Structure):
    """sseg_sum_term_t: one term of an HRNet exchange-unit sum."""
    _fields_ = [("x", c_void_p), ("h", c_int), ("w", c_int), ("ld", c_long), ("scale", c_void_p), ("shift", c_void_p)]


class CoopPeer(Structure):
    """sseg_coop_peer_t: the peer arenas the cooperative conv+BN kernels pool their partial sums from (world > 1)."""
    _fields_ = [("bases", POINTER(c_void_p)), ("world", c_int), ("rank", c_int), ("data_off", c_long), ("data_stride", c_long),
                ("flag_off", c_long), ("step", c_void_p)]


class BnFused(Structure):
    """sseg_bn_fused_t: the BatchNorm half of sseg_conv_bn_train."""
    _fields_ = [("gamma", c_void_p), ("beta", c_void_p), ("eps", c_float), ("momentum", c_float), ("count", c_float),
                ("stat_sum", c_void_p), ("stat_sqsum", c_void_p), ("counter", c_void_p), ("mean_out", c_void_p),
                ("invstd_out", c_void_p), ("scale_out", c_void_p), ("shift_out", c_void_p), ("running_mean", c_void_p),
                ("running_var", c_void_p), ("res", POINTER(Act)), ("rscale", c_void_p), ("rshift", c_void_p),
                ("chanmul", c_void_p), ("relu", c_int), ("res_after_relu", c_int), ("peer", POINTER(CoopPeer)),
                ("tmp_running_mean", c_void_p), ("tmp_running_var", c_void_p), ("running_iter", c_void_p),
                ("count_out", c_void_p)]


class SgdChunk(Structure):
    """sseg_sgd_chunk_t"""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("momentum_buf", c_void_p), ("weight_decay", c_float),
                ("n", c_int), ("vec4", c_int), ("reserved", c_int)]


_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SsegError(
                "libsseg_b200.so not found at %s - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C semantic-segmentation-pytorch_b200/csrc`). There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        _declare(L)
        _lib = L
    return _lib


def _declare(L):
    L.sseg_last_error.restype = c_char_p
    L.sseg_last_error.argtypes = []
    L.sseg_version.restype = c_int
    L.sseg_launch_count.restype = c_long
    L.sseg_launch_count_reset.restype = None
    for name, args in _SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = c_int
        fn.argtypes = args


def check(rc):
    if rc != 0:
        raise SsegError("libsseg_b200 error %d: %s" % (rc, lib().sseg_last_error().decode()))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


_p = c_void_p
_ip = POINTER(c_int)
_SIGNATURES = {
    "sseg_conv_igemm": [POINTER(Geom), _p, c_long, c_int, POINTER(Act), c_int, _p, POINTER(Act), _p, _p, _p],
    "sseg_conv_igemm_affine": [POINTER(Geom), _p, c_long, c_int, POINTER(Act), _p, _p, c_int, POINTER(Act), _p],
    "sseg_conv_igemm_bnfin": [POINTER(Geom), _p, c_long, c_int, POINTER(Act), POINTER(BnFused), _p],
    "sseg_conv_bn_train": [POINTER(Geom), _p, c_long, c_int, POINTER(Act), POINTER(Act), POINTER(BnFused), _p],
    "sseg_conv_bn_train_fits": [POINTER(Geom), _p, c_long, c_int, POINTER(Act), POINTER(Act), POINTER(BnFused)],
    "sseg_conv_dgrad_bn": [POINTER(Geom), _p, c_long, c_int, POINTER(Act), POINTER(Act), _p, _p, _p, _p, c_float, _p, _p, _p,
                           _p, POINTER(CoopPeer), _p, _p, _p],
    "sseg_bn_running_from_tmp": [_p, _p, _p, _p, _p, c_int, _p],
    "sseg_conv_dgrad_bn_fits": [POINTER(Geom), _p, c_long, c_int, POINTER(Act), POINTER(Act), _p, _p, _p, _p, c_float, _p, _p,
                                _p, _p],
    "sseg_conv_igemm_bnbwd": [POINTER(Geom), _p, c_long, c_int, POINTER(Act), POINTER(Act), POINTER(Act), _p, _p, _p, _p, _p],
    "sseg_conv_igemm_bnbwd_res": [POINTER(Geom), _p, c_long, c_int, POINTER(Act), POINTER(Act), POINTER(Act), POINTER(Act), _p, _p, _p],
    "sseg_conv_wgrad": [POINTER(Geom), POINTER(Act), c_int, _p, c_long, _p],
    "sseg_prep_conv_weight": [_p, c_int, c_int, c_int, _p, c_long, _p, c_long, c_int, _p],
    "sseg_prep_conv_weights_batched": [_p, c_int, c_int, _p],
    "sseg_prep_conv_weights_batched_ex": [_p, c_int, c_int, c_int, _p],
    "sseg_grads_to_oihw_batched": [_p, c_int, c_int, c_float, _p],
    "sseg_sgd_step": [_p, c_int, c_float, c_float, c_int, _p],
    "sseg_scale_by_scalar": [_p, c_long, _p, _p],
    "sseg_grad_to_oihw": [_p, c_long, c_int, c_int, c_int, _p, c_float, c_int, _p],
    "sseg_stem_conv_fwd": [_p, c_int, c_int, c_int, _p, _p, _p, _p, _p],
    "sseg_stem_conv_wgrad": [_p, c_int, c_int, c_int, _p, _p, _p],
    "sseg_bn_finalize": [_p, _p, _p, c_float, _p, _p, c_float, c_float, c_int, c_int, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                         c_int, _p],
    "sseg_bn_apply": [_p, c_long, _p, _p, _p, c_long, _p, _p, _p, _p, c_long, c_long, c_long, c_int, c_int, c_int, _p],
    "sseg_bn_finalize_apply": [_p, _p, c_float, _p, _p, c_float, c_float, _p, _p, _p, _p, _p, _p, _p, c_long, _p, c_long, _p,
                               _p, _p, _p, c_long, c_long, c_long, c_int, c_int, c_int, _p],
    "sseg_bn_bwd_reduce": [_p, c_long, _p, c_long, _p, c_long, _p, _p, _p, _p, _p, _p, _p, c_long, c_long, c_int, _p],
    "sseg_bn_bwd_apply": [_p, c_long, _p, c_long, _p, c_long, _p, _p, _p, _p, _p, _p, _p, _p, c_float, _p, c_long, _p,
                          c_long, c_long, c_long, c_int, c_int, c_int, _p, _p],
    "sseg_image_transform": [_p, c_int, c_int, c_int, _p, POINTER(c_float), _p, _p],
    "sseg_label_transform": [_p, c_int, c_int, c_int, _p, c_int, _p, _p],
    "sseg_peer_alloc": [ctypes.c_size_t, POINTER(c_void_p), _p],
    "sseg_peer_open": [_p, POINTER(c_void_p)],
    "sseg_peer_close": [_p],
    "sseg_peer_free": [_p],
    "sseg_peer_step": [_p, _p],
    "sseg_bn_finalize_peer": [POINTER(c_void_p), c_int, c_int, c_long, c_long, _p, _p, _p, c_float, c_float, c_int, _p, _p,
                              _p, _p, _p, _p, _p, _p, _p, _p, c_int, _p],
    "sseg_bn_bwd_apply_peer": [POINTER(c_void_p), c_int, c_int, c_long, c_long, _p, _p, c_long, _p, c_long, _p, c_long, _p, _p, _p,
                               _p, _p, _p, _p, c_long, _p, c_long, c_long, c_long, c_int, c_int, _p, _p, _p],
    "sseg_bn_bwd_peer_sum": [POINTER(c_void_p), c_int, c_int, c_long, c_long, _p, _p, _p, _p, _p, _p, _p, c_int, c_int, _p],
    "sseg_bn_finalize_peer_ll": [POINTER(c_void_p), c_int, c_int, c_long, c_long, _p, _p, _p, c_float, c_float, c_int, _p, _p,
                              _p, _p, _p, _p, _p, _p, _p, _p, c_int, _p],
    "sseg_bn_bwd_peer_sum_ll": [POINTER(c_void_p), c_int, c_int, c_long, c_long, _p, _p, _p, _p, _p, _p, _p, c_int, c_int, _p],
    "sseg_maxpool_fwd": [_p, c_int, c_int, c_int, c_int, _p, _p, _p],
    "sseg_maxpool_bwd": [_p, _p, _p, c_int, c_int, c_int, c_int, _p],
    "sseg_avgpool_fwd": [_p, c_long, c_int, c_int, c_int, c_int, c_int, _p, _p],
    "sseg_avgpool_bwd": [_p, c_long, POINTER(c_void_p), _ip, c_int, _p, c_long, c_int, c_int, c_int, c_int, _p],
    "sseg_bilinear_fwd": [_p, c_long, c_int, c_int, c_int, c_int, _p, c_long, c_int, c_int, _p],
    "sseg_bilinear_bwd": [_p, c_long, c_int, c_int, c_int, c_int, _p, c_long, c_int, c_int, c_int, _p, _p],
    "sseg_sum_terms": [POINTER(SumTerm), c_int, c_int, c_int, c_int, c_int, _p, c_long, c_int, _p],
    "sseg_relu_mask_bwd": [_p, c_long, _p, c_long, _p, c_long, _p, c_long, c_int, c_long, c_int, _p],
    "sseg_stem_conv_affine": [_p, c_int, c_int, c_int, _p, c_int, _p, _p, c_int, _p, _p],
    "sseg_dwconv_affine": [_p, c_int, c_int, c_int, c_int, _p, c_int, c_int, _p, _p, c_int, _p, _p],
    "sseg_split_affine": [POINTER(Act), _p, _p, _p, _p, c_long, _p, _p, c_long, c_int, c_int, _p],
    "sseg_stem_conv_fwd_f32": [_p, c_int, c_int, c_int, _p, _p, _p],
    "sseg_maxpool_pair_fwd": [_p, _p, c_int, c_int, c_int, c_int, _p, _p, _p],
    "sseg_avgpool_pair_fwd": [_p, _p, c_long, c_int, c_int, c_int, c_int, c_int, _p, _p, _p],
    "sseg_bilinear_pair_fwd": [_p, _p, c_long, c_int, c_int, c_int, c_int, _p, _p, c_long, c_int, c_int, _p],
    "sseg_prep_conv_weight_split": [_p, c_int, c_int, c_int, _p, c_long, _p],
    "sseg_softmax_nll_fwd": [_p, c_long, c_int, _p, c_long, _p, _p, _p],
    "sseg_nll_finalize": [_p, _p, c_float, _p, _p],
    "sseg_softmax_nll_bwd": [_p, c_long, c_int, _p, _p, _p, c_float, c_long, _p, c_long, c_int, _p],
    "sseg_colsum": [_p, c_long, c_long, c_int, _p, _p],
    "sseg_upsample_softmax": [_p, c_long, c_int, c_int, c_int, c_int, _p, c_int, c_int, c_float, c_int, c_int, _p],
    "sseg_nhwc_bf16_to_nchw_f32": [_p, c_long, c_int, c_int, c_int, c_int, _p, _p],
    "sseg_nchw_f32_to_nhwc_bf16": [_p, c_int, c_int, c_int, c_int, _p, c_long, _p],
}

EXPORTED_SYMBOLS = ["sseg_last_error", "sseg_version", "sseg_launch_count", "sseg_launch_count_reset"] + list(
    _SIGNATURES)


def int_array(vals):
    return (c_int * len(vals))(*vals)


def act_array(acts):
    return (Act * len(acts))(*acts)
