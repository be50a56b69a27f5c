"""Eager, one-call-per-kernel Python wrappers over the C ABI (include/sseg_b200.h).

These are the building blocks the step programs in `engine/program.py` are assembled from and what the
`-m gpu` parity tests call.  Every wrapper takes torch CUDA tensors, passes raw device pointers + the current
CUDA stream, and raises on any error.  Activations are NHWC bf16 tensors of shape [N, H, W, C] (a channel
slice `t[..., a:b]` of a wider buffer is fine: the pixel stride is taken from `t.stride(2)`).
"""
import torch

from . import _C


def _stream():
    return _C.c_void_p(torch.cuda.current_stream().cuda_stream)


def act(t):
    """NHWC tensor (possibly a channel slice) -> sseg_act_t."""
    assert t.dim() == 4 and t.stride(3) == 1, "expected NHWC with contiguous channels"
    n, h, w, c = t.shape
    ld = t.stride(2)
    assert t.stride(1) == w * ld and t.stride(0) == h * w * ld, "expected dense NHWC pixels"
    return _C.Act(_C.c_void_p(t.data_ptr()), n, h, w, c, ld)


def conv_taps(ksize, dilation):
    """Tap offsets (dh, dw) of a stride-1 'same' convolution, in the weight's (kh, kw) order."""
    r = ksize // 2
    dh, dw = [], []
    for i in range(ksize):
        for j in range(ksize):
            dh.append((i - r) * dilation)
            dw.append((j - r) * dilation)
    return dh, dw


def make_geom(srcs, taps, tap_src=None, tap_koff=None):
    """sseg_conv_geom_t from NHWC source tensors and (dh, dw) tap lists. Default: every tap reads the channel
    concatenation of all sources and tap t's weights start at t * cin_total."""
    dh, dw = taps
    g = _C.Geom()
    g.nsrc = len(srcs)
    for i, s in enumerate(srcs):
        g.srcs[i] = act(s)
    g.ntaps = len(dh)
    cin = sum(s.shape[3] for s in srcs)
    for t in range(len(dh)):
        g.tap_dh[t], g.tap_dw[t] = dh[t], dw[t]
        g.tap_src[t] = -1 if tap_src is None else tap_src[t]
        g.tap_koff[t] = t * cin if tap_koff is None else tap_koff[t]
    return g


def conv_igemm(geom, w_bf16, cout, out, n_store=None, bias=None, addend=None, stat_sum=None, stat_sqsum=None):
    """out[n,h,w,:n_store] = conv(geom, w) (+bias) (+addend); optional BN statistics accumulation.
    w_bf16: 2-D [cout, K] bf16 (row = taps x channels)."""
    if n_store is None:
        n_store = (cout + 7) // 8 * 8
    assert out.dtype in (torch.float32, torch.bfloat16) and w_bf16.dim() == 2 and w_bf16.stride(1) == 1
    out_f32 = 1 if out.dtype == torch.float32 else 0
    ld_add = addend.stride(2) if addend is not None else 0
    _C.check(_C.lib().sseg_conv_igemm(geom, _C.ptr(w_bf16), w_bf16.stride(0), cout, _C.ptr(out), out_f32,
                                      out.stride(2), n_store, _C.ptr(bias), _C.ptr(addend), ld_add,
                                      _C.ptr(stat_sum), _C.ptr(stat_sqsum), _stream()))
    return out


def conv_wgrad(geom, dy, cout, dw):
    """dw[co, koff_t + ci] += sum_pixels dy[.., co] * x_t[.., ci]; dw: fp32 2-D [cout, K], pre-zeroed by the caller."""
    assert dw.dtype == torch.float32 and dw.dim() == 2 and dw.stride(1) == 1
    a = act(dy)
    _C.check(_C.lib().sseg_conv_wgrad(geom, a, cout, _C.ptr(dw), dw.stride(0), _stream()))
    return dw
