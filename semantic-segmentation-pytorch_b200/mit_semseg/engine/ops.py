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


def conv_igemm(srcs, w_bf16, cout, taps, out, n_store=None, bias=None, addend=None, stat_sum=None, stat_sqsum=None):
    """out[n,h,w,:n_store] = conv(concat(srcs), w) (+bias) (+addend); optional BN statistics accumulation."""
    dh, dw = taps
    if n_store is None:
        n_store = (cout + 7) // 8 * 8
    acts = _C.act_array([act(s) for s in srcs])
    out_f32 = 1 if out.dtype == torch.float32 else 0
    assert out.dtype in (torch.float32, torch.bfloat16)
    ld_out = out.stride(2)
    ld_add = addend.stride(2) if addend is not None else 0
    _C.check(_C.lib().sseg_conv_igemm(acts, len(srcs), _C.ptr(w_bf16), cout, len(dh), _C.int_array(dh),
                                      _C.int_array(dw), _C.ptr(out), out_f32, ld_out, n_store, _C.ptr(bias),
                                      _C.ptr(addend), ld_add, _C.ptr(stat_sum), _C.ptr(stat_sqsum), _stream()))
    return out
