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
    return _C.Act(_C.c_void_p(t.data_ptr()), n, h, w, c, t.stride(2), t.stride(1), t.stride(0))


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
    o = act(out[..., :n_store])
    a = act(addend) if addend is not None else None
    _C.check(_C.lib().sseg_conv_igemm(geom, _C.ptr(w_bf16), w_bf16.stride(0), cout, o, out_f32, _C.ptr(bias), a,
                                      _C.ptr(stat_sum), _C.ptr(stat_sqsum), _stream()))
    return out


RELU_NONE, RELU_AFTER_ADD, RELU_BEFORE_ADD = 0, 1, 2
RELU6 = 4   # or-ed to one of the above: clip at 6 (nn.ReLU6)


def conv_igemm_affine(geom, w_bf16, cout, out, scale, shift, relu=RELU_AFTER_ADD, addend=None):
    """out = relu?(conv(geom, w) * scale + shift (+ addend)) in one kernel (eval-mode BN folded into the epilogue)."""
    assert out.dtype == torch.bfloat16 and w_bf16.dim() == 2 and w_bf16.stride(1) == 1
    n_store = (cout + 7) // 8 * 8
    o = act(out[..., :n_store])
    a = act(addend) if addend is not None else None
    _C.check(_C.lib().sseg_conv_igemm_affine(geom, _C.ptr(w_bf16), w_bf16.stride(0), cout, o, _C.ptr(scale), _C.ptr(shift),
                                             int(relu), a, _stream()))
    return out


def make_coop_peer(arena, data_off, data_stride, flag_off, step):
    """sseg_coop_peer_t over a PeerArena (engine/peer.py); keep the returned object alive while it is in use."""
    p = _C.CoopPeer()
    p.bases = _C.ctypes.cast(arena.bases, _C.POINTER(_C.c_void_p))
    p.world, p.rank = arena.world, arena.rank
    p.data_off, p.data_stride, p.flag_off = int(data_off), int(data_stride), int(flag_off)
    p.step = _C.ptr(step)
    p._keep = (arena, step)
    return p


def bn_running_from_tmp(tmp_mean, tmp_var, running_iter, running_mean, running_var):
    _C.check(_C.lib().sseg_bn_running_from_tmp(_C.ptr(tmp_mean), _C.ptr(tmp_var), _C.ptr(running_iter), _C.ptr(running_mean),
                                               _C.ptr(running_var), running_mean.numel(), _stream()))


def make_bn_fused(gamma, beta, eps, momentum, count, stat_sum, stat_sqsum, counter, mean, invstd, scale, shift,
                  running_mean=None, running_var=None, res=None, rscale=None, rshift=None, chanmul=None, relu=True,
                  res_after_relu=False, peer=None, tmp_running_mean=None, tmp_running_var=None, running_iter=None,
                  count_out=None):
    """sseg_bn_fused_t for conv_bn_train (keep the returned object and the tensors alive while it is in use)."""
    b = _C.BnFused()
    b.gamma, b.beta = _C.ptr(gamma), _C.ptr(beta)
    b.eps, b.momentum, b.count = float(eps), float(momentum), float(count)
    b.stat_sum, b.stat_sqsum, b.counter = _C.ptr(stat_sum), _C.ptr(stat_sqsum), _C.ptr(counter)
    b.mean_out, b.invstd_out, b.scale_out, b.shift_out = _C.ptr(mean), _C.ptr(invstd), _C.ptr(scale), _C.ptr(shift)
    b.running_mean, b.running_var = _C.ptr(running_mean), _C.ptr(running_var)
    b._res_act = act(res) if res is not None else None   # the struct only holds a pointer to it
    b.res = _C.ctypes.pointer(b._res_act) if res is not None else None
    b.rscale, b.rshift, b.chanmul = _C.ptr(rscale), _C.ptr(rshift), _C.ptr(chanmul)
    b.relu, b.res_after_relu = int(relu), int(res_after_relu)
    b._peer = peer
    b.peer = _C.ctypes.pointer(peer) if peer is not None else None
    b.tmp_running_mean, b.tmp_running_var = _C.ptr(tmp_running_mean), _C.ptr(tmp_running_var)
    b.running_iter, b.count_out = _C.ptr(running_iter), _C.ptr(count_out)
    return b


def conv_igemm_bnfin(geom, w_bf16, cout, y, bn):
    """conv -> y (bf16) with the BatchNorm statistics and their finalisation (train mode, one GPU) in the same launch: the
    last CTA to finish writes mean / inv_std / scale / shift and the running statistics. bn: make_bn_fused(...)."""
    n_store = (cout + 7) // 8 * 8
    _C.check(_C.lib().sseg_conv_igemm_bnfin(geom, _C.ptr(w_bf16), w_bf16.stride(0), cout, act(y[..., :n_store]), bn, _stream()))
    return y


def conv_bn_train(geom, w_bf16, cout, y, a_out, bn):
    """conv + train-mode BN (+shortcut, ReLU, dropout mask) in one kernel; y may be None. bn: make_bn_fused(...)."""
    n_store = (cout + 7) // 8 * 8
    yo = act(y[..., :n_store]) if y is not None else None
    _C.check(_C.lib().sseg_conv_bn_train(geom, _C.ptr(w_bf16), w_bf16.stride(0), cout, yo, act(a_out[..., :n_store]), bn,
                                         _stream()))
    return a_out


def conv_bn_train_fits(geom, w_bf16, cout, y, a_out, bn):
    """True when the layer's tiles fit the tensor memory of one persistent CTA per SM (needs a CUDA device)."""
    n_store = (cout + 7) // 8 * 8
    yo = act(y[..., :n_store]) if y is not None else None
    rc = _C.lib().sseg_conv_bn_train_fits(geom, _C.ptr(w_bf16), w_bf16.stride(0), cout, yo, act(a_out[..., :n_store]), bn)
    if rc < 0:
        _C.check(rc)
    return rc == 1


def conv_dgrad_bn(geom, w_bf16, cout, y, dy_out, fscale, fshift, mean, invstd, count, s1, s2_raw, dgamma_out, counter,
                  query=False, peer=None, count_dev=None, dbeta_out=None):
    """Data gradient of the consumer conv + the producer layer's whole BN backward in one kernel: only dy_out is written.
    query=True: returns whether the layer fits (no launch)."""
    n_store = (cout + 7) // 8 * 8
    args = (geom, _C.ptr(w_bf16), w_bf16.stride(0), cout, act(y), act(dy_out[..., :n_store]), _C.ptr(fscale), _C.ptr(fshift),
            _C.ptr(mean), _C.ptr(invstd), float(count), _C.ptr(s1), _C.ptr(s2_raw), _C.ptr(dgamma_out), _C.ptr(counter))
    if query:
        rc = _C.lib().sseg_conv_dgrad_bn_fits(*args)
        if rc < 0:
            _C.check(rc)
        return rc == 1
    _C.check(_C.lib().sseg_conv_dgrad_bn(*args, peer, _C.ptr(count_dev), _C.ptr(dbeta_out), _stream()))
    return dy_out


def stem_conv_affine(img, w, out, scale=None, shift=None, relu6=True):
    """3x3 stride-2 conv from the fp32 NCHW image (3 -> cout <= 64) + eval-BN affine + ReLU6 -> bf16 NHWC."""
    n, c, h, w_ = img.shape
    cout = w.shape[0]
    assert c == 3 and img.is_contiguous() and w.is_contiguous() and out.is_contiguous() and out.shape[3] == cout
    _C.check(_C.lib().sseg_stem_conv_affine(_C.ptr(img), n, h, w_, _C.ptr(w), cout, _C.ptr(scale), _C.ptr(shift), int(relu6),
                                            _C.ptr(out), _stream()))
    return out


def dwconv_affine(x, w, out, stride=1, dilation=1, scale=None, shift=None, relu6=True):
    """Depthwise 3x3 conv ('same' padding = dilation) + eval-BN affine + ReLU6; x/out dense NHWC bf16, w fp32 [C,1,3,3]."""
    n, h, w_, c = x.shape
    assert x.is_contiguous() and out.is_contiguous() and w.is_contiguous() and w.dtype == torch.float32 and w.shape[0] == c
    assert out.shape == (n, (h - 1) // stride + 1, (w_ - 1) // stride + 1, c)
    _C.check(_C.lib().sseg_dwconv_affine(_C.ptr(x), n, h, w_, c, _C.ptr(w), stride, dilation, _C.ptr(scale), _C.ptr(shift),
                                         int(relu6), _C.ptr(out), _stream()))
    return out


def conv_igemm_bnbwd(geom, w_bf16, cout, out, y, fscale, fshift, s1, s2_raw, addend=None):
    """Data gradient into `out` + the producer layer's BN-backward partial sums (see sseg_conv_igemm_bnbwd)."""
    assert out.dtype == torch.bfloat16 and w_bf16.dim() == 2 and w_bf16.stride(1) == 1
    n_store = (cout + 7) // 8 * 8
    o = act(out[..., :n_store])
    a = act(addend) if addend is not None else None
    _C.check(_C.lib().sseg_conv_igemm_bnbwd(geom, _C.ptr(w_bf16), w_bf16.stride(0), cout, o, a, act(y), _C.ptr(fscale),
                                            _C.ptr(fshift), _C.ptr(s1), _C.ptr(s2_raw), _stream()))
    return out


def conv_igemm_bnbwd_res(geom, w_bf16, cout, out, y, a_saved, s1, s2_raw, addend=None):
    """Data gradient into `out` (the COMPLETE gradient of the producer's output) + the producer's BN-backward partial sums
    with the ReLU mask taken from its saved output `a_saved` (shortcut layers; see sseg_conv_igemm_bnbwd_res)."""
    assert out.dtype == torch.bfloat16 and w_bf16.dim() == 2 and w_bf16.stride(1) == 1
    n_store = (cout + 7) // 8 * 8
    o = act(out[..., :n_store])
    ad = act(addend) if addend is not None else None
    _C.check(_C.lib().sseg_conv_igemm_bnbwd_res(geom, _C.ptr(w_bf16), w_bf16.stride(0), cout, o, ad, act(y), act(a_saved),
                                                _C.ptr(s1), _C.ptr(s2_raw), _stream()))
    return out


def conv_wgrad(geom, dy, cout, dw):
    """dw[co, koff_t + ci] += sum_pixels dy[.., co] * x_t[.., ci]; dw: fp32 2-D [cout, K], pre-zeroed by the caller."""
    assert dw.dtype == torch.float32 and dw.dim() == 2 and dw.stride(1) == 1
    a = act(dy)
    _C.check(_C.lib().sseg_conv_wgrad(geom, a, cout, _C.ptr(dw), dw.stride(0), _stream()))
    return dw


# ---------------------------------------------------------------------------------------------------------
# helpers for dense [pixels][ld] views
def _pix(t):
    """(P, pixels-per-image, ld) of a pixel-dense NHWC tensor (channel slices allowed)."""
    n, h, w, _ = t.shape
    ld = t.stride(2)
    assert t.stride(3) == 1 and t.stride(1) == w * ld and t.stride(0) == h * w * ld, "expected pixel-dense NHWC"
    return n * h * w, h * w, ld


def conv_s2_taps(ksize):
    """Stride-2 'same' conv (pad = ksize//2) over the 4 parity planes x[:, hp::2, wp::2]: per tap (dh, dw, plane)."""
    r = ksize // 2
    dh, dw, src = [], [], []
    for i in range(ksize):
        for j in range(ksize):
            oh, ow = i - r, j - r  # input row = 2*ho + oh
            hp, wp = oh % 2, ow % 2
            dh.append((oh - hp) // 2)
            dw.append((ow - wp) // 2)
            src.append(hp * 2 + wp)
    return dh, dw, src


def parity_planes(x):
    """The 4 strided views x[:, hp::2, wp::2, :] (no copy), index hp*2+wp."""
    return [x[:, hp::2, wp::2, :] for hp in (0, 1) for wp in (0, 1)]


def prep_conv_weight(w, w_fwd=None, w_dgrad=None, o_pad=None):
    """fp32 OIHW -> bf16 [O, T*I] and/or bf16 [I, T*o_pad] (w_dgrad must have been zero-initialised)."""
    O, I, kh, kw = w.shape
    T = kh * kw
    assert w.dtype == torch.float32
    w = w.contiguous()   # a channels-last master (engine/program.py::ConvW) is re-packed to OIHW for this single-conv entry
    if o_pad is None:
        o_pad = (O + 63) // 64 * 64
    _C.check(_C.lib().sseg_prep_conv_weight(_C.ptr(w), O, I, T, _C.ptr(w_fwd), w_fwd.stride(0) if w_fwd is not None else 0,
                                            _C.ptr(w_dgrad), w_dgrad.stride(0) if w_dgrad is not None else 0, o_pad,
                                            _stream()))


def grad_to_oihw(g, O, I, T, out, scale=1.0, accumulate=False):
    _C.check(_C.lib().sseg_grad_to_oihw(_C.ptr(g), g.stride(0), O, I, T, _C.ptr(out), scale, int(accumulate), _stream()))


def stem_conv_fwd(img, w, out, stat_sum=None, stat_sqsum=None):
    n, c, h, w_ = img.shape
    assert c == 3 and img.is_contiguous() and img.dtype == torch.float32 and w.shape == (64, 3, 3, 3)
    _C.check(_C.lib().sseg_stem_conv_fwd(_C.ptr(img), n, h, w_, _C.ptr(w), _C.ptr(out), _C.ptr(stat_sum),
                                         _C.ptr(stat_sqsum), _stream()))


def stem_conv_wgrad(img, dy, dw):
    n, c, h, w_ = img.shape
    assert dy.is_contiguous() and dw.is_contiguous()
    _C.check(_C.lib().sseg_stem_conv_wgrad(_C.ptr(img), n, h, w_, _C.ptr(dy), _C.ptr(dw), _stream()))


BN_TRAIN, BN_TRAIN_SYNC, BN_EVAL = 0, 1, 2


def bn_finalize(ssum, ssq, count, gamma, beta, eps, momentum, mode, mean, invstd, scale, shift, running=None,
                update_running=False, count_dev=None):
    """running = (running_mean, running_var, tmp_running_mean, tmp_running_var, running_iter) or None."""
    rm = rv = tm = tv = it = None
    if running is not None:
        rm, rv, tm, tv, it = running
    _C.check(_C.lib().sseg_bn_finalize(_C.ptr(ssum), _C.ptr(ssq), _C.ptr(count_dev), float(count), _C.ptr(gamma),
                                       _C.ptr(beta), eps, momentum, mode, int(update_running), _C.ptr(rm), _C.ptr(rv),
                                       _C.ptr(tm), _C.ptr(tv), _C.ptr(it), _C.ptr(mean), _C.ptr(invstd), _C.ptr(scale),
                                       _C.ptr(shift), scale.numel(), _stream()))


def peer_step(step):
    _C.check(_C.lib().sseg_peer_step(_C.ptr(step), _stream()))


def bn_finalize_peer(arena, stats_off, flag_off, step, gamma, beta, eps, momentum, mean, invstd, scale, shift, count_out,
                     running=None, update_running=False, defer_running=False, ll=False):
    """defer_running: advance the accumulators only; the caller refreshes running_mean / running_var later
    (bn_running_from_tmp)."""
    rm = rv = tm = tv = it = None
    if running is not None:
        rm, rv, tm, tv, it = running
    upd = (2 if defer_running else 1) if update_running else 0
    fn = _C.lib().sseg_bn_finalize_peer_ll if ll else _C.lib().sseg_bn_finalize_peer   # ll: flag_off is the inbox offset
    _C.check(fn(arena.bases, arena.world, arena.rank, stats_off, flag_off, _C.ptr(step), _C.ptr(gamma), _C.ptr(beta), eps,
                momentum, upd, _C.ptr(rm), _C.ptr(rv), _C.ptr(tm), _C.ptr(tv), _C.ptr(it), _C.ptr(mean), _C.ptr(invstd),
                _C.ptr(scale), _C.ptr(shift), _C.ptr(count_out), scale.numel(), _stream()))


def bn_bwd_peer_sum(arena, part_off, flag_off, step, s1_tot, s2_tot, dbeta, dgamma, mean=None, invstd=None, s2_raw=False,
                    ll=False):
    fn = _C.lib().sseg_bn_bwd_peer_sum_ll if ll else _C.lib().sseg_bn_bwd_peer_sum   # ll: flag_off is the inbox offset
    _C.check(fn(arena.bases, arena.world, arena.rank, part_off, flag_off, _C.ptr(step), _C.ptr(s1_tot), _C.ptr(s2_tot),
                _C.ptr(dbeta), _C.ptr(dgamma), _C.ptr(mean), _C.ptr(invstd), int(s2_raw), s1_tot.numel(), _stream()))


def bn_apply(y, scale, shift, out, relu=True, res=None, rscale=None, rshift=None, chanmul=None, res_after_relu=False):
    P, ppi, y_ld = _pix(y)
    _, _, out_ld = _pix(out)
    res_ld = _pix(res)[2] if res is not None else 0
    _C.check(_C.lib().sseg_bn_apply(_C.ptr(y), y_ld, _C.ptr(scale), _C.ptr(shift), _C.ptr(res), res_ld, _C.ptr(rscale),
                                    _C.ptr(rshift), _C.ptr(chanmul), _C.ptr(out), out_ld, P, ppi, y.shape[3], int(relu),
                                    int(res_after_relu), _stream()))
    return out


def bn_finalize_apply(ssum, ssq, count, gamma, beta, eps, momentum, mean, invstd, scale, shift, y, out, relu=True,
                      res=None, rscale=None, rshift=None, chanmul=None, res_after_relu=False, running_mean=None,
                      running_var=None):
    """bn_finalize(BN_TRAIN) + bn_apply in one launch."""
    P, ppi, y_ld = _pix(y)
    _, _, out_ld = _pix(out)
    res_ld = _pix(res)[2] if res is not None else 0
    _C.check(_C.lib().sseg_bn_finalize_apply(_C.ptr(ssum), _C.ptr(ssq), float(count), _C.ptr(gamma), _C.ptr(beta), eps,
                                             momentum, _C.ptr(running_mean), _C.ptr(running_var), _C.ptr(mean),
                                             _C.ptr(invstd), _C.ptr(scale), _C.ptr(shift), _C.ptr(y), y_ld, _C.ptr(res),
                                             res_ld, _C.ptr(rscale), _C.ptr(rshift), _C.ptr(chanmul), _C.ptr(out), out_ld,
                                             P, ppi, y.shape[3], int(relu), int(res_after_relu), _stream()))
    return out


def bn_bwd_reduce(g, a, y, mean, invstd, s1, s2, chanmul=None, scale=None, fshift=None):
    """a: saved output (ReLU mask) or None; (scale, fshift) with a=None: mask recomputed from y; all None: no ReLU."""
    P, ppi, g_ld = _pix(g)
    a_ld = _pix(a)[2] if a is not None else 0
    _C.check(_C.lib().sseg_bn_bwd_reduce(_C.ptr(g), g_ld, _C.ptr(a), a_ld, _C.ptr(y), _pix(y)[2], _C.ptr(mean),
                                         _C.ptr(invstd), _C.ptr(scale), _C.ptr(fshift), _C.ptr(chanmul), _C.ptr(s1),
                                         _C.ptr(s2), P, ppi, g.shape[3], _stream()))


def bn_bwd_apply(g, a, y, mean, invstd, scale, s1, s2, count, dy, dres=None, chanmul=None, eval_mode=False,
                 count_dev=None, fshift=None, s2_raw=False, dgamma_out=None):
    P, ppi, g_ld = _pix(g)
    a_ld = _pix(a)[2] if a is not None else 0
    y_ld = _pix(y)[2] if y is not None else 0
    dres_ld = _pix(dres)[2] if dres is not None else 0
    _C.check(_C.lib().sseg_bn_bwd_apply(_C.ptr(g), g_ld, _C.ptr(a), a_ld, _C.ptr(y), y_ld, _C.ptr(mean), _C.ptr(invstd),
                                        _C.ptr(scale), _C.ptr(fshift), _C.ptr(chanmul), _C.ptr(s1), _C.ptr(s2),
                                        _C.ptr(count_dev), float(count), _C.ptr(dy), _pix(dy)[2], _C.ptr(dres), dres_ld, P,
                                        ppi, g.shape[3], int(eval_mode), int(s2_raw), _C.ptr(dgamma_out), _stream()))


def bn_bwd_apply_peer(arena, part_off, flag_off, step, g, a, y, mean, invstd, scale, count_dev, dy, dbeta, dgamma, dres=None,
                      chanmul=None, fshift=None, s2_raw=False):
    """bn_bwd_peer_sum + bn_bwd_apply in one launch (multi-GPU SyncBN backward; see sseg_bn_bwd_apply_peer)."""
    P, ppi, g_ld = _pix(g)
    a_ld = _pix(a)[2] if a is not None else 0
    dres_ld = _pix(dres)[2] if dres is not None else 0
    _C.check(_C.lib().sseg_bn_bwd_apply_peer(arena.bases, arena.world, arena.rank, part_off, flag_off, _C.ptr(step), _C.ptr(g),
                                             g_ld, _C.ptr(a), a_ld, _C.ptr(y), _pix(y)[2], _C.ptr(mean), _C.ptr(invstd),
                                             _C.ptr(scale), _C.ptr(fshift), _C.ptr(chanmul), _C.ptr(count_dev), _C.ptr(dy),
                                             _pix(dy)[2], _C.ptr(dres), dres_ld, P, ppi, g.shape[3], int(s2_raw),
                                             _C.ptr(dbeta), _C.ptr(dgamma), _stream()))


def maxpool_fwd(x, out, idx):
    n, h, w, c = x.shape
    assert x.is_contiguous() and out.is_contiguous()
    _C.check(_C.lib().sseg_maxpool_fwd(_C.ptr(x), n, h, w, c, _C.ptr(out), _C.ptr(idx), _stream()))


def maxpool_bwd(dout, idx, dx):
    n, h, w, c = dx.shape
    assert dout.is_contiguous() and dx.is_contiguous()
    _C.check(_C.lib().sseg_maxpool_bwd(_C.ptr(dout), _C.ptr(idx), _C.ptr(dx), n, h, w, c, _stream()))


def avgpool_fwd(x, S, out):
    n, h, w, c = x.shape
    assert out.is_contiguous() and out.shape == (n, S, S, c)
    _C.check(_C.lib().sseg_avgpool_fwd(_C.ptr(x), _pix(x)[2], n, h, w, c, S, _C.ptr(out), _stream()))


def avgpool_bwd(base, dpools, scales, dx):
    n, h, w, c = dx.shape
    arr = (_C.c_void_p * len(dpools))(*[d.data_ptr() for d in dpools])
    _C.check(_C.lib().sseg_avgpool_bwd(_C.ptr(base), _pix(base)[2] if base is not None else 0, arr,
                                       _C.int_array(list(scales)), len(dpools), _C.ptr(dx), _pix(dx)[2], n, h, w, c,
                                       _stream()))


def bilinear_fwd(x, out):
    n, hi, wi, c = x.shape
    _, ho, wo, _ = out.shape
    _C.check(_C.lib().sseg_bilinear_fwd(_C.ptr(x), _pix(x)[2], n, hi, wi, c, _C.ptr(out), _pix(out)[2], ho, wo, _stream()))


def bilinear_bwd(dout, dx, accumulate=False, scratch=None):
    """scratch: fp32 tensor with >= N*Ho*Wi*C elements (allocated here when omitted - pass one on hot paths)."""
    n, ho, wo, c = dout.shape
    _, hi, wi, _ = dx.shape
    if scratch is None:
        scratch = torch.empty(n * ho * wi * c, device=dout.device, dtype=torch.float32)
    assert scratch.numel() >= n * ho * wi * c and scratch.dtype == torch.float32
    _C.check(_C.lib().sseg_bilinear_bwd(_C.ptr(dout), _pix(dout)[2], n, ho, wo, c, _C.ptr(dx), _pix(dx)[2], hi, wi,
                                        int(accumulate), _C.ptr(scratch), _stream()))


def make_sum_terms(terms):
    """[(x NHWC bf16, scale or None, shift or None), ...] -> ctypes array of sseg_sum_term_t (keep the tensors alive)."""
    arr = (_C.SumTerm * len(terms))()
    for k, (x, scale, shift) in enumerate(terms):
        _, _, ld = _pix(x)
        arr[k].x, arr[k].h, arr[k].w, arr[k].ld = x.data_ptr(), x.shape[1], x.shape[2], ld
        arr[k].scale = scale.data_ptr() if scale is not None else None
        arr[k].shift = shift.data_ptr() if shift is not None else None
    return arr


def sum_terms(terms, out, relu=True):
    """out = relu?(sum_k scale_k * resample(x_k) + shift_k); `terms` from make_sum_terms or a list of tuples."""
    if isinstance(terms, (list, tuple)):
        terms = make_sum_terms(terms)
    n, ho, wo, c = out.shape
    _C.check(_C.lib().sseg_sum_terms(terms, len(terms), n, ho, wo, c, _C.ptr(out), _pix(out)[2], int(relu), _stream()))
    return out


def relu_mask_bwd(g, out, ds, acc_out=None, accumulate=False):
    """ds = g * [out > 0]; acc_out (+)= ds."""
    P, _, g_ld = _pix(g)
    _C.check(_C.lib().sseg_relu_mask_bwd(_C.ptr(g), g_ld, _C.ptr(out), _pix(out)[2], _C.ptr(ds), _pix(ds)[2],
                                         _C.ptr(acc_out), _pix(acc_out)[2] if acc_out is not None else 0,
                                         int(accumulate), P, g.shape[3], _stream()))
    return ds


# ---------------------------------------------------------------------------------------------------------
# fp32-accurate inference: activations / weights as (hi, lo) bf16 pairs (csrc/accurate.cu)
def split_affine(z, out_hi, out_lo, scale=None, shift=None, res=None, relu=True, res_after_relu=False):
    """(out_hi, out_lo) = split(relu?(z*scale + shift (+ res_hi + res_lo))); z: fp32 NHWC view, res: (hi, lo) or None."""
    assert z.dtype == torch.float32
    rh, rl = res if res is not None else (None, None)
    _C.check(_C.lib().sseg_split_affine(act(z), _C.ptr(scale), _C.ptr(shift), _C.ptr(rh), _C.ptr(rl),
                                        _pix(rh)[2] if rh is not None else 0, _C.ptr(out_hi), _C.ptr(out_lo), _pix(out_hi)[2],
                                        int(relu), int(res_after_relu), _stream()))


def stem_conv_fwd_f32(img, w, out):
    n, c, h, w_ = img.shape
    assert c == 3 and img.is_contiguous() and out.dtype == torch.float32 and out.is_contiguous()
    _C.check(_C.lib().sseg_stem_conv_fwd_f32(_C.ptr(img), n, h, w_, _C.ptr(w), _C.ptr(out), _stream()))


def maxpool_pair_fwd(x, out):
    n, h, w, c = x[0].shape
    assert all(t.is_contiguous() for t in x + out)
    _C.check(_C.lib().sseg_maxpool_pair_fwd(_C.ptr(x[0]), _C.ptr(x[1]), n, h, w, c, _C.ptr(out[0]), _C.ptr(out[1]), _stream()))


def avgpool_pair_fwd(x, S, out):
    n, h, w, c = x[0].shape
    assert out[0].is_contiguous() and out[1].is_contiguous() and out[0].shape == (n, S, S, c)
    _C.check(_C.lib().sseg_avgpool_pair_fwd(_C.ptr(x[0]), _C.ptr(x[1]), _pix(x[0])[2], n, h, w, c, S, _C.ptr(out[0]),
                                            _C.ptr(out[1]), _stream()))


def bilinear_pair_fwd(x, out):
    n, hi, wi, c = x[0].shape
    _, ho, wo, _ = out[0].shape
    _C.check(_C.lib().sseg_bilinear_pair_fwd(_C.ptr(x[0]), _C.ptr(x[1]), _pix(x[0])[2], n, hi, wi, c, _C.ptr(out[0]),
                                             _C.ptr(out[1]), _pix(out[0])[2], ho, wo, _stream()))


def prep_conv_weight_split(w, out):
    O, I, kh, kw = w.shape
    assert w.dtype == torch.float32 and out.dtype == torch.bfloat16
    w = w.contiguous()   # (a channels-last master is re-packed to OIHW first)
    _C.check(_C.lib().sseg_prep_conv_weight_split(_C.ptr(w), O, I, kh * kw, _C.ptr(out), out.stride(0), _stream()))


def softmax_nll_fwd(logits, num_class, label, lse, accum):
    P, _, ld = _pix(logits)
    assert label.dtype == torch.int64 and label.is_contiguous() and label.numel() == P
    _C.check(_C.lib().sseg_softmax_nll_fwd(_C.ptr(logits), ld, num_class, _C.ptr(label), P, _C.ptr(lse), _C.ptr(accum),
                                           _stream()))


def nll_finalize(accum_main, accum_ds, ds_scale, out):
    _C.check(_C.lib().sseg_nll_finalize(_C.ptr(accum_main), _C.ptr(accum_ds), float(ds_scale or 0.0), _C.ptr(out),
                                        _stream()))


def softmax_nll_bwd(logits, num_class, label, lse, accum, weight, dlogits):
    P, _, ld = _pix(logits)
    _C.check(_C.lib().sseg_softmax_nll_bwd(_C.ptr(logits), ld, num_class, _C.ptr(label), _C.ptr(lse), _C.ptr(accum),
                                           float(weight), P, _C.ptr(dlogits), _pix(dlogits)[2], dlogits.shape[3],
                                           _stream()))


def scale_by_scalar(x, scalar_dev):
    """x (fp32, contiguous, 16-byte aligned) *= the device scalar; free when the scalar is exactly 1."""
    assert x.dtype == torch.float32 and x.is_contiguous() and scalar_dev.dtype == torch.float32
    _C.check(_C.lib().sseg_scale_by_scalar(_C.ptr(x), x.numel(), _C.ptr(scalar_dev), _stream()))


def colsum(x, C, out):
    P, _, ld = _pix(x)
    _C.check(_C.lib().sseg_colsum(_C.ptr(x), ld, P, C, _C.ptr(out), _stream()))


def upsample_softmax(logits, num_class, probs, weight=1.0, accumulate=False, log_output=False):
    n, hi, wi, _ = logits.shape
    _, c, ho, wo = probs.shape
    assert c == num_class and probs.is_contiguous() and probs.dtype == torch.float32
    _C.check(_C.lib().sseg_upsample_softmax(_C.ptr(logits), _pix(logits)[2], n, hi, wi, num_class, _C.ptr(probs), ho, wo,
                                            float(weight), int(accumulate), int(log_output), _stream()))


def nhwc_bf16_to_nchw_f32(x, out):
    n, h, w, c = x.shape
    _C.check(_C.lib().sseg_nhwc_bf16_to_nchw_f32(_C.ptr(x), _pix(x)[2], n, h, w, c, _C.ptr(out), _stream()))


def nchw_f32_to_nhwc_bf16(x, out):
    n, c, h, w = x.shape
    _C.check(_C.lib().sseg_nchw_f32_to_nhwc_bf16(_C.ptr(x), n, h, w, c, _C.ptr(out), _pix(out)[2], _stream()))


def image_transform(img_u8, valid_hw, out, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """uint8 [N,H,W,3] -> normalised fp32 [N,3,H,W] (the reference's img_transform, dataset.py:53-58), zeros outside each
    image's valid rows / columns (valid_hw: device int32 [N,2])."""
    n, h, w, c = img_u8.shape
    assert c == 3 and img_u8.dtype == torch.uint8 and img_u8.is_contiguous() and out.is_contiguous()
    assert tuple(out.shape) == (n, 3, h, w) and out.dtype == torch.float32 and valid_hw.dtype == torch.int32
    ms = (_C.c_float * 6)(*[float(torch.tensor(v, dtype=torch.float32)) for v in tuple(mean) + tuple(std)])
    _C.check(_C.lib().sseg_image_transform(_C.ptr(img_u8), n, h, w, _C.ptr(valid_hw), ms, _C.ptr(out), _stream()))
    return out


def label_transform(seg_u8, valid_hw, rate, out):
    """uint8 [N,Hs,Ws] stored ids -> int64 labels id - 1 (the reference's segm_transform, dataset.py:60-63); 0 in the padding."""
    n, hs, ws = seg_u8.shape
    assert seg_u8.dtype == torch.uint8 and seg_u8.is_contiguous() and out.is_contiguous() and out.dtype == torch.int64
    assert tuple(out.shape) == (n, hs, ws) and valid_hw.dtype == torch.int32
    _C.check(_C.lib().sseg_label_transform(_C.ptr(seg_u8), n, hs, ws, _C.ptr(valid_hw), int(rate), _C.ptr(out), _stream()))
    return out


class WeightTable:
    """Device table of sseg_weight_desc_t for the batched weight re-layout / gradient re-layout kernels.
    entries: list of dicts with keys w, wf, wd, g_src, g_dst (tensors or None), O, I, T, o_pad."""

    def __init__(self, entries, device):
        import ctypes
        n = len(entries)
        arr = (_C.WeightDesc * n)()
        tiles = 0
        for k, e in enumerate(entries):
            d = arr[k]
            for name in ("w", "wf", "wd", "g_src", "g_dst"):
                t = e.get(name)
                setattr(d, name, t.data_ptr() if t is not None else None)
            d.fwd_ld = e["wf"].stride(0) if e.get("wf") is not None else 0
            d.dgrad_ld = e["wd"].stride(0) if e.get("wd") is not None else 0
            d.g_ld = e["g_src"].stride(0) if e.get("g_src") is not None else 0
            d.O, d.I, d.T, d.o_pad = e["O"], e["I"], e["T"], e["o_pad"]
            assert d.T <= 9
            w = e.get("w")
            if e.get("channels_last") and d.T > 1:
                # master weight [O][kh][kw][I] in memory (torch.channels_last): the kernel reads rows of I
                assert w is not None and w.is_contiguous(memory_format=torch.channels_last), "expected a channels-last master"
                assert e.get("g_dst") is None, "channels-last masters need no gradient re-layout"
                d.reserved = 1
            else:
                assert w is None or w.is_contiguous(), "expected an OIHW-contiguous master weight"
                d.reserved = 0
            d.first_tile = tiles
            rep = 8 if d.T == 1 else 1  # i-tiles per CTA, must match i_tiles_per_cta() in csrc/weights.cu
            tiles += ((d.O + 31) // 32) * ((((d.I + 31) // 32) + rep - 1) // rep)
        raw = bytes(arr)
        self.n, self.tiles = n, tiles
        self.dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
        self.keep = entries

    def prep(self, max_blocks=0):
        """max_blocks > 0: a thin grid (blocks walk the tiles) - for a pass that runs next to the main stream's kernels"""
        if max_blocks > 0:
            _C.check(_C.lib().sseg_prep_conv_weights_batched_ex(_C.ptr(self.dev), self.n, self.tiles, int(max_blocks), _stream()))
        else:
            _C.check(_C.lib().sseg_prep_conv_weights_batched(_C.ptr(self.dev), self.n, self.tiles, _stream()))

    def grads(self, scale=1.0):
        _C.check(_C.lib().sseg_grads_to_oihw_batched(_C.ptr(self.dev), self.n, self.tiles, float(scale), _stream()))
