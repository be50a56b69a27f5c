"""TEST INFRASTRUCTURE ONLY — a Python restatement of the C ABI's documented semantics (include/sseg_b200.h).

`EmuLib` stands in for `ctypes.CDLL("libsseg_b200.so")` in CPU tests: every entry point a step program calls receives
the same raw pointers / sizes / structs the real library would get, re-creates tensor views over that memory and
performs the documented operation with torch CPU ops (bf16 storage, fp32 accumulation, rounding at the stores the
kernels round at). It exists so that the SCHEDULES (engine/program.py: which kernel, which buffers, which order, which
coefficient vector goes where) can be checked numerically against the oracle without a GPU — including the opt-in
schedules that have not run on hardware yet. It says nothing about the CUDA kernels themselves (the `-m gpu` tests do
that) and nothing in the product imports it: there is no CPU execution path in the engine.
"""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

_CT = {torch.bfloat16: ctypes.c_uint16, torch.float32: ctypes.c_float, torch.int64: ctypes.c_int64,
       torch.uint8: ctypes.c_uint8, torch.int32: ctypes.c_int32}


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    v = p.value if hasattr(p, "value") else ctypes.cast(p, ctypes.c_void_p).value
    return v or 0


def flat(p, span, dtype):
    """1-D tensor sharing the `span` elements of type `dtype` that start at raw address p."""
    a = _addr(p)
    assert a != 0 and span > 0
    arr = np.ctypeslib.as_array((_CT[dtype] * int(span)).from_address(a))
    t = torch.from_numpy(arr)
    return t.view(torch.bfloat16) if dtype == torch.bfloat16 else t


def vec(p, n):
    return None if _addr(p) == 0 else flat(p, n, torch.float32)


def act_view(a, dtype=torch.bfloat16, c=None):
    """sseg_act_t -> strided [n,h,w,c] view."""
    c = a.c if c is None else c
    span = (a.n - 1) * a.img_stride + (a.h - 1) * a.row_stride + (a.w - 1) * a.ld + c
    return flat(a.ptr, span, dtype).as_strided((a.n, a.h, a.w, c), (a.img_stride, a.row_stride, a.ld, 1))


def pix_view(p, ld, P, C, dtype=torch.bfloat16):
    """[P][ld] pixel-dense matrix -> [P, C] view."""
    return flat(p, (P - 1) * ld + C, dtype).as_strided((P, C), (ld, 1))


def _shift(x, dh, dw):
    """y[n,h,w] = x[n,h+dh,w+dw], zeros outside (the TMA out-of-bounds fill = conv padding)."""
    n, h, w, c = x.shape
    ph, pw = abs(dh), abs(dw)
    xp = F.pad(x, (0, 0, pw, pw, ph, ph))
    return xp[:, ph + dh:ph + dh + h, pw + dw:pw + dw + w, :]


def _taps(g):
    srcs = [act_view(g.srcs[i]).float() for i in range(g.nsrc)]
    cat = torch.cat(srcs, 3) if g.nsrc > 1 else srcs[0]
    for t in range(g.ntaps):
        x = srcs[g.tap_src[t]] if g.tap_src[t] >= 0 else cat
        yield _shift(x, g.tap_dh[t], g.tap_dw[t]), g.tap_koff[t]


def _conv(g, w_ptr, w_ld, cout):
    a0 = g.srcs[0]
    out = torch.zeros(a0.n, a0.h, a0.w, cout)
    kmax = max(g.tap_koff[t] for t in range(g.ntaps)) + sum(g.srcs[i].c for i in range(g.nsrc))
    W = flat(w_ptr, (cout - 1) * w_ld + min(w_ld, kmax), torch.bfloat16).as_strided((cout, min(w_ld, kmax)), (w_ld, 1)).float()
    for xs, koff in _taps(g):
        cin = xs.shape[3]
        out += xs @ W[:, koff:koff + cin].t()
    return out


def _bf(x):
    return x.to(torch.bfloat16)


def _store(out_act, val, cout, dtype=torch.bfloat16):
    """channels [0,cout) <- val, channels [cout, out.c) <- 0"""
    o = act_view(out_act, dtype)
    o[..., :cout] = val.to(dtype)
    if out_act.c > cout:
        o[..., cout:] = 0


class EmuLib:
    def __init__(self):
        self.calls = {}
        self._maxpool_idx = {}
        self._err = b"emulator"

    # ---- library
    def sseg_last_error(self):
        return self._err

    def sseg_launch_count(self):
        return sum(self.calls.values())

    def sseg_launch_count_reset(self):
        self.calls.clear()

    def __getattribute__(self, name):
        attr = object.__getattribute__(self, name)
        if name.startswith("sseg_") and name not in ("sseg_last_error", "sseg_launch_count", "sseg_launch_count_reset"):
            def counted(*a):
                calls = object.__getattribute__(self, "calls")
                calls[name] = calls.get(name, 0) + 1
                with torch.no_grad():
                    return attr(*a)
            return counted
        return attr

    # ---- convolution family
    def _epilogue(self, g, w, w_ld, cout, out, out_f32, bias, addend, ssum, ssq, ep=None):
        v = _conv(g, w, w_ld, cout)
        if _addr(bias):
            v = v + vec(bias, cout)
        if ep is not None:
            scale, shift, relu = ep
            if _addr(scale):
                v = v * vec(scale, cout) + vec(shift, cout)
            if relu & 3 == 2:
                v = torch.relu(v)
        else:
            relu = 0
        if addend is not None:
            v = v + act_view(addend).float()[..., :cout]
        if relu & 3 == 1:
            v = torch.relu(v)
        if relu & 4:
            v = v.clamp(max=6.0)
        dtype = torch.float32 if out_f32 else torch.bfloat16
        _store(out, v, cout, dtype)
        if _addr(ssum):
            r = v.to(dtype).float()
            vec(ssum, cout).add_(r.sum((0, 1, 2)))
            vec(ssq, cout).add_((r * r).sum((0, 1, 2)))
        return v.to(dtype).float()

    def sseg_conv_igemm(self, g, w, w_ld, cout, out, out_f32, bias, addend, ssum, ssq, stream):
        self._epilogue(g, w, w_ld, cout, out, out_f32, bias, addend, ssum, ssq)
        return 0

    def sseg_conv_igemm_affine(self, g, w, w_ld, cout, out, scale, shift, relu, addend, stream):
        self._epilogue(g, w, w_ld, cout, out, 0, None, addend, None, None, ep=(scale, shift, relu))
        return 0

    def sseg_conv_igemm_bnbwd(self, g, w, w_ld, cout, out, addend, y, fscale, fshift, s1, s2raw, stream):
        gr = self._epilogue(g, w, w_ld, cout, out, 0, None, addend, None, None)
        yv = act_view(y).float()[..., :cout]
        gp = gr * ((yv * vec(fscale, cout) + vec(fshift, cout)) > 0)
        vec(s1, cout).add_(gp.sum((0, 1, 2)))
        vec(s2raw, cout).add_((gp * yv).sum((0, 1, 2)))
        return 0

    def sseg_conv_igemm_bnbwd_res(self, g, w, w_ld, cout, out, addend, y, a, s1, s2raw, stream):
        gr = self._epilogue(g, w, w_ld, cout, out, 0, None, addend, None, None)
        yv = act_view(y).float()[..., :cout]
        gp = gr * (act_view(a).float()[..., :cout] > 0)
        vec(s1, cout).add_(gp.sum((0, 1, 2)))
        vec(s2raw, cout).add_((gp * yv).sum((0, 1, 2)))
        return 0

    def sseg_conv_wgrad(self, g, dy, cout, dw, dw_ld, stream):
        d = act_view(dy).float()[..., :cout]
        kmax = max(g.tap_koff[t] for t in range(g.ntaps)) + sum(g.srcs[i].c for i in range(g.nsrc))
        D = flat(dw, (cout - 1) * dw_ld + min(dw_ld, kmax), torch.float32).as_strided((cout, min(dw_ld, kmax)), (dw_ld, 1))
        d2 = d.reshape(-1, cout)
        for xs, koff in _taps(g):
            cin = xs.shape[3]
            D[:, koff:koff + cin] += d2.t() @ xs.reshape(-1, cin)
        return 0

    def sseg_conv_bn_train_fits(self, *a):
        return 1

    def sseg_conv_dgrad_bn_fits(self, *a):
        return 1

    @staticmethod
    def _pool(peer, off, n):
        """sum over the ranks of the n floats at `off` (4-byte units) inside every rank's arena"""
        return sum(flat(peer.bases[r] + 4 * off, n, torch.float32).clone() for r in range(peer.world))

    def sseg_conv_igemm_bnfin(self, g, w, w_ld, cout, out, bn, stream):
        bn = bn.contents if hasattr(bn, "contents") else bn
        v = _bf(_conv(g, w, w_ld, cout)).float()                      # y as stored
        _store(out, v, cout)
        ssum, ssq = vec(bn.stat_sum, cout), vec(bn.stat_sqsum, cout)
        ssum.add_(v.sum((0, 1, 2)))
        ssq.add_((v * v).sum((0, 1, 2)))
        flat(bn.counter, 1, torch.int32).add_(1)                      # the tickets leave a non-zero counter
        mean = ssum / bn.count
        sumvar = ssq - ssum * mean
        inv = torch.rsqrt((sumvar / bn.count).clamp(min=0) + bn.eps)
        gm = vec(bn.gamma, cout) if _addr(bn.gamma) else torch.ones(cout)
        bt = vec(bn.beta, cout) if _addr(bn.beta) else torch.zeros(cout)
        for ptr, val in ((bn.mean_out, mean), (bn.invstd_out, inv), (bn.scale_out, gm * inv), (bn.shift_out, bt - mean * gm * inv)):
            vec(ptr, cout).copy_(val)
        if _addr(bn.running_mean):
            rm, rv = vec(bn.running_mean, cout), vec(bn.running_var, cout)
            rm.mul_(1 - bn.momentum).add_(bn.momentum * mean)
            rv.mul_(1 - bn.momentum).add_(bn.momentum * sumvar / (bn.count - 1))
        return 0

    def sseg_conv_bn_train(self, g, w, w_ld, cout, y, a_out, bn, stream):
        v = _bf(_conv(g, w, w_ld, cout)).float()                      # y as stored
        if y is not None:
            _store(y, v, cout)
        ssum, ssq = vec(bn.stat_sum, cout), vec(bn.stat_sqsum, cout)
        ssum.add_(v.sum((0, 1, 2)))
        ssq.add_((v * v).sum((0, 1, 2)))
        flat(bn.counter, 1, torch.int32).add_(1)                      # the grid barrier leaves a non-zero counter
        if bool(bn.peer):   # synchronised branch: pooled over the ranks' arenas, clamp(var, eps), accumulator running stats
            pr = bn.peer.contents
            ssum = self._pool(pr, pr.data_off, cout)
            ssq = self._pool(pr, pr.data_off + pr.data_stride, cout)
            cnt = self._pool(pr, pr.data_off + 2 * pr.data_stride, 1).item()
            mean = ssum / cnt
            sumvar = ssq - ssum * mean
            inv = torch.rsqrt((sumvar / cnt).clamp(min=bn.eps))
            flat(bn.count_out, 1, torch.float32).fill_(cnt)
            if _addr(bn.tmp_running_mean):
                frac = 1 - bn.momentum
                vec(bn.tmp_running_mean, cout).mul_(frac).add_(mean)
                vec(bn.tmp_running_var, cout).mul_(frac).add_(sumvar / (cnt - 1))
                flat(bn.running_iter, 1, torch.float32).mul_(frac).add_(1)
        else:
            mean = ssum / bn.count
            sumvar = ssq - ssum * mean
            inv = torch.rsqrt((sumvar / bn.count).clamp(min=0) + bn.eps)
        gm = vec(bn.gamma, cout) if _addr(bn.gamma) else torch.ones(cout)
        bt = vec(bn.beta, cout) if _addr(bn.beta) else torch.zeros(cout)
        sc, sh = gm * inv, bt - mean * gm * inv
        for ptr, val in ((bn.mean_out, mean), (bn.invstd_out, inv), (bn.scale_out, sc), (bn.shift_out, sh)):
            vec(ptr, cout).copy_(val)
        if _addr(bn.running_mean) and not bool(bn.peer):
            rm, rv = vec(bn.running_mean, cout), vec(bn.running_var, cout)
            rm.mul_(1 - bn.momentum).add_(bn.momentum * mean)
            rv.mul_(1 - bn.momentum).add_(bn.momentum * sumvar / (bn.count - 1))
        t = v * sc + sh
        rr = 0
        if bool(bn.res):
            rr = act_view(bn.res.contents).float()[..., :cout]
            if _addr(bn.rscale):
                rr = rr * vec(bn.rscale, cout) + vec(bn.rshift, cout)
        if not bn.res_after_relu:
            t = t + rr
        if bn.relu:
            t = torch.relu(t)
        if bn.res_after_relu:
            t = t + rr
        if _addr(bn.chanmul):
            n = a_out.n
            t = t * flat(bn.chanmul, n * cout, torch.float32).view(n, 1, 1, cout)
        _store(a_out, t, cout)
        return 0

    def sseg_bn_running_from_tmp(self, tm, tv, it, rm, rv, C, stream):
        i = flat(it, 1, torch.float32)
        vec(rm, C).copy_(vec(tm, C) / i)
        vec(rv, C).copy_(vec(tv, C) / i)
        return 0

    def sseg_conv_dgrad_bn(self, g, w, w_ld, cout, y, dy_out, fscale, fshift, mean, invstd, count, s1, s2raw, dgamma,
                           counter, peer, count_dev, dbeta_out, stream):
        gr = _bf(_conv(g, w, w_ld, cout)).float()
        yv = act_view(y).float()[..., :cout]
        fs, fb = vec(fscale, cout), vec(fshift, cout)
        gp = gr * ((yv * fs + fb) > 0)
        a, b = vec(s1, cout), vec(s2raw, cout)
        a.add_(gp.sum((0, 1, 2)))
        b.add_((gp * yv).sum((0, 1, 2)))
        flat(counter, 1, torch.int32).add_(1)
        mu, inv = vec(mean, cout), vec(invstd, cout)
        inv_w = 1.0
        if peer is not None:   # totals over the ranks; dbeta / dgamma leave divided by world
            a = self._pool(peer, peer.data_off, cout)
            b = self._pool(peer, peer.data_off + peer.data_stride, cout)
            count = flat(count_dev, 1, torch.float32).item()
            inv_w = 1.0 / peer.world
            vec(dbeta_out, cout).copy_(a * inv_w)
        s2 = inv * (b - mu * a)
        if _addr(dgamma):
            vec(dgamma, cout).copy_(s2 * inv_w)
        tt = fs * inv * s2 / count
        _store(dy_out, fs * gp - tt * yv + (tt * mu - fs * a / count), cout)
        return 0

    # ---- weights
    def _descs(self, table, n):
        from mit_semseg.engine import _C
        return (_C.WeightDesc * n).from_address(_addr(table))

    def sseg_prep_conv_weight(self, w, O, I, T, wf, fwd_ld, wd, dgrad_ld, o_pad, stream):
        wt = flat(w, O * I * T, torch.float32).view(O, I, T)
        if _addr(wf):
            flat(wf, (O - 1) * fwd_ld + T * I, torch.bfloat16).as_strided((O, T, I), (fwd_ld, I, 1)).copy_(wt.permute(0, 2, 1))
        if _addr(wd):
            flat(wd, (I - 1) * dgrad_ld + (T - 1) * o_pad + O, torch.bfloat16).as_strided((I, T, O), (dgrad_ld, o_pad, 1)).copy_(
                wt.permute(1, 2, 0))
        return 0

    def sseg_grad_to_oihw(self, g, g_ld, O, I, T, out, scale, accumulate, stream):
        gs = flat(g, (O - 1) * g_ld + T * I, torch.float32).as_strided((O, T, I), (g_ld, I, 1)).permute(0, 2, 1) * scale
        o = flat(out, O * I * T, torch.float32).view(O, I, T)
        o.copy_(o + gs if accumulate else gs)
        return 0

    def sseg_prep_conv_weights_batched_ex(self, table, n, tiles, max_blocks, stream):
        return self.sseg_prep_conv_weights_batched(table, n, tiles, stream)

    def sseg_prep_conv_weights_batched(self, table, n, tiles, stream):
        for d in self._descs(table, n):
            w = flat(d.w, d.O * d.I * d.T, torch.float32)
            w = w.view(d.O, d.T, d.I).permute(0, 2, 1) if d.reserved == 1 else w.view(d.O, d.I, d.T)   # 1: channels-last master
            if d.wf:
                wf = flat(d.wf, (d.O - 1) * d.fwd_ld + d.T * d.I, torch.bfloat16).as_strided((d.O, d.T, d.I), (d.fwd_ld, d.I, 1))
                wf.copy_(w.permute(0, 2, 1))
            if d.wd:
                wd = flat(d.wd, (d.I - 1) * d.dgrad_ld + (d.T - 1) * d.o_pad + d.O, torch.bfloat16).as_strided(
                    (d.I, d.T, d.O), (d.dgrad_ld, d.o_pad, 1))
                wd.copy_(w.permute(1, 2, 0))
        return 0

    def sseg_grads_to_oihw_batched(self, table, n, tiles, scale, stream):
        for d in self._descs(table, n):
            gs = flat(d.g_src, (d.O - 1) * d.g_ld + d.T * d.I, torch.float32).as_strided((d.O, d.T, d.I), (d.g_ld, d.I, 1))
            flat(d.g_dst, d.O * d.I * d.T, torch.float32).view(d.O, d.I, d.T).copy_(gs.permute(0, 2, 1) * scale)
        return 0

    # ---- stem
    def sseg_stem_conv_fwd(self, img, N, H, W, w, out, ssum, ssq, stream):
        x = flat(img, N * 3 * H * W, torch.float32).view(N, 3, H, W)
        wt = flat(w, 64 * 27, torch.float32).view(64, 3, 3, 3)
        y = _bf(F.conv2d(x, wt, stride=2, padding=1).permute(0, 2, 3, 1))
        flat(out, y.numel(), torch.bfloat16).view_as(y).copy_(y)
        if _addr(ssum):
            r = y.float()
            vec(ssum, 64).add_(r.sum((0, 1, 2)))
            vec(ssq, 64).add_((r * r).sum((0, 1, 2)))
        return 0

    def sseg_stem_conv_wgrad(self, img, N, H, W, dy, dw, stream):
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        x = flat(img, N * 3 * H * W, torch.float32).view(N, 3, H, W)
        d = flat(dy, N * Ho * Wo * 64, torch.bfloat16).view(N, Ho, Wo, 64).float().permute(0, 3, 1, 2)
        with torch.enable_grad():
            wt = torch.zeros(64, 3, 3, 3, requires_grad=True)
            (gw,) = torch.autograd.grad(F.conv2d(x, wt, stride=2, padding=1), wt, d)
        flat(dw, 64 * 27, torch.float32).view(64, 3, 3, 3).add_(gw)
        return 0

    # ---- batch norm
    def sseg_bn_finalize(self, ssum, ssq, count_dev, count_host, gamma, beta, eps, momentum, mode, update, rm, rv, tm, tv, it,
                         mean_out, invstd_out, scale, shift, C, stream):
        if mode == 2:
            mean, inv = vec(rm, C).clone(), torch.rsqrt(vec(rv, C) + eps)
        else:
            cnt = flat(count_dev, 1, torch.float32).item() if _addr(count_dev) else count_host
            s, q = vec(ssum, C), vec(ssq, C)
            mean = s / cnt
            sumvar = q - s * mean
            bias, unbias = sumvar / cnt, sumvar / (cnt - 1)
            if mode == 0:
                inv = torch.rsqrt(bias.clamp(min=0) + eps)
                if update:
                    vec(rm, C).mul_(1 - momentum).add_(momentum * mean)
                    vec(rv, C).mul_(1 - momentum).add_(momentum * unbias)
            else:
                inv = torch.rsqrt(bias.clamp(min=eps))
                if update:
                    frac = 1 - momentum
                    i = flat(it, 1, torch.float32)
                    i.mul_(frac).add_(1)
                    vec(tm, C).mul_(frac).add_(mean)
                    vec(tv, C).mul_(frac).add_(unbias)
                    vec(rm, C).copy_(vec(tm, C) / i)
                    vec(rv, C).copy_(vec(tv, C) / i)
        g = vec(gamma, C) if _addr(gamma) else torch.ones(C)
        b = vec(beta, C) if _addr(beta) else torch.zeros(C)
        vec(mean_out, C).copy_(mean)
        vec(invstd_out, C).copy_(inv)
        vec(scale, C).copy_(g * inv)
        vec(shift, C).copy_(b - mean * g * inv)
        return 0

    def sseg_bn_apply(self, y, y_ld, scale, shift, res, res_ld, rscale, rshift, chanmul, out, out_ld, P, ppi, C, relu,
                      res_after_relu, stream):
        t = pix_view(y, y_ld, P, C).float() * vec(scale, C) + vec(shift, C)
        rr = 0
        if _addr(res):
            rr = pix_view(res, res_ld, P, C).float()
            if _addr(rscale):
                rr = rr * vec(rscale, C) + vec(rshift, C)
        if not res_after_relu:
            t = t + rr
        if relu:
            t = torch.relu(t)
        if res_after_relu:
            t = t + rr
        if _addr(chanmul):
            n = P // ppi
            t = (t.view(n, ppi, C) * flat(chanmul, n * C, torch.float32).view(n, 1, C)).view(P, C)
        pix_view(out, out_ld, P, C).copy_(_bf(t))
        return 0

    def sseg_bn_finalize_apply(self, ssum, ssq, count, gamma, beta, eps, momentum, rm, rv, mean_out, invstd_out, scale_out,
                               shift_out, y, y_ld, res, res_ld, rscale, rshift, chanmul, out, out_ld, P, ppi, C, relu,
                               res_after_relu, stream):
        self.sseg_bn_finalize(ssum, ssq, None, count, gamma, beta, eps, momentum, 0, 1 if _addr(rm) else 0, rm, rv, None,
                              None, None, mean_out, invstd_out, scale_out, shift_out, C, stream)
        return self.sseg_bn_apply(y, y_ld, scale_out, shift_out, res, res_ld, rscale, rshift, chanmul, out, out_ld, P, ppi, C,
                                  relu, res_after_relu, stream)

    def _gprime(self, g, g_ld, a, a_ld, y, y_ld, scale, fshift, chanmul, P, ppi, C):
        gp = pix_view(g, g_ld, P, C).float()
        if _addr(chanmul):
            n = P // ppi
            gp = (gp.view(n, ppi, C) * flat(chanmul, n * C, torch.float32).view(n, 1, C)).reshape(P, C)
        if _addr(a):
            gp = gp * (pix_view(a, a_ld, P, C).float() > 0)
        elif _addr(fshift):
            gp = gp * ((pix_view(y, y_ld, P, C).float() * vec(scale, C) + vec(fshift, C)) > 0)
        return gp

    def sseg_bn_bwd_reduce(self, g, g_ld, a, a_ld, y, y_ld, mean, invstd, scale, fshift, chanmul, s1, s2, P, ppi, C, stream):
        gp = self._gprime(g, g_ld, a, a_ld, y, y_ld, scale, fshift, chanmul, P, ppi, C)
        xhat = (pix_view(y, y_ld, P, C).float() - vec(mean, C)) * vec(invstd, C)
        vec(s1, C).add_(gp.sum(0))
        vec(s2, C).add_((gp * xhat).sum(0))
        return 0

    def sseg_bn_bwd_apply(self, g, g_ld, a, a_ld, y, y_ld, mean, invstd, scale, fshift, chanmul, s1, s2, count_dev, count_host,
                          dy, dy_ld, dres, dres_ld, P, ppi, C, eval_mode, s2_raw, dgamma_out, stream):
        gp = self._gprime(g, g_ld, a, a_ld, y, y_ld, scale, fshift, chanmul, P, ppi, C)
        if _addr(dres):
            pix_view(dres, dres_ld, P, C).copy_(_bf(gp))
        sc = vec(scale, C)
        a2 = None
        if _addr(s2):
            a2 = vec(s2, C).clone()
            if s2_raw:   # the conversion (and the dgamma store) happens in eval mode as well: frozen BN keeps trainable affines
                a2 = vec(invstd, C) * (a2 - vec(mean, C) * vec(s1, C))
                if _addr(dgamma_out):
                    vec(dgamma_out, C).copy_(a2)
        if eval_mode:
            out = sc * gp
        else:
            m = flat(count_dev, 1, torch.float32).item() if _addr(count_dev) else count_host
            mu, inv = vec(mean, C), vec(invstd, C)
            xhat = (pix_view(y, y_ld, P, C).float() - mu) * inv
            out = sc * (gp - vec(s1, C) / m - xhat * a2 / m)
        pix_view(dy, dy_ld, P, C).copy_(_bf(out))
        return 0

    # ---- pooling / resize
    def sseg_maxpool_fwd(self, x, N, H, W, C, out, idx, stream):
        xv = flat(x, N * H * W * C, torch.bfloat16).view(N, H, W, C).float().permute(0, 3, 1, 2)
        o, ind = F.max_pool2d(xv, 3, 2, 1, return_indices=True)
        flat(out, o.numel(), torch.bfloat16).view(N, o.shape[2], o.shape[3], C).copy_(_bf(o.permute(0, 2, 3, 1)))
        if _addr(idx):
            self._maxpool_idx[_addr(idx)] = ind
        return 0

    def sseg_maxpool_bwd(self, dout, idx, dx, N, H, W, C, stream):
        ind = self._maxpool_idx[_addr(idx)]
        Ho, Wo = ind.shape[2], ind.shape[3]
        d = flat(dout, N * Ho * Wo * C, torch.bfloat16).view(N, Ho, Wo, C).float().permute(0, 3, 1, 2)
        gx = torch.zeros(N, C, H * W).scatter_add_(2, ind.reshape(N, C, -1), d.reshape(N, C, -1)).view(N, C, H, W)
        flat(dx, N * H * W * C, torch.bfloat16).view(N, H, W, C).copy_(_bf(gx.permute(0, 2, 3, 1)))
        return 0

    def sseg_avgpool_fwd(self, x, x_ld, N, H, W, C, S, out, stream):
        xv = pix_view(x, x_ld, N * H * W, C).float().view(N, H, W, C).permute(0, 3, 1, 2)
        o = F.adaptive_avg_pool2d(xv, S).permute(0, 2, 3, 1)
        flat(out, N * S * S * C, torch.bfloat16).view(N, S, S, C).copy_(_bf(o))
        return 0

    def sseg_avgpool_bwd(self, base, base_ld, dpool, scales, nscales, dx, dx_ld, N, H, W, C, stream):
        acc = pix_view(base, base_ld, N * H * W, C).float().view(N, H, W, C).clone() if _addr(base) else torch.zeros(N, H, W, C)
        for k in range(nscales):
            S = scales[k]
            d = flat(dpool[k], N * S * S * C, torch.bfloat16).view(N, S, S, C).float().permute(0, 3, 1, 2)
            with torch.enable_grad():
                z = torch.zeros(N, C, H, W, requires_grad=True)
                (gz,) = torch.autograd.grad(F.adaptive_avg_pool2d(z, S), z, d)
            acc += gz.permute(0, 2, 3, 1)
        pix_view(dx, dx_ld, N * H * W, C).copy_(_bf(acc.reshape(-1, C)))
        return 0

    def sseg_bilinear_fwd(self, x, x_ld, N, Hi, Wi, C, out, out_ld, Ho, Wo, stream):
        xv = pix_view(x, x_ld, N * Hi * Wi, C).float().view(N, Hi, Wi, C).permute(0, 3, 1, 2)
        o = F.interpolate(xv, size=(Ho, Wo), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
        pix_view(out, out_ld, N * Ho * Wo, C).copy_(_bf(o.reshape(-1, C)))
        return 0

    def sseg_bilinear_bwd(self, dout, dout_ld, N, Ho, Wo, C, dx, dx_ld, Hi, Wi, accumulate, scratch, stream):
        d = pix_view(dout, dout_ld, N * Ho * Wo, C).float().view(N, Ho, Wo, C).permute(0, 3, 1, 2)
        with torch.enable_grad():
            z = torch.zeros(N, C, Hi, Wi, requires_grad=True)
            (gz,) = torch.autograd.grad(F.interpolate(z, size=(Ho, Wo), mode="bilinear", align_corners=False), z, d)
        gz = gz.permute(0, 2, 3, 1).reshape(-1, C)
        o = pix_view(dx, dx_ld, N * Hi * Wi, C)
        o.copy_(_bf(gz + o.float()) if accumulate else _bf(gz))
        return 0

    def sseg_sum_terms(self, terms, nterms, N, Ho, Wo, C, out, out_ld, relu, stream):
        acc = torch.zeros(N, Ho, Wo, C)
        for k in range(nterms):
            t = terms[k]
            x = pix_view(t.x, t.ld, N * t.h * t.w, C).float().view(N, t.h, t.w, C)
            if (t.h, t.w) != (Ho, Wo):
                x = F.interpolate(x.permute(0, 3, 1, 2), size=(Ho, Wo), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
            if t.scale:
                x = x * vec(t.scale, C) + vec(t.shift, C)
            acc += x
        if relu:
            acc = torch.relu(acc)
        pix_view(out, out_ld, N * Ho * Wo, C).copy_(_bf(acc.reshape(-1, C)))
        return 0

    def sseg_relu_mask_bwd(self, g, g_ld, out, out_ld, ds, ds_ld, acc_out, acc_ld, accumulate, P, C, stream):
        d = pix_view(g, g_ld, P, C).float() * (pix_view(out, out_ld, P, C).float() > 0)
        pix_view(ds, ds_ld, P, C).copy_(_bf(d))
        if _addr(acc_out):
            o = pix_view(acc_out, acc_ld, P, C)
            o.copy_(_bf(o.float() + _bf(d).float()) if accumulate else _bf(d))
        return 0

    # ---- loss / head / layout
    def sseg_softmax_nll_fwd(self, logits, ld, C, label, P, lse, accum, stream):
        lg = pix_view(logits, ld, P, C, torch.float32)
        lab = flat(label, P, torch.int64)
        l = torch.logsumexp(lg, 1)
        flat(lse, P, torch.float32).copy_(l)
        valid = (lab >= 0) & (lab < C)
        idx = lab.clamp(0, C - 1)
        acc = vec(accum, 3)
        acc[0] += ((l - lg.gather(1, idx[:, None])[:, 0]) * valid).sum()
        acc[1] += valid.sum()
        acc[2] += (valid & (lg.argmax(1) == lab)).sum()
        return 0

    def sseg_nll_finalize(self, main, ds, ds_scale, out, stream):
        m = vec(main, 3)
        o = vec(out, 2)
        o[0] = m[0] / m[1]
        if _addr(ds):
            d = vec(ds, 3)
            o[0] += ds_scale * d[0] / d[1]
        o[1] = m[2] / (m[1] + 1e-10)
        return 0

    def sseg_softmax_nll_bwd(self, logits, ld, C, label, lse, accum, weight, P, dlogits, ld_out, c_store, stream):
        lg = pix_view(logits, ld, P, C, torch.float32)
        lab = flat(label, P, torch.int64)
        valid = (lab >= 0) & (lab < C)
        sm = torch.exp(lg - flat(lse, P, torch.float32)[:, None])
        sm[torch.arange(P)[valid], lab[valid]] -= 1
        d = sm * (weight / vec(accum, 3)[1]) * valid[:, None]
        o = pix_view(dlogits, ld_out, P, c_store)
        o.zero_()
        o[:, :C] = _bf(d)
        return 0

    def sseg_colsum(self, x, ld, P, C, out, stream):
        vec(out, C).add_(pix_view(x, ld, P, C).float().sum(0))
        return 0

    def sseg_upsample_softmax(self, logits, ld, N, Hi, Wi, C, probs, Ho, Wo, weight, accumulate, log_output, stream):
        lg = pix_view(logits, ld, N * Hi * Wi, C, torch.float32).view(N, Hi, Wi, C).permute(0, 3, 1, 2)
        up = F.interpolate(lg, size=(Ho, Wo), mode="bilinear", align_corners=False) if (Hi, Wi) != (Ho, Wo) else lg
        p = F.log_softmax(up, 1) if log_output else F.softmax(up, 1)
        o = flat(probs, N * C * Ho * Wo, torch.float32).view(N, C, Ho, Wo)
        o.copy_(o + weight * p if accumulate else weight * p)
        return 0

    def sseg_nhwc_bf16_to_nchw_f32(self, x, ld, N, H, W, C, out, stream):
        flat(out, N * C * H * W, torch.float32).view(N, C, H, W).copy_(
            pix_view(x, ld, N * H * W, C).float().view(N, H, W, C).permute(0, 3, 1, 2))
        return 0

    def sseg_nchw_f32_to_nhwc_bf16(self, x, N, H, W, C, out, ld, stream):
        pix_view(out, ld, N * H * W, C).copy_(_bf(flat(x, N * C * H * W, torch.float32).view(N, C, H, W).permute(0, 2, 3, 1)
                                                  .reshape(-1, C)))
        return 0

    # ---- input pipeline
    def sseg_image_transform(self, img, N, H, W, valid, mean_std, out, stream):
        if W % 4:
            self._err = b"sseg_image_transform: W must be a multiple of 4"
            return -1
        x = flat(img, N * H * W * 3, torch.uint8).view(N, H, W, 3)
        v = flat(valid, 2 * N, torch.int32).view(N, 2)
        ms = torch.tensor([mean_std[i] for i in range(6)], dtype=torch.float32)
        o = flat(out, N * 3 * H * W, torch.float32).view(N, 3, H, W)
        o.zero_()
        for n in range(N):
            h, w = int(v[n, 0]), int(v[n, 1])
            o[n, :, :h, :w] = ((x[n, :h, :w].float() / 255.).permute(2, 0, 1) - ms[:3].view(3, 1, 1)) / ms[3:].view(3, 1, 1)
        return 0

    def sseg_label_transform(self, seg, N, Hs, Ws, valid, rate, out, stream):
        x = flat(seg, N * Hs * Ws, torch.uint8).view(N, Hs, Ws)
        v = flat(valid, 2 * N, torch.int32).view(N, 2)
        o = flat(out, N * Hs * Ws, torch.int64).view(N, Hs, Ws)
        o.zero_()
        for n in range(N):
            h, w = -(-int(v[n, 0]) // rate), -(-int(v[n, 1]) // rate)
            o[n, :h, :w] = x[n, :h, :w].long() - 1
        return 0

    # ---- fp32-accurate inference (pairs of bf16 tensors)
    @staticmethod
    def _pair_val(hi, lo, ld, P, C):
        return pix_view(hi, ld, P, C).float() + pix_view(lo, ld, P, C).float()

    @staticmethod
    def _pair_store(hi, lo, ld, P, C, val):
        h = _bf(val)
        pix_view(hi, ld, P, C).copy_(h)
        pix_view(lo, ld, P, C).copy_(_bf(val - h.float()))

    def sseg_split_affine(self, z, scale, shift, res_hi, res_lo, res_ld, out_hi, out_lo, out_ld, relu, res_after_relu, stream):
        v = act_view(z, torch.float32).clone()
        n, h, w, c = v.shape
        P = n * h * w
        if _addr(scale):
            v = v * vec(scale, c) + vec(shift, c)
        v = v.reshape(P, c)
        rr = self._pair_val(res_hi, res_lo, res_ld, P, c) if _addr(res_hi) else 0
        if not res_after_relu:
            v = v + rr
        if relu:
            v = torch.relu(v)
        if res_after_relu:
            v = v + rr
        self._pair_store(out_hi, out_lo, out_ld, P, c, v)
        return 0

    def sseg_stem_conv_fwd_f32(self, img, N, H, W, w, out, stream):
        x = flat(img, N * 3 * H * W, torch.float32).view(N, 3, H, W)
        wt = flat(w, 64 * 27, torch.float32).view(64, 3, 3, 3)
        y = F.conv2d(x, wt, stride=2, padding=1).permute(0, 2, 3, 1)
        flat(out, y.numel(), torch.float32).view(y.shape).copy_(y)
        return 0

    def sseg_maxpool_pair_fwd(self, xh, xl, N, H, W, C, oh, ol, stream):
        v = self._pair_val(xh, xl, C, N * H * W, C).view(N, H, W, C).permute(0, 3, 1, 2)
        o = F.max_pool2d(v, 3, 2, 1).permute(0, 2, 3, 1)
        self._pair_store(oh, ol, C, o.shape[0] * o.shape[1] * o.shape[2], C, o.reshape(-1, C))
        return 0

    def sseg_avgpool_pair_fwd(self, xh, xl, x_ld, N, H, W, C, S, oh, ol, stream):
        v = self._pair_val(xh, xl, x_ld, N * H * W, C).view(N, H, W, C).permute(0, 3, 1, 2)
        o = F.adaptive_avg_pool2d(v, S).permute(0, 2, 3, 1)
        self._pair_store(oh, ol, C, N * S * S, C, o.reshape(-1, C))
        return 0

    def sseg_bilinear_pair_fwd(self, xh, xl, x_ld, N, Hi, Wi, C, oh, ol, out_ld, Ho, Wo, stream):
        v = self._pair_val(xh, xl, x_ld, N * Hi * Wi, C).view(N, Hi, Wi, C).permute(0, 3, 1, 2)
        o = F.interpolate(v, size=(Ho, Wo), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
        self._pair_store(oh, ol, out_ld, N * Ho * Wo, C, o.reshape(-1, C))
        return 0

    def sseg_prep_conv_weight_split(self, w, O, I, T, out, ld, stream):
        wt = flat(w, O * I * T, torch.float32).view(O, I, T)
        hi = _bf(wt)
        lo = _bf(wt - hi.float())
        o = flat(out, (O - 1) * ld + 3 * T * I, torch.bfloat16).as_strided((O, T, 3, I), (ld, 3 * I, I, 1))
        o[:, :, 0, :] = hi.permute(0, 2, 1)
        o[:, :, 1, :] = hi.permute(0, 2, 1)
        o[:, :, 2, :] = lo.permute(0, 2, 1)
        return 0

    def sseg_dwconv_affine(self, x, N, H, W, C, w, stride, dilation, scale, shift, relu6, out, stream):
        xv = flat(x, N * H * W * C, torch.bfloat16).view(N, H, W, C).float().permute(0, 3, 1, 2)
        wt = flat(w, C * 9, torch.float32).view(C, 1, 3, 3)
        y = F.conv2d(xv, wt, None, stride, dilation, dilation, C).permute(0, 2, 3, 1)
        if _addr(scale):
            y = y * vec(scale, C) + vec(shift, C)
        if relu6:
            y = y.clamp(0.0, 6.0)
        flat(out, y.numel(), torch.bfloat16).view(y.shape).copy_(_bf(y))
        return 0

    def sseg_stem_conv_affine(self, img, N, H, W, w, cout, scale, shift, relu6, out, stream):
        x = flat(img, N * 3 * H * W, torch.float32).view(N, 3, H, W)
        wt = flat(w, cout * 27, torch.float32).view(cout, 3, 3, 3)
        y = F.conv2d(x, wt, stride=2, padding=1).permute(0, 2, 3, 1)
        if _addr(scale):
            y = y * vec(scale, cout) + vec(shift, cout)
        if relu6:
            y = y.clamp(0.0, 6.0)
        flat(out, y.numel(), torch.bfloat16).view(y.shape).copy_(_bf(y))
        return 0

    # ---- SyncBN over peer arenas (csrc/peer.cu); the handshake itself (flags over NVLink) has no CPU counterpart
    class _Peers:
        def __init__(self, bases, world):
            self.bases, self.world = [_addr(bases[r]) for r in range(world)], world

    def sseg_peer_step(self, step, stream):
        flat(step, 1, torch.int32).add_(1)
        return 0

    def sseg_bn_finalize_peer(self, bases, world, rank, stats_off, flag_off, step, gamma, beta, eps, momentum, update, rm, rv,
                              tm, tv, it, mean_out, invstd_out, scale, shift, count_out, C, stream):
        pr = self._Peers(bases, world)
        s, q = self._pool(pr, stats_off, C), self._pool(pr, stats_off + C, C)
        cnt = self._pool(pr, stats_off + 2 * C, 1).item()
        flat(count_out, 1, torch.float32).fill_(cnt)
        mean = s / cnt
        sumvar = q - s * mean
        inv = torch.rsqrt((sumvar / cnt).clamp(min=eps))
        if update:
            frac = 1 - momentum
            i = flat(it, 1, torch.float32)
            i.mul_(frac).add_(1)
            vec(tm, C).mul_(frac).add_(mean)
            vec(tv, C).mul_(frac).add_(sumvar / (cnt - 1))
            if update == 1:
                vec(rm, C).copy_(vec(tm, C) / i)
                vec(rv, C).copy_(vec(tv, C) / i)
        g = vec(gamma, C) if _addr(gamma) else torch.ones(C)
        b = vec(beta, C) if _addr(beta) else torch.zeros(C)
        for ptr, val in ((mean_out, mean), (invstd_out, inv), (scale, g * inv), (shift, b - mean * g * inv)):
            vec(ptr, C).copy_(val)
        return 0

    # push-protocol variants: same pooled results (the emulator reads the partials out of every rank's arena either way)
    def sseg_bn_finalize_peer_ll(self, bases, world, rank, stats_off, inbox_off, *rest):
        return self.sseg_bn_finalize_peer(bases, world, rank, stats_off, 0, *rest)

    def sseg_bn_bwd_peer_sum_ll(self, bases, world, rank, part_off, inbox_off, *rest):
        return self.sseg_bn_bwd_peer_sum(bases, world, rank, part_off, 0, *rest)

    def sseg_bn_bwd_peer_sum(self, bases, world, rank, part_off, flag_off, step, s1_tot, s2_tot, dbeta, dgamma, mean, invstd,
                             s2_raw, C, stream):
        pr = self._Peers(bases, world)
        a, b = self._pool(pr, part_off, C), self._pool(pr, part_off + C, C)
        if s2_raw:
            b = vec(invstd, C) * (b - vec(mean, C) * a)
        vec(s1_tot, C).copy_(a)
        vec(s2_tot, C).copy_(b)
        vec(dbeta, C).copy_(a / world)
        vec(dgamma, C).copy_(b / world)
        return 0

    def sseg_bn_bwd_apply_peer(self, bases, world, rank, part_off, flag_off, step, g, g_ld, a, a_ld, y, y_ld, mean, invstd,
                               scale, fshift, chanmul, count_dev, dy, dy_ld, dres, dres_ld, P, ppi, C, s2_raw, dbeta_out,
                               dgamma_out, stream):
        pr = self._Peers(bases, world)
        s1 = self._pool(pr, part_off, C).clone()
        s2 = self._pool(pr, part_off + C, C).clone()
        if s2_raw:
            s2 = vec(invstd, C) * (s2 - vec(mean, C) * s1)
        vec(dbeta_out, C).copy_(s1 / world)
        vec(dgamma_out, C).copy_(s2 / world)
        keep = (s1.contiguous(), s2.contiguous())
        return self.sseg_bn_bwd_apply(g, g_ld, a, a_ld, y, y_ld, mean, invstd, scale, fshift, chanmul,
                                      ctypes.c_void_p(keep[0].data_ptr()), ctypes.c_void_p(keep[1].data_ptr()), count_dev, 1.0, dy,
                                      dy_ld, dres, dres_ld, P, ppi, C, 0, 0, None, stream)

    # ---- optimiser
    def sseg_scale_by_scalar(self, x, n, scalar, stream):
        v = flat(scalar, 1, torch.float32)[0].item()
        if v != 1.0:
            flat(x, n, torch.float32).mul_(v)
        return 0

    def sseg_sgd_step(self, chunks, nchunks, lr, momentum, first_step, stream):
        from mit_semseg.engine import _C
        for c in (_C.SgdChunk * nchunks).from_address(_addr(chunks)):
            p, g, b = flat(c.param, c.n, torch.float32), flat(c.grad, c.n, torch.float32), flat(c.momentum_buf, c.n, torch.float32)
            d = g + c.weight_decay * p
            b.copy_(d if first_step else momentum * b + d)
            p.sub_(lr * b)
        return 0
