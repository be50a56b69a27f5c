"""-m gpu: the HBM-bound kernels (BN, pooling, resize, stem conv, loss, layout) against torch fp32 on the same device.
Inputs are bf16-rounded first; outputs are compared within bf16 rounding (2^-8 relative to the tensor scale) unless
the kernel's output is fp32, where 1e-4 relative applies."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _gen(seed=0):
    return torch.Generator(device=DEV).manual_seed(seed)


def _close(got, ref, rel, what=""):
    scale = max(ref.abs().max().item(), 1e-6)
    err = (got.float() - ref.float()).abs().max().item()
    assert err <= rel * scale, "%s: max err %g > %g (scale %g)" % (what, err, rel * scale, scale)


def nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def test_prep_weight_and_grad_layout():
    from mit_semseg.engine import ops
    g = _gen()
    w = torch.randn(150, 128, 3, 3, device=DEV, generator=g)
    wf = torch.empty(150, 9 * 128, device=DEV, dtype=torch.bfloat16)
    wd = torch.zeros(128, 9 * 192, device=DEV, dtype=torch.bfloat16)
    ops.prep_conv_weight(w, wf, wd, o_pad=192)
    ref_f = w.permute(0, 2, 3, 1).reshape(150, -1).bfloat16()
    assert torch.equal(wf, ref_f)
    ref_d = torch.zeros(128, 9, 192, device=DEV)
    ref_d[:, :, :150] = w.permute(1, 2, 3, 0).reshape(128, 9, 150)
    assert torch.equal(wd, ref_d.reshape(128, -1).bfloat16())
    gsrc = torch.randn(150, 9 * 128, device=DEV, generator=g)
    out = torch.ones(150, 128, 3, 3, device=DEV)
    ops.grad_to_oihw(gsrc, 150, 128, 9, out, scale=0.5, accumulate=True)
    ref = 1.0 + 0.5 * gsrc.reshape(150, 3, 3, 128).permute(0, 3, 1, 2)
    _close(out, ref, 1e-6, "grad_to_oihw")


@pytest.mark.parametrize("H,W", [(70, 96), (33, 300), (64, 256)])   # one partial tile of the 64-pixel row tiles / several
def test_stem_conv_fwd_and_wgrad(H, W):
    from mit_semseg.engine import ops
    g = _gen(1)
    img = torch.randn(2, 3, H, W, device=DEV, generator=g)
    w = torch.randn(64, 3, 3, 3, device=DEV, generator=g) * 0.2
    ref = F.conv2d(img, w, stride=2, padding=1)
    ho, wo = ref.shape[2:]
    out = torch.empty(2, ho, wo, 64, device=DEV, dtype=torch.bfloat16)
    ssum, ssq = torch.zeros(64, device=DEV), torch.zeros(64, device=DEV)
    ops.stem_conv_fwd(img, w, out, ssum, ssq)
    _close(nchw(out), ref, 2 ** -8, "stem fwd")
    # statistics are those of the bf16-rounded stored values: compare with the rounded reference
    rr = ref.bfloat16().float()
    _close(ssum, rr.sum(dim=(0, 2, 3)), 2e-3, "stem sum")
    _close(ssq, (rr * rr).sum(dim=(0, 2, 3)), 2e-3, "stem sqsum")
    dy = (torch.randn(2, ho, wo, 64, device=DEV, generator=g) * 0.1).bfloat16()
    wp = w.clone().requires_grad_(True)
    (gref,) = torch.autograd.grad(F.conv2d(img, wp, stride=2, padding=1), wp, nchw(dy))
    dw = torch.zeros(64, 3, 3, 3, device=DEV)
    ops.stem_conv_wgrad(img, dy, dw)
    _close(dw, gref, 1e-4, "stem wgrad")


@pytest.mark.parametrize("C,hw", [(64, 40), (256, 24), (2048, 8)])
@pytest.mark.parametrize("mode", [0, 1])
def test_bn_forward_backward(C, hw, mode):
    """conv-output y -> BN(train) -> (+res) -> ReLU -> dropout-mask; backward against autograd of the same graph."""
    from mit_semseg.engine import ops
    g = _gen(C + mode)
    n = 2
    y = (torch.randn(n, hw, hw, C, device=DEV, generator=g) * 2 + 0.5).bfloat16()
    res = torch.randn(n, hw, hw, C, device=DEV, generator=g).bfloat16()
    gamma = torch.rand(C, device=DEV, generator=g) + 0.5
    beta = torch.randn(C, device=DEV, generator=g) * 0.1
    mask = (torch.rand(n, C, device=DEV, generator=g) > 0.1).float() / 0.9
    eps, mom = 1e-5, 0.001
    yf = y.float()
    ssum, ssq = yf.sum(dim=(0, 1, 2)), (yf * yf).sum(dim=(0, 1, 2))
    cnt = n * hw * hw
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    tm, tv, it = rm.clone(), rv.clone(), torch.ones(1, device=DEV)
    mean, invstd, scale, shift = [torch.empty(C, device=DEV) for _ in range(4)]
    ops.bn_finalize(ssum, ssq, cnt, gamma, beta, eps, mom, mode, mean, invstd, scale, shift, running=(rm, rv, tm, tv, it),
                    update_running=True)
    # reference statistics
    m_ref = yf.mean(dim=(0, 1, 2))
    v_ref = yf.var(dim=(0, 1, 2), unbiased=False)
    inv_ref = (v_ref + eps).rsqrt() if mode == 0 else v_ref.clamp(min=eps).rsqrt()
    _close(mean, m_ref, 1e-4, "mean")
    _close(invstd, inv_ref, 1e-3, "invstd")
    unb = v_ref * cnt / (cnt - 1)
    if mode == 0:
        _close(rm, mom * m_ref, 1e-3, "running_mean")
        _close(rv, (1 - mom) + mom * unb, 1e-4, "running_var")
    else:
        it_ref = 1 * (1 - mom) + 1
        _close(rm, (0 * (1 - mom) + m_ref) / it_ref, 1e-3, "sync running_mean")
        _close(rv, (1 * (1 - mom) + unb) / it_ref, 1e-3, "sync running_var")
        _close(it, torch.full((1,), it_ref, device=DEV), 1e-6, "running_iter")
    out = torch.empty_like(y)
    ops.bn_apply(y, scale, shift, out, relu=True, res=res, chanmul=mask)
    # autograd reference of the same graph, using the kernel's own (mean, invstd) definition
    yv = yf.clone().requires_grad_(True)
    rr = res.float().clone().requires_grad_(True)
    mu = yv.mean(dim=(0, 1, 2))
    var = yv.var(dim=(0, 1, 2), unbiased=False)
    inv = (var + eps).rsqrt() if mode == 0 else var.clamp(min=eps).rsqrt()
    z = torch.relu((yv - mu) * inv * gamma + beta + rr) * mask.view(n, 1, 1, C)
    _close(out, z.detach(), 2 ** -7, "bn_apply")
    gout = (torch.randn(n, hw, hw, C, device=DEV, generator=g)).bfloat16()
    dy_ref, dres_ref = torch.autograd.grad(z, (yv, rr), gout.float())
    s1, s2 = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    ops.bn_bwd_reduce(gout, out, y, mean, invstd, s1, s2, chanmul=mask)
    dy, dres = torch.empty_like(y), torch.empty_like(y)
    ops.bn_bwd_apply(gout, out, y, mean, invstd, scale, s1, s2, cnt, dy, dres=dres, chanmul=mask)
    # note: the kernel masks with the *stored bf16* output (a > 0); sub-ulp outputs can flip, hence the tolerance
    _close(dres, dres_ref, 2 ** -6, "dres")
    _close(dy, dy_ref, 2 ** -5, "dy")
    # no-shortcut layers: the ReLU mask recomputed from y must give the same result as the saved output
    out2 = torch.empty_like(y)
    ops.bn_apply(y, scale, shift, out2, relu=True, chanmul=mask)
    s1a, s2a, s1b, s2b = [torch.zeros(C, device=DEV) for _ in range(4)]
    ops.bn_bwd_reduce(gout, out2, y, mean, invstd, s1a, s2a, chanmul=mask)
    ops.bn_bwd_reduce(gout, None, y, mean, invstd, s1b, s2b, chanmul=mask, scale=scale, fshift=shift)
    _close(s1b, s1a, 1e-4, "mask-from-y s1")
    _close(s2b, s2a, 1e-4, "mask-from-y s2")
    dya, dyb = torch.empty_like(y), torch.empty_like(y)
    ops.bn_bwd_apply(gout, out2, y, mean, invstd, scale, s1a, s2a, cnt, dya, chanmul=mask)
    ops.bn_bwd_apply(gout, None, y, mean, invstd, scale, s1a, s2a, cnt, dyb, chanmul=mask, fshift=shift)
    assert torch.equal(dya, dyb)
    # dgamma / dbeta are s2 / s1
    gam = gamma.clone().requires_grad_(True)
    bet = beta.clone().requires_grad_(True)
    z2 = torch.relu((yf - m_ref) * inv_ref * gam + bet + res.float()) * mask.view(n, 1, 1, C)
    dg, db = torch.autograd.grad(z2, (gam, bet), gout.float())
    _close(s2, dg, 2e-2, "dgamma")
    _close(s1, db, 2e-2, "dbeta")


def test_bn_eval_mode():
    from mit_semseg.engine import ops
    g = _gen(5)
    C = 128
    y = torch.randn(2, 16, 16, C, device=DEV, generator=g).bfloat16()
    gamma, beta = torch.rand(C, device=DEV, generator=g) + 0.5, torch.randn(C, device=DEV, generator=g)
    rm, rv = torch.randn(C, device=DEV, generator=g) * 0.1, torch.rand(C, device=DEV, generator=g) + 0.5
    mean, invstd, scale, shift = [torch.empty(C, device=DEV) for _ in range(4)]
    ops.bn_finalize(None, None, 1, gamma, beta, 1e-5, 0.001, 2, mean, invstd, scale, shift, running=(rm, rv, None, None, None))
    out = torch.empty_like(y)
    ops.bn_apply(y, scale, shift, out, relu=False)
    ref = F.batch_norm(nchw(y), rm, rv, gamma, beta, False, 0.0, 1e-5)
    _close(nchw(out), ref, 2 ** -8, "bn eval")


def test_maxpool():
    from mit_semseg.engine import ops
    g = _gen(2)
    x = torch.randn(2, 37, 50, 128, device=DEV, generator=g).bfloat16()
    xr = nchw(x).requires_grad_(True)
    ref = F.max_pool2d(xr, 3, 2, 1)
    ho, wo = ref.shape[2:]
    out = torch.empty(2, ho, wo, 128, device=DEV, dtype=torch.bfloat16)
    idx = torch.empty(2, ho, wo, 128, device=DEV, dtype=torch.uint8)
    ops.maxpool_fwd(x, out, idx)
    assert torch.equal(nchw(out), ref.detach())
    dout = torch.randn(2, ho, wo, 128, device=DEV, generator=g).bfloat16()
    (gref,) = torch.autograd.grad(ref, xr, nchw(dout))
    dx = torch.empty_like(x)
    ops.maxpool_bwd(dout, idx, dx)
    _close(nchw(dx), gref, 2 ** -7, "maxpool bwd")


@pytest.mark.parametrize("H,W", [(64, 64), (38, 50)])
def test_adaptive_avgpool(H, W):
    from mit_semseg.engine import ops
    g = _gen(3)
    C = 256
    x = torch.randn(2, H, W, C, device=DEV, generator=g).bfloat16()
    xr = nchw(x).requires_grad_(True)
    scales = (1, 2, 3, 6)
    dps, refs = [], []
    for s in scales:
        ref = F.adaptive_avg_pool2d(xr, s)
        out = torch.empty(2, s, s, C, device=DEV, dtype=torch.bfloat16)
        ops.avgpool_fwd(x, s, out)
        _close(nchw(out), ref.detach(), 2 ** -8, "avgpool %d" % s)
        refs.append(ref)
        dps.append(torch.randn(2, s, s, C, device=DEV, generator=g).bfloat16())
    base = torch.randn(2, H, W, C, device=DEV, generator=g).bfloat16()
    gref = torch.autograd.grad(refs, xr, [nchw(d) for d in dps])[0] + nchw(base)
    dx = torch.empty_like(x)
    ops.avgpool_bwd(base, dps, scales, dx)
    _close(nchw(dx), gref, 2 ** -7, "avgpool bwd")


@pytest.mark.parametrize("hi,ho", [(1, 64), (2, 64), (3, 64), (6, 64), (6, 38), (16, 32), (8, 8)])
def test_bilinear(hi, ho):
    from mit_semseg.engine import ops
    g = _gen(hi * 100 + ho)
    C = 64
    wi, wo = hi + (1 if hi > 2 else 0), ho + 4
    x = torch.randn(2, hi, wi, C, device=DEV, generator=g).bfloat16()
    xr = nchw(x).requires_grad_(True)
    ref = F.interpolate(xr, (ho, wo), mode="bilinear", align_corners=False)
    out = torch.empty(2, ho, wo, C, device=DEV, dtype=torch.bfloat16)
    ops.bilinear_fwd(x, out)
    _close(nchw(out), ref.detach(), 2 ** -7, "bilinear fwd")
    dout = torch.randn(2, ho, wo, C, device=DEV, generator=g).bfloat16()
    (gref,) = torch.autograd.grad(ref, xr, nchw(dout))
    dx = torch.empty_like(x)
    ops.bilinear_bwd(dout, dx)
    _close(nchw(dx), gref, 2 ** -7, "bilinear bwd")
    ops.bilinear_bwd(dout, dx, accumulate=True)
    _close(nchw(dx), 2 * gref, 2 ** -6, "bilinear bwd accumulate")


def test_softmax_nll_and_acc():
    from mit_semseg.engine import ops
    g = _gen(4)
    n, h, w, C, ld = 2, 64, 64, 150, 160
    logits = torch.zeros(n, h, w, ld, device=DEV)
    logits[..., :C] = torch.randn(n, h, w, C, device=DEV, generator=g) * 3
    lg = logits[..., :152]
    label = torch.randint(-1, C, (n, h, w), device=DEV, generator=g)
    lse = torch.empty(n * h * w, device=DEV)
    acc = torch.zeros(3, device=DEV)
    ops.softmax_nll_fwd(lg, C, label, lse, acc)
    lr = logits[..., :C].permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    logp = F.log_softmax(lr, dim=1)
    loss_ref = F.nll_loss(logp, label, ignore_index=-1)
    valid = label >= 0
    acc_ref = ((logp.argmax(1) == label) & valid).sum().float() / (valid.sum().float() + 1e-10)
    out = torch.empty(2, device=DEV)
    ops.nll_finalize(acc, acc, 0.4, out)
    _close(out[0], 1.4 * loss_ref.detach(), 1e-5, "loss")
    _close(out[1], acc_ref, 1e-6, "acc")
    dl = torch.full((n, h, w, 192), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.softmax_nll_bwd(lg, C, label, lse, acc, 0.4, dl)
    (gref,) = torch.autograd.grad(0.4 * loss_ref, lr)
    _close(dl[..., :C].float().permute(0, 3, 1, 2), gref, 2 ** -8, "dlogits")
    assert dl[..., C:].float().abs().max().item() == 0.0
    bias_g = torch.zeros(C, device=DEV)
    ops.colsum(dl, C, bias_g)
    _close(bias_g, dl[..., :C].float().sum(dim=(0, 1, 2)), 1e-4, "colsum")


def test_upsample_softmax_and_layout():
    from mit_semseg.engine import ops
    g = _gen(6)
    n, h, w, C = 1, 24, 32, 150
    logits = torch.zeros(n, h, w, 160, device=DEV)
    logits[..., :C] = torch.randn(n, h, w, C, device=DEV, generator=g) * 2
    probs = torch.zeros(n, C, 100, 131, device=DEV)
    ops.upsample_softmax(logits[..., :152], C, probs, weight=0.2, accumulate=False)
    ops.upsample_softmax(logits[..., :152], C, probs, weight=0.2, accumulate=True)
    ref = F.softmax(F.interpolate(logits[..., :C].permute(0, 3, 1, 2), size=(100, 131), mode="bilinear",
                                  align_corners=False), dim=1) * 0.4
    _close(probs, ref, 1e-4, "upsample_softmax")
    x = torch.randn(2, 19, 23, 72, device=DEV, generator=g).bfloat16()
    o = torch.empty(2, 72, 19, 23, device=DEV)
    ops.nhwc_bf16_to_nchw_f32(x, o)
    assert torch.equal(o, nchw(x))
    back = torch.empty_like(x)
    ops.nchw_f32_to_nhwc_bf16(o, back)
    assert torch.equal(back, x)


def test_batched_weight_prep_and_grad_layout():
    from mit_semseg.engine import ops
    g = _gen(9)
    specs = [(150, 512, 1), (64, 64, 9), (256, 128, 9), (512, 4096, 9)]
    entries, refs = [], []
    for O_, I_, T_ in specs:
        k = int(T_ ** 0.5)
        w = torch.randn(O_, I_, k, k, device=DEV, generator=g)
        opad = (O_ + 63) // 64 * 64
        e = dict(w=w, wf=torch.empty(O_, T_ * I_, device=DEV, dtype=torch.bfloat16),
                 wd=torch.zeros(I_, T_ * opad, device=DEV, dtype=torch.bfloat16),
                 g_src=torch.randn(O_, T_ * I_, device=DEV, generator=g), g_dst=torch.empty_like(w),
                 O=O_, I=I_, T=T_, o_pad=opad)
        entries.append(e)
    tab = ops.WeightTable(entries, DEV)
    tab.prep()
    tab.grads(0.5)
    torch.cuda.synchronize()
    for e in entries:
        w, O_, I_, T_, opad = e["w"], e["O"], e["I"], e["T"], e["o_pad"]
        assert torch.equal(e["wf"], w.permute(0, 2, 3, 1).reshape(O_, -1).bfloat16())
        ref_d = torch.zeros(I_, T_, opad, device=DEV)
        ref_d[:, :, :O_] = w.permute(1, 2, 3, 0).reshape(I_, T_, O_)
        assert torch.equal(e["wd"], ref_d.reshape(I_, -1).bfloat16())
        k = int(T_ ** 0.5)
        ref_g = 0.5 * e["g_src"].reshape(O_, k, k, I_).permute(0, 3, 1, 2)
        assert torch.equal(e["g_dst"], ref_g.contiguous())


def test_batched_weight_prep_from_channels_last_masters():
    """The engine keeps its convolution masters channels-last ([O][kh][kw][I] in memory): same operands out of the batched
    prep kernel as from OIHW masters, including ragged 32-tiles (I = 48, O = 150) and a pointwise entry next to 3x3 ones."""
    from mit_semseg.engine import ops
    g = _gen(10)
    specs = [(150, 512, 1), (64, 48, 9), (256, 128, 9), (96, 200, 9), (96, 200, 1), (40, 72, 1), (2048, 1024, 1)]
    entries = []
    for O_, I_, T_ in specs:
        k = int(T_ ** 0.5)
        w = torch.randn(O_, I_, k, k, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
        opad = (O_ + 63) // 64 * 64
        entries.append(dict(w=w, wf=torch.empty(O_, T_ * I_, device=DEV, dtype=torch.bfloat16),
                            wd=torch.zeros(I_, T_ * opad, device=DEV, dtype=torch.bfloat16), g_src=None, g_dst=None,
                            O=O_, I=I_, T=T_, o_pad=opad, channels_last=True))
    tab = ops.WeightTable(entries, DEV)
    tab.prep()
    torch.cuda.synchronize()
    for e in entries:
        w, O_, I_, T_, opad = e["w"], e["O"], e["I"], e["T"], e["o_pad"]
        assert torch.equal(e["wf"], w.permute(0, 2, 3, 1).reshape(O_, -1).bfloat16())
        ref_d = torch.zeros(I_, T_, opad, device=DEV)
        ref_d[:, :, :O_] = w.permute(1, 2, 3, 0).reshape(I_, T_, O_)
        assert torch.equal(e["wd"], ref_d.reshape(I_, -1).bfloat16())
    # the split the step programs use: forward operands alone (the direct cast path where the tile allows it) and the
    # data-gradient operands alone on a thin grid (3 blocks walking all tiles)
    fwd_only = [dict(e, wf=torch.zeros_like(e["wf"]), wd=None) for e in entries]
    dgrad_only = [dict(e, wf=None, wd=torch.zeros_like(e["wd"])) for e in entries]
    ops.WeightTable(fwd_only, DEV).prep()
    ops.WeightTable(dgrad_only, DEV).prep(max_blocks=3)
    torch.cuda.synchronize()
    for e, f, d in zip(entries, fwd_only, dgrad_only):
        assert torch.equal(f["wf"], e["wf"]) and torch.equal(d["wd"], e["wd"])


def test_fused_sgd_matches_torch_sgd():
    """One-launch multi-tensor SGD == torch.optim.SGD(momentum, weight_decay) with train.py's two parameter groups."""
    from mit_semseg.engine.optim import FusedSGD
    g = _gen(11)
    shapes = [(512, 4096, 3, 3), (150, 512, 1, 1), (150,), (64,), (2048,), (64, 3, 3, 3), (7,)]
    ref_p = [torch.randn(s, device=DEV, generator=g).requires_grad_(True) for s in shapes]
    my_p = [p.detach().clone().requires_grad_(True) for p in ref_p]
    # the engine keeps its conv masters channels-last ([O][kh][kw][I] in memory) and hands out gradients that are VIEWS of
    # the GEMM's [O][tap * I] buffer (for a 1x1 conv: strides that differ from the parameter's only on size-1 dimensions)
    my_p[0].data = my_p[0].data.contiguous(memory_format=torch.channels_last)
    decay = [0, 1, 5]
    mk = lambda ps: [dict(params=[ps[i] for i in decay]), dict(params=[ps[i] for i in range(len(ps)) if i not in decay],
                                                            weight_decay=0.0)]
    ref = torch.optim.SGD(mk(ref_p), lr=0.02, momentum=0.9, weight_decay=1e-4)
    mine = FusedSGD(mk(my_p), lr=0.02, momentum=0.9, weight_decay=1e-4)
    grads = [torch.empty_like(p) for p in my_p]     # static gradient buffers, as the engine provides (empty_like keeps strides)
    O1, I1 = shapes[1][0], shapes[1][1]
    grads[1] = torch.empty(O1, I1, device=DEV).as_strided((O1, I1, 1, 1), (I1, 1, I1, I1))
    for step in range(4):
        for i, (a, b) in enumerate(zip(ref_p, my_p)):
            gr = torch.randn(a.shape, device=DEV, generator=g)
            a.grad = gr.clone()
            grads[i].copy_(gr)
            b.grad = grads[i]
        for grp_r, grp_m in zip(ref.param_groups, mine.param_groups):
            grp_r["lr"] = grp_m["lr"] = 0.02 * (1 - step / 10) ** 0.9   # train.py:130-139 poly schedule
        ref.step()
        mine.step()
    torch.cuda.synchronize()
    for a, b in zip(ref_p, my_p):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), (a - b).abs().max().item()


def test_bn_finalize_apply_fused_equals_two_kernels():
    from mit_semseg.engine import ops
    g = _gen(13)
    n, hw, C = 2, 24, 256
    y = (torch.randn(n, hw, hw, C, device=DEV, generator=g) * 1.5 + 0.3).bfloat16()
    res = torch.randn(n, hw, hw, C, device=DEV, generator=g).bfloat16()
    gamma, beta = torch.rand(C, device=DEV, generator=g) + 0.5, torch.randn(C, device=DEV, generator=g) * 0.1
    yf = y.float()
    ssum, ssq, cnt = yf.sum(dim=(0, 1, 2)), (yf * yf).sum(dim=(0, 1, 2)), n * hw * hw
    outs = []
    for fused in (False, True):
        rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
        mean, invstd, scale, shift = [torch.empty(C, device=DEV) for _ in range(4)]
        out = torch.empty_like(y)
        if fused:
            ops.bn_finalize_apply(ssum, ssq, cnt, gamma, beta, 1e-5, 0.001, mean, invstd, scale, shift, y, out, relu=True,
                                  res=res, running_mean=rm, running_var=rv)
        else:
            ops.bn_finalize(ssum, ssq, cnt, gamma, beta, 1e-5, 0.001, ops.BN_TRAIN, mean, invstd, scale, shift,
                            running=(rm, rv, None, None, None), update_running=True)
            ops.bn_apply(y, scale, shift, out, relu=True, res=res)
        torch.cuda.synchronize()
        outs.append((out, mean, invstd, scale, shift, rm, rv))
    for a, b in zip(outs[0], outs[1]):
        assert torch.allclose(a.float(), b.float(), rtol=1e-5, atol=1e-6)


def test_exchange_unit_sum_terms_and_relu_backward():
    """sseg_sum_terms / sseg_relu_mask_bwd against torch (models/hrnet.py:225-250): identity + affine same-resolution
    term + two bilinearly sampled lower-resolution affine terms, ReLU; then the ReLU adjoint with the identity share."""
    from mit_semseg.engine import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    n, h, w, c = 2, 24, 40, 48
    x0 = torch.randn(n, h, w, c, device="cuda", generator=g).bfloat16()
    y1 = torch.randn(n, h, w, c, device="cuda", generator=g).bfloat16()
    y2 = torch.randn(n, h // 2, w // 2, c, device="cuda", generator=g).bfloat16()
    y3buf = torch.randn(n, h // 8, w // 8, c + 16, device="cuda", generator=g).bfloat16()
    y3 = y3buf[..., :c]                                        # pixel stride wider than the channel count
    sc = [torch.rand(c, device="cuda", generator=g) + 0.5 for _ in range(3)]
    sh = [torch.randn(c, device="cuda", generator=g) * 0.3 for _ in range(3)]
    out = torch.empty(n, h, w, c, device="cuda", dtype=torch.bfloat16)
    ops.sum_terms([(x0, None, None), (y1, sc[0], sh[0]), (y2, sc[1], sh[1]), (y3, sc[2], sh[2])], out, relu=True)

    def up(t):
        return F.interpolate(t.float().permute(0, 3, 1, 2), size=(h, w), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    ref = x0.float() + (y1.float() * sc[0] + sh[0]) + (up(y2) * sc[1] + sh[1]) + (up(y3) * sc[2] + sh[2])
    ref = torch.relu(ref)
    torch.cuda.synchronize()
    assert (out.float() - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item()
    # two terms, no ReLU
    out2 = torch.empty_like(out)
    ops.sum_terms([(x0, None, None), (y2, sc[1], sh[1])], out2, relu=False)
    ref2 = x0.float() + up(y2) * sc[1] + sh[1]
    assert (out2.float() - ref2).abs().max().item() <= 2 ** -8 * ref2.abs().max().item()
    # ReLU adjoint + identity share (overwrite, then accumulate)
    gr = (torch.randn(n, h, w, c, device="cuda", generator=g) * 0.1).bfloat16()
    ds = torch.empty_like(gr)
    acc = torch.full_like(gr, float("nan"))
    ops.relu_mask_bwd(gr, out, ds, acc, accumulate=False)
    ref_ds = gr.float() * (out.float() > 0)
    assert torch.equal(ds.float(), ref_ds) and torch.equal(acc, ds)
    ops.relu_mask_bwd(gr, out, ds, acc, accumulate=True)
    assert (acc.float() - 2 * ref_ds).abs().max().item() <= 2 ** -8 * ref_ds.abs().max().item()
    ds2 = torch.empty_like(gr)
    ops.relu_mask_bwd(gr, out, ds2)
    assert torch.equal(ds2, ds)
