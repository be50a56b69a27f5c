"""-m gpu: the tcgen05 implicit-GEMM convolution (sseg_conv_igemm) against torch's fp32 convolution.

Inputs are rounded to bf16 first, so the only differences are fp32 accumulation order and (for bf16
outputs) the final rounding: tolerance 2^-8 relative to the output scale (bf16 half-ulp is 2^-9).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref_conv(xs, w, dilation, bias=None):
    x = torch.cat([t.float() for t in xs], dim=3).permute(0, 3, 1, 2)
    k = w.shape[-1]
    y = F.conv2d(x, w.float(), bias=bias, padding=dilation * (k // 2), dilation=dilation)
    return y.permute(0, 2, 3, 1).contiguous()


def _acts(n, h, w_, cins, g, scale=1.0):
    """bf16 NHWC sources; channel counts that are not multiples of 8 live in a wider (x8) zero-padded buffer."""
    xs = []
    for c in cins:
        buf = torch.zeros(n, h, w_, (c + 7) // 8 * 8, device="cuda", dtype=torch.bfloat16)
        buf[..., :c] = (torch.randn(n, h, w_, c, device="cuda", generator=g) * scale).bfloat16()
        xs.append(buf[..., :c])
    return xs


def _run(n, h, w_, cins, cout, k, d, out_f32=False, bias=False, addend=False, stats=False, seed=0):
    from mit_semseg.engine import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(seed)
    xs = _acts(n, h, w_, cins, g)
    cin = sum(cins)
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) * (2.0 / (cin * k * k)) ** 0.5).bfloat16()
    b = torch.randn(cout, device="cuda", generator=g) if bias else None
    ref = _ref_conv(xs, wt, d, b)
    n_store = (cout + 7) // 8 * 8
    ld = n_store + 8 if out_f32 else n_store
    out = torch.full((n, h, w_, ld), float("nan"), device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    add = None
    if addend:
        add = torch.randn(n, h, w_, n_store, device="cuda", generator=g).bfloat16()
        ref = ref + add[..., :cout].float()
    ssum = torch.zeros(cout, device="cuda") if stats else None
    ssq = torch.zeros(cout, device="cuda") if stats else None
    K = k * k * cin
    w_ohwi = torch.zeros(cout, (K + 7) // 8 * 8, device="cuda", dtype=torch.bfloat16)   # row pitch: 16-byte multiple
    w_ohwi[:, :K] = wt.permute(0, 2, 3, 1).reshape(cout, -1)
    w_ohwi = w_ohwi[:, :K]
    ops.conv_igemm(ops.make_geom(xs, ops.conv_taps(k, d)), w_ohwi, cout, out, n_store=n_store, bias=b, addend=add,
                   stat_sum=ssum, stat_sqsum=ssq)
    torch.cuda.synchronize()
    got = out[..., :cout].float()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert torch.isfinite(got).all()
    tol = (2 ** -8 if not out_f32 else 1e-4) * scale
    assert err <= tol, "max err %g > tol %g (scale %g)" % (err, tol, scale)
    if n_store > cout:
        assert (out[..., cout:n_store].float().abs().max().item() == 0.0)
    if stats:
        rr = ref if out_f32 else ref.bfloat16().float()  # statistics describe the stored (rounded) tensor
        rs, rq = rr.sum(dim=(0, 1, 2)), (rr * rr).sum(dim=(0, 1, 2))
        assert (ssum - rs).abs().max().item() <= 1e-3 * max(1.0, rs.abs().max().item()) + 1e-2 * (n * h * w_) ** 0.5
        assert (ssq - rq).abs().max().item() <= 1e-3 * rq.abs().max().item()


def test_pointwise_basic():
    _run(2, 64, 64, [256], 128, 1, 1)


def test_pointwise_wide_k_and_stats():
    _run(2, 64, 64, [2048], 512, 1, 1, stats=True)


def test_pointwise_cout64_large_map():
    _run(1, 128, 128, [128], 64, 1, 1, stats=True)


@pytest.mark.parametrize("d", [1, 2, 4])
def test_3x3_dilated(d):
    _run(2, 64, 64, [128], 128, 3, d, stats=True)


def test_3x3_cout256_cin512():
    _run(2, 64, 64, [512], 256, 3, 1)


def test_3x3_wide_rows():
    _run(1, 32, 256, [64], 64, 3, 1, stats=True)


def test_virtual_concat_3x3():
    _run(2, 64, 64, [256, 64, 64, 128, 64], 128, 3, 1, stats=True)


def test_classifier_f32_bias():
    _run(2, 64, 64, [512], 150, 1, 1, out_f32=True, bias=True)


def test_ragged_spatial():
    _run(2, 38, 50, [64], 128, 3, 2, stats=True)
    _run(3, 6, 6, [128], 64, 3, 1, stats=True)
    _run(1, 5, 300, [64], 64, 1, 1, stats=True)


def test_addend():
    _run(2, 64, 64, [128], 256, 3, 1, addend=True)


def test_channel_counts_not_multiple_of_64():
    """HRNetV2 widths (48 / 96 / 192 / 384, concat 720, C1 hidden 180): partial 64-channel K blocks are zero-filled by
    TMA, partial N tiles masked in the epilogue (models/hrnet.py:257-262, models/models.py:327-346)."""
    _run(2, 32, 32, [48], 48, 3, 1, stats=True)
    _run(2, 16, 16, [96], 96, 3, 1, stats=True)
    _run(2, 16, 16, [96], 48, 1, 1, stats=True)            # exchange unit, branch 1 -> branch 0
    _run(1, 32, 32, [48, 96, 192, 384], 180, 3, 1, stats=True)   # C1 on the 720-channel concat
    _run(2, 16, 16, [180], 150, 1, 1, out_f32=True, bias=True)   # classifier over the 180-channel hidden layer
    _run(2, 8, 8, [384], 384, 3, 1, addend=True)


# ------------------------------------------------------------------ weight gradient (MN-major operands, split-K)
def _run_wgrad(n, h, w_, cins, cout, k, d, cpad=0, seed=0):
    from mit_semseg.engine import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    xs = _acts(n, h, w_, cins, g)
    cin = sum(cins)
    dy = torch.zeros(n, h, w_, cout + cpad, device="cuda", dtype=torch.bfloat16)
    dy[..., :cout] = (torch.randn(n, h, w_, cout, device="cuda", generator=g) * 0.1).bfloat16()
    x = torch.cat([t.float() for t in xs], dim=3).permute(0, 3, 1, 2).contiguous().requires_grad_(False)
    wt = torch.zeros(cout, cin, k, k, device="cuda", requires_grad=True)
    y = F.conv2d(x, wt, padding=d * (k // 2), dilation=d)
    (ref,) = torch.autograd.grad(y, wt, dy[..., :cout].float().permute(0, 3, 1, 2))
    ref = ref.permute(0, 2, 3, 1).reshape(cout, -1)
    dw = torch.zeros(cout, k * k * cin, device="cuda")
    ops.conv_wgrad(ops.make_geom(xs, ops.conv_taps(k, d)), dy, cout, dw)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (dw - ref).abs().max().item()
    assert err <= 2e-4 * scale + 1e-5, "wgrad max err %g (scale %g)" % (err, scale)


def test_wgrad_pointwise():
    _run_wgrad(2, 64, 64, [256], 128, 1, 1)


@pytest.mark.parametrize("d", [1, 2, 4])
def test_wgrad_3x3(d):
    _run_wgrad(2, 64, 64, [128], 256, 3, d)


def test_wgrad_cout64_cin64():
    _run_wgrad(1, 128, 128, [64], 64, 3, 1)


def test_wgrad_concat():
    _run_wgrad(2, 32, 32, [256, 128, 128], 128, 3, 1)


def test_wgrad_classifier_padded():
    _run_wgrad(2, 64, 64, [512], 150, 1, 1, cpad=10)


def test_wgrad_ragged():
    _run_wgrad(2, 38, 50, [64], 128, 3, 2)
    _run_wgrad(3, 6, 6, [128], 64, 3, 1)


def test_wgrad_channel_counts_not_multiple_of_64():
    _run_wgrad(2, 32, 32, [48], 48, 3, 1)
    _run_wgrad(2, 16, 16, [96], 48, 1, 1)
    _run_wgrad(2, 16, 16, [192], 96, 3, 1)
    _run_wgrad(1, 32, 32, [48, 96, 192, 384], 180, 3, 1, cpad=4)
    _run_wgrad(2, 16, 16, [180], 150, 1, 1, cpad=42)


# ------------------------------------------------------------------ stride-2 convs over parity-plane views
@pytest.mark.parametrize("k,cin,cout", [(1, 128, 256), (3, 128, 256), (3, 48, 96), (3, 96, 96), (3, 192, 384)])
def test_stride2_via_parity_planes(k, cin, cout):
    from mit_semseg.engine import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    n, h, w_ = 2, 64, 64
    x = torch.randn(n, h, w_, cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) * 0.05).bfloat16()
    xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wr = wt.float().requires_grad_(True)
    ref = F.conv2d(xr, wr, stride=2, padding=k // 2)
    dh, dw, src = ops.conv_s2_taps(k)
    geom = ops.make_geom(ops.parity_planes(x), (dh, dw), tap_src=src, tap_koff=[t * cin for t in range(k * k)])
    out = torch.empty(n, h // 2, w_ // 2, cout, device="cuda", dtype=torch.bfloat16)
    ops.conv_igemm(geom, wt.permute(0, 2, 3, 1).reshape(cout, -1).contiguous(), cout, out)
    ref_nhwc = ref.detach().permute(0, 2, 3, 1)
    assert (out.float() - ref_nhwc).abs().max().item() <= 2 ** -8 * ref_nhwc.abs().max().item()
    # weight gradient through the same geometry
    dy = (torch.randn(n, h // 2, w_ // 2, cout, device="cuda", generator=g) * 0.1).bfloat16()
    gx, gw = torch.autograd.grad(ref, (xr, wr), dy.float().permute(0, 3, 1, 2))
    dwbuf = torch.zeros(cout, k * k * cin, device="cuda")
    ops.conv_wgrad(geom, dy, cout, dwbuf)
    gw_ref = gw.permute(0, 2, 3, 1).reshape(cout, -1)
    assert (dwbuf - gw_ref).abs().max().item() <= 2e-4 * gw_ref.abs().max().item() + 1e-5
    # data gradient: one launch per input parity plane, writing straight into the strided plane view of dx
    wd = torch.zeros(cin, k * k * cout, device="cuda", dtype=torch.bfloat16)
    ops.prep_conv_weight(wt.float().contiguous(), None, wd, o_pad=cout)
    dx = torch.zeros(n, h, w_, cin, device="cuda", dtype=torch.bfloat16)
    planes = ops.parity_planes(dx)
    for pl in range(4):
        taps = [t for t in range(k * k) if src[t] == pl]
        if not taps:
            continue
        geom_d = ops.make_geom([dy], ([-dh[t] for t in taps], [-dw[t] for t in taps]),
                               tap_koff=[t * cout for t in taps])
        ops.conv_igemm(geom_d, wd, cin, planes[pl])
    gx_nhwc = gx.permute(0, 2, 3, 1)
    assert (dx.float() - gx_nhwc).abs().max().item() <= 2 ** -7 * gx_nhwc.abs().max().item()


# ------------------------------------------------------------------ dgrad with the producer's BN-backward reduce fused
@pytest.mark.parametrize("k,hw,cprod,cout_next", [(1, 64, 128, 256), (3, 64, 128, 256), (3, 38, 128, 256),
                                                  (3, 32, 48, 48), (3, 16, 96, 192)])
def test_dgrad_with_fused_bn_backward_reduce(k, hw, cprod, cout_next):
    from mit_semseg.engine import ops
    g = torch.Generator(device="cuda").manual_seed(21)
    n = 2          # producer layer has cprod channels; the consumer conv maps cprod -> cout_next
    y = torch.randn(n, hw, hw, cprod, device="cuda", generator=g).bfloat16()       # producer's saved conv output
    fscale = torch.rand(cprod, device="cuda", generator=g) + 0.5
    fshift = torch.randn(cprod, device="cuda", generator=g) * 0.3
    dy_next = (torch.randn(n, hw, hw, cout_next, device="cuda", generator=g) * 0.1).bfloat16()
    w = (torch.randn(cout_next, cprod, k, k, device="cuda", generator=g) * 0.05).bfloat16()
    wd = torch.zeros(cprod, k * k * cout_next, device="cuda", dtype=torch.bfloat16)
    ops.prep_conv_weight(w.float().contiguous(), None, wd, o_pad=cout_next)
    dh, dw = ops.conv_taps(k, 1)
    gd = ops.make_geom([dy_next], ([-v for v in dh], [-v for v in dw]), tap_koff=[t * cout_next for t in range(k * k)])
    gout = torch.empty(n, hw, hw, cprod, device="cuda", dtype=torch.bfloat16)
    s1, s2 = torch.zeros(cprod, device="cuda"), torch.zeros(cprod, device="cuda")
    ops.conv_igemm_bnbwd(gd, wd, cprod, gout, y, fscale, fshift, s1, s2)
    ref_g = torch.empty_like(gout)
    ops.conv_igemm(gd, wd, cprod, ref_g, n_store=cprod)
    torch.cuda.synchronize()
    assert torch.equal(gout, ref_g)
    mask = (y.float() * fscale + fshift) > 0
    gp = ref_g.float() * mask
    r1, r2 = gp.sum(dim=(0, 1, 2)), (gp * y.float()).sum(dim=(0, 1, 2))
    assert (s1 - r1).abs().max().item() <= 1e-3 * r1.abs().max().item() + 1e-4
    assert (s2 - r2).abs().max().item() <= 1e-3 * r2.abs().max().item() + 1e-4
    # raw -> xhat conversion inside bn_bwd_apply reproduces the two-pass result (dy and dgamma)
    mean = torch.randn(cprod, device="cuda", generator=g) * 0.2
    invstd = torch.rand(cprod, device="cuda", generator=g) + 0.5
    s1b, s2b = torch.zeros(cprod, device="cuda"), torch.zeros(cprod, device="cuda")
    ops.bn_bwd_reduce(ref_g, None, y, mean, invstd, s1b, s2b, scale=fscale, fshift=fshift)
    dya, dyb, dgam = torch.empty_like(y), torch.empty_like(y), torch.zeros(cprod, device="cuda")
    cnt = n * hw * hw
    ops.bn_bwd_apply(ref_g, None, y, mean, invstd, fscale, s1b, s2b, cnt, dya, fshift=fshift)
    ops.bn_bwd_apply(gout, None, y, mean, invstd, fscale, s1, s2, cnt, dyb, fshift=fshift, s2_raw=True, dgamma_out=dgam)
    torch.cuda.synchronize()
    assert (dgam - s2b).abs().max().item() <= 2e-3 * s2b.abs().max().item() + 1e-3
    assert (dya.float() - dyb.float()).abs().max().item() <= 2 ** -6 * dya.float().abs().max().item()


@pytest.mark.parametrize("k,hw,cprod,cout_next", [(1, 64, 256, 64), (1, 38, 128, 256), (3, 32, 64, 64), (1, 16, 512, 128)])
def test_dgrad_with_fused_bn_backward_reduce_of_a_shortcut_layer(k, hw, cprod, cout_next):
    """sseg_conv_igemm_bnbwd_res: the data-gradient launch that ADDS THE LAST CONTRIBUTION to the gradient of a residual
    block's output a = relu(bn(y) + shortcut) also reduces that block's BN backward, the ReLU mask coming from the saved
    output: out = dgrad + addend (bit-identical to the plain accumulate launch), s1 = sum out*[a>0], s2 = sum out*[a>0]*y;
    then sseg_bn_bwd_apply(a=..., dres=..., s2_raw) reproduces the separate reduce + apply pair (dy, shortcut gradient,
    dgamma). Reference: autograd of models/resnet.py:37-53,72-92 through lib/nn/modules/batchnorm.py:58-61."""
    from mit_semseg.engine import ops
    g = torch.Generator(device="cuda").manual_seed(22)
    n = 2
    y = torch.randn(n, hw, hw, cprod, device="cuda", generator=g).bfloat16()        # producer's saved conv output
    a = torch.relu(torch.randn(n, hw, hw, cprod, device="cuda", generator=g)).bfloat16()   # its saved block output
    prev = (torch.randn(n, hw, hw, cprod, device="cuda", generator=g) * 0.1).bfloat16()    # gradient accumulated so far
    dy_next = (torch.randn(n, hw, hw, cout_next, device="cuda", generator=g) * 0.1).bfloat16()
    w = (torch.randn(cout_next, cprod, k, k, device="cuda", generator=g) * 0.05).bfloat16()
    wd = torch.zeros(cprod, k * k * cout_next, device="cuda", dtype=torch.bfloat16)
    ops.prep_conv_weight(w.float().contiguous(), None, wd, o_pad=cout_next)
    dh, dw = ops.conv_taps(k, 1)
    gd = ops.make_geom([dy_next], ([-v for v in dh], [-v for v in dw]), tap_koff=[t * cout_next for t in range(k * k)])
    gout = prev.clone()
    s1, s2 = torch.zeros(cprod, device="cuda"), torch.zeros(cprod, device="cuda")
    ops.conv_igemm_bnbwd_res(gd, wd, cprod, gout, y, a, s1, s2, addend=gout)
    ref_g = prev.clone()
    ops.conv_igemm(gd, wd, cprod, ref_g, n_store=cprod, addend=ref_g)
    torch.cuda.synchronize()
    assert torch.equal(gout, ref_g)
    gp = ref_g.float() * (a.float() > 0)
    r1, r2 = gp.sum(dim=(0, 1, 2)), (gp * y.float()).sum(dim=(0, 1, 2))
    assert (s1 - r1).abs().max().item() <= 1e-3 * r1.abs().max().item() + 1e-4
    assert (s2 - r2).abs().max().item() <= 1e-3 * r2.abs().max().item() + 1e-4
    mean = torch.randn(cprod, device="cuda", generator=g) * 0.2
    invstd = torch.rand(cprod, device="cuda", generator=g) + 0.5
    scale = torch.rand(cprod, device="cuda", generator=g) + 0.5
    s1b, s2b = torch.zeros(cprod, device="cuda"), torch.zeros(cprod, device="cuda")
    ops.bn_bwd_reduce(ref_g, a, y, mean, invstd, s1b, s2b)
    dya, dyb, dra, drb = torch.empty_like(y), torch.empty_like(y), torch.empty_like(y), torch.empty_like(y)
    dgam = torch.zeros(cprod, device="cuda")
    cnt = n * hw * hw
    ops.bn_bwd_apply(ref_g, a, y, mean, invstd, scale, s1b, s2b, cnt, dya, dres=dra)
    ops.bn_bwd_apply(gout, a, y, mean, invstd, scale, s1, s2, cnt, dyb, dres=drb, s2_raw=True, dgamma_out=dgam)
    torch.cuda.synchronize()
    assert torch.equal(dra, drb)
    assert (dgam - s2b).abs().max().item() <= 2e-3 * s2b.abs().max().item() + 1e-3
    assert (dya.float() - dyb.float()).abs().max().item() <= 2 ** -6 * dya.float().abs().max().item()


@pytest.mark.parametrize("n,hw,cin,cout,k", [(2, 32, 64, 128, 1), (2, 24, 96, 200, 3), (1, 64, 256, 512, 1)])
def test_conv_with_bn_finalize_by_the_last_cta(n, hw, cin, cout, k):
    """sseg_conv_igemm_bnfin = sseg_conv_igemm(stats) + sseg_bn_finalize(SSEG_BN_TRAIN): same y (bit for bit), same
    statistics, and mean / inv_std / scale / shift / running statistics equal to the separate finalize kernel's up to the
    summation order of the atomics (F.batch_norm training branch, lib/nn/modules/batchnorm.py:58-61)."""
    from mit_semseg.engine import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    xs = _acts(n, hw, hw, [cin], g)
    K = k * k * cin
    wt = (torch.randn(cout, K, device="cuda", generator=g) * (2.0 / K) ** 0.5).bfloat16()
    cp = (cout + 7) // 8 * 8
    gamma, beta = 0.5 + torch.rand(cout, device="cuda", generator=g), torch.randn(cout, device="cuda", generator=g) * 0.1
    geom = ops.make_geom(xs, ops.conv_taps(k, 1))
    count = float(n * hw * hw)
    # reference: two launches
    y0 = torch.zeros(n, hw, hw, cp, device="cuda", dtype=torch.bfloat16)
    s0, q0 = torch.zeros(cout, device="cuda"), torch.zeros(cout, device="cuda")
    ops.conv_igemm(geom, wt, cout, y0, stat_sum=s0, stat_sqsum=q0)
    v0 = torch.zeros(4, cout, device="cuda")
    rm0, rv0 = torch.full((cout,), 0.25, device="cuda"), torch.full((cout,), 1.5, device="cuda")
    ops.bn_finalize(s0, q0, count, gamma, beta, 1e-5, 0.1, ops.BN_TRAIN, v0[0], v0[1], v0[2], v0[3],
                    running=(rm0, rv0, None, None, None), update_running=True)
    # fused
    y1 = torch.zeros_like(y0)
    s1, q1 = torch.zeros(cout, device="cuda"), torch.zeros(cout, device="cuda")
    v1 = torch.zeros(4, cout, device="cuda")
    rm1, rv1 = torch.full((cout,), 0.25, device="cuda"), torch.full((cout,), 1.5, device="cuda")
    counter = torch.zeros(1, device="cuda", dtype=torch.float32)
    bn = ops.make_bn_fused(gamma, beta, 1e-5, 0.1, count, s1, q1, counter, v1[0], v1[1], v1[2], v1[3], running_mean=rm1,
                           running_var=rv1)
    ops.conv_igemm_bnfin(geom, wt, cout, y1, bn)
    torch.cuda.synchronize()
    assert torch.equal(y0, y1)
    assert torch.allclose(s0, s1, rtol=1e-5, atol=1e-3) and torch.allclose(q0, q1, rtol=1e-5, atol=1e-3)
    assert counter.view(torch.int32).item() > 0
    for a, b, what in ((v0[0], v1[0], "mean"), (v0[1], v1[1], "inv_std"), (v0[2], v1[2], "scale"), (v0[3], v1[3], "shift"),
                       (rm0, rm1, "running_mean"), (rv0, rv1, "running_var")):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), what
