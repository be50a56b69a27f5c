"""CPU: whole step programs EXECUTED against `tests/abi_emulator.py` (a Python restatement of the C ABI's documented
semantics) and compared with the oracle — the schedules' arithmetic wiring without a GPU.

 * default schedule vs the bf16-emulating oracle: frozen BN pins every gradient route; train-mode BN pins loss / features
   (two bf16 pipelines decorrelate through ReLU-mask flips);
 * every opt-in schedule (fused conv+BN kernels, folded eval BN, branch streams, overlapped re-layout, fused multi-scale
   head) vs the default schedule on the same emulator: same loss, same gradients, same probabilities.
The CUDA kernels themselves are covered by the `-m gpu` tests; this file covers which kernel gets which buffer and
coefficient vector, in which order."""
import statistics

import pytest
import torch
import torch.nn as nn

from abi_emulator import EmuLib
from oracle import segnet_oracle as O
from test_program_dry import _seg

SWITCHES = ("SSEG_BRANCH_STREAMS", "SSEG_FOLD_BN_EVAL", "SSEG_OVERLAP_RELAYOUT", "SSEG_COOP_BN")


@pytest.fixture
def emu(monkeypatch):
    from mit_semseg.engine import _C, ops
    lib = EmuLib()
    monkeypatch.setattr(_C, "lib", lambda: lib)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    for name in SWITCHES:
        monkeypatch.delenv(name, raising=False)
    return lib


def _load(seg, enc_arch, dec_arch, fc, gain=0.25, bias_shift=0.0, calibrate_on=None):
    esd = O.synth_state_dict(O.encoder_param_shapes(enc_arch), 304, gain)
    dsd = O.synth_state_dict(O.decoder_param_shapes(dec_arch, fc), 305)
    if bias_shift:
        for sd in (esd, dsd):
            for k in sd:
                if k.endswith(".bias") and (k[:-5] + ".running_mean") in sd:
                    sd[k] = sd[k] + bias_shift
    if calibrate_on is not None:
        with torch.no_grad():
            st = O.BNState(True, update_running=True, momentum=1.0)
            O.decoder_forward(O.encoder_forward(calibrate_on["img_data"], esd, enc_arch, st), dsd, dec_arch, st, dropout_p=0.0)
    seg.encoder.load_state_dict(esd)
    seg.decoder.load_state_dict(dsd)
    return esd, dsd


def _run_train(seg, feed):
    from mit_semseg.engine import program as PR
    P = PR.SegProgram(seg, tuple(feed["img_data"].shape), training=True, with_grad=True, dry_run=True)
    P.dry_run, P.serial = False, True
    P.load_inputs(feed["img_data"], feed["seg_label"])
    P.run_eager()
    return P


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _freeze(seg):
    for m in seg.modules():
        if isinstance(m, nn.modules.batchnorm._BatchNorm):
            m.eval()
        if isinstance(m, nn.Dropout2d):
            m.p = 0.0


@pytest.mark.parametrize("enc,dec,fc,stride,hw", [("resnet50dilated", "ppm_deepsup", 2048, 8, 64),   # the bench network
                                                  ("resnet18dilated", "ppm_deepsup", 512, 8, 64),
                                                  ("resnet50", "upernet", 2048, 4, 64),
                                                  ("hrnetv2", "c1", 720, 4, 64)])
def test_default_schedule_frozen_bn_matches_the_oracle(enc, dec, fc, stride, hw, emu):
    seg = _seg(enc, dec, fc)
    feed = O.synth_batch(2, hw, hw, stride, 1)
    esd, dsd = _load(seg, enc, dec, fc, bias_shift=2.0 if enc == "hrnetv2" else 0.0,
                     calibrate_on=feed if enc == "hrnetv2" else None)
    seg.train()
    _freeze(seg)
    P = _run_train(seg, feed)
    e = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in esd.items()}
    d = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in dsd.items()}
    ds = seg.deep_sup_scale
    l_ref, a_ref = O.segmentation_forward(feed, e, d, enc, dec, O.BNState(False, emulate="bf16"), ds, dropout_p=0.0)
    l_ref.backward()
    assert abs(P.out[0].item() - l_ref.item()) <= 2e-3 * abs(l_ref.item())
    grads = P.param_grads()
    rels = {}
    for prefix, net, sd in (("enc.", seg.encoder, e), ("dec.", seg.decoder, d)):
        for name, p in net.named_parameters():
            rels[prefix + name] = _rel(grads[p], sd[name].grad)
    worst = max(rels, key=rels.get)
    print(enc, dec, "loss", P.out[0].item(), l_ref.item(), "grad rel median %.4f max %.4f (%s)" % (
        statistics.median(rels.values()), rels[worst], worst))
    # a mis-routed gradient is an O(1) error; bf16 rounding noise of these tiny (2x2-pixel layer4) runs stays below 10 %
    assert statistics.median(rels.values()) <= 0.08 and rels[worst] <= 0.4, (worst, rels[worst])


@pytest.mark.parametrize("enc,dec,fc,stride", [("resnet18dilated", "ppm_deepsup", 512, 8), ("hrnetv2", "c1", 720, 4)])
def test_default_schedule_train_mode_bn_matches_the_oracle(enc, dec, fc, stride, emu):
    seg = _seg(enc, dec, fc)
    feed = O.synth_batch(2, 64, 64, stride, 1)
    esd, dsd = _load(seg, enc, dec, fc)
    seg.train()
    for m in seg.modules():
        if isinstance(m, nn.Dropout2d):
            m.p = 0.0
    rm0 = seg.encoder.bn1.running_mean.clone()
    P = _run_train(seg, feed)
    l_ref, a_ref, feats, out = O.segmentation_forward(feed, dict(esd), dict(dsd), enc, dec, O.BNState(True, emulate="bf16"),
                                                      seg.deep_sup_scale, dropout_p=0.0, return_aux=True)
    assert abs(P.out[0].item() - l_ref.item()) <= 5e-3 * abs(l_ref.item())
    f = P.feats[-1]
    got = torch.cat([q.t for q in f], 3) if isinstance(f, list) else f.t
    assert _rel(got.float().permute(0, 3, 1, 2), feats[-1]) <= 0.1
    assert not torch.equal(seg.encoder.bn1.running_mean, rm0)       # running statistics moved (F.batch_norm semantics)


@pytest.mark.parametrize("enc,dec,fc,stride", [("resnet18dilated", "ppm_deepsup", 512, 8), ("resnet50", "upernet", 2048, 4),
                                               ("hrnetv2", "c1", 720, 4)])
@pytest.mark.parametrize("switch", ["SSEG_COOP_BN", "SSEG_BRANCH_STREAMS", "SSEG_OVERLAP_RELAYOUT"])
def test_opt_in_training_schedules_equal_the_default_schedule(enc, dec, fc, stride, switch, emu, monkeypatch):
    """Train-mode BN, same emulator, same weights: an opt-in schedule must reproduce the default schedule's loss, running
    statistics and gradients (only the split of the arithmetic over kernels differs)."""
    feed = O.synth_batch(2, 64, 64, stride, 2)
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv(switch, flag)
        seg = _seg(enc, dec, fc)
        _load(seg, enc, dec, fc)
        seg.train()
        for m in seg.modules():
            if isinstance(m, nn.Dropout2d):
                m.p = 0.0
        P = _run_train(seg, feed)
        res[flag] = (P.out.clone(), {n: g.clone() for n, g in zip([n for n, _ in seg.named_parameters()],
                                                                    [P.param_grads()[p] for p in seg.parameters()])},
                     {n: b.clone() for n, b in seg.named_buffers() if "running" in n})
        if switch == "SSEG_COOP_BN" and flag == "1":
            assert emu.calls.get("sseg_conv_bn_train", 0) > 0 and emu.calls.get("sseg_conv_dgrad_bn", 0) > 0
    (o0, g0, b0), (o1, g1, b1) = res["0"], res["1"]
    assert torch.allclose(o0, o1, rtol=1e-4, atol=1e-6), (o0, o1)
    for n in b0:
        assert torch.allclose(b0[n], b1[n], rtol=1e-4, atol=1e-6), n
    rels = {n: _rel(g1[n], g0[n]) for n in g0}
    worst = max(rels, key=rels.get)
    print(switch, enc, "grad rel vs default schedule: median %.2e max %.2e (%s)" % (statistics.median(rels.values()), rels[worst], worst))
    # identical arithmetic up to fp32 association; train-mode BN amplifies the resulting rounding flips
    assert statistics.median(rels.values()) <= 1e-2 and rels[worst] <= 0.15, (worst, rels[worst])


@pytest.mark.parametrize("enc,dec,fc", [("resnet18dilated", "ppm_deepsup", 512), ("resnet50", "upernet", 2048), ("hrnetv2", "c1", 720)])
def test_inference_schedules_match_the_oracle_and_each_other(enc, dec, fc, emu, monkeypatch):
    from mit_semseg.engine import program as PR
    feed = O.synth_batch(1, 64, 96, 8, 5)
    hr = enc == "hrnetv2"   # HRNet needs calibrated statistics (its synthetic ones explode); the ResNets run as they are
    probs = {}
    for fold in ("0", "1"):
        monkeypatch.setenv("SSEG_FOLD_BN_EVAL", fold)
        seg = _seg(enc, dec, fc)
        esd, dsd = _load(seg, enc, dec, fc, bias_shift=1.0 if hr else 0.0, calibrate_on=feed if hr else None)
        seg.eval()
        P = PR.SegProgram(seg, (1, 3, 64, 96), training=False, with_grad=False, seg_size=(64, 96), dry_run=True)
        P.dry_run, P.serial = False, True
        P.load_inputs(feed["img_data"])
        P.run_eager()
        probs[fold] = P.probs.clone()
    with torch.no_grad():
        ref = O.segmentation_forward(feed, esd, dsd, enc, dec, O.BNState(False, emulate="bf16"), None, segSize=(64, 96))
    for fold, p in probs.items():
        assert (p.sum(1) - 1).abs().max().item() < 1e-3
        agree = (p.argmax(1) == ref.argmax(1)).float().mean().item()
        # (uncalibrated synthetic statistics saturate the softmax: a flipped pixel is a probability error of 1)
        assert agree >= 0.9 and (not hr or (p - ref).abs().max().item() <= 5e-2), (fold, agree)
    assert (probs["0"].argmax(1) == probs["1"].argmax(1)).float().mean().item() >= 0.9
    # fused multi-scale head == the reference loop (eval.py:63-72), through the program interface
    seg.eval()
    imgs = [torch.randn(1, 3, 64, 96), torch.randn(1, 3, 96, 128)]
    scores = torch.zeros(1, 150, 64, 96)
    loop = torch.zeros(1, 150, 64, 96)
    for im in imgs:
        for head in (dict(), dict(head_out=scores, head_weight=0.5)):
            P = PR.SegProgram(seg, tuple(im.shape), training=False, with_grad=False, seg_size=(64, 96), dry_run=True, **head)
            P.dry_run, P.serial = False, True
            P.load_inputs(im)
            P.run_eager()
            if not head:
                loop += 0.5 * P.probs
    assert torch.allclose(scores, loop, atol=1e-6)


def test_training_loss_curve_follows_the_oracle(emu, monkeypatch):
    """north_star: "the training loss curve overlays the reference". Six SGD steps (train.py's optimiser settings) through
    the public API - SegmentationModule(feed) -> loss.backward() -> torch.optim.SGD.step() - on the emulated ABI, against
    the oracle trained with autograd from the same weights on the same batches."""
    from mit_semseg.engine import functional as EF
    from mit_semseg.engine import program as PR
    real = PR.SegProgram

    def factory(*a, **k):
        k["dry_run"] = True
        prog = real(*a, **k)
        prog.dry_run, prog.serial = False, True
        return prog
    monkeypatch.setattr(EF, "SegProgram", factory)
    monkeypatch.setattr(real, "capture", lambda self, warm=True: None)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    enc, dec, fc = "resnet18dilated", "c1_deepsup", 512
    seg = _seg(enc, dec, fc)
    esd, dsd = _load(seg, enc, dec, fc)
    seg.train()
    opt = torch.optim.SGD(seg.parameters(), lr=0.02, momentum=0.9, weight_decay=1e-4)
    e = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in esd.items()}
    d = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in dsd.items()}
    params = [v for v in list(e.values()) + list(d.values()) if v.requires_grad]
    opt_ref = torch.optim.SGD(params, lr=0.02, momentum=0.9, weight_decay=1e-4)
    ours, ref = [], []
    for step in range(6):
        feed = O.synth_batch(2, 64, 64, 8, 100 + step % 2)          # two alternating batches: the loss must come down
        opt.zero_grad()
        loss, acc = seg(feed)
        loss.backward()
        opt.step()
        ours.append(loss.item())
        opt_ref.zero_grad()
        l_ref, _ = O.segmentation_forward(feed, e, d, enc, dec, O.BNState(True, emulate="bf16", update_running=True), 0.4,
                                          dropout_p=0.0)
        l_ref.backward()
        opt_ref.step()
        ref.append(l_ref.item())
    print("engine schedule:", ["%.4f" % v for v in ours])
    print("oracle         :", ["%.4f" % v for v in ref])
    assert ours[4] < ours[0] and ours[5] < ours[1]                  # it trains (same batch, two / three updates later)
    assert all(abs(a - b) <= 2e-2 * abs(b) for a, b in zip(ours, ref)), (ours, ref)


@pytest.mark.parametrize("enc,dec,fc", [("resnet18dilated", "ppm_deepsup", 512), ("resnet18", "c1", 512)])
def test_accurate_inference_schedule_reaches_fp32_accuracy(enc, dec, fc, emu):
    """BASELINE config 2 (ResNet18dilated + PPM_deepsup inference): logits within 1e-3 of the fp32 reference and the same
    arg-max map. The bf16-pair schedule (engine/accurate.py: three-term products as ONE virtual-concat GEMM launch per
    convolution) against the plain fp32 oracle - and, for scale, the ordinary bf16 schedule on the same input."""
    from mit_semseg.engine import accurate as ACC
    from mit_semseg.engine import program as PR
    seg = _seg(enc, dec, fc)
    esd, dsd = _load(seg, enc, dec, fc)
    seg.eval()
    feed = O.synth_batch(2, 64, 96, 8, 9)
    with torch.no_grad():
        feats = O.encoder_forward(feed["img_data"], esd, enc, O.BNState(False))
        lg = O.decoder_forward(feats, dsd, dec, O.BNState(False), return_logits=True)
        lg = lg[0] if isinstance(lg, tuple) else lg
        ref = O.segmentation_forward(feed, esd, dsd, enc, dec, O.BNState(False), None, segSize=(64, 96))
    A = ACC.AccurateInference(seg, (2, 3, 64, 96), (64, 96), dry_run=True)
    A.dry_run = False
    A.load_inputs(feed["img_data"])
    A.run_eager()
    got = A.logits.permute(0, 3, 1, 2)
    rel = _rel(got, lg)
    agree = (A.probs.argmax(1) == ref.argmax(1)).float().mean().item()
    B = PR.SegProgram(seg, (2, 3, 64, 96), training=False, with_grad=False, seg_size=(64, 96), dry_run=True)
    B.dry_run, B.serial = False, True
    B.load_inputs(feed["img_data"])
    B.run_eager()
    rel_bf16 = _rel(B.logits[..., :150].permute(0, 3, 1, 2), lg)
    print("%s+%s logits rel-L2 vs fp32 oracle: accurate %.2e (bf16 schedule %.2e), arg-max agreement %.6f" % (enc, dec, rel, rel_bf16, agree))
    assert rel <= 1e-3 and rel < rel_bf16 / 20
    assert agree >= 0.9995    # two fp32-grade implementations differ only at exact ties of random-init logits
    assert (A.probs - ref).abs().max().item() <= 1e-3 and (A.probs.sum(1) - 1).abs().max().item() < 1e-4


@pytest.mark.parametrize("coop", ["0", "1"])
def test_dropout2d_masks_reach_the_oracle_loss(coop, emu, monkeypatch):
    """nn.Dropout2d(0.1) active (models/models.py:460,464): injected keep-masks, default and fused conv+BN schedules."""
    from mit_semseg.engine import program as PR
    monkeypatch.setenv("SSEG_COOP_BN", coop)
    enc, dec, fc = "resnet18dilated", "ppm_deepsup", 512
    seg = _seg(enc, dec, fc)
    esd, dsd = _load(seg, enc, dec, fc)
    seg.train()
    feed = O.synth_batch(2, 64, 64, 8, 9)
    g = torch.Generator().manual_seed(5)
    masks = {"main": (torch.rand(2, 512, generator=g) >= 0.3).float(), "deepsup": (torch.rand(2, 128, generator=g) >= 0.3).float()}
    P = PR.SegProgram(seg, (2, 3, 64, 64), training=True, with_grad=True, dry_run=True, dropout_masks=masks)
    P.dry_run, P.serial = False, True
    P.load_inputs(feed["img_data"], feed["seg_label"])
    P.run_eager()
    l_ref, _ = O.segmentation_forward(feed, dict(esd), dict(dsd), enc, dec, O.BNState(True, emulate="bf16"), 0.4,
                                      dropout_p=0.1, masks=masks)
    assert abs(P.out[0].item() - l_ref.item()) <= 5e-3 * abs(l_ref.item()), (P.out[0].item(), l_ref.item())


def test_mobilenetv2dilated_inference_matches_the_oracle(emu):
    """BASELINE configs[0]: MobileNetV2dilated + C1_deepsup, single-image forward (eval.py) - on the engine as an
    inference-only schedule (folded BN + ReLU6, depthwise kernel) against the fp32 oracle pinned to the reference."""
    from mit_semseg.engine import program as PR
    from mit_semseg.models import ModelBuilder, SegmentationModule
    from mit_semseg.models import mobilenet as MB, models as M
    enc_arch, dec_arch, fc = "mobilenetv2dilated", "c1_deepsup", 320
    enc = M.MobileNetV2Dilated(MB.mobilenetv2(pretrained=False), dilate_scale=8)
    dec = ModelBuilder.build_decoder(dec_arch, fc_dim=fc, num_class=150, use_softmax=True)
    seg = SegmentationModule(enc, dec, nn.NLLLoss(ignore_index=-1), 0.4)
    feed = O.synth_batch(1, 96, 128, 8, 5)
    esd, dsd = _load(seg, enc_arch, dec_arch, fc, gain=None, calibrate_on=O.synth_batch(2, 96, 128, 8, 6))
    seg.eval()
    P = PR.SegProgram(seg, (1, 3, 96, 128), training=False, with_grad=False, seg_size=(96, 128), dry_run=True)
    P.dry_run, P.serial = False, True
    P.load_inputs(feed["img_data"])
    P.run_eager()
    with torch.no_grad():
        # MobileNetV2's linear bottlenecks amplify perturbations ~1.4x per block at random init (bf16 vs fp32: 23 % on the
        # last feature map), so the schedule is compared with the oracle rounding where the engine's kernels round
        ref = O.segmentation_forward(feed, esd, dsd, enc_arch, dec_arch, O.BNState(False, emulate="bf16"), None, segSize=(96, 128))
        rfeats = O.encoder_forward(feed["img_data"], esd, enc_arch, O.BNState(False, emulate="bf16"))
    frel = [_rel(f.t.float().permute(0, 3, 1, 2), r) for f, r in zip(P.feats, rfeats)]
    agree = (P.probs.argmax(1) == ref.argmax(1)).float().mean().item()
    err = (P.probs - ref).abs().max().item()
    print("mobilenetv2dilated+c1_deepsup: feature rel %s, arg-max agreement %.4f, max |dp| %.3e" % (
        ["%.4f" % v for v in frel], agree, err))
    assert [tuple(f.t.shape[1:]) for f in P.feats] == [(24, 32, 24), (12, 16, 32), (12, 16, 64), (12, 16, 160), (12, 16, 320)]
    # identical arithmetic up to fp32 association: 1e-4 after two blocks; the growth afterwards is the network's own
    # amplification of those last-bit differences (a wiring error is O(1) at the block where it happens)
    assert max(frel[:3]) <= 1e-2 and max(frel) <= 0.15
    assert emu.calls["sseg_dwconv_affine"] == 17 and emu.calls["sseg_stem_conv_affine"] == 1
    assert agree >= 0.85 and err <= 5e-2
    seg.train()
    with pytest.raises(NotImplementedError, match="eval mode only"):
        PR.SegProgram(seg, (1, 3, 96, 128), training=True, with_grad=True, dry_run=True)


def test_fused_sgd_equals_torch_sgd(emu, monkeypatch):
    """engine/optim.py::FusedSGD (one multi-tensor launch over a chunk table) against torch.optim.SGD - train.py's optimiser
    (train.py:115-127: two parameter groups, weight decay only on conv weights) - over three steps with changing
    gradients, including train.py's adjust_learning_rate rewriting group['lr']."""
    import types
    from mit_semseg.engine.optim import FusedSGD
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: types.SimpleNamespace(cuda_stream=0))
    g = torch.Generator().manual_seed(0)
    shapes = [(64, 3, 3, 3), (64,), (64,), (150, 512, 1, 1), (150,), (70001,)]     # the last one spans two chunks, not x4
    pa = [nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    pb = [nn.Parameter(p.detach().clone()) for p in pa]

    def groups(ps):
        return [{"params": [p for p in ps if p.dim() > 1]}, {"params": [p for p in ps if p.dim() <= 1], "weight_decay": 0.0}]
    oa = FusedSGD(groups(pa), lr=0.02, momentum=0.9, weight_decay=1e-4)
    ob = torch.optim.SGD(groups(pb), lr=0.02, momentum=0.9, weight_decay=1e-4)
    grads = [torch.empty_like(p) for p in pa]          # static gradient buffers, like the engine's
    for p, gr in zip(pa, grads):
        p.grad = gr
    for step in range(3):
        for p, q, gr in zip(pa, pb, grads):
            gr.copy_(torch.randn(p.shape, generator=g))
            q.grad = gr.clone()
        if step == 2:
            for o in (oa, ob):
                for grp in o.param_groups:
                    grp["lr"] = 0.01
        oa.step()
        ob.step()
        for p, q in zip(pa, pb):
            assert torch.allclose(p, q, rtol=1e-6, atol=1e-7)
    assert emu.calls["sseg_sgd_step"] == 6             # one launch per parameter group and step
