"""CPU: the oracle (oracle/segnet_oracle.py) against the committed golden vectors produced by the imported reference
(oracle/make_golden.py), the numpy primitive restatements against torch, and the host-side mirror of the reference API
(module tree, state-dict keys, initialisation, conv hyper-parameters) against the reference's own recorded values."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import np_primitives as NP
from oracle import segnet_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def _oracle_train(enc_arch, dec_arch, fc, n, hw, stride):
    esd = O.synth_state_dict(O.encoder_param_shapes(enc_arch), 304)
    dsd = O.synth_state_dict(O.decoder_param_shapes(dec_arch, fc), 305)
    e = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in esd.items()}
    d = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in dsd.items()}
    ds = 0.4 if dec_arch.endswith("deepsup") else None
    feed = O.synth_batch(n, hw, hw, stride, 1)
    loss, acc, feats, out = O.segmentation_forward(feed, e, d, enc_arch, dec_arch, O.BNState(True), ds, dropout_p=0.0,
                                                   return_aux=True)
    loss.backward()
    return loss, acc, feats, out, e, d


@pytest.mark.parametrize("name,enc,dec,fc,hw,stride", [
    ("train_r50dilated_ppm_deepsup_96", "resnet50dilated", "ppm_deepsup", 2048, 96, 8),
    ("train_r18dilated_c1_deepsup_96", "resnet18dilated", "c1_deepsup", 512, 96, 8),
    ("train_r50_upernet_128", "resnet50", "upernet", 2048, 128, 4),
    ("train_hrnetv2_c1_64", "hrnetv2", "c1", 720, 64, 4),
    ("train_mobilenetv2dilated_c1_deepsup_96", "mobilenetv2dilated", "c1_deepsup", 320, 96, 8),
])
def test_oracle_training_matches_reference_golden(name, enc, dec, fc, hw, stride):
    g = _gold(name)
    loss, acc, feats, out, e, d = _oracle_train(enc, dec, fc, 2, hw, stride)
    # same ATen kernels, same order: allow only thread-count dependent reduction-order noise
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert abs(acc.item() - float(g["acc"])) <= 1e-6
    pred = out[0] if isinstance(out, tuple) else out
    np.testing.assert_allclose(pred.detach().numpy(), g["pred"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose([f.mean().item() for f in feats], g["feat_mean"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(feats[-1].detach()[:, ::64, ::3, ::3].numpy(), g["feat3_sample"], rtol=1e-3, atol=1e-4)
    for key in [k for k in g.files if k.startswith("gradnorm:")]:
        pname = key[len("gradnorm:"):]
        sd = e if pname.startswith("enc.") else d
        gr = sd[pname[4:]].grad
        assert abs(gr.norm().item() - float(g[key])) <= 1e-3 * float(g[key]), pname
        ref = g["grad:" + pname]
        got = (gr if gr.numel() <= 4096 else gr.flatten()[:: max(1, gr.numel() // 4096)][:4096]).numpy()
        np.testing.assert_allclose(got, ref, rtol=2e-3, atol=1e-5 * float(g[key]))


@pytest.mark.parametrize("name,enc,dec,fc,n,h,w", [
    ("infer_r18dilated_ppm_deepsup_96x128", "resnet18dilated", "ppm_deepsup", 512, 2, 96, 128),
    ("infer_hrnetv2_c1_64x96", "hrnetv2", "c1", 720, 1, 64, 96),
    ("infer_mobilenetv2dilated_c1_deepsup_384", "mobilenetv2dilated", "c1_deepsup", 320, 1, 384, 384),   # BASELINE configs[0]
])
def test_oracle_inference_matches_reference_golden(name, enc, dec, fc, n, h, w):
    g = _gold(name)
    esd = O.synth_state_dict(O.encoder_param_shapes(enc), 304)
    dsd = O.synth_state_dict(O.decoder_param_shapes(dec, fc), 305)
    feed = O.synth_batch(n, h, w, 8, 2)
    with torch.no_grad():
        probs = O.segmentation_forward(feed, esd, dsd, enc, dec, O.BNState(False), None, segSize=(h, w))
    assert (probs.argmax(1).numpy().astype(np.uint8) == g["argmax"]).mean() > 0.9999
    np.testing.assert_allclose(probs[:, ::7, ::5, ::5].numpy(), g["probs_sample"], rtol=1e-4, atol=1e-6)


def test_oracle_dropout_semantics_match_reference():
    g = _gold("train_r18_dropout_seed11")
    esd = O.synth_state_dict(O.encoder_param_shapes("resnet18dilated"), 304)
    dsd = O.synth_state_dict(O.decoder_param_shapes("ppm_deepsup", 512), 305)
    feed = O.synth_batch(2, 64, 64, 8, 3)
    torch.manual_seed(11)
    with torch.no_grad():
        loss, acc = O.segmentation_forward(feed, esd, dsd, "resnet18dilated", "ppm_deepsup", O.BNState(True), 0.4)
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * float(g["loss"])
    # injected masks reproduce F.dropout2d's scaling
    x = torch.ones(2, 4, 3, 3)
    m = torch.tensor([[1, 0, 1, 1], [0, 1, 1, 0]])
    y = O._dropout2d(x, 0.1, True, m)
    assert torch.allclose(y[0, 0], torch.full((3, 3), 1 / 0.9)) and y[0, 1].abs().sum() == 0


def test_sync_batchnorm_formula_matches_reference_compute_mean_std():
    """Known-answer test from the reference's own _compute_mean_std (batchnorm.py:123-139), two consecutive updates."""
    g = _gold("syncbn_compute_mean_std")
    x = torch.from_numpy(g["x"])
    sd = {"bn.weight": torch.ones(16), "bn.bias": torch.zeros(16), "bn.running_mean": torch.zeros(16),
          "bn.running_var": torch.ones(16), "bn._tmp_running_mean": torch.zeros(16), "bn._tmp_running_var": torch.ones(16),
          "bn._running_iter": torch.ones(1)}
    st = O.BNState(True, sync=True, update_running=True)
    y = O.batch_norm(x, sd, "bn", st)
    mean, inv_std = torch.from_numpy(g["mean"]), torch.from_numpy(g["inv_std"])
    ref = (x - mean.view(1, -1, 1, 1)) * inv_std.view(1, -1, 1, 1)
    assert torch.allclose(y, ref, atol=1e-5)
    # numpy restatement agrees on mean / inv_std / unbiased variance bookkeeping
    _, m_np, is_np, unb = NP.batch_norm_train(g["x"], np.ones(16), np.zeros(16), sync_formula=True)
    np.testing.assert_allclose(m_np, g["mean"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(is_np, g["inv_std"], rtol=1e-5)
    # the reference then ran a SECOND update with scaled sums; its buffers after both updates are in the fixture
    frac = 1 - 0.001
    xv = x.view(6, 16, -1)
    s, ss, size = xv.sum(0).sum(-1) * 0.5, (xv ** 2).sum(0).sum(-1) * 0.7, 6 * 35
    mean2 = s / size
    unb2 = (ss - s * mean2) / (size - 1)
    tmp_mean = sd["bn._tmp_running_mean"] * frac + mean2
    tmp_var = sd["bn._tmp_running_var"] * frac + unb2
    it = sd["bn._running_iter"] * frac + 1
    np.testing.assert_allclose(tmp_mean.numpy(), g["tmp_mean"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(tmp_var.numpy(), g["tmp_var"], rtol=1e-5)
    np.testing.assert_allclose(it.numpy(), g["running_iter"], rtol=1e-6)
    np.testing.assert_allclose((tmp_mean / it).numpy(), g["running_mean"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose((tmp_var / it).numpy(), g["running_var"], rtol=1e-5)


def test_sync_bn_equals_plain_bn_on_pooled_batch():
    """Structure of the reference's test_sync_batchnorm.py:44-65: pooled statistics == nn.BatchNorm2d on the full batch
    (output and input gradient, atol 1e-3 like TorchTestCase.assertTensorClose)."""
    torch.manual_seed(0)
    x = torch.rand(16, 10, 16, 16)
    sd = {"bn.weight": torch.ones(10), "bn.bias": torch.zeros(10), "bn.running_mean": torch.zeros(10),
          "bn.running_var": torch.ones(10), "bn._tmp_running_mean": torch.zeros(10), "bn._tmp_running_var": torch.ones(10),
          "bn._running_iter": torch.ones(1)}
    xa = x.clone().requires_grad_(True)
    ya = O.batch_norm(xa, sd, "bn", O.BNState(True, sync=True))
    ya.sum().backward()
    ref = nn.BatchNorm2d(10)
    xb = x.clone().requires_grad_(True)
    yb = ref(xb)
    yb.sum().backward()
    assert torch.allclose(ya, yb, atol=1e-3) and torch.allclose(xa.grad, xb.grad, atol=1e-3)


# ------------------------------------------------------------------------------------------- numpy primitives
def test_numpy_primitives_against_torch():
    rng = np.random.RandomState(0)
    x = rng.randn(2, 5, 13, 11).astype(np.float32)
    w = rng.randn(7, 5, 3, 3).astype(np.float32)
    b = rng.randn(7).astype(np.float32)
    for stride, dil in ((1, 1), (1, 2), (1, 4), (2, 1)):
        ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), stride, dil, dil).numpy()
        np.testing.assert_allclose(NP.conv2d(x, w, b, stride, dil, dil), ref, rtol=1e-4, atol=1e-4)
    g, be = rng.rand(5).astype(np.float32) + 0.5, rng.randn(5).astype(np.float32)
    ref = F.batch_norm(torch.from_numpy(x), None, None, torch.from_numpy(g), torch.from_numpy(be), True, 0.1, 1e-5).numpy()
    np.testing.assert_allclose(NP.batch_norm_train(x, g, be)[0], ref, rtol=1e-4, atol=1e-4)
    np.testing.assert_array_equal(NP.max_pool_3x3_s2(x), F.max_pool2d(torch.from_numpy(x), 3, 2, 1).numpy())
    big = rng.randn(1, 3, 64, 64).astype(np.float32)
    for s in (1, 2, 3, 6):
        np.testing.assert_allclose(NP.adaptive_avg_pool(big, s), F.adaptive_avg_pool2d(torch.from_numpy(big), s).numpy(),
                                   rtol=1e-4, atol=1e-5)
        small = NP.adaptive_avg_pool(big, s)
        np.testing.assert_allclose(NP.bilinear(small, 64, 64),
                                   F.interpolate(torch.from_numpy(small), (64, 64), mode="bilinear",
                                                 align_corners=False).numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(NP.bilinear(x, 29, 40), F.interpolate(torch.from_numpy(x), (29, 40), mode="bilinear",
                                                                    align_corners=False).numpy(), rtol=1e-4, atol=1e-5)
    logits = rng.randn(2, 9, 6, 5).astype(np.float32) * 3
    label = rng.randint(-1, 9, (2, 6, 5))
    loss, acc = NP.log_softmax_nll(logits, label)
    ref = F.nll_loss(F.log_softmax(torch.from_numpy(logits), 1), torch.from_numpy(label), ignore_index=-1)
    assert abs(loss - ref.item()) < 1e-5
    assert abs(acc - O.pixel_acc(torch.from_numpy(logits), torch.from_numpy(label)).item()) < 1e-6


def test_adaptive_pool_bins_overlap_as_surveyed():
    """64 -> 3 bins are [0,22),[21,43),[42,64)."""
    bins = [((i * 64) // 3, -((-(i + 1) * 64) // 3)) for i in range(3)]
    assert bins == [(0, 22), (21, 43), (42, 64)]


# ------------------------------------------------------------------------------------------- host-side API mirror
API = json.load(open(os.path.join(GOLD, "reference_api.json")))


def _build_mine(enc_arch, dec_arch, fc):
    from mit_semseg.models import ModelBuilder
    from mit_semseg.models import models as M, resnet as R
    base, dil = O.parse_encoder_arch(enc_arch)
    if base == "hrnetv2":
        from mit_semseg.models import hrnet as HR
        enc = HR.hrnetv2(pretrained=False)
    elif base == "mobilenetv2":
        from mit_semseg.models import mobilenet as MB
        enc = M.MobileNetV2Dilated(MB.mobilenetv2(pretrained=False), dilate_scale=8)
    else:
        net = R.__dict__[base](pretrained=False)
        enc = M.ResnetDilated(net, 8) if dil else M.Resnet(net)
    dec = ModelBuilder.build_decoder(dec_arch, fc_dim=fc, num_class=150)
    return enc, dec


@pytest.mark.parametrize("combo,fc", [("resnet50dilated+ppm_deepsup", 2048), ("resnet18dilated+ppm_deepsup", 512),
                                      ("resnet101+c1_deepsup", 2048), ("resnet50+ppm", 2048), ("resnet18+c1", 512),
                                      ("resnet50+upernet", 2048), ("resnet18+upernet_lite", 512),
                                      ("hrnetv2+c1", 720), ("mobilenetv2dilated+c1_deepsup", 320)])
def test_module_tree_matches_reference_state_dict_init_and_hparams(combo, fc):
    enc_arch, dec_arch = combo.split("+")
    ref = API[combo]
    torch.manual_seed(304)
    enc, dec = _build_mine(enc_arch, dec_arch, fc)
    assert {k: list(v.shape) for k, v in enc.state_dict().items()} == ref["enc_keys"]
    assert {k: list(v.shape) for k, v in dec.state_dict().items()} == ref["dec_keys"]
    # same constructor order + same initialisers => same RNG stream => identical weights under the same seed
    for k, (s, a) in ref["enc_init"].items():
        v = enc.state_dict()[k].double()
        assert abs(v.sum().item() - s) <= 1e-6 * max(1.0, abs(a)) and abs(v.abs().sum().item() - a) <= 1e-6 * max(1.0, a), k
    for k, (s, a) in ref["dec_init"].items():
        v = dec.state_dict()[k].double()
        assert abs(v.abs().sum().item() - a) <= 1e-6 * max(1.0, a), k
    hp = {k: [list(m.stride), list(m.dilation), list(m.padding)] + ([m.groups] if m.groups != 1 else [])
          for k, m in enc.named_modules() if isinstance(m, nn.Conv2d)}
    assert hp == ref["conv_hparams"]
    # oracle's own table of hyper-parameters (used by encoder_forward) agrees as well
    base, dil = O.parse_encoder_arch(enc_arch)
    # the oracle's parameter table names exactly the reference's parameters
    shapes = O.synth_state_dict(O.encoder_param_shapes(enc_arch), 1)
    assert {k: list(v.shape) for k, v in shapes.items()} == ref["enc_keys"]
    if base in ("hrnetv2", "mobilenetv2"):
        return
    block, counts = O.RESNET_LAYERS[base]
    for li, nb in enumerate(counts, start=1):
        for bi in range(nb):
            for cname in ("conv1", "conv2", "downsample.0"):
                key = "layer%d.%d.%s" % (li, bi, cname)
                if key in ref["conv_hparams"]:
                    s, d, p = O._conv_hparams(block, li, bi, cname, dil)
                    assert ref["conv_hparams"][key] == [[s, s], [d, d], [p, p]], key


def test_builder_contract():
    from mit_semseg.models import ModelBuilder, SegmentationModule
    with pytest.raises(Exception, match="Architecture undefined"):
        ModelBuilder.build_encoder("nope", weights="x")
    with pytest.raises(Exception, match="Architecture undefined"):
        ModelBuilder.build_decoder("nope")
    with pytest.raises(NotImplementedError):
        ModelBuilder.build_encoder("resnet34", weights="x")
    dec = ModelBuilder.build_decoder("ppm_deepsup", fc_dim=2048, num_class=150)
    # weights_init: conv kaiming, BN weight 1 / bias 1e-4 (reference models.py:52-59)
    assert float(dec.conv_last[1].weight.min()) == 1.0 and abs(float(dec.conv_last[1].bias[0]) - 1e-4) < 1e-9
    enc = ModelBuilder.build_encoder  # pretrained download path must fail loudly offline, never silently random-init
    with pytest.raises(FileNotFoundError):
        enc("resnet18dilated", weights="")
    seg = SegmentationModule(nn.Identity(), dec, nn.NLLLoss(ignore_index=-1), 0.4)
    assert seg.deep_sup_scale == 0.4 and seg.decoder is dec


def test_train_py_group_weight_contract():
    """train.py:92-112 asserts every parameter lives in a Conv / BatchNorm / Linear module."""
    enc, dec = _build_mine("resnet50dilated", "ppm_deepsup", 2048)
    for module in (enc, dec):
        n = 0
        for m in module.modules():
            if isinstance(m, (nn.Linear, nn.modules.conv._ConvNd)):
                n += 1 + (m.bias is not None)
            elif isinstance(m, nn.modules.batchnorm._BatchNorm):
                n += (m.weight is not None) + (m.bias is not None)
        assert n == len(list(module.parameters()))


def test_checkpoint_round_trip_through_the_builders(tmp_path):
    """train.py:74-89 saves `encoder_epoch_N.pth` / `decoder_epoch_N.pth` as raw state dicts; eval.py / test.py hand the
    paths to build_encoder / build_decoder (models.py:104-108,151-155: torch.load to CPU, load_state_dict strict=False).
    The synthetic state dicts carry exactly the reference's keys (checked against reference_api.json above), so this is
    the reference's checkpoint format."""
    from mit_semseg.models import ModelBuilder
    for enc_arch, dec_arch, fc in (("resnet18dilated", "ppm_deepsup", 512), ("hrnetv2", "c1", 720)):
        esd = O.synth_state_dict(O.encoder_param_shapes(enc_arch), 7)
        dsd = O.synth_state_dict(O.decoder_param_shapes(dec_arch, fc), 8)
        pe, pd = str(tmp_path / ("encoder_%s.pth" % enc_arch)), str(tmp_path / ("decoder_%s.pth" % dec_arch))
        torch.save(esd, pe)
        torch.save(dsd, pd)
        enc = ModelBuilder.build_encoder(enc_arch, fc_dim=fc, weights=pe)
        dec = ModelBuilder.build_decoder(dec_arch, fc_dim=fc, num_class=150, weights=pd, use_softmax=True)
        for net, sd in ((enc, esd), (dec, dsd)):
            got = net.state_dict()
            assert set(got) == set(sd)
            assert all(torch.equal(got[k], sd[k]) for k in sd)
        # and back: what train.py's checkpoint() writes loads into a fresh net bit for bit
        torch.save(enc.state_dict(), pe)
        enc2 = ModelBuilder.build_encoder(enc_arch, fc_dim=fc, weights=pe)
        assert all(torch.equal(v, enc2.state_dict()[k]) for k, v in enc.state_dict().items())
        assert dec.use_softmax
