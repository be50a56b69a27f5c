"""Generates tests/golden/* by running the UNMODIFIED reference (imported read-only from /root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):   python oracle/make_golden.py
It also asserts, while generating, that oracle/segnet_oracle.py reproduces every reference output bit-for-bit
(both sides are torch-CPU fp32 calling the same ATen kernels in the same order), so the fixtures pin the oracle.

Work-arounds needed to import the reference here: `collections.Mapping/Sequence` aliases for py>=3.10;
encoders are built with pretrained=False (the builders would download ImageNet weights otherwise).
"""
import collections
import collections.abc
import json
import os
import sys

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True
collections.Mapping = collections.abc.Mapping
collections.Sequence = collections.abc.Sequence
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from mit_semseg.lib.nn import SynchronizedBatchNorm2d  # noqa: E402  (reference)
from mit_semseg.models import ModelBuilder, SegmentationModule  # noqa: E402  (reference)
from mit_semseg.models import models as rmodels, resnet as rresnet, hrnet as rhrnet, mobilenet as rmobilenet  # noqa: E402

from oracle import segnet_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
assert "/root/reference" in rmodels.__file__, "must import the REFERENCE mit_semseg here"


def build_ref(enc_arch, dec_arch, fc_dim, use_softmax=False):
    base, dil = O.parse_encoder_arch(enc_arch)
    if base == "hrnetv2":
        enc = rhrnet.hrnetv2(pretrained=False)
    elif base == "mobilenetv2":
        enc = rmodels.MobileNetV2Dilated(rmobilenet.mobilenetv2(pretrained=False), dilate_scale=8)
    else:
        net = rresnet.__dict__[base](pretrained=False)
        enc = rmodels.ResnetDilated(net, 8) if dil else rmodels.Resnet(net)
    dec = ModelBuilder.build_decoder(dec_arch, fc_dim=fc_dim, num_class=150, use_softmax=use_softmax)
    return enc, dec


def grads_of(module):
    return {k: p.grad.clone() for k, p in module.named_parameters()}


def train_case(name, enc_arch, dec_arch, fc, n, hw, label_stride, keep_grads):
    enc, dec = build_ref(enc_arch, dec_arch, fc)
    esd = O.synth_state_dict(O.encoder_param_shapes(enc_arch), 304)
    dsd = O.synth_state_dict(O.decoder_param_shapes(dec_arch, fc), 305)
    enc.load_state_dict(esd)
    dec.load_state_dict(dsd)
    ds = 0.4 if dec_arch.endswith("deepsup") else None
    seg = SegmentationModule(enc, dec, nn.NLLLoss(ignore_index=-1), ds)
    seg.train()
    for m in seg.modules():
        if isinstance(m, nn.Dropout2d):
            m.p = 0.0
    feed = O.synth_batch(n, hw, hw, label_stride, 1)
    feats = enc(feed["img_data"], return_feature_maps=True)
    out = dec(feats)
    # reload: the forward above moved the running statistics
    enc.load_state_dict(esd)
    dec.load_state_dict(dsd)
    loss, acc = seg(feed)
    loss.backward()
    ge, gd = grads_of(enc), grads_of(dec)
    # ---- the oracle must agree exactly
    e = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in esd.items()}
    d = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in dsd.items()}
    lo, ao, fo, oo = O.segmentation_forward(feed, e, d, enc_arch, dec_arch, O.BNState(True), ds, dropout_p=0.0,
                                            return_aux=True)
    lo.backward()
    assert lo.item() == loss.item() and ao.item() == acc.item(), (name, lo.item(), loss.item())
    for a, b in zip(fo, feats):
        assert torch.equal(a, b)
    for k, g in ge.items():
        assert torch.equal(e[k].grad, g), k
    for k, g in gd.items():
        assert torch.equal(d[k].grad, g), k
    pred = out[0] if isinstance(out, tuple) else out
    rec = {"loss": np.float32(loss.item()), "acc": np.float32(acc.item()), "pred": pred.detach().numpy(),
           "feat_mean": np.array([f.mean().item() for f in feats], np.float32),
           "feat_absmean": np.array([f.abs().mean().item() for f in feats], np.float32),
           "feat3_sample": feats[-1].detach()[:, ::64, ::3, ::3].numpy()}
    if isinstance(out, tuple):
        rec["pred_deepsup_sample"] = out[1].detach()[:, ::5].numpy()
    for k in keep_grads:
        src = ge if k.startswith("enc.") else gd
        g = src[k[4:]]
        rec["grad:" + k] = (g if g.numel() <= 4096 else g.flatten()[:: max(1, g.numel() // 4096)][:4096]).numpy()
        rec["gradnorm:" + k] = np.float32(g.norm().item())
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(name, "loss", loss.item(), "acc", acc.item())


def infer_case(name, enc_arch, dec_arch, fc, n, h, w):
    enc, dec = build_ref(enc_arch, dec_arch, fc, use_softmax=True)
    esd = O.synth_state_dict(O.encoder_param_shapes(enc_arch), 304)
    dsd = O.synth_state_dict(O.decoder_param_shapes(dec_arch, fc), 305)
    enc.load_state_dict(esd)
    dec.load_state_dict(dsd)
    seg = SegmentationModule(enc, dec, nn.NLLLoss(ignore_index=-1))
    seg.eval()
    feed = O.synth_batch(n, h, w, 8, 2)
    with torch.no_grad():
        probs = seg(feed, segSize=(h, w))
        po = O.segmentation_forward(feed, esd, dsd, enc_arch, dec_arch, O.BNState(False), None, segSize=(h, w))
    assert torch.equal(probs, po)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), argmax=probs.argmax(1).numpy().astype(np.uint8),
                        probs_sample=probs[:, ::7, ::5, ::5].numpy(), maxprob_mean=np.float32(probs.max(1)[0].mean().item()))
    print(name, "ok")


def syncbn_case():
    """_compute_mean_std (reference batchnorm.py:123-139) is a pure function of (sum, ssum, size) + buffers: call it."""
    torch.manual_seed(0)
    bn = SynchronizedBatchNorm2d(16)
    x = torch.randn(6, 16, 5, 7) * 2 + 1
    xv = x.view(6, 16, -1)
    s, ss, size = xv.sum(0).sum(-1), (xv ** 2).sum(0).sum(-1), 6 * 35
    mean, inv_std = bn._compute_mean_std(s, ss, size)
    mean2, inv_std2 = bn._compute_mean_std(s * 0.5, ss * 0.7, size)  # second update: accumulator state matters
    sd = {"bn.weight": bn.weight.detach().clone(), "bn.bias": bn.bias.detach().clone(), "bn.running_mean": torch.zeros(16),
          "bn.running_var": torch.ones(16), "bn._tmp_running_mean": torch.zeros(16), "bn._tmp_running_var": torch.ones(16),
          "bn._running_iter": torch.ones(1)}
    st = O.BNState(True, sync=True, update_running=True)
    y = O.batch_norm(x, sd, "bn", st)
    y_ref = (xv - mean.view(1, -1, 1)) * (inv_std * bn.weight).view(1, -1, 1) + bn.bias.view(1, -1, 1)
    assert torch.allclose(y, y_ref.view_as(x).detach(), atol=1e-6)
    np.savez_compressed(os.path.join(GOLD, "syncbn_compute_mean_std.npz"), x=x.numpy(), mean=mean.numpy(),
                        inv_std=inv_std.numpy(), running_mean=bn.running_mean.numpy(), running_var=bn.running_var.numpy(),
                        tmp_mean=bn._tmp_running_mean.numpy(), tmp_var=bn._tmp_running_var.numpy(),
                        running_iter=bn._running_iter.numpy(), mean2=mean2.numpy(), inv_std2=inv_std2.numpy())
    # the reference's own unit test structure (test_sync_batchnorm.py:44-65): vs nn.BatchNorm2d, matching momentum
    bn2 = SynchronizedBatchNorm2d(10, momentum=0.1)
    ref = nn.BatchNorm2d(10, momentum=0.1)
    xi = torch.rand(16, 10, 16, 16)
    a = bn2(xi.clone().requires_grad_(True))
    b = ref(xi.clone().requires_grad_(True))
    assert torch.allclose(a, b, atol=1e-5) and torch.allclose(bn2.running_var, ref.running_var, atol=1e-5)
    print("syncbn ok")


def dropout_case():
    """Dropout2d active: reference vs oracle under the same torch seed (both CPU) -> identical loss."""
    enc_arch, dec_arch, fc = "resnet18dilated", "ppm_deepsup", 512
    enc, dec = build_ref(enc_arch, dec_arch, fc)
    esd = O.synth_state_dict(O.encoder_param_shapes(enc_arch), 304)
    dsd = O.synth_state_dict(O.decoder_param_shapes(dec_arch, fc), 305)
    enc.load_state_dict(esd)
    dec.load_state_dict(dsd)
    seg = SegmentationModule(enc, dec, nn.NLLLoss(ignore_index=-1), 0.4)
    seg.train()
    feed = O.synth_batch(2, 64, 64, 8, 3)
    torch.manual_seed(11)
    loss, acc = seg(feed)
    torch.manual_seed(11)
    lo, ao = O.segmentation_forward(feed, dict(esd), dict(dsd), enc_arch, dec_arch, O.BNState(True), 0.4)
    assert lo.item() == loss.item(), (lo.item(), loss.item())
    np.savez_compressed(os.path.join(GOLD, "train_r18_dropout_seed11.npz"), loss=np.float32(loss.item()),
                        acc=np.float32(acc.item()))
    print("dropout ok", loss.item())


def api_case():
    """State-dict keys/shapes and seed-304 initialisation checksums of the reference builders."""
    out = {}
    for enc_arch, dec_arch, fc in (("resnet50dilated", "ppm_deepsup", 2048), ("resnet18dilated", "ppm_deepsup", 512),
                                   ("resnet101", "c1_deepsup", 2048), ("resnet50", "ppm", 2048), ("resnet18", "c1", 512),
                                   ("resnet50", "upernet", 2048), ("resnet18", "upernet_lite", 512),
                                   ("hrnetv2", "c1", 720), ("mobilenetv2dilated", "c1_deepsup", 320)):
        torch.manual_seed(304)
        enc, dec = build_ref(enc_arch, dec_arch, fc)
        rec = {"enc_keys": {k: list(v.shape) for k, v in enc.state_dict().items()},
               "dec_keys": {k: list(v.shape) for k, v in dec.state_dict().items()},
               "enc_init": {k: [float(v.double().sum()), float(v.double().abs().sum())]
                            for k, v in list(enc.state_dict().items())[:: 37] if v.is_floating_point()},
               "dec_init": {k: [float(v.double().sum()), float(v.double().abs().sum())]
                            for k, v in dec.state_dict().items() if v.is_floating_point() and v.dim() > 1},
               "conv_hparams": {k: [list(m.stride), list(m.dilation), list(m.padding)] + ([m.groups] if m.groups != 1 else [])
                                for k, m in enc.named_modules() if isinstance(m, nn.Conv2d)}}
        out["%s+%s" % (enc_arch, dec_arch)] = rec
    json.dump(out, open(os.path.join(GOLD, "reference_api.json"), "w"), indent=0, sort_keys=True)
    print("api ok")


if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "api":
        api_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mobilenet":
        # BASELINE configs[0]: MobileNetV2dilated + C1_deepsup, single-image forward at 384x384 (the reference's CPU case)
        infer_case("infer_mobilenetv2dilated_c1_deepsup_384", "mobilenetv2dilated", "c1_deepsup", 320, 1, 384, 384)
        train_case("train_mobilenetv2dilated_c1_deepsup_96", "mobilenetv2dilated", "c1_deepsup", 320, 2, 96, 8,
                   ["enc.features.0.0.weight", "enc.features.7.conv.3.weight", "enc.features.17.conv.6.weight",
                    "dec.cbr_deepsup.0.weight"])
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "hrnet":
        train_case("train_hrnetv2_c1_64", "hrnetv2", "c1", 720, 2, 64, 4,
                   ["enc.conv1.weight", "enc.stage2.0.fuse_layers.0.1.0.weight", "enc.stage4.2.fuse_layers.3.0.2.0.weight",
                    "enc.stage3.1.branches.2.3.conv2.weight", "enc.transition2.2.0.0.weight", "dec.cbr.0.weight",
                    "dec.conv_last.bias"])
        infer_case("infer_hrnetv2_c1_64x96", "hrnetv2", "c1", 720, 1, 64, 96)
        sys.exit(0)
    train_case("train_r50dilated_ppm_deepsup_96", "resnet50dilated", "ppm_deepsup", 2048, 2, 96, 8,
               ["enc.conv1.weight", "enc.bn1.weight", "enc.layer2.0.conv2.weight", "enc.layer3.1.conv2.weight",
                "enc.layer4.2.conv3.weight", "dec.conv_last.0.weight", "dec.conv_last.4.bias", "dec.ppm.0.2.weight",
                "dec.cbr_deepsup.0.weight"])
    train_case("train_r18dilated_c1_deepsup_96", "resnet18dilated", "c1_deepsup", 512, 2, 96, 8,
               ["enc.conv1.weight", "dec.cbr.0.weight"])
    train_case("train_r50_upernet_128", "resnet50", "upernet", 2048, 2, 128, 4, ["dec.conv_last.1.bias"])
    infer_case("infer_r18dilated_ppm_deepsup_96x128", "resnet18dilated", "ppm_deepsup", 512, 2, 96, 128)
    train_case("train_hrnetv2_c1_64", "hrnetv2", "c1", 720, 2, 64, 4,
               ["enc.conv1.weight", "enc.stage2.0.fuse_layers.0.1.0.weight", "enc.stage4.2.fuse_layers.3.0.2.0.weight",
                "enc.stage3.1.branches.2.3.conv2.weight", "enc.transition2.2.0.0.weight", "dec.cbr.0.weight",
                "dec.conv_last.bias"])
    infer_case("infer_hrnetv2_c1_64x96", "hrnetv2", "c1", 720, 1, 64, 96)
    infer_case("infer_mobilenetv2dilated_c1_deepsup_384", "mobilenetv2dilated", "c1_deepsup", 320, 1, 384, 384)
    train_case("train_mobilenetv2dilated_c1_deepsup_96", "mobilenetv2dilated", "c1_deepsup", 320, 2, 96, 8,
               ["enc.features.0.0.weight", "enc.features.7.conv.3.weight", "enc.features.17.conv.6.weight",
                "dec.cbr_deepsup.0.weight"])
    syncbn_case()
    dropout_case()
    api_case()
