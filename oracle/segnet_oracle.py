"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A CPU, fp32 restatement of the reference hot path (CSAILVision/semantic-segmentation-pytorch @ 8f27c9b):
deep-stem (dilated) ResNet, HRNetV2-W48 or MobileNetV2dilated -> PPM / PPM_deepsup / C1 / UPerNet decoder -> log-softmax / NLL / pixel accuracy, with both
batch-norm formulas of SynchronizedBatchNorm2d.  It is written functionally over a flat state dict that uses the
reference's parameter names, so one weights file loads into the reference, this oracle and the B200 engine.

Where the arithmetic lives: the reference has no kernels of its own; every primitive is a call into PyTorch ATen
(CPU: oneDNN) — third-party, pinned only as `torch>=0.4.1` (reference setup.py:22); this container has torch 2.11.0.
The primitives (conv2d, batch_norm, max_pool2d, adaptive_avg_pool2d, bilinear interpolate, log_softmax, nll_loss)
are therefore the SAME library calls the reference makes, composed exactly as the reference composes them; their
published algorithms are restated independently in numpy in `oracle/np_primitives.py` and cross-checked in
tests/test_oracle.py.  Backward is torch autograd through these ops, as in the reference (batchnorm.py has no
custom backward).

Pinning: the reference ships no golden vectors for this path ("parity unpinned" by its tests), so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF imported from /root/reference in the build container:
`oracle/make_golden.py` generated tests/golden/*.npz; tests/test_oracle.py checks this file against them.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# architecture tables (reference: models/resnet.py:165-203 layer counts; models/models.py:63-110 arch names)
RESNET_LAYERS = {"resnet18": ("basic", [2, 2, 2, 2]), "resnet50": ("bottleneck", [3, 4, 6, 3]),
                 "resnet101": ("bottleneck", [3, 4, 23, 3])}


# HRNetV2-W48 (reference hrnet.py:257-262, :271): (modules, branch widths) per stage, BasicBlocks per branch, Bottlenecks
# in layer1. Module-level so a test can also run a shallower net of the same topology (the reference hard-codes these).
HRNET_STAGES = ((1, (48, 96)), (4, (48, 96, 192)), (3, (48, 96, 192, 384)))
HRNET_BRANCH_BLOCKS = 4
HRNET_LAYER1_BLOCKS = 4


# MobileNetV2 (reference models/mobilenet.py:86-95): (expansion t, channels c, repeats n, stride s) per group
MOBILENET_SETTING = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1))
MOBILENET_DOWN_IDX = (2, 4, 7, 14)   # models/models.py:280


def parse_encoder_arch(arch):
    arch = arch.lower()
    if arch == "hrnetv2":
        return "hrnetv2", False
    if arch == "mobilenetv2dilated":
        return "mobilenetv2", True
    dilated = arch.endswith("dilated")
    base = arch[:-len("dilated")] if dilated else arch
    if base not in RESNET_LAYERS:
        raise Exception("Architecture undefined!")
    return base, dilated


def _conv_hparams(block, layer_idx, block_idx, conv_name, dilated):
    """stride/dilation/padding of one ResNet conv after ResnetDilated._nostride_dilate
    (reference models/models.py:238-251 applied with dilate=2 to layer3 and dilate=4 to layer4)."""
    # original stride: first block of layers 2..4 carries stride 2 on conv1 (basic) / conv2 (bottleneck) / downsample
    strided_name = "conv1" if block == "basic" else "conv2"
    stride = 2 if (layer_idx >= 2 and block_idx == 0 and conv_name in (strided_name, "downsample.0")) else 1
    is3x3 = (conv_name in ("conv1", "conv2")) if block == "basic" else (conv_name == "conv2")
    dilation, padding = 1, (1 if is3x3 else 0)
    if dilated and layer_idx in (3, 4):
        dilate = 2 if layer_idx == 3 else 4
        if stride == 2:
            stride = 1
            if is3x3:
                dilation = padding = dilate // 2
        elif is3x3:
            dilation = padding = dilate
    return stride, dilation, padding


# ----------------------------------------------------------------------------------------------
# optional storage-precision emulation: the engine keeps activations / activation-gradients / GEMM operands in bf16
# (fp32 accumulation). `BNState(emulate="bf16")` rounds at the same points (conv output, applied block output, pooled /
# resized maps, weights; and the corresponding gradients in backward), so engine-vs-oracle comparisons isolate kernel
# errors from the (large, chaotic at random init) sensitivity of a 50-layer BN-ReLU net to bf16 rounding.
class _RoundFB(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _RoundF(torch.autograd.Function):
    """round in forward only (weights: the fp32 master receives the full-precision gradient)"""
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundB(torch.autograd.Function):
    """identity in forward, round the gradient (the engine stores that gradient tensor in bf16)"""
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _qb(x, st):
    return _RoundB.apply(x) if getattr(st, "emulate", None) == "bf16" else x


def _q(x, st):
    return _RoundFB.apply(x) if getattr(st, "emulate", None) == "bf16" else x


def _qw(w, st):
    return _RoundF.apply(w) if getattr(st, "emulate", None) == "bf16" else w


# ----------------------------------------------------------------------------------------------
# batch norm: the two formulas of _SynchronizedBatchNorm.forward (reference lib/nn/modules/batchnorm.py:56-86)
class BNState:
    """How batch norm behaves for this forward pass.

    training=False               -> running statistics (F.batch_norm eval; batchnorm.py:58-61)
    training=True,  sync=False   -> F.batch_norm training: biased var + eps, momentum update (batchnorm.py:58-61)
    training=True,  sync=True    -> the data-parallel branch: statistics pooled over ALL replicas' batches,
                                    inv_std = clamp(var, eps)^-0.5 (batchnorm.py:123-139). The oracle is given the
                                    concatenation of every replica's batch, which is the same pooling.
    """

    def __init__(self, training, sync=False, eps=1e-5, momentum=0.001, update_running=False, emulate=None):
        self.training, self.sync, self.eps, self.momentum, self.update_running = training, sync, eps, momentum, update_running
        self.emulate = emulate  # None (fp32, the reference) | "bf16" (engine storage precision, see _RoundFB)


def batch_norm(x, sd, prefix, st):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if not st.training:
        return F.batch_norm(x, rm, rv, w, b, False, st.momentum, st.eps)
    if not st.sync:
        if st.update_running:
            return F.batch_norm(x, rm, rv, w, b, True, st.momentum, st.eps)
        return F.batch_norm(x, None, None, w, b, True, st.momentum, st.eps)
    # batchnorm.py:63-81 + _compute_mean_std :123-139
    shape = x.shape
    xv = x.reshape(x.size(0), x.size(1), -1)
    size = xv.size(0) * xv.size(2)
    s = xv.sum(dim=0).sum(dim=-1)
    ss = (xv ** 2).sum(dim=0).sum(dim=-1)
    assert size > 1
    mean = s / size
    sumvar = ss - s * mean
    bias_var = sumvar / size
    if st.update_running:
        unbias_var = sumvar / (size - 1)
        frac = 1.0 - st.momentum
        sd[prefix + "._tmp_running_mean"] = sd[prefix + "._tmp_running_mean"] * frac + mean.detach()
        sd[prefix + "._tmp_running_var"] = sd[prefix + "._tmp_running_var"] * frac + unbias_var.detach()
        sd[prefix + "._running_iter"] = sd[prefix + "._running_iter"] * frac + 1
        sd[prefix + ".running_mean"] = sd[prefix + "._tmp_running_mean"] / sd[prefix + "._running_iter"]
        sd[prefix + ".running_var"] = sd[prefix + "._tmp_running_var"] / sd[prefix + "._running_iter"]
    inv_std = bias_var.clamp(st.eps) ** -0.5
    out = (xv - mean.view(1, -1, 1)) * (inv_std * w).view(1, -1, 1) + b.view(1, -1, 1)
    return out.view(shape)


def _cbr(x, sd, conv, bn, st, stride=1, dilation=1, padding=0, relu=True):
    """conv -> BN -> (ReLU). With relu=False the caller finishes the block (residual add + ReLU) and rounds there."""
    is_img = x.shape[1] == 3
    w = sd[conv + ".weight"] if is_img else _qw(sd[conv + ".weight"], st)  # the stem kernel reads fp32 weights
    x = _q(F.conv2d(x, w, sd.get(conv + ".bias"), stride, padding, dilation), st)
    x = batch_norm(x, sd, bn, st)
    return _q(F.relu(x), st) if relu else x


# ----------------------------------------------------------------------------------------------
# encoder (reference models/resnet.py:24-92 blocks, :96-160 ResNet; models/models.py:170-268 wrappers)
def _residual_block(x, sd, p, st, kind, stride=1):
    """BasicBlock / Bottleneck with stride-1 3x3s and 'same' padding (the HRNet blocks, hrnet.py:34-102)."""
    residual = x
    if kind == "basic":
        out = _cbr(x, sd, p + "conv1", p + "bn1", st, stride, 1, 1)
        out = _cbr(out, sd, p + "conv2", p + "bn2", st, 1, 1, 1, relu=False)
    else:
        out = _cbr(x, sd, p + "conv1", p + "bn1", st)
        out = _cbr(out, sd, p + "conv2", p + "bn2", st, stride, 1, 1)
        out = _cbr(out, sd, p + "conv3", p + "bn3", st, relu=False)
    if (p + "downsample.0.weight") in sd:
        residual = _cbr(x, sd, p + "downsample.0", p + "downsample.1", st, stride, relu=False)
    return _q(F.relu(out + residual), st)


def hrnet_forward(x, sd, st, prefix=""):
    """HRNetV2.forward (reference models/hrnet.py:395-437) with HighResolutionModule.forward (:225-250) inlined.
    Rounding emulation follows the engine: an exchange output is ONE fused kernel (sum of the branch tensor, affine
    BN terms and bilinearly sampled affine BN terms, then ReLU), so only its result is rounded."""
    P = prefix
    x = _cbr(x, sd, P + "conv1", P + "bn1", st, stride=2, padding=1)
    x = _cbr(x, sd, P + "conv2", P + "bn2", st, stride=2, padding=1)
    for b in range(HRNET_LAYER1_BLOCKS):
        x = _residual_block(x, sd, "%slayer1.%d." % (P, b), st, "bottleneck")
    ys = [x]
    for si, (nmod, widths) in enumerate(HRNET_STAGES, start=2):
        # transition (hrnet.py:307-341, :405-432): new branches hang off the LAST previous output
        tr = "%stransition%d." % (P, si - 1)
        xs = []
        for i in range(len(widths)):
            if i < len(ys):
                if (tr + "%d.0.weight" % i) in sd:
                    xs.append(_cbr(ys[-1] if si > 2 else ys[i], sd, tr + "%d.0" % i, tr + "%d.1" % i, st, 1, 1, 1))
                else:
                    xs.append(ys[i])
            else:
                t = ys[-1]
                for k in range(i + 1 - len(ys)):
                    t = _cbr(t, sd, tr + "%d.%d.0" % (i, k), tr + "%d.%d.1" % (i, k), st, 2, 1, 1)
                xs.append(t)
        for m in range(nmod):
            mp = "%sstage%d.%d." % (P, si, m)
            for i in range(len(widths)):
                for b in range(HRNET_BRANCH_BLOCKS):
                    xs[i] = _residual_block(xs[i], sd, "%sbranches.%d.%d." % (mp, i, b), st, "basic")
            fused = []
            for i in range(len(widths)):
                size = xs[i].shape[2:]
                y = None
                for j in range(len(widths)):
                    fp = "%sfuse_layers.%d.%d." % (mp, i, j)
                    if j == i:
                        t = xs[j]
                    elif j > i:
                        t = _cbr(xs[j], sd, fp + "0", fp + "1", st, relu=False)
                        t = F.interpolate(_qb(t, st), size=size, mode="bilinear", align_corners=False)
                    else:
                        t = xs[j]
                        for k in range(i - j):
                            t = _cbr(t, sd, fp + "%d.0" % k, fp + "%d.1" % k, st, 2, 1, 1, relu=(k != i - j - 1))
                    y = t if y is None else y + t
                fused.append(_q(F.relu(y), st))
            xs = fused
        ys = xs
    size = ys[0].shape[2:]
    ups = [ys[0]] + [_q(F.interpolate(t, size=size, mode="bilinear", align_corners=False), st) for t in ys[1:]]
    return [torch.cat(ups, 1)]


def _mobilenet_blocks():
    """[(index in `features`, inp, oup, stride, expand_ratio)] of the InvertedResidual blocks (mobilenet.py:104-112)."""
    out, inp, idx = [], 32, 1
    for t, c, n, s in MOBILENET_SETTING:
        for i in range(n):
            out.append((idx, inp, c, s if i == 0 else 1, t))
            inp, idx = c, idx + 1
    return out


def _mobilenet_dilate(idx, stride, is3x3):
    """MobileNetV2Dilated._nostride_dilate with dilate_scale 8 (models/models.py:282-310): features[7:14] dilate 2,
    features[14:] dilate 4 -> (stride, dilation = padding) of a 3x3 conv, stride of a 1x1 conv."""
    dilate = 2 if 7 <= idx < 14 else (4 if idx >= 14 else 0)
    dil = 1
    if dilate:
        if stride == 2:
            stride = 1
            if is3x3:
                dil = dilate // 2
        elif is3x3:
            dil = dilate
    return stride, dil


def mobilenet_forward(x, sd, st, prefix="", return_feature_maps=True):
    """MobileNetV2Dilated.forward (reference models/models.py:312-323) over mobilenet.py's InvertedResidual blocks
    (:38-76): ReLU6 activations, depthwise 3x3 convolutions (groups = channels)."""
    P = prefix

    def cbr6(x, conv, bn, stride=1, dil=1, groups=1, k=3, act=True, add=None):
        # storage-precision emulation follows the engine's inference kernels: GEMM weights in bf16 (the depthwise / first
        # layer kernels read fp32 weights), ONE rounding per layer - after BN, shortcut and ReLU6, which ride in the epilogue
        pad = dil if k == 3 else 0
        w = sd[conv + ".weight"]
        if groups == 1 and x.shape[1] != 3:
            w = _qw(w, st)
        y = F.conv2d(x, w, None, stride, pad, dil, groups)
        y = batch_norm(y, sd, bn, st)
        if add is not None:
            y = y + add
        return _q(F.relu6(y) if act else y, st)
    x = cbr6(x, P + "features.0.0", P + "features.0.1", stride=2)
    outs = []
    for idx, inp, oup, stride, t in _mobilenet_blocks():
        p = "%sfeatures.%d.conv." % (P, idx)
        hidden = round(inp * t)
        s3, d3 = _mobilenet_dilate(idx, stride, True)
        y = x
        k = 0
        if t != 1:
            y = cbr6(y, p + "0", p + "1", k=1)
            k = 3
        y = cbr6(y, p + "%d" % k, p + "%d" % (k + 1), stride=s3, dil=d3, groups=hidden)
        x = cbr6(y, p + "%d" % (k + 3), p + "%d" % (k + 4), k=1, act=False, add=x if (stride == 1 and inp == oup) else None)
        if idx in MOBILENET_DOWN_IDX:
            outs.append(x)
    outs.append(x)
    return outs if return_feature_maps else [x]


def encoder_forward(x, sd, arch, st, prefix=""):
    base, dilated = parse_encoder_arch(arch)
    if base == "hrnetv2":
        return hrnet_forward(x, sd, st, prefix)
    if base == "mobilenetv2":
        return mobilenet_forward(x, sd, st, prefix)
    block, counts = RESNET_LAYERS[base]
    P = prefix
    x = _cbr(x, sd, P + "conv1", P + "bn1", st, stride=2, padding=1)
    x = _cbr(x, sd, P + "conv2", P + "bn2", st, padding=1)
    x = _cbr(x, sd, P + "conv3", P + "bn3", st, padding=1)
    x = F.max_pool2d(x, 3, 2, 1)  # max of bf16 values is exact
    outs = []
    for li, nblocks in enumerate(counts, start=1):
        for bi in range(nblocks):
            p = "%slayer%d.%d." % (P, li, bi)
            residual = x
            if block == "basic":
                s, d, pad = _conv_hparams(block, li, bi, "conv1", dilated)
                out = _cbr(x, sd, p + "conv1", p + "bn1", st, s, d, pad)
                s, d, pad = _conv_hparams(block, li, bi, "conv2", dilated)
                out = _cbr(out, sd, p + "conv2", p + "bn2", st, s, d, pad, relu=False)
            else:
                out = _cbr(x, sd, p + "conv1", p + "bn1", st)
                s, d, pad = _conv_hparams(block, li, bi, "conv2", dilated)
                out = _cbr(out, sd, p + "conv2", p + "bn2", st, s, d, pad)
                out = _cbr(out, sd, p + "conv3", p + "bn3", st, relu=False)
            if (p + "downsample.0.weight") in sd:
                s, _, _ = _conv_hparams(block, li, bi, "downsample.0", dilated)
                residual = _cbr(x, sd, p + "downsample.0", p + "downsample.1", st, s, relu=False)
            x = _q(F.relu(out + residual), st)
        outs.append(x)
    return outs


# ----------------------------------------------------------------------------------------------
# decoders (reference models/models.py:327-586)
def _dropout2d(x, p, training, mask=None):
    """nn.Dropout2d: one Bernoulli(1-p) draw per (n, c), survivors scaled by 1/(1-p). `mask` (N x C of 0/1) injects
    the draw so GPU and CPU runs can share it."""
    if not training or p == 0.0:
        return x
    if mask is None:
        return F.dropout2d(x, p, True)
    return x * (mask.to(x.dtype) / (1.0 - p)).view(x.size(0), x.size(1), 1, 1)


def _head(x, segSize, use_softmax):
    if use_softmax:
        x = F.interpolate(x, size=segSize, mode="bilinear", align_corners=False)
        return F.softmax(x, dim=1)
    return F.log_softmax(x, dim=1)


def decoder_forward(conv_out, sd, arch, st, segSize=None, use_softmax=False, dropout_p=0.1, masks=None, prefix="",
                    return_logits=False):
    arch = arch.lower()
    P = prefix
    masks = masks or {}
    conv5 = conv_out[-1]
    if arch in ("ppm", "ppm_deepsup"):
        H, W = conv5.shape[2:]
        ppm_out = [conv5]
        for i, scale in enumerate((1, 2, 3, 6)):
            y = _q(F.adaptive_avg_pool2d(conv5, scale), st)
            y = _cbr(y, sd, "%sppm.%d.1" % (P, i), "%sppm.%d.2" % (P, i), st)
            ppm_out.append(_q(F.interpolate(y, (H, W), mode="bilinear", align_corners=False), st))
        x = torch.cat(ppm_out, 1)
        x = _cbr(x, sd, P + "conv_last.0", P + "conv_last.1", st, padding=1)
        x = _dropout2d(x, dropout_p, st.training, masks.get("main"))
        logits = F.conv2d(x, _qw(sd[P + "conv_last.4.weight"], st), sd[P + "conv_last.4.bias"])
        if use_softmax:
            return _head(logits, segSize, True)
        if arch == "ppm":
            return logits if return_logits else F.log_softmax(logits, dim=1)
        y = _cbr(conv_out[-2], sd, P + "cbr_deepsup.0", P + "cbr_deepsup.1", st, padding=1)
        y = _dropout2d(y, dropout_p, st.training, masks.get("deepsup"))
        logits_ds = F.conv2d(y, _qw(sd[P + "conv_last_deepsup.weight"], st), sd[P + "conv_last_deepsup.bias"])
        if return_logits:
            return logits, logits_ds
        return F.log_softmax(logits, dim=1), F.log_softmax(logits_ds, dim=1)
    if arch in ("c1", "c1_deepsup"):
        x = _cbr(conv5, sd, P + "cbr.0", P + "cbr.1", st, padding=1)
        logits = F.conv2d(x, _qw(sd[P + "conv_last.weight"], st), sd[P + "conv_last.bias"])
        if use_softmax:
            return _head(logits, segSize, True)
        if arch == "c1":
            return logits if return_logits else F.log_softmax(logits, dim=1)
        y = _cbr(conv_out[-2], sd, P + "cbr_deepsup.0", P + "cbr_deepsup.1", st, padding=1)
        logits_ds = F.conv2d(y, _qw(sd[P + "conv_last_deepsup.weight"], st), sd[P + "conv_last_deepsup.bias"])
        if return_logits:
            return logits, logits_ds
        return F.log_softmax(logits, dim=1), F.log_softmax(logits_ds, dim=1)
    if arch in ("upernet", "upernet_lite"):
        H, W = conv5.shape[2:]
        ppm_out = [conv5]
        for i, scale in enumerate((1, 2, 3, 6)):
            y = _q(F.interpolate(_q(F.adaptive_avg_pool2d(conv5, scale), st), (H, W), mode="bilinear",
                                 align_corners=False), st)
            ppm_out.append(_cbr(y, sd, "%sppm_conv.%d.0" % (P, i), "%sppm_conv.%d.1" % (P, i), st))
        f = _cbr(torch.cat(ppm_out, 1), sd, P + "ppm_last_conv.0", P + "ppm_last_conv.1", st, padding=1)
        feats = [f]
        for i in reversed(range(len(conv_out) - 1)):
            lat = _cbr(conv_out[i], sd, "%sfpn_in.%d.0" % (P, i), "%sfpn_in.%d.1" % (P, i), st, relu=False)
            up = _q(F.interpolate(f, size=lat.shape[2:], mode="bilinear", align_corners=False), st)
            f = _q(F.relu(lat) + up, st)   # conv_x + f (models.py:563); the engine stores only the sum
            feats.append(_cbr(f, sd, "%sfpn_out.%d.0.0" % (P, i), "%sfpn_out.%d.0.1" % (P, i), st, padding=1))
        feats.reverse()
        size = feats[0].shape[2:]
        fusion = [feats[0]] + [_q(F.interpolate(t, size, mode="bilinear", align_corners=False), st) for t in feats[1:]]
        x = _cbr(torch.cat(fusion, 1), sd, P + "conv_last.0.0", P + "conv_last.0.1", st, padding=1)
        logits = F.conv2d(x, _qw(sd[P + "conv_last.1.weight"], st), sd[P + "conv_last.1.bias"])
        if use_softmax:
            return _head(logits, segSize, True)
        return logits if return_logits else F.log_softmax(logits, dim=1)
    raise Exception("Architecture undefined!")


# ----------------------------------------------------------------------------------------------
# SegmentationModule.forward (reference models/models.py:29-47) and pixel_acc (:12-18)
def pixel_acc(pred, label):
    _, preds = torch.max(pred, dim=1)
    valid = (label >= 0).long()
    acc_sum = torch.sum(valid * (preds == label).long())
    pixel_sum = torch.sum(valid)
    return acc_sum.float() / (pixel_sum.float() + 1e-10)


def segmentation_forward(feed, enc_sd, dec_sd, enc_arch, dec_arch, st, deep_sup_scale=None, segSize=None,
                         dropout_p=0.1, masks=None, return_aux=False):
    """Training branch (segSize None) -> (loss, acc); inference branch -> probabilities [N, C, *segSize]."""
    feats = encoder_forward(feed["img_data"], enc_sd, enc_arch, st)
    if segSize is not None:
        return decoder_forward(feats, dec_sd, dec_arch, st, segSize=segSize, use_softmax=True, dropout_p=dropout_p)
    out = decoder_forward(feats, dec_sd, dec_arch, st, dropout_p=dropout_p, masks=masks)
    label = feed["seg_label"]
    if deep_sup_scale is not None:
        pred, pred_ds = out
        loss = F.nll_loss(pred, label, ignore_index=-1) + F.nll_loss(pred_ds, label, ignore_index=-1) * deep_sup_scale
    else:
        pred = out
        loss = F.nll_loss(pred, label, ignore_index=-1)
    acc = pixel_acc(pred, label)
    if return_aux:
        return loss, acc, feats, out
    return loss, acc


# ----------------------------------------------------------------------------------------------
# deterministic synthetic weights shared by reference / oracle / engine (see oracle/make_golden.py)
def hrnet_param_shapes():
    shapes = {}

    def cb(name_conv, name_bn, co, ci, k):
        shapes[name_conv] = ("conv", (co, ci, k, k))
        shapes[name_bn] = ("bn", co)

    cb("conv1", "bn1", 64, 3, 3), cb("conv2", "bn2", 64, 64, 3)
    for b in range(HRNET_LAYER1_BLOCKS):
        p = "layer1.%d." % b
        cb(p + "conv1", p + "bn1", 64, 64 if b == 0 else 256, 1)
        cb(p + "conv2", p + "bn2", 64, 64, 3), cb(p + "conv3", p + "bn3", 256, 64, 1)
        if b == 0:
            cb(p + "downsample.0", p + "downsample.1", 256, 64, 1)
    pre = [256]
    for si, (nmod, widths) in enumerate(HRNET_STAGES, start=2):
        tr = "transition%d." % (si - 1)
        for i, c in enumerate(widths):
            if i < len(pre):
                if pre[i] != c:
                    cb(tr + "%d.0" % i, tr + "%d.1" % i, c, pre[i], 3)
            else:
                steps = i + 1 - len(pre)
                for k in range(steps):
                    cb(tr + "%d.%d.0" % (i, k), tr + "%d.%d.1" % (i, k), c if k == steps - 1 else pre[-1], pre[-1], 3)
        for m in range(nmod):
            mp = "stage%d.%d." % (si, m)
            for i, c in enumerate(widths):
                for b in range(HRNET_BRANCH_BLOCKS):
                    p = "%sbranches.%d.%d." % (mp, i, b)
                    cb(p + "conv1", p + "bn1", c, c, 3), cb(p + "conv2", p + "bn2", c, c, 3)
            for i, ci_ in enumerate(widths):
                for j, cj in enumerate(widths):
                    fp = "%sfuse_layers.%d.%d." % (mp, i, j)
                    if j > i:
                        cb(fp + "0", fp + "1", ci_, cj, 1)
                    elif j < i:
                        for k in range(i - j):
                            cb(fp + "%d.0" % k, fp + "%d.1" % k, ci_ if k == i - j - 1 else cj, cj, 3)
        pre = list(widths)
    return shapes


def mobilenet_param_shapes():
    shapes = {"features.0.0": ("conv", (32, 3, 3, 3)), "features.0.1": ("bn", 32)}
    for idx, inp, oup, stride, t in _mobilenet_blocks():
        p = "features.%d.conv." % idx
        hidden, k = round(inp * t), 0
        if t != 1:
            shapes[p + "0"], shapes[p + "1"] = ("conv", (hidden, inp, 1, 1)), ("bn", hidden)
            k = 3
        shapes[p + "%d" % k], shapes[p + "%d" % (k + 1)] = ("conv", (hidden, 1, 3, 3)), ("bn", hidden)   # depthwise
        shapes[p + "%d" % (k + 3)], shapes[p + "%d" % (k + 4)] = ("conv", (oup, hidden, 1, 1)), ("bn", oup)
    return shapes


def encoder_param_shapes(arch):
    base, _ = parse_encoder_arch(arch)
    if base == "hrnetv2":
        return hrnet_param_shapes()
    if base == "mobilenetv2":
        return mobilenet_param_shapes()
    block, counts = RESNET_LAYERS[base]
    exp = 1 if block == "basic" else 4
    shapes = {}

    def bn(name, c):
        shapes[name] = ("bn", c)

    def conv(name, co, ci, k):
        shapes[name] = ("conv", (co, ci, k, k))

    conv("conv1", 64, 3, 3), bn("bn1", 64), conv("conv2", 64, 64, 3), bn("bn2", 64), conv("conv3", 128, 64, 3), bn("bn3", 128)
    inplanes = 128
    for li, (planes, nb) in enumerate(zip((64, 128, 256, 512), counts), start=1):
        for bi in range(nb):
            p = "layer%d.%d." % (li, bi)
            if block == "basic":
                conv(p + "conv1", planes, inplanes, 3), bn(p + "bn1", planes)
                conv(p + "conv2", planes, planes, 3), bn(p + "bn2", planes)
            else:
                conv(p + "conv1", planes, inplanes, 1), bn(p + "bn1", planes)
                conv(p + "conv2", planes, planes, 3), bn(p + "bn2", planes)
                conv(p + "conv3", planes * 4, planes, 1), bn(p + "bn3", planes * 4)
            if bi == 0 and (li > 1 or inplanes != planes * exp):
                conv(p + "downsample.0", planes * exp, inplanes, 1), bn(p + "downsample.1", planes * exp)
            inplanes = planes * exp
    return shapes


def decoder_param_shapes(arch, fc_dim, num_class=150, fpn_inplanes=(256, 512, 1024, 2048)):
    arch = arch.lower()
    shapes = {}

    def bn(name, c):
        shapes[name] = ("bn", c)

    def conv(name, co, ci, k, bias=False):
        shapes[name] = ("convb" if bias else "conv", (co, ci, k, k))

    if arch in ("ppm", "ppm_deepsup"):
        for i in range(4):
            conv("ppm.%d.1" % i, 512, fc_dim, 1), bn("ppm.%d.2" % i, 512)
        conv("conv_last.0", 512, fc_dim + 4 * 512, 3), bn("conv_last.1", 512)
        conv("conv_last.4", num_class, 512, 1, bias=True)
        if arch == "ppm_deepsup":
            conv("cbr_deepsup.0", fc_dim // 4, fc_dim // 2, 3), bn("cbr_deepsup.1", fc_dim // 4)
            conv("conv_last_deepsup", num_class, fc_dim // 4, 1, bias=True)
    elif arch in ("c1", "c1_deepsup"):
        conv("cbr.0", fc_dim // 4, fc_dim, 3), bn("cbr.1", fc_dim // 4)
        conv("conv_last", num_class, fc_dim // 4, 1, bias=True)
        if arch == "c1_deepsup":
            conv("cbr_deepsup.0", fc_dim // 4, fc_dim // 2, 3), bn("cbr_deepsup.1", fc_dim // 4)
            conv("conv_last_deepsup", num_class, fc_dim // 4, 1, bias=True)
    elif arch in ("upernet", "upernet_lite"):
        fpn_dim = 512 if arch == "upernet" else 256
        for i in range(4):
            conv("ppm_conv.%d.0" % i, 512, fc_dim, 1), bn("ppm_conv.%d.1" % i, 512)
        conv("ppm_last_conv.0", fpn_dim, fc_dim + 4 * 512, 3), bn("ppm_last_conv.1", fpn_dim)
        for i, c in enumerate(fpn_inplanes[:-1]):
            conv("fpn_in.%d.0" % i, fpn_dim, c, 1), bn("fpn_in.%d.1" % i, fpn_dim)
            conv("fpn_out.%d.0.0" % i, fpn_dim, fpn_dim, 3), bn("fpn_out.%d.0.1" % i, fpn_dim)
        conv("conv_last.0.0", fpn_dim, 4 * fpn_dim, 3), bn("conv_last.0.1", fpn_dim)
        conv("conv_last.1", num_class, fpn_dim, 1, bias=True)
    else:
        raise Exception("Architecture undefined!")
    return shapes


def synth_state_dict(shapes, seed, residual_gain=None):
    """Deterministic weights for a {name: (kind, shape)} table: He-scaled convs, non-trivial BN affine + running stats.
    Uses a private CPU torch.Generator so it is reproducible wherever the same torch build runs.
    residual_gain: if given, the gamma of the LAST BatchNorm of every residual block (bn3 of a Bottleneck, bn2 of a
    BasicBlock) is multiplied by it.  Random-init BN-ReLU ResNets amplify perturbations ~8 %/layer (bf16 storage alone
    moves layer4 by 55 % and decorrelates early-layer gradients); a gain of 0.25 gives the trained-network-like,
    well-conditioned regime in which kernel bugs are distinguishable from rounding chaos (tools/debug_parity.py)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for name in sorted(shapes):
        kind, shp = shapes[name]
        if kind in ("conv", "convb"):
            co, ci, k, _ = shp
            sd[name + ".weight"] = torch.randn(shp, generator=g) * math.sqrt(2.0 / (ci * k * k))
            if kind == "convb":
                sd[name + ".bias"] = torch.randn(co, generator=g) * 0.1
        else:
            c = shp
            sd[name + ".weight"] = 0.5 + torch.rand(c, generator=g)
            sd[name + ".bias"] = torch.randn(c, generator=g) * 0.1
            sd[name + ".running_mean"] = torch.randn(c, generator=g) * 0.1
            sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
            sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
            sd[name + "._tmp_running_mean"] = sd[name + ".running_mean"].clone()
            sd[name + "._tmp_running_var"] = sd[name + ".running_var"].clone()
            sd[name + "._running_iter"] = torch.ones(1)
    if residual_gain is not None:
        blocks = {}
        for name in shapes:
            if (name.startswith("layer") or ".branches." in name) and ".bn" in name and "downsample" not in name:
                blk, bn = name.rsplit(".", 1)
                blocks[blk] = max(blocks.get(blk, ""), bn)
        for blk, bn in blocks.items():
            sd["%s.%s.weight" % (blk, bn)] = sd["%s.%s.weight" % (blk, bn)] * residual_gain
    return sd


def synth_batch(n, h, w, label_stride, seed, num_class=150):
    """Synthetic ADE20K-shaped batch: img ~ N(0,1), labels uniform in {-1..num_class-1}."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    img = torch.randn(n, 3, h, w, generator=g)
    label = torch.randint(-1, num_class, (n, h // label_stride, w // label_stride), generator=g)
    return {"img_data": img, "seg_label": label}
