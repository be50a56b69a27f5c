"""ORACLE — TEST INFRASTRUCTURE ONLY.  Independent numpy restatement of the primitives the reference delegates to
PyTorch ATen (third-party; reference pins torch>=0.4.1, this container runs 2.11.0).  Pure numpy loops / einsum, meant
for SMALL cases: tests/test_oracle.py cross-checks them against torch and they document the exact index conventions the
CUDA kernels implement.

    conv2d                nn.Conv2d call sites: models/resnet.py:18-21,61-66,130-131; models/models.py:160-167,449-462
    batch_norm_train      F.batch_norm (lib/nn/modules/batchnorm.py:58-61) and the sync formula (:123-139)
    max_pool_3x3_s2       nn.MaxPool2d(3, 2, 1)   models/resnet.py:109
    adaptive_avg_pool     nn.AdaptiveAvgPool2d    models/models.py:447   bins [floor(i*H/s), ceil((i+1)*H/s))
    bilinear              F.interpolate(bilinear, align_corners=False)   models/models.py:472-475
    log_softmax_nll       F.log_softmax + nn.NLLLoss(ignore_index=-1)    models/models.py:492-493, train.py:154
"""
import math

import numpy as np


def conv2d(x, w, bias=None, stride=1, padding=0, dilation=1):
    n, ci, h, wd = x.shape
    co, _, kh, kw = w.shape
    ho = (h + 2 * padding - dilation * (kh - 1) - 1) // stride + 1
    wo = (wd + 2 * padding - dilation * (kw - 1) - 1) // stride + 1
    xp = np.zeros((n, ci, h + 2 * padding, wd + 2 * padding), x.dtype)
    xp[:, :, padding:padding + h, padding:padding + wd] = x
    y = np.zeros((n, co, ho, wo), np.float64)
    for r in range(kh):
        for s in range(kw):
            patch = xp[:, :, r * dilation: r * dilation + stride * (ho - 1) + 1: stride,
                       s * dilation: s * dilation + stride * (wo - 1) + 1: stride]
            y += np.einsum("nchw,oc->nohw", patch.astype(np.float64), w[:, :, r, s].astype(np.float64))
    if bias is not None:
        y += bias.reshape(1, -1, 1, 1)
    return y.astype(np.float32)


def batch_norm_train(x, gamma, beta, eps=1e-5, sync_formula=False):
    """returns (y, mean, inv_std, unbiased_var).  sync_formula: clamp(var, eps)^-0.5 instead of (var+eps)^-0.5."""
    n, c, h, w = x.shape
    xs = x.transpose(1, 0, 2, 3).reshape(c, -1).astype(np.float64)
    size = xs.shape[1]
    s, ss = xs.sum(1), (xs ** 2).sum(1)
    mean = s / size
    sumvar = ss - s * mean
    var = sumvar / size
    inv_std = 1.0 / np.sqrt(np.maximum(var, eps)) if sync_formula else 1.0 / np.sqrt(var + eps)
    y = (x - mean.reshape(1, c, 1, 1)) * (inv_std * gamma).reshape(1, c, 1, 1) + beta.reshape(1, c, 1, 1)
    return y.astype(np.float32), mean, inv_std, sumvar / (size - 1)


def max_pool_3x3_s2(x):
    n, c, h, w = x.shape
    ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    xp = np.full((n, c, h + 2, w + 2), -np.inf, x.dtype)
    xp[:, :, 1:h + 1, 1:w + 1] = x
    y = np.full((n, c, ho, wo), -np.inf, x.dtype)
    for r in range(3):
        for s in range(3):
            y = np.maximum(y, xp[:, :, r: r + 2 * (ho - 1) + 1: 2, s: s + 2 * (wo - 1) + 1: 2])
    return y


def adaptive_avg_pool(x, s):
    n, c, h, w = x.shape
    y = np.zeros((n, c, s, s), np.float64)
    for i in range(s):
        h0, h1 = (i * h) // s, -((-(i + 1) * h) // s)
        for j in range(s):
            w0, w1 = (j * w) // s, -((-(j + 1) * w) // s)
            y[:, :, i, j] = x[:, :, h0:h1, w0:w1].astype(np.float64).mean(axis=(2, 3))
    return y.astype(np.float32)


def _lin_coeff(dst, n_in, n_out):
    src = max((dst + 0.5) * (n_in / n_out) - 0.5, 0.0)
    i0 = int(math.floor(src))
    i1 = min(i0 + 1, n_in - 1)
    return i0, i1, src - i0


def bilinear(x, ho, wo):
    n, c, h, w = x.shape
    y = np.zeros((n, c, ho, wo), np.float64)
    for a in range(ho):
        h0, h1, lh = _lin_coeff(a, h, ho)
        for b in range(wo):
            w0, w1, lw = _lin_coeff(b, w, wo)
            y[:, :, a, b] = (1 - lh) * ((1 - lw) * x[:, :, h0, w0] + lw * x[:, :, h0, w1]) + \
                lh * ((1 - lw) * x[:, :, h1, w0] + lw * x[:, :, h1, w1])
    return y.astype(np.float32)


def log_softmax_nll(logits, label, ignore_index=-1):
    """returns (mean NLL over valid pixels, pixel accuracy as SegmentationModuleBase.pixel_acc computes it)."""
    m = logits.max(axis=1, keepdims=True)
    lse = m + np.log(np.exp(logits - m).sum(axis=1, keepdims=True))
    logp = logits - lse
    valid = label != ignore_index
    n, c, h, w = logits.shape
    idx = np.where(valid, label, 0)
    picked = np.take_along_axis(logp, idx[:, None], axis=1)[:, 0]
    loss = -(picked * valid).sum() / valid.sum()
    acc = ((logp.argmax(1) == label) & (label >= 0)).sum() / ((label >= 0).sum() + 1e-10)
    return np.float32(loss), np.float32(acc)
