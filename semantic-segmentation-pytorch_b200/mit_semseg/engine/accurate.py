"""fp32-accurate inference on the bf16 tensor cores (BASELINE config 2: forward logits within 1e-3 of the reference's fp32
path, arg-max map bit-exact — one bf16 rounding per layer gives 6e-3, measured with the oracle).

Every activation and every weight is a PAIR of bf16 tensors (hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits), and

    conv(x, w)  ~=  conv(x_hi, w_hi) + conv(x_lo, w_hi) + conv(x_hi, w_lo)                (logit error 1e-5)

is ONE launch of the ordinary tcgen05 implicit-GEMM kernel: the three products are the virtual channel concatenation
[x_hi | x_lo | x_hi] against the K-concatenated weight [w_hi | w_hi | w_lo], accumulated in fp32 in tensor memory and
written as fp32. BatchNorm (running statistics), shortcut and ReLU are applied by `sseg_split_affine`, which also splits
the result into the next layer's pair. Stride-2 convolutions run as their stride-1 twin whose output is read through a
`[::2, ::2]` view. 3x the tensor-core FLOPs of the bf16 path, for evaluation runs that must reproduce the reference.

Scope: ResNet / ResnetDilated encoders with PPM / PPMDeepsup / C1 / C1DeepSup decoders (eval mode), i.e. config 2.
Enabled by `SSEG_ACCURATE_INFERENCE=1` (engine/functional.py::segmentation_inference). Written after the last GPU run of
round 1: the schedule is verified against the fp32 oracle on the emulated ABI (tests/test_program_emulated.py), the
CUDA side (csrc/accurate.cu + the unchanged igemm kernel) has its first run next round.
"""
import torch
import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm

from . import ops


def _pad(x, m):
    return (x + m - 1) // m * m


class Pair:
    """An activation as (hi, lo) bf16 NHWC tensors of equal shape (possibly channel slices of wider buffers)."""
    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo):
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape


class AccurateInference:
    def __init__(self, seg, img_shape, seg_size, dry_run=False):
        from ..models import models as M
        self.seg, self.enc, self.dec = seg, seg.encoder, seg.decoder
        assert not seg.training, "the accurate mode is an inference mode (module.eval())"
        if not isinstance(self.enc, (M.Resnet, M.ResnetDilated)) or not isinstance(self.dec, (M.PPM, M.PPMDeepsup, M.C1, M.C1DeepSup)):
            raise NotImplementedError("accurate inference covers ResNet(-dilated) encoders with PPM / C1 decoders")
        self.N, _, self.H, self.W = img_shape
        self.seg_size = tuple(seg_size)
        self.dev = next(self.enc.parameters()).device
        assert self.dev.type == "cuda" or dry_run, "the B200 engine runs on CUDA devices only (no CPU fallback)"
        self.dry_run = bool(dry_run)
        self.fwd, self.keep = [], []
        self.graph = None
        self.img = torch.zeros(img_shape, device=self.dev, dtype=torch.float32)
        self._build()

    # ------------------------------------------------------------------------------------------ helpers
    def _new(self, *shape, dtype=torch.bfloat16):
        return torch.empty(shape, device=self.dev, dtype=dtype)

    def _pair(self, n, h, w, c):
        return Pair(self._new(n, h, w, c), self._new(n, h, w, c))

    def _bn_affine(self, bn):
        """Eval-mode BatchNorm as (scale, shift) float vectors, recomputed every run from the module's buffers."""
        C = bn.num_features
        vecs = torch.zeros(4, C, device=self.dev, dtype=torch.float32)
        w = bn.weight.detach() if bn.weight is not None else None
        b = bn.bias.detach() if bn.bias is not None else None
        running = (bn.running_mean, bn.running_var, None, None, None)
        self.fwd.append(lambda: ops.bn_finalize(None, None, 1.0, w, b, bn.eps, 0.0, ops.BN_EVAL, vecs[0], vecs[1], vecs[2],
                                                vecs[3], running=running))
        return vecs[2], vecs[3]

    def _conv(self, x, conv, bias=None):
        """fp32 NHWC output of `conv` over the pair x (stride-2: the stride-1 twin, returned as a [::2, ::2] view)."""
        O, I, kh, kw = conv.weight.shape
        assert kh == kw and conv.groups == 1 and conv.padding[0] == conv.dilation[0] * (kh // 2)
        T, dil, stride = kh * kw, conv.dilation[0], conv.stride[0]
        n, h, w, c = x.shape
        assert c == I
        w3 = torch.zeros(O, _pad(3 * T * I, 8), device=self.dev, dtype=torch.bfloat16)
        wsrc = conv.weight.detach()
        self.fwd.append(lambda: ops.prep_conv_weight_split(wsrc, w3))
        dh, dw = ops.conv_taps(kh, dil)
        geom = ops.make_geom([x.hi, x.lo, x.hi], (dh, dw), tap_koff=[t * 3 * I for t in range(T)])
        Op = _pad(O, 8)
        z = self._new(n, h, w, Op + (8 if Op != O else 0), dtype=torch.float32)   # classifier logits keep a padded pitch
        self.keep.append((geom, w3))
        wv = w3[:, :3 * T * I]
        self.fwd.append(lambda: ops.conv_igemm(geom, wv, O, z, n_store=Op, bias=bias))
        self._last_z = z
        if stride == 2:
            assert h % 2 == 0 and w % 2 == 0
            return z[:, ::2, ::2, :O], O
        assert stride == 1
        return z[..., :O], O

    def _cbr(self, x, conv, bn, relu=True, res=None, out=None):
        z, O = self._conv(x, conv)
        n, h, w, _ = z.shape
        scale, shift = self._bn_affine(bn)
        out = out if out is not None else self._pair(n, h, w, O)
        rr = (res.hi, res.lo) if res is not None else None
        self.fwd.append(lambda: ops.split_affine(z, out.hi, out.lo, scale=scale, shift=shift, res=rr, relu=relu))
        return out

    # ------------------------------------------------------------------------------------------ schedule
    def _build(self):
        from ..models import models as M
        enc, dec = self.enc, self.dec
        N, H, W = self.N, self.H, self.W
        # stem: conv1 in fp32 from the fp32 image, then two 3x3 convs and the max pool on pairs (models/resnet.py:100-109)
        ho, wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        z1 = self._new(N, ho, wo, 64, dtype=torch.float32)
        w1 = enc.conv1.weight.detach()
        self.fwd.append(lambda: ops.stem_conv_fwd_f32(self.img, w1, z1))
        s1, b1 = self._bn_affine(enc.bn1)
        x = self._pair(N, ho, wo, 64)
        self.fwd.append(lambda x=x: ops.split_affine(z1, x.hi, x.lo, scale=s1, shift=b1, relu=True))
        x = self._cbr(x, enc.conv2, enc.bn2)
        x = self._cbr(x, enc.conv3, enc.bn3)
        n, h, w, c = x.shape
        p = self._pair(n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c)
        self.fwd.append(lambda x=x, p=p: ops.maxpool_pair_fwd((x.hi, x.lo), (p.hi, p.lo)))
        x = p
        feats = []
        layers = (enc.layer1, enc.layer2, enc.layer3, enc.layer4)
        for li, layer in enumerate(layers):
            for bi, block in enumerate(layer):
                last_block = li == 3 and bi == len(layer) - 1
                res = x
                if block.downsample is not None:
                    res = self._cbr(x, block.downsample[0], block.downsample[1], relu=False)
                stages = block.stages()
                y = x
                for si, (cv, bn) in enumerate(stages):
                    final = si == len(stages) - 1
                    out = None
                    if final and last_block and isinstance(dec, (M.PPM, M.PPMDeepsup)):
                        # conv5 is written straight into the first channels of the pyramid's concat buffers
                        n_, h_, w_, _ = y.shape
                        ho_, wo_ = (h_ // cv.stride[0], w_ // cv.stride[0])
                        c5 = cv.out_channels
                        ctot = c5 + len(dec.pool_scales) * 512
                        self.cat = Pair(self._new(n_, ho_, wo_, ctot), self._new(n_, ho_, wo_, ctot))
                        out = Pair(self.cat.hi[..., :c5], self.cat.lo[..., :c5])
                    y = self._cbr(y, cv, bn, relu=True, res=res if final else None, out=out)
                x = y
            feats.append(x)
        conv5 = feats[-1]
        n, h, w, c5 = conv5.shape
        if isinstance(dec, (M.PPM, M.PPMDeepsup)):
            off = c5
            for scale, branch in zip(dec.pool_scales, dec.ppm):
                pooled = self._pair(n, scale, scale, c5)
                self.fwd.append(lambda pooled=pooled, s=scale: ops.avgpool_pair_fwd((conv5.hi, conv5.lo), s, (pooled.hi, pooled.lo)))
                y = self._cbr(pooled, branch[1], branch[2])
                dst = Pair(self.cat.hi[..., off:off + 512], self.cat.lo[..., off:off + 512])
                self.fwd.append(lambda y=y, dst=dst: ops.bilinear_pair_fwd((y.hi, y.lo), (dst.hi, dst.lo)))
                off += 512
            x = self._cbr(self.cat, dec.conv_last[0], dec.conv_last[1])       # Dropout2d is the identity in eval mode
            cls = dec.conv_last[4]
        else:
            x = self._cbr(conv5, dec.cbr[0], dec.cbr[1])
            cls = dec.conv_last
        bias = cls.bias.detach() if cls.bias is not None else None
        logits, C = self._conv(x, cls, bias=bias)
        self.num_class = C
        self.logits = logits
        hs, ws = self.seg_size
        self.probs = self._new(n, C, hs, ws, dtype=torch.float32)
        full = self._last_z[..., :_pad(C, 8)]     # the head kernel reads whole 8-channel groups (pad channels are zero)
        self.fwd.append(lambda: ops.upsample_softmax(full, C, self.probs))

    # ------------------------------------------------------------------------------------------ execution
    def load_inputs(self, img):
        self.img.copy_(img, non_blocking=True)

    def run_eager(self):
        if self.dry_run:
            raise RuntimeError("a dry-run program only describes the schedule; there is no CPU execution path")
        for f in self.fwd:
            f()

    def capture(self):
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self.run_eager()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run_eager()
        self.graph = g
        return g

    def run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.run_eager()
