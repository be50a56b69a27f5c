"""Model API of the reference (mit_semseg/models/models.py) over the B200 engine.

`ModelBuilder.build_encoder / build_decoder` and `SegmentationModule` keep the reference's signatures, module
attribute names, class-name conventions ('Conv' / 'BatchNorm' substrings drive `weights_init` and
`_nostride_dilate`) and state-dict keys.  Modules here only OWN parameters and hyper-parameters; every forward —
and, for `SegmentationModule`, the backward as well — is a schedule of sm_100a kernels assembled by
`mit_semseg.engine.program` (implicit-GEMM convolutions on tcgen05, fused BN/ReLU/residual, PPM cascade with a
virtual concat, fused log-softmax/NLL/accuracy).  There is no PyTorch-operator fallback.

Supported on the engine in this build: resnet18/50/101 (+dilated), hrnetv2 and (inference only) mobilenetv2dilated encoders; ppm, ppm_deepsup, c1, c1_deepsup,
upernet, upernet_lite decoders.  Other reference arch names are recognised and raise NotImplementedError with an explanation
(unknown names raise the reference's Exception('Architecture undefined!')).
"""
from functools import partial

import torch
import torch.nn as nn

from ..lib.nn import SynchronizedBatchNorm2d
from . import hrnet, mobilenet, resnet

BatchNorm2d = SynchronizedBatchNorm2d


class SegmentationModuleBase(nn.Module):
    def pixel_acc(self, pred, label):
        """Reference models.py:12-18 (kept for API parity; the training path computes it inside the loss kernel)."""
        _, preds = torch.max(pred, dim=1)
        valid = (label >= 0).long()
        acc_sum = torch.sum(valid * (preds == label).long())
        pixel_sum = torch.sum(valid)
        return acc_sum.float() / (pixel_sum.float() + 1e-10)


class SegmentationModule(SegmentationModuleBase):
    """forward(feed_dict, *, segSize=None): training -> (loss, acc) 0-d tensors; inference -> probs [N,C,*segSize].

    Reference models.py:21-47.  The whole step (encoder, decoder, loss, accuracy and — when gradients are enabled —
    every weight/data gradient) runs as one engine program; `loss.backward()` then only hands the already computed
    gradients to autograd.
    """

    def __init__(self, net_enc, net_dec, crit, deep_sup_scale=None):
        super().__init__()
        self.encoder = net_enc
        self.decoder = net_dec
        self.crit = crit
        self.deep_sup_scale = deep_sup_scale

    def forward(self, feed_dict, *, segSize=None):
        from ..engine import functional as EF
        if isinstance(feed_dict, (list, tuple)):
            # `user_scattered_collate` hands the loader's list through; with one GPU train.py passes it on as it is
            # (train.py:176-183 only wraps the module for len(gpus) > 1). Later upstream revisions unwrap it here too.
            assert len(feed_dict) == 1, "one process drives one GPU: expected this GPU's batch only"
            feed_dict = feed_dict[0]
        if segSize is None:
            return EF.segmentation_train_step(self, feed_dict['img_data'], feed_dict['seg_label'])
        return EF.segmentation_inference(self, feed_dict['img_data'], segSize)

    def zero_grad(self, set_to_none=True):
        """nn.Module.zero_grad (train.py:41), over a cached flat parameter list: the generic implementation walks the
        module tree through nested generators, 0.3 ms of host time per step during which the GPU is idle."""
        from ..engine import functional as EF
        EF.fast_zero_grad(self, set_to_none)


class ModelBuilder:
    @staticmethod
    def weights_init(m):
        """Reference models.py:52-59."""
        classname = m.__class__.__name__
        if classname.find('Conv') != -1:
            nn.init.kaiming_normal_(m.weight.data)
        elif classname.find('BatchNorm') != -1:
            m.weight.data.fill_(1.)
            m.bias.data.fill_(1e-4)

    @staticmethod
    def build_encoder(arch='resnet50dilated', fc_dim=512, weights=''):
        pretrained = len(weights) == 0
        arch = arch.lower()
        resnets = {'resnet18': False, 'resnet18dilated': True, 'resnet50': False, 'resnet50dilated': True,
                   'resnet101': False, 'resnet101dilated': True}
        if arch in resnets:
            base = arch.replace('dilated', '')
            orig = resnet.__dict__[base](pretrained=pretrained)
            net_encoder = ResnetDilated(orig, dilate_scale=8) if resnets[arch] else Resnet(orig)
        elif arch in ('resnet34', 'resnet34dilated'):
            raise NotImplementedError
        elif arch == 'hrnetv2':
            net_encoder = hrnet.__dict__['hrnetv2'](pretrained=pretrained)
        elif arch == 'mobilenetv2dilated':
            orig_mobilenet = mobilenet.__dict__['mobilenetv2'](pretrained=pretrained)
            net_encoder = MobileNetV2Dilated(orig_mobilenet, dilate_scale=8)
        elif arch == 'resnext101':
            raise NotImplementedError(
                "encoder '%s' is part of the reference API but not built on the B200 engine "
                "(grouped convolutions; see DESIGN.md 'out of scope')" % arch)
        else:
            raise Exception('Architecture undefined!')
        if len(weights) > 0:
            print('Loading weights for net_encoder')
            net_encoder.load_state_dict(torch.load(weights, map_location=lambda storage, loc: storage), strict=False)
        return net_encoder

    @staticmethod
    def build_decoder(arch='ppm_deepsup', fc_dim=512, num_class=150, weights='', use_softmax=False):
        arch = arch.lower()
        if arch == 'c1_deepsup':
            net_decoder = C1DeepSup(num_class=num_class, fc_dim=fc_dim, use_softmax=use_softmax)
        elif arch == 'c1':
            net_decoder = C1(num_class=num_class, fc_dim=fc_dim, use_softmax=use_softmax)
        elif arch == 'ppm':
            net_decoder = PPM(num_class=num_class, fc_dim=fc_dim, use_softmax=use_softmax)
        elif arch == 'ppm_deepsup':
            net_decoder = PPMDeepsup(num_class=num_class, fc_dim=fc_dim, use_softmax=use_softmax)
        elif arch == 'upernet_lite':
            net_decoder = UPerNet(num_class=num_class, fc_dim=fc_dim, use_softmax=use_softmax, fpn_dim=256)
        elif arch == 'upernet':
            net_decoder = UPerNet(num_class=num_class, fc_dim=fc_dim, use_softmax=use_softmax, fpn_dim=512)
        else:
            raise Exception('Architecture undefined!')
        net_decoder.apply(ModelBuilder.weights_init)
        if len(weights) > 0:
            print('Loading weights for net_decoder')
            net_decoder.load_state_dict(torch.load(weights, map_location=lambda storage, loc: storage), strict=False)
        return net_decoder


def conv3x3_bn_relu(in_planes, out_planes, stride=1):
    return nn.Sequential(
        nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False),
        BatchNorm2d(out_planes),
        nn.ReLU(inplace=True),
    )


# ------------------------------------------------------------------------------------------------ encoders
class _ResnetTrunk(nn.Module):
    """Owns the stem + four stages of an `orig_resnet` (everything but avgpool/fc), reference models.py:170-268."""

    def _adopt(self, orig_resnet):
        for name in ('conv1', 'bn1', 'relu1', 'conv2', 'bn2', 'relu2', 'conv3', 'bn3', 'relu3', 'maxpool',
                     'layer1', 'layer2', 'layer3', 'layer4'):
            setattr(self, name, getattr(orig_resnet, name))

    def forward(self, x, return_feature_maps=False):
        from ..engine import functional as EF
        conv_out = EF.encoder_forward(self, x)
        return conv_out if return_feature_maps else [conv_out[-1]]


class Resnet(_ResnetTrunk):
    def __init__(self, orig_resnet):
        super().__init__()
        self._adopt(orig_resnet)


class ResnetDilated(_ResnetTrunk):
    def __init__(self, orig_resnet, dilate_scale=8):
        super().__init__()
        if dilate_scale == 8:
            orig_resnet.layer3.apply(partial(self._nostride_dilate, dilate=2))
            orig_resnet.layer4.apply(partial(self._nostride_dilate, dilate=4))
        elif dilate_scale == 16:
            orig_resnet.layer4.apply(partial(self._nostride_dilate, dilate=2))
        self._adopt(orig_resnet)

    def _nostride_dilate(self, m, dilate):
        """Reference models.py:238-251: stride-2 convs lose their stride (3x3 get dilate//2), other 3x3 get dilate."""
        if m.__class__.__name__.find('Conv') == -1:
            return
        if m.stride == (2, 2):
            m.stride = (1, 1)
            if m.kernel_size == (3, 3):
                m.dilation = (dilate // 2, dilate // 2)
                m.padding = (dilate // 2, dilate // 2)
        elif m.kernel_size == (3, 3):
            m.dilation = (dilate, dilate)
            m.padding = (dilate, dilate)


class MobileNetV2Dilated(nn.Module):
    """Reference models.py:271-323: MobileNetV2 features without the last 1x1 conv, features[7:14] dilated by 2 and
    features[14:] by 4 (stride-2 convs lose their stride). Feature maps are returned after features[2, 4, 7, 14] and at
    the end (5 maps). Inference only on the engine (see models/mobilenet.py)."""

    def __init__(self, orig_net, dilate_scale=8):
        super().__init__()
        self.features = orig_net.features[:-1]
        self.total_idx = len(self.features)
        self.down_idx = [2, 4, 7, 14]
        if dilate_scale == 8:
            for i in range(self.down_idx[-2], self.down_idx[-1]):
                self.features[i].apply(partial(self._nostride_dilate, dilate=2))
            for i in range(self.down_idx[-1], self.total_idx):
                self.features[i].apply(partial(self._nostride_dilate, dilate=4))
        elif dilate_scale == 16:
            for i in range(self.down_idx[-1], self.total_idx):
                self.features[i].apply(partial(self._nostride_dilate, dilate=2))

    _nostride_dilate = ResnetDilated._nostride_dilate

    def forward(self, x, return_feature_maps=False):
        from ..engine import functional as EF
        conv_out = EF.encoder_forward(self, x)
        return conv_out if return_feature_maps else [conv_out[-1]]


# ------------------------------------------------------------------------------------------------ decoders
class _Decoder(nn.Module):
    def forward(self, conv_out, segSize=None):
        from ..engine import functional as EF
        return EF.decoder_forward(self, conv_out, segSize)


class C1DeepSup(_Decoder):
    def __init__(self, num_class=150, fc_dim=2048, use_softmax=False):
        super().__init__()
        self.use_softmax = use_softmax
        self.cbr = conv3x3_bn_relu(fc_dim, fc_dim // 4, 1)
        self.cbr_deepsup = conv3x3_bn_relu(fc_dim // 2, fc_dim // 4, 1)
        self.conv_last = nn.Conv2d(fc_dim // 4, num_class, 1, 1, 0)
        self.conv_last_deepsup = nn.Conv2d(fc_dim // 4, num_class, 1, 1, 0)


class C1(_Decoder):
    def __init__(self, num_class=150, fc_dim=2048, use_softmax=False):
        super().__init__()
        self.use_softmax = use_softmax
        self.cbr = conv3x3_bn_relu(fc_dim, fc_dim // 4, 1)
        self.conv_last = nn.Conv2d(fc_dim // 4, num_class, 1, 1, 0)


def _ppm_branches(fc_dim, pool_scales):
    return nn.ModuleList([
        nn.Sequential(
            nn.AdaptiveAvgPool2d(scale),
            nn.Conv2d(fc_dim, 512, kernel_size=1, bias=False),
            BatchNorm2d(512),
            nn.ReLU(inplace=True),
        ) for scale in pool_scales])


def _ppm_head(fc_dim, pool_scales, num_class):
    return nn.Sequential(
        nn.Conv2d(fc_dim + len(pool_scales) * 512, 512, kernel_size=3, padding=1, bias=False),
        BatchNorm2d(512),
        nn.ReLU(inplace=True),
        nn.Dropout2d(0.1),
        nn.Conv2d(512, num_class, kernel_size=1),
    )


class PPM(_Decoder):
    def __init__(self, num_class=150, fc_dim=4096, use_softmax=False, pool_scales=(1, 2, 3, 6)):
        super().__init__()
        self.use_softmax = use_softmax
        self.pool_scales = tuple(pool_scales)
        self.ppm = _ppm_branches(fc_dim, pool_scales)
        self.conv_last = _ppm_head(fc_dim, pool_scales, num_class)


class PPMDeepsup(_Decoder):
    def __init__(self, num_class=150, fc_dim=4096, use_softmax=False, pool_scales=(1, 2, 3, 6)):
        super().__init__()
        self.use_softmax = use_softmax
        self.pool_scales = tuple(pool_scales)
        self.ppm = _ppm_branches(fc_dim, pool_scales)
        self.cbr_deepsup = conv3x3_bn_relu(fc_dim // 2, fc_dim // 4, 1)
        self.conv_last = _ppm_head(fc_dim, pool_scales, num_class)
        self.conv_last_deepsup = nn.Conv2d(fc_dim // 4, num_class, 1, 1, 0)
        self.dropout_deepsup = nn.Dropout2d(0.1)


class UPerNet(_Decoder):
    """Reference models.py:499-586: PPM on conv5 (1x1 conv AFTER the up-sampling) -> top-down FPN with lateral 1x1 convs
    -> all levels up-sampled to the 1/4-resolution map, concatenated, fused by a 3x3 conv -> classifier."""

    def __init__(self, num_class=150, fc_dim=4096, use_softmax=False, pool_scales=(1, 2, 3, 6),
                 fpn_inplanes=(256, 512, 1024, 2048), fpn_dim=256):
        super().__init__()
        self.use_softmax = use_softmax
        self.pool_scales = tuple(pool_scales)
        # construction order = the reference's (same RNG stream for the default initialisers)
        ppm_pooling, ppm_conv = [], []
        for scale in pool_scales:
            ppm_pooling.append(nn.AdaptiveAvgPool2d(scale))
            ppm_conv.append(nn.Sequential(
                nn.Conv2d(fc_dim, 512, kernel_size=1, bias=False),
                BatchNorm2d(512),
                nn.ReLU(inplace=True)))
        self.ppm_pooling = nn.ModuleList(ppm_pooling)
        self.ppm_conv = nn.ModuleList(ppm_conv)
        self.ppm_last_conv = conv3x3_bn_relu(fc_dim + len(pool_scales) * 512, fpn_dim, 1)
        fpn_in = []
        for fpn_inplane in fpn_inplanes[:-1]:
            fpn_in.append(nn.Sequential(
                nn.Conv2d(fpn_inplane, fpn_dim, kernel_size=1, bias=False),
                BatchNorm2d(fpn_dim),
                nn.ReLU(inplace=True)))
        self.fpn_in = nn.ModuleList(fpn_in)
        fpn_out = []
        for _ in range(len(fpn_inplanes) - 1):
            fpn_out.append(nn.Sequential(conv3x3_bn_relu(fpn_dim, fpn_dim, 1)))
        self.fpn_out = nn.ModuleList(fpn_out)
        self.conv_last = nn.Sequential(
            conv3x3_bn_relu(len(fpn_inplanes) * fpn_dim, fpn_dim, 1),
            nn.Conv2d(fpn_dim, num_class, kernel_size=1))
