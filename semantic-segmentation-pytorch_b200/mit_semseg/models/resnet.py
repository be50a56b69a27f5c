"""Deep-stem ResNet module tree (parameter containers + hyper-parameters) behind the B200 engine.

Same attribute names, parameter shapes and initialisation as the reference (mit_semseg/models/resnet.py:24-160) so
state dicts are interchangeable; the arithmetic is NOT here — `forward` hands the tree to `mit_semseg.engine`, which
walks it (reading each conv's stride / dilation / padding, so ResnetDilated's in-place rewrites are honoured) and
launches the sm_100a kernels.

Module order matters twice: it is the state-dict key order, and it is the order in which the seeded initialisers draw
(tests/test_oracle.py compares both with the reference): stem conv1/bn1/relu1 .. conv3/bn3/relu3, maxpool, layer1..4
(in each first block the projection shortcut is created BEFORE the block), avgpool, fc.
"""
import math

import torch.nn as nn

from ..lib.nn import SynchronizedBatchNorm2d
from .utils import load_url

BatchNorm2d = SynchronizedBatchNorm2d

__all__ = ['ResNet', 'resnet18', 'resnet50', 'resnet101']

_URL = 'http://sceneparsing.csail.mit.edu/model/pretrained_resnet/%s-imagenet.pth'
model_urls = {name: _URL % name for name in ('resnet18', 'resnet50', 'resnet101')}

# name -> (block kind, blocks per stage, strict checkpoint loading)
_ARCHS = {'resnet18': ('basic', (2, 2, 2, 2), True), 'resnet50': ('bottleneck', (3, 4, 6, 3), False),
          'resnet101': ('bottleneck', (3, 4, 23, 3), False)}


def _conv(cin, cout, k, stride=1):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=False)


def conv3x3(in_planes, out_planes, stride=1):
    return _conv(in_planes, out_planes, 3, stride)


class _Block(nn.Module):
    """A residual block = numbered (conv_i, bn_i) stages, one shared ReLU and an optional (conv, bn) shortcut. `_SPEC`
    lists (kernel size, output width as a multiple of `planes`, carries the block's stride) per stage."""
    expansion = 1
    _SPEC = ()

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        cin, relu_at = inplanes, len(self._SPEC) if self.expansion > 1 else 1
        for i, (k, mult, strided) in enumerate(self._SPEC, start=1):
            setattr(self, 'conv%d' % i, _conv(cin, planes * mult, k, stride if strided else 1))
            setattr(self, 'bn%d' % i, BatchNorm2d(planes * mult))
            if i == relu_at:
                self.relu = nn.ReLU(inplace=True)
            cin = planes * mult
        self.downsample = downsample
        self.stride = stride

    def stages(self):
        return [(getattr(self, 'conv%d' % i), getattr(self, 'bn%d' % i)) for i in range(1, len(self._SPEC) + 1)]

    def forward(self, x):
        # A block owns parameters and describes its stages; it is scheduled (fused with its neighbours' BN / shortcut / ReLU)
        # by the step program of the network that contains it. The smallest callable units on the engine are the encoder and
        # the decoder (engine/functional.py: encoder_forward / decoder_forward); the reference's per-block forward
        # (models/resnet.py:37-53,72-92) has no stand-alone counterpart.
        raise RuntimeError("%s is executed as part of its network's step program (call the encoder, the decoder or the "
                           "SegmentationModule); a residual block cannot be called on its own on the B200 engine"
                           % type(self).__name__)


class BasicBlock(_Block):
    expansion = 1
    _SPEC = ((3, 1, True), (3, 1, False))


class Bottleneck(_Block):
    expansion = 4
    _SPEC = ((1, 1, False), (3, 1, True), (1, 4, False))


class ResNet(nn.Module):
    def __init__(self, block, layers, num_classes=1000):
        super().__init__()
        for i, (cin, cout, stride) in enumerate(((3, 64, 2), (64, 64, 1), (64, 128, 1)), start=1):
            setattr(self, 'conv%d' % i, conv3x3(cin, cout, stride))
            setattr(self, 'bn%d' % i, BatchNorm2d(cout))
            setattr(self, 'relu%d' % i, nn.ReLU(inplace=True))
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.inplanes = 128
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers), start=1):
            setattr(self, 'layer%d' % i, self._make_layer(block, planes, n, stride=1 if i == 1 else 2))
        self.avgpool = nn.AvgPool2d(7, stride=1)
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        self._init_like_the_reference()

    def _init_like_the_reference(self):
        """resnet.py:118-124: He-normal over the fan-OUT of every conv, unit BN; the Linear keeps torch's default."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / fan_out))
            elif isinstance(m, BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1):
        width = planes * block.expansion
        shortcut = None
        if stride != 1 or self.inplanes != width:
            shortcut = nn.Sequential(_conv(self.inplanes, width, 1, stride), BatchNorm2d(width))
        chain = [block(self.inplanes, planes, stride, shortcut)]
        self.inplanes = width
        chain.extend(block(width, planes) for _ in range(1, blocks))
        return nn.Sequential(*chain)

    def forward(self, x):
        raise NotImplementedError("the ImageNet classification head (avgpool+fc) is outside the segmentation path; "
                                  "wrap this net in models.Resnet / models.ResnetDilated")


def _build(name, pretrained, **kwargs):
    kind, layers, strict = _ARCHS[name]
    model = ResNet(BasicBlock if kind == 'basic' else Bottleneck, list(layers), **kwargs)
    if pretrained:
        model.load_state_dict(load_url(model_urls[name]), strict=strict)
    return model


def resnet18(pretrained=False, **kwargs):
    return _build('resnet18', pretrained, **kwargs)


def resnet50(pretrained=False, **kwargs):
    return _build('resnet50', pretrained, **kwargs)


def resnet101(pretrained=False, **kwargs):
    return _build('resnet101', pretrained, **kwargs)
