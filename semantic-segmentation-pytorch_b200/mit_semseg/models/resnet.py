"""Deep-stem ResNet module tree (parameter containers + hyper-parameters) behind the B200 engine.

Same attribute names, parameter shapes and initialisation as the reference (mit_semseg/models/resnet.py:24-160) so
state dicts are interchangeable; the arithmetic is NOT here — `forward` hands the tree to `mit_semseg.engine`, which
walks it (reading each conv's stride / dilation / padding, so ResnetDilated's in-place rewrites are honoured) and
launches the sm_100a kernels.
"""
import math

import torch.nn as nn

from ..lib.nn import SynchronizedBatchNorm2d
from .utils import load_url

BatchNorm2d = SynchronizedBatchNorm2d

__all__ = ['ResNet', 'resnet18', 'resnet50', 'resnet101']

model_urls = {
    'resnet18': 'http://sceneparsing.csail.mit.edu/model/pretrained_resnet/resnet18-imagenet.pth',
    'resnet50': 'http://sceneparsing.csail.mit.edu/model/pretrained_resnet/resnet50-imagenet.pth',
    'resnet101': 'http://sceneparsing.csail.mit.edu/model/pretrained_resnet/resnet101-imagenet.pth',
}


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


class _Block(nn.Module):
    """A residual block is a list of (conv, bn) stages + an optional (conv, bn) shortcut; see engine/program.py."""

    def stages(self):
        raise NotImplementedError

    def forward(self, x):
        from ..engine import functional as EF
        return EF.run_block(self, x)


class BasicBlock(_Block):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def stages(self):
        return [(self.conv1, self.bn1), (self.conv2, self.bn2)]


class Bottleneck(_Block):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def stages(self):
        return [(self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)]


class ResNet(nn.Module):
    def __init__(self, block, layers, num_classes=1000):
        super().__init__()
        self.inplanes = 128
        self.conv1 = conv3x3(3, 64, stride=2)
        self.bn1 = BatchNorm2d(64)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(64, 64)
        self.bn2 = BatchNorm2d(64)
        self.relu2 = nn.ReLU(inplace=True)
        self.conv3 = conv3x3(64, 128)
        self.bn3 = BatchNorm2d(128)
        self.relu3 = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.avgpool = nn.AvgPool2d(7, stride=1)
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        # reference initialisation (resnet.py:118-124), same module order => same RNG stream
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                BatchNorm2d(planes * block.expansion),
            )
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):
        raise NotImplementedError("the ImageNet classification head (avgpool+fc) is outside the segmentation path; "
                                  "wrap this net in models.Resnet / models.ResnetDilated")


def _build(block, layers, name, pretrained, strict=True, **kwargs):
    model = ResNet(block, layers, **kwargs)
    if pretrained:
        model.load_state_dict(load_url(model_urls[name]), strict=strict)
    return model


def resnet18(pretrained=False, **kwargs):
    return _build(BasicBlock, [2, 2, 2, 2], 'resnet18', pretrained, **kwargs)


def resnet50(pretrained=False, **kwargs):
    return _build(Bottleneck, [3, 4, 6, 3], 'resnet50', pretrained, strict=False, **kwargs)


def resnet101(pretrained=False, **kwargs):
    return _build(Bottleneck, [3, 4, 23, 3], 'resnet101', pretrained, strict=False, **kwargs)
