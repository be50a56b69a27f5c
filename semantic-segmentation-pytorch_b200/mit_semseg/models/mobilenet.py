"""MobileNetV2 module tree (parameter containers + hyper-parameters) behind the B200 engine.

Same attribute names, parameter shapes, initialisation and construction order as the reference
(mit_semseg/models/mobilenet.py:22-155), so state dicts and seeded initialisations are interchangeable. The engine runs
this network in INFERENCE only (BASELINE configs[0] is a single-image forward): 1x1 convolutions on the tcgen05
implicit-GEMM kernel with BatchNorm + ReLU6 folded into its epilogue, depthwise 3x3 convolutions on their own HBM-bound
kernel (csrc/depthwise.cu). Training this encoder is not built (no ReLU6 / depthwise backward kernels).

Topology: features[0] = 3x3 stride-2 conv-BN-ReLU6 (3 -> 32); features[1..17] = InvertedResidual blocks
(expand 1x1 -> depthwise 3x3 -> linear 1x1, shortcut when stride 1 and inp == oup); features[18] = 1x1 conv to 1280
(dropped by MobileNetV2Dilated); classifier (unused by the segmentation path, kept for state-dict parity).
"""
import math

import torch.nn as nn

from ..lib.nn import SynchronizedBatchNorm2d
from .utils import load_url

BatchNorm2d = SynchronizedBatchNorm2d

__all__ = ['mobilenetv2']

model_urls = {
    'mobilenetv2': 'http://sceneparsing.csail.mit.edu/model/pretrained_resnet/mobilenet_v2.pth.tar',
}

# (expansion t, output channels c, repeats n, stride s) — mobilenet.py:86-95
_SETTING = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1))


def _cbr6(inp, oup, k, stride=1, groups=1):
    return [nn.Conv2d(inp, oup, k, stride, k // 2, groups=groups, bias=False), BatchNorm2d(oup), nn.ReLU6(inplace=True)]


def conv_bn(inp, oup, stride):
    return nn.Sequential(*_cbr6(inp, oup, 3, stride))


def conv_1x1_bn(inp, oup):
    return nn.Sequential(*_cbr6(inp, oup, 1))


class InvertedResidual(nn.Module):
    """conv = [1x1 expand-BN-ReLU6 (absent when expand_ratio == 1)] + depthwise 3x3-BN-ReLU6 + linear 1x1-BN."""

    def __init__(self, inp, oup, stride, expand_ratio):
        super().__init__()
        assert stride in (1, 2)
        self.stride = stride
        hidden_dim = round(inp * expand_ratio)
        self.use_res_connect = self.stride == 1 and inp == oup
        layers = [] if expand_ratio == 1 else _cbr6(inp, hidden_dim, 1)
        layers += _cbr6(hidden_dim, hidden_dim, 3, stride, groups=hidden_dim)
        layers += [nn.Conv2d(hidden_dim, oup, 1, 1, 0, bias=False), BatchNorm2d(oup)]
        self.conv = nn.Sequential(*layers)

    def stages(self):
        """[(conv, bn, relu6?)] in execution order (read by the engine)."""
        mods = list(self.conv)
        out, i = [], 0
        while i < len(mods):
            act = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU6)
            out.append((mods[i], mods[i + 1], act))
            i += 3 if act else 2
        return out


class MobileNetV2(nn.Module):
    def __init__(self, n_class=1000, input_size=224, width_mult=1.):
        super().__init__()
        assert input_size % 32 == 0
        input_channel = int(32 * width_mult)
        self.last_channel = int(1280 * width_mult) if width_mult > 1.0 else 1280
        features = [conv_bn(3, input_channel, 2)]
        for t, c, n, s in _SETTING:
            output_channel = int(c * width_mult)
            for i in range(n):
                features.append(InvertedResidual(input_channel, output_channel, s if i == 0 else 1, expand_ratio=t))
                input_channel = output_channel
        features.append(conv_1x1_bn(input_channel, self.last_channel))
        self.features = nn.Sequential(*features)
        self.classifier = nn.Sequential(nn.Dropout(0.2), nn.Linear(self.last_channel, n_class))
        self._initialize_weights()

    def _initialize_weights(self):
        """mobilenet.py:129-143, same module order => same RNG stream."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()

    def forward(self, x):
        raise NotImplementedError("the ImageNet classification head is outside the segmentation path; wrap this net in "
                                  "models.MobileNetV2Dilated")


def mobilenetv2(pretrained=False, **kwargs):
    model = MobileNetV2(n_class=1000, **kwargs)
    if pretrained:
        model.load_state_dict(load_url(model_urls['mobilenetv2']), strict=False)
    return model
