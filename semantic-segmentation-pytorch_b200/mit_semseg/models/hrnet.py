"""HRNetV2-W48 module tree (parameter containers + hyper-parameters) behind the B200 engine.

Same attribute names, parameter shapes, default initialisers and construction order as the reference
(mit_semseg/models/hrnet.py:34-445) so state dicts and seeded initialisations are interchangeable.  The arithmetic is
not here: `HRNetV2.forward` hands the tree to `mit_semseg.engine`, which turns the parallel-branch / exchange-unit
structure into a kernel schedule (implicit-GEMM convolutions, one fused "sum of affine / bilinear terms + ReLU" kernel
per exchange output).

Topology (hrnet.py:254-437): stem conv1/conv2 (3x3, stride 2 each) -> layer1 = 4 Bottlenecks (64 -> 256) ->
transition1 -> stage2 (1 module, 2 branches: 48, 96 ch) -> transition2 -> stage3 (4 modules, 3 branches: +192) ->
transition3 -> stage4 (3 modules, 4 branches: +384) -> all branches bilinearly up-sampled to the 1/4-resolution map and
concatenated (720 ch).  Every branch is 4 BasicBlocks; every module ends in an exchange unit where output i =
ReLU(sum_j f_ij(x_j)): identity (j = i), 1x1 conv + BN + bilinear up (j > i), a chain of stride-2 3x3 conv + BN
(+ReLU between links) (j < i).
"""
import torch.nn as nn

from ..lib.nn import SynchronizedBatchNorm2d
from .utils import load_url

BatchNorm2d = SynchronizedBatchNorm2d
BN_MOMENTUM = 0.1

__all__ = ['hrnetv2']

model_urls = {
    'hrnetv2': 'http://sceneparsing.csail.mit.edu/model/pretrained_resnet/hrnetv2_w48-imagenet.pth',
}

# (modules, blocks per branch, channels per branch) of stages 2..4 — hrnet.py:257-262; Bottlenecks in layer1 — :271
_LAYER1_BLOCKS = 4
_STAGES = {
    'STAGE2': dict(NUM_MODULES=1, NUM_BRANCHES=2, BLOCK='BASIC', NUM_BLOCKS=(4, 4), NUM_CHANNELS=(48, 96),
                   FUSE_METHOD='SUM'),
    'STAGE3': dict(NUM_MODULES=4, NUM_BRANCHES=3, BLOCK='BASIC', NUM_BLOCKS=(4, 4, 4), NUM_CHANNELS=(48, 96, 192),
                   FUSE_METHOD='SUM'),
    'STAGE4': dict(NUM_MODULES=3, NUM_BRANCHES=4, BLOCK='BASIC', NUM_BLOCKS=(4, 4, 4, 4),
                   NUM_CHANNELS=(48, 96, 192, 384), FUSE_METHOD='SUM'),
}


def _bn(c):
    return BatchNorm2d(c, momentum=BN_MOMENTUM)


def _conv(cin, cout, k, stride=1):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=False)


def _conv_bn(cin, cout, k, stride=1, relu=False):
    mods = [_conv(cin, cout, k, stride), _bn(cout)]
    if relu:
        mods.append(nn.ReLU(inplace=True))
    return nn.Sequential(*mods)


class _Block(nn.Module):
    """(conv, bn) stages + optional (conv, bn) projection shortcut; the engine reads them through `stages()`."""

    def stages(self):
        raise NotImplementedError


class BasicBlock(_Block):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 3, stride)
        self.bn1 = _bn(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _conv(planes, planes, 3)
        self.bn2 = _bn(planes)
        self.downsample = downsample
        self.stride = stride

    def stages(self):
        return [(self.conv1, self.bn1), (self.conv2, self.bn2)]


class Bottleneck(_Block):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 1)
        self.bn1 = _bn(planes)
        self.conv2 = _conv(planes, planes, 3, stride)
        self.bn2 = _bn(planes)
        self.conv3 = _conv(planes, planes * self.expansion, 1)
        self.bn3 = _bn(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def stages(self):
        return [(self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)]


blocks_dict = {'BASIC': BasicBlock, 'BOTTLENECK': Bottleneck}


def _block_chain(block, inplanes, planes, n, stride=1):
    """n residual blocks; the first one projects its shortcut when the shape changes (shortcut built first, as the
    reference does, so seeded default initialisers draw in the same order)."""
    shortcut = None
    if stride != 1 or inplanes != planes * block.expansion:
        shortcut = _conv_bn(inplanes, planes * block.expansion, 1, stride)
    chain = [block(inplanes, planes, stride, shortcut)]
    chain += [block(planes * block.expansion, planes) for _ in range(1, n)]
    return nn.Sequential(*chain)


class HighResolutionModule(nn.Module):
    """Parallel branches followed by an exchange unit (hrnet.py:105-245).
    fuse_layers[i][j]: None (j == i) | Sequential(1x1 conv, BN) (j > i; up-sampled by the engine) |
    Sequential of (i - j) stride-2 3x3 conv+BN links, ReLU after all but the last (j < i)."""

    def __init__(self, num_branches, blocks, num_blocks, num_inchannels, num_channels, fuse_method,
                 multi_scale_output=True):
        super().__init__()
        for what, seq in (('NUM_BLOCKS', num_blocks), ('NUM_CHANNELS', num_channels), ('NUM_INCHANNELS', num_inchannels)):
            if num_branches != len(seq):
                raise ValueError('NUM_BRANCHES({}) <> {}({})'.format(num_branches, what, len(seq)))
        self.num_inchannels = num_inchannels
        self.fuse_method = fuse_method
        self.num_branches = num_branches
        self.multi_scale_output = multi_scale_output
        branches = []
        for b in range(num_branches):
            branches.append(_block_chain(blocks, self.num_inchannels[b], num_channels[b], num_blocks[b]))
            self.num_inchannels[b] = num_channels[b] * blocks.expansion
        self.branches = nn.ModuleList(branches)
        self.fuse_layers = self._exchange_unit()
        self.relu = nn.ReLU(inplace=True)

    def _exchange_unit(self):
        if self.num_branches == 1:
            return None
        ch = self.num_inchannels
        rows = []
        for i in range(self.num_branches if self.multi_scale_output else 1):
            row = []
            for j in range(self.num_branches):
                if j > i:
                    row.append(_conv_bn(ch[j], ch[i], 1))
                elif j == i:
                    row.append(None)
                else:
                    links = [_conv_bn(ch[j], ch[j], 3, 2, relu=True) for _ in range(i - j - 1)]
                    links.append(_conv_bn(ch[j], ch[i], 3, 2))
                    row.append(nn.Sequential(*links))
            rows.append(nn.ModuleList(row))
        return nn.ModuleList(rows)

    def get_num_inchannels(self):
        return self.num_inchannels


class HRNetV2(nn.Module):
    def __init__(self, n_class, **kwargs):
        super().__init__()
        self.conv1 = _conv(3, 64, 3, 2)
        self.bn1 = _bn(64)
        self.conv2 = _conv(64, 64, 3, 2)
        self.bn2 = _bn(64)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = _block_chain(Bottleneck, 64, 64, _LAYER1_BLOCKS)

        pre = [256]
        for idx, key in ((2, 'STAGE2'), (3, 'STAGE3'), (4, 'STAGE4')):
            cfg = dict(_STAGES[key])
            setattr(self, 'stage%d_cfg' % idx, cfg)
            block = blocks_dict[cfg['BLOCK']]
            cur = [c * block.expansion for c in cfg['NUM_CHANNELS']]
            setattr(self, 'transition%d' % (idx - 1), self._make_transition_layer(pre, cur))
            stage, pre = self._make_stage(cfg, cur)
            setattr(self, 'stage%d' % idx, stage)

    def _make_transition_layer(self, num_channels_pre_layer, num_channels_cur_layer):
        """hrnet.py:307-341: existing branches get a 3x3 conv only if their width changes; each new branch is a chain
        of stride-2 3x3 convs from the LAST previous branch."""
        npre = len(num_channels_pre_layer)
        layers = []
        for i, cout in enumerate(num_channels_cur_layer):
            if i < npre:
                cin = num_channels_pre_layer[i]
                layers.append(_conv_bn(cin, cout, 3, 1, relu=True) if cin != cout else None)
            else:
                cin = num_channels_pre_layer[-1]
                steps = i + 1 - npre
                layers.append(nn.Sequential(*[_conv_bn(cin, cout if s == steps - 1 else cin, 3, 2, relu=True)
                                              for s in range(steps)]))
        return nn.ModuleList(layers)

    def _make_stage(self, layer_config, num_inchannels, multi_scale_output=True):
        block = blocks_dict[layer_config['BLOCK']]
        n = layer_config['NUM_MODULES']
        modules = []
        for i in range(n):
            multi = multi_scale_output or i != n - 1
            modules.append(HighResolutionModule(layer_config['NUM_BRANCHES'], block, layer_config['NUM_BLOCKS'],
                                                num_inchannels, layer_config['NUM_CHANNELS'],
                                                layer_config['FUSE_METHOD'], multi))
            num_inchannels = modules[-1].get_num_inchannels()
        return nn.Sequential(*modules), num_inchannels

    def forward(self, x, return_feature_maps=False):
        from ..engine import functional as EF
        return EF.encoder_forward(self, x)   # one 720-channel map either way (hrnet.py:437)


def hrnetv2(pretrained=False, **kwargs):
    model = HRNetV2(n_class=1000, **kwargs)
    if pretrained:
        model.load_state_dict(load_url(model_urls['hrnetv2']), strict=False)
    return model
