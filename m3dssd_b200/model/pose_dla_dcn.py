"""DLA backbone + DLAUp/IDAUp aggregation (mirror of model/pose_dla_dcn.py).

Same public classes, constructor arguments and state-dict key names as the
reference (BasicBlock :93-121, Bottleneck :162-204, Root :251-269, Tree :272-327,
DLA :330-397, dla34 :419-425, dla102 :435-441, DeformConv :471-485, IDAUp
:519-552, DLAUp :556-578, DLASeg :641-696) so reference checkpoints load
unchanged.  These modules are parameter containers plus a module-by-module
forward (torch ops for the dense convs, the C-ABI DCNv2 operator for the
deformable ones) that training uses; inference goes through the fused engine
(m3dssd_b200/engine.py), which reads the parameters from here.
"""
import math

import numpy as np
import torch
from torch import nn

from .DCNv2.dcn_v2 import DCN

BN_MOMENTUM = 0.1


def _act():
    return nn.LeakyReLU(inplace=True)  # slope 0.01 everywhere (pose_dla_dcn.py:100)


class BasicBlock(nn.Module):
    def __init__(self, inplanes, planes, stride=1, dilation=1):
        super().__init__()
        # both convs carry a bias in this model (differs from stock DLA)
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=True, dilation=dilation)
        self.bn1 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.relu = _act()
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=1, padding=1, bias=True, dilation=dilation)
        self.bn2 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.stride = stride

    def forward(self, x, residual=None):
        skip = x if residual is None else residual
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return self.relu(y + skip)


class Bottleneck(nn.Module):
    expansion = 2

    def __init__(self, inplanes, planes, stride=1, dilation=1):
        super().__init__()
        mid = planes // Bottleneck.expansion
        self.conv1 = nn.Conv2d(inplanes, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid, momentum=BN_MOMENTUM)
        self.conv2 = nn.Conv2d(mid, mid, 3, stride=stride, padding=dilation, bias=False, dilation=dilation)
        self.bn2 = nn.BatchNorm2d(mid, momentum=BN_MOMENTUM)
        self.conv3 = nn.Conv2d(mid, planes, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.relu = _act()
        self.stride = stride

    def forward(self, x, residual=None):
        skip = x if residual is None else residual
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return self.relu(y + skip)


class Root(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, residual):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, 1, stride=1, bias=False, padding=(kernel_size - 1) // 2)
        self.bn = nn.BatchNorm2d(out_channels, momentum=BN_MOMENTUM)
        self.relu = _act()
        self.residual = residual

    def forward(self, *x):
        y = self.bn(self.conv(torch.cat(x, 1)))
        if self.residual:
            y = y + x[0]
        return self.relu(y)


class Tree(nn.Module):
    def __init__(self, levels, block, in_channels, out_channels, stride=1, level_root=False, root_dim=0,
                 root_kernel_size=1, dilation=1, root_residual=False):
        super().__init__()
        if root_dim == 0:
            root_dim = 2 * out_channels
        if level_root:
            root_dim += in_channels
        if levels == 1:
            self.tree1 = block(in_channels, out_channels, stride, dilation=dilation)
            self.tree2 = block(out_channels, out_channels, 1, dilation=dilation)
            self.root = Root(root_dim, out_channels, root_kernel_size, root_residual)
        else:
            self.tree1 = Tree(levels - 1, block, in_channels, out_channels, stride, root_dim=0,
                              root_kernel_size=root_kernel_size, dilation=dilation, root_residual=root_residual)
            self.tree2 = Tree(levels - 1, block, out_channels, out_channels, root_dim=root_dim + out_channels,
                              root_kernel_size=root_kernel_size, dilation=dilation, root_residual=root_residual)
        self.level_root = level_root
        self.root_dim = root_dim
        self.levels = levels
        self.downsample = nn.MaxPool2d(stride, stride=stride) if stride > 1 else None
        self.project = None
        if in_channels != out_channels:
            self.project = nn.Sequential(nn.Conv2d(in_channels, out_channels, 1, stride=1, bias=False),
                                         nn.BatchNorm2d(out_channels, momentum=BN_MOMENTUM))

    def forward(self, x, residual=None, children=None):
        children = [] if children is None else children
        bottom = self.downsample(x) if self.downsample is not None else x
        residual = self.project(bottom) if self.project is not None else bottom
        if self.level_root:
            children.append(bottom)
        x1 = self.tree1(x, residual)
        if self.levels == 1:
            x2 = self.tree2(x1)
            return self.root(x2, x1, *children)
        children.append(x1)
        return self.tree2(x1, children=children)


class DLA(nn.Module):
    def __init__(self, levels, channels, num_classes=1000, block=BasicBlock, residual_root=False, linear_root=False):
        super().__init__()
        self.channels = channels
        self.num_classes = num_classes
        self.base_layer = nn.Sequential(nn.Conv2d(3, channels[0], 7, stride=1, padding=3, bias=False),
                                        nn.BatchNorm2d(channels[0], momentum=BN_MOMENTUM), _act())
        self.level0 = self._make_conv_level(channels[0], channels[0], levels[0])
        self.level1 = self._make_conv_level(channels[0], channels[1], levels[1], stride=2)
        for i in range(2, 6):
            setattr(self, "level%d" % i, Tree(levels[i], block, channels[i - 1], channels[i], 2, level_root=i > 2,
                                              root_residual=residual_root))

    @staticmethod
    def _make_conv_level(inplanes, planes, convs, stride=1, dilation=1):
        mods = []
        for i in range(convs):
            mods += [nn.Conv2d(inplanes, planes, 3, stride=stride if i == 0 else 1, padding=dilation, bias=False,
                               dilation=dilation),
                     nn.BatchNorm2d(planes, momentum=BN_MOMENTUM), _act()]
            inplanes = planes
        return nn.Sequential(*mods)

    def forward(self, x):
        y = []
        x = self.base_layer(x)
        for i in range(6):
            x = getattr(self, "level%d" % i)(x)
            y.append(x)
        return y

    def load_pretrained_model(self, data="imagenet", name="dla34", hash="ba72cf86"):
        """The reference downloads ImageNet weights here (model_zoo.load_url, pose_dla_dcn.py:399-416) and, as a
        side effect, registers `self.fc` for good -- so every checkpoint trained with conf.pre_train=True (all the
        shipped configs) carries base.base.fc.weight / .bias.  There is no network in this environment: the
        download is skipped with a warning, the `fc` module is registered all the same so those checkpoints load
        with strict=True; the detector's own checkpoint then overwrites every trunk weight anyway."""
        import warnings
        warnings.warn("m3dssd_b200: pre_train requested but ImageNet weights (%s/%s-%s) cannot be downloaded here; "
                      "the trunk keeps its initialisation until a checkpoint is loaded" % (data, name, hash))
        self.fc = nn.Conv2d(self.channels[-1], self.num_classes, kernel_size=1, stride=1, padding=0, bias=True)


def dla34(pretrained=True, **kwargs):
    model = DLA([1, 1, 1, 2, 2, 1], [16, 32, 64, 128, 256, 512], block=BasicBlock, **kwargs)
    if pretrained:
        model.load_pretrained_model(data="imagenet", name="dla34", hash="ba72cf86")
    return model


def dla102(pretrained=None, **kwargs):
    Bottleneck.expansion = 2
    model = DLA([1, 1, 1, 3, 4, 1], [16, 32, 128, 256, 512, 1024], block=Bottleneck, residual_root=True, **kwargs)
    if pretrained:  # the reference tests `is not None` (quirk, pose_dla_dcn.py:439); any falsy value means "no download" here
        model.load_pretrained_model(data="imagenet", name="dla102", hash="d94d9790")
    return model


def fill_up_weights(up):
    """Bilinear kernel for the depthwise ConvTranspose2d (pose_dla_dcn.py:459-468)."""
    w = up.weight.data
    k = w.size(2)
    f = math.ceil(k / 2)
    c = (2 * f - 1 - f % 2) / (2.0 * f)
    ramp = torch.tensor([1 - abs(i / f - c) for i in range(k)], dtype=w.dtype)
    w[:, 0] = ramp[:, None] * ramp[None, :]


class DeformConv(nn.Module):
    def __init__(self, chi, cho):
        super().__init__()
        self.actf = nn.Sequential(nn.BatchNorm2d(cho, momentum=BN_MOMENTUM), _act())
        self.conv = DCN(chi, cho, kernel_size=3, stride=1, padding=1, dilation=1, deformable_groups=1)

    def forward(self, x):
        return self.actf(self.conv(x))


class IDAUp(nn.Module):
    def __init__(self, o, channels, up_f, conf):
        super().__init__()
        self.out_channels = channels
        for i in range(1, len(channels)):
            c, f = channels[i], int(up_f[i])
            if conf.ida_dcnv2:
                proj, node = DeformConv(c, o), DeformConv(o, o)
            else:
                proj, node = nn.Conv2d(c, o, 3, 1, 1), nn.Conv2d(o, o, 3, 1, 1)
            up = nn.ConvTranspose2d(o, o, f * 2, stride=f, padding=f // 2, output_padding=0, groups=o, bias=False)
            fill_up_weights(up)
            setattr(self, "proj_%d" % i, proj)
            setattr(self, "up_%d" % i, up)
            setattr(self, "node_%d" % i, node)

    def forward(self, layers, startp, endp):
        for i in range(startp + 1, endp):
            k = i - startp
            up, proj, node = (getattr(self, "%s_%d" % (n, k)) for n in ("up", "proj", "node"))
            layers[i] = node(up(proj(layers[i])) + layers[i - 1])


class DLAUp(nn.Module):
    def __init__(self, startp, channels, scales, in_channels=None, conf=None):
        super().__init__()
        self.startp = startp
        in_channels = list(channels) if in_channels is None else in_channels
        self.channels = channels
        channels = list(channels)
        scales = np.array(scales, dtype=int)
        for i in range(len(channels) - 1):
            j = -i - 2
            setattr(self, "ida_%d" % i, IDAUp(channels[j], in_channels[j:], scales[j:] // scales[j], conf=conf))
            scales[j + 1:] = scales[j]
            in_channels[j + 1:] = [channels[j] for _ in channels[j + 1:]]

    def forward(self, layers):
        out = [layers[-1]]
        for i in range(len(layers) - self.startp - 1):
            getattr(self, "ida_%d" % i)(layers, len(layers) - i - 2, len(layers))
            out.insert(0, layers[-1])
        return out


class DLASeg(nn.Module):
    def __init__(self, base_name, pretrained, down_ratio, final_kernel, last_level, head_conv, conf, out_channel=0):
        super().__init__()
        assert down_ratio in [2, 4, 8, 16]
        self.first_level = int(np.log2(down_ratio))
        self.last_level = last_level
        self.base_name = base_name
        self.base = {"dla34": dla34, "dla102": dla102}[base_name](pretrained=pretrained)
        channels = self.base.channels
        scales = [2 ** i for i in range(len(channels[self.first_level:]))]
        self.dla_up = DLAUp(self.first_level, channels[self.first_level:], scales, conf=conf)
        if out_channel == 0:
            out_channel = channels[self.first_level]
        self.out_channels = out_channel
        self.ida_up = IDAUp(out_channel, channels[self.first_level:self.last_level],
                            [2 ** i for i in range(self.last_level - self.first_level)], conf)

    def forward(self, x):
        x = self.dla_up(self.base(x))
        y = [x[i].clone() for i in range(self.last_level - self.first_level)]
        self.ida_up(y, 0, len(y))
        return y[-1]
