"""DCNv2 / DCN modules (mirror of model/DCNv2/dcn_v2.py:14-70): same constructor
arguments, parameter names (weight, bias, conv_offset_mask.*) and initialisation;
the arithmetic runs in libm3dssd_b200.so."""
import math

import torch
from torch import nn
from torch.nn.modules.utils import _pair

from .dcn_v2_func import DCNv2Function


class DCNv2(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.deformable_groups = deformable_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        fan_in = self.in_channels * self.kernel_size[0] * self.kernel_size[1]
        bound = 1.0 / math.sqrt(fan_in)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            self.bias.zero_()

    def _op(self):
        # `precision` (None = dcn_v2_func.default_precision = "fp32"): set per module, e.g. by the mixed-precision
        # training path (m3dssd_b200.train.enable)
        return DCNv2Function(self.stride, self.padding, self.dilation, self.deformable_groups,
                             precision=getattr(self, "precision", None))

    def forward(self, input, offset, mask):
        return self._op()(input, offset, mask, self.weight, self.bias)


class DCN(DCNv2):
    """DCNv2 that predicts its own offsets and mask with a zero-initialised conv."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, deformable_groups)
        taps = self.kernel_size[0] * self.kernel_size[1]
        self.conv_offset_mask = nn.Conv2d(in_channels, deformable_groups * 3 * taps, kernel_size=self.kernel_size,
                                          stride=(stride, stride), padding=(padding, padding), bias=True)
        self.init_offset()

    def init_offset(self):
        with torch.no_grad():
            self.conv_offset_mask.weight.zero_()
            self.conv_offset_mask.bias.zero_()

    def forward(self, input):
        om = self.conv_offset_mask(input)
        third = om.shape[1] // 3
        offset, mask = om[:, :2 * third], torch.sigmoid(om[:, 2 * third:])
        return self._op()(input, offset.contiguous(), mask.contiguous(), self.weight, self.bias)
