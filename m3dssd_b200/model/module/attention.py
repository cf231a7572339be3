"""ANAB asymmetric non-local attention (mirror of model/module/attention.py:120-216).

Queries stay at full resolution (H*W tokens, key_ch = 168 channels); keys and
values are reduced to 1 + 16 + 64 + 256 = 337 tokens by attention-weighted
adaptive average pooling, so softmax(Q K) V costs O(337 * H * W).  There is no
1/sqrt(d) scaling and `num_psp` is unused, as in the reference.
"""
import torch
from torch import nn


class PAPAModule(nn.Module):
    def __init__(self, sizes=(1, 4, 8), dimension=2):
        super().__init__()
        pool = {1: nn.AdaptiveAvgPool1d, 2: nn.AdaptiveAvgPool2d, 3: nn.AdaptiveAvgPool3d}[dimension]
        self.stages = nn.ModuleList([pool(output_size=(s,) * dimension) for s in sizes])

    def forward(self, feats, atten):
        n, c = feats.shape[:2]
        tokens = []
        for i, stage in enumerate(self.stages):
            weighted = feats if atten is None else feats * atten[:, i:i + 1]
            tokens.append(stage(weighted).view(n, c, -1))
        return torch.cat(tokens, -1)


class ANAB(nn.Module):
    def __init__(self, ch, num_psp, psp_size=[1, 4, 8, 16], with_atten=True):
        super().__init__()
        self.inch = self.outch = ch
        self.psp_size = list(psp_size)
        self.key_num = sum(s * s for s in psp_size)
        self.key_ch = self.key_num // 2
        self.with_atten = with_atten
        self.value_conv = nn.Conv2d(ch, ch, kernel_size=1, bias=False)
        if with_atten:
            self.spatial_conv = nn.Conv2d(ch, len(psp_size), kernel_size=1, bias=False)
        self.key_conv = nn.Conv2d(ch, self.key_ch, kernel_size=1, bias=False)
        self.query_conv = nn.Conv2d(ch, self.key_ch, kernel_size=1, bias=False)
        self.key_papa = PAPAModule(sizes=psp_size)
        self.value_papa = PAPAModule(sizes=psp_size)
        self.softmax = nn.Softmax(dim=-1)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        B, C, H, W = x.shape
        query = self.query_conv(x).view(B, self.key_ch, H * W).permute(0, 2, 1)
        atten = self.sigmoid(self.spatial_conv(x)) if self.with_atten else None
        key = self.key_papa(self.key_conv(x), atten)
        value = self.value_papa(self.value_conv(x), atten).permute(0, 2, 1)
        weights = self.softmax(torch.bmm(query, key))
        new_value = torch.bmm(weights, value).permute(0, 2, 1).reshape(B, self.outch, H, W)
        return (new_value + x).contiguous()
