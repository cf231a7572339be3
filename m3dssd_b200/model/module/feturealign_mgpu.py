"""Feature alignment (mirror of model/module/feturealign_mgpu.py:7-208).

shape_align: a 3x3 DCNv2 whose tap offsets stretch the kernel to the shape of the
top-1 foreground anchor; center_align: a 1x1 DCNv2 that re-samples the feature
at the predicted box centre.  Offsets are *computed*, not learned; both are
zeroed where the foreground probability is <= thresh and the DCN modulation mask
is the foreground probability itself; output = DCNv2(x) + x.
"""
import torch
from torch import nn
from torch.nn.modules.utils import _pair

from ..DCNv2.dcn_v2 import DCNv2


def _top1(prob, k):
    prob_k, ind = torch.topk(prob, k=k, dim=1)
    soft = torch.softmax(prob_k, dim=1)
    mask = prob_k.max(dim=1, keepdim=True)[0]
    return ind, soft, mask


class center_align(nn.Module):
    def __init__(self, ch, anchors, xy_mean, xy_std, feat_stride, feat_size, kernel_size=1, k=1, thresh=0.5):
        super().__init__()
        self.ch = ch
        self.kernel_size = _pair(kernel_size)
        self.anchors = torch.as_tensor(anchors).clone().float().detach()
        self.num_anchors = self.anchors.shape[0]
        self.k = k
        self.feat_stride = feat_stride
        self.thresh = thresh
        self.xy_mean = torch.tensor([float(v) for v in xy_mean])
        self.xy_std = torch.tensor([float(v) for v in xy_std])
        self.feat_size = feat_size
        self.anchors_w = ((self.anchors[:, 2] - self.anchors[:, 0]) / feat_stride).view(1, -1, 1, 1)
        self.anchors_h = ((self.anchors[:, 3] - self.anchors[:, 1]) / feat_stride).view(1, -1, 1, 1)
        self.align = DCNv2(ch, ch, self.kernel_size, 1, kernel_size // 2, dilation=1, deformable_groups=1)

    def forward(self, x, bbox_x, bbox_y, prob):
        dev = x.device
        aw, ah = self.anchors_w.to(dev), self.anchors_h.to(dev)
        mean, std = self.xy_mean.to(dev), self.xy_std.to(dev)
        ind, soft, mask = _top1(prob, self.k)
        hard = (mask > self.thresh).float()
        off_x = torch.gather((bbox_x * std[0] + mean[0]) * aw, 1, ind)
        off_y = torch.gather((bbox_y * std[1] + mean[1]) * ah, 1, ind)
        off_x = (off_x * soft).sum(dim=1, keepdim=True) * hard
        off_y = (off_y * soft).sum(dim=1, keepdim=True) * hard
        taps = self.kernel_size[0] * self.kernel_size[1]
        offset = torch.cat([off_y, off_x], dim=1).repeat(1, taps, 1, 1)
        return self.align(x, offset, mask.repeat(1, taps, 1, 1)) + x


class shape_align(nn.Module):
    def __init__(self, ch, anchors, feat_stride, feat_size, kernel_size=3, k=1, thresh=0.5):
        super().__init__()
        self.ch = ch
        self.kernel_size = _pair(kernel_size)
        self.anchors = torch.as_tensor(anchors).clone().float().detach()
        self.num_anchors = self.anchors.shape[0]
        self.feat_stride = feat_stride
        self.feat_size = feat_size
        self.k = k
        self.thresh = thresh
        kh, kw = self.kernel_size
        h_step = (self.anchors[:, 3] - self.anchors[:, 1]) / feat_stride / kh
        w_step = (self.anchors[:, 2] - self.anchors[:, 0]) / feat_stride / kw
        table = torch.zeros(self.num_anchors, 2 * kh * kw)  # per-anchor tap offsets (the reference tiles this over H x W)
        for i in range(kh):
            for j in range(kw):
                t = i * kw + j
                table[:, 2 * t] = (h_step - 1) * (i - kh / 2 + 0.5)
                table[:, 2 * t + 1] = (w_step - 1) * (j - kw / 2 + 0.5)
        self.offset_table = table
        self.align = DCNv2(ch, ch, self.kernel_size, 1, kernel_size // 2, 1, deformable_groups=1)
        self.proj = nn.Conv2d(ch * 2, ch, 1, bias=False)  # present in checkpoints, never used in forward (reference quirk)

    def forward(self, x, prob):
        ind, soft, mask = _top1(prob, self.k)
        hard = (mask > self.thresh).float()
        table = self.offset_table.to(x.device)
        # [B, k, H, W] anchor ids -> [B, k, H, W, 2*taps] -> weighted sum over k -> [B, 2*taps, H, W]
        offset = (table[ind] * soft.unsqueeze(-1)).sum(dim=1).permute(0, 3, 1, 2) * hard
        taps = self.kernel_size[0] * self.kernel_size[1]
        return self.align(x, offset.contiguous(), mask.repeat(1, taps, 1, 1)) + x
