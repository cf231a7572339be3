"""RPN_3D_loss_smp on the device with static shapes (reference: lib/loss/rpn_3d.py:659-1360).

The reference walks the batch in a Python loop, pulls index lists out of boolean masks (`torch.nonzero`), sorts the
per-image candidates for online hard-example mining and indexes the predictions with those lists: every step is a
device -> host synchronisation and a fresh shape.  Here the same arithmetic is written on [B, M] masks:

  * sampling (rpn_3d.py:826-868, 964-987): "keep the n lowest-scored of the masked anchors" = rank of every anchor in a
    per-image ascending sort of (score where masked, +inf elsewhere) compared with a per-image count tensor;
  * weights, cross-entropy, smooth-L1, IoU loss (rpn_3d.py:1110-1330): masked sums divided by masked counts.

No `.item()`, no boolean indexing, no data-dependent shape: the whole loss is CUDA-graph capturable together with the
forward / backward passes (m3dssd_b200.train.TrainStep).  Same constructor, forward signature, return value
(loss, stats) and conf fields as the reference class; targets arrive in the reference's `imobjs` dict
(lib/dataloader.py:959-982).  Differences, all outside the shipped configs (scripts/config/*.py): `focal_loss` and
`bbox_2d_lambda` follow the formulas the reference intends (its own code paths for them reference undefined names);
`hard_negatives=False` draws its random subset with torch.rand keys, not torch.randperm (same distribution, different
stream); stats whose presence depends on the data in the reference (`acc fg`, `bbox3d`, ...) are always present and
their values are detached (the reference's keep the whole autograd graph of the iteration alive until the next one).
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

IGN_FLAG = 3000


def _as(t, device, dtype=None):
    t = torch.as_tensor(t)
    return t.to(device=device, dtype=dtype) if dtype is not None else t.to(device)


def bbox_transform_inv_new(boxes, deltas, means=None, stds=None):
    """lib/rpn_util.py:1188-1276 without the in-place scaling of `deltas` (the reference multiplies the network
    output through a view; the values and gradients are the same)."""
    widths = boxes[..., 2] - boxes[..., 0] + 1.0
    heights = boxes[..., 3] - boxes[..., 1] + 1.0
    ctr_x = boxes[..., 0] + 0.5 * widths
    ctr_y = boxes[..., 1] + 0.5 * heights
    dx, dy, dw, dh = deltas[..., 0], deltas[..., 1], deltas[..., 2], deltas[..., 3]
    if stds is not None:
        dx, dy, dw, dh = dx * stds[0], dy * stds[1], dw * stds[2], dh * stds[3]
    if means is not None:
        dx, dy, dw, dh = dx + means[0], dy + means[1], dw + means[2], dh + means[3]
    pcx, pcy = dx * widths + ctr_x, dy * heights + ctr_y
    pw, ph = torch.exp(dw) * widths, torch.exp(dh) * heights
    return torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], dim=-1)


def iou_list(a, b):
    """lib/core.py:249-300, 341-400, mode='list' (areas without the +1 convention, union + 1e-8)."""
    inter = torch.clamp(torch.min(a[..., 2:4], b[..., 2:4]) - torch.max(a[..., 0:2], b[..., 0:2]), min=0)
    inter = inter[..., 0] * inter[..., 1]
    area_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    return inter / (area_a + area_b - inter + 1e-8)


def _lowest(mask, key, num):
    """mask [B, M] bool, key [B, M] float, num [B] long: the `num` masked anchors of every row with the smallest key
    (rpn_3d.py:838-852: sort ascending, keep the head)."""
    B, M = mask.shape
    order = torch.where(mask, key, torch.full_like(key, float("inf"))).argsort(dim=1)
    rank = torch.empty_like(order).scatter_(1, order, torch.arange(M, device=mask.device).expand(B, M))
    return mask & (rank < num.unsqueeze(1))


class RPN_3D_loss_smp(nn.Module):

    def __init__(self, conf):
        super().__init__()
        self.num_classes = len(conf.lbls) + 1
        self.device = conf.get("device", None) if hasattr(conf, "get") else getattr(conf, "device", None)
        self.num_anchors = conf.anchors.shape[0]
        self.register_buffer("bbox_means", torch.as_tensor(conf.bbox_means).float(), persistent=False)
        self.register_buffer("bbox_stds", torch.as_tensor(conf.bbox_stds).float(), persistent=False)
        self.register_buffer("anchors", torch.as_tensor(conf.anchors).float(), persistent=False)
        for k in ("feat_stride", "fg_fraction", "box_samples", "hard_negatives", "focal_loss", "cls_2d_lambda",
                  "iou_2d_lambda", "bbox_2d_lambda", "bbox_3d_lambda"):
            setattr(self, k, conf[k] if isinstance(conf, dict) else getattr(conf, k))
        for k in ("ign_thresh", "nms_thres", "fg_thresh", "bg_thresh_lo", "bg_thresh_hi", "best_thresh", "crop_size",
                  "bbox_3d_proj_lambda", "lbls", "ilbls", "min_gt_vis", "min_gt_h", "max_gt_h"):  # carried, unused here too
            setattr(self, k, (conf.get(k) if hasattr(conf, "get") else getattr(conf, k, None)))

    def forward(self, cls, prob, bbox_2d, bbox_3d, imobjs, feat_size=None):
        dev = cls.device
        B, M, K = cls.shape
        means = self.bbox_means.to(dev)[0]
        stds = self.bbox_stds.to(dev)[0]
        anchors = self.anchors.to(dev)
        lf = _as(imobjs["labels_fg"], dev).reshape(B, M) != 0
        lb = _as(imobjs["labels_bg"], dev).reshape(B, M) != 0
        li = _as(imobjs["labels_ign"], dev).reshape(B, M) != 0
        labels = _as(imobjs["labels"], dev, torch.long).reshape(B, M)
        t2 = _as(imobjs["bbox_2d"], dev, torch.float32)
        t3 = _as(imobjs["bbox_3d"], dev, torch.float32)
        rois_b = _as(imobjs["meta"]["rois"], dev, torch.float32)
        any_val = _as(imobjs["meta"]["any_val"], dev).reshape(B) != 0
        rois = rois_b[0]
        cls, bbox_2d, bbox_3d = cls.float(), bbox_2d.float(), bbox_3d.float()
        prob_d = prob.detach().float()
        stats = []

        # ---- box sampling (rpn_3d.py:811-987)
        own = prob_d.gather(2, labels.clamp(0, K - 1).unsqueeze(2)).squeeze(2)  # score of every anchor for its own label
        n_fg, n_bg, n_ign = lf.sum(1), lb.sum(1), li.sum(1)
        some = (n_fg > 0) | (n_ign > 0)
        case_a, case_b = any_val & some, any_val & ~some
        if math.isinf(self.box_samples):
            fg_num, bg_num = n_fg, n_bg
            bgb_num = torch.full_like(n_bg, M)
        else:
            fg_num = n_fg.clamp(max=round(M * self.box_samples * self.fg_fraction))
            bg_num = torch.minimum(torch.round(M * self.box_samples - fg_num.double()).long(), n_bg)
            bgb_num = torch.full_like(n_bg, min(round(self.box_samples * (1 - self.fg_fraction)), M))
        key = own if self.hard_negatives else torch.rand_like(own)
        ones = torch.ones_like(lf)
        # (a count of 0 or the full set means "no sub-sampling" in the reference: `if num > 0 and num != len(inds)`)
        fg_sel = torch.where((fg_num > 0).unsqueeze(1), _lowest(lf, key, fg_num), lf)
        bg_sel = torch.where((bg_num > 0).unsqueeze(1), _lowest(lb, key, bg_num), lb)
        bgb_sel = torch.where((bgb_num > 0).unsqueeze(1), _lowest(ones, key, bgb_num), ones)
        fg_mask = case_a.unsqueeze(1) & fg_sel
        bg_mask = (case_a.unsqueeze(1) & bg_sel & ~fg_sel) | (case_b.unsqueeze(1) & bgb_sel)

        # ---- accuracy statistics (rpn_3d.py:1085-1107)
        cls_pred = cls.argmax(dim=2)
        valid = labels != IGN_FLAG
        for name, m in (("fg", (labels > 0) & valid), ("bg", (labels == 0) & valid)):
            if self.cls_2d_lambda:
                acc = ((cls_pred == labels) & m).sum().float() / m.sum().clamp(min=1).float()
                stats.append({"name": name, "val": (acc).detach(), "format": "{:0.2f}", "group": "acc"})

        # ---- box weighting (rpn_3d.py:1110-1175)
        fg_cnt, bg_cnt = fg_mask.sum().float(), bg_mask.sum().float()
        weight = (fg_mask | bg_mask).float()
        if self.fg_fraction is not None:
            fg_w = (self.fg_fraction / (1 - self.fg_fraction)) * (bg_cnt / fg_cnt.clamp(min=1.0))
            weight = torch.where(fg_mask, fg_w.expand_as(weight), weight)
        if self.focal_loss:
            score = torch.where(any_val.unsqueeze(1) & valid, own, torch.zeros_like(own))  # labels_scores
            weight = weight * (1 - score) ** self.focal_loss

        loss = torch.zeros((), device=dev)
        # ---- classification loss (rpn_3d.py:1180-1196)
        if self.cls_2d_lambda:
            active = weight > 0
            ce = F.cross_entropy(cls.reshape(-1, K), labels.reshape(-1), reduction="none", ignore_index=IGN_FLAG).view(B, M)
            loss_cls = (ce * weight).clamp(min=0, max=2000)
            loss_cls = torch.where(active, loss_cls, torch.zeros_like(loss_cls)).sum() / active.sum().clamp(min=1).float()
            loss_cls = loss_cls * self.cls_2d_lambda
            loss = loss + loss_cls
            stats.append({"name": "cls", "val": (loss_cls).detach(), "format": "{:0.4f}", "group": "loss"})

        # ---- regression losses over the sampled foreground anchors (rpn_3d.py:1198-1353)
        fgf = fg_mask.float()
        nfg = fg_cnt.clamp(min=1.0)

        def fg_mean(x):  # mean over the foreground anchors; 0 when there are none (the reference skips the block)
            return (x * fgf).sum() / nfg

        if self.bbox_2d_lambda:
            l2 = F.smooth_l1_loss(bbox_2d, t2, reduction="none")
            bbox_2d_loss = sum(fg_mean(l2[..., j]) for j in range(4)) * self.bbox_2d_lambda
            loss = loss + bbox_2d_loss
            stats.append({"name": "bbox2d", "val": (bbox_2d_loss).detach(), "format": "{:0.4f}", "group": "loss"})
        if self.bbox_3d_lambda:
            l3 = F.smooth_l1_loss(bbox_3d, t3, reduction="none")
            # the reference adds x, y, z first and then (w + h + l + ry): same association here
            bbox_3d_loss = (fg_mean(l3[..., 0]) + fg_mean(l3[..., 1]) + fg_mean(l3[..., 2]))
            bbox_3d_loss = bbox_3d_loss + (fg_mean(l3[..., 3]) + fg_mean(l3[..., 4]) + fg_mean(l3[..., 5]) + fg_mean(l3[..., 6]))
            bbox_3d_loss = bbox_3d_loss * self.bbox_3d_lambda
            loss = loss + bbox_3d_loss
            stats.append({"name": "bbox3d", "val": (bbox_3d_loss).detach(), "format": "{:0.4f}", "group": "loss"})

        # depth / rotation error in absolute units (rpn_3d.py:786-806, 1056-1066)
        src = anchors[rois[:, 4].long()]
        z_dn = bbox_3d[..., 2] * stds[6] + means[6] + src[:, 4].unsqueeze(0)
        z_tar = t3[..., 2] * stds[6] + means[6] + src[:, 4].unsqueeze(0)
        src_b = anchors[rois_b[..., 4].long()]
        ry_dn = src[:, 8].unsqueeze(0) + (bbox_3d[..., 6] * stds[10] + means[10])
        ry_tar = src_b[..., 8] + (t3[..., 6] * stds[10] + means[10])
        stats.append({"name": "z", "val": fg_mean((z_tar - z_dn).abs().detach()), "format": "{:0.2f}", "group": "misc"})
        stats.append({"name": "ry", "val": fg_mean((ry_tar - ry_dn).abs().detach()), "format": "{:0.2f}", "group": "misc"})

        # 2D IoU of the decoded boxes (rpn_3d.py:1020-1030, 1335-1347)
        c2 = bbox_transform_inv_new(rois_b, bbox_2d, means=means, stds=stds)
        c2t = bbox_transform_inv_new(rois_b, t2, means=means, stds=stds)
        ious = iou_list(c2, c2t)
        stats.append({"name": "iou", "val": fg_mean(ious.detach()), "format": "{:0.2f}", "group": "acc"})
        if self.iou_2d_lambda:
            # (anchors outside the sample take IoU 1 -> log 0: no value, no gradient, no 0 * inf)
            iou_loss = fg_mean(-torch.log(torch.where(fg_mask, ious, torch.ones_like(ious)))) * self.iou_2d_lambda
            loss = loss + iou_loss
            stats.append({"name": "iou", "val": (iou_loss).detach(), "format": "{:0.4f}", "group": "loss"})
        stats.append({"name": "ttloss", "val": (loss).detach(), "format": "{:0.4f}", "group": "loss"})
        return loss, stats
