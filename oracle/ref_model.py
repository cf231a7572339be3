"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32/fp64, functional) of the
reference's dense forward path, driven by a state_dict with the reference's key
names.  It exists because /root/reference cannot travel to the GPU box: this
file can, and oracle/ref_harness.py proves (in the build container) that it
reproduces the unmodified reference modules (tests/test_oracle_model.py and the
committed fixtures in tests/golden/).

Follows:
  DLA-34 trunk ............ model/pose_dla_dcn.py:93-121 (BasicBlock), 251-269 (Root),
                            272-327 (Tree), 330-397 (DLA), 419-425 (dla34)
  DLAUp / IDAUp / DLASeg .. model/pose_dla_dcn.py:471-485, 519-552, 556-578, 641-696
  DCN wrapper ............. model/DCNv2/dcn_v2.py:64-70
  RPN heads / forward ..... model/M3d_inference_align.py:66-210, 215-313
  shape/center align ...... model/module/feturealign_mgpu.py:48-99, 153-208
  ANAB .................... model/module/attention.py:120-147, 183-216
  decode + NMS ............ lib/rpn_util.py:1416-1563, 1137-1186, 1329-1398
Dense convs / BN / pooling are torch CPU ops (the reference's own third-party
arithmetic); DCNv2 is the C oracle (or torchvision.ops.deform_conv2d, which the
C oracle is pinned against, when speed matters).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as O

# (levels, channels, block, residual_root): model/pose_dla_dcn.py:419-425 (dla34), 435-441 (dla102)
LEVELS = {"dla34": ([1, 1, 1, 2, 2, 1], [16, 32, 64, 128, 256, 512], "basic", False),
          "dla102": ([1, 1, 1, 3, 4, 1], [16, 32, 128, 256, 512, 1024], "bottleneck", True)}


def _dcn_c(x, offset, mask, w, b, stride, pad):
    out = O.dcn_v2_forward(x.numpy(), offset.numpy(), mask.numpy(), w.numpy(), b.numpy(), stride, pad, 1, 1,
                           dtype=np.float64 if x.dtype == torch.float64 else np.float32)
    return torch.from_numpy(out)


def _dcn_tv(x, offset, mask, w, b, stride, pad):
    import torchvision
    return torchvision.ops.deform_conv2d(x, offset, w, b, stride=stride, padding=pad, mask=mask)


class RefModel:
    def __init__(self, state_dict, conf, dcn="c", dtype=torch.float32):
        self.sd = {k: v.detach().to(dtype) if v.is_floating_point() else v for k, v in state_dict.items()}
        self.conf = conf
        self.dtype = dtype
        self.dcn = _dcn_c if dcn == "c" else _dcn_tv
        self.num_anchors = conf.anchors.shape[0]
        self.num_classes = len(conf.lbls) + 1
        self.anchors = torch.tensor(conf.anchors, dtype=dtype)
        self.taps = {}  # intermediate activations by name, for layer-wise parity checks

    # ------------------------------------------------------------- primitives
    def conv(self, x, p, stride=1, pad=0):
        return F.conv2d(x, self.sd[p + ".weight"], self.sd.get(p + ".bias"), stride=stride, padding=pad)

    def bn(self, x, p):
        return F.batch_norm(x, self.sd[p + ".running_mean"], self.sd[p + ".running_var"], self.sd[p + ".weight"],
                            self.sd[p + ".bias"], False, 0.0, 1e-5)

    @staticmethod
    def act(x):
        return F.leaky_relu(x, 0.01)

    # ------------------------------------------------------------------ trunk
    def basic_block(self, x, p, stride, residual=None):
        if residual is None:
            residual = x
        out = self.act(self.bn(self.conv(x, p + ".conv1", stride, 1), p + ".bn1"))
        out = self.bn(self.conv(out, p + ".conv2", 1, 1), p + ".bn2")
        return self.act(out + residual)

    def bottleneck(self, x, p, stride, residual=None):
        """model/pose_dla_dcn.py:162-204 (expansion 2): 1x1 -> 3x3 (stride) -> 1x1, + residual."""
        if residual is None:
            residual = x
        out = self.act(self.bn(self.conv(x, p + ".conv1"), p + ".bn1"))
        out = self.act(self.bn(self.conv(out, p + ".conv2", stride, 1), p + ".bn2"))
        out = self.bn(self.conv(out, p + ".conv3"), p + ".bn3")
        return self.act(out + residual)

    def block(self, x, p, stride, residual=None):
        kind = LEVELS[self.conf.back_bone][2]
        return (self.bottleneck if kind == "bottleneck" else self.basic_block)(x, p, stride, residual)

    def root(self, xs, p):
        x = self.bn(self.conv(torch.cat(xs, 1), p + ".conv"), p + ".bn")
        if LEVELS[self.conf.back_bone][3]:  # residual_root (dla102), model/pose_dla_dcn.py:262-266
            x = x + xs[0]
        return self.act(x)

    def tree(self, x, p, levels, cin, cout, stride, level_root, residual=None, children=None):
        children = [] if children is None else children
        bottom = F.max_pool2d(x, stride, stride) if stride > 1 else x
        residual = self.bn(self.conv(bottom, p + ".project.0"), p + ".project.1") if cin != cout else bottom
        if level_root:
            children.append(bottom)
        if levels == 1:
            x1 = self.block(x, p + ".tree1", stride, residual)
            x2 = self.block(x1, p + ".tree2", 1)
            return self.root([x2, x1] + children, p + ".root")
        x1 = self.tree(x, p + ".tree1", levels - 1, cin, cout, stride, False, residual)
        children.append(x1)
        return self.tree(x1, p + ".tree2", levels - 1, cout, cout, 1, False, None, children)

    def dla(self, x):
        levels, ch = LEVELS[self.conf.back_bone][:2]
        p = "base.base"
        x = self.act(self.bn(self.conv(x, p + ".base_layer.0", 1, 3), p + ".base_layer.1"))
        y = []
        x = self.act(self.bn(self.conv(x, p + ".level0.0", 1, 1), p + ".level0.1"))
        y.append(x)
        x = self.act(self.bn(self.conv(x, p + ".level1.0", 2, 1), p + ".level1.1"))
        y.append(x)
        for i in range(2, 6):
            x = self.tree(x, "%s.level%d" % (p, i), levels[i], ch[i - 1], ch[i], 2, i > 2)
            y.append(x)
        return y

    # ------------------------------------------------------------ aggregation
    def deform_conv(self, x, p):
        om = self.conv(x, p + ".conv.conv_offset_mask", 1, 1)
        o1, o2, mask = torch.chunk(om, 3, dim=1)
        offset = torch.cat((o1, o2), dim=1)
        y = self.dcn(x, offset, torch.sigmoid(mask), self.sd[p + ".conv.weight"], self.sd[p + ".conv.bias"], 1, 1)
        self.taps[p + ".conv"] = y
        self.taps[p + ".om"] = om
        out = self.act(self.bn(y, p + ".actf.0"))
        self.taps[p] = out
        return out

    def ida_up(self, layers, p, startp, endp):
        for i in range(startp + 1, endp):
            k = i - startp
            w = self.sd["%s.up_%d.weight" % (p, k)]
            f = w.shape[2] // 2
            x = self.deform_conv(layers[i], "%s.proj_%d" % (p, k))
            x = F.conv_transpose2d(x, w, None, stride=f, padding=f // 2, groups=w.shape[0])
            self.taps["%s.up_%d" % (p, k)] = x + layers[i - 1]
            layers[i] = self.deform_conv(x + layers[i - 1], "%s.node_%d" % (p, k))

    def dla_seg(self, x):
        layers = self.dla(x)
        for i, t in enumerate(layers):
            self.taps["level%d" % i] = t
        first, last = 3, 5  # log2(down_ratio=8), last_level
        layers = list(layers)
        out = [layers[-1]]
        for i in range(len(layers) - first - 1):
            self.ida_up(layers, "base.dla_up.ida_%d" % i, len(layers) - i - 2, len(layers))
            out.insert(0, layers[-1])
        y = [out[i].clone() for i in range(last - first)]
        self.ida_up(y, "base.ida_up", 0, len(y))
        return y[-1]

    # ------------------------------------------------------------------ heads
    def head(self, x, p, k0=1):
        x = self.act(self.bn(self.conv(x, p + ".0", 1, k0 // 2), p + ".1"))
        x = self.act(self.bn(self.conv(x, p + ".3"), p + ".4"))
        return self.conv(x, p + ".6")

    def _top1(self, prob):
        mask, ind = torch.max(prob, dim=1, keepdim=True)  # topk(k=1); softmax over one element = 1
        return mask, ind, (mask > 0.5).to(prob.dtype)

    def shape_align_offsets(self, prob):
        """(offset [B,18,H,W], mask [B,9,H,W]) of shape_align (feturealign_mgpu.py:119-136, 160-172)."""
        stride = self.conf.feat_stride
        aw = (self.anchors[:, 2] - self.anchors[:, 0]) / stride / 3
        ah = (self.anchors[:, 3] - self.anchors[:, 1]) / stride / 3
        mask, ind, hard = self._top1(prob)
        offs = []
        for i in range(3):
            for j in range(3):
                offs.append((ah[ind] - 1) * (i - 3 / 2 + 0.5))
                offs.append((aw[ind] - 1) * (j - 3 / 2 + 0.5))
        return torch.cat(offs, dim=1) * hard, mask.repeat(1, 9, 1, 1)

    def shape_align(self, x, prob):
        offset, mask = self.shape_align_offsets(prob)
        y = self.dcn(x, offset, mask, self.sd["shape_align.align.weight"], self.sd["shape_align.align.bias"], 1, 1)
        return y + x

    def center_align_offsets(self, bx, by, prob, mean, std):
        """(offset [B,2,H,W] = (dy, dx), mask [B,1,H,W]) of center_align (feturealign_mgpu.py:58-77)."""
        stride = self.conf.feat_stride
        aw = ((self.anchors[:, 2] - self.anchors[:, 0]) / stride).view(1, -1, 1, 1)
        ah = ((self.anchors[:, 3] - self.anchors[:, 1]) / stride).view(1, -1, 1, 1)
        mask, ind, hard = self._top1(prob)
        ox = torch.gather((bx * std[0] + mean[0]) * aw, 1, ind) * hard
        oy = torch.gather((by * std[1] + mean[1]) * ah, 1, ind) * hard
        return torch.cat([oy, ox], dim=1), mask

    def center_align(self, x, bx, by, prob, p, mean, std):
        offset, mask = self.center_align_offsets(bx, by, prob, mean, std)
        y = self.dcn(x, offset, mask, self.sd[p + ".align.weight"], self.sd[p + ".align.bias"], 1, 0)
        return y + x

    @staticmethod
    def papa(feats, att, sizes=(1, 4, 8, 16)):
        """PAPAModule (model/module/attention.py:120-147): attention-weighted pyramid pooling -> [n, c, T]."""
        n, c = feats.shape[:2]
        return torch.cat([F.adaptive_avg_pool2d(feats * att[:, i:i + 1], (s, s)).view(n, c, -1)
                          for i, s in enumerate(sizes)], -1)

    @staticmethod
    def anab_attend(q, key, value, x):
        """softmax(Q K) V + x (attention.py:205-214); q [B,HW,ck], key [B,ck,T], value [B,T,cv]."""
        B, C, H, W = x.shape
        a = torch.softmax(torch.bmm(q, key), dim=-1)
        return torch.bmm(a, value).permute(0, 2, 1).reshape(B, C, H, W) + x

    def anab(self, x, p):
        B, C, H, W = x.shape
        q = self.conv(x, p + ".query_conv").view(B, -1, H * W).permute(0, 2, 1)
        att = torch.sigmoid(self.conv(x, p + ".spatial_conv"))
        key = self.papa(self.conv(x, p + ".key_conv"), att)
        value = self.papa(self.conv(x, p + ".value_conv"), att).permute(0, 2, 1)
        return self.anab_attend(q, key, value, x)

    @staticmethod
    def flatten(t):
        return t.permute(0, 2, 3, 1).contiguous().view(t.shape[0], -1, t.shape[1])

    def forward(self, x):
        """Returns (cls, prob, bbox_2d, bbox_3d, feat_size, rois) like RPN.forward in eval mode."""
        conf = self.conf
        x = x.to(self.dtype)
        B = x.shape[0]
        A, K = self.num_anchors, self.num_classes
        feat = self.dla_seg(x)
        self.taps["feat"] = feat
        H, W = feat.shape[2:]
        cls = self.head(feat, "cls", 3).view(B, K, H * A, W)
        prob = torch.softmax(cls, dim=1)
        fg = (1 - prob[:, 0]).view(B, A, H, W)
        self.taps["fg_prob"] = fg
        means = torch.tensor(conf.bbox_means[0], dtype=self.dtype)
        stds = torch.tensor(conf.bbox_stds[0], dtype=self.dtype)
        feats = self.shape_align(feat, fg) if conf.shape_align else feat
        self.taps["feats_shape"] = feats
        bx, by = self.head(feats, "bbox_x"), self.head(feats, "bbox_y")
        f2 = self.center_align(feats, bx, by, fg, "center_align2d", means[0:2], stds[0:2]) if conf.center_align else feats
        self.taps["feats_align2d"] = f2
        bw, bh = self.head(f2, "bbox_w"), self.head(f2, "bbox_h")
        bx3, by3 = self.head(feats, "bbox_x3d"), self.head(feats, "bbox_y3d")
        f3 = self.center_align(feats, bx3, by3, fg, "center_align3d", means[4:6], stds[4:6]) if conf.center_align else feats
        self.taps["feats_align3d"] = f3
        bw3, bh3, bl3, br3 = (self.head(f3, "bbox_" + n) for n in ("w3d", "h3d", "l3d", "rY3d"))
        fz = f3
        if conf.get("attention") == "ANAB":
            fz = self.act(self.bn(self.anab(f3, "bbox_z3d_gl.0"), "bbox_z3d_gl.1"))
        self.taps["feats_gl"] = fz
        bz3 = self.head(fz, "bbox_z3d")

        def fl(t):
            return self.flatten(t.reshape(B, 1, H * A, W))

        bbox_2d = torch.cat([fl(t) for t in (bx, by, bw, bh)], dim=2)
        bbox_3d = torch.cat([fl(t) for t in (bx3, by3, bz3, bw3, bh3, bl3, br3)], dim=2)
        feat_size = torch.tensor([H, W], dtype=torch.float32)
        return self.flatten(cls), self.flatten(prob), bbox_2d, bbox_3d, feat_size, self.rois(H, W)

    # ----------------------------------------------------------------- decode
    def rois(self, H, W):
        """locate_anchors (lib/rpn_util.py:1329-1398): [(A*H*W), 5], anchor-major then (h, w)."""
        stride = float(self.conf.feat_stride)
        a = self.anchors[:, 0:4].double()
        sx = (torch.arange(W, dtype=torch.float64) * stride).view(1, 1, W)
        sy = (torch.arange(H, dtype=torch.float64) * stride).view(1, H, 1)
        A = a.shape[0]
        x1 = (sx + a[:, 0].view(A, 1, 1)).expand(A, H, W)
        y1 = (sy + a[:, 1].view(A, 1, 1)).expand(A, H, W)
        x2 = (sx + a[:, 2].view(A, 1, 1)).expand(A, H, W)
        y2 = (sy + a[:, 3].view(A, 1, 1)).expand(A, H, W)
        tr = torch.arange(A, dtype=torch.float64).view(A, 1, 1).expand(A, H, W)
        return torch.stack([t.reshape(-1) for t in (x1, y1, x2, y2, tr)], dim=1).float()

    def detect(self, outputs, image_index=0, scale_factor=1.0):
        """im_detect_3d (lib/rpn_util.py:1444-1555) for one image of the batch:
        returns (aboxes_pre_nms [<=3000, 14], keep indices, aboxes_kept)."""
        conf = self.conf
        cls, prob, bbox_2d, bbox_3d, feat_size, rois = outputs
        b2 = bbox_2d[image_index].float().clone()
        b3 = bbox_3d[image_index].float()
        pr = prob[image_index].float()
        rois = rois.float()
        means = torch.tensor(conf.bbox_means[0]).float()
        stds = torch.tensor(conf.bbox_stds[0]).float()
        d3 = [b3[:, i] * stds[4 + i] + means[4 + i] for i in range(7)]
        tracker = rois[:, 4].long()
        src = self.anchors.float()[tracker, 4:]
        widths = rois[:, 2] - rois[:, 0] + 1.0
        heights = rois[:, 3] - rois[:, 1] + 1.0
        ctr_x = rois[:, 0] + 0.5 * widths
        ctr_y = rois[:, 1] + 0.5 * heights
        x3d = d3[0] * widths + ctr_x
        y3d = d3[1] * heights + ctr_y
        z3d = src[:, 0] + d3[2]
        w3d = torch.exp(d3[3]) * src[:, 1]
        h3d = torch.exp(d3[4]) * src[:, 2]
        l3d = torch.exp(d3[5]) * src[:, 3]
        ry3d = src[:, 4] + d3[6]
        coords_3d = torch.stack((x3d, y3d, z3d, w3d, h3d, l3d, ry3d), dim=1)
        dx = b2[:, 0] * stds[0] + means[0]
        dy = b2[:, 1] * stds[1] + means[1]
        dw = b2[:, 2] * stds[2] + means[2]
        dh = b2[:, 3] * stds[3] + means[3]
        pcx = dx * widths + ctr_x
        pcy = dy * heights + ctr_y
        pw = torch.exp(dw) * widths
        ph = torch.exp(dh) * heights
        coords_2d = torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=1)
        coords_2d = coords_2d / scale_factor
        coords_3d[:, 0:2] = coords_3d[:, 0:2] / scale_factor
        cls_pred = torch.argmax(pr[:, 1:], dim=1) + 1
        scores = torch.max(pr[:, 1:], dim=1)[0]
        aboxes = torch.cat((coords_2d, scores[:, None]), dim=1)
        # argsort(-score); ties are implementation-defined in the reference (unstable sort):
        # the oracle breaks them by lower index first, and parity inputs are tie-free.
        order = torch.argsort(-aboxes[:, 4], stable=True)
        n = min(conf.nms_topN_pre, order.shape[0])
        order = order[:n]
        pre = torch.cat((aboxes[order], cls_pred[order].float()[:, None], coords_3d[order],
                         tracker[order].float()[:, None]), dim=1)
        keep = O.gpu_nms(pre[:, 0:5].numpy().astype(np.float32), conf.nms_thres)
        keep = np.asarray(keep, dtype=np.int64)
        return pre, keep, pre[keep]
