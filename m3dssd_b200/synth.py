"""Synthetic KITTI-shaped configuration, weights and inputs.

There is no dataset or checkpoint in the build/bench environment, so the
benchmark and the parity tests use: the reference's own anchor recipe
(lib/rpn_util.py:39-52,167-183; scripts/config/kitti_3d_base.py:75-79,130-132)
with fixed 3-D priors, zero/one bbox statistics, and deterministic random
weights in which the (zero-initialised, model/DCNv2/dcn_v2.py:60-62)
conv_offset_mask layers are randomised so the deformable gather is exercised.
"""
import contextlib
import copy
import math

import numpy as np
import torch
import torch.nn.functional as F


class Conf(dict):
    """Attribute dict (the reference uses easydict.EasyDict, absent from this image)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def anchor_center(w, h, stride):
    # lib/rpn_util.py:167-183
    a = np.zeros([4], dtype=np.float32)
    a[0] = -w / 2 + (stride - 1) / 2
    a[1] = -h / 2 + (stride - 1) / 2
    a[2] = w / 2 + (stride - 1) / 2
    a[3] = h / 2 + (stride - 1) / 2
    return a


def make_anchors(test_scale=(384, 1280), percent_anc_h=(0.0625, 0.75), n_scales=12, ratios=(0.5, 1.0, 1.5), stride=8):
    min_h, max_h = test_scale[0] * percent_anc_h[0], test_scale[0] * percent_anc_h[1]
    base = (max_h / min_h) ** (1 / (n_scales - 1))
    scales = np.array([min_h * (base ** i) for i in range(n_scales)])
    anchors = np.zeros([n_scales * len(ratios), 9], dtype=np.float32)
    k = 0
    for s in scales:
        for r in ratios:
            anchors[k, 0:4] = anchor_center(s * r, s, stride)
            # fixed 3-D priors: z = pinhole depth of a 1.5 m object at KITTI focal ~721 px
            anchors[k, 4:9] = [721.0 * 1.5 / s, 1.6, 1.5, 3.9, 0.0]
            k += 1
    return anchors, scales, np.array(ratios)


def make_conf(attention=None, center_align=True, shape_align=True, back_bone="dla34", batch_size=8,
              crop_size=(384, 1280), device="cpu"):
    c = Conf()
    c.model = "M3d_inference_align"
    c.ida_dcnv2 = True
    c.attention = attention
    c.center_align = center_align
    c.shape_align = shape_align
    c.back_bone = back_bone
    c.pre_train = False
    c.feat_stride = 8
    c.has_3d = True
    c.test_scale = list(crop_size)
    c.crop_size = list(crop_size)
    c.image_means = [0.485, 0.456, 0.406]  # scripts/config/kitti_3d_base.py:42-43
    c.image_stds = [0.229, 0.224, 0.225]
    c.lbls = ["Car", "Pedestrian", "Cyclist"]
    c.ilbls = ["Van", "ignore"]
    c.batch_size = batch_size
    c.nms_topN_pre = 3000
    c.nms_topN_post = 40
    c.nms_thres = 0.4
    c.clip_boxes = False
    c.rng_seed = 2
    c.cuda_seed = 2
    anchors, scales, ratios = make_anchors(stride=c.feat_stride)
    c.anchors = anchors
    c.anchor_scales = scales
    c.anchor_ratios = ratios
    c.bbox_means = np.zeros([1, 11], dtype=np.float32)
    c.bbox_stds = np.ones([1, 11], dtype=np.float32)
    c.device = device
    return c


# Conditioning of the synthetic network (see randomize_weights).  A He-initialised BatchNorm + LeakyReLU network is
# chaotic: every conv-BN-activation layer multiplies the ratio perturbation / signal fluctuation by
# sqrt(E[phi'^2] / Var[phi(z)]) ~ 1.21 (the BatchNorm mean subtraction), i.e. x10^2..10^4 over the ~70 layers between
# the stem and the boxes -- fp32 round-off became 1.5e-2 rms at 384x1280, bf16 rounding 40 %.  On top of that the
# deformable offsets were driven entirely by rough random features and the fg probabilities crowded the 0.5 hard mask
# of the align modules.  Trained detectors are not like that, and neither is this network (measured with
# oracle/ref_model.py on the CPU: fp32 vs fp64 3.5e-6 rms at the boxes at 384x1280; a bf16-rounding emulation of the
# same forward 0.8e-2 on the class logits, 1.4e-2 on the aggregated feature map):
BN_BIAS_MEAN = 1.0       # BatchNorm beta ~ N(1, 0.1^2): ~16 % of the pre-activations in the negative branch (gain ~1.03 / layer)
RES_BRANCH_GAIN = 0.25   # BasicBlock bn2.weight scale: block = skip + 0.25 * branch (zero-gamma-style residuals)
OFFSET_DATA_SIGMA = 0.5  # px: data-dependent part of every DCN offset; the rest of offset_sigma_px is a fixed per-tap bias
CLS_LOGIT_GAIN = 1.0     # class-logit spread (He init)
CENTER_DELTA_GAIN = 0.1  # bbox_x/y/x3d/y3d outputs ~0.1 (they are multiplied by anchor sizes of up to 54 feature px)
FG_FRACTION = 0.002      # anchors with fg prob > 0.5 (~1/3 of the positions then have a foreground top-1 anchor)


@torch.no_grad()
def randomize_weights(model, seed=2, offset_sigma_px=2.0, fg_fraction=FG_FRACTION, calibrate=True):
    """Deterministic synthetic weights for a reference-shaped RPN (reference modules or ours).

    Draws every tensor from its own generator keyed by the parameter name, so the
    result does not depend on module construction order; then (calibrate=True) runs
    calibrate_statistics.  `model` must be one of OUR modules for calibration (the reference's
    modules take the returned state_dict through load_state_dict).
    """
    sd = model.state_dict()
    mods = dict(model.named_modules())
    new = {}
    for name in sorted(sd.keys()):
        t = sd[name]
        g = torch.Generator().manual_seed((hash_name(name) + seed) % (2 ** 31))
        mod = mods.get(name.rsplit(".", 1)[0])
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            new[name] = t
            continue
        if isinstance(mod, torch.nn.BatchNorm2d):
            if leaf == "running_mean":
                v = torch.randn(t.shape, generator=g) * 0.1
            elif leaf == "running_var":
                v = torch.rand(t.shape, generator=g) + 0.5
            elif leaf == "weight":
                v = torch.rand(t.shape, generator=g) + 0.5
                if name.endswith((".bn2.weight", ".bn3.weight")) and ".tree" in name:
                    v = v * RES_BRANCH_GAIN
            else:
                v = torch.randn(t.shape, generator=g) * 0.1 + BN_BIAS_MEAN
        elif "conv_offset_mask" in name and leaf == "weight":
            fan_in = t.shape[1] * t.shape[2] * t.shape[3]
            # activations are O(1); offsets ~ N(0, sigma^2) need w ~ sigma / sqrt(fan_in)
            v = torch.randn(t.shape, generator=g) * (offset_sigma_px / math.sqrt(fan_in))
            v[18:] *= 0.5 / offset_sigma_px  # mask logits ~ N(0, 0.5^2): sigmoid spread over ~(0.25, 0.75)
        elif "conv_offset_mask" in name:
            v = torch.randn(t.shape, generator=g) * 0.2
        elif isinstance(mod, torch.nn.ConvTranspose2d):
            # bilinear kernels from fill_up_weights (model/pose_dla_dcn.py:459-468); trainable, so perturbed
            v = t.float() * (1.0 + 0.05 * torch.randn(t.shape, generator=g))
        elif t.dim() == 4:
            fan_in = t.shape[1] * t.shape[2] * t.shape[3]
            v = torch.randn(t.shape, generator=g) * math.sqrt(2.0 / fan_in)
            if name == "cls.6.weight":
                v = v * CLS_LOGIT_GAIN
            elif name in ("bbox_x.6.weight", "bbox_y.6.weight", "bbox_x3d.6.weight", "bbox_y3d.6.weight"):
                v = v * CENTER_DELTA_GAIN
        else:
            v = torch.randn(t.shape, generator=g) * 0.1
        new[name] = v.to(t.dtype)
    model.load_state_dict(new)
    if calibrate:
        calibrate_statistics(model, seed, offset_sigma_px, fg_fraction)
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


@contextlib.contextmanager
def _surrogate_dcn():
    """Weight synthesis only.  DCNv2 has no CPU implementation (here as in the reference), so while
    activation statistics are measured on the CPU the operator is stood in for by the plain
    convolution it reduces to at zero offsets and mask 0.5 -- its state at initialisation
    (model/DCNv2/dcn_v2.py:60-62).  Second-order statistics are all that is needed."""
    from .model.DCNv2 import dcn_v2

    def plain(self, input):
        return F.conv2d(input, self.weight * 0.5, self.bias, stride=self.stride, padding=self.padding,
                        dilation=self.dilation)

    def v2_forward(self, input, offset, mask):
        return plain(self, input)

    def dcn_forward(self, input):
        self.conv_offset_mask(input)  # so the calibration hook sees its output
        return plain(self, input)

    saved = dcn_v2.DCNv2.forward, dcn_v2.DCN.forward
    dcn_v2.DCNv2.forward, dcn_v2.DCN.forward = v2_forward, dcn_forward
    try:
        yield
    finally:
        dcn_v2.DCNv2.forward, dcn_v2.DCN.forward = saved


@torch.no_grad()
def calibrate_statistics(model, seed=2, offset_sigma_px=2.0, fg_fraction=FG_FRACTION, crop=(96, 320), batch=2):
    """Make the synthetic network well-conditioned: one fp64 CPU pass over a seeded batch sets every
    BatchNorm's running statistics to the statistics it actually sees (so activations stay O(1) at
    any depth), rescales each conv_offset_mask so offsets have sigma ~ offset_sigma_px and mask logits
    sigma ~ 0.5, and shifts the background logit so ~fg_fraction of the anchors are foreground.
    fp64 keeps the result identical across host CPUs after rounding to fp32."""
    m = copy.deepcopy(model).double().eval()
    hooks = []
    fixes = {}

    def bn_pre(mod, inp):
        x = inp[0]
        mod.running_mean.copy_(x.mean(dim=(0, 2, 3)))
        mod.running_var.copy_(x.var(dim=(0, 2, 3), unbiased=False).clamp_min(1e-6))

    def om_post(name):
        def hook(mod, inp, out):
            # offsets = fixed per-tap bias pattern + a data-dependent part of sigma OFFSET_DATA_SIGMA (total sigma
            # ~ offset_sigma_px); mask logits = bias N(0, 0.4^2) + data part of sigma 0.3
            n_off = out.shape[1] // 3 * 2
            data = out - mod.bias.view(1, -1, 1, 1)
            sig_d = min(OFFSET_DATA_SIGMA, offset_sigma_px)
            s_off = sig_d / float(data[:, :n_off].std().clamp_min(1e-9))
            s_msk = 0.3 / float(data[:, n_off:].std().clamp_min(1e-9))
            scale = torch.cat([torch.full((n_off,), s_off), torch.full((out.shape[1] - n_off,), s_msk)]).double()
            mod.weight.mul_(scale.view(-1, 1, 1, 1))
            gb = torch.Generator().manual_seed((hash_name(name) + seed + 7) % (2 ** 31))
            pat = torch.randn(out.shape[1], generator=gb, dtype=torch.float64)
            sig_b = math.sqrt(max(offset_sigma_px ** 2 - sig_d ** 2, 0.0))
            mod.bias.copy_(torch.cat([pat[:n_off] * sig_b, pat[n_off:] * 0.4]))
            fixes[name] = mod
        return hook

    def cls_post(mod, inp, out):
        B, KA, H, W = out.shape
        K = m.num_classes
        lo = out.view(B, K, KA // K, H, W)
        t = torch.logsumexp(lo[:, 1:], dim=1) - lo[:, 0]
        shift = torch.quantile(t.flatten()[:: max(1, t.numel() // 200000)], 1.0 - fg_fraction)
        mod.bias[: KA // K] += shift

    for name, mod in m.named_modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            hooks.append(mod.register_forward_pre_hook(bn_pre))
        elif name.endswith("conv_offset_mask"):
            hooks.append(mod.register_forward_hook(om_post(name)))
    hooks.append(m.cls[6].register_forward_hook(cls_post))
    g = torch.Generator().manual_seed(seed + 12345)
    x = torch.randn(batch, 3, crop[0], crop[1], generator=g, dtype=torch.float64)
    with _surrogate_dcn():
        m(x)
    for h in hooks:
        h.remove()
    src = m.state_dict()
    dst = model.state_dict()
    for k, v in src.items():
        if k.endswith(("running_mean", "running_var")) or "conv_offset_mask" in k or k == "cls.6.bias":
            dst[k].copy_(v.to(dst[k].dtype))


def hash_name(name):
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % 1000003
    return h


def make_images_u8(batch, crop_size=(384, 1280), seed=0):
    """uint8 HWC (BGR) images as cv2.imread returns them: the input of the device-side pipeline."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, crop_size[0], crop_size[1], 3), generator=g, dtype=torch.uint8)


def make_images(batch, crop_size=(384, 1280), seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, crop_size[0], crop_size[1], generator=g)


def loss_conf(conf):
    """The loss hyper-parameters of scripts/config/kitti_3d_base.py:89-142 on top of make_conf()."""
    conf.box_samples, conf.fg_fraction = 0.20, 0.20
    conf.bg_thresh_lo, conf.bg_thresh_hi, conf.fg_thresh, conf.ign_thresh, conf.best_thresh = 0, 0.5, 0.5, 0.5, 0.35
    conf.hard_negatives, conf.focal_loss = True, 0
    conf.cls_2d_lambda, conf.iou_2d_lambda, conf.bbox_2d_lambda, conf.bbox_3d_lambda = 1, 1, 0, 1
    conf.bbox_3d_proj_lambda, conf.bbox_3d_iou_lambda = 0.0, 0
    conf.min_gt_vis, conf.min_gt_h, conf.max_gt_h = 0.65, 0, 10e10
    return conf


def make_targets(conf, batch, feat_hw=None, seed=0, fg_per_image=300, ign_fraction=0.05, empty_images=(), box_sigma=0.1):
    """Synthetic training targets in the reference's `imobjs` layout (lib/dataloader.py:959-982):
    labels_fg / labels_bg / labels_ign [B, M] bool, labels [B, M] int64 (0 background, 1..3 classes, 3000 ignored),
    bbox_2d [B, M, 4], bbox_3d [B, M, 7] regression targets, meta.rois [B, M, 5], meta.any_val [B], meta.p2 [B, 4, 4].
    Images listed in `empty_images` have no foreground and no ignored anchors (the loss's all-background branch)."""
    from .lib.rpn_util import locate_anchors
    g = torch.Generator().manual_seed(seed)
    A = conf.anchors.shape[0]
    if feat_hw is None:
        feat_hw = (conf.crop_size[0] // conf.feat_stride, conf.crop_size[1] // conf.feat_stride)
    Hf, Wf = feat_hw
    M = A * Hf * Wf
    labels = torch.zeros(batch, M, dtype=torch.long)
    fg = torch.zeros(batch, M, dtype=torch.bool)
    ign = torch.zeros(batch, M, dtype=torch.bool)
    for b in range(batch):
        if b in empty_images:
            continue
        perm = torch.randperm(M, generator=g)
        nf = min(fg_per_image, M // 8)
        ni = int(M * ign_fraction)
        fg[b, perm[:nf]] = True
        ign[b, perm[nf:nf + ni]] = True
        labels[b, perm[:nf]] = torch.randint(1, len(conf.lbls) + 1, (nf,), generator=g)
        labels[b, perm[nf:nf + ni]] = 3000
    bg = ~fg & ~ign
    rois = torch.from_numpy(locate_anchors(conf.anchors, (Hf, Wf), conf.feat_stride)).float()
    p2 = torch.eye(4).repeat(batch, 1, 1)
    p2[:, 0, 0] = p2[:, 1, 1] = 721.5
    p2[:, 0, 2], p2[:, 1, 2] = 609.5, 172.8
    return {
        "labels_fg": fg, "labels_bg": bg, "labels_ign": ign, "labels": labels,
        "bbox_2d": torch.randn(batch, M, 4, generator=g) * box_sigma, "bbox_3d": torch.randn(batch, M, 7, generator=g) * 0.3,
        "meta": {"rois": rois.unsqueeze(0).repeat(batch, 1, 1), "any_val": torch.ones(batch, dtype=torch.bool), "p2": p2},
    }


@torch.no_grad()
def condition_for_training(net, box_gain=0.05, center_gain=0.3):
    """Training-mode conditioning of the synthetic network: RPN_3D_loss_smp's IoU term is -log(IoU) with no epsilon
    (lib/loss/rpn_3d.py:1338), so ONE sampled foreground anchor whose decoded box misses its target makes the loss inf --
    a freshly initialised bbox_w / bbox_h head (log-size deltas ~ N(0, 2^2)) does that at once, in the reference as
    here.  Scale the last layers of the 2D box heads so every decoded box overlaps its target, as in a network a few
    hundred iterations into training.  Not used by the inference goldens."""
    for name, gain in (("bbox_w", box_gain), ("bbox_h", box_gain), ("bbox_x", center_gain), ("bbox_y", center_gain)):
        last = getattr(net, name)[-1]
        last.weight.mul_(gain)
        if last.bias is not None:
            last.bias.mul_(gain)
    return net


def make_gts(conf, n_val=8, n_ign=2, seed=0, image_hw=None):
    """Seeded KITTI-like annotations of one image in the form Dataset._targets hands to compute_targets
    (lib/dataloader.py:1060-1084): gts_val / gts_ign [n, 4] boxes (x1, y1, x2, y2) whose sizes follow the anchors',
    box_lbls (class index 1..len(lbls)), gts_3d [n, 7] = projected centre, depth, w3d, h3d, l3d, rotY."""
    rng = np.random.default_rng(seed)
    H, W = image_hw if image_hw is not None else conf.crop_size

    def boxes(n):
        h = np.exp(rng.uniform(np.log(20.0), np.log(0.7 * H), n))
        w = h * rng.uniform(0.4, 1.8, n)
        cx, cy = rng.uniform(0, W, n), rng.uniform(0.3 * H, 0.9 * H, n)
        return np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1)
    val = boxes(n_val)
    g3d = np.stack([(val[:, 0] + val[:, 2]) / 2 + rng.normal(0, 3, n_val), (val[:, 1] + val[:, 3]) / 2 + rng.normal(0, 3, n_val),
                    rng.uniform(4, 60, n_val), rng.uniform(0.5, 2.0, n_val), rng.uniform(1.2, 2.0, n_val),
                    rng.uniform(0.8, 4.5, n_val), rng.uniform(-3.1, 3.1, n_val)], axis=1) if n_val else np.zeros((0, 7))
    return {"gts_val": val, "gts_ign": boxes(n_ign), "box_lbls": rng.integers(1, len(conf.lbls) + 1, n_val), "gts_3d": g3d}
