"""RPN detector (mirror of model/M3d_inference_align.py): `build(conf, phase)` returns an
nn.Module with the reference's attributes, state-dict keys and forward contract

    train:  cls, prob, bbox_2d, bbox_3d, feat_size
    eval:   cls, prob, bbox_2d, bbox_3d, feat_size, rois

(M3d_inference_align.py:215-313).  In eval mode on a CUDA device the forward runs
through the fused engine (m3dssd_b200/engine.py): hand-written sm_100a kernels
behind the C ABI.  In training mode the modules run one by one under autograd,
with torch's dense convolutions -- as in the reference -- and the C-ABI DCNv2
forward/backward for the deformable ones.  There is no CPU path for DCNv2.
"""
import threading

import numpy as np
import torch
from torch import nn

from ..lib.rpn_util import calc_output_size, flatten_tensor, locate_anchors
from .module.attention import ANAB
from .module.feturealign_mgpu import center_align, shape_align
from .pose_dla_dcn import DeformConv, DLASeg  # noqa: F401  (DeformConv re-exported like the reference)

_ENGINE_LOCK = threading.Lock()  # module level: an nn.Module attribute would break copy.deepcopy / pickling of the net
_REG_HEADS = ["bbox_x", "bbox_y", "bbox_w", "bbox_h", "bbox_x3d", "bbox_y3d", "bbox_z3d", "bbox_w3d", "bbox_h3d",
              "bbox_l3d", "bbox_rY3d"]


def _head(cin, mid, cout, first_kernel=1):
    """conv-BN-LeakyReLU x2 then a 1x1 predictor; indices 0,1,3,4,6 carry parameters (checkpoint keys)."""
    return nn.Sequential(
        nn.Conv2d(cin, mid, first_kernel, padding=first_kernel // 2), nn.BatchNorm2d(mid), nn.LeakyReLU(inplace=True),
        nn.Conv2d(mid, mid, 1), nn.BatchNorm2d(mid), nn.LeakyReLU(inplace=True),
        nn.Conv2d(mid, cout, 1))


class RPN(nn.Module):
    def __init__(self, phase, base, conf):
        super().__init__()
        self.base = base
        self.conf = conf
        self.phase = phase
        self.device = conf.get("device", "cuda")
        self.num_classes = len(conf["lbls"]) + 1
        self.num_anchors = conf["anchors"].shape[0]
        self.anchors = torch.tensor(np.asarray(conf.anchors), dtype=torch.float)
        self.bbox_means = conf.bbox_means[0]
        self.bbox_stds = conf.bbox_stds[0]
        self.base_channels = self.base.out_channels
        self.head_channels = 256
        self.back_bone = conf.back_bone
        self.batch_size = conf.batch_size
        self.align_type = conf.get("align_type", "max")
        self.attention = conf.get("attention", None)
        # "fp32": bf16x3 split arithmetic (reference-accurate); "bf16": throughput mode
        self.precision = conf.get("precision", "fp32")
        self.feat_stride = conf.feat_stride
        self.feat_size = calc_output_size(np.array(conf.crop_size), self.feat_stride)
        self.rois = locate_anchors(conf.anchors, self.feat_size, conf.feat_stride, convert_tensor=True).float()

        ch, mid, A = self.base_channels, self.head_channels, self.num_anchors
        self.cls = _head(ch, mid, A * self.num_classes, first_kernel=3)
        for name in ("bbox_x", "bbox_y", "bbox_w", "bbox_h", "bbox_x3d", "bbox_y3d"):
            setattr(self, name, _head(ch, mid, A))
        if conf.center_align:
            kw = dict(feat_stride=self.feat_stride, feat_size=self.feat_size, kernel_size=1, k=1, thresh=0.5)
            self.center_align2d = center_align(ch, self.anchors, xy_mean=self.bbox_means[0:2],
                                               xy_std=self.bbox_stds[0:2], **kw)
            self.center_align3d = center_align(ch, self.anchors, xy_mean=self.bbox_means[4:6],
                                               xy_std=self.bbox_stds[4:6], **kw)
        else:
            self.center_align2d = self.center_align3d = None
        if conf.shape_align:
            self.shape_align = shape_align(ch, self.anchors, feat_stride=self.feat_stride, feat_size=self.feat_size,
                                           kernel_size=3, k=1, thresh=0.5)
        else:
            self.shape_align = None
        self.bbox_z3d = _head(ch, mid, A)
        if self.attention == "ANAB":
            self.bbox_z3d_gl = nn.Sequential(ANAB(ch, 1), nn.BatchNorm2d(ch), nn.LeakyReLU(inplace=True))
        for name in ("bbox_w3d", "bbox_h3d", "bbox_l3d", "bbox_rY3d"):
            setattr(self, name, _head(ch, mid, A))
        self.softmax = nn.Softmax(dim=1)
        self._feat_size_cache = {}
        self._engines = {}

    # ---------------------------------------------------------------- engine
    def engine(self, batch, height, width, precision=None, **kw):
        """Fused inference engine for a fixed input shape (built on first use, cached)."""
        from ..engine import Engine
        precision = precision or self.precision
        dev = next(self.parameters()).device
        key = (str(dev), batch, height, width, precision) + tuple(sorted(kw.items()))
        with _ENGINE_LOCK:  # nn.DataParallel replicas share this dict by reference and build from threads
            if key not in self._engines:
                self._engines[key] = Engine(self, batch, height, width, precision=precision, **kw)
            return self._engines[key]

    def engine_supported(self):
        """The fused engine covers the trunks the reference ships (BasicBlock = dla34, Bottleneck = dla102,
        model/pose_dla_dcn.py:419-441) with DCNv2 aggregation; anything else runs module by module."""
        from .pose_dla_dcn import BasicBlock, Bottleneck, Tree
        blocks = [m.tree1 for m in self.base.base.modules() if isinstance(m, Tree) and m.levels == 1]
        return bool(self.conf.get("ida_dcnv2", True)) and all(isinstance(b, (BasicBlock, Bottleneck)) for b in blocks)

    def invalidate_engines(self):
        """Call after changing parameters (load_state_dict / optimizer step) so weights are re-packed."""
        self._engines = {}

    def load_state_dict(self, *a, **k):
        self.invalidate_engines()
        return super().load_state_dict(*a, **k)

    def train(self, mode=True):
        if mode:
            self.invalidate_engines()
        return super().train(mode)

    def detect(self, x, scale_factor=1.0, max_out=None):
        """Forward + decode + top-K + batched NMS on the device: (dets [B, max_out, 14], num_kept [B])."""
        if not x.is_cuda:
            raise NotImplementedError("m3dssd_b200 has no CPU path")
        kw = {} if max_out is None else dict(max_out=int(max_out))
        eng = self.engine(x.shape[0], x.shape[2], x.shape[3], **kw)
        kept, num = eng.detect(x.float().contiguous(), scale_factor)
        return kept.clone(), num.clone()  # fresh tensors, like any nn.Module (the engine's buffers are reused per call)

    def detect_images(self, images, size=None, scale_factor=1.0, max_out=None):
        """detect() from the raw frames: a list of uint8 HWC (BGR, as cv2.imread returns) images of different sizes.
        The reference's test loader applies Preprocess(conf.test_scale, means, stds) = zero Padding to `size` +
        Normalize, then BGR->RGB + CHW (lib/dataloader.py:904,942-950; lib/augmentations.py:472-492) on the CPU, one
        image at a time; here the packed bytes are copied once and transformed on the device (bit-identical)."""
        size = size if size is not None else self.conf.get("test_scale", self.conf.get("crop_size"))
        H, W = (int(size[0]), int(size[1]))
        kw = {} if max_out is None else dict(max_out=int(max_out))
        eng = self.engine(len(images), H, W, **kw)
        kept, num = eng.detect(list(images), scale_factor)
        return kept.clone(), num.clone()

    # --------------------------------------------------------------- forward
    def forward(self, x):
        if not self.training and x.is_cuda:
            if self.engine_supported():
                return self._forward_engine(x)
            with torch.no_grad():
                return self._forward_modules(x)
        return self._forward_modules(x)

    def _forward_engine(self, x):
        B, _, H, W = x.shape
        eng = self.engine(B, H, W)
        # the engine returns views of its persistent buffers (overwritten by the next call); the nn.Module surface
        # hands out fresh tensors like the reference's forward does
        cls, prob, bbox_2d, bbox_3d = (t.clone() for t in eng.forward(x.float().contiguous()))
        feat_size = eng.feat_size.clone()
        if self.feat_size[0] != eng.Hf or self.feat_size[1] != eng.Wf or self.rois.device != x.device:
            self.feat_size = [eng.Hf, eng.Wf]
            self.rois = locate_anchors(self.conf.anchors, self.feat_size, self.feat_stride,
                                       convert_tensor=True).float().to(x.device)
        return cls, prob, bbox_2d, bbox_3d, feat_size, self.rois

    def _forward_modules(self, x):
        B = x.size(0)
        A, K = self.num_anchors, self.num_classes
        x = self.base(x)
        Hf, Wf = x.shape[2:]
        cls = self.cls(x).contiguous().view(B, K, Hf * A, Wf)  # (contiguous: the native training path is channels_last)
        prob = self.softmax(cls)
        fg_prob = (1 - prob.detach()[:, 0]).view(B, A, Hf, Wf)
        feats = self.shape_align(x, fg_prob) if self.shape_align is not None else x
        out = {}
        out["bbox_x"], out["bbox_y"] = self.bbox_x(feats), self.bbox_y(feats)
        f2d = feats
        if self.center_align2d is not None:
            f2d = self.center_align2d(feats, out["bbox_x"].detach(), out["bbox_y"].detach(), fg_prob)
        out["bbox_w"], out["bbox_h"] = self.bbox_w(f2d), self.bbox_h(f2d)
        out["bbox_x3d"], out["bbox_y3d"] = self.bbox_x3d(feats), self.bbox_y3d(feats)
        f3d = feats
        if self.center_align3d is not None:
            f3d = self.center_align3d(feats, out["bbox_x3d"].detach(), out["bbox_y3d"].detach(), fg_prob)
        for n in ("bbox_w3d", "bbox_h3d", "bbox_l3d", "bbox_rY3d"):
            out[n] = getattr(self, n)(f3d)
        fz = self.bbox_z3d_gl(f3d) if self.attention == "ANAB" else f3d
        out["bbox_z3d"] = self.bbox_z3d(fz)
        flat = {n: flatten_tensor(t.reshape(B, 1, Hf * A, Wf)) for n, t in out.items()}
        bbox_2d = torch.cat([flat[n] for n in _REG_HEADS[:4]], dim=2)
        bbox_3d = torch.cat([flat[n] for n in _REG_HEADS[4:]], dim=2)
        key = (Hf, Wf, str(x.device))  # cached: a host->device copy per call would also break CUDA-graph capture
        if key not in self._feat_size_cache:
            self._feat_size_cache[key] = torch.tensor([Hf, Wf], dtype=torch.float, device=x.device)
        feat_size = self._feat_size_cache[key]
        cls, prob = flatten_tensor(cls), flatten_tensor(prob)
        if self.training:
            return cls, prob, bbox_2d, bbox_3d, feat_size
        if self.feat_size[0] != Hf or self.feat_size[1] != Wf or self.rois.device != x.device:
            self.feat_size = [Hf, Wf]
            self.rois = locate_anchors(self.conf.anchors, self.feat_size, self.feat_stride,
                                       convert_tensor=True).float().to(x.device)
        return cls, prob, bbox_2d, bbox_3d, feat_size, self.rois


def build(conf, phase="train"):
    train = phase.lower() == "train"
    base_name = conf.back_bone
    if base_name[0:3] != "dla":
        raise NotImplementedError
    base = DLASeg(base_name, pretrained=conf.pre_train, down_ratio=conf.feat_stride, final_kernel=1, last_level=5,
                  head_conv=256, conf=conf)
    rpn_net = RPN(phase, base, conf)
    rpn_net.train() if train else rpn_net.eval()
    return rpn_net
