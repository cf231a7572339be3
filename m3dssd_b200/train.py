"""Native training path (BASELINE config 4: kitti_3d_base train step, forward + backward + SGD).

The reference trains through torch autograd with cuDNN convolutions (scripts/train_rpn_3d.py:204-218,
lib/core.py:73-83).  Here every nn.Conv2d of the detector -- trunk, heads, the DCNs' offset/mask predictors -- runs, in
training mode, through the C ABI in both directions:

    forward   m3d_conv2d_nhwc      (tcgen05 implicit GEMM, bf16 operands, fp32 accumulation, bias fused)
    dgrad     m3d_conv2d_nhwc      on the output gradient with the flipped / transposed weights (a strided convolution's
                                   gradient is taken on the zero-inserted gradient map)
    wgrad     m3d_conv2d_wgrad     (tcgen05, MN-major operands straight from the NHWC tensors, deterministic split-K)

and the deformable layers through m3d_dcn_v2_forward / m3d_dcn_v2_backward (model/DCNv2/dcn_v2_func.py).  Activations are
bf16 in channels_last (= NHWC) storage, so torch's elementwise / BatchNorm / pooling kernels and the C-ABI kernels share
buffers without layout copies; master weights, BatchNorm statistics, gradients of the parameters and the optimizer
state stay fp32 (mixed precision).  The depthwise ConvTranspose2d up-sampling of IDAUp runs through m3d_upsample_add_nhwc /
m3d_upsample_backward.  No cuDNN / cuBLAS convolution is called: `enable(net)` also switches cuDNN off for the
process so that a stray nn.functional.conv2d cannot silently fall back to it.
"""
import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function

from . import ops


def _cl(t):
    """bf16, channels_last storage."""
    if t.dtype != torch.bfloat16:
        t = t.to(torch.bfloat16)
    return t.contiguous(memory_format=torch.channels_last)


def _pad_channels(t_nhwc, mult):
    c = t_nhwc.shape[-1]
    cp = (c + mult - 1) // mult * mult
    if cp == c:
        return t_nhwc.contiguous(), c
    out = torch.zeros(t_nhwc.shape[:-1] + (cp,), dtype=t_nhwc.dtype, device=t_nhwc.device)
    out[..., :c] = t_nhwc
    return out, c


def _pack(w, cin_pad):
    """[Cout, Cin, R, S] fp32 -> bf16 [Cout, R*S*cin_pad] (tap-major, channel-minor), zero channel padding."""
    co, ci, r, s = w.shape
    wp = w.permute(0, 2, 3, 1)
    if cin_pad != ci:
        wp = F.pad(wp, (0, cin_pad - ci))
    return wp.reshape(co, r * s * cin_pad).to(torch.bfloat16).contiguous()


def _kmult(c):
    return 64 if c % 64 == 0 else (32 if c % 32 == 0 else 16)


class _ConvFn(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad):
        co, ci, k, _ = weight.shape
        xn, _ = _pad_channels(_cl(x).permute(0, 2, 3, 1), 16)  # NHWC view of the channels_last storage
        N, H, W, cip = xn.shape
        P, Q = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        cop = (co + 7) // 8 * 8
        out = torch.empty(N, P, Q, cop, dtype=torch.bfloat16, device=x.device)
        b = bias.detach().float().contiguous() if bias is not None else None
        ops.conv2d_nhwc([(xn, 0, cip)], _pack(weight.detach(), cip), out, R=k, S=k, stride=stride, pad=pad, Cout=co, bias=b,
                        slope=1.0)
        ctx.save_for_backward(xn, weight)
        ctx.cfg = (stride, pad, ci, bias is not None)
        return out[..., :co].permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        xn, weight = ctx.saved_tensors
        stride, pad, ci, has_bias = ctx.cfg
        co, _, k, _ = weight.shape
        N, H, W, cip = xn.shape
        gyn, _ = _pad_channels(_cl(gy).permute(0, 2, 3, 1), 16)
        P, Q, cop = gyn.shape[1], gyn.shape[2], gyn.shape[3]
        dw = ops.conv2d_wgrad(xn, gyn, ci, co, k, k, stride, pad) if ctx.needs_input_grad[1] else None
        db = ops.channel_sum(gyn, co) if has_bias and ctx.needs_input_grad[2] else None
        dx = None
        if ctx.needs_input_grad[0]:
            # dx = conv(U, flip(w)^T, pad k-1-pad), U = gy with stride-1 zeros inserted, sized so that the result is H x W
            if stride == 1 and P == H - k + 2 * pad + 1:
                u = gyn
            else:
                u = torch.zeros(N, H - k + 2 * pad + 1, W - k + 2 * pad + 1, cop, dtype=torch.bfloat16, device=gy.device)
                u[:, ::stride, ::stride][:, :P, :Q] = gyn
            wt = weight.detach().flip(2, 3).transpose(0, 1)  # [Cin, Cout, k, k]
            cin_out = (ci + 7) // 8 * 8
            dxn = torch.empty(N, H, W, cin_out, dtype=torch.bfloat16, device=gy.device)
            ops.conv2d_nhwc([(u, 0, cop)], _pack(wt, cop), dxn, R=k, S=k, stride=1, pad=k - 1 - pad, Cout=ci, slope=1.0)
            dx = dxn[..., :ci].permute(0, 3, 1, 2)
        return dx, dw, db, None, None


class _DCNFn(Function):
    """DCNv2 in the training graph on the NHWC / bf16 tensors themselves: forward = the fused deformable kernel
    (m3d_conv2d_nhwc with offsets / mask: bilinear gather + tcgen05 GEMM, no NCHW round trip), backward =
    m3d_dcn_v2_backward in its bf16 tensor-core mode.  `mask` is the modulation after the sigmoid, as DCNv2.forward
    receives it (model/DCNv2/dcn_v2.py:39-41)."""

    @staticmethod
    def forward(ctx, x, offset, mask, weight, bias, stride, pad):
        co, ci, k, _ = weight.shape
        xn = _cl(x).permute(0, 2, 3, 1)
        om = torch.cat([offset, mask], dim=1).permute(0, 2, 3, 1).float().contiguous()  # [N, P, Q, 3 k^2]
        N, P, Q, _ = om.shape
        out = torch.empty(N, P, Q, co, dtype=torch.bfloat16, device=x.device)
        ops.conv2d_nhwc([(xn, 0, ci)], _pack(weight.detach(), ci), out, R=k, S=k, stride=stride, pad=pad, Cout=co,
                        bias=bias.detach().float().contiguous(), slope=1.0, om=om, sigmoid_mask=False)
        ctx.save_for_backward(x, offset, mask, weight, bias)
        ctx.cfg = (stride, pad)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        from ._lib import M3D_BF16
        x, offset, mask, weight, bias = ctx.saved_tensors
        stride, pad = ctx.cfg
        gi, go, gm, gw, gb = ops.dcn_v2_backward(x, offset, mask, weight, gy, stride, pad, 1, 1, precision=M3D_BF16)
        return gi.to(x.dtype), go.to(offset.dtype), gm.to(mask.dtype), gw, gb, None, None


def _dcn_op(self):
    """Replacement of DCNv2._op for modules in the native training graph."""
    ok = (self.training and self.dilation == 1 and self.deformable_groups == 1 and self.in_channels % 64 == 0 and
          self.out_channels % 8 == 0 and self.kernel_size == (3, 3))
    if not ok:
        return type(self)._op(self)
    return lambda x, offset, mask, weight, bias: _DCNFn.apply(x, offset, mask, weight, bias, self.stride, self.padding)


class _UpsampleFn(Function):
    """IDAUp.up_i: depthwise ConvTranspose2d(2f, stride f, pad f // 2, groups = C, no bias) (model/pose_dla_dcn.py:536-539):
    forward m3d_upsample_add_nhwc (no skip), backward m3d_upsample_backward (input gradient = depthwise strided
    correlation, weight gradient reduced deterministically)."""

    @staticmethod
    def forward(ctx, x, weight, f):
        xn = _cl(x).permute(0, 2, 3, 1)
        N, H, W, C = xn.shape
        wt = ops.pack_upsample_weight(weight)  # tap-major [(2f)^2, C] fp32
        out = torch.empty(N, H * f, W * f, C, dtype=torch.bfloat16, device=x.device)
        ops.upsample_add(xn, wt, None, out, f)
        ctx.save_for_backward(xn, wt)
        ctx.f = f
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        xn, wt = ctx.saved_tensors
        gx, gw = ops.upsample_backward(_cl(gy).permute(0, 2, 3, 1).contiguous(), xn, wt, ctx.f)
        k = 2 * ctx.f
        return gx.permute(0, 3, 1, 2), gw.t().reshape(-1, 1, k, k).contiguous(), None


def _conv_forward(self, x):
    if self.training and x.is_cuda and self.groups == 1 and self.dilation == (1, 1) and \
            self.kernel_size[0] == self.kernel_size[1] and self.stride[0] == self.stride[1] and \
            self.padding[0] == self.padding[1] and isinstance(self.padding[0], int):
        return _ConvFn.apply(x, self.weight, self.bias, self.stride[0], self.padding[0])
    return nn.Conv2d.forward(self, x)


def _upsample_forward(self, x):
    if self.training and x.is_cuda and self.groups == self.in_channels == self.out_channels and self.bias is None:
        return _UpsampleFn.apply(x, self.weight, self.stride[0])
    return nn.ConvTranspose2d.forward(self, x)


def enable(net, dcn_precision="bf16"):
    """Route every nn.Conv2d / depthwise nn.ConvTranspose2d of `net` through the native kernels while it is in training
    mode (eval mode keeps using the fused engine).  Returns the number of patched modules."""
    import types
    n = 0
    from .model.DCNv2.dcn_v2 import DCNv2
    for m in net.modules():
        if isinstance(m, DCNv2):
            m.precision = dcn_precision  # tensor-core arithmetic for the deformable layers of the training graph
            m._op = types.MethodType(_dcn_op, m)
        if type(m) is nn.Conv2d:
            m.forward = types.MethodType(_conv_forward, m)
            n += 1
        elif type(m) is nn.ConvTranspose2d:
            m.forward = types.MethodType(_upsample_forward, m)
            n += 1
    torch.backends.cudnn.enabled = False  # nothing may fall back to a cuDNN convolution behind our back
    net._native_training = True
    return n


def surrogate_targets(conf, batch, device, seed=0, fg_per_image=300):
    """Synthetic training targets shaped like the reference's (lib/dataloader.py:959-982): labels [B, M] (0 background,
    1..3 classes, -1 ignored), regression targets for the foreground anchors, ~300 foreground anchors per image."""
    g = torch.Generator().manual_seed(seed)
    A = conf.anchors.shape[0]
    Hf, Wf = conf.crop_size[0] // conf.feat_stride, conf.crop_size[1] // conf.feat_stride
    M = A * Hf * Wf
    labels = torch.zeros(batch, M, dtype=torch.long)
    labels[torch.rand(batch, M, generator=g) < 0.7] = -1  # most background anchors are not sampled (ignored)
    for b in range(batch):
        idx = torch.randperm(M, generator=g)[:fg_per_image]
        labels[b, idx] = torch.randint(1, len(conf.lbls) + 1, (fg_per_image,), generator=g)
    t2 = torch.randn(batch, M, 4, generator=g) * 0.3
    t3 = torch.randn(batch, M, 7, generator=g) * 0.3
    return labels.to(device), t2.to(device), t3.to(device)


def surrogate_loss(cls, bbox_2d, bbox_3d, labels, t2, t3):
    """Detection loss with the structure of RPN_3D_loss_smp (lib/loss/rpn_3d.py:811-994): softmax cross-entropy over the
    sampled anchors + smooth-L1 on the 2D / 3D regressions of the foreground anchors (the reference's per-image python
    loops, OHEM sort and IoU weighting are out of scope here: SURVEY 8f rank 3)."""
    K = cls.shape[-1]
    ce = F.cross_entropy(cls.float().reshape(-1, K), labels.reshape(-1), ignore_index=-1)
    fg = (labels > 0).unsqueeze(-1).float()  # masked means instead of boolean indexing: static shapes (CUDA-graph safe)
    nfg = fg.sum().clamp(min=1.0)
    l2 = (F.smooth_l1_loss(bbox_2d.float(), t2, reduction="none") * fg).sum() / (nfg * bbox_2d.shape[-1])
    l3 = (F.smooth_l1_loss(bbox_3d.float(), t3, reduction="none") * fg).sum() / (nfg * bbox_3d.shape[-1])
    return ce + l2 + l3


class TrainStep:
    """One training iteration of scripts/train_rpn_3d.py:196-218 on device tensors: forward, loss, backward, SGD.

    graph=True: after `warmup` eager iterations (momentum buffers, workspaces and kernel attributes exist) the whole
    iteration -- ~5000 kernel launches, most of them tiny -- is captured once into a CUDA graph and replayed: the step
    is then bound by the device, not by Python / ctypes / dispatcher time (measured: 65 ms of host time per eager step
    against 44 ms of device time).  Inputs are copied into static buffers; the returned loss is a static tensor."""

    def __init__(self, net, conf, lr=0.004, momentum=0.9, weight_decay=0.0005, native=True, graph=False, warmup=3,
                 criterion=None):
        """criterion: None = surrogate_loss(cls, bbox_2d, bbox_3d, labels, t2, t3) on the three target tensors passed to
        __call__; or an RPN_3D_loss_smp instance (m3dssd_b200.lib.loss.rpn_3d: the reference's loss with static shapes),
        __call__ then takes the reference's `imobjs` target dict (tensors on the device)."""
        self.net, self.conf = net, conf
        if native:
            enable(net)
        self.opt = torch.optim.SGD(net.parameters(), lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.use_graph, self.warmup = graph, warmup
        self.criterion = criterion
        self._n, self._graph, self._static = 0, None, None

    def _iteration(self, images, *targets):
        self.net.train()
        cls, prob, bbox_2d, bbox_3d, feat_size = self.net(images)
        if self.criterion is not None:
            loss, self.stats = self.criterion(cls, prob, bbox_2d, bbox_3d, targets[0], feat_size)
        else:
            loss = surrogate_loss(cls, bbox_2d, bbox_3d, *targets)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return loss

    @staticmethod
    def _tensors(obj):
        """The tensors of a (nested dict of) target(s), in a fixed order."""
        if isinstance(obj, dict):
            return [t for k in sorted(obj) for t in TrainStep._tensors(obj[k])]
        return [obj]

    @staticmethod
    def _clone(obj):
        if isinstance(obj, dict):
            return {k: TrainStep._clone(v) for k, v in obj.items()}
        return obj.clone()

    def __call__(self, images, *targets):
        if not self.use_graph:
            return self._iteration(images, *targets)
        if self._graph is None:
            if self._n < self.warmup:
                self._n += 1
                return self._iteration(images, *targets)
            self._static = [images.clone()] + [self._clone(t) for t in targets]
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._loss = self._iteration(*self._static)
            self._graph = g
        self._static[0].copy_(images, non_blocking=True)
        dsts = [t for tgt in self._static[1:] for t in self._tensors(tgt)]
        srcs = [t for tgt in targets for t in self._tensors(tgt)]
        for dst, src in zip(dsts, srcs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._loss


def targets_to(imobjs, device):
    """The reference's target dict (synth.make_targets / lib/dataloader.py:959-982) with every tensor on `device`."""
    return {k: (targets_to(v, device) if isinstance(v, dict) else v.to(device)) for k, v in imobjs.items()}
