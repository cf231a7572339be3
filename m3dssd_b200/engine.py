"""Fused inference engine for the M3DSSD dense forward path.

Reads the parameters of an RPN module (m3dssd_b200/model/M3d_inference_align.py,
same state-dict layout as the reference), folds every eval-mode BatchNorm into
the preceding convolution, packs weights for the tcgen05 kernels, pre-allocates
all NHWC activations for a fixed (batch, H, W) and records the layer sequence
as a list of C-ABI calls that is replayed per batch -- optionally as one CUDA
graph.  Torch provides device memory and streams only.

Forward path covered (reference file:line):
  DLA trunk ............. model/pose_dla_dcn.py:330-397 (+ Tree :272-327, BasicBlock :93-121, Root :251-269)
  DLAUp / IDAUp ......... model/pose_dla_dcn.py:519-578, 687-696; DeformConv :471-485; DCN model/DCNv2/dcn_v2.py:64-70
  heads, softmax, align . model/M3d_inference_align.py:215-301; model/module/feturealign_mgpu.py
  ANAB .................. model/module/attention.py:183-216
  decode + NMS .......... lib/rpn_util.py:1444-1555; lib/nms/nms_kernel.cu
"""
import os

import numpy as np
import torch
from torch import nn

from . import ops
from .model import pose_dla_dcn as dla

HEAD_ORDER = ["bbox_x", "bbox_y", "bbox_x3d", "bbox_y3d", "bbox_w", "bbox_h", "bbox_w3d", "bbox_h3d", "bbox_l3d",
              "bbox_rY3d", "bbox_z3d"]  # slot order in the heads buffer (grouped by shared input)
# output columns x,y,w,h | x3d,y3d,z3d,w3d,h3d,l3d,rY3d -> slot
OUT_SLOTS = [0, 1, 4, 5, 2, 3, 10, 6, 7, 8, 9]


_CONV_KINDS = ("conv3x3", "conv_tma", "conv_gather", "dcn_fused", "dcn_gather")  # ops that go through m3d_conv2d_nhwc
_KIND_KERNEL = {"stem": "stem_conv7x7_kernel", "stem_s2d": "stem_s2d_kernel", "head_mlp": "head_mlp_kernel<48>",
                "maxpool": "maxpool2x2_kernel", "upsample": "upsample_add_kernel", "softmax": "cls_softmax4_kernel",
                "align_om": "align_om_kernel", "flatten": "flatten_heads_kernel", "anab_pool": "anab_pool_region_kernel",
                "anab_attention": "anab_attention_tc_kernel"}


class Act:
    """An NHWC activation: `c` real channels inside a buffer with t.shape[-1] channels per pixel."""

    def __init__(self, t, c, coff=0, s2d=False):
        self.t, self.c, self.coff, self.s2d = t, c, coff, s2d  # s2d: 2x2 space-to-depth packed ((dy*2+dx)*c/4 + ch)


def _fold(conv_w, conv_b, bn):
    """Eval-mode BatchNorm folded into (weight, bias), on the HOST (fp32): weight preparation launches no device
    kernels, so an engine build shows up in a profile as H2D copies only."""
    w = conv_w.detach().float().cpu()
    b = conv_b.detach().float().cpu() if conv_b is not None else torch.zeros(w.shape[0])
    if bn is not None:
        scale = bn.weight.detach().float().cpu() / torch.sqrt(bn.running_var.detach().float().cpu() + bn.eps)
        w = w * scale.view(-1, 1, 1, 1)
        b = (b - bn.running_mean.detach().float().cpu()) * scale + bn.bias.detach().float().cpu()
    return w, b


class Engine:
    def __init__(self, net, batch, height, width, precision="bf16", use_graph=True, topk=None, max_out=None):
        assert precision in ("bf16", "bf16x3", "fp32")
        if not torch.cuda.is_available():
            raise RuntimeError("m3dssd_b200.Engine needs a CUDA device: there is no CPU fallback")
        pdev = next(net.parameters()).device
        self.dev = pdev if pdev.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        self.net = net
        self._names = {id(m): n for n, m in net.named_modules()}  # module -> state-dict prefix (layer specs)
        self.conf = net.conf
        self.B, self.H, self.W = batch, height, width
        self.fp32 = precision != "bf16"  # fp32 activation storage ("fp32": IEEE FMA; "bf16x3": tensor-core split)
        self.precision = precision
        self.adt = torch.float32 if self.fp32 else torch.bfloat16
        self.ops = []
        self.meta = []
        self.bufs = {}
        self._pool_cache = {}
        self.named = {}  # name -> Act of notable intermediate activations (parity checks)
        self.n_launches = 0
        self.detect_alt = {}  # op index -> cheaper callable used by the detection stages
        self.use_graph = use_graph
        self.graph = None
        self.A = net.num_anchors
        self.K = net.num_classes
        self.topk = int(topk or self.conf.nms_topN_pre)
        self.max_out = int(max_out or self.conf.nms_topN_post)
        self.image = torch.zeros(batch, 3, height, width, dtype=torch.float32, device=self.dev)
        with torch.no_grad(), torch.cuda.device(self.dev):
            self._build()

    # ------------------------------------------------------------------ utils
    def _cpad(self, c):
        return (c + 63) // 64 * 64 if self.fp32 else c

    def _new(self, name, n, h, w, c, dtype=None):
        t = torch.zeros(n, h, w, c, dtype=dtype or self.adt, device=self.dev)
        self.bufs[name] = t
        return t

    def _add(self, fn, launches=1, name="", kind="misc", flops=0.0, bytes_=0.0, spec=None):
        """Register one step of the plan.  flops / bytes_ are ALGORITHMIC (real channels, every
        operand read or written once), the numerators of the roofline figures bench.py reports.
        spec: what the step computes in terms of the reference's state-dict keys and the engine's own
        buffers (inputs / outputs as Act) -- the teacher-forced parity test replays every step's inputs
        through the oracle and compares with the step's output buffer (tests/test_teacher_forced_gpu.py)."""
        self.ops.append(fn)
        self.meta.append(dict(name=name, kind=kind, launches=launches, flops=float(flops), bytes=float(bytes_),
                              spec=spec))
        self.n_launches += launches

    def _key(self, mod):
        return self._names[id(mod)] if mod is not None else None

    def _conv(self, name, inputs, w, b, bn, k, stride=1, pad=None, slope=0.01, res=None, out=None, out_dtype=None,
              om=None, sigmoid_mask=False, out_coff=0, out_hw=None, keys=None):
        """Register conv(+BN)(+res)(+LeakyReLU) over concatenated NHWC inputs; returns the output Act."""
        pad = k // 2 if pad is None else pad
        wf, bf = _fold(w, b, bn)
        cout = wf.shape[0]
        splits = [(a.c, self._cpad(a.c)) for a in inputs]
        hi, lo = ops.pack_conv_weight(wf.cpu(), in_splits=splits, mode=self.precision)
        kz = ops.k16_zero_mask(hi) if (hi is not None and lo is None) else (0, 0)  # bf16 mode: zero k-steps may be skipped
        hi = hi.to(self.dev) if hi is not None else None
        if isinstance(lo, tuple):
            lo = tuple(t.to(self.dev) for t in lo)
        elif lo is not None:
            lo = lo.to(self.dev)
        bias = bf.to(self.dev).contiguous()
        x0 = inputs[0].t
        N, H, W = x0.shape[:3]
        P = (H + 2 * pad - k) // stride + 1
        Q = (W + 2 * pad - k) // stride + 1
        if out_hw is not None:
            P, Q = out_hw
        if out is None:
            odt = out_dtype or self.adt
            cbuf = self._cpad(cout) if odt == self.adt else cout
            out = self._new(name, N, P, Q, cbuf, odt)
        ins = [(a.t, a.coff, self._cpad(a.c)) for a in inputs]
        res_t = res.t if res is not None else None
        res_coff = res.coff if res is not None else 0
        om_t = om

        def run():
            ops.conv2d_nhwc(ins, hi, out, R=k, S=k, stride=stride, pad=pad, Cout=cout, bias=bias, res=res_t,
                            res_coff=res_coff, slope=slope, weight_lo=lo, om=om_t, sigmoid_mask=sigmoid_mask,
                            out_coff=out_coff, out_hw=out_hw, k16_zero=kz)

        esz = 4 if self.fp32 else 2
        k_real = k * k * sum(a.c for a in inputs)
        flops = 2.0 * N * P * Q * cout * k_real
        nbytes = sum(N * H * W * a.c * esz for a in inputs) + N * P * Q * cout * out.element_size() + cout * k_real * 2
        if res is not None:
            nbytes += N * P * Q * cout * esz
        if om is not None:
            nbytes += N * P * Q * 3 * k * k * 4
        if om is not None:
            kind = "dcn_fused" if (k == 3 and not self.fp32) else "dcn_gather"  # dcn_fused.cu / conv_gather_kernel
        elif self.fp32:
            kind = "conv_gather"
        elif k == 3 and stride == 1 and pad == 1 and len(inputs) == 1 and inputs[0].c % 64 == 0:
            kind = "conv3x3"  # conv_halo.cu / conv_halo2.cu (CTA pairs)
        else:
            kind = "conv_tma"
        oact = Act(out, cout, out_coff)
        spec = None
        if keys is not None:  # keys = (conv state-dict prefix or list of prefixes, bn prefix or None[, geometry])
            spec = dict(op="dcn" if om is not None else "conv", inputs=list(inputs), conv=keys[0], bn=keys[1], k=k,
                        stride=stride, pad=pad, slope=slope, res=res, out=oact, om=om, sigmoid_mask=sigmoid_mask)
            if len(keys) > 2:
                spec.update(keys[2])
        self._add(run, 1, name, kind, flops, nbytes, spec)
        return oact

    def _conv_module(self, name, inputs, conv, bn, slope=0.01, res=None, **kw):
        return self._conv(name, inputs, conv.weight, conv.bias, bn, conv.kernel_size[0], conv.stride[0],
                          conv.padding[0], slope, res, keys=(self._key(conv), self._key(bn)), **kw)

    def _maxpool(self, name, a):
        key = (a.t.data_ptr(), a.coff, a.c)
        if key in self._pool_cache:  # nested Trees pool the same tensor twice in the reference
            return self._pool_cache[key]
        r = self._maxpool_new(name, a)
        self._pool_cache[key] = r
        return r

    def _maxpool_new(self, name, a):
        N, H, W, cs = a.t.shape
        out = self._new(name, N, H // 2, W // 2, cs)
        x = a.t
        r = Act(out, a.c)
        self._add(lambda: ops.maxpool2x2(x, out), 1, name, "maxpool", 0, x.numel() * x.element_size() * 1.25,
                  dict(op="maxpool", x=a, out=r))
        return r

    # ------------------------------------------------------------- DLA trunk
    def _basic_block(self, name, blk, x, residual):
        res = residual if residual is not None else x
        if isinstance(blk, dla.Bottleneck):  # dla102 (model/pose_dla_dcn.py:162-204): 1x1 -> 3x3 (stride) -> 1x1 + residual
            y = self._conv_module(name + ".conv1", [x], blk.conv1, blk.bn1)
            y = self._conv_module(name + ".conv2", [y], blk.conv2, blk.bn2)
            return self._conv_module(name + ".conv3", [y], blk.conv3, blk.bn3, res=res)
        if not isinstance(blk, dla.BasicBlock):
            raise NotImplementedError("fused engine implements BasicBlock / Bottleneck trunks (dla34, dla102); got %s"
                                      % type(blk).__name__)
        y = self._conv_module(name + ".conv1", [x], blk.conv1, blk.bn1)
        return self._conv_module(name + ".conv2", [y], blk.conv2, blk.bn2, res=res)

    def _tree(self, name, tree, x, children=None):
        children = [] if children is None else children
        bottom = self._maxpool(name + ".down", x) if tree.downsample is not None else x
        if tree.level_root:
            children.append(bottom)
        if tree.levels == 1:
            residual = bottom
            if tree.project is not None:
                residual = self._conv_module(name + ".project", [bottom], tree.project[0], tree.project[1], slope=1.0)
            x1 = self._basic_block(name + ".tree1", tree.tree1, x, residual)
            x2 = self._basic_block(name + ".tree2", tree.tree2, x1, None)
            root = tree.root
            ins = [x2, x1] + children
            return self._conv_module(name + ".root", ins, root.conv, root.bn, res=x2 if root.residual else None)
        # levels > 1: the reference also evaluates self.project(bottom) here, but the nested
        # tree1 recomputes its own residual and never reads it (model/pose_dla_dcn.py:314-320): skipped.
        x1 = self._tree(name + ".tree1", tree.tree1, x)
        children.append(x1)
        return self._tree(name + ".tree2", tree.tree2, x1, children)

    def _trunk(self):
        base = self.net.base.base
        B, H, W = self.B, self.H, self.W
        w, b = _fold(base.base_layer[0].weight, None, base.base_layer[1])
        c0 = w.shape[0]
        assert c0 == 16, "stem kernel is specialised for 16 output channels"
        l0, l1 = base.level0, base.level1
        if (not self.fp32 and H % 2 == 0 and W % 2 == 0 and len(l0) == 3 and len(l1) == 3
                and l0[0].out_channels == 16 and l1[0].stride[0] == 2):
            return self._trunk_s2d(base, w, b)
        w, b = w.to(self.dev).contiguous(), b.to(self.dev).contiguous()
        s0 = self._new("stem", B, H, W, self._cpad(c0))
        img = self.image
        x = Act(s0, c0)
        stem_spec = dict(op="stem", conv=self._key(base.base_layer[0]), bn=self._key(base.base_layer[1]), out=x)
        self._add(lambda: ops.stem_conv7x7(img, w, b, s0, 0.01), 1, "stem", "stem", 2.0 * B * H * W * c0 * 147,
                  B * H * W * (3 * 4 + c0 * s0.element_size()), stem_spec)
        x = self._conv_module("level0", [x], base.level0[0], base.level0[1])
        levels = [x]
        x = self._conv_module("level1", [x], base.level1[0], base.level1[1])
        levels.append(x)
        for i in range(2, 6):
            x = self._tree("level%d" % i, getattr(base, "level%d" % i), x)
            levels.append(x)
        for i, a in enumerate(levels):
            self.named["level%d" % i] = a
        return levels

    def _trunk_s2d(self, base, w_stem, b_stem):
        """bf16 trunk head in 2x2 space-to-depth form: the three full-resolution 16-channel layers (stem 7x7,
        level0 3x3, level1 3x3 s2) become well-shaped GEMMs on [N, H/2, W/2, 64] tensors -- same arithmetic,
        zero-padded weights -- instead of 32-byte-per-pixel tiles the TMA / tensor pipes handle poorly."""
        B, H, W = self.B, self.H, self.W
        wp, bp = ops.pack_stem_s2d(w_stem, b_stem)
        wp, bp = wp.to(self.dev), bp.to(self.dev)
        s0 = self._new("stem", B, H // 2, W // 2, 64)
        img = self.image
        x = Act(s0, 64, s2d=True)
        stem_spec = dict(op="stem", conv=self._key(base.base_layer[0]), bn=self._key(base.base_layer[1]), out=x)
        self._add(lambda: ops.stem_conv7x7_s2d(img, wp, bp, s0, 0.01), 1, "stem", "stem_s2d",
                  2.0 * B * H * W * 16 * 147, B * H * W * (3 * 4 + 16 * 2), stem_spec)
        w0, b0 = _fold(base.level0[0].weight, None, base.level0[1])
        w0p, b0p = ops.s2d_conv3x3_weight(w0, b0)
        # spec geometry = the ORIGINAL layer (3x3 / stride 1 / pad 1 on the un-packed tensor), not the s2d rewrite
        x = self._conv("level0", [x], w0p, b0p, None, 3, 1, 1, 0.01,
                       keys=(self._key(base.level0[0]), self._key(base.level0[1]), dict(k=3, stride=1, pad=1)))
        x.s2d = True
        self.meta[-1]["spec"]["out"] = x
        self._fix_meta("level0", 2.0 * B * H * W * 16 * 144, B * H * W * 16 * 2 * 2)
        levels = [x]
        w1, b1 = _fold(base.level1[0].weight, None, base.level1[1])
        x = self._conv("level1", [x], ops.s2d_conv3x3_s2_weight(w1), b1, None, 2, 1, 1, 0.01, out_hw=(H // 2, W // 2),
                       keys=(self._key(base.level1[0]), self._key(base.level1[1]), dict(k=3, stride=2, pad=1)))
        self._fix_meta("level1", 2.0 * B * (H // 2) * (W // 2) * w1.shape[0] * 144,
                       B * H * W * 16 * 2 + B * (H // 2) * (W // 2) * w1.shape[0] * 2)
        levels.append(x)
        for i in range(2, 6):
            x = self._tree("level%d" % i, getattr(base, "level%d" % i), x)
            levels.append(x)
        for i, a in enumerate(levels):
            self.named["level%d" % i] = a
        return levels

    def _fix_meta(self, name, flops, bytes_):
        """Roofline numerators stay ALGORITHMIC: the zero-padded s2d weights do not count as work."""
        for m in self.meta:
            if m["name"] == name:
                m["flops"], m["bytes"] = float(flops), float(bytes_)

    # ----------------------------------------------------------- aggregation
    def _deform_conv(self, name, m, x):
        """DeformConv = DCN 3x3 -> BN -> LeakyReLU (model/pose_dla_dcn.py:471-485), BN folded into the DCN."""
        dcn = m.conv
        N, H, W = x.t.shape[:3]
        om = self._new(name + ".om", N, H, W, 32, torch.float32)
        self._conv_module(name + ".offset", [x], dcn.conv_offset_mask, None, slope=1.0, out=om)
        return self._conv(name, [x], dcn.weight, dcn.bias, m.actf[0], 3, 1, 1, 0.01, om=om, sigmoid_mask=True,
                          keys=(self._key(dcn), self._key(m.actf[0])))

    def _ida_up(self, name, ida, layers, startp, endp):
        for i in range(startp + 1, endp):
            k = i - startp
            proj, up, node = (getattr(ida, "%s_%d" % (n, k)) for n in ("proj", "up", "node"))
            p = self._deform_conv("%s.proj_%d" % (name, k), proj, layers[i])
            f = up.stride[0]
            N, H, W, cs = p.t.shape
            skip = layers[i - 1]
            u = self._new("%s.up_%d" % (name, k), N, H * f, W * f, cs)
            wt = ops.pack_upsample_weight(up.weight).to(self.dev)
            pt, st = p.t, skip.t
            ua = Act(u, p.c)
            self._add(lambda pt=pt, wt=wt, st=st, u=u, f=f: ops.upsample_add(pt, wt, st, u, f), 1,
                      "%s.up_%d" % (name, k), "upsample", 2.0 * u.numel() * 4,
                      (pt.numel() + 2 * u.numel()) * u.element_size(),
                      dict(op="upsample", x=p, skip=skip, up=self._key(up), f=f, out=ua))
            layers[i] = self._deform_conv("%s.node_%d" % (name, k), node, ua)

    def _dla_seg(self):
        seg = self.net.base
        layers = self._trunk()
        first, last = seg.first_level, seg.last_level
        out = [layers[-1]]
        for i in range(len(layers) - first - 1):
            self._ida_up("dla_up.ida_%d" % i, getattr(seg.dla_up, "ida_%d" % i), layers, len(layers) - i - 2,
                         len(layers))
            out.insert(0, layers[-1])
        y = [out[i] for i in range(last - first)]  # the reference clones; nothing writes in place here
        self._ida_up("ida_up", seg.ida_up, y, 0, len(y))
        return y[-1]

    # ------------------------------------------------------------------ heads
    def _head_group(self, gname, names, x, heads_buf):
        """Three-layer 1x1 heads that share input x; layer 1 is one wide GEMM, layers 2-3 are grouped GEMMs."""
        net, A = self.net, self.A
        mods = [getattr(net, n) for n in names]
        slot0 = HEAD_ORDER.index(names[0])
        assert [HEAD_ORDER.index(n) for n in names] == list(range(slot0, slot0 + len(names)))
        N, H, W = x.t.shape[:3]
        G = len(names)
        mid = mods[0][0].out_channels
        if self.fp32 or not (mid == 256 and A <= 48 and x.c in (64, 128)):
            # layer by layer: the fp32 parity modes, and head shapes the fused kernel does not cover (dla102: 256 channels)
            for g, m in enumerate(mods):
                h1 = self._conv_module("%s.%d.l1" % (gname, g), [x], m[0], m[1])
                h2 = self._conv_module("%s.%d.l2" % (gname, g), [h1], m[3], m[4])
                self._conv_module("%s.%d.l3" % (gname, g), [h2], m[6], None, slope=1.0, out=heads_buf,
                                  out_coff=(slot0 + g) * A)
            return
        # one fused kernel: the 256-channel intermediates never leave the SM (csrc/heads.cu)
        rows3 = 48
        assert mid == 256 and A <= rows3 and x.c in (64, 128)
        f1 = [_fold(m[0].weight, m[0].bias, m[1]) for m in mods]
        f2 = [_fold(m[3].weight, m[3].bias, m[4]) for m in mods]
        w1 = torch.cat([ops.pack_conv_weight(w.cpu())[0] for w, _ in f1]).to(self.dev).contiguous()
        b1 = torch.cat([b for _, b in f1]).to(self.dev).float().contiguous()
        w2 = torch.cat([ops.pack_conv_weight(w.cpu())[0] for w, _ in f2]).to(self.dev).contiguous()
        b2 = torch.cat([b for _, b in f2]).to(self.dev).float().contiguous()
        w3 = torch.zeros(G * rows3, mid, dtype=torch.bfloat16)
        b3 = torch.zeros(G * rows3, dtype=torch.float32)
        for g, m in enumerate(mods):
            w3[g * rows3:g * rows3 + A] = ops.pack_conv_weight(m[6].weight.detach().float().cpu())[0]
            b3[g * rows3:g * rows3 + A] = m[6].bias.detach().float().cpu()
        w3, b3 = w3.to(self.dev), b3.to(self.dev)
        xt, xc, xoff = x.t, x.c, x.coff

        def run():
            ops.head_mlp(xt, xoff, xc, w1, b1, w2, b2, w3, b3, G, A, rows3, heads_buf, slot0 * A, 0.01)

        fl = 2.0 * N * H * W * G * (xc * mid + mid * mid + A * mid)
        by = N * H * W * (xc * 2 + G * A * 4) + G * (xc * mid + mid * mid + A * mid) * 2
        self._add(run, 1, gname + ".mlp", "head_mlp", fl, by,
                  dict(op="head_mlp", x=x, heads=[self._key(m) for m in mods], slot0=slot0, out=heads_buf))

    def _align(self, name, m, x, om):
        return self._conv(name, [x], m.align.weight, m.align.bias, None, m.align.kernel_size[0], 1, m.align.padding,
                          1.0, res=x, om=om, keys=(self._key(m.align), None))

    def _anab(self, x):
        """bbox_z3d_gl = ANAB -> BN -> LeakyReLU (model/M3d_inference_align.py:168-173); BN folded into the epilogue."""
        seq = self.net.bbox_z3d_gl
        anab, bn = seq[0], seq[1]
        if not anab.with_atten:
            raise NotImplementedError("ANAB(with_atten=False) is not on the reference's path")
        ck, cv, T = anab.key_ch, anab.outch, anab.key_num
        sizes = list(anab.psp_size)
        N, H, W = x.t.shape[:3]
        f32 = dict(dtype=torch.float32, device=self.dev)
        q = self._conv_module("anab.q", [x], anab.query_conv, None, slope=1.0)
        w = torch.cat([anab.key_conv.weight, anab.value_conv.weight, anab.spatial_conv.weight]).detach()
        ckvs = (ck + cv + len(sizes) + 3) // 4 * 4
        kvs = self._new("anab.kvs", N, H, W, ckvs, torch.float32)
        self._conv("anab.kvs", [x], w, None, None, 1, 1, 0, 1.0, out=kvs,
                   keys=([self._key(anab.key_conv), self._key(anab.value_conv), self._key(anab.spatial_conv)], None))
        ktok = torch.zeros(N, T, ck, **f32)
        vtok = torch.zeros(N, T, cv, **f32)
        ws = torch.zeros(ops.anab_pool_workspace(N, H, sizes, ck, cv), dtype=torch.uint8, device=self.dev)
        self._add(lambda: ops.anab_pool(kvs, ck, cv, sizes, ktok, vtok, ws), 2, "anab.pool", "anab_pool",
                  2.0 * N * H * W * (ck + cv) * len(sizes), kvs.numel() * 4,
                  dict(op="anab_pool", kvs=kvs, ck=ck, cv=cv, sizes=sizes, ktok=ktok, vtok=vtok))
        scale = (bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)).to(self.dev)
        shift = (bn.bias.detach().float().to(self.dev) - bn.running_mean.detach().float().to(self.dev) * scale)
        scale, shift = scale.contiguous(), shift.contiguous()
        out = self._new("anab.out", N, H, W, x.t.shape[-1])
        qt, xt = q.t, x.t
        aws = ops.anab_attention_workspace(N, xt)  # owned by this engine (baked into its graphs)
        oa = Act(out, x.c)
        self._add(lambda: ops.anab_attention(qt, ktok, vtok, xt, scale, shift, 0.01, out, ck, cv, aws), 1,
                  "anab.attention", "anab_attention", 2.0 * N * H * W * T * (ck + cv),
                  N * H * W * (ck + 2 * cv) * out.element_size() + N * T * (ck + cv) * 4,
                  dict(op="anab_attention", q=q, ktok=ktok, vtok=vtok, x=x, bn=self._key(bn), out=oa))
        self.bufs["anab.ktok"], self.bufs["anab.vtok"], self.bufs["anab.ws"] = ktok, vtok, aws
        return oa

    def _build(self):
        net, conf = self.net, self.conf
        A, K, B = self.A, self.K, self.B
        feat = self._dla_seg()
        self.n_trunk_ops = len(self.ops)  # ops [0, n_trunk_ops) never touch the buffers the detection tail reads
        self.named["feat"] = feat
        N, Hf, Wf = feat.t.shape[:3]
        self.Hf, self.Wf = Hf, Wf
        M = A * Hf * Wf
        self.M = M
        # --- classification head
        c = net.cls
        h = self._conv_module("cls.l1", [feat], c[0], c[1])
        h = self._conv_module("cls.l2", [h], c[3], c[4])
        logits = self._new("cls.logits", N, Hf, Wf, K * A, torch.float32)
        self._conv_module("cls.l3", [h], c[6], None, slope=1.0, out=logits)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.cls_out = torch.zeros(B, M, K, **f32)
        self.prob_out = torch.zeros(B, M, K, **f32)
        self.fg_max = torch.zeros(B, Hf, Wf, **f32)
        self.fg_arg = torch.zeros(B, Hf, Wf, dtype=torch.int32, device=self.dev)
        self.score = torch.zeros(B, M, **f32)
        self.cls_pred = torch.zeros(B, M, dtype=torch.uint8, device=self.dev)
        self._add(lambda: ops.cls_softmax(logits, A, K, self.cls_out, self.prob_out, self.fg_max, self.fg_arg,
                                          self.score, self.cls_pred), 1, "cls_softmax", "softmax", 0, B * M * K * 4 * 3,
                  dict(op="softmax", logits=logits))
        # detection stages: same kernel without the flattened cls / prob copies (70 MB of stores nobody reads there)
        i_softmax = len(self.ops) - 1
        self.detect_alt[i_softmax] = lambda: ops.cls_softmax(logits, A, K, None, None, self.fg_max, self.fg_arg,
                                                             self.score, self.cls_pred)
        anchors = torch.tensor(np.asarray(conf.anchors), **f32).contiguous()
        self.anchors = anchors
        means = [float(v) for v in np.asarray(conf.bbox_means)[0]]
        stds = [float(v) for v in np.asarray(conf.bbox_stds)[0]]
        stride = float(conf.feat_stride)
        # --- shape align
        feats = feat
        if net.shape_align is not None:
            om_s = torch.zeros(B, Hf, Wf, 27, **f32)
            thr_s = float(net.shape_align.thresh)
            self._add(lambda: ops.shape_align_om(self.fg_max, self.fg_arg, anchors, stride, thr_s, om_s), 1,
                      "shape_align.om", "align_om", 0, om_s.numel() * 4, dict(op="shape_align_om", om=om_s))
            # detection stages: the offset builder rides in the softmax kernel's per-pixel tail
            self.detect_alt[i_softmax] = lambda: ops.cls_softmax_shape_om(logits, A, K, None, None, self.fg_max, self.fg_arg,
                                                                          self.score, self.cls_pred, anchors, stride, thr_s, om_s)
            self.detect_alt[len(self.ops) - 1] = None
            feats = self._align("shape_align", net.shape_align, feat, om_s)
        # --- regression heads (slots follow HEAD_ORDER)
        heads = self._new("heads", B, Hf, Wf, 11 * A, torch.float32)
        self.heads = heads
        self._head_group("headsA", ["bbox_x", "bbox_y", "bbox_x3d", "bbox_y3d"], feats, heads)
        f2d = f3d = feats
        if net.center_align2d is not None:
            om2 = torch.zeros(B, Hf, Wf, 4, **f32)
            om3 = torch.zeros(B, Hf, Wf, 4, **f32)
            thr = float(net.center_align2d.thresh)
            sx, sy, sx3, sy3 = (HEAD_ORDER.index(n) * A for n in ("bbox_x", "bbox_y", "bbox_x3d", "bbox_y3d"))
            self._add(lambda: ops.center_align_om(self.fg_max, self.fg_arg, heads, sx, sy, anchors, stride, means[0:2],
                                                  stds[0:2], thr, om2), 1, "center_align2d.om", "align_om", 0, om2.numel() * 4,
                      dict(op="center_align_om", om=om2, hx="bbox_x", hy="bbox_y", mean=means[0:2], std=stds[0:2]))
            self._add(lambda: ops.center_align_om(self.fg_max, self.fg_arg, heads, sx3, sy3, anchors, stride,
                                                  means[4:6], stds[4:6], thr, om3), 1, "center_align3d.om", "align_om", 0, om3.numel() * 4,
                      dict(op="center_align_om", om=om3, hx="bbox_x3d", hy="bbox_y3d", mean=means[4:6], std=stds[4:6]))
            # detection stages: one launch builds both
            self.detect_alt[len(self.ops) - 2] = lambda: ops.center_align_om2(
                self.fg_max, self.fg_arg, heads, (sx, sy, sx3, sy3), anchors, stride, means[0:2] + means[4:6],
                stds[0:2] + stds[4:6], thr, om2, om3)
            self.detect_alt[len(self.ops) - 1] = None
            f2d = self._align("center_align2d", net.center_align2d, feats, om2)
            f3d = self._align("center_align3d", net.center_align3d, feats, om3)
        self.named["feats_shape"], self.named["feats_align2d"], self.named["feats_align3d"] = feats, f2d, f3d
        self._head_group("headsB", ["bbox_w", "bbox_h"], f2d, heads)
        self._head_group("headsC", ["bbox_w3d", "bbox_h3d", "bbox_l3d", "bbox_rY3d"], f3d, heads)
        fz = f3d
        if net.attention == "ANAB":
            fz = self._anab(f3d)
        self.named["feats_gl"] = fz
        self._head_group("headsZ", ["bbox_z3d"], fz, heads)
        self.bbox_2d = torch.zeros(B, M, 4, **f32)
        self.bbox_3d = torch.zeros(B, M, 7, **f32)
        # the reference's flattened outputs (lib/rpn_util.py:892-901): part of "forward" (what RPN.forward returns) only;
        # the detection stages decode their <= topk rows straight from `heads` (m3d_decode_topk_heads)
        self.n_detect_ops = len(self.ops)
        self._add(lambda: ops.flatten_heads(heads, A, OUT_SLOTS, self.bbox_2d, self.bbox_3d), 1, "flatten_heads",
                  "flatten", 0, B * M * 11 * 4 * 2, dict(op="flatten"))
        self.n_forward_ops = len(self.ops)
        # --- detection tail
        self.means_t, self.stds_t = means, stds  # host lists (the C ABI takes them by value)
        self.dets = torch.zeros(B, self.topk, 14, **f32)
        self.det_idx = torch.zeros(B, self.topk, dtype=torch.int32, device=self.dev)
        self.det_num = torch.zeros(B, dtype=torch.int32, device=self.dev)
        self.keep = torch.zeros(B, self.topk, dtype=torch.int32, device=self.dev)
        self.num_keep = torch.zeros(B, dtype=torch.int32, device=self.dev)
        self.nms_ws = torch.zeros(ops.nms_workspace_bytes(B, self.topk), dtype=torch.uint8, device=self.dev)
        self.topk_ws = ops.decode_topk_workspace(B, self.dev)  # per engine: its pointer is baked into the captured graphs
        self.kept = torch.zeros(B, self.max_out, 14, **f32)
        self.scale_factor = 1.0
        self.feat_size = torch.tensor([Hf, Wf], dtype=torch.float32, device=self.dev)

    # ------------------------------------------------------------------- run
    def _detect_ops(self):
        """The op list of the detection stages: no flattened copies of the network outputs (cls / prob / bbox_2d /
        bbox_3d are RPN.forward's return values; decode reads score, class and the head buffer)."""
        alt = [self.detect_alt.get(i, op) for i, op in enumerate(self.ops[:self.n_detect_ops])]
        return [op for op in alt if op is not None]  # None: folded into a neighbour's kernel

    def _side_tasks(self):
        """Independent branches of the aggregation network that may run beside the small-grid layers of the trunk:
        (after, [ops], before) = once op `after` has been issued the listed ops go to a side stream; the main stream
        waits for them before op `before`.  dla_up.ida_1.proj_1 only reads level4's output, so its offset conv + DCN
        (120 tiles, L1-bound) run under level5 + ida_0 (48..120 CTAs per launch, tensor-bound)."""
        if os.environ.get("M3D_SIDE", "1") == "0":
            return []
        names = [m["name"] for m in self.meta]
        tasks = []

        want = os.environ.get("M3D_SIDE", "1")  # "0": none, "1": all, or a comma list of agg / pool / heads (development)

        def add(tag, after, side, before):
            if want != "1" and tag not in want.split(","):
                return
            try:
                a = max(i for i, n in enumerate(names) if n == after or n.startswith(after + "."))
                ops_ = [names.index(n) for n in side]
                b = names.index(before)
            except ValueError:
                return
            if all(a < i < b for i in ops_):
                tasks.append(dict(after=a, ops=ops_, before=b))
        # dla_up.ida_1.proj_1 reads level4 only: beside level5 + ida_0
        add("agg", "level4", ["dla_up.ida_1.proj_1.offset", "dla_up.ida_1.proj_1"], "dla_up.ida_1.up_1")
        # dla_up.ida_1.proj_2 and ida_up.proj_1 read ida_0's output only: beside ida_1's first up / node pair
        add("agg", "dla_up.ida_0.node_1", ["dla_up.ida_1.proj_2.offset", "dla_up.ida_1.proj_2"], "dla_up.ida_1.up_2")
        add("agg", "dla_up.ida_0.node_1", ["ida_up.proj_1.offset", "ida_up.proj_1"], "ida_up.up_1")
        # Tree: max-pool + 1x1 project of the residual (model/pose_dla_dcn.py:303-309) beside the block's first conv
        for lvl, prev in ((2, "level1"), (3, "level2"), (4, "level3"), (5, "level4")):
            L = "level%d" % lvl
            for proj, conv2 in ((L + ".project", L + ".tree1.conv2"), (L + ".tree1.project", L + ".tree1.tree1.conv2")):
                add("pool", prev, [L + ".down", proj], conv2)
        # the two centre alignments and their head groups are independent chains
        add("heads", "center_align3d.om", ["center_align2d", "headsB.mlp"], "flatten_heads")
        # ... and so is the depth head (input: the 3D-aligned features, or ANAB's output when the attention block is on)
        add("heads", "anab.attention" if "anab.attention" in names else "center_align3d", ["headsZ.mlp"], "flatten_heads")
        return tasks

    def _exec(self, lo, hi, detect=False):
        """Issue ops [lo, hi) of the plan (detect: the detection stages' variants), branches of _side_tasks on the side
        stream (fork / join by stream events: captured into the CUDA graph as parallel branches)."""
        if not hasattr(self, "_tasks"):
            self._tasks = self._side_tasks()
            self._side = torch.cuda.Stream() if self._tasks else None
        tasks = [t for t in self._tasks if lo <= t["after"] and t["before"] <= hi]
        on_side = {i for t in tasks for i in t["ops"]}
        cur = torch.cuda.current_stream()
        pending = set()

        def pick(i):
            return self.detect_alt.get(i, self.ops[i]) if detect else self.ops[i]

        def join(i):
            for k, t in enumerate(tasks):
                if t["before"] == i and k in pending:
                    cur.wait_event(t["done"])
                    pending.discard(k)
        safe = os.environ.get("M3D_SIDE_PDL", "0") == "0"  # no PDL while a branch is in flight (see m3d_set_pdl)
        try:
            for i in range(lo, hi):
                join(i)
                op = pick(i)
                if op is not None and i not in on_side:
                    op()  # (the first op after a join is still launched without PDL: it waits for its predecessor,
                    #        which may have had unscheduled CTAs, to complete)
                if safe and not pending:
                    ops.set_pdl(True)
                for k, t in enumerate(tasks):
                    if t["after"] == i:
                        if safe:
                            ops.set_pdl(False)
                        self._side.wait_stream(cur)
                        with torch.cuda.stream(self._side):
                            for j in t["ops"]:
                                if pick(j) is not None:
                                    pick(j)()
                            t["done"] = torch.cuda.Event()
                            t["done"].record(self._side)
                        pending.add(k)
        finally:
            ops.set_pdl(True)
        join(hi)  # branches that run to the end of the range

    def _run_forward(self, flatten=True):
        self._exec(0, len(self.ops) if flatten else self.n_detect_ops, detect=not flatten)

    def _run_decode(self):
        ops.decode_topk_heads(self.score, self.cls_pred, self.heads, OUT_SLOTS, self.anchors, self.means_t, self.stds_t,
                              self.A, self.Hf, self.Wf, float(self.conf.feat_stride), self.scale_factor, self.topk,
                              self.dets, self.det_idx, self.det_num, self.topk_ws)

    def _run_nms(self):
        ops.nms_batched(self.dets, self.det_num, float(self.conf.nms_thres), self.nms_ws, self.keep, self.num_keep)
        ops.gather_kept(self.dets, self.keep, self.num_keep, self.max_out, self.kept)

    _STAGES = {"forward": 0, "decode": 1, "detect": 2}

    def _run_stage(self, stage):
        # bbox_2d / bbox_3d (the flattened copies) are refreshed by "forward" only
        self._run_forward(flatten=self._STAGES[stage] == 0)
        if self._STAGES[stage] >= 1:
            self._run_decode()
        if self._STAGES[stage] >= 2:
            self._run_nms()

    def launches_per_step(self, stage="detect"):
        # decode = 2 histogram passes + compaction + finish (m3d_decode_topk); NMS = mask + sweep + gather of the kept rows
        st = self._STAGES[stage]
        folded = 1 + sum(1 for v in self.detect_alt.values() if v is None)  # flatten_heads + launches folded into neighbours
        return self.n_launches - (folded if st >= 1 else 0) + (0, 4, 7)[st]

    def activation_nchw(self, name):
        """fp32 NCHW copy of a named intermediate activation (testing aid)."""
        a = self.named[name]
        t = a.t[..., a.coff:a.coff + a.c].float()
        if a.s2d:  # [N, Y, X, (dy, dx, c)] -> [N, 2Y+dy, 2X+dx, c]
            n, hh, ww, c4 = t.shape
            t = t.view(n, hh, ww, 2, 2, c4 // 4).permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * hh, 2 * ww, c4 // 4)
        return t.permute(0, 3, 1, 2).contiguous()

    def run(self, images=None, stage="forward"):
        """One pass over the engine's batch.  images: [B,3,H,W] fp32 CUDA tensor copied into the input
        buffer (None = reuse what is there).  stage: "forward" (network outputs), "decode" (+ top-K
        decode into self.dets) or "detect" (+ batched NMS into self.kept / self.num_keep)."""
        if images is not None:
            self._set_input(images)
        if not self.use_graph:
            self._run_stage(stage)
            return
        if self.graph is None:
            self.graph = {}
        if stage not in self.graph:
            self._run_stage(stage)  # eager pass first: module load, shared-memory attributes
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run_stage(stage)
            self.graph[stage] = g
        self.graph[stage].replay()

    def _set_input(self, images):
        """fp32 NCHW [B,3,H,W] (already normalised, as the reference's DataLoader yields) is copied into the input
        buffer; uint8 HWC [B,H,W,3] (what cv2.imread returns) goes through the device-side input pipeline:
        Normalize + BGR->RGB + HWC->CHW of lib/augmentations.py:44-57 / lib/dataloader.py:942-950, bit-identical;
        a list of B uint8 HWC frames of different sizes is zero-padded to H x W first (the reference's Preprocess)."""
        mean = self.conf.get("image_means", (0.485, 0.456, 0.406))
        std = self.conf.get("image_stds", (0.229, 0.224, 0.225))
        if isinstance(images, (list, tuple)):
            # ragged uint8 HWC frames (KITTI: 370-376 x 1224-1242): the reference's Preprocess = zero Padding on the
            # bottom / right to the test size, then Normalize (lib/augmentations.py:472-492), on the device
            assert len(images) == self.B, (len(images), self.B)
            buf, off, hh, ww = ops.pack_ragged_u8([im.cpu() if torch.is_tensor(im) else im for im in images])
            ops.preprocess_u8_pad(buf.to(self.dev, non_blocking=True), off, hh, ww, self.image, mean, std, swap_rb=True)
            return
        if images.dtype == torch.uint8:
            assert tuple(images.shape) == (self.B, self.H, self.W, 3), (images.shape, self.image.shape)
            ops.preprocess_u8(images.contiguous(), self.image, mean, std, swap_rb=True)
            return
        assert tuple(images.shape) == tuple(self.image.shape), (images.shape, self.image.shape)
        self.image.copy_(images, non_blocking=True)

    def flatten_outputs(self):
        """Refresh the reference's flattened network outputs (cls, prob, bbox_2d, bbox_3d) from the logits / head
        buffers after a "decode" / "detect" / pipelined step, which do not write them.  Returns the four tensors."""
        self.ops[min(self.detect_alt)]()  # the full softmax (flattened cls / prob copies)
        self.ops[self.n_detect_ops]()
        return self.cls_out, self.prob_out, self.bbox_2d, self.bbox_3d

    def forward(self, images=None):
        """Returns (cls, prob, bbox_2d, bbox_3d): views of the engine's output buffers."""
        self.run(images, "forward")
        return self.cls_out, self.prob_out, self.bbox_2d, self.bbox_3d

    def detect(self, images=None, scale_factor=1.0):
        """Full path: forward + decode/top-K + batched NMS.  Returns (kept [B,max_out,14], num_keep [B])."""
        if float(scale_factor) != self.scale_factor:
            self.scale_factor = float(scale_factor)
            self.graph = None
        self.run(images, "detect")
        return self.kept, self.num_keep

    # ------------------------------------------------------------- pipelined
    def _pipe_init(self):
        """Three CUDA graphs and a side stream: trunk (DLA + aggregation), heads (everything that writes
        the score / box buffers) and the detection tail.  The tail of batch i runs on the side stream
        under the trunk of batch i+1; only the heads of batch i+1 wait for it (they overwrite its inputs)."""
        self._pipe = {}
        self._tail_stream = torch.cuda.Stream()
        self._ev_fwd = torch.cuda.Event()
        self._ev_tail = torch.cuda.Event()
        self._ev_tail.record(self._tail_stream)

        def capture(fn):
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            return g

        def trunk():
            self._exec(0, self.n_trunk_ops, detect=True)

        def heads():
            self._exec(self.n_trunk_ops, self.n_detect_ops, detect=True)

        if self.use_graph:
            # the trunk of batch i+1 shares the device with the detection tail of batch i (one CTA per image on
            # the side stream): leave those SMs free, or every persistent kernel of the trunk runs as two waves
            nsm = torch.cuda.get_device_properties(self.dev).multi_processor_count
            reserve = int(os.environ.get("M3D_TAIL_SMS", str(self.B)))
            ops.set_sm_limit(nsm - reserve if 0 < reserve < nsm else 0)
            try:
                self._pipe["trunk"] = capture(trunk)
            finally:
                ops.set_sm_limit(0)
            self._pipe["heads"] = capture(heads)
            self._pipe["decode"], self._pipe["nms"] = capture(self._run_decode), capture(self._run_nms)
            self._pipe = {k: g.replay for k, g in self._pipe.items()}
        else:
            self._pipe = dict(trunk=trunk, heads=heads, decode=self._run_decode, nms=self._run_nms)

    def detect_pipelined(self, images=None, between=None):
        """Asynchronous detect: enqueues batch i and returns immediately.  self.kept / self.num_keep hold
        batch i's detections once `self.tail_done` (an event on self.tail_stream) has fired; the caller
        must consume them (e.g. enqueue a D2H copy on self.tail_stream) before the next call's tail runs.
        `between(dets, det_num) -> (dets, det_num)` runs on the side stream after decode (the multi-GPU
        all-gather hooks in here) and must then drive the NMS itself by returning None."""
        if not hasattr(self, "_pipe"):
            self._pipe_init()
        cur = torch.cuda.current_stream()
        if images is not None:
            self._set_input(images)
        self._pipe["trunk"]()
        cur.wait_event(self._ev_tail)  # the previous tail has finished reading score / box buffers
        self._pipe["heads"]()
        self._ev_fwd.record(cur)
        ts = self._tail_stream
        ts.wait_event(self._ev_fwd)
        with torch.cuda.stream(ts):
            self._pipe["decode"]()
            if between is not None:
                between()
            else:
                self._pipe["nms"]()
            self._ev_tail.record(ts)
        return self.kept, self.num_keep

    @property
    def tail_stream(self):
        return self._tail_stream

    @property
    def tail_done(self):
        return self._ev_tail

    def profile(self, iters=3):
        """Per-op device time (CUDA events on the launching stream, eager replay) with each op's
        algorithmic FLOPs / bytes.  Returns a list of dicts (name, kind, ms, flops, bytes, launches)."""
        st = torch.cuda.current_stream()
        self._run_forward()
        torch.cuda.synchronize()
        acc = [0.0] * len(self.ops)
        kernels = [None] * len(self.ops)
        for _ in range(iters):
            evs = []
            # keep the device busy while the host enqueues the whole step (a ctypes call costs more than most
            # of these kernels run): the events then bracket device time, not launch latency
            torch.cuda._sleep(40_000_000)
            for i, op in enumerate(self.ops):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                op()
                e1.record(st)
                evs.append((e0, e1))
                if self.meta[i]["kind"] in _CONV_KINDS:  # which kernel instantiation the C-side dispatch picked
                    kernels[i] = ops.last_kernel()
            torch.cuda.synchronize()
            for i, (e0, e1) in enumerate(evs):
                acc[i] += e0.elapsed_time(e1)
        return [dict({k: v for k, v in m.items() if k != "spec"}, ms=acc[i] / iters,
                     kernel=kernels[i] or _KIND_KERNEL.get(m["kind"], m["kind"])) for i, m in enumerate(self.meta)]
