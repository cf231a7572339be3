"""Tensor-level wrappers over the C ABI (torch supplies memory and streams only)."""
import ctypes as C

import torch

from . import _lib
from ._lib import ConvDesc, M3D_BF16, M3D_BF16X3, M3D_F32, check, lib


LAUNCHES = 0  # kernels launched through the C ABI by this process (bench.py's gpu_launches claim for the training step)


def _count(n):
    global LAUNCHES
    LAUNCHES += n


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dt(t):
    if t.dtype == torch.bfloat16:
        return M3D_BF16
    if t.dtype == torch.float32:
        return M3D_F32
    raise TypeError("unsupported dtype %s" % t.dtype)


def split_bf16(w):
    """fp32 -> (hi, mid, lo) bf16 parts, 8 mantissa bits each: hi + mid + lo == w to ~2^-24 relative."""
    hi = w.to(torch.bfloat16)
    r1 = w - hi.float()
    mid = r1.to(torch.bfloat16)
    lo = (r1 - mid.float()).to(torch.bfloat16)
    return hi, mid, lo


def pack_conv_weight(weight, in_splits=None, k_pad_to=None, fp32_mode=False, mode=None):
    """[Cout, Cin, R, S] fp32 -> packed [Cout, K] with K = concat_i (tap-major, channel-minor).

    in_splits: channel counts of the concatenated inputs (Root convs); each may be
    given as (c, c_padded) to zero-pad that input's channels in K.
    mode "bf16" -> (bf16 weight, None); "bf16x3" (or fp32_mode=True) -> (hi, (mid, lo)) bf16 parts;
    "fp32" -> (None, fp32 weight) for the reference-accuracy path.
    """
    cout, cin, r, s = weight.shape
    if in_splits is None:
        in_splits = [cin]
    parts, c0 = [], 0
    for sp in in_splits:
        c, cp = (sp, sp) if isinstance(sp, int) else sp
        w = weight[:, c0:c0 + c].permute(0, 2, 3, 1)  # [Cout, R, S, c]
        if cp != c:
            w = torch.nn.functional.pad(w, (0, cp - c))
        parts.append(w.reshape(cout, r * s * cp))
        c0 += c
    assert c0 == cin
    w = torch.cat(parts, dim=1).contiguous().float()
    mode = mode or ("bf16x3" if fp32_mode else "bf16")
    if mode == "fp32":
        return None, w
    if mode == "bf16x3":
        hi, mid, lo = split_bf16(w)
        return hi.contiguous(), (mid.contiguous(), lo.contiguous())
    return w.to(torch.bfloat16).contiguous(), None


def k16_zero_mask(packed):
    """m3d_conv_desc.k16_zero for a packed bf16 weight matrix [rows, K]: bit j set = columns 16 j .. 16 j + 15 are zero
    in every row (the kernel may skip that k-step).  (0, 0) when K has more than 128 such slices or none is zero."""
    rows, K = packed.shape
    if K % 16 or K // 16 > 128:
        return 0, 0
    dead = (packed.reshape(rows, K // 16, 16) == 0).all(dim=2).all(dim=0).cpu().tolist()
    bits = sum(1 << j for j, z in enumerate(dead) if z)
    return bits & ((1 << 64) - 1), bits >> 64


def conv2d_nhwc(inputs, weight, out, *, R, S, stride=1, pad=0, dil=1, Cout=None, bias=None, res=None,
                slope=1.0, weight_lo=None, om=None, sigmoid_mask=False, groups=1, in_goff=None,
                weight_goff=0, bias_goff=0, out_coff=0, out_goff=0, res_coff=0, res_goff=0,
                force_gather=False, out_hw=None, k16_zero=(0, 0)):
    """inputs: list of (tensor[N,H,W,Cbuf], coff, c) or bare tensors; out: tensor[N,P,Q,Cbuf_out]."""
    d = ConvDesc()
    ins = []
    for x in inputs:
        if isinstance(x, torch.Tensor):
            x = (x, 0, x.shape[-1])
        ins.append(x)
    t0 = ins[0][0]
    d.act_dtype = _dt(t0)
    d.out_dtype = _dt(out)
    d.num_inputs = len(ins)
    for k, (t, coff, c) in enumerate(ins):
        assert t.is_cuda and t.is_contiguous() and t.dim() == 4 and t.dtype == t0.dtype
        assert t.shape[:3] == t0.shape[:3]
        d.in_[k] = t.data_ptr()
        d.in_c[k] = c
        d.in_cstride[k] = t.shape[-1]
        d.in_coff[k] = coff
        d.in_goff[k] = 0 if in_goff is None else in_goff[k]
    d.N, d.H, d.W = t0.shape[0], t0.shape[1], t0.shape[2]
    d.R, d.S, d.stride, d.pad, d.dil = R, S, stride, pad, dil
    if out_hw is not None:
        d.out_h, d.out_w = out_hw
    d.Cout = Cout if Cout is not None else (weight if weight is not None else weight_lo).shape[0]
    d.groups = groups
    if weight is not None:
        assert weight.dtype == torch.bfloat16 and weight.is_contiguous()
        d.weight = weight.data_ptr()
    if isinstance(weight_lo, torch.Tensor):  # reference-accuracy fp32 weights
        assert weight_lo.dtype == torch.float32 and weight_lo.is_contiguous()
        d.weight_f32 = weight_lo.data_ptr()
    elif weight_lo is not None:  # bf16x3: (mid, lo) parts
        d.weight_mid = weight_lo[0].data_ptr()
        d.weight_lo = weight_lo[1].data_ptr()
    d.weight_rows = (weight if weight is not None else weight_lo).shape[0]
    d.weight_goff = weight_goff
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        d.bias = bias.data_ptr()
    d.bias_goff = bias_goff
    if res is not None:
        assert res.dtype == t0.dtype and res.is_contiguous()
        d.res = res.data_ptr()
        d.res_cstride = res.shape[-1]
    d.res_coff, d.res_goff = res_coff, res_goff
    assert out.is_contiguous() and out.is_cuda
    d.out = out.data_ptr()
    d.out_cstride = out.shape[-1]
    d.out_coff, d.out_goff = out_coff, out_goff
    d.slope = slope
    if om is not None:
        assert om.dtype == torch.float32 and om.is_contiguous()
        d.om = om.data_ptr()
        d.om_cstride = om.shape[-1]
    d.sigmoid_mask = int(sigmoid_mask)
    d.force_gather = int(force_gather)
    d.k16_zero[0], d.k16_zero[1] = int(k16_zero[0]), int(k16_zero[1])
    check(lib().m3d_conv2d_nhwc(C.byref(d), _stream()))
    _count(1)
    return out


def last_kernel():
    """Kernel instantiation the last conv2d_nhwc call on this thread dispatched to (e.g. 'conv_halo2_kernel<128>')."""
    return lib().m3d_last_kernel().decode()


# --------------------------------------------------------------------------- other ops
def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stem_conv7x7(image, weight, bias, out, slope=0.01):
    """image [N,3,H,W] fp32 NCHW; weight [16,3,7,7] fp32 (BN folded); out NHWC [N,H,W,>=16]."""
    assert image.dtype == torch.float32 and image.is_contiguous() and image.shape[1] == 3
    N, _, H, W = image.shape
    check(lib().m3d_stem_conv7x7(_p(image), _p(weight), _p(bias), _p(out), _dt(out), out.shape[-1], N, H, W,
                                 float(slope), _stream()))
    return out


def preprocess_u8(image_hwc, out_nchw, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), swap_rb=True):
    """uint8 [N,H,W,3] (cv2 BGR) CUDA tensor -> normalised fp32 NCHW (RGB): the reference's Normalize transform and
    channel swap (lib/augmentations.py:44-57, lib/dataloader.py:942-950) on the device."""
    assert image_hwc.dtype == torch.uint8 and image_hwc.is_cuda and image_hwc.is_contiguous() and image_hwc.shape[-1] == 3
    N, H, W, _ = image_hwc.shape
    assert tuple(out_nchw.shape) == (N, 3, H, W) and out_nchw.dtype == torch.float32 and out_nchw.is_contiguous()
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    check(lib().m3d_preprocess_u8(_p(image_hwc), _p(out_nchw), N, H, W, m, s, int(bool(swap_rb)), _stream()))
    return out_nchw


def pack_ragged_u8(images, pin=True):
    """Host side of preprocess_u8_pad: uint8 HWC images of different sizes (numpy arrays or CPU tensors, as cv2.imread
    returns them) -> one (pinned) byte buffer + (offsets, heights, widths) lists."""
    arrs = [torch.as_tensor(a) for a in images]
    for a in arrs:
        if a.dtype != torch.uint8 or a.dim() != 3 or a.shape[-1] != 3:
            raise ValueError("every image must be uint8 [h, w, 3], got %s %s" % (a.dtype, tuple(a.shape)))
    sizes = [a.numel() for a in arrs]
    offsets, total = [], 0
    for n in sizes:
        offsets.append(total)
        total += n
    # (torch's caching host allocator: after the first call a pinned block of this size is reused, no cudaHostAlloc)
    buf = torch.empty(max(total, 1), dtype=torch.uint8, pin_memory=bool(pin and torch.cuda.is_available()))
    for a, o, n in zip(arrs, offsets, sizes):
        buf[o:o + n] = a.contiguous().view(-1)
    return buf, offsets, [int(a.shape[0]) for a in arrs], [int(a.shape[1]) for a in arrs]


def preprocess_u8_pad(packed, offsets, heights, widths, out_nchw, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225),
                      swap_rb=True):
    """The reference's test-time transform Preprocess(size, mean, stds) (lib/augmentations.py:472-492: ConvertToFloat,
    Padding = zero border on the bottom / right, Normalize) + BGR->RGB + CHW (lib/dataloader.py:942-950) for a ragged
    batch on the device.  packed: 1-D uint8 CUDA tensor holding image n (HWC) at offsets[n]; out_nchw [N,3,H,W] fp32.
    Raises M3DError for an image larger than H x W (cv2.copyMakeBorder raises there)."""
    assert packed.dtype == torch.uint8 and packed.is_cuda and packed.is_contiguous() and packed.dim() == 1
    N, _, H, W = out_nchw.shape
    assert len(offsets) == len(heights) == len(widths) == N and out_nchw.shape[1] == 3
    assert out_nchw.dtype == torch.float32 and out_nchw.is_contiguous() and out_nchw.is_cuda
    for o, h, w in zip(offsets, heights, widths):
        assert 0 <= o and o + h * w * 3 <= packed.numel(), "image outside the packed buffer"
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    off = (C.c_longlong * max(N, 1))(*[int(v) for v in offsets])
    hh = (C.c_int * max(N, 1))(*[int(v) for v in heights])
    ww = (C.c_int * max(N, 1))(*[int(v) for v in widths])
    check(lib().m3d_preprocess_u8_pad(_p(packed), off, hh, ww, _p(out_nchw), N, H, W, m, s, int(bool(swap_rb)), _stream()))
    return out_nchw


def pack_stem_s2d(w, b):
    """Stem weights [16,3,7,7] (BN folded) + bias [16] -> bf16 [64, 192] and bias [64] for m3d_stem_conv7x7_s2d:
    output channel (ey*2+ex)*16 + co, K index c*64 + r8*8 + s8 over the 8x8 stride-2 window."""
    co, ci, kh, kw = w.shape
    assert (co, ci, kh, kw) == (16, 3, 7, 7)
    wp = torch.zeros(2, 2, co, ci, 8, 8, dtype=torch.float32)
    for ey in range(2):
        for ex in range(2):
            wp[ey, ex, :, :, ey:ey + 7, ex:ex + 7] = w.float().cpu()
    return wp.reshape(4 * co, ci * 64).to(torch.bfloat16).contiguous(), b.float().cpu().repeat(4).contiguous()


def s2d_conv3x3_weight(w, b):
    """3x3 / stride 1 / pad 1 conv [Co, Ci, 3, 3] acting on a 2x2 space-to-depth tensor: the equivalent dense
    3x3 conv [4*Co, 4*Ci, 3, 3] (75 % zeros) whose in/out channels are (dy*2+dx)*C + c."""
    co, ci = w.shape[:2]
    wp = torch.zeros(2, 2, co, 2, 2, ci, 3, 3, dtype=torch.float32)  # [ey, ex, co, dy, dx, ci, R, S]
    w = w.float().cpu()
    for ey in range(2):
        for dy in range(2):
            for R in (-1, 0, 1):
                r = 2 * R + dy - ey
                if not -1 <= r <= 1:
                    continue
                for ex in range(2):
                    for dx in range(2):
                        for S in (-1, 0, 1):
                            s = 2 * S + dx - ex
                            if -1 <= s <= 1:
                                wp[ey, ex, :, dy, dx, :, R + 1, S + 1] = w[:, :, r + 1, s + 1]
    return wp.reshape(4 * co, 4 * ci, 3, 3), b.float().cpu().repeat(4)


def s2d_conv3x3_s2_weight(w):
    """3x3 / stride 2 / pad 1 conv [Co, Ci, 3, 3] reading a 2x2 space-to-depth tensor and writing a plain one:
    the equivalent 2x2 conv [Co, 4*Ci, 2, 2] with top/left padding 1 (taps R, S in {-1, 0})."""
    co, ci = w.shape[:2]
    wp = torch.zeros(co, 2, 2, ci, 2, 2, dtype=torch.float32)  # [co, dy, dx, ci, R+1, S+1]
    w = w.float().cpu()
    for dy in range(2):
        for R in (-1, 0):
            r = 2 * R + dy
            if not -1 <= r <= 1:
                continue
            for dx in range(2):
                for S in (-1, 0):
                    s = 2 * S + dx
                    if -1 <= s <= 1:
                        wp[:, dy, dx, :, R + 1, S + 1] = w[:, :, r + 1, s + 1]
    return wp.reshape(co, 4 * ci, 2, 2)


def stem_conv7x7_s2d(image, weight, bias, out, slope=0.01):
    """image [N,3,H,W] fp32 NCHW; weight/bias from pack_stem_s2d; out bf16 [N,H/2,W/2,64]."""
    assert image.dtype == torch.float32 and image.is_contiguous() and image.shape[1] == 3
    assert out.dtype == torch.bfloat16 and out.shape[-1] == 64
    N, _, H, W = image.shape
    check(lib().m3d_stem_conv7x7_s2d(_p(image), _p(weight), _p(bias), _p(out), N, H, W, float(slope), _stream()))
    return out


def maxpool2x2(x, out, C_=None):
    N, H, W, cs = x.shape
    check(lib().m3d_maxpool2x2_nhwc(_p(x), _p(out), _dt(x), N, H, W, C_ or cs, cs, out.shape[-1], _stream()))
    return out


def upsample_add(x, weight, skip, out, f=2):
    """depthwise ConvTranspose2d(2f, stride f, pad f//2) of x plus skip; weight tap-major [(2f)^2, C] fp32
    (pack with pack_upsample_weight)."""
    N, H, W, cs = x.shape
    Cc = weight.shape[1]
    assert weight.shape[0] == 4 * f * f and weight.is_contiguous()
    check(lib().m3d_upsample_add_nhwc(_p(x), _p(weight), _p(skip), _p(out), _dt(x), N, H, W, Cc, f, cs,
                                      skip.shape[-1] if skip is not None else 0, out.shape[-1], _stream()))
    return out


def pack_upsample_weight(w):
    """ConvTranspose2d weight [C, 1, k, k] -> tap-major [k*k, C] fp32."""
    C_ = w.shape[0]
    return w.detach().float().reshape(C_, -1).t().contiguous()


def cls_softmax(logits, A, K, cls_out, prob_out, fg_max, fg_arg, score, cls_pred):
    N, H, W, cs = logits.shape
    check(lib().m3d_cls_softmax(_p(logits), cs, N, H, W, A, K, _p(cls_out), _p(prob_out), _p(fg_max), _p(fg_arg),
                                _p(score), _p(cls_pred), _stream()))


def cls_softmax_shape_om(logits, A, K, cls_out, prob_out, fg_max, fg_arg, score, cls_pred, anchors, feat_stride, thresh, om):
    """cls_softmax + shape_align_om in one launch (identical values)."""
    N, H, W, cs = logits.shape
    check(lib().m3d_cls_softmax_shape_om(_p(logits), cs, N, H, W, A, K, _p(cls_out), _p(prob_out), _p(fg_max), _p(fg_arg),
                                         _p(score), _p(cls_pred), _p(anchors), anchors.shape[1], float(feat_stride),
                                         float(thresh), _p(om), _stream()))


def center_align_om2(fg_max, fg_arg, heads, coffs4, anchors, feat_stride, mean4, std4, thresh, om_a, om_b):
    """The 2D- and 3D-centre offset builders in one launch (identical values to two center_align_om calls)."""
    assert om_a.shape == om_b.shape
    check(lib().m3d_center_align_om2(_p(fg_max), _p(fg_arg), _p(heads), heads.shape[-1], (C.c_int * 4)(*[int(v) for v in coffs4]),
                                     (C.c_float * 4)(*[float(v) for v in mean4]), (C.c_float * 4)(*[float(v) for v in std4]),
                                     _p(anchors), anchors.shape[1], float(feat_stride), float(thresh), _p(om_a), _p(om_b),
                                     om_a.shape[-1], fg_max.numel(), _stream()))


def shape_align_om(fg_max, fg_arg, anchors, feat_stride, thresh, om):
    check(lib().m3d_shape_align_om(_p(fg_max), _p(fg_arg), _p(anchors), anchors.shape[1], float(feat_stride),
                                   float(thresh), _p(om), fg_max.numel(), _stream()))


def center_align_om(fg_max, fg_arg, heads, x_coff, y_coff, anchors, feat_stride, mean_xy, std_xy, thresh, om):
    check(lib().m3d_center_align_om(_p(fg_max), _p(fg_arg), _p(heads), heads.shape[-1], x_coff, y_coff, _p(anchors),
                                    anchors.shape[1], float(feat_stride), float(mean_xy[0]), float(mean_xy[1]),
                                    float(std_xy[0]), float(std_xy[1]), float(thresh), _p(om), om.shape[-1],
                                    fg_max.numel(), _stream()))


def refine_3d(kept, num_keep, p2, score_thresh=0.75, hill_climbing=True, step_r_init=None, r_lim=0.01):
    """Post-NMS 3D refinement of kept[B, max_out, >=13] (fp32, device) with projection matrices p2 [B, 4, 4] or [4, 4]
    (any float dtype / device): returns (out float64 [B, max_out, 14], valid int32 [B, max_out]) on the device."""
    import math
    B, max_out, row_len = kept.shape
    p2 = torch.as_tensor(p2, dtype=torch.float64).cpu().reshape(-1, 4, 4)
    if p2.shape[0] == 1 and B > 1:
        p2 = p2.expand(B, 4, 4)
    assert p2.shape[0] == B
    p2_inv = torch.from_numpy(__import__("numpy").linalg.inv(p2.numpy()))  # the reference inverts with numpy on the host
    p2d, p2i = p2.contiguous().to(kept.device), p2_inv.contiguous().to(kept.device)
    out = torch.empty(B, max_out, 14, dtype=torch.float64, device=kept.device)
    valid = torch.empty(B, max_out, dtype=torch.int32, device=kept.device)
    step = 0.3 * math.pi if step_r_init is None else float(step_r_init)
    check(lib().m3d_refine_3d(_p(kept), _p(num_keep), B, max_out, row_len, _p(p2d), _p(p2i), float(score_thresh),
                              int(bool(hill_climbing)), step, float(r_lim), _p(out), _p(valid), _stream()))
    return out, valid


def set_sm_limit(sms):
    """SMs the persistent kernels launched / captured from now on may use (0 = all)."""
    check(lib().m3d_set_sm_limit(int(sms)))


def set_pdl(on):
    """Programmatic dependent launch for the kernels launched / captured from now on (see m3d_set_pdl)."""
    check(lib().m3d_set_pdl(int(bool(on))))


def head_mlp(x, x_coff, cx, w1, b1, w2, b2, w3, b3, G, A, rows3, out, out_coff, slope=0.01):
    """G fused three-layer 1x1 heads on x[..., x_coff:x_coff+cx] (bf16 NHWC) -> out[..., out_coff + g*A + a] (fp32 NHWC)."""
    N, H, W, cs = x.shape
    assert x.dtype == torch.bfloat16 and out.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous()
    assert w1.shape == (G * 256, cx) and w2.shape == (G * 256, 256) and w3.shape == (G * rows3, 256)
    check(lib().m3d_head_mlp(_p(x), N, H, W, cs, x_coff, cx, _p(w1), _p(b1), _p(w2), _p(b2), _p(w3), _p(b3), G, A, rows3,
                             _p(out), out.shape[-1], out_coff, float(slope), _stream()))


def flatten_heads(heads, A, slots, bbox_2d, bbox_3d):
    N, H, W, cs = heads.shape
    arr = (C.c_int * 11)(*slots)
    check(lib().m3d_flatten_heads(_p(heads), cs, N, H, W, A, arr, _p(bbox_2d), _p(bbox_3d), _stream()))


def nchw_to_nhwc(x, out, coff=0):
    N, Cc, H, W = x.shape
    check(lib().m3d_nchw_to_nhwc(_p(x), _dt(x), _p(out), _dt(out), N, Cc, H, W, out.shape[-1], coff, _stream()))
    return out


def nhwc_to_nchw(x, out, coff=0):
    N, Cc, H, W = out.shape
    check(lib().m3d_nhwc_to_nchw(_p(x), _dt(x), _p(out), _dt(out), N, Cc, H, W, x.shape[-1], coff, _stream()))
    return out


def decode_topk_workspace(batch, device):
    """Scratch of the multi-CTA top-K for `batch` images: owned by the caller (the engine keeps one per instance)."""
    return torch.zeros(lib().m3d_decode_topk_workspace(batch), dtype=torch.uint8, device=device)


def decode_topk(score, cls_pred, bbox_2d, bbox_3d, anchors, means, stds, A, H, W, feat_stride, scale_factor, topk,
                dets, det_idx, det_num, workspace=None):
    B = score.shape[0]
    if workspace is None:
        workspace = decode_topk_workspace(B, score.device)
    means = (C.c_float * 11)(*[float(v) for v in means])  # host arrays in the C ABI
    stds = (C.c_float * 11)(*[float(v) for v in stds])
    check(lib().m3d_decode_topk(_p(score), _p(cls_pred), _p(bbox_2d), _p(bbox_3d), _p(anchors), means, stds,
                                B, A, H, W, float(feat_stride), float(scale_factor), topk, _p(dets), _p(det_idx),
                                _p(det_num), _p(workspace), workspace.numel(), _stream()))


def decode_topk_heads(score, cls_pred, heads, slots, anchors, means, stds, A, H, W, feat_stride, scale_factor, topk,
                      dets, det_idx, det_num, workspace=None):
    """decode_topk with the regression outputs read from the NHWC head buffer [B,H,W,>=11*A] (output j of anchor a =
    channel slots[j]*A + a): the detection path skips flatten_heads; same results bit for bit."""
    B = score.shape[0]
    assert heads.dtype == torch.float32 and heads.is_contiguous() and tuple(heads.shape[:3]) == (B, H, W)
    if workspace is None:
        workspace = decode_topk_workspace(B, score.device)
    means = (C.c_float * 11)(*[float(v) for v in means])
    stds = (C.c_float * 11)(*[float(v) for v in stds])
    sl = (C.c_int * 11)(*[int(v) for v in slots])
    check(lib().m3d_decode_topk_heads(_p(score), _p(cls_pred), _p(heads), heads.shape[-1], sl, _p(anchors), means, stds,
                                      B, A, H, W, float(feat_stride), float(scale_factor), topk, _p(dets), _p(det_idx),
                                      _p(det_num), _p(workspace), workspace.numel(), _stream()))


def nms_workspace_bytes(batch, max_n):
    return lib().m3d_nms_workspace_bytes(batch, max_n)


def nms_batched(boxes, num, thresh, workspace, keep, num_keep):
    """boxes [B, max_n, stride] fp32 (x1,y1,x2,y2,...), sorted by score within each image."""
    B, max_n, stride = boxes.shape
    check(lib().m3d_nms_batched(_p(boxes), stride, _p(num), B, max_n, float(thresh), _p(workspace),
                                workspace.numel() * workspace.element_size(), _p(keep), _p(num_keep), _stream()))


def gather_kept(dets, keep, num_keep, max_out, out):
    B, max_n, row = dets.shape
    check(lib().m3d_gather_kept(_p(dets), row, B, max_n, _p(keep), _p(num_keep), max_out, _p(out), _stream()))


def dcn_v2_forward(input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups, precision=M3D_F32):
    """NCHW fp32 CUDA tensors -> NCHW fp32 output (DCNv2Function.forward, model/DCNv2/dcn_v2_func.py:22-38)."""
    B, Cin, H, W = input.shape
    Cout, Ck, kh, kw = weight.shape
    if Ck != Cin:
        raise RuntimeError("Input shape and kernel channels wont match: (%d vs %d)." % (Cin, Ck))
    Ho = (H + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
    if tuple(offset.shape) != (B, 2 * deformable_groups * kh * kw, Ho, Wo) or \
            tuple(mask.shape) != (B, deformable_groups * kh * kw, Ho, Wo):
        raise RuntimeError("offset/mask shape does not match the output size %dx%d" % (Ho, Wo))
    if Cin % deformable_groups != 0:
        raise RuntimeError("input channels (%d) must be a multiple of deformable_groups (%d)" % (Cin, deformable_groups))
    ws_bytes = lib().m3d_dcn_v2_forward_workspace(B, Cin, H, W, Cout, kh, kw, stride, padding, dilation,
                                                  deformable_groups, precision)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=input.device)
    out = torch.empty(B, Cout, Ho, Wo, dtype=torch.float32, device=input.device)
    args = [t.contiguous().float() for t in (input, weight, bias, offset, mask)]
    check(lib().m3d_dcn_v2_forward(_p(args[0]), _p(args[1]), _p(args[2]), _p(args[3]), _p(args[4]), _p(out), B, Cin, H,
                                   W, Cout, kh, kw, stride, stride, padding, padding, dilation, dilation,
                                   deformable_groups, precision, _p(ws), ws_bytes, _stream()))
    _count(6 * deformable_groups)  # 3 layout conversions in, weight pack, fused gather + GEMM, layout conversion out
    return out


def anab_pool(kvs, ck, cv, sizes, ktok, vtok, workspace=None):
    N, H, W, cs = kvs.shape
    arr = (C.c_int * len(sizes))(*sizes)
    need = lib().m3d_anab_pool_workspace(N, H, len(sizes), arr, ck, cv)
    if workspace is None:
        workspace = torch.empty(need, dtype=torch.uint8, device=kvs.device)
    check(lib().m3d_anab_pool(_p(kvs), cs, N, H, W, ck, cv, len(sizes), arr, _p(workspace), workspace.numel(), _p(ktok),
                              _p(vtok), _stream()))
    return workspace


def anab_pool_workspace(N, H, sizes, ck, cv):
    arr = (C.c_int * len(sizes))(*sizes)
    return lib().m3d_anab_pool_workspace(N, H, len(sizes), arr, ck, cv)


def anab_attention_workspace(N, x):
    """bf16 token operands of the tensor-core attention kernel (caller-owned; empty in fp32 mode)."""
    return torch.zeros(max(1, lib().m3d_anab_attention_workspace(N, _dt(x))), dtype=torch.uint8, device=x.device)


def anab_attention(q, ktok, vtok, x, scale, shift, slope, out, ck, cv, workspace=None):
    N, H, W, _ = x.shape
    T = ktok.shape[1]
    if workspace is None:
        workspace = anab_attention_workspace(N, x)
    check(lib().m3d_anab_attention(_p(q), q.shape[-1], _p(ktok), _p(vtok), _p(x), x.shape[-1], _dt(x), _p(scale),
                                   _p(shift), float(slope), _p(out), out.shape[-1], N, H * W, ck, cv, T, _p(workspace),
                                   workspace.numel(), _stream()))
    return out


def dcn_v2_backward(input, offset, mask, weight, grad_output, stride, padding, dilation, deformable_groups,
                    precision=M3D_F32):
    """DCNv2Function.backward (model/DCNv2/dcn_v2_func.py:40-62): returns (grad_input, grad_offset, grad_mask,
    grad_weight, grad_bias), fp32 NCHW CUDA tensors."""
    B, Cin, H, W = input.shape
    Cout, _, kh, kw = weight.shape
    args = [t.contiguous().float() for t in (input, weight, offset, mask, grad_output)]
    gi, go, gm = torch.empty_like(args[0]), torch.empty_like(args[2]), torch.empty_like(args[3])
    gw = torch.empty_like(args[1])
    gb = torch.empty(Cout, dtype=torch.float32, device=input.device)
    n = lib().m3d_dcn_v2_backward_workspace(B, Cin, H, W, Cout, kh, kw, stride, padding, dilation, deformable_groups)
    ws = torch.empty(n, dtype=torch.uint8, device=input.device)
    check(lib().m3d_dcn_v2_backward(*[_p(t) for t in args], _p(gi), _p(gw), _p(gb), _p(go), _p(gm), B, Cin, H, W, Cout,
                                    kh, kw, stride, stride, padding, padding, dilation, dilation, deformable_groups,
                                    precision, _p(ws), n, _stream()))
    _count(12 * deformable_groups)  # layout conversions, W^T dY GEMM, coordinate / input gradient kernels, dW, bias, repack
    return gi, go, gm, gw, gb


def conv2d_wgrad(x, gy, Cin, Cout, R, S, stride=1, pad=0, dil=1, x_coff=0, gy_coff=0):
    """Weight gradient of a convolution (training path): x [N,H,W,Cx] and gy [N,P,Q,Cy] bf16 NHWC CUDA tensors (channel
    strides multiples of 8; channels [x_coff, x_coff+Cin) / [gy_coff, gy_coff+Cout) are used) -> dW fp32 [Cout,Cin,R,S]."""
    assert x.dtype == torch.bfloat16 and gy.dtype == torch.bfloat16 and x.is_contiguous() and gy.is_contiguous()
    N, H, W, cx = x.shape
    _, P, Q, cy = gy.shape
    dw = torch.empty(Cout, Cin, R, S, dtype=torch.float32, device=x.device)
    n = lib().m3d_conv2d_wgrad_workspace(N, P, Q, Cin, Cout, R, S)
    ws = torch.empty(n, dtype=torch.uint8, device=x.device)
    check(lib().m3d_conv2d_wgrad(_p(x), cx, x_coff, _p(gy), cy, gy_coff, _p(dw), N, H, W, Cin, P, Q, Cout, R, S, stride, pad,
                                 dil, _p(ws), n, _stream()))
    _count(2)  # wgrad + slice reduction
    return dw


def channel_sum(x_nhwc, C, coff=0):
    """Per-channel sum over all pixels of a bf16 NHWC tensor -> fp32 [C] (the bias gradient of a convolution)."""
    assert x_nhwc.dtype == torch.bfloat16 and x_nhwc.is_contiguous()
    cs = x_nhwc.shape[-1]
    npix = x_nhwc.numel() // cs
    out = torch.empty(C, dtype=torch.float32, device=x_nhwc.device)
    n = lib().m3d_channel_sum_workspace(C)
    ws = torch.empty(n, dtype=torch.uint8, device=x_nhwc.device)
    check(lib().m3d_channel_sum(_p(x_nhwc), npix, C, cs, coff, _p(out), _p(ws), n, _stream()))
    _count(2)
    return out


def upsample_backward(gy, x, weight, f):
    """Backward of upsample_add without the skip (training path): gy [N,H*f,W*f,C], x [N,H,W,C] bf16 NHWC dense,
    weight tap-major [(2f)^2, C] fp32 -> (gx bf16 [N,H,W,C], gw fp32 [(2f)^2, C])."""
    N, H, W, Cc = x.shape
    assert gy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and gy.is_contiguous() and x.is_contiguous()
    gx = torch.empty_like(x)
    gw = torch.empty(4 * f * f, Cc, dtype=torch.float32, device=x.device)
    n = lib().m3d_upsample_backward_workspace(Cc, f)
    ws = torch.empty(n, dtype=torch.uint8, device=x.device)
    check(lib().m3d_upsample_backward(_p(gy), _p(x), _p(weight), _p(gx), _p(gw), N, H, W, Cc, f, _p(ws), n, _stream()))
    _count(2)
    return gx, gw
