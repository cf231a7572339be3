"""Tensor-level wrappers over the C ABI (torch supplies memory and streams only)."""
import ctypes as C

import torch

from . import _lib
from ._lib import ConvDesc, M3D_BF16, M3D_F32, check, lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dt(t):
    if t.dtype == torch.bfloat16:
        return M3D_BF16
    if t.dtype == torch.float32:
        return M3D_F32
    raise TypeError("unsupported dtype %s" % t.dtype)


def split_bf16(w):
    """fp32 -> (hi, lo) bf16 parts with hi + lo == w to ~2^-17 relative."""
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return hi, lo


def pack_conv_weight(weight, in_splits=None, k_pad_to=None, fp32_mode=False):
    """[Cout, Cin, R, S] fp32 -> packed [Cout, K] with K = concat_i (tap-major, channel-minor).

    in_splits: channel counts of the concatenated inputs (Root convs); each may be
    given as (c, c_padded) to zero-pad that input's channels in K.
    Returns (hi, lo|None) bf16 tensors.
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
    if fp32_mode:
        hi, lo = split_bf16(w)
        return hi.contiguous(), lo.contiguous()
    return w.to(torch.bfloat16).contiguous(), None


def conv2d_nhwc(inputs, weight, out, *, R, S, stride=1, pad=0, dil=1, Cout=None, bias=None, res=None,
                slope=1.0, weight_lo=None, om=None, sigmoid_mask=False, groups=1, in_goff=None,
                weight_goff=0, bias_goff=0, out_coff=0, out_goff=0, res_coff=0, res_goff=0,
                force_gather=False):
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
    d.Cout = Cout if Cout is not None else weight.shape[0]
    d.groups = groups
    assert weight.dtype == torch.bfloat16 and weight.is_contiguous()
    d.weight = weight.data_ptr()
    d.weight_lo = weight_lo.data_ptr() if weight_lo is not None else None
    d.weight_rows = weight.shape[0]
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
    check(lib().m3d_conv2d_nhwc(C.byref(d), _stream()))
    return out
