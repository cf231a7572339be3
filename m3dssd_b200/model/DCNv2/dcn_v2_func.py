"""DCNv2 autograd operator over the C ABI (mirror of model/DCNv2/dcn_v2_func.py).

The reference's DCNv2Function is a legacy *instance-style* autograd Function:
`DCNv2Function(stride, padding, dilation, deformable_groups)(input, offset, mask,
weight, bias)` (dcn_v2_func.py:13-38, call sites dcn_v2.py:40-41,69-70).  That
calling convention is kept; underneath it is a modern static Function whose
forward/backward call m3d_dcn_v2_forward / m3d_dcn_v2_backward.
"""
import torch
from torch.autograd import Function

from ... import ops
from ..._lib import M3D_BF16, M3D_BF16X3, M3D_F32

# "fp32":   IEEE fp32 FMA on the CUDA cores (reference accuracy; the default, what parity is judged on)
# "bf16x3": fp32 tensors, 3-part bf16 split on the tensor cores (~3e-6 per layer, much faster)
# "bf16":   bf16 operands, fp32 accumulate (throughput mode)
_PRECISION = {"fp32": M3D_F32, "bf16x3": M3D_BF16X3, "bf16": M3D_BF16}
default_precision = "fp32"


class _DCNv2Op(Function):
    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups, precision):
        if not input.is_cuda:
            raise NotImplementedError  # same contract as the reference (dcn_v2_func.py:23-24): no CPU path
        ctx.cfg = (stride, padding, dilation, deformable_groups, _PRECISION[precision])
        ctx.save_for_backward(input, offset, mask, weight, bias)
        return ops.dcn_v2_forward(input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups,
                                  _PRECISION[precision])

    @staticmethod
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        input, offset, mask, weight, bias = ctx.saved_tensors
        stride, padding, dilation, dg, prec = ctx.cfg
        gi, go, gm, gw, gb = ops.dcn_v2_backward(input, offset, mask, weight, grad_output.contiguous(), stride, padding,
                                                 dilation, dg, precision=prec)
        # (the C ABI computes in fp32; autograd wants each gradient in its tensor's dtype: the mixed-precision training
        #  path feeds bf16 activations)
        return (gi.to(input.dtype), go.to(offset.dtype), gm.to(mask.dtype), gw.to(weight.dtype), gb.to(bias.dtype),
                None, None, None, None, None)


class DCNv2Function(object):
    """Callable with the reference's constructor signature."""

    def __init__(self, stride, padding, dilation=1, deformable_groups=1, precision=None):
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.deformable_groups = deformable_groups
        self.precision = precision

    def __call__(self, input, offset, mask, weight, bias):
        return _DCNv2Op.apply(input, offset, mask, weight, bias, self.stride, self.padding, self.dilation,
                              self.deformable_groups, self.precision or default_precision)

    forward = __call__

    def _infer_shape(self, input, weight):
        n, _, h, w = input.shape
        kh, kw = weight.shape[2:4]
        ho = (h + 2 * self.padding - (self.dilation * (kh - 1) + 1)) // self.stride + 1
        wo = (w + 2 * self.padding - (self.dilation * (kw - 1) + 1)) // self.stride + 1
        return (n, weight.size(0), ho, wo)
