"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference Python modules from
/root/reference (build container only; the checkout does not exist on the GPU box).

The reference's model code runs on CPU once the packages missing from this
image are stubbed in sys.modules and its CUDA-only DCNv2 extension
(model/DCNv2/dcn_v2_func.py:23-24 raises NotImplementedError on CPU tensors) is
replaced by a module with the same classes whose arithmetic is the CPU oracle.
Used to (a) validate oracle/ref_model.py, (b) generate tests/golden/*.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch
from torch import nn

REF = os.environ.get("M3D_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "model"))


class _EasyDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


class _Permissive(types.ModuleType):
    """Module whose unknown attributes read as 0 (default-argument constants such as cv2.FONT_*)."""

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return 0


def _stub(name, **attrs):
    m = _Permissive(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    sys.modules[name] = m
    return m


def _dcn_module(dcn_impl):
    """Stand-in for model.DCNv2.dcn_v2 with the reference's constructors/forward
    (model/DCNv2/dcn_v2.py:14-70); only the FFI call is replaced."""
    import math
    from torch.nn.modules.utils import _pair

    class DCNv2(nn.Module):
        def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
            super().__init__()
            self.in_channels, self.out_channels = in_channels, out_channels
            self.kernel_size = _pair(kernel_size)
            self.stride, self.padding, self.dilation = stride, padding, dilation
            self.deformable_groups = deformable_groups
            self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels, *self.kernel_size))
            self.bias = nn.Parameter(torch.Tensor(out_channels))
            n = in_channels * self.kernel_size[0] * self.kernel_size[1]
            self.weight.data.uniform_(-1. / math.sqrt(n), 1. / math.sqrt(n))
            self.bias.data.zero_()

        def forward(self, input, offset, mask):
            return dcn_impl(input, offset, mask, self.weight, self.bias, self.stride, self.padding, self.dilation,
                            self.deformable_groups)

    class DCN(DCNv2):
        def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
            super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, deformable_groups)
            self.conv_offset_mask = nn.Conv2d(in_channels, deformable_groups * 3 * self.kernel_size[0] * self.kernel_size[1],
                                              kernel_size=self.kernel_size, stride=(stride, stride),
                                              padding=(padding, padding), bias=True)
            self.conv_offset_mask.weight.data.zero_()
            self.conv_offset_mask.bias.data.zero_()

        def forward(self, input):
            out = self.conv_offset_mask(input)
            o1, o2, mask = torch.chunk(out, 3, dim=1)
            offset = torch.cat((o1, o2), dim=1)
            mask = torch.sigmoid(mask)
            return dcn_impl(input, offset, mask, self.weight, self.bias, self.stride, self.padding, self.dilation,
                            self.deformable_groups)

    return _stub("model.DCNv2.dcn_v2", DCNv2=DCNv2, DCN=DCN)


def dcn_oracle_c(input, offset, mask, weight, bias, stride, padding, dilation, dg):
    from . import oracle as O
    out = O.dcn_v2_forward(input.detach().numpy(), offset.detach().numpy(), mask.detach().numpy(),
                           weight.detach().numpy(), bias.detach().numpy(), stride, padding, dilation, dg)
    return torch.from_numpy(out)


def dcn_torchvision(input, offset, mask, weight, bias, stride, padding, dilation, dg):
    import torchvision
    return torchvision.ops.deform_conv2d(input, offset, weight, bias, stride=stride, padding=padding,
                                         dilation=dilation, mask=mask)


_loaded = {}


def load_reference(dcn_impl=dcn_torchvision):
    """Returns a namespace with the reference's modules (model.*, lib.rpn_util, py_cpu_nms)."""
    key = dcn_impl.__name__
    if key in _loaded:
        return _loaded[key]
    assert available(), "reference checkout not found at %s" % REF
    for m in [k for k in sys.modules if k == "model" or k.startswith("model.") or k == "lib" or k.startswith("lib.")]:
        del sys.modules[m]
    _stub("easydict", EasyDict=_EasyDict)
    _stub("shapely")
    _stub("shapely.geometry", Polygon=object)
    mpl = _stub("matplotlib", use=lambda *a, **k: None)
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    _stub("matplotlib.backends")
    _stub("matplotlib.backends.backend_agg", FigureCanvasAgg=object)
    _stub("matplotlib.patches")
    _stub("mpl_toolkits")
    _stub("mpl_toolkits.mplot3d", Axes3D=object)
    _stub("skimage")
    _stub("skimage.io")
    _stub("tensorboardX", SummaryWriter=object)
    _stub("fire")
    _stub("cv2")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    importlib.invalidate_caches()
    # numba-CUDA initialisation in lib.eval.eval fails without a driver
    _stub("lib.eval.eval", get_official_eval_result=lambda *a, **k: None)
    from . import oracle as O
    _stub("lib.nms.gpu_nms", gpu_nms=O.gpu_nms)
    import model  # noqa: F401  (namespace package rooted at REF)
    _dcn_module(dcn_impl)
    ns = types.SimpleNamespace()
    try:
        ns.rpn_util = importlib.import_module("lib.rpn_util")
    except Exception:
        # lib.rpn_util star-imports helpers that need cv2/etc.; retry with more stubs if needed
        raise
    ns.pose_dla_dcn = importlib.import_module("model.pose_dla_dcn")
    ns.attention = importlib.import_module("model.module.attention")
    ns.align = importlib.import_module("model.module.feturealign_mgpu")
    ns.rpn = importlib.import_module("model.M3d_inference_align")
    sys.path.insert(0, os.path.join(REF, "lib", "nms"))
    ns.py_cpu_nms = importlib.import_module("py_cpu_nms").py_cpu_nms
    ns.EasyDict = _EasyDict
    _loaded[key] = ns
    return ns
