"""GPU parity of the DCNv2 operator behind the reference's Python surface (DCNv2 / DCN /
DCNv2Function) against the C oracle; plus the reference's own known-answer test and the
reference's unmodified im2col CUDA kernel (oracle/_ref) against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_check_zero_offset_known_answer():
    """model/DCNv2/test.py:32-65, verbatim procedure on our modules: |x - 2*DCNv2(x)| < 1e-10."""
    from torch import nn
    from m3dssd_b200.model.DCNv2.dcn_v2 import DCNv2
    N, inC, inH, inW, outC, kH, kW = 2, 2, 4, 4, 2, 3, 3
    conv_offset = nn.Conv2d(inC, 2 * kH * kW, (kH, kW), stride=1, padding=1).cuda()
    conv_mask = nn.Conv2d(inC, kH * kW, (kH, kW), stride=1, padding=1).cuda()
    dcn = DCNv2(inC, outC, (kH, kW), stride=1, padding=1, dilation=1, deformable_groups=1).cuda()
    for m in (conv_offset, conv_mask):
        m.weight.data.zero_()
        m.bias.data.zero_()
    dcn.weight.data.zero_()
    dcn.bias.data.zero_()
    for p in range(inC):
        dcn.weight.data[p, p, kH // 2, kW // 2] = 1.0
    x = torch.randn(N, inC, inH, inW).cuda()
    with torch.no_grad():
        out = dcn(x, conv_offset(x), torch.sigmoid(conv_mask(x)))
    assert (x - 2 * out).abs().max().item() < 1e-10


@pytest.mark.parametrize("shape", [
    (2, 64, 12, 40, 64, 3, 1, 1), (1, 128, 24, 80, 128, 3, 1, 1), (2, 2, 4, 4, 2, 3, 1, 1),
    (1, 128, 12, 40, 128, 1, 1, 0), (1, 70, 9, 11, 30, 3, 2, 1), (1, 512, 6, 20, 256, 3, 1, 1)])
@pytest.mark.parametrize("precision,tol", [("fp32", 3e-6), ("bf16x3", 3e-5), ("bf16", 1.2e-2)])
def test_dcn_v2_function_vs_oracle(shape, precision, tol):
    from m3dssd_b200.model.DCNv2.dcn_v2_func import DCNv2Function
    B, Cin, H, W, Cout, k, s, p = shape
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal((B, Cin, H, W)).astype(np.float32)
    Ho, Wo = O.dcn_out_shape(H, W, k, k, s, p, 1)
    off = (rng.standard_normal((B, 2 * k * k, Ho, Wo)) * 2.5).astype(np.float32)
    m = rng.random((B, k * k, Ho, Wo)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    ref = O.dcn_v2_forward(x, off, m, w, b, s, p, 1, 1)
    fn = DCNv2Function(s, p, 1, 1, precision=precision)
    with torch.no_grad():
        out = fn(*[torch.from_numpy(a).cuda() for a in (x, off, m, w, b)])
    err = np.abs(out.cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1.0)
    assert out.shape == ref.shape and err < tol, err


def test_dcn_module_matches_oracle_and_errors():
    from m3dssd_b200.model.DCNv2.dcn_v2 import DCN, DCNv2
    torch.manual_seed(0)
    dcn = DCN(64, 64, 3, 1, 1).cuda()
    dcn.conv_offset_mask.weight.data.normal_(0, 0.05)
    dcn.conv_offset_mask.bias.data.normal_(0, 0.2)
    x = torch.randn(2, 64, 10, 14).cuda()
    with torch.no_grad():
        y = dcn(x)
        om = dcn.conv_offset_mask(x)
    off, mask = om[:, :18].cpu().numpy(), torch.sigmoid(om[:, 18:]).cpu().numpy()
    ref = O.dcn_v2_forward(x.cpu().numpy(), off, mask, dcn.weight.detach().cpu().numpy(),
                           dcn.bias.detach().cpu().numpy(), 1, 1, 1, 1)
    assert np.abs(y.cpu().numpy() - ref).max() < 3e-6 * max(np.abs(ref).max(), 1)
    # error behaviour of the reference surface
    cpu_dcn = DCNv2(4, 4, 3, 1, 1)
    with pytest.raises(NotImplementedError):  # dcn_v2_func.py:23-24
        cpu_dcn(torch.randn(1, 4, 5, 5), torch.zeros(1, 18, 5, 5), torch.ones(1, 9, 5, 5))
    with pytest.raises(RuntimeError):  # channel mismatch -> THError in dcn_v2_cuda.c:36-38
        dcn(torch.randn(1, 32, 8, 8).cuda())


def test_oracle_im2col_vs_reference_cuda_kernel():
    """modulated_deformable_im2col_cuda from the reference's unmodified dcn_v2_im2col_cuda.cu
    (oracle/_ref) on the GPU vs the oracle's CPU restatement: pins the sampling arithmetic."""
    path = os.path.join(ROOT, "oracle", "_ref", "libref_dcn_im2col.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_dcn_im2col.so not built")
    ref = C.CDLL(path)
    fn = ref.modulated_deformable_im2col_cuda
    fn.restype = None
    fn.argtypes = [C.c_void_p] * 4 + [C.c_int] * 15 + [C.c_void_p]
    rng = np.random.default_rng(5)
    for (Cc, H, W, k, s, p) in ((16, 12, 17, 3, 1, 1), (8, 9, 11, 3, 2, 1), (12, 7, 9, 1, 1, 0)):
        Ho, Wo = O.dcn_out_shape(H, W, k, k, s, p, 1)
        x = rng.standard_normal((Cc, H, W)).astype(np.float32)
        off = (rng.standard_normal((2 * k * k, Ho, Wo)) * 3).astype(np.float32)
        m = rng.random((k * k, Ho, Wo)).astype(np.float32)
        xd, od, md = (torch.from_numpy(a).cuda() for a in (x, off, m))
        col = torch.zeros(Cc * k * k, Ho, Wo, device="cuda")
        fn(None, xd.data_ptr(), od.data_ptr(), md.data_ptr(), 1, Cc, H, W, Ho, Wo, k, k, p, p, s, s, 1, 1, 1,
           col.data_ptr())
        torch.cuda.synchronize()
        exp = O.dcn_v2_im2col(x, off, m, k, k, s, p, 1, 1)
        # same fp32 expressions; nvcc may contract a*b+c into FMA, so allow an ulp-level difference
        assert np.abs(col.cpu().numpy() - exp).max() < 2e-6 * max(np.abs(exp).max(), 1.0)


@pytest.mark.parametrize("shape", [(2, 16, 9, 11, 8, 3, 1, 1), (1, 64, 12, 20, 32, 3, 1, 1), (1, 6, 7, 9, 5, 1, 1, 0),
                                   (1, 32, 10, 13, 20, 3, 2, 1)])
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16x3", 6e-5), ("bf16", 1.5e-2)])
def test_dcn_v2_backward_vs_oracle(shape, precision, tol):
    """DCNv2Function under autograd vs the C oracle's restatement of dcn_v2_cuda_backward (fp32 inputs,
    the oracle accumulates in double): all five gradients."""
    from m3dssd_b200.model.DCNv2.dcn_v2_func import DCNv2Function
    B, Cin, H, W, Cout, k, s, p = shape
    rng = np.random.default_rng(sum(shape) + 1)
    x = rng.standard_normal((B, Cin, H, W)).astype(np.float32)
    Ho, Wo = O.dcn_out_shape(H, W, k, k, s, p, 1)
    off = (rng.standard_normal((B, 2 * k * k, Ho, Wo)) * 2.5).astype(np.float32)
    m = rng.random((B, k * k, Ho, Wo)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    gy = rng.standard_normal((B, Cout, Ho, Wo)).astype(np.float32)
    ref = O.dcn_v2_backward(x, off, m, w, gy, s, p, 1, 1)
    ts = [torch.from_numpy(a).cuda().requires_grad_(True) for a in (x, off, m, w, b)]
    out = DCNv2Function(s, p, 1, 1, precision=precision)(*ts)  # bf16x3: forward and the W^T dY GEMM on the tensor cores
    out.backward(torch.from_numpy(gy).cuda())
    for name, t, r in zip(("input", "offset", "mask", "weight", "bias"), ts, ref):
        err = np.abs(t.grad.cpu().numpy() - r).max() / max(np.abs(r).max(), 1e-6)
        assert err < tol, (name, err)


@pytest.mark.parametrize("shape", [(2, 64, 12, 16, 64, 3, 1, 1, 2), (1, 12, 7, 9, 5, 3, 1, 1, 3), (1, 8, 6, 7, 4, 1, 1, 0, 4)])
@pytest.mark.parametrize("precision,tol", [("fp32", 3e-6), ("bf16x3", 3e-5)])
def test_dcn_v2_deformable_groups_forward_backward_vs_oracle(shape, precision, tol):
    """deformable_groups > 1 (the reference's example_dconv uses dg = 2, model/DCNv2/test.py:169-179; index math
    dcn_v2_im2col_cuda.cu:139-149): forward and all five gradients vs the C oracle."""
    from m3dssd_b200.model.DCNv2.dcn_v2_func import DCNv2Function
    B, Cin, H, W, Cout, k, s, p, dg = shape
    rng = np.random.default_rng(sum(shape) + 7)
    x = rng.standard_normal((B, Cin, H, W)).astype(np.float32)
    Ho, Wo = O.dcn_out_shape(H, W, k, k, s, p, 1)
    off = (rng.standard_normal((B, 2 * dg * k * k, Ho, Wo)) * 2.5).astype(np.float32)
    m = rng.random((B, dg * k * k, Ho, Wo)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    gy = rng.standard_normal((B, Cout, Ho, Wo)).astype(np.float32)
    ref = O.dcn_v2_forward(x, off, m, w, b, s, p, 1, dg)
    ts = [torch.from_numpy(a).cuda().requires_grad_(True) for a in (x, off, m, w, b)]
    out = DCNv2Function(s, p, 1, dg, precision=precision)(*ts)
    err = np.abs(out.detach().cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1.0)
    assert out.shape == ref.shape and err < tol, err
    if precision == "fp32":
        out.backward(torch.from_numpy(gy).cuda())
        gref = O.dcn_v2_backward(x, off, m, w, gy, s, p, 1, dg)
        for name, t, r in zip(("input", "offset", "mask", "weight", "bias"), ts, gref):
            e = np.abs(t.grad.cpu().numpy() - r).max() / max(np.abs(r).max(), 1e-6)
            assert e < 2e-5, (name, e)


def test_dcn_module_with_two_deformable_groups():
    """DCN(64, 64, 3, 1, 1, deformable_groups=2) as in the reference's example_dconv: conv_offset_mask has
    dg * 27 channels; the module output matches the oracle on the module's own offsets / masks."""
    from m3dssd_b200.model.DCNv2.dcn_v2 import DCN
    torch.manual_seed(1)
    dcn = DCN(64, 64, 3, 1, 1, deformable_groups=2).cuda()
    assert dcn.conv_offset_mask.out_channels == 2 * 27
    dcn.conv_offset_mask.weight.data.normal_(0, 0.05)
    dcn.conv_offset_mask.bias.data.normal_(0, 0.2)
    x = torch.randn(2, 64, 16, 16).cuda()
    with torch.no_grad():
        y = dcn(x)
        om = dcn.conv_offset_mask(x)
    off, mask = om[:, :36].cpu().numpy(), torch.sigmoid(om[:, 36:]).cpu().numpy()
    ref = O.dcn_v2_forward(x.cpu().numpy(), off, mask, dcn.weight.detach().cpu().numpy(),
                           dcn.bias.detach().cpu().numpy(), 1, 1, 1, 2)
    assert np.abs(y.cpu().numpy() - ref).max() < 3e-6 * max(np.abs(ref).max(), 1)


def test_dcn_v2_backward_bf16_mode_is_deterministic():
    """M3D_BF16 backward (the training mode): fixed-point col2im accumulation, ordered split-K wgrad and channel sums ->
    two runs give bit-identical gradients (the reference's float atomicAdd col2im does not)."""
    from m3dssd_b200 import ops
    from m3dssd_b200._lib import M3D_BF16
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 64, 24, 40, generator=g).cuda()
    off = (torch.randn(2, 18, 24, 40, generator=g) * 2.5).cuda()
    m = torch.rand(2, 9, 24, 40, generator=g).cuda()
    w = (torch.randn(64, 64, 3, 3, generator=g) / 24).cuda()
    gy = torch.randn(2, 64, 24, 40, generator=g).cuda()
    a = ops.dcn_v2_backward(x, off, m, w, gy, 1, 1, 1, 1, precision=M3D_BF16)
    b = ops.dcn_v2_backward(x, off, m, w, gy, 1, 1, 1, 1, precision=M3D_BF16)
    for name, ta, tb in zip(("input", "offset", "mask", "weight", "bias"), a, b):
        assert torch.equal(ta, tb), name
