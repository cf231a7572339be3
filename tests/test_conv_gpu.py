"""GPU parity of the tcgen05 implicit-GEMM convolution / fused DCNv2 kernels (through the C ABI
m3d_conv2d_nhwc) against torch CPU conv2d and torchvision deform_conv2d, which the C oracle is
pinned to (tests/test_oracle.py).  Tolerances are set in conv_cases.run_conv_case:
fp32 mode (bf16x3 split) 2e-5 of the output scale; bf16 mode = one bf16 ulp of the output scale."""
import pytest

from conv_cases import CASES, run_conv_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,kw", CASES, ids=[c[0] for c in CASES])
def test_conv_case(name, kw):
    err, scale, tol = run_conv_case(name, **kw)
    assert err <= tol, "%s: max abs err %.3e > tol %.3e (output scale %.3f)" % (name, err, tol, scale)


@pytest.mark.parametrize("hw", [(12, 40), (96, 320)])
def test_k16_zero_hint_skips_dead_k_steps_bit_identically(hw):
    """m3d_conv_desc.k16_zero: the 2x2 space-to-depth rewrite of a 16 -> 16 3x3 conv (level0: 64 -> 64 dense with 20 of
    its 36 16-channel k-steps structurally zero) through the resident-weight halo kernel with the hint == without it,
    bit for bit (a skipped step only ever added 0), and == the original conv on the un-packed tensor."""
    import torch
    import torch.nn.functional as F
    from m3dssd_b200 import ops
    H, W = hw
    g = torch.Generator().manual_seed(H)
    x = torch.randn(2, 16, 2 * H, 2 * W, generator=g).bfloat16().float()
    w = (torch.randn(16, 16, 3, 3, generator=g) / 12.0).bfloat16().float()
    b = torch.randn(16, generator=g)
    wp, bp = ops.s2d_conv3x3_weight(w, b)
    hi, _ = ops.pack_conv_weight(wp, in_splits=[(64, 64)], mode="bf16")
    kz = ops.k16_zero_mask(hi)
    assert bin(kz[0]).count("1") == 20 and kz[1] == 0
    # [N, C, 2Y+dy, 2X+dx] -> [N, Y, X, (dy, dx, c)]
    xs = x.reshape(2, 16, H, 2, W, 2).permute(0, 2, 4, 3, 5, 1).reshape(2, H, W, 64).bfloat16().contiguous().cuda()
    outs = []
    for hint in ((0, 0), kz):
        out = torch.full((2, H, W, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
        ops.conv2d_nhwc([(xs, 0, 64)], hi.cuda(), out, R=3, S=3, stride=1, pad=1, Cout=64, bias=bp.cuda(), slope=0.01,
                        k16_zero=hint)
        assert ops.last_kernel().startswith("conv_halo_kernel<64,bf16,1,1>"), ops.last_kernel()
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.01)
    got = outs[1].float().cpu().reshape(2, H, W, 2, 2, 16).permute(0, 5, 1, 3, 2, 4).reshape(2, 16, 2 * H, 2 * W)
    assert (got - ref).abs().max().item() <= 2 ** -8 * max(1.0, ref.abs().max().item())
    # a hint claiming every step dead is ignored (nothing would initialise the accumulator)
    out = torch.full((2, H, W, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.conv2d_nhwc([(xs, 0, 64)], hi.cuda(), out, R=3, S=3, stride=1, pad=1, Cout=64, bias=bp.cuda(), slope=0.01,
                    k16_zero=((1 << 36) - 1, 0))
    assert torch.equal(out, outs[0])
