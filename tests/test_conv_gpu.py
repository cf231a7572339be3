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
