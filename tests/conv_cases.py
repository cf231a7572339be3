"""Shared case runner for the tcgen05 conv / DCNv2 kernels (used by the GPU tests
and by tools/gpu_probe.py, which prints every case instead of stopping at the
first failure)."""
import torch
import torch.nn.functional as F


def _nhwc(x, dtype):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def run_conv_case(name, N, H, W, cins, cout, k=3, stride=1, pad=None, dtype=torch.bfloat16, out_dtype=None,
                  bias=True, res=False, slope=0.01, deform=False, sigmoid_mask=False, force_gather=False,
                  groups=1, seed=0, offset_sigma=2.0, mode=None):
    """Returns (max_abs_err, ref_scale, tolerance) comparing the CUDA kernel with a CPU reference."""
    import torchvision
    from m3dssd_b200 import ops

    g = torch.Generator().manual_seed(seed)
    pad = k // 2 if pad is None else pad
    fp32 = dtype == torch.float32
    out_dtype = out_dtype or dtype
    cin = sum(cins)
    P = (H + 2 * pad - k) // stride + 1
    Q = (W + 2 * pad - k) // stride + 1

    xs = [torch.randn(N, c * groups, H, W, generator=g) for c in cins]
    w = torch.randn(groups * cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(groups * cout, generator=g) if bias else None
    r = torch.randn(N, groups * cout, P, Q, generator=g) if res else None
    if not fp32:  # the kernel sees bf16-rounded operands; give the reference the same values
        xs = [x.bfloat16().float() for x in xs]
        w = w.bfloat16().float()
        if r is not None:
            r = r.bfloat16().float()

    om = None
    if deform:
        kk = k * k
        off = torch.randn(N, 2 * kk, P, Q, generator=g) * offset_sigma
        mlogit = torch.randn(N, kk, P, Q, generator=g)
        mask = torch.sigmoid(mlogit) if sigmoid_mask else torch.rand(N, kk, P, Q, generator=g)
        om = torch.cat([off, mlogit if sigmoid_mask else mask], dim=1)

    # ---- CPU reference (fp32)
    refs = []
    for gi in range(groups):
        xg = torch.cat([x[:, gi * c:(gi + 1) * c] for x, c in zip(xs, cins)], dim=1)
        wg = w[gi * cout:(gi + 1) * cout]
        bg = b[gi * cout:(gi + 1) * cout] if b is not None else None
        if deform:
            y = torchvision.ops.deform_conv2d(xg, off, wg, bg, stride=stride, padding=pad, mask=mask)
        else:
            y = F.conv2d(xg, wg, bg, stride=stride, padding=pad)
        refs.append(y)
    ref = torch.cat(refs, dim=1)
    if r is not None:
        ref = ref + r
    ref = F.leaky_relu(ref, slope) if slope != 1.0 else ref

    # ---- CUDA kernel
    dev = "cuda"
    ins = [_nhwc(x, dtype).to(dev) for x in xs]
    whi_l, wlo_l = [], []
    for gi in range(groups):
        hi, lo = ops.pack_conv_weight(w[gi * cout:(gi + 1) * cout], in_splits=list(cins),
                                      mode=mode or ("bf16x3" if fp32 else "bf16"))
        whi_l.append(hi)
        wlo_l.append(lo)
    whi = torch.cat(whi_l).to(dev) if whi_l[0] is not None else None
    if mode == "fp32":
        wlo = torch.cat(wlo_l).to(dev)
    else:
        wlo = tuple(torch.cat([w_[i] for w_ in wlo_l]).to(dev) for i in range(2)) if fp32 else None
    out = torch.full((N, P, Q, groups * cout), float("nan"), dtype=out_dtype, device=dev)
    inputs = [(t, 0, c) for t, c in zip(ins, cins)]
    ops.conv2d_nhwc(
        inputs, whi, out, R=k, S=k, stride=stride, pad=pad, Cout=cout,
        bias=b.to(dev) if b is not None else None,
        res=_nhwc(r, dtype).to(dev) if r is not None else None,
        slope=slope, weight_lo=wlo,
        om=_nhwc(om, torch.float32).to(dev) if om is not None else None,
        sigmoid_mask=sigmoid_mask, groups=groups,
        in_goff=[c for c in cins] if groups > 1 else None,
        weight_goff=cout, bias_goff=cout, out_goff=cout, res_goff=cout,
        force_gather=force_gather)
    torch.cuda.synchronize()
    got = out.float().cpu().permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    if fp32:
        # "fp32": IEEE FMA accumulation, two fp32 summation orders differ by ~sqrt(K) ulp;
        # "bf16x3": tensor-core accumulators truncate, error grows with the number of accumulated terms
        tol = (3e-6 if mode == "fp32" else 2e-5) * max(scale, 1.0)
    else:
        # fp32 accumulation of exact bf16 products; bf16 output rounding 2^-9;
        # the deformable blend rounds the sampled value to bf16 (2^-9 of |x|)
        tol = (2 ** -8 if out_dtype == torch.bfloat16 else 1e-4) * max(scale, 1.0)
        if deform:
            tol += 2 ** -8 * 4.0
    return err, scale, tol


CASES = [
    # name, kwargs
    ("gemm_1x1_c64", dict(N=1, H=16, W=16, cins=[64], cout=128, k=1, bias=False, slope=1.0)),
    ("gemm_1x1_c128_f32out", dict(N=1, H=16, W=16, cins=[128], cout=128, k=1, out_dtype=torch.float32, slope=1.0)),
    ("conv3x3_c64", dict(N=2, H=12, W=40, cins=[64], cout=64)),
    ("conv3x3_c128_res", dict(N=2, H=24, W=80, cins=[128], cout=128, res=True)),
    ("conv3x3_s2_c64", dict(N=2, H=24, W=80, cins=[64], cout=128, stride=2)),
    ("conv3x3_c16_bk16", dict(N=1, H=32, W=64, cins=[16], cout=16)),
    ("conv3x3_s2_c16_bk16", dict(N=1, H=32, W=64, cins=[16], cout=32, stride=2)),
    ("conv3x3_s2_c32_bk32", dict(N=1, H=32, W=64, cins=[32], cout=64, stride=2)),
    ("conv1x1_c32_bk32", dict(N=1, H=16, W=32, cins=[32], cout=64, k=1)),
    ("root_concat_4", dict(N=1, H=12, W=40, cins=[128, 128, 64, 128], cout=128, k=1)),
    ("conv1x1_cout256", dict(N=4, H=24, W=80, cins=[256], cout=256, k=1)),
    ("conv3x3_cout256_big", dict(N=8, H=24, W=80, cins=[256], cout=256)),
    ("conv3x3_cout256_big_res", dict(N=8, H=24, W=80, cins=[256], cout=256, res=True)),  # level4 conv2: pair kernel, BN = 256
    ("head_cout36_f32", dict(N=1, H=12, W=40, cins=[256], cout=36, k=1, out_dtype=torch.float32, slope=1.0)),
    ("logits_cout144_f32", dict(N=2, H=24, W=80, cins=[256], cout=144, k=1, out_dtype=torch.float32, slope=1.0)),
    ("logits_cout160_f32_ragged", dict(N=3, H=13, W=37, cins=[128], cout=160, k=1, out_dtype=torch.float32)),
    ("heads_grouped", dict(N=1, H=12, W=40, cins=[256], cout=256, k=1, groups=3)),
    ("heads_grouped_36", dict(N=1, H=12, W=40, cins=[256], cout=36, k=1, groups=4, out_dtype=torch.float32, slope=1.0)),
    ("offsetconv_27_f32", dict(N=1, H=24, W=80, cins=[128], cout=27, out_dtype=torch.float32, slope=1.0)),
    ("offsetconv_27_c64_f32", dict(N=2, H=13, W=37, cins=[64], cout=27, out_dtype=torch.float32, slope=1.0)),
    ("offsetconv_27_c256_f32", dict(N=1, H=12, W=40, cins=[256], cout=27, out_dtype=torch.float32, slope=1.0)),
    ("conv3x3_c64_ragged_res", dict(N=3, H=13, W=37, cins=[64], cout=64, res=True)),
    ("conv7x1ish_5x5", dict(N=1, H=16, W=16, cins=[64], cout=64, k=5)),
    ("gather_plain_bf16", dict(N=1, H=12, W=40, cins=[64], cout=64, force_gather=True)),
    ("gather_plain_s2_bf16", dict(N=1, H=24, W=80, cins=[64], cout=128, stride=2, force_gather=True)),
    ("dcn_bf16_c64", dict(N=1, H=12, W=40, cins=[64], cout=64, deform=True)),
    ("dcn_bf16_c128_sig", dict(N=2, H=24, W=80, cins=[128], cout=128, deform=True, sigmoid_mask=True)),
    ("dcn_bf16_c512_256", dict(N=1, H=12, W=40, cins=[512], cout=256, deform=True, sigmoid_mask=True)),
    ("dcn_bf16_1x1", dict(N=1, H=12, W=40, cins=[128], cout=128, k=1, deform=True, res=False, slope=1.0)),
    ("dcn_bf16_bigoff", dict(N=1, H=12, W=40, cins=[64], cout=64, deform=True, offset_sigma=20.0)),
    # fused DCN kernel, staged-window mode (dcn_fused.cu, BN = 128): node / proj / shape_align shapes, far offsets that
    # leave the window (global fallback, entry by entry), image sizes that are not multiples of the tile
    ("dcn_bf16_c128_node_res", dict(N=2, H=48, W=160, cins=[128], cout=128, deform=True, sigmoid_mask=True, res=True)),
    ("dcn_bf16_c256_128_proj", dict(N=2, H=24, W=80, cins=[256], cout=128, deform=True, sigmoid_mask=True)),
    ("dcn_bf16_c128_faroff", dict(N=1, H=24, W=80, cins=[128], cout=128, deform=True, offset_sigma=12.0)),
    ("dcn_bf16_c128_ragged", dict(N=3, H=13, W=37, cins=[128], cout=128, deform=True, sigmoid_mask=True, offset_sigma=3.0)),
    ("f32_plain_c64", dict(N=1, H=12, W=40, cins=[64], cout=64, dtype=torch.float32)),
    ("f32_plain_s2_res", dict(N=1, H=24, W=80, cins=[64], cout=128, stride=2, dtype=torch.float32)),
    ("f32_concat", dict(N=1, H=12, W=40, cins=[64, 128], cout=64, k=1, dtype=torch.float32)),
    ("f32_dcn_c128", dict(N=1, H=24, W=80, cins=[128], cout=128, deform=True, sigmoid_mask=True, dtype=torch.float32)),
    ("f32_dcn_c256_256", dict(N=1, H=12, W=40, cins=[256], cout=256, deform=True, dtype=torch.float32, res=True)),
    ("f32_7x7", dict(N=1, H=16, W=24, cins=[64], cout=16, k=7, dtype=torch.float32)),
    ("ref32_plain_c64", dict(N=1, H=12, W=40, cins=[64], cout=64, dtype=torch.float32, mode="fp32")),
    ("ref32_plain_s2_res", dict(N=2, H=24, W=80, cins=[64], cout=128, stride=2, dtype=torch.float32, res=True, mode="fp32")),
    ("ref32_concat_c16", dict(N=1, H=12, W=40, cins=[16, 64, 128], cout=36, k=1, dtype=torch.float32, mode="fp32", slope=1.0)),
    ("ref32_dcn_c128", dict(N=1, H=24, W=80, cins=[128], cout=128, deform=True, sigmoid_mask=True, dtype=torch.float32, mode="fp32")),
    ("ref32_dcn_1x1_res", dict(N=1, H=12, W=40, cins=[128], cout=128, k=1, deform=True, dtype=torch.float32, res=True, mode="fp32", slope=1.0)),
    ("ref32_dcn_c512_256", dict(N=1, H=12, W=40, cins=[512], cout=256, deform=True, dtype=torch.float32, mode="fp32")),
    ("ref32_7x7", dict(N=1, H=16, W=24, cins=[64], cout=16, k=7, dtype=torch.float32, mode="fp32")),
]
