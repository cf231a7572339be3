"""Per-shape timing + parity of the plain 3x3 / 1x1 conv kernels (development aid; A/B builds via M3D_LIB=<path>).

Each shape runs over a rotation of 8 input / residual / output sets (> L2), 40 back-to-back launches timed with CUDA
events; the result of the first set is checked against torch's fp32 convolution of the same bf16 operands."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from m3dssd_b200 import ops

# (name, N, H, W, Cin, Cout, R, stride, residual)
SHAPES = [
    ("level3 3x3 128->128 +res", 8, 48, 160, 128, 128, 3, 1, True),
    ("level3 3x3 128->128", 8, 48, 160, 128, 128, 3, 1, False),
    ("level4 3x3 256->256 +res", 8, 24, 80, 256, 256, 3, 1, True),
    ("level5 3x3 512->512 +res", 8, 12, 40, 512, 512, 3, 1, True),
    ("level5 3x3 512->512", 8, 12, 40, 512, 512, 3, 1, False),
    ("level2 3x3 64->64 +res", 8, 96, 320, 64, 64, 3, 1, True),
    ("level0 3x3 64->64 s2d", 8, 192, 640, 64, 64, 3, 1, False),
    ("offset 3x3 128->27 f32", 8, 48, 160, 128, 27, 3, 1, False),
    ("level3 3x3 s2 64->128", 8, 96, 320, 64, 128, 3, 2, False),
    ("level4 3x3 s2 128->256", 8, 48, 160, 128, 256, 3, 2, False),
    ("root 1x1 256->128", 8, 48, 160, 256, 128, 1, 1, False),
    ("cls.l1 3x3 64->256", 8, 48, 160, 64, 256, 3, 1, False),
    ("root 1x1 512->256 +res", 8, 24, 80, 512, 256, 1, 1, True),
    ("root2 1x1 128->64", 8, 96, 320, 128, 64, 1, 1, False),
    ("proj 1x1 64->128", 8, 48, 160, 64, 128, 1, 1, False),
]
only = sys.argv[1:]
SETS = 8
g = torch.Generator(device="cuda").manual_seed(0)
print("lib:", os.environ.get("M3D_LIB", "default"))
for name, N, H, W, Cin, Cout, R, stride, has_res in SHAPES:
    if only and not any(o in name for o in only):
        continue
    pad = R // 2
    P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    xs = [torch.randn(N, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16) for _ in range(SETS)]
    rs = [torch.randn(N, P, Q, Cout, device="cuda", generator=g).to(torch.bfloat16) for _ in range(SETS)] if has_res else None
    f32out = "f32" in name
    outs = [torch.empty(N, P, Q, 32 if f32out else Cout, device="cuda", dtype=torch.float32 if f32out else torch.bfloat16)
            for _ in range(SETS)]
    w = torch.randn(Cout, Cin, R, R, device="cuda", generator=g) / (Cin * R * R) ** 0.5
    wp, _ = ops.pack_conv_weight(w.cpu())
    if f32out:  # weight rows padded to the N tile
        wpad = torch.zeros(32, wp.shape[1], dtype=wp.dtype)
        wpad[:Cout] = wp
        wp = wpad
    wp = wp.cuda()
    b = torch.randn(Cout, device="cuda", generator=g)

    def run(i):
        ops.conv2d_nhwc([xs[i]], wp, outs[i], R=R, S=R, stride=stride, pad=pad, Cout=Cout, bias=b, slope=1.0 if f32out else 0.01,
                        res=rs[i] if has_res else None)
    for i in range(SETS):
        run(i)
    torch.cuda.synchronize()
    kern = ops.last_kernel()
    ref = F.conv2d(xs[0].float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), b, stride=stride, padding=pad)
    if has_res:
        ref = ref + rs[0].float().permute(0, 3, 1, 2)
    ref = (ref if f32out else F.leaky_relu(ref, 0.01)).permute(0, 2, 3, 1)
    err = (outs[0][..., :Cout].float() - ref).abs().max().item() / ref.abs().max().item()
    # replay through a CUDA graph: the eager ctypes call costs ~20 us of host time, more than most of these kernels
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(SETS):
                run(i)
    torch.cuda.synchronize()
    graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / (reps * SETS)
    fl = 2.0 * N * P * Q * Cout * Cin * R * R
    print("%-28s %-40s %7.2f us  %7.1f TF/s  relerr %.2e %s" % (name, kern, us, fl / us * 1e-6, err,
                                                               "OK" if err < 1.2e-2 else "MISMATCH"))
