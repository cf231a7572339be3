"""Time the fused DCNv2 layers in isolation (graph replays) for the ring/L1 variants (development aid)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from m3dssd_b200 import ops

SHAPES = [("node 128->128 @48x160", 8, 48, 160, 128, 128, 3), ("proj 256->128 @24x80", 8, 24, 80, 256, 128, 3),
          ("node 256->256 @24x80", 8, 24, 80, 256, 256, 3), ("proj 512->256 @12x40", 8, 12, 40, 512, 256, 3),
          ("center 128->128 1x1 @48x160", 8, 48, 160, 128, 128, 1)]


def run_case(N, H, W, Cin, Cout, R, legacy, iters=20):
    if legacy:
        os.environ["M3D_DCN_LEGACY"] = "1"
    else:
        os.environ.pop("M3D_DCN_LEGACY", None)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, R, R, device="cuda", generator=g) / (Cin * R * R) ** 0.5
    wp, _ = ops.pack_conv_weight(w)
    b = torch.zeros(Cout, device="cuda")
    om = torch.zeros(N, H, W, 32, device="cuda")
    om[..., :2 * R * R] = torch.randn(N, H, W, 2 * R * R, device="cuda", generator=g) * 2.0
    om[..., 2 * R * R:3 * R * R] = torch.randn(N, H, W, R * R, device="cuda", generator=g)
    out = torch.empty(N, H, W, Cout, device="cuda", dtype=torch.bfloat16)

    def run():
        ops.conv2d_nhwc([x], wp, out, R=R, S=R, stride=1, pad=R // 2, Cout=Cout, bias=b, slope=0.01, om=om, sigmoid_mask=True)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ref = out.clone()
    g1 = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        run()
        with torch.cuda.graph(g1, stream=s):
            for _ in range(iters):
                run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g1.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / iters)
    return min(ts), ref


for name, N, H, W, Cin, Cout, R in SHAPES:
    base = None
    for legacy in (1, 0, -1):
        os.environ["M3D_DCN_HALF"] = "1" if legacy < 0 else "0"
        t, out = run_case(N, H, W, Cin, Cout, R, legacy > 0)
        if base is None:
            base = out
        same = bool(torch.equal(base, out))
        print("%-30s legacy=%d  %7.1f us  bit-identical to the legacy kernel: %s" % (name, legacy, t, same), flush=True)
