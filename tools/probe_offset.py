"""Time the DCN offset conv (3x3, Cout = 27, fp32 out) with the halo kernel's probe bits (development aid)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from m3dssd_b200 import ops

def run(N, H, W, Cin, dbg, iters=20):
    os.environ["M3D_DBG"] = str(dbg)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(27, Cin, 3, 3, device="cuda", generator=g) / (Cin * 9) ** 0.5
    wp, _ = ops.pack_conv_weight(w)
    wpad = torch.zeros(32, wp.shape[1], dtype=wp.dtype, device="cuda")
    wpad[:27] = wp
    b = torch.zeros(27, device="cuda")
    out = torch.zeros(N, H, W, 32, device="cuda")
    f = lambda: ops.conv2d_nhwc([x], wpad, out, R=3, S=3, stride=1, pad=1, Cout=27, bias=b, slope=1.0)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    g1 = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        f()
        with torch.cuda.graph(g1, stream=s):
            for _ in range(iters):
                f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g1.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / iters)
    return min(ts)

for name, shp in (("node 128ch @48x160", (8, 48, 160, 128)), ("proj 256ch @24x80", (8, 24, 80, 256)), ("proj 512ch @12x40", (8, 12, 40, 512))):
    print(name, "  ".join("dbg=%d: %.1f us" % (d, run(*shp, d)) for d in (0, 4, 8, 1, 13)), flush=True)
