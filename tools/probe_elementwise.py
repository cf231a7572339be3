"""Graph-replayed timing of the elementwise / layout kernels at the benchmark shapes (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from m3dssd_b200 import ops

def timeit(name, fns, bytes_):
    for f in fns:
        f()
    torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for f in fns:
                f()
    torch.cuda.synchronize(); g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / (5 * len(fns))
    print("%-34s %7.2f us  %7.1f GB/s" % (name, us, bytes_ / us * 1e-3))

S = 8
bf = dict(device="cuda", dtype=torch.bfloat16)
# upsample 24x80 -> 48x160, 128 ch + skip
xs = [torch.randn(8, 24, 80, 128, **bf) for _ in range(S)]
sk = [torch.randn(8, 48, 160, 128, **bf) for _ in range(S)]
out = [torch.empty(8, 48, 160, 128, **bf) for _ in range(S)]
w = ops.pack_upsample_weight(torch.rand(128, 1, 4, 4)).cuda()
timeit("upsample_add 128ch 24x80->48x160", [lambda i=i: ops.upsample_add(xs[i], w, sk[i], out[i], 2) for i in range(S)],
       (xs[0].numel() + 2 * sk[0].numel()) * 2)
# maxpool 96x320x64 -> 48x160
xi = [torch.randn(8, 96, 320, 64, **bf) for _ in range(S)]
xo = [torch.empty(8, 48, 160, 64, **bf) for _ in range(S)]
timeit("maxpool2x2 64ch 96x320", [lambda i=i: ops.maxpool2x2(xi[i], xo[i]) for i in range(S)], (xi[0].numel() + xo[0].numel()) * 2)
# softmax (detect variant)
A, K, H, W, B = 36, 4, 48, 160, 8
M = A * H * W
lg = [torch.randn(B, H, W, K * A, device="cuda") for _ in range(S)]
f32 = dict(dtype=torch.float32, device="cuda")
fm, fa = torch.zeros(B, H, W, **f32), torch.zeros(B, H, W, dtype=torch.int32, device="cuda")
sc, cp = torch.zeros(B, M, **f32), torch.zeros(B, M, dtype=torch.uint8, device="cuda")
co, po = torch.zeros(B, M, K, **f32), torch.zeros(B, M, K, **f32)
timeit("cls_softmax full", [lambda i=i: ops.cls_softmax(lg[i], A, K, co, po, fm, fa, sc, cp) for i in range(S)], lg[0].numel() * 4 * 3)
timeit("cls_softmax detect (no copies)", [lambda i=i: ops.cls_softmax(lg[i], A, K, None, None, fm, fa, sc, cp) for i in range(S)], lg[0].numel() * 4 * 1.3)
