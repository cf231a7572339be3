"""Phase timeline of the fused head kernel (development aid)."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from m3dssd_b200 import _lib
# probe build: tools/build_probe.sh
_lib.LIB_PATH = os.path.join(ROOT, "build", "libm3d_probe.so")
from m3dssd_b200 import ops

os.environ["M3D_HEAD_DBG"] = "1"
N, H, W, Cx, G, A, rows3 = 8, 48, 160, 128, 4, 36, 48
g = torch.Generator().manual_seed(0)
x = torch.randn(N, H, W, Cx, generator=g).bfloat16().cuda()
w1 = (torch.randn(G * 256, Cx, generator=g) / 11).bfloat16().cuda()
w2 = (torch.randn(G * 256, 256, generator=g) / 16).bfloat16().cuda()
w3 = (torch.randn(G * rows3, 256, generator=g) / 16).bfloat16().cuda()
b1 = torch.randn(G * 256, generator=g).cuda(); b2 = torch.randn(G * 256, generator=g).cuda(); b3 = torch.randn(G * rows3, generator=g).cuda()
out = torch.zeros(N, H, W, 11 * A, device="cuda")
for _ in range(3):
    ops.head_mlp(x, 0, Cx, w1, b1, w2, b2, w3, b3, G, A, rows3, out, 0, 0.01)
torch.cuda.synchronize()
buf = np.zeros(1024, dtype=np.int64)
L = _lib.lib()
L.m3d_head_debug_read.argtypes = [C.c_void_p, C.c_int]
L.m3d_head_debug_read(buf.ctypes.data, buf.size)
a = buf.reshape(32, 32)
t0 = a[0, 8]
names = {0: "G2 start", 1: "G2 issued", 2: "G3 start", 3: "G3 issued", 8: "E1 wait", 9: "E1 go", 10: "E2 wait", 11: "E2 go", 12: "E3 wait", 13: "E3 go", 14: "E3 done"}
for kb in range(4):
    names[16 + kb] = "w2[%d] in" % kb       # MMA warp: weight slot of G2 slab kb has landed
    names[20 + kb] = "y1[%d] rdy" % kb      # MMA warp: Y slab kb written by E1 -> MMAs issued right after
for i, n in enumerate(["W2[0]", "W2[1]", "W2[2]", "W2[3]", "W1'[0]", "W1'[1]", "W3"]):
    names[24 + i] = "slot for " + n         # producer: ring slot freed (MMAs that read it have completed)
for li in range(6):
    ev = sorted((a[li, k] - t0, names[k]) for k in names if a[li, k])
    print("item %d: " % li + "  ".join("%s@%d" % (n, t) for t, n in ev))
