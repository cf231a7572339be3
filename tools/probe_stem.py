"""Phase timeline of the stem (7x7, space-to-depth) gather kernel (development aid; needs the -DM3D_PROBE build,
see tools/probe_heads.py)."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from m3dssd_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "build", "libm3d_probe.so")
from m3dssd_b200 import ops

N, H, W = 8, 384, 1280
g = torch.Generator().manual_seed(0)
img = torch.randn(N, 3, H, W, generator=g).cuda()
w = torch.randn(16, 3, 7, 7, generator=g) / 12
b = torch.randn(16, generator=g)
wp, bp = ops.pack_stem_s2d(w, b)
wp, bp = wp.cuda(), bp.cuda()
out = torch.zeros(N, H // 2, W // 2, 64, dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    ops.stem_conv7x7_s2d(img, wp, bp, out, 0.01)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.stem_conv7x7_s2d(img, wp, bp, out, 0.01)
e1.record()
torch.cuda.synchronize()
print("stem: %.1f us / launch" % (e0.elapsed_time(e1) * 100))
buf = np.zeros(6 * 32, dtype=np.int64)
L = _lib.lib()
L.m3d_gather_debug_read.argtypes = [C.c_void_p, C.c_int]
L.m3d_gather_debug_read(buf.ctypes.data, buf.size)
a = buf.reshape(6, 32)
t0 = a[0, 0]
names = {0: "P bar", 1: "P img", 2: "P kb0", 3: "P kb1", 4: "P kb2", 8: "M acc free", 9: "M kb0", 10: "M kb1", 11: "M kb2",
         12: "M done", 16: "E start", 17: "E end"}
for li in range(6):
    ev = sorted((a[li, k] - t0, names[k]) for k in names if a[li, k])
    print("tile %d: " % li + "  ".join("%s@%d" % (n, t) for t, n in ev))
