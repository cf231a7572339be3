"""Pipeline timeline of the CTA-pair 3x3 conv kernel on a level-3 layer (128 -> 128 @ 48x160, batch 8); development aid,
needs the probe build: M3D_VARIANT=probe M3D_NVCC_EXTRA=-DM3D_PROBE python -m m3dssd_b200.build"""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from m3dssd_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "m3dssd_b200", "libm3dssd_b200.probe.so")
from m3dssd_b200 import ops

N, H, W, Cin, Cout = 8, 48, 160, 128, 128
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(N, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16)
w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (Cin * 9) ** 0.5
wp, _ = ops.pack_conv_weight(w.cpu())
wp = wp.cuda()
b = torch.zeros(Cout, device="cuda")
res = torch.randn(N, H, W, Cout, device="cuda", generator=g).to(torch.bfloat16)
out = torch.empty(N, H, W, Cout, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.conv2d_nhwc([x], wp, out, R=3, S=3, stride=1, pad=1, Cout=Cout, bias=b, slope=0.01, res=res)
torch.cuda.synchronize()
print("kernel:", ops.last_kernel())
buf = np.zeros(64 * 4, dtype=np.int64)
L = _lib.lib()
L.m3d_halo2_debug_read.argtypes = [C.c_void_p, C.c_int]
L.m3d_halo2_debug_read(buf.ctypes.data, buf.size)
a = buf.reshape(64, 4)
t0 = a[0, 0]
print("stage: wait-start  full-seen  issued   (clk since first stage; leader MMA warp)")
for i in range(30):
    r = a[i]
    if r[0]:
        print("%2d: %8d %8d %8d   wait %5d  issue %4d" % (i, r[0] - t0, r[1] - t0, r[2] - t0, r[1] - r[0], r[2] - r[1]))
print("item: tempty-wait-start  tempty-free | epilogue(warp 2): tfull-wait-start  tfull-seen")
for i in range(6):
    r = a[48 + i]
    if r[0]:
        print("%2d: %8d %8d | %8d %8d" % (i, r[0] - t0, r[1] - t0, r[2] - t0, r[3] - t0))
r = a[47]
print("kernel entry %d  setup done %d  grid-dep passed %d  all roles done %d (clk since first stage)" % tuple(int(v - t0) for v in r))
eb = np.zeros(64, dtype=np.int64)
L.m3d_halo2_epi_read.argtypes = [C.c_void_p]
L.m3d_halo2_epi_read(eb.ctypes.data)
print("epilogue thread 0, per tile: tmem read | ready | math done | fence | barrier | store read | tile end")
for i in range(7):
    r = eb[i * 8:i * 8 + 7]
    if r[0]:
        print("%2d: " % i + " ".join("%8d" % (v - t0 if v else -1) for v in r))
