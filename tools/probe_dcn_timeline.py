"""Per-k-block timeline of the fused DCN kernel on a node layer (128 -> 128 @ 48x160, batch 8); development aid,
needs the -DM3D_PROBE build (build/libm3d_probe.so, see tools/probe_heads.py)."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from m3dssd_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "m3dssd_b200", "libm3dssd_b200.probe.so")  # M3D_VARIANT=probe M3D_NVCC_EXTRA=-DM3D_PROBE python -m m3dssd_b200.build
from m3dssd_b200 import ops

N, H, W, Cin, Cout, R = 8, 48, 160, 128, 128, 3
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(N, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16)
w = torch.randn(Cout, Cin, R, R, device="cuda", generator=g) / (Cin * R * R) ** 0.5
wp, _ = ops.pack_conv_weight(w)
b = torch.zeros(Cout, device="cuda")
om = torch.zeros(N, H, W, 32, device="cuda")
om[..., :18] = torch.randn(N, H, W, 18, device="cuda", generator=g) * 2.0
om[..., 18:27] = torch.randn(N, H, W, 9, device="cuda", generator=g)
out = torch.empty(N, H, W, Cout, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.conv2d_nhwc([x], wp, out, R=R, S=R, stride=1, pad=1, Cout=Cout, bias=b, slope=0.01, om=om, sigmoid_mask=True)
torch.cuda.synchronize()
buf = np.zeros(32 * 8, dtype=np.int64)
L = _lib.lib()
L.m3d_dcn_debug_read.argtypes = [C.c_void_p, C.c_int]
L.m3d_dcn_debug_read(buf.ctypes.data, buf.size)
a = buf.reshape(32, 8)
t0 = a[31, 0]
print("table wait at tile start: %d clk" % (a[31, 1] - a[31, 0]))
print("epilogue of tile 0 (run by producer warp 0 after the first k-block of tile 1): entry %d, accumulator ready +%d, drained +%d clk"
      % (a[31, 2] - t0, a[31, 3] - a[31, 2], a[31, 4] - a[31, 3]))
print("kb:  P start  P slot free  P half0 done  P arrived | M full seen  M issued   (clk since tile start)")
for kb in range(31):
    r = a[kb]
    print("%2d: %8d %8d %8d %8d | %8d %8d" % ((kb,) + tuple(int(v - t0) if v else -1 for v in r[:6])))
