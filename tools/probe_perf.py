"""Time single conv layers of the DLA-34 trunk in isolation, with the kernel's development probes
(M3D_DBG bits: 1 skip A loads, 2 skip B loads, 4 skip the epilogue body, 8 skip the MMAs) to see which
pipe bounds them.  Development aid; results go to gpurun_out/probe_perf.log."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from m3dssd_b200 import ops  # noqa: E402

SHAPES = [
    ("l0 64->64 3x3 @192x640", 8, 192, 640, 64, 64, 3, 1),
    # name, N, H, W, Cin, Cout, R, stride
    ("l3 128->128 3x3 @48x160", 8, 48, 160, 128, 128, 3, 1),
    ("l4 256->256 3x3 @24x80", 8, 24, 80, 256, 256, 3, 1),
    ("l5 512->512 3x3 @12x40", 8, 12, 40, 512, 512, 3, 1),
    ("l2 64->64 3x3 @96x320", 8, 96, 320, 64, 64, 3, 1),
    ("head 128->256 1x1 @48x160", 8, 48, 160, 128, 256, 1, 1),
    ("head 256->256 1x1 @48x160", 8, 48, 160, 256, 256, 1, 1),
]


def time_conv(N, H, W, Cin, Cout, R, stride, ksub, iters=20, res=False):
    if ksub:
        os.environ["M3D_KSUB"] = str(ksub)
    else:
        os.environ.pop("M3D_KSUB", None)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, R, R, device="cuda", generator=g) / (Cin * R * R) ** 0.5)
    wp, _ = ops.pack_conv_weight(w)
    b = torch.zeros(Cout, device="cuda")
    P, Q = H // stride, W // stride
    out = torch.empty(N, P, Q, Cout, device="cuda", dtype=torch.bfloat16)
    r = torch.randn(N, P, Q, Cout, device="cuda", generator=g).to(torch.bfloat16) if res else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def run():
        ops.conv2d_nhwc([x], wp, out, R=R, S=R, stride=stride, pad=R // 2, Cout=Cout, bias=b, slope=0.01, res=r)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    # the host path (ctypes + tensor-map encoding) costs more than these kernels: time graph replays
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
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g1.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / iters)
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "probe_perf.log"), "w")
    for name, N, H, W, Cin, Cout, R, stride in SHAPES:
        gf = 2.0 * N * (H // stride) * (W // stride) * Cin * Cout * R * R / 1e9
        for dbg in (0, 1):
            if dbg:
                os.environ.pop("M3D_NO_PAIR", None)
            else:
                os.environ["M3D_NO_PAIR"] = "1"
            ksub = dbg
            cold, warm = time_conv(N, H, W, Cin, Cout, R, stride, 0)
            line = "%-28s pair=%d  best %7.1f us  median %7.1f us  (%.0f TF/s)" % (name, ksub, cold, warm, gf / cold * 1e-3)
            print(line, flush=True)
            log.write(line + "\n")
    os.environ.pop("M3D_KSUB", None)
    log.close()


if __name__ == "__main__":
    main()
