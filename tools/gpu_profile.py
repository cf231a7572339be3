"""Per-op device times of the fused engine at the benchmark shape (development aid).

    python tools/gpu_profile.py [--batch 8] [--precision bf16] [--attention ANAB]
Writes gpurun_out/profile_ops.txt (sorted by time, with achieved TFLOP/s and GB/s per op).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from m3dssd_b200 import synth  # noqa: E402
from m3dssd_b200.model.M3d_inference_align import build  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--attention", default=None)
    ap.add_argument("--crop", default="384x1280")
    ap.add_argument("--offset-sigma", type=float, default=2.0, help="sigma of the synthetic DCN offsets, pixels")
    a = ap.parse_args()
    H, W = [int(v) for v in a.crop.split("x")]
    conf = synth.make_conf(attention=a.attention, crop_size=(H, W))
    net = build(conf, "test")
    synth.randomize_weights(net, offset_sigma_px=a.offset_sigma)
    net = net.cuda()
    eng = net.engine(a.batch, H, W, precision=a.precision, use_graph=False)
    eng.forward(synth.make_images(a.batch, (H, W)).cuda())
    prof = eng.profile(iters=5)
    total = sum(p["ms"] for p in prof)
    lines = ["%-34s %-12s %9s %7s %9s %9s" % ("op", "kind", "ms", "share", "TFLOP/s", "GB/s")]
    for p in sorted(prof, key=lambda q: -q["ms"]):
        t = p["ms"] * 1e-3
        lines.append("%-34s %-12s %9.4f %6.1f%% %9.1f %9.1f" % (p["name"], p["kind"], p["ms"], 100 * p["ms"] / total,
                                                                 p["flops"] / t / 1e12, p["bytes"] / t / 1e9))
    lines.append("total forward (eager, per-op events): %.3f ms for batch %d" % (total, a.batch))
    # end-to-end timings
    for stage, graph in (("forward", True), ("detect", True), ("detect", False)):
        e = net.engine(a.batch, H, W, precision=a.precision, use_graph=graph)
        for _ in range(3):
            e.run(None, stage)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            e.run(None, stage)
        e1.record()
        torch.cuda.synchronize()
        lines.append("stage=%s graph=%s: %.3f ms/step" % (stage, graph, e0.elapsed_time(e1) / 10))
    out = "\n".join(lines)
    print(out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "profile_ops_%s%s.txt" % (a.precision, "_anab" if a.attention else "")), "w") as fh:
        fh.write(out + "\n")


if __name__ == "__main__":
    main()
