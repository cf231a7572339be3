"""Layer-wise parity of the fused engine against the CPU oracle model (development aid).

    python tools/gpu_probe_model.py [--crop 96x320] [--batch 2] [--precision fp32|bf16] [--attention ANAB]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from m3dssd_b200 import synth  # noqa: E402
from m3dssd_b200.model.M3d_inference_align import build  # noqa: E402
from oracle import ref_model as RM  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def rms(a, b):
    return float(((a - b).double().pow(2).mean().sqrt()) / (b.double().pow(2).mean().sqrt() + 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--crop", default="96x320")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--attention", default=None)
    ap.add_argument("--no-align", action="store_true")
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    H, W = [int(v) for v in a.crop.split("x")]
    conf = synth.make_conf(attention=a.attention, center_align=not a.no_align, shape_align=not a.no_align,
                           crop_size=(H, W))
    net = build(conf, "test")
    sd = synth.randomize_weights(net)
    x = synth.make_images(a.batch, (H, W))
    t0 = time.time()
    oracle = RM.RefModel(sd, conf, dcn="tv")
    ref = oracle.forward(x)
    print("oracle forward %.2fs" % (time.time() - t0))
    net = net.cuda()
    eng = net.engine(a.batch, H, W, precision=a.precision, use_graph=a.graph)
    out = eng.forward(x.cuda())
    torch.cuda.synchronize()
    print("engine launches/step:", eng.launches_per_step())
    for name in ("level0", "level1", "level2", "level3", "level4", "level5", "feat", "feats_shape", "feats_align2d",
                 "feats_align3d", "feats_gl"):
        if name in oracle.taps and name in eng.named:
            got = eng.activation_nchw(name).cpu()
            print("%-14s rel-max-err %.3e  rel-rms-err %.3e" % (name, rel(got, oracle.taps[name]), rms(got, oracle.taps[name])))
    for n, t in eng.bufs.items():
        key = "base." + n
        if key in oracle.taps:
            r = oracle.taps[key]
            got = t[..., :r.shape[1]].float().permute(0, 3, 1, 2).cpu()
            d = (got - r).abs()
            print("  %-28s rel-max-err %.3e  mean-abs-err %.3e (ref mean abs %.3e)" % (
                n, float(d.max() / (r.abs().max() + 1e-12)), float(d.mean()), float(r.abs().mean())))
    fg = eng.fg_max.cpu()
    fgo = oracle.taps["fg_prob"].max(dim=1)[0]
    print("%-14s abs-max-err %.3e  argmax agree %.5f" % ("fg_prob", float((fg - fgo).abs().max()),
          float((eng.fg_arg.cpu().long() == oracle.taps["fg_prob"].max(dim=1)[1]).float().mean())))
    for n, o, r in zip(("cls", "prob", "bbox_2d", "bbox_3d"), out, ref):
        d = (o.cpu() - r).abs()
        tol = 1e-3 * r.abs().max()
        print("%-14s rel-max-err %.3e  rel-rms-err %.3e  frac(|err| > 1e-3*scale) = %.2e" % (
            n, rel(o.cpu(), r), rms(o.cpu(), r), float((d > tol).float().mean())))
    # detection tail
    kept, num = eng.detect(x.cuda())
    torch.cuda.synchronize()
    for b in range(a.batch):
        pre, keep, kept_ref = oracle.detect(ref, b)
        n = int(num[b].item())
        mo = min(eng.max_out, kept_ref.shape[0])
        d = (kept[b, :mo].cpu() - kept_ref[:mo]).abs().max().item() if mo else 0.0
        print("image %d: kept %d (oracle %d); first %d rows max abs diff %.3e; idx agree(top-k) %.4f" % (
            b, n, kept_ref.shape[0], mo, d,
            float((eng.dets[b, :, 13].cpu() == pre[:, 13]).float().mean())))


if __name__ == "__main__":
    main()
