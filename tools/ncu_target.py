"""Short target for Nsight Compute: build the bf16 engine at the benchmark shape and run the forward
path eagerly (no CUDA graph) `--iters` times, plus the detection tail once."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from m3dssd_b200 import synth  # noqa: E402
from m3dssd_b200.model.M3d_inference_align import build  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--attention", default=None)
ap.add_argument("--ops", default="", help="comma-separated engine op names: profile only these (cudaProfilerStart/Stop range)")
a = ap.parse_args()
conf = synth.make_conf(attention=a.attention, crop_size=(384, 1280))
net = build(conf, "test")
synth.randomize_weights(net)
net = net.cuda()
eng = net.engine(a.batch, 384, 1280, precision="bf16", use_graph=False)
x = synth.make_images(a.batch, (384, 1280)).cuda()
for _ in range(a.iters):
    eng.detect(x)
torch.cuda.synchronize()
if a.ops == "step":  # one whole warm step inside the profiler range (ncu --profile-from-start off)
    torch.cuda.profiler.start()
    eng.detect(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
elif a.ops:
    want = a.ops.split(",")
    torch.cuda.profiler.start()
    for name in want:
        hits = [i for i, m in enumerate(eng.meta) if m["name"] == name]
        if name == "tail":
            eng._run_decode()
            eng._run_nms()
            continue
        assert hits, "no op named %s; have %s" % (name, [m["name"] for m in eng.meta])
        eng.ops[hits[0]]()
    if "tail" in a.ops:
        pass
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("launches per step:", eng.launches_per_step())
