"""Development aid: top CUDA kernels of the native training step (torch profiler)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from m3dssd_b200 import synth, train
from m3dssd_b200.model.M3d_inference_align import build
B, CROP = 4, (384, 1280)
conf = synth.make_conf(attention=None, center_align=False, shape_align=False, crop_size=CROP, batch_size=B)
net = build(conf, "train"); synth.randomize_weights(net, calibrate=False); net = net.cuda()
step = train.TrainStep(net, conf, native=True)
x = synth.make_images(B, CROP).cuda()
labels, t2, t3 = train.surrogate_targets(conf, B, "cuda")
for _ in range(3): step(x, labels, t2, t3)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2): step(x, labels, t2, t3)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
