import sys; sys.path.insert(0, "/root/repo")
import torch
from m3dssd_b200 import synth
from m3dssd_b200.model.M3d_inference_align import build
for cal in (True, False):
    conf = synth.make_conf(attention=None, crop_size=(384, 1280))
    net = build(conf, "test"); synth.randomize_weights(net, calibrate=cal); net = net.cuda()
    eng = net.engine(8, 384, 1280, precision="bf16", use_graph=False)
    x = synth.make_images(8, (384, 1280)).cuda()
    eng.detect(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(20_000_000)
    e0.record(); eng._run_decode(); e1.record(); torch.cuda.synchronize()
    print("calibrated=%s decode/top-K stage: %.3f ms, kept %s" % (cal, e0.elapsed_time(e1), eng.num_keep.tolist()))
