"""Small end-to-end run for compute-sanitizer (development aid): bf16 engine incl. ANAB, tail, refinement, 2 steps."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from m3dssd_b200 import synth
from m3dssd_b200.model.M3d_inference_align import build
from m3dssd_b200.lib.rpn_util import refine_detections
for att, crop, batch in ((None, (96, 320), 2), ("ANAB", (96, 320), 1), (None, (192, 640), 1)):
    conf = synth.make_conf(attention=att, center_align=True, shape_align=True, crop_size=crop)
    net = build(conf, "test"); synth.randomize_weights(net); net = net.cuda().eval()
    eng = net.engine(batch, crop[0], crop[1], precision="bf16", use_graph=False)
    x = synth.make_images(batch, crop).cuda()
    for _ in range(2):
        kept, num = eng.detect(x)
    rows, valid = refine_detections(kept, num, np.eye(4) + np.array([[720, 0, 600, 45], [0, 720, 170, 0], [0, 0, 0, 0], [0, 0, 0, 0.0]]))
    torch.cuda.synchronize()
    print(att, crop, "kept", num.tolist(), "valid", int(valid.sum()))
# round-2 additions: flattened outputs on demand, device-side targets + loss
from m3dssd_b200.lib.targets import compute_targets_batch
from m3dssd_b200.lib.loss.rpn_3d import RPN_3D_loss_smp
eng.flatten_outputs()
conf = synth.loss_conf(synth.make_conf(crop_size=(96, 320)))
tar = compute_targets_batch(conf, [synth.make_gts(conf, 5, 1, seed=s) for s in range(2)], (12, 40))
M = tar["labels"].shape[1]
cls = torch.randn(2, M, 4, device="cuda", requires_grad=True)
loss, _ = RPN_3D_loss_smp(conf).cuda()(cls, torch.softmax(cls, 2), tar["bbox_2d"] + 0.01, torch.zeros(2, M, 7, device="cuda"), tar)
loss.backward()
torch.cuda.synchronize()
print("targets fg", int(tar["labels_fg"].sum()), "loss", float(loss))
print("done")
