"""Development aid: gradient agreement of the native training path with torch autograd (fp32 and bf16 autocast)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from m3dssd_b200 import synth, train
from m3dssd_b200.model.M3d_inference_align import build

conf = synth.make_conf(attention=None, center_align=False, shape_align=False, crop_size=(96, 320), batch_size=2)
net = build(conf, "train")
sd = synth.randomize_weights(net)
x = synth.make_images(2, (96, 320)).cuda()
labels, t2, t3 = train.surrogate_targets(conf, 2, "cuda", fg_per_image=60)


def grads(model, autocast=False):
    model.train()
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        cls, prob, b2, b3, _ = model(x)
    loss = train.surrogate_loss(cls, b2, b3, labels, t2, t3)
    loss.backward()
    return float(loss), {n: p.grad.float().flatten().clone() for n, p in model.named_parameters() if p.grad is not None}


ref = build(conf, "train").cuda(); ref.load_state_dict(sd)
l32, g32 = grads(ref)
l16, g16 = grads(ref, autocast=True)
net = net.cuda(); train.enable(net)
ln, gn = grads(net)
print("loss fp32 %.5f  autocast %.5f  native %.5f" % (l32, l16, ln))
def cos(a, b): return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))
names = [n for n in g32 if n.endswith("weight") and g32[n].numel() > 64]
for n in names[::6]:
    print("%-58s native~fp32 %.4f  autocast~fp32 %.4f  native~autocast %.4f  |g| %.2e" % (n, cos(gn[n], g32[n]), cos(g16[n], g32[n]), cos(gn[n], g16[n]), float(g32[n].norm())))
