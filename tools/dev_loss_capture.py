import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from m3dssd_b200 import synth, train
from m3dssd_b200.lib.loss.rpn_3d import RPN_3D_loss_smp
from m3dssd_b200.model.M3d_inference_align import build

def attempt(tag, native=True, surrogate=False, **over):
    conf = synth.loss_conf(synth.make_conf(attention=None, center_align=False, shape_align=False, crop_size=(96, 320), batch_size=2))
    for k, v in over.items():
        conf[k] = v
    net = build(conf, "train")
    synth.randomize_weights(net)
    synth.condition_for_training(net)
    net = net.cuda()
    x = synth.make_images(2, (96, 320)).cuda()
    tar = train.targets_to(synth.make_targets(conf, 2, fg_per_image=60), "cuda")
    crit = RPN_3D_loss_smp(conf).cuda()
    step = train.TrainStep(net, conf, lr=0.002, graph=True, warmup=2, criterion=None if surrogate else crit, native=native)
    if surrogate:
        tar = None
        tgs = train.surrogate_targets(conf, 2, "cuda", fg_per_image=60)
    else:
        tgs = (tar,)
    try:
        losses = [float(step(x, *tgs).detach()) for _ in range(5)]
        print(tag, "OK", ["%.4f" % v for v in losses], flush=True)
    except Exception as e:
        print(tag, "FAILED:", str(e).splitlines()[0], flush=True)

which = sys.argv[1]
if which == "all":
    attempt("all terms")
elif which == "noiou":
    attempt("iou_2d_lambda=0", iou_2d_lambda=0)
elif which == "no3d":
    attempt("bbox_3d_lambda=0", bbox_3d_lambda=0)
elif which == "nocls":
    attempt("cls_2d_lambda=0", cls_2d_lambda=0)
elif which == "surrogate":
    attempt("surrogate loss", surrogate=True)
elif which == "torchconv":
    attempt("criterion, torch convs (native=False)", native=False)
elif which == "only_cls":
    attempt("only cls", iou_2d_lambda=0, bbox_3d_lambda=0)
