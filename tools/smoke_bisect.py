"""Run the smoke comparison (96x320 engine vs oracle) under development toggles (development aid)."""
import os
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys; sys.path.insert(0, %r)
import torch
from m3dssd_b200 import synth
from m3dssd_b200.model.M3d_inference_align import build as build_net
from oracle import ref_model as RM
conf = synth.make_conf(attention=None, center_align=True, shape_align=True, crop_size=(96, 320))
net = build_net(conf, "test")
sd = synth.randomize_weights(net)
img = synth.make_images(1, (96, 320))
ref_out = RM.RefModel(sd, conf, dcn="tv").forward(img)
net = net.cuda()
eng = net.engine(1, 96, 320, precision="bf16", use_graph=False)
outs = eng.forward(img.cuda())
torch.cuda.synchronize()
res = []
for name, o, r in zip(("cls", "prob", "bbox_2d", "bbox_3d"), outs, ref_out):
    d = (o.cpu() - r).abs()
    res.append("%%s %%.3f" %% (name, float((d > 0.05 * r.abs().max()).float().mean())))
print("  ".join(res))
''' % ROOT
for env in ({}, {"M3D_NO_HALO": "1"}, {"M3D_DCN_LEGACY": "1"}, {"M3D_KSUB": "1"}, {"M3D_PDL": "0"},
            {"M3D_NO_HALO": "1", "M3D_DCN_LEGACY": "1", "M3D_KSUB": "1"}):
    e = dict(os.environ)
    e.update(env)
    p = subprocess.run([sys.executable, "-c", CODE], env=e, capture_output=True, text=True)
    print(env, "->", p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-400:], flush=True)
