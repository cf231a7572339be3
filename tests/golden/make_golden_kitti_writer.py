"""Golden KITTI result files written by the UNMODIFIED reference evaluation driver, lib/rpn_util.py test_kitti_3d
(:1753-1852): the reference network (align configuration, synthetic weights) on two synthetic 96x320 images through
the reference's own im_detect_3d, score cut, hill_climb and format string, on the CPU.  The AP evaluation the function
runs after the files are written needs the KITTI label directory and is cut off there (the exception is expected).
Run in the build container only:   python tests/golden/make_golden_kitti_writer.py
-> tests/golden/kitti_writer_align_96x320.npz (the two files' text + the calibration used)."""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from m3dssd_b200 import synth  # noqa: E402
from oracle import hill_climb as HC  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402

CROP = (96, 320)


def main():
    ns = RH.load_reference()
    torch.Tensor.cuda = lambda self, *a, **k: self  # im_detect_3d calls .cuda(); stay on CPU
    torch.cuda.FloatTensor = torch.FloatTensor
    kw = dict(attention=None, center_align=True, shape_align=True)
    conf = ns.EasyDict(synth.make_conf(crop_size=CROP, **kw))
    conf.hill_climbing = True
    conf.pre_compute_target = False
    from m3dssd_b200.model.M3d_inference_align import build as our_build
    sd = synth.randomize_weights(our_build(synth.make_conf(crop_size=CROP, **kw), "test"))
    net = ns.rpn.build(conf, "test")  # the unmodified reference modules, fed the same state_dict
    net.load_state_dict(sd)
    net.eval()
    x = synth.make_images(2, CROP)
    _, p2 = HC.synthetic_detections(4, 0, hw=CROP)
    p2s = [p2.copy(), p2.copy()]
    p2s[1][0, 0] *= 1.02
    dataset = [(x[i:i + 1].clone(), ns.EasyDict(dict(id="%06d" % i, p2=p2s[i], scale_factor=1.0, imH=CROP[0], imW=CROP[1])))
               for i in range(2)]
    out_dir = tempfile.mkdtemp()
    try:
        with torch.no_grad():
            ns.rpn_util.test_kitti_3d(dataset, net, conf, out_dir, "/nonexistent", use_log=False)
    except Exception as e:  # the label directory of the AP evaluation does not exist here
        print("stopped after the result files, as expected: %s: %s" % (type(e).__name__, str(e)[:100]))
    fix = dict(p2=np.stack(p2s))
    for i in range(2):
        text = open(os.path.join(out_dir, "%06d.txt" % i)).read()
        fix["text_%d" % i] = np.array(text)
        print(i, len(text.splitlines()), "lines;", text.splitlines()[0][:100] if text else "")
    np.savez_compressed(os.path.join(HERE, "kitti_writer_align_96x320.npz"), **fix)


if __name__ == "__main__":
    main()
