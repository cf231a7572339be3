"""Generate the committed golden fixtures from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Runs /root/reference's own Python modules on CPU through oracle/ref_harness.py
(its CUDA-only DCNv2 extension replaced by torchvision.ops.deform_conv2d, which
the C oracle is pinned against) and its own lib/nms/py_cpu_nms.py, on seeded
synthetic inputs that tests regenerate from the same seeds.  Outputs are stored
sub-sampled (every STRIDE-th row) plus whole-tensor checksums to stay small.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from m3dssd_b200 import synth  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
STRIDE = 11
CONFIGS = {
    "base": dict(attention=None, center_align=False, shape_align=False),
    "align": dict(attention=None, center_align=True, shape_align=True),
    "anab": dict(attention="ANAB", center_align=True, shape_align=True),
    # the backbone of the reference's shipped configs (scripts/config/kitti_3d_base.py:46): Bottleneck blocks, residual roots
    "dla102": dict(attention=None, center_align=False, shape_align=False, back_bone="dla102"),
}
CROP = (96, 320)


def nms_boxes(n, seed):
    rng = np.random.default_rng(seed)
    xy = rng.random((n, 2)) * np.array([1200, 350])
    wh = rng.random((n, 2)) * 120 + 4
    sc = rng.permutation(n).astype(np.float32) / max(n, 1)  # tie-free
    return np.concatenate([xy, xy + wh, sc[:, None]], 1).astype(np.float32)


def main():
    ns = RH.load_reference()
    # --- DCNv2 known-answer test of the reference (model/DCNv2/test.py:32-65), stated as data
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 2, 4, 4, generator=g)
    np.savez(os.path.join(HERE, "dcn_zero_offset.npz"), input=x.numpy(), expected_2x_output=x.numpy())

    # --- NMS: the reference's own py_cpu_nms on seeded boxes
    out = {}
    for n, seed in [(1, 1), (63, 2), (64, 3), (65, 4), (500, 5), (3000, 6)]:
        d = nms_boxes(n, seed)
        out["keep_%d" % n] = np.asarray(ns.py_cpu_nms(d, 0.4), dtype=np.int32)
        out["seed_%d" % n] = np.int32(seed)
    np.savez_compressed(os.path.join(HERE, "nms_py_cpu.npz"), **out)

    # --- input transform: the reference's own Normalize class (lib/augmentations.py:44-57) + its channel swap
    import importlib
    aug = importlib.import_module("lib.augmentations")
    rng = np.random.default_rng(17)
    im = rng.integers(0, 256, (24, 40, 3), dtype=np.uint8)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]  # scripts/config/kitti_3d_base.py:42-43
    out, _ = aug.Normalize(mean, std)(im.copy())
    out = out[:, :, (2, 1, 0)]  # lib/dataloader.py:942-947 (cv2.cvtColor BGR2RGB == this permutation)
    np.savez_compressed(os.path.join(HERE, "preprocess_u8.npz"), image=im, mean=np.float32(mean), std=np.float32(std),
                        expected_chw=np.ascontiguousarray(out.transpose(2, 0, 1)).astype(np.float32))

    # --- model forward + decode through the unmodified reference modules
    torch.Tensor.cuda = lambda self, *a, **k: self  # im_detect_3d calls .cuda(); stay on CPU
    torch.cuda.FloatTensor = torch.FloatTensor
    for name, kw in CONFIGS.items():
        conf = ns.EasyDict(synth.make_conf(crop_size=CROP, **kw))
        if kw.get("back_bone") == "dla102":
            conf.pre_train = None  # the reference's dla102() downloads ImageNet weights unless this `is None` (pose_dla_dcn.py:439)
        from m3dssd_b200.model.M3d_inference_align import build as our_build
        sd = synth.randomize_weights(our_build(synth.make_conf(crop_size=CROP, **kw), "test"))
        net = ns.rpn.build(conf, "test")  # the unmodified reference modules, fed the same state_dict
        net.load_state_dict(sd)
        x = synth.make_images(2, CROP)
        with torch.no_grad():
            cls, prob, b2, b3, feat_size, rois = net(x)
        fix = dict(stride=np.int32(STRIDE))
        for k, t in (("cls", cls), ("prob", prob), ("bbox_2d", b2), ("bbox_3d", b3)):
            a = t.numpy()
            fix[k] = a[:, ::STRIDE].copy()
            fix[k + "_sum"] = np.float64(a.astype(np.float64).sum())
            fix[k + "_abssum"] = np.float64(np.abs(a.astype(np.float64)).sum())
        fix["rois"] = rois.numpy()[::STRIDE].copy()
        fix["fg_frac"] = np.float64(((1 - prob[..., 0]) > 0.5).float().mean().item())
        # unmodified im_detect_3d (lib/rpn_util.py:1416-1563) on image 0
        obj = types.SimpleNamespace(imH=CROP[0], imW=CROP[1], p2=np.eye(4), scale_factor=1.0)
        ab = ns.rpn_util.im_detect_3d(x[0:1].clone(), net, conf, obj)
        fix["aboxes"] = ab.astype(np.float32)
        np.savez_compressed(os.path.join(HERE, "ref_model_%s_96x320.npz" % name), **fix)
        print(name, "fg_frac=%.4f" % fix["fg_frac"], "aboxes", ab.shape,
              {k: v.shape for k, v in fix.items() if hasattr(v, "shape") and v.ndim > 0})
    hill_climb_golden(ns)


def hill_climb_golden(ns):
    """Post-NMS refinement: the unmodified reference functions (lib/rpn_util.py hill_climb / test_projection /
    project_3d, lib/util.py convertAlpha2Rot / convertRot2Alpha) driven by the loop body of test_kitti_3d
    (lib/rpn_util.py:1801-1852) on seeded KITTI-like detections (oracle.hill_climb.synthetic_detections).
    NOTE: run here under NumPy 2 (NEP 50) the reference keeps float32 in a few scalar expressions that its pinned
    NumPy 1.x promoted to float64; the float64 oracle therefore matches these vectors to ~2e-6, not bit for bit."""
    import importlib
    import math
    from oracle import hill_climb as HC
    ru = ns.rpn_util
    util = importlib.import_module("lib.util")
    out = {}
    for seed in (5, 6, 7):
        rows, p2 = HC.synthetic_detections(64, seed)
        p2_inv = np.linalg.inv(p2)
        res = []
        for i in range(40):
            box = rows[i]
            if not box[4] >= 0.75:
                continue
            x1, y1, x2, y2 = box[0], box[1], box[2], box[3]
            width, height = (x2 - x1 + 1), (y2 - y1 + 1)
            x3d, y3d, z3d, w3d, h3d, l3d, ry3d = box[6], box[7], box[8], box[9], box[10], box[11], box[12]
            coord3d = np.linalg.inv(p2).dot(np.array([x3d * z3d, y3d * z3d, 1 * z3d, 1]))
            ry3d = util.convertAlpha2Rot(ry3d, coord3d[2], coord3d[0])
            z3d, ry3d, _ = ru.hill_climb(p2, p2_inv, np.array([x1, y1, width, height]), x3d, y3d, z3d, w3d, h3d, l3d, ry3d,
                                         step_r_init=0.3 * math.pi, r_lim=0.01)
            coord3d = np.linalg.inv(p2).dot(np.array([x3d * z3d, y3d * z3d, 1 * z3d, 1]))
            alpha = util.convertRot2Alpha(ry3d, coord3d[2], coord3d[0])
            res.append([box[5] - 1, alpha, x1, y1, x2, y2, h3d, w3d, l3d, coord3d[0], coord3d[1] + h3d / 2, coord3d[2],
                        ry3d, box[4]])
        out["refined_%d" % seed] = np.asarray(res, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "hill_climb.npz"), **out)
    print("hill_climb", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if "--hill-climb-only" in sys.argv:
        hill_climb_golden(RH.load_reference())
    else:
        main()
