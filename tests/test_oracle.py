"""CPU: the oracle against the reference's golden vectors (tests/golden/, generated from the
unmodified reference by tests/golden/make_golden.py) and against independent implementations."""
import os

import numpy as np
import pytest
import torch

from m3dssd_b200 import synth
from oracle import oracle as O
from oracle import ref_model as RM

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_dcn_zero_offset_known_answer():
    """model/DCNv2/test.py:32-65: zero offsets, mask = sigmoid(0), identity taps => 2*DCNv2(x) == x."""
    g = np.load(os.path.join(GOLD, "dcn_zero_offset.npz"))
    x = g["input"]
    n, c, h, w = x.shape
    wt = np.zeros((c, c, 3, 3), np.float32)
    for i in range(c):
        wt[i, i, 1, 1] = 1.0
    out = O.dcn_v2_forward(x, np.zeros((n, 18, h, w), np.float32), np.full((n, 9, h, w), 0.5, np.float32), wt,
                           np.zeros(c, np.float32), 1, 1, 1, 1)
    assert np.abs(g["expected_2x_output"] - 2 * out).max() < 1e-10


@pytest.mark.parametrize("shape", [(2, 8, 9, 11, 6, 3, 1, 1, 1), (1, 8, 10, 13, 4, 3, 2, 1, 2),
                                   (2, 4, 7, 9, 5, 1, 1, 0, 1), (1, 6, 8, 8, 3, 3, 1, 2, 1)])
def test_dcn_forward_backward_vs_torchvision(shape):
    import torchvision
    B, C, H, W, Co, k, s, p, dg = shape
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    Ho, Wo = O.dcn_out_shape(H, W, k, k, s, p, 1)
    off = (rng.standard_normal((B, 2 * dg * k * k, Ho, Wo)) * 3).astype(np.float32)
    m = rng.random((B, dg * k * k, Ho, Wo)).astype(np.float32)
    w = rng.standard_normal((Co, C, k, k)).astype(np.float32)
    b = rng.standard_normal(Co).astype(np.float32)
    o = O.dcn_v2_forward(x, off, m, w, b, s, p, 1, dg)
    t = torchvision.ops.deform_conv2d(torch.from_numpy(x), torch.from_numpy(off), torch.from_numpy(w),
                                      torch.from_numpy(b), stride=s, padding=p, mask=torch.from_numpy(m)).numpy()
    assert np.abs(o - t).max() < 2e-5
    xd, offd, md, wd, bd = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, off, m, w, b)]
    y = torchvision.ops.deform_conv2d(xd, offd, wd, bd, stride=s, padding=p, mask=md)
    gy = torch.tensor(rng.standard_normal(tuple(y.shape)))
    y.backward(gy)
    grads = O.dcn_v2_backward(x, off, m, w, gy.numpy(), s, p, 1, dg, dtype=np.float64)
    for a, ref in zip(grads, (xd, offd, md, wd, bd)):
        assert np.abs(a - ref.grad.numpy()).max() < 1e-10


def test_dcn_shape_error_like_reference():
    with pytest.raises(RuntimeError):  # THError in dcn_v2_cuda.c:36-38
        O.dcn_v2_forward(np.zeros((1, 4, 5, 5), np.float32), np.zeros((1, 18, 5, 5), np.float32),
                         np.zeros((1, 9, 5, 5), np.float32), np.zeros((2, 3, 3, 3), np.float32), np.zeros(2, np.float32),
                         1, 1, 1, 1)


def _boxes(n, seed):
    rng = np.random.default_rng(seed)
    xy = rng.random((n, 2)) * np.array([1200, 350])
    wh = rng.random((n, 2)) * 120 + 4
    sc = rng.permutation(n).astype(np.float32) / max(n, 1)
    return np.concatenate([xy, xy + wh, sc[:, None]], 1).astype(np.float32)


def test_nms_vs_reference_py_cpu_nms_golden():
    g = np.load(os.path.join(GOLD, "nms_py_cpu.npz"))
    for n in (1, 63, 64, 65, 500, 3000):
        d = _boxes(n, int(g["seed_%d" % n]))
        assert list(O.gpu_nms(d, 0.4)) == list(g["keep_%d" % n])
    assert O.gpu_nms(np.zeros((0, 5), np.float32), 0.4) == []


def test_nms_mask_consistent_with_sweep():
    d = _boxes(300, 11)
    order = d[:, 4].argsort()[::-1]
    sd = d[order]
    mask = O.nms_mask(sd, 0.4)
    removed = np.zeros(300, bool)
    keep = []
    for i in range(300):
        if removed[i]:
            continue
        keep.append(i)
        for j in range(i + 1, 300):
            if (int(mask[i, j // 64]) >> (j % 64)) & 1:
                removed[j] = True
    assert keep == list(O.nms_sorted(sd, 0.4))


@pytest.mark.parametrize("name,kw", [
    ("base", dict(attention=None, center_align=False, shape_align=False)),
    ("align", dict(attention=None, center_align=True, shape_align=True)),
    ("anab", dict(attention="ANAB", center_align=True, shape_align=True)),
    ("dla102", dict(attention=None, center_align=False, shape_align=False, back_bone="dla102")),
])
def test_model_restatement_vs_reference_golden(name, kw):
    """oracle/ref_model.py reproduces the unmodified reference modules' outputs (fixtures)."""
    from m3dssd_b200.model.M3d_inference_align import build  # state-dict source: same key names as the reference
    g = np.load(os.path.join(GOLD, "ref_model_%s_96x320.npz" % name))
    conf = synth.make_conf(crop_size=(96, 320), **kw)
    net = build(conf, "test")
    sd = synth.randomize_weights(net)
    m = RM.RefModel(sd, conf, dcn="tv")
    x = synth.make_images(2, (96, 320))
    out = m.forward(x)
    st = int(g["stride"])
    for k, t in zip(("cls", "prob", "bbox_2d", "bbox_3d"), out[:4]):
        a = t.numpy()
        assert np.abs(a[:, ::st] - g[k]).max() < 1e-4, k
        assert abs(a.astype(np.float64).sum() - float(g[k + "_sum"])) < 1e-3 * float(g[k + "_abssum"])
    assert np.abs(out[5].numpy()[::st] - g["rois"]).max() == 0
    # decode + NMS restatement vs the reference's own im_detect_3d
    # (the fixture ran im_detect_3d on image 0 alone, so do the same: batch size changes oneDNN's
    #  summation order by ~1e-7, enough to swap the rank of near-tied low scores)
    pre, keep, kept = m.detect(m.forward(x[0:1]), 0)
    ab = g["aboxes"]
    assert kept.shape == ab.shape
    if name == "anab":  # ANAB's bmm is not bit-reproducible between the two formulations (~6e-6)
        rows_ok = np.isclose(kept.numpy(), ab, rtol=1e-4, atol=1e-2).all(axis=1)
        assert rows_ok.mean() > 0.99
    else:
        assert np.abs(kept.numpy() - ab).max() == 0.0
    # the C DCN oracle and torchvision agree through the whole network to fp32 noise
    out_c = RM.RefModel(sd, conf, dcn="c").forward(x)
    for a, b in zip(out_c[:4], out[:4]):
        assert (a - b).abs().max() < 2e-2
        assert ((a - b).abs() > 1e-3 * b.abs().max()).float().mean() < 1e-3


def test_hill_climb_oracle_vs_reference_golden():
    """oracle/hill_climb.py (float64 restatement of lib/rpn_util.py:652-708, 921-970, 1801-1852, 2015-2050) against
    the unmodified reference functions run by tests/golden/make_golden.py.  Tolerance 2e-5: under NumPy 2 the
    reference keeps float32 in a few scalar expressions (see make_golden.hill_climb_golden)."""
    from oracle import hill_climb as HC
    fix = np.load(os.path.join(GOLD, "hill_climb.npz"))
    for seed in (5, 6, 7):
        rows, p2 = HC.synthetic_detections(64, seed)
        got = HC.refine_detections(rows, p2)
        ref = fix["refined_%d" % seed]
        assert got.shape == ref.shape and got.shape[0] > 5
        assert np.abs(got - ref).max() < 2e-5
        # the search really moves the rotation: otherwise the comparison is vacuous
        alpha0 = rows[:40][rows[:40, 4] >= 0.75][:, 12]
        assert np.abs(got[:, 1] - alpha0).max() > 0.05


def test_preprocess_oracle_vs_reference_normalize_golden():
    """oracle.preprocess_u8 == the reference's Normalize transform + BGR->RGB + CHW, bit for bit (fixture generated by
    tests/golden/make_golden.py from the unmodified lib/augmentations.py)."""
    g = np.load(os.path.join(GOLD, "preprocess_u8.npz"))
    got = O.preprocess_u8(g["image"], g["mean"], g["std"])
    assert got.dtype == np.float32 and np.array_equal(got, g["expected_chw"])


def test_preprocess_pad_oracle_vs_reference_preprocess_golden():
    """oracle.preprocess_pad_u8 == the reference's Preprocess (ConvertToFloat + cv2 Padding + Normalize) + BGR->RGB +
    CHW on ragged frames, bit for bit (fixture generated by tests/golden/make_golden_preprocess.py from the unmodified
    lib/augmentations.py with the real cv2); an over-sized frame is an error like cv2's negative border."""
    g = np.load(os.path.join(GOLD, "preprocess_pad_u8.npz"))
    ims = [g["image_%d" % k] for k in range(int(g["n"]))]
    got = O.preprocess_pad_u8(ims, g["size"], g["mean"], g["std"])
    assert got.dtype == np.float32
    for k in range(len(ims)):
        assert np.array_equal(got[k], g["expected_%d" % k]), k
    # a padded pixel is Normalize(0), not 0 (the reference pads before it normalises)
    h, w = ims[1].shape[:2]
    pad = got[1][:, h:, :]
    assert pad.size and np.all(pad[0] == pad[0].flat[0]) and abs(float(pad[0].flat[0])) > 1.0
    with pytest.raises(ValueError):
        O.preprocess_pad_u8([np.zeros((49, 64, 3), np.uint8)], g["size"], g["mean"], g["std"])



def test_kitti_writer_chain_vs_unmodified_reference_driver_golden():
    """The evaluation loop around the path, pinned to the UNMODIFIED reference driver: lib/rpn_util.py test_kitti_3d
    (:1753-1852) run on the reference network by tests/golden/make_golden_kitti_writer.py wrote two KITTI result files;
    the oracle chain (ref_model forward + im_detect_3d restatement, hill_climb.refine_detections) formatted by the
    product's kitti_result_lines must give the same files: same lines, same classes, numbers to float32 noise."""
    from m3dssd_b200.lib import rpn_util as RU
    from m3dssd_b200.model.M3d_inference_align import build
    from oracle import hill_climb as HC
    g = np.load(os.path.join(GOLD, "kitti_writer_align_96x320.npz"))
    conf = synth.make_conf(crop_size=(96, 320), attention=None, center_align=True, shape_align=True)
    sd = synth.randomize_weights(build(conf, "test"))
    m = RM.RefModel(sd, conf, dcn="tv")
    x = synth.make_images(2, (96, 320))
    n_lines = 0
    for i in range(2):
        _, _, kept = m.detect(m.forward(x[i:i + 1]), 0)
        rows = HC.refine_detections(kept.numpy(), g["p2"][i], hill_climbing=True, max_out=int(conf.nms_topN_post))
        text = RU.kitti_result_lines(rows, np.ones(len(rows), dtype=bool), conf.lbls)
        ref = str(g["text_%d" % i])
        got_l, ref_l = text.splitlines(), ref.splitlines()
        assert len(got_l) == len(ref_l) and len(ref_l) > 0, (i, len(got_l), len(ref_l))
        assert text.endswith("\n") and ref.endswith("\n")
        for a, b in zip(got_l, ref_l):
            ta, tb = a.split(" "), b.split(" ")
            assert len(ta) == len(tb) == 16 and ta[:3] == tb[:3], (a, b)  # class, "-1", "-1"
            assert all(len(t.split(".")[1]) == 6 for t in ta[3:])  # '{:.6f}'
            va, vb = np.array(ta[3:], dtype=np.float64), np.array(tb[3:], dtype=np.float64)
            assert np.allclose(va, vb, rtol=2e-5, atol=2e-5), (a, b)
        n_lines += len(ref_l)
        exact = sum(a == b for a, b in zip(got_l, ref_l))
        assert exact >= 0.5 * len(ref_l), (i, exact, len(ref_l))  # most lines agree to the last printed digit
    assert n_lines == 50
