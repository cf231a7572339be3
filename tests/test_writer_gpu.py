"""The evaluation driver around the path: m3dssd_b200.lib.rpn_util.test_kitti_3d (the reference's lib/rpn_util.py:1753-1852
loop: detect -> top 40 -> score cut -> hill_climb -> KITTI txt per image), batched, against the per-image composition
of im_detect_3d and the oracle's restatement of that loop."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CROP = (96, 320)


def _parse(path):
    rows = []
    with open(path) as f:
        for line in f:
            t = line.split()
            assert len(t) == 16 and t[1] == "-1" and t[2] == "-1", line
            rows.append((t[0], [float(v) for v in t[3:]]))
    return rows


def _setup(n_images=5):
    from m3dssd_b200 import synth
    from m3dssd_b200.model.M3d_inference_align import build
    from oracle import hill_climb as HC
    conf = synth.make_conf(crop_size=CROP)
    conf.hill_climbing = True
    net = build(conf, "test")
    synth.randomize_weights(net)
    net = net.cuda().eval()
    _, p2 = HC.synthetic_detections(4, 0, hw=CROP)
    frames = synth.make_images_u8(n_images, CROP, seed=11).numpy()
    sizes = [(96, 320), (90, 301), (96, 316), (93, 320), (88, 310)][:n_images]
    frames = [np.ascontiguousarray(frames[i][:h, :w]) for i, (h, w) in enumerate(sizes)]
    objs = []
    for i, (h, w) in enumerate(sizes):
        q = p2.copy()
        q[0, 0] *= 1.0 + 0.01 * i  # a different camera per image
        objs.append(types.SimpleNamespace(id="%06d" % i, p2=q, scale_factor=1.0, imH=h, imW=w))
    return conf, net, frames, objs


def test_test_kitti_3d_batched_writer(tmp_path):
    from m3dssd_b200.lib import rpn_util as RU
    from oracle import hill_climb as HC
    from oracle import oracle as O
    conf, net, frames, objs = _setup()
    mean, std = conf.image_means, conf.image_stds
    pre = O.preprocess_pad_u8(frames, CROP, mean, std)  # what the reference's test loader yields
    loader = [(torch.from_numpy(pre[i])[None], objs[i]) for i in range(len(frames))]

    # (1) batched, from the loader's normalised tensors; 5 images in batches of 2 (short last batch)
    d2 = str(tmp_path / "b2")
    written = RU.test_kitti_3d(loader, net, conf, d2, batch_size=2)
    assert [os.path.basename(p) for p in written] == ["%06d.txt" % i for i in range(5)]
    # (2) from the raw ragged uint8 frames (Preprocess on the device): the same files, byte for byte
    d2u = str(tmp_path / "b2u8")
    RU.test_kitti_3d([(frames[i], objs[i]) for i in range(5)], net, conf, d2u, batch_size=2)
    for i in range(5):
        assert open(os.path.join(d2, "%06d.txt" % i)).read() == open(os.path.join(d2u, "%06d.txt" % i)).read(), i
    # (3) one image at a time through the dict form of the loader items
    d1 = str(tmp_path / "b1")
    RU.test_kitti_3d([{"input": loader[i][0], "target": {"meta": vars(objs[i])}} for i in range(5)], net, conf, d1,
                     batch_size=1)

    total = 0
    for i in range(5):
        got = _parse(os.path.join(d2, "%06d.txt" % i))
        one = _parse(os.path.join(d1, "%06d.txt" % i))
        # per-image composition the reference runs: im_detect_3d, then the loop body restated by the oracle
        aboxes = RU.im_detect_3d(loader[i][0], net, conf, objs[i])
        ref = HC.refine_detections(aboxes, objs[i].p2, hill_climbing=True, max_out=int(conf.nms_topN_post))
        assert len(got) == len(one) == ref.shape[0], (i, len(got), len(one), ref.shape)
        for (c, v), (c1, v1), r in zip(got, one, ref):
            assert c == c1 == conf.lbls[int(r[0])]
            assert np.allclose(v, v1, rtol=1e-5, atol=2e-5)
            assert np.allclose(v, r[1:], rtol=1e-5, atol=2e-5), (i, np.abs(np.array(v) - r[1:]).max())
        total += len(got)
    assert total > 0  # the score cut leaves something to write


def test_test_kitti_3d_clip_boxes_and_empty(tmp_path):
    """clip_boxes clips before the refinement like im_detect_3d (:1552-1556); an empty loader writes nothing."""
    from m3dssd_b200.lib import rpn_util as RU
    conf, net, frames, objs = _setup(2)
    assert RU.test_kitti_3d([], net, conf, str(tmp_path / "none")) == []
    conf.clip_boxes = True
    RU.test_kitti_3d([(frames[i], objs[i]) for i in range(2)], net, conf, str(tmp_path / "clip"), batch_size=2)
    for i in range(2):
        for _, v in _parse(os.path.join(str(tmp_path / "clip"), "%06d.txt" % i)):
            x1, y1, x2, y2 = v[1:5]
            assert 0 <= x1 <= objs[i].imW - 1 and 0 <= x2 <= objs[i].imW - 1
            assert 0 <= y1 <= objs[i].imH - 1 and 0 <= y2 <= objs[i].imH - 1
