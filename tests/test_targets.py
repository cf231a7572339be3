"""Training-target assignment: oracle/targets.py against the fixture the UNMODIFIED reference compute_targets produced
(tests/golden/make_golden_targets.py), and m3d_compute_targets (csrc/targets.cu, through
m3dssd_b200.lib.targets.compute_targets_batch) against both."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_targets as G  # noqa: E402  (case definitions; imports no reference code at import time)
from oracle import targets as OT  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "targets.npz"))


def _rois(conf, feat):
    from m3dssd_b200.lib.rpn_util import locate_anchors
    return locate_anchors(conf.anchors, feat, conf.feat_stride).astype(np.float32)


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_matches_reference_golden(name):
    conf, _ = G.case_conf(name)
    rois = _rois(conf, G.CASES[name]["feat"])
    for i, g in enumerate(G.case_gts(name)):
        t = OT.compute_targets(g["gts_val"], g["gts_ign"], g["box_lbls"], rois, conf.fg_thresh, conf.ign_thresh, conf.bg_thresh_lo,
                               conf.bg_thresh_hi, conf.best_thresh, g["gts_3d"], np.asarray(conf.anchors))
        assert np.array_equal(t[:, 4].astype(np.int8), GOLD["%s.%d.code" % (name, i)])
        fg = GOLD["%s.%d.fg_rows" % (name, i)]
        assert np.array_equal(t[fg], GOLD["%s.%d.fg_transforms" % (name, i)])


def _ulp_close(a, b, ulps=2):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.all(np.abs(a - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32)))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.CASES))
def test_compute_targets_kernel_vs_reference_golden(name):
    """Labels bit for bit (every anchor of every image), regression targets to the last float32 bit or two (the only
    inexact step is log(): CUDA's and numpy's float64 logarithms may differ in the last place before the rounding)."""
    from m3dssd_b200.lib.targets import compute_targets_batch
    conf, _ = G.case_conf(name)
    feat = G.CASES[name]["feat"]
    gts = G.case_gts(name)
    gts.append({"gts_val": np.zeros((0, 4)), "gts_ign": gts[0]["gts_ign"], "box_lbls": np.zeros(0, int), "gts_3d": np.zeros((0, 7))})
    out = compute_targets_batch(conf, gts, feat)
    rois = _rois(conf, feat)
    means, stds = np.asarray(conf.bbox_means), np.asarray(conf.bbox_stds)
    assert torch.equal(out["meta"]["rois"][0].cpu(), torch.from_numpy(rois))
    for i, g in enumerate(gts):
        fg, bg, ign, labels, b2, b3, any_val = OT.targets_image(g, rois, conf)
        if i < len(gts) - 1:  # the golden pins the oracle's transforms for this image
            code = GOLD["%s.%d.code" % (name, i)]
            assert np.array_equal(fg, code > 0) and np.array_equal(bg, code < 0) and np.array_equal(ign, code == 0)
            rows = GOLD["%s.%d.fg_rows" % (name, i)]
            t = GOLD["%s.%d.fg_transforms" % (name, i)]
            assert np.array_equal(b2[rows], (t[:, 0:4] - means[:, 0:4]) / stds[:, 0:4])
        assert np.array_equal(out["labels_fg"][i].cpu().numpy(), fg)
        assert np.array_equal(out["labels_bg"][i].cpu().numpy(), bg)
        assert np.array_equal(out["labels_ign"][i].cpu().numpy(), ign)
        assert np.array_equal(out["labels"][i].cpu().numpy(), labels)
        assert bool(out["meta"]["any_val"][i]) == bool(any_val)
        assert _ulp_close(out["bbox_2d"][i].cpu().numpy(), b2) and _ulp_close(out["bbox_3d"][i].cpu().numpy(), b3)
        exact = np.mean(out["bbox_3d"][i].cpu().numpy() == b3)
        assert exact > 0.999, exact


@pytest.mark.gpu
def test_targets_feed_the_loss():
    """The device-built targets go straight into RPN_3D_loss_smp: finite loss with all terms present."""
    from m3dssd_b200.lib.loss.rpn_3d import RPN_3D_loss_smp
    from m3dssd_b200.lib.targets import compute_targets_batch
    conf, _ = G.case_conf("small")
    feat = G.CASES["small"]["feat"]
    tar = compute_targets_batch(conf, G.case_gts("small"), feat)
    B, M = tar["labels"].shape
    g = torch.Generator().manual_seed(0)
    cls = torch.randn(B, M, 4, generator=g).cuda().requires_grad_(True)
    b2 = (tar["bbox_2d"] + 0.05 * torch.randn(B, M, 4, generator=g).cuda()).requires_grad_(True)
    b3 = (torch.randn(B, M, 7, generator=g) * 0.3).cuda().requires_grad_(True)
    loss, stats = RPN_3D_loss_smp(conf).cuda()(cls, torch.softmax(cls, 2), b2, b3, tar)
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(cls.grad).all() and torch.isfinite(b2.grad).all()
    assert {s["name"] for s in stats} >= {"cls", "bbox3d", "iou", "z", "ry"}
