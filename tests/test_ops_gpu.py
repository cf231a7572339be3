"""GPU parity of the bandwidth-bound kernels against torch CPU fp32 (the reference's own
third-party arithmetic for these layers) on seeded inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _nhwc(x, dtype=torch.float32):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()


def _nchw(x):
    return x.float().cpu().permute(0, 3, 1, 2).contiguous()


def _g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_stem_conv7x7(dtype):
    from m3dssd_b200 import ops
    g = _g(1)
    x = torch.randn(2, 3, 37, 70, generator=g)
    w = torch.randn(16, 3, 7, 7, generator=g) * 0.1
    b = torch.randn(16, generator=g)
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=3), 0.01)
    out = torch.zeros(2, 37, 70, 64 if dtype == torch.float32 else 16, dtype=dtype, device="cuda")
    ops.stem_conv7x7(x.cuda(), w.cuda(), b.cuda(), out, 0.01)
    got = _nchw(out)[:, :16]
    tol = 1e-4 if dtype == torch.float32 else 2 ** -7 * ref.abs().max().item()
    assert (got - ref).abs().max().item() < tol


@pytest.mark.parametrize("shape", [(2, 36, 72), (1, 96, 320), (3, 50, 44)])
@pytest.mark.parametrize("legacy", [False, True])
def test_stem_conv7x7_s2d(shape, legacy, monkeypatch):
    """bf16 trunk stem in 2x2 space-to-depth form (dedicated kernel, and the shared gather kernel it replaced) vs
    conv2d on the bf16-rounded operands: output channel (ey*2+ex)*16 + c of pixel (Y, X) = channel c at (2Y+ey, 2X+ex)."""
    from m3dssd_b200 import ops
    if legacy:
        monkeypatch.setenv("M3D_STEM_LEGACY", "1")
    else:
        monkeypatch.delenv("M3D_STEM_LEGACY", raising=False)
    N, H, W = shape
    g = _g(21)
    x = torch.randn(N, 3, H, W, generator=g)
    w = torch.randn(16, 3, 7, 7, generator=g) * 0.1
    b = torch.randn(16, generator=g)
    wp, bp = ops.pack_stem_s2d(w, b)
    ref = F.leaky_relu(F.conv2d(x.bfloat16().float(), w.bfloat16().float(), b, padding=3), 0.01)
    out = torch.full((N, H // 2, W // 2, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.stem_conv7x7_s2d(x.cuda(), wp.cuda(), bp.cuda(), out, 0.01)
    got = out.float().cpu().reshape(N, H // 2, W // 2, 2, 2, 16).permute(0, 5, 1, 3, 2, 4).reshape(N, 16, H, W)
    assert (got - ref).abs().max().item() < 2 ** -7 * ref.abs().max().item()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_maxpool2x2(dtype):
    from m3dssd_b200 import ops
    x = torch.randn(2, 32, 12, 20, generator=_g(2))
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    ref = F.max_pool2d(x, 2, 2)
    out = torch.zeros(2, 6, 10, 32, dtype=dtype, device="cuda")
    ops.maxpool2x2(_nhwc(x, dtype), out)
    assert torch.equal(_nchw(out), ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("f", [2, 4])
def test_upsample_add(dtype, f):
    from m3dssd_b200 import ops
    g = _g(3)
    C = 64
    x = torch.randn(2, C, 6, 10, generator=g)
    skip = torch.randn(2, C, 6 * f, 10 * f, generator=g)
    w = torch.rand(C, 1, 2 * f, 2 * f, generator=g)
    if dtype == torch.bfloat16:
        x, skip = x.bfloat16().float(), skip.bfloat16().float()
    ref = F.conv_transpose2d(x, w, None, stride=f, padding=f // 2, groups=C) + skip
    out = torch.zeros(2, 6 * f, 10 * f, C, dtype=dtype, device="cuda")
    ops.upsample_add(_nhwc(x, dtype), ops.pack_upsample_weight(w).cuda(), _nhwc(skip, dtype), out, f)
    tol = 1e-5 if dtype == torch.float32 else 2 ** -7 * ref.abs().max().item()
    assert (_nchw(out) - ref).abs().max().item() < tol


def test_cls_softmax_and_flatten():
    from m3dssd_b200 import ops
    g = _g(4)
    B, A, K, H, W = 2, 36, 4, 5, 41
    logits = torch.randn(B, K * A, H, W, generator=g) * 2
    cls = logits.view(B, K, H * A, W)
    prob = torch.softmax(cls, dim=1)
    fg = (1 - prob[:, 0]).view(B, A, H, W)

    def flat(t):
        return t.permute(0, 2, 3, 1).contiguous().view(B, -1, t.shape[1])

    M = A * H * W
    f32 = dict(dtype=torch.float32, device="cuda")
    cls_o, prob_o = torch.zeros(B, M, K, **f32), torch.zeros(B, M, K, **f32)
    fg_max, fg_arg = torch.zeros(B, H, W, **f32), torch.zeros(B, H, W, dtype=torch.int32, device="cuda")
    score, cp = torch.zeros(B, M, **f32), torch.zeros(B, M, dtype=torch.uint8, device="cuda")
    ops.cls_softmax(_nhwc(logits), A, K, cls_o, prob_o, fg_max, fg_arg, score, cp)
    assert torch.equal(cls_o.cpu(), flat(cls))
    assert (prob_o.cpu() - flat(prob)).abs().max().item() < 1e-6
    mx, arg = fg.max(dim=1)
    assert (fg_max.cpu() - mx).abs().max().item() < 1e-6
    assert (fg_arg.cpu().long() == arg).float().mean().item() > 0.999
    p = flat(prob)
    assert (score.cpu() - p[..., 1:].max(dim=2)[0]).abs().max().item() < 1e-6
    assert (cp.cpu().long() == p[..., 1:].argmax(dim=2) + 1).float().mean().item() > 0.999
    # heads flatten
    from m3dssd_b200.engine import HEAD_ORDER, OUT_SLOTS
    names = ["bbox_x", "bbox_y", "bbox_w", "bbox_h", "bbox_x3d", "bbox_y3d", "bbox_z3d", "bbox_w3d", "bbox_h3d",
             "bbox_l3d", "bbox_rY3d"]
    assert [HEAD_ORDER[s] for s in OUT_SLOTS] == names
    heads = {n: torch.randn(B, A, H, W, generator=g) for n in names}
    buf = torch.cat([heads[n] for n in HEAD_ORDER], dim=1)
    b2, b3 = torch.zeros(B, M, 4, **f32), torch.zeros(B, M, 7, **f32)
    ops.flatten_heads(_nhwc(buf), A, OUT_SLOTS, b2, b3)
    ref2 = torch.cat([flat(heads[n].reshape(B, 1, H * A, W)) for n in names[:4]], dim=2)
    ref3 = torch.cat([flat(heads[n].reshape(B, 1, H * A, W)) for n in names[4:]], dim=2)
    assert torch.equal(b2.cpu(), ref2) and torch.equal(b3.cpu(), ref3)


def test_align_offset_builders():
    from m3dssd_b200 import ops, synth
    g = _g(5)
    conf = synth.make_conf()
    anchors = torch.tensor(conf.anchors)
    B, A, H, W = 2, 36, 6, 9
    fg = torch.rand(B, A, H, W, generator=g)
    mask, ind = fg.max(dim=1, keepdim=True)
    hard = (mask > 0.5).float()
    f32 = dict(dtype=torch.float32, device="cuda")
    fg_max, fg_arg = mask[:, 0].contiguous().cuda(), ind[:, 0].int().contiguous().cuda()
    # shape align (feturealign_mgpu.py:119-136, 160-172)
    ah = (anchors[:, 3] - anchors[:, 1]) / 8 / 3
    aw = (anchors[:, 2] - anchors[:, 0]) / 8 / 3
    offs = []
    for i in range(3):
        for j in range(3):
            offs += [(ah[ind] - 1) * (i - 1.5 + 0.5), (aw[ind] - 1) * (j - 1.5 + 0.5)]
    ref = torch.cat([torch.cat(offs, dim=1) * hard, mask.repeat(1, 9, 1, 1)], dim=1)
    om = torch.zeros(B, H, W, 27, **f32)
    ops.shape_align_om(fg_max, fg_arg, anchors.cuda(), 8.0, 0.5, om)
    assert (_nchw(om) - ref).abs().max().item() < 1e-5
    # center align (feturealign_mgpu.py:58-77)
    heads = torch.randn(B, 11 * A, H, W, generator=g)
    bx, by = heads[:, 0:A], heads[:, A:2 * A]
    mean, std = [0.1, -0.2], [1.5, 0.7]
    ox = torch.gather((bx * std[0] + mean[0]) * ((anchors[:, 2] - anchors[:, 0]) / 8).view(1, -1, 1, 1), 1, ind) * hard
    oy = torch.gather((by * std[1] + mean[1]) * ((anchors[:, 3] - anchors[:, 1]) / 8).view(1, -1, 1, 1), 1, ind) * hard
    ref = torch.cat([oy, ox, mask], dim=1)
    om = torch.zeros(B, H, W, 4, **f32)
    ops.center_align_om(fg_max, fg_arg, _nhwc(heads), 0, A, anchors.cuda(), 8.0, mean, std, 0.5, om)
    assert (_nchw(om)[:, :3] - ref).abs().max().item() < 1e-5
    # the two-in-one launch of the detection stages: same bits as two calls
    mean3, std3 = [0.3, 0.05], [0.9, 1.1]
    om_b = torch.zeros(B, H, W, 4, **f32)
    ops.center_align_om(fg_max, fg_arg, _nhwc(heads), 2 * A, 3 * A, anchors.cuda(), 8.0, mean3, std3, 0.5, om_b)
    o2a, o2b = torch.zeros_like(om), torch.zeros_like(om)
    ops.center_align_om2(fg_max, fg_arg, _nhwc(heads), (0, A, 2 * A, 3 * A), anchors.cuda(), 8.0, mean + mean3, std + std3,
                         0.5, o2a, o2b)
    assert torch.equal(o2a, om) and torch.equal(o2b, om_b)


def test_softmax_detect_variants():
    """The detection stages' softmax (no flattened cls / prob copies, shape_align offsets built in its tail) writes the
    same score / class / fg / om bits as cls_softmax + shape_align_om."""
    from m3dssd_b200 import ops, synth
    g = _g(14)
    B, A, K, H, W = 2, 36, 4, 5, 41
    logits = _nhwc(torch.randn(B, K * A, H, W, generator=g) * 2)
    anchors = torch.tensor(synth.make_conf().anchors).cuda()
    M = A * H * W
    f32 = dict(dtype=torch.float32, device="cuda")

    def bufs():
        return (torch.zeros(B, H, W, **f32), torch.zeros(B, H, W, dtype=torch.int32, device="cuda"), torch.zeros(B, M, **f32),
                torch.zeros(B, M, dtype=torch.uint8, device="cuda"))
    fm, fa, sc, cp = bufs()
    cls_o, prob_o = torch.zeros(B, M, K, **f32), torch.zeros(B, M, K, **f32)
    ops.cls_softmax(logits, A, K, cls_o, prob_o, fm, fa, sc, cp)
    om = torch.zeros(B, H, W, 27, **f32)
    ops.shape_align_om(fm, fa, anchors, 8.0, 0.5, om)
    for fused in (False, True):
        fm2, fa2, sc2, cp2 = bufs()
        om2 = torch.zeros_like(om)
        if fused:
            ops.cls_softmax_shape_om(logits, A, K, None, None, fm2, fa2, sc2, cp2, anchors, 8.0, 0.5, om2)
            assert torch.equal(om2, om)
        else:
            ops.cls_softmax(logits, A, K, None, None, fm2, fa2, sc2, cp2)
        assert torch.equal(fm2, fm) and torch.equal(fa2, fa) and torch.equal(sc2, sc) and torch.equal(cp2, cp)


def test_layout_round_trip():
    from m3dssd_b200 import ops
    x = torch.randn(2, 37, 5, 9, generator=_g(6)).cuda()
    nhwc = torch.zeros(2, 5, 9, 64, dtype=torch.float32, device="cuda")
    ops.nchw_to_nhwc(x, nhwc)
    assert torch.equal(nhwc[..., :37].permute(0, 3, 1, 2), x)
    back = torch.zeros_like(x)
    ops.nhwc_to_nchw(nhwc, back)
    assert torch.equal(back, x)


@pytest.mark.parametrize("hw", [(12, 40), (48, 160)])  # uneven adaptive bins / the exact-bin fast pooling path
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_anab_pool_and_attention(dtype, hw):
    from m3dssd_b200 import ops
    g = _g(7)
    B, C, ck, cv = 2, 128, 168, 128
    H, W = hw
    sizes = [1, 4, 8, 16]
    T = sum(s * s for s in sizes)
    x = torch.randn(B, C, H, W, generator=g)
    q = torch.randn(B, ck, H, W, generator=g) * 0.3
    k = torch.randn(B, ck, H, W, generator=g) * 0.3
    v = torch.randn(B, cv, H, W, generator=g)
    s = torch.randn(B, 4, H, W, generator=g)
    if dtype == torch.bfloat16:
        x, q = x.bfloat16().float(), q.bfloat16().float()
    att = torch.sigmoid(s)

    def papa(f):
        return torch.cat([F.adaptive_avg_pool2d(f * att[:, i:i + 1], (sz, sz)).view(B, f.shape[1], -1)
                          for i, sz in enumerate(sizes)], -1)

    key, val = papa(k), papa(v).permute(0, 2, 1)
    a = torch.softmax(torch.bmm(q.view(B, ck, H * W).permute(0, 2, 1), key), dim=-1)
    new = torch.bmm(a, val).permute(0, 2, 1).reshape(B, cv, H, W) + x
    scale, shift = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    ref = F.leaky_relu(new * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), 0.01)

    kvs = _nhwc(torch.cat([k, v, s], dim=1))
    ktok = torch.zeros(B, T, ck, device="cuda")
    vtok = torch.zeros(B, T, cv, device="cuda")
    ops.anab_pool(kvs, ck, cv, sizes, ktok, vtok)
    assert (ktok.cpu() - key.permute(0, 2, 1)).abs().max().item() < 1e-5
    assert (vtok.cpu() - val).abs().max().item() < 1e-5
    out = torch.zeros(B, H, W, C, dtype=dtype, device="cuda")
    ops.anab_attention(_nhwc(q, dtype), ktok, vtok, _nhwc(x, dtype), scale.cuda(), shift.cuda(), 0.01, out, ck, cv)
    tol = 2e-5 * ref.abs().max().item() if dtype == torch.float32 else 2 ** -7 * ref.abs().max().item()
    assert (_nchw(out) - ref).abs().max().item() < tol


@pytest.mark.parametrize("shape", [(2, 12, 20, 128, 3), (1, 9, 37, 128, 1), (2, 48, 160, 128, 4), (1, 8, 16, 64, 2)])
def test_head_mlp(shape):
    """Fused three-layer 1x1 heads vs the layer-by-layer fp32 evaluation with the same bf16 roundings
    (inputs / weights bf16, fp32 accumulation, intermediates rounded to bf16 after bias + LeakyReLU)."""
    from m3dssd_b200 import ops
    N, H, W, Cx, G = shape
    A, rows3 = 36, 48
    g = _g(11)
    x = torch.randn(N, H, W, Cx + 64, generator=g).bfloat16()  # a channel window of a wider buffer
    xoff = 64
    w1 = (torch.randn(G * 256, Cx, generator=g) / Cx ** 0.5).bfloat16()
    w2 = (torch.randn(G * 256, 256, generator=g) / 16).bfloat16()
    w3 = torch.zeros(G * rows3, 256)
    b3 = torch.zeros(G * rows3)
    for k in range(G):
        w3[k * rows3:k * rows3 + A] = torch.randn(A, 256, generator=g) / 16
        b3[k * rows3:k * rows3 + A] = torch.randn(A, generator=g)
    w3 = w3.bfloat16()
    b1 = torch.randn(G * 256, generator=g)
    b2 = torch.randn(G * 256, generator=g)
    out = torch.full((N, H, W, 11 * A), 7.0, device="cuda")
    coff = 2 * A
    ops.head_mlp(x.cuda(), xoff, Cx, w1.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), w3.cuda(), b3.cuda(), G, A, rows3, out,
                 coff, 0.01)
    got = out.cpu()
    xs = x[..., xoff:xoff + Cx].float().reshape(-1, Cx).double()
    for k in range(G):
        h1 = F.leaky_relu(xs @ w1[k * 256:(k + 1) * 256].double().t() + b1[k * 256:(k + 1) * 256].double(), 0.01)
        h1 = h1.float().bfloat16().double()
        h2 = F.leaky_relu(h1 @ w2[k * 256:(k + 1) * 256].double().t() + b2[k * 256:(k + 1) * 256].double(), 0.01)
        h2 = h2.float().bfloat16().double()
        ref = (h2 @ w3[k * rows3:k * rows3 + A].double().t() + b3[k * rows3:k * rows3 + A].double()).float()
        o = got[..., coff + k * A:coff + (k + 1) * A].reshape(-1, A)
        # a bf16 rounding of an intermediate can flip on an fp32-accumulation-order difference: 1 bf16 ulp of one
        # of 256 terms, i.e. ~2^-9 * |h| * |w| ~ 1e-3 absolute; typical error is ~1e-5
        assert (o - ref).abs().max().item() < 5e-3 * max(1.0, ref.abs().max().item()), k
        assert (o - ref).abs().mean().item() < 2e-4
    # untouched channels stay untouched
    assert torch.all(got[..., :coff] == 7.0) and torch.all(got[..., coff + G * A:] == 7.0)


@pytest.mark.parametrize("shape", [(1, 24, 40), (2, 96, 320), (8, 384, 1280)])
def test_preprocess_u8_bit_exact_vs_oracle(shape):
    """Device-side input pipeline (m3d_preprocess_u8) == the reference's Normalize + BGR->RGB + CHW: bit-exact."""
    from m3dssd_b200 import ops, synth
    from oracle import oracle as O
    N, H, W = shape
    im = synth.make_images_u8(N, (H, W), seed=N)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    out = torch.empty(N, 3, H, W, dtype=torch.float32, device="cuda")
    ops.preprocess_u8(im.cuda(), out, mean, std, swap_rb=True)
    ref = O.preprocess_u8(im.numpy(), mean, std)
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("size,n", [((48, 64), 6), ((384, 1280), 8), ((96, 320), 70)])
def test_preprocess_u8_pad_ragged_bit_exact(size, n):
    """Ragged frames -> zero Padding + Normalize + BGR->RGB + CHW on the device (m3d_preprocess_u8_pad) == the
    reference's Preprocess: bit-exact against the committed golden (unmodified reference + cv2) and the oracle,
    incl. full-size, 1x1 and empty frames, > 64 frames (two launches), and the error for an over-sized frame."""
    import os
    from m3dssd_b200 import ops
    from m3dssd_b200._lib import M3DError
    from oracle import oracle as O
    H, W = size
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    if size == (48, 64):
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess_pad_u8.npz"))
        ims = [g["image_%d" % k] for k in range(int(g["n"]))]
        ref = np.stack([g["expected_%d" % k] for k in range(int(g["n"]))])
    else:
        rng = np.random.default_rng(n)
        hs = rng.integers(H - 14, H + 1, n)  # KITTI: 370-376 x 1224-1242 under 384 x 1280
        ws = rng.integers(W - 56, W + 1, n)
        hs[0], ws[0] = H, W
        hs[1], ws[1] = 1, 1
        hs[2], ws[2] = 0, 0
        ims = [rng.integers(0, 256, (int(h), int(w), 3), dtype=np.uint8) for h, w in zip(hs, ws)]
        ref = O.preprocess_pad_u8(ims, size, mean, std)
    buf, off, hh, ww = ops.pack_ragged_u8(ims)
    out = torch.full((len(ims), 3, H, W), float("nan"), dtype=torch.float32, device="cuda")
    ops.preprocess_u8_pad(buf.cuda(), off, hh, ww, out, mean, std, swap_rb=True)
    assert np.array_equal(out.cpu().numpy(), ref)
    with pytest.raises(M3DError):
        ops.preprocess_u8_pad(buf.cuda(), off, [H + 1] + hh[1:], ww, out, mean, std)


def test_detect_images_ragged_equals_padded_batch():
    """RPN.detect_images(list of ragged uint8 frames) == RPN.engine.detect on the zero-padded uint8 batch... except the
    padding: the reference pads BEFORE Normalize, so the padded uint8 batch (0 bytes) gives the same tensor."""
    from m3dssd_b200 import synth
    from m3dssd_b200.model.M3d_inference_align import build
    conf = synth.make_conf(crop_size=(96, 320))
    net = build(conf, "test")
    synth.randomize_weights(net)
    net = net.cuda().eval()
    rng = np.random.default_rng(3)
    full = synth.make_images_u8(2, (96, 320), seed=5).numpy()
    ims = [full[0][:90, :301].copy(), full[1][:96, :316].copy()]
    padded = np.zeros_like(full)
    padded[0, :90, :301] = ims[0]
    padded[1, :96, :316] = ims[1]
    kept_a, num_a = net.detect_images(ims)
    eng = net.engine(2, 96, 320)
    kept_b, num_b = eng.detect(torch.from_numpy(padded).cuda())
    assert torch.equal(num_a, num_b) and torch.equal(kept_a, kept_b)
    assert int(num_a.sum()) > 0
