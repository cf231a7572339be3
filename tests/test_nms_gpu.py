"""GPU parity of gpu_nms / batched NMS / decode+top-K: bit-exact keep indices against the C oracle,
the reference's py_cpu_nms golden fixture and (when oracle/_ref was built) the reference's own
unmodified nms_kernel.cu running on the same GPU."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _boxes(n, seed, spread=(1200, 350)):
    rng = np.random.default_rng(seed)
    xy = rng.random((n, 2)) * np.array(spread)
    wh = rng.random((n, 2)) * 120 + 4
    sc = rng.permutation(n).astype(np.float32) / max(n, 1)
    return np.concatenate([xy, xy + wh, sc[:, None]], 1).astype(np.float32)


@pytest.mark.parametrize("n", [0, 1, 2, 63, 64, 65, 127, 500, 3000, 4096, 4097, 9000])
def test_gpu_nms_vs_oracle(n):
    from m3dssd_b200.lib.nms.gpu_nms import gpu_nms
    d = _boxes(n, 100 + n) if n else np.zeros((0, 5), np.float32)
    assert list(gpu_nms(d, 0.4)) == list(O.gpu_nms(d, 0.4))


def test_gpu_nms_vs_reference_py_cpu_nms_golden():
    from m3dssd_b200.lib.nms.gpu_nms import gpu_nms
    g = np.load(os.path.join(GOLD, "nms_py_cpu.npz"))
    for n in (1, 63, 64, 65, 500, 3000):
        d = _boxes(n, int(g["seed_%d" % n]))
        assert list(gpu_nms(d, 0.4)) == list(g["keep_%d" % n])


def test_gpu_nms_dense_overlaps_and_thresholds():
    """Crowded boxes (many IoUs near the threshold) at several thresholds; 14-column rows like aboxes."""
    from m3dssd_b200.lib.nms.gpu_nms import gpu_nms
    for seed, thr in ((1, 0.3), (2, 0.4), (3, 0.5), (4, 0.75)):
        d = _boxes(2000, seed, spread=(300, 120))
        assert list(gpu_nms(d, thr)) == list(O.gpu_nms(d, thr))
        wide = np.concatenate([d, np.zeros((2000, 9), np.float32)], 1)
        assert list(gpu_nms(wide, thr)) == list(O.gpu_nms(d, thr))


def test_gpu_nms_vs_reference_kernel_binary():
    """The reference's unmodified lib/nms/nms_kernel.cu (_nms), built from /root/reference into oracle/_ref."""
    from m3dssd_b200._lib import lib
    path = os.path.join(ROOT, "oracle", "_ref", "libref_nms.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_nms.so not built (reference checkout absent at build time)")
    ref = C.CDLL(path)
    fn = getattr(ref, "_Z4_nmsPiS_PKfiifi")
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int]
    torch.cuda.init()
    for n, seed, spread in ((3000, 7, (1200, 350)), (3000, 8, (300, 120)), (777, 9, (200, 100))):
        d = _boxes(n, seed, spread)
        sd = np.ascontiguousarray(d[d[:, 4].argsort()[::-1]])
        k_ref = np.zeros(n, np.int32)
        n_ref = C.c_int(0)
        fn(k_ref.ctypes.data, C.byref(n_ref), sd.ctypes.data, n, 5, 0.4, 0)
        k_our = np.zeros(n, np.int32)
        n_our = C.c_int(0)
        assert lib().m3d_nms(k_our.ctypes.data, C.byref(n_our), sd.ctypes.data, n, 5, C.c_float(0.4), 0) == 0
        assert n_ref.value == n_our.value
        assert np.array_equal(k_ref[:n_ref.value], k_our[:n_our.value])
        assert np.array_equal(k_ref[:n_ref.value], O.nms_sorted(sd, 0.4))  # pins the oracle too


def test_nms_batched_variable_counts():
    from m3dssd_b200 import ops
    B, max_n = 3, 3000
    nums = [3000, 1234, 0]
    boxes = np.zeros((B, max_n, 14), np.float32)
    for b, n in enumerate(nums):
        if n:
            d = _boxes(n, 40 + b)
            boxes[b, :n, :5] = d[d[:, 4].argsort()[::-1]]
    t = torch.from_numpy(boxes).cuda()
    num = torch.tensor(nums, dtype=torch.int32, device="cuda")
    ws = torch.zeros(ops.nms_workspace_bytes(B, max_n), dtype=torch.uint8, device="cuda")
    keep = torch.zeros(B, max_n, dtype=torch.int32, device="cuda")
    nk = torch.zeros(B, dtype=torch.int32, device="cuda")
    ops.nms_batched(t, num, 0.4, ws, keep, nk)
    for b, n in enumerate(nums):
        exp = O.nms_sorted(boxes[b, :n, :5], 0.4) if n else np.zeros(0, np.int32)
        assert int(nk[b]) == len(exp)
        assert np.array_equal(keep[b, :len(exp)].cpu().numpy(), exp)
    # the bitmask itself (upper triangle) against the oracle's restatement of nms_kernel.cu:61-77
    cb = (max_n + 63) // 64
    mask = ws.view(torch.int64)[:B * max_n * cb].view(B, max_n, cb).cpu().numpy().astype(np.uint64)
    exp = O.nms_mask(boxes[1, :nums[1], :5], 0.4)
    for i in range(0, nums[1], 97):
        rb = i // 64
        assert np.array_equal(mask[1, i, rb:exp.shape[1]], exp[i, rb:])


def test_decode_topk_with_ties():
    """Top-K order = score descending, lower anchor index first on ties (the oracle's stable argsort);
    decode arithmetic vs the oracle's restatement of lib/rpn_util.py:1462-1521."""
    from m3dssd_b200 import ops, synth
    from oracle import ref_model as RM
    conf = synth.make_conf(crop_size=(96, 320))
    A, H, W, K = 36, 12, 40, 4
    M = A * H * W
    g = torch.Generator().manual_seed(3)
    B = 2
    prob = torch.rand(B, M, K, generator=g)
    prob[0, :, 1:] = (prob[0, :, 1:] * 50).round() / 50  # image 0: heavy score ties
    prob = prob / prob.sum(dim=2, keepdim=True)
    b2 = torch.randn(B, M, 4, generator=g) * 0.5
    b3 = torch.randn(B, M, 7, generator=g) * 0.5
    score, cls_pred = prob[..., 1:].max(dim=2)
    cls_pred = (cls_pred + 1).to(torch.uint8)
    conf.bbox_means = np.random.default_rng(0).normal(size=(1, 11)).astype(np.float32) * 0.1
    conf.bbox_stds = (np.random.default_rng(1).random((1, 11)).astype(np.float32) + 0.5)
    m = RM.RefModel({}, conf)
    rois = m.rois(H, W)
    topk = 3000
    dets = torch.zeros(B, topk, 14, device="cuda")
    didx = torch.zeros(B, topk, dtype=torch.int32, device="cuda")
    dnum = torch.zeros(B, dtype=torch.int32, device="cuda")
    ops.decode_topk(score.contiguous().cuda(), cls_pred.contiguous().cuda(), b2.cuda(), b3.cuda(),
                    torch.tensor(conf.anchors).cuda(), conf.bbox_means[0], conf.bbox_stds[0], A, H, W, 8.0, 1.0, topk, dets, didx, dnum)
    for b in range(B):
        pre, keep, kept = m.detect((None, prob, b2, b3, None, rois), b)
        order = torch.argsort(-score[b], stable=True)[:topk]
        assert torch.equal(didx[b].cpu().long(), order)
        assert np.allclose(dets[b].cpu().numpy(), pre.numpy(), rtol=2e-6, atol=1e-4)
        assert int(dnum[b]) == topk
    # the same rows decoded from the NHWC head buffer (what the detection stages do): bit-identical
    slots = [3, 7, 0, 9, 1, 10, 4, 2, 8, 5, 6]  # any permutation: output j lives in slot slots[j]
    heads = torch.zeros(B, H, W, 11 * A + 4)
    flat = torch.cat([b2, b3], dim=2).view(B, A, H, W, 11)  # row = (a*H + h)*W + w
    for j in range(11):
        heads[..., slots[j] * A:(slots[j] + 1) * A] = flat[..., j].permute(0, 2, 3, 1)
    dets2, didx2, dnum2 = torch.zeros_like(dets), torch.zeros_like(didx), torch.zeros_like(dnum)
    ops.decode_topk_heads(score.contiguous().cuda(), cls_pred.contiguous().cuda(), heads.cuda(), slots,
                          torch.tensor(conf.anchors).cuda(), conf.bbox_means[0], conf.bbox_stds[0], A, H, W, 8.0, 1.0, topk,
                          dets2, didx2, dnum2)
    assert torch.equal(dets2, dets) and torch.equal(didx2, didx) and torch.equal(dnum2, dnum)


@pytest.mark.parametrize("seed", [5, 6, 7, 11])
def test_refine_3d_vs_oracle(seed):
    """m3d_refine_3d (post-NMS hill-climb refinement, one thread per box, float64) vs oracle/hill_climb.py and, for
    the golden seeds, vs the unmodified reference functions (tests/golden/hill_climb.npz)."""
    import os
    import torch
    from m3dssd_b200 import ops
    from oracle import hill_climb as HC
    B, max_out = 3, 40
    kept = np.zeros((B, max_out, 14), dtype=np.float32)
    nk = np.zeros(B, dtype=np.int32)
    p2s, refs = [], []
    for b in range(B):
        rows, p2 = HC.synthetic_detections(64, seed + 100 * b)
        if b == 1:  # a different camera and a short list
            p2 = p2.copy()
            p2[0, 0] *= 1.1
            p2[1, 1] *= 1.1
            rows = rows[:17]
        if b == 2:  # boxes behind the camera / tiny depth: the `invalid` branch
            rows = rows.copy()
            rows[::5, 8] = 0.3
        n = min(max_out, rows.shape[0])
        kept[b, :n] = rows[:n]
        nk[b] = n
        p2s.append(p2)
        refs.append(HC.refine_detections(rows[:n], p2, max_out=max_out))
    out, valid = ops.refine_3d(torch.from_numpy(kept).cuda(), torch.from_numpy(nk).cuda(), np.stack(p2s))
    out, valid = out.cpu().numpy(), valid.cpu().numpy().astype(bool)
    for b in range(B):
        exp_valid = (np.arange(max_out) < nk[b]) & (kept[b, :, 4] >= 0.75)
        assert np.array_equal(valid[b], exp_valid)
        got = out[b][valid[b]]
        assert got.shape == refs[b].shape
        assert np.allclose(got, refs[b], rtol=1e-9, atol=1e-9), np.abs(got - refs[b]).max()
        assert np.all(out[b][~valid[b]] == 0)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hill_climb.npz"))
    if "refined_%d" % seed in gold:
        assert np.abs(out[0][valid[0]] - gold["refined_%d" % seed]).max() < 2e-5
    # no hill climbing: only the alpha <-> rotation round trip and the back-projection
    out2, valid2 = ops.refine_3d(torch.from_numpy(kept).cuda(), torch.from_numpy(nk).cuda(), np.stack(p2s),
                                 hill_climbing=False)
    ref2 = HC.refine_detections(kept[0, :nk[0]], p2s[0], hill_climbing=False, max_out=max_out)
    assert np.allclose(out2.cpu().numpy()[0][valid2.cpu().numpy()[0].astype(bool)], ref2, rtol=1e-9, atol=1e-9)


def test_batched_nms_more_than_4096_boxes_per_image():
    """m3d_nms_batched above the fast sweep's 4096-box limit (generic exact sweep): bit-exact keep lists per image,
    ragged box counts."""
    from m3dssd_b200 import ops
    B, max_n = 2, 5000
    nums = [5000, 4321]
    boxes = torch.zeros(B, max_n, 5)
    exp = []
    for b in range(B):
        d = _boxes(nums[b], 40 + b, spread=(900, 300))
        sd = np.ascontiguousarray(d[d[:, 4].argsort()[::-1]])
        boxes[b, :nums[b]] = torch.from_numpy(sd)
        exp.append(O.nms_sorted(sd, 0.4))
    ws = torch.zeros(ops.nms_workspace_bytes(B, max_n), dtype=torch.uint8, device="cuda")
    keep = torch.zeros(B, max_n, dtype=torch.int32, device="cuda")
    nk = torch.zeros(B, dtype=torch.int32, device="cuda")
    ops.nms_batched(boxes.cuda(), torch.tensor(nums, dtype=torch.int32, device="cuda"), 0.4, ws, keep, nk)
    torch.cuda.synchronize()
    for b in range(B):
        n = int(nk[b])
        assert n == len(exp[b]) and np.array_equal(keep[b, :n].cpu().numpy(), np.asarray(exp[b]))
