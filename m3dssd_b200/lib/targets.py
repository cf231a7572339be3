"""Training targets for a batch on the device: the reference's compute_targets (lib/rpn_util.py:430-532) and the
per-image part of Dataset._targets (lib/dataloader.py:1014-1144) in one C-ABI call (csrc/targets.cu).

The reference builds them with numpy on the data-loader workers and ships five [M]-sized arrays per image to the GPU;
here only the ground-truth boxes travel (a few hundred bytes per image) and the result is the `imobjs` dict
RPN_3D_loss_smp takes (m3dssd_b200/lib/loss/rpn_3d.py), already on the device.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .rpn_util import locate_anchors

IGN_FLAG = 3000
_rois_cache = {}


def _rois(conf, feat_size, device):
    key = (hash(np.asarray(conf.anchors).tobytes()), tuple(int(v) for v in feat_size), float(conf.feat_stride), str(device))
    if key not in _rois_cache:
        r = torch.from_numpy(locate_anchors(conf.anchors, feat_size, conf.feat_stride).astype(np.float32))
        _rois_cache[key] = r.to(device)
    return _rois_cache[key]


def compute_targets_batch(conf, gts, feat_size, device="cuda", p2=None):
    """gts: one dict per image with numpy arrays gts_val [G,4] (x1,y1,x2,y2), gts_ign [Gi,4], box_lbls [G] (class index
    >= 1, clsName2Ind of the reference) and gts_3d [G,7] (cx, cy, z, w3d, h3d, l3d, rotY) -- what Dataset._targets derives
    from the image's annotations before it calls compute_targets.  conf: fg_thresh, ign_thresh, bg_thresh_lo/hi,
    best_thresh, anchors, bbox_means / bbox_stds, feat_stride.  Returns the reference's target dict on `device`."""
    if not torch.cuda.is_available():
        raise NotImplementedError("compute_targets_batch needs a CUDA device (no CPU fallback)")
    L = _lib.lib()
    B = len(gts)
    A = conf.anchors.shape[0]
    H, W = int(feat_size[0]), int(feat_size[1])
    M = A * H * W
    gmax = max([len(g["gts_val"]) for g in gts] + [0])
    imax = max([len(g["gts_ign"]) for g in gts] + [0])
    val = np.zeros((B, max(gmax, 1), 4), dtype=np.float64)
    g3d = np.zeros((B, max(gmax, 1), 7), dtype=np.float64)
    lbl = np.zeros((B, max(gmax, 1)), dtype=np.int32)
    ign = np.zeros((B, max(imax, 1), 4), dtype=np.float64)
    nv = np.zeros(B, dtype=np.int32)
    ni = np.zeros(B, dtype=np.int32)
    for b, g in enumerate(gts):
        n, k = len(g["gts_val"]), len(g["gts_ign"])
        nv[b], ni[b] = n, k
        if n:
            val[b, :n] = np.asarray(g["gts_val"], dtype=np.float64)[:, :4]
            g3d[b, :n] = np.asarray(g["gts_3d"], dtype=np.float64)[:, :7]
            lbl[b, :n] = np.asarray(g["box_lbls"], dtype=np.int32)
            assert (lbl[b, :n] >= 1).all(), "box_lbls are class indices >= 1 (lib/rpn_util.py:504)"
        if k:
            ign[b, :k] = np.asarray(g["gts_ign"], dtype=np.float64)[:, :4]
    dv = {k: torch.from_numpy(v).to(device) for k, v in dict(val=val, g3d=g3d, lbl=lbl, ign=ign, nv=nv, ni=ni).items()}
    anchors = torch.as_tensor(np.asarray(conf.anchors), dtype=torch.float32).contiguous().to(device)
    u8 = dict(dtype=torch.uint8, device=device)
    out = {
        "labels_fg": torch.empty(B, M, **u8), "labels_bg": torch.empty(B, M, **u8), "labels_ign": torch.empty(B, M, **u8),
        "labels": torch.empty(B, M, dtype=torch.int64, device=device),
        "bbox_2d": torch.empty(B, M, 4, dtype=torch.float32, device=device),
        "bbox_3d": torch.empty(B, M, 7, dtype=torch.float32, device=device),
    }
    any_val = torch.empty(B, **u8)
    ws = torch.empty(L.m3d_compute_targets_workspace(B, gmax), dtype=torch.uint8, device=device)
    means = (C.c_float * 11)(*[float(v) for v in np.asarray(conf.bbox_means).reshape(-1)[:11]])
    stds = (C.c_float * 11)(*[float(v) for v in np.asarray(conf.bbox_stds).reshape(-1)[:11]])
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    _lib.check(L.m3d_compute_targets(
        p(dv["val"]), p(dv["g3d"]), p(dv["lbl"]), p(dv["nv"]), gmax, p(dv["ign"]), p(dv["ni"]), imax, p(anchors), B, A, H, W,
        float(conf.feat_stride), float(conf.fg_thresh), float(conf.ign_thresh), float(conf.bg_thresh_lo),
        float(conf.bg_thresh_hi), float(conf.best_thresh), means, stds, p(out["labels_fg"]), p(out["labels_bg"]),
        p(out["labels_ign"]), p(out["labels"]), p(out["bbox_2d"]), p(out["bbox_3d"]), p(any_val), p(ws), ws.numel(),
        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    for k in ("labels_fg", "labels_bg", "labels_ign"):
        out[k] = out[k].bool()
    rois = _rois(conf, (H, W), device)
    out["meta"] = {"rois": rois.unsqueeze(0).expand(B, M, 5), "any_val": any_val.bool(),
                   "p2": None if p2 is None else torch.as_tensor(p2).to(device)}
    if out["meta"]["p2"] is None:
        del out["meta"]["p2"]
    return out
