"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's training-target assignment, used to check
m3d_compute_targets (csrc/targets.cu).  Never imported by the product.

Follows lib/rpn_util.py:430-532 (compute_targets), lib/core.py:249-300, 341-372, 402-430 (intersect / iou / iou_ign, numpy
branches), lib/rpn_util.py:1059-1134 (bbox_transform_3d, bbox_transform) and the per-image post-processing of
lib/dataloader.py:1086-1125 (Dataset._targets).  Pinned by tests/golden/targets.npz, which the UNMODIFIED reference
functions produced (tests/golden/make_golden_targets.py).
"""
import numpy as np

IGN_FLAG = 3000


def _intersect(a, b):
    max_xy = np.minimum(a[:, 2:4], np.expand_dims(b[:, 2:4], axis=1))
    min_xy = np.maximum(a[:, 0:2], np.expand_dims(b[:, 0:2], axis=1))
    inter = np.clip(max_xy - min_xy, a_min=0, a_max=None)
    return inter[:, :, 0] * inter[:, :, 1]


def iou(a, b):
    inter = _intersect(a, b)
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return (inter / (np.expand_dims(area_a, 0) + np.expand_dims(area_b, 1) - inter)).T


def iou_ign(a, b):
    inter = _intersect(a, b)
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return (inter / (np.expand_dims(area_a, 0) + np.expand_dims(area_b, 1) * 0 - inter * 0)).T


def bbox_transform(ex, gt):
    ew = ex[:, 2] - ex[:, 0] + 1.0
    eh = ex[:, 3] - ex[:, 1] + 1.0
    ecx = ex[:, 0] + 0.5 * (ew - 1)
    ecy = ex[:, 1] + 0.5 * (eh - 1)
    gw = gt[:, 2] - gt[:, 0] + 1.0
    gh = gt[:, 3] - gt[:, 1] + 1.0
    gcx = gt[:, 0] + 0.5 * (gw - 1.0)
    gcy = gt[:, 1] + 0.5 * (gh - 1.0)
    return np.vstack(((gcx - ecx) / ew, (gcy - ecy) / eh, np.log(gw / ew), np.log(gh / eh))).transpose()


def bbox_transform_3d(ex2d, ex3d, gt):
    ew = ex2d[:, 2] - ex2d[:, 0] + 1.0
    eh = ex2d[:, 3] - ex2d[:, 1] + 1.0
    ecx = ex2d[:, 0] + 0.5 * (ew - 1)
    ecy = ex2d[:, 1] + 0.5 * (eh - 1)
    return np.vstack(((gt[:, 0] - ecx) / ew, (gt[:, 1] - ecy) / eh, gt[:, 2] - ex3d[:, 0], np.log(gt[:, 3] / ex3d[:, 1]),
                      np.log(gt[:, 4] / ex3d[:, 2]), np.log(gt[:, 5] / ex3d[:, 3]), gt[:, 6] - ex3d[:, 4])).transpose()


def compute_targets(gts_val, gts_ign, box_lbls, rois, fg_thresh, ign_thresh, bg_lo, bg_hi, best_thresh, gts_3d, anchors):
    transforms = np.zeros([len(rois), 12], dtype=np.float32)
    if gts_val.shape[0] > 0 or gts_ign.shape[0] > 0:
        ols_ign_max = np.amax(iou_ign(rois, gts_ign), axis=1) if gts_ign.shape[0] > 0 else np.zeros([rois.shape[0]], dtype=np.float32)
        if gts_val.shape[0] > 0:
            ols = iou(rois, gts_val)
            ols_max = np.amax(ols, axis=1)
            targets = np.argmax(ols, axis=1)
            gt_best_rois = np.argmax(ols, axis=0)
            gt_best_ols = np.amax(ols, axis=0)
            gt_best_rois = gt_best_rois[gt_best_ols >= best_thresh]
            fg_inds = np.unique(np.concatenate((np.flatnonzero(ols_max >= fg_thresh), gt_best_rois)))
            if len(fg_inds) > 0:
                src = rois[fg_inds, :]
                transforms[fg_inds, 0:4] = bbox_transform(src, gts_val[targets[fg_inds], :])
                src_3d = anchors[rois[fg_inds, 4].astype(np.int64), 4:]
                transforms[fg_inds, 5:] = bbox_transform_3d(src, src_3d, gts_3d[targets[fg_inds]])
                transforms[fg_inds, 4] = [box_lbls[x] for x in targets[fg_inds]]
        else:
            ols_max = np.zeros(rois.shape[0], dtype=int)
            fg_inds = np.empty(shape=[0])
            gt_best_rois = np.empty(shape=[0])
        ign_inds = np.flatnonzero(ols_ign_max >= ign_thresh)
        bg_inds = np.flatnonzero((ols_max >= bg_lo) & (ols_max < bg_hi))
        bg_inds = np.setdiff1d(np.setdiff1d(np.setdiff1d(bg_inds, ign_inds), fg_inds), gt_best_rois)
        transforms[bg_inds, 4] = -1
    else:
        transforms[:, 4] = -1
    return transforms


def postprocess(transforms, n_val, means, stds):
    """lib/dataloader.py:1086-1125: normalisation and the label arrays of one image (transforms is None without a valid box)."""
    M = len(transforms) if transforms is not None else None
    if n_val > 0:
        t = transforms.copy()
        t[:, 0:4] -= means[:, 0:4]
        t[:, 0:4] /= stds[:, 0:4]
        t[:, 5:12] -= means[:, 4:]
        t[:, 5:12] /= stds[:, 4:]
        fg, bg, ign = t[:, 4] > 0, t[:, 4] < 0, t[:, 4] == 0
        labels = np.zeros(M, dtype=np.int64)
        labels[fg] = t[fg, 4]
        labels[ign] = IGN_FLAG
        return fg, bg, ign, labels, t[:, 0:4].copy(), t[:, 5:12].copy()
    raise ValueError("no valid box: the caller fills background-only arrays (lib/dataloader.py:1127-1131)")


def targets_image(g, rois, conf):
    """One image: dict of gts_val / gts_ign / box_lbls / gts_3d -> (labels_fg, labels_bg, labels_ign, labels, bbox_2d, bbox_3d, any_val)."""
    M = len(rois)
    if len(g["gts_val"]) > 0:
        t = compute_targets(np.asarray(g["gts_val"], dtype=np.float64), np.asarray(g["gts_ign"], dtype=np.float64).reshape(-1, 4),
                            np.asarray(g["box_lbls"]), rois, conf.fg_thresh, conf.ign_thresh, conf.bg_thresh_lo, conf.bg_thresh_hi,
                            conf.best_thresh, np.asarray(g["gts_3d"], dtype=np.float64), np.asarray(conf.anchors))
        return postprocess(t, len(g["gts_val"]), np.asarray(conf.bbox_means), np.asarray(conf.bbox_stds)) + (1,)
    z = np.zeros(M, dtype=bool)
    return z, ~z, z.copy(), np.zeros(M, dtype=np.int64), np.zeros((M, 4), np.float32), np.zeros((M, 7), np.float32), 0


def synthetic_gts(conf, n_val, n_ign, seed, image_hw=(384, 1280)):
    """Seeded KITTI-like annotations of one image: boxes whose sizes follow the anchors' (so that some anchors clear the
    0.5 IoU bar and others only the best-anchor rule), classes 1..3, 3D centre / depth / size / rotation."""
    rng = np.random.default_rng(seed)
    H, W = image_hw

    def boxes(n):
        h = np.exp(rng.uniform(np.log(20.0), np.log(0.7 * H), n))
        w = h * rng.uniform(0.4, 1.8, n)
        cx, cy = rng.uniform(0, W, n), rng.uniform(0.3 * H, 0.9 * H, n)
        return np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1)
    val = boxes(n_val)
    g3d = np.stack([(val[:, 0] + val[:, 2]) / 2 + rng.normal(0, 3, n_val), (val[:, 1] + val[:, 3]) / 2 + rng.normal(0, 3, n_val),
                    rng.uniform(4, 60, n_val), rng.uniform(0.5, 2.0, n_val), rng.uniform(1.2, 2.0, n_val),
                    rng.uniform(0.8, 4.5, n_val), rng.uniform(-3.1, 3.1, n_val)], axis=1) if n_val else np.zeros((0, 7))
    return {"gts_val": val, "gts_ign": boxes(n_ign), "box_lbls": rng.integers(1, len(conf.lbls) + 1, n_val), "gts_3d": g3d}
