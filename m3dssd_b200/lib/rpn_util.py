"""The RPN helpers the model imports (mirror of the hot-path subset of lib/rpn_util.py):
calc_output_size (:1401-1413), locate_anchors (:1329-1398), flatten_tensor (:892-901),
bbox_transform_inv (:1137-1186), im_detect_3d (:1416-1563) and the result-writing loop of test_kitti_3d (:1753-1852).
The dataset / evaluation / plotting parts of that file are out of scope."""
import os

import numpy as np
import torch

from .nms.gpu_nms import gpu_nms


def calc_output_size(res, stride):
    return np.ceil(np.array(res) / stride).astype(int)


def locate_anchors(anchors, feat_size, stride, convert_tensor=False):
    """[(A*H*W), 5] rows (x1, y1, x2, y2, anchor index), anchor-major then row-major over the feature map."""
    if torch.is_tensor(anchors):
        anchors = anchors.detach().cpu().numpy()
    H, W = int(feat_size[0]), int(feat_size[1])
    sx = (np.arange(W, dtype=np.float64) * float(stride))[None, None, :]
    sy = (np.arange(H, dtype=np.float64) * float(stride))[None, :, None]
    a = anchors[:, 0:4]
    cols = [np.broadcast_to(s + a[:, k][:, None, None], (a.shape[0], H, W))
            for k, s in ((0, sx), (1, sy), (2, sx), (3, sy))]
    cols.append(np.broadcast_to(np.arange(a.shape[0], dtype=np.float64)[:, None, None], (a.shape[0], H, W)))
    rois = np.stack([c.reshape(-1) for c in cols], axis=1)
    return torch.from_numpy(rois) if convert_tensor else rois


def flatten_tensor(input):
    """[B, C, H, W] -> [B, H*W, C]"""
    return input.permute(0, 2, 3, 1).contiguous().view(input.shape[0], -1, input.shape[1])


def bbox_transform_inv(boxes, deltas, means=None, stds=None):
    if boxes.shape[0] == 0:
        return torch.zeros((0, deltas.shape[1]), dtype=deltas.dtype, device=deltas.device)
    widths = boxes[:, 2] - boxes[:, 0] + 1.0
    heights = boxes[:, 3] - boxes[:, 1] + 1.0
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    d = deltas
    if stds is not None:
        d = d * stds[0:4]
    if means is not None:
        d = d + means[0:4]
    pcx, pcy = d[:, 0] * widths + ctr_x, d[:, 1] * heights + ctr_y
    pw, ph = torch.exp(d[:, 2]) * widths, torch.exp(d[:, 3]) * heights
    return torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=1)


def im_detect_3d(im, net, rpn_conf, obj, gpu=0, synced=False):
    """Single-image detection with the reference's signature and return value
    (numpy [n_kept, 14]); decode, top-K and NMS all run on the device through the
    network's detection tail (no host round trip before NMS)."""
    if im.dim() == 3:
        im = im[None]
    im = im.cuda()
    with torch.no_grad():
        net.eval()
        kept, num = net.detect(im, scale_factor=float(obj.scale_factor), max_out=int(rpn_conf.nms_topN_pre))
    aboxes = kept[0, :int(num[0].item())].cpu().numpy()
    if rpn_conf.clip_boxes:
        aboxes[:, 0] = np.clip(aboxes[:, 0], 0, obj.imW - 1)
        aboxes[:, 1] = np.clip(aboxes[:, 1], 0, obj.imH - 1)
        aboxes[:, 2] = np.clip(aboxes[:, 2], 0, obj.imW - 1)
        aboxes[:, 3] = np.clip(aboxes[:, 3], 0, obj.imH - 1)
    return aboxes


# ---------------------------------------------------------------------------------------------------------
# Post-NMS 3D refinement + KITTI result lines: the loop body of test_kitti_3d
# (lib/rpn_util.py:1801-1852) with hill_climb (:652-708) on the device, all kept boxes of a batch in one launch.
# ---------------------------------------------------------------------------------------------------------
def refine_detections(kept, num_keep, p2, rpn_conf=None, score_thresh=0.75, hill_climbing=None):
    """kept [B, max_out, >=13] fp32 CUDA rows (x1, y1, x2, y2, score, cls, x3d, y3d, z3d, w3d, h3d, l3d, alpha, ...)
    as the engine's NMS gather leaves them, num_keep [B] int32 CUDA, p2 [4, 4] or [B, 4, 4].
    Returns (rows float64 CUDA [B, max_out, 14], valid bool CUDA [B, max_out]); rows = (class index, alpha, x1, y1,
    x2, y2, h3d, w3d, l3d, x3d, y3d, z3d, ry3d, score), the numbers of the reference's KITTI line.
    Raises NotImplementedError on CPU tensors (no CPU fallback, like the DCN op)."""
    from .. import ops
    if not (torch.is_tensor(kept) and kept.is_cuda):
        raise NotImplementedError("refine_detections needs CUDA tensors: m3dssd_b200 has no CPU fallback")
    if hill_climbing is None:
        hill_climbing = bool(getattr(rpn_conf, "hill_climbing", True)) if rpn_conf is not None else True
    out, valid = ops.refine_3d(kept.float().contiguous(), num_keep.int().contiguous(), p2, score_thresh=score_thresh,
                               hill_climbing=hill_climbing)
    return out, valid.bool()


def kitti_result_lines(rows, valid, lbls):
    """Format one image's refined rows exactly like lib/rpn_util.py:1846-1847."""
    rows = rows.detach().cpu().numpy() if torch.is_tensor(rows) else np.asarray(rows)
    valid = valid.detach().cpu().numpy() if torch.is_tensor(valid) else np.asarray(valid)
    text = ''
    for r, ok in zip(rows, valid):
        if not ok:
            continue
        cls = lbls[int(r[0])]
        text += ('{} -1 -1 {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} '
                 + '{:.6f} {:.6f}\n').format(cls, r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11],
                                             r[12], r[13])
    return text


def _iter_test_items(dataset_test, rpn_conf):
    """(image [3,H,W] or [1,3,H,W] tensor / uint8 HWC frame, imobj) pairs out of the reference's test loader items:
    `(im, imobj)` tuples, or {'input', 'target': {'meta'}} dicts when conf.pre_compute_target
    (lib/rpn_util.py:1776-1780)."""
    for batch in dataset_test:
        if isinstance(batch, dict):
            im, imobj = batch["input"], batch["target"]["meta"]
        else:
            im, imobj = batch
        yield im, imobj


def _field(imobj, name, default=None):
    v = imobj[name] if isinstance(imobj, dict) and name in imobj else getattr(imobj, name, default)
    if isinstance(v, (list, tuple)) and len(v) == 1:  # DataLoader(batch_size=1) collation wraps ids in a list
        v = v[0]
    return v


def test_kitti_3d(dataset_test, net, rpn_conf, results_path, test_path=None, use_log=True, writer=None,
                  phase="validation", batch_size=8, **_unused):
    """The reference's evaluation driver (lib/rpn_util.py:1753-1860) up to and including the KITTI result files:
    for every test image, detect -> first nms_topN_post kept boxes -> score cut 0.75 -> alpha -> rotation, hill_climb,
    back-projection (:1801-1844) -> one `<id>.txt` per image (:1846-1850).  Same signature and file contents; the
    images go through the network `batch_size` at a time, and decode / NMS / refinement stay on the device (one D2H
    of <= 40 rows per image).  Images arrive as the reference's loader yields them (normalised [3,H,W] / [1,3,H,W]
    tensors) or as raw uint8 HWC frames (then Preprocess runs on the device).  The KITTI AP evaluation the reference
    runs on those files afterwards (lib/eval get_official_eval_result, :1866-1900) is outside the path; extra keyword
    arguments (scripts/test_rpn_3d.py:59 passes val_train=) are accepted and ignored.  Returns the files written."""
    os.makedirs(results_path, exist_ok=True)
    net.eval()
    written, pending = [], []

    def flush():
        if not pending:
            return
        ims = [p[0] for p in pending]
        objs = [p[1] for p in pending]
        n = len(ims)
        while len(ims) < batch_size:  # a short last batch reuses the engine of the full ones
            ims.append(ims[-1])
        scale = float(_field(objs[0], "scale_factor", 1.0))
        post = int(rpn_conf.nms_topN_post)
        kw = {} if post == int(net.conf.nms_topN_post) else dict(max_out=post)  # (the engine's default: one engine, not two)
        with torch.no_grad():
            if torch.is_tensor(ims[0]) and ims[0].dtype != torch.uint8:
                x = torch.stack([im.reshape(im.shape[-3:]) for im in ims]).cuda(non_blocking=True)
                kept, num = net.detect(x, scale_factor=scale, **kw)
            else:
                kept, num = net.detect_images(ims, scale_factor=scale, **kw)
            if rpn_conf.clip_boxes:  # im_detect_3d clips before the refinement (:1552-1556)
                for b in range(n):
                    imW, imH = float(_field(objs[b], "imW")), float(_field(objs[b], "imH"))
                    kept[b, :, 0].clamp_(0, imW - 1), kept[b, :, 2].clamp_(0, imW - 1)
                    kept[b, :, 1].clamp_(0, imH - 1), kept[b, :, 3].clamp_(0, imH - 1)
            p2 = np.stack([np.asarray(_field(o, "p2"), dtype=np.float64).reshape(4, 4) for o in objs] +
                          [np.asarray(_field(objs[-1], "p2"), dtype=np.float64).reshape(4, 4)] * (batch_size - n))
            rows, valid = refine_detections(kept, num, p2, rpn_conf)
        rows, valid = rows.cpu().numpy(), valid.cpu().numpy()
        for b in range(n):
            path = os.path.join(results_path, str(_field(objs[b], "id")) + ".txt")
            with open(path, "w") as f:
                f.write(kitti_result_lines(rows[b], valid[b], rpn_conf.lbls))
            written.append(path)
        pending.clear()

    for im, imobj in _iter_test_items(dataset_test, rpn_conf):
        if pending and (float(_field(imobj, "scale_factor", 1.0)) != float(_field(pending[0][1], "scale_factor", 1.0))):
            flush()  # the decode takes one scale factor per launch
        pending.append((im, imobj))
        if len(pending) == batch_size:
            flush()
    flush()
    return written

