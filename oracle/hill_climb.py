"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy float64) of the reference's post-NMS 3D refinement, the step
right after the detection path:

  * project_3d            lib/rpn_util.py:921-970
  * test_projection       lib/rpn_util.py:2015-2050
  * hill_climb            lib/rpn_util.py:652-708
  * convertAlpha2Rot / convertRot2Alpha   lib/util.py:516-535
  * the per-box loop of test_kitti_3d that feeds the KITTI writer   lib/rpn_util.py:1801-1852

Pinned against the unmodified reference functions by tests/golden/make_golden.py (fixture
tests/golden/hill_climb.npz).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.
"""
import math

import numpy as np


def convert_alpha2rot(alpha, z3d, x3d):
    ry3d = alpha + math.atan2(-z3d, x3d) + 0.5 * math.pi
    while ry3d > math.pi:
        ry3d -= math.pi * 2
    while ry3d < (-math.pi):
        ry3d += math.pi * 2
    return ry3d


def convert_rot2alpha(ry3d, z3d, x3d):
    alpha = ry3d - math.atan2(-z3d, x3d) - 0.5 * math.pi
    while alpha > math.pi:
        alpha -= math.pi * 2
    while alpha < (-math.pi):
        alpha += math.pi * 2
    return alpha


def project_3d(p2, x3d, y3d, z3d, w3d, h3d, l3d, ry3d):
    """-> (2D bounding extent of the 8 projected corners, min corner depth)."""
    c, s = math.cos(ry3d), math.sin(ry3d)
    R = np.array([[+c, 0, +s], [0, 1, 0], [-s, 0, +c]])
    x_corners = np.array([0, l3d, l3d, l3d, l3d, 0, 0, 0], dtype=np.float64) - l3d / 2
    y_corners = np.array([0, 0, h3d, h3d, 0, 0, h3d, h3d], dtype=np.float64) - h3d / 2
    z_corners = np.array([0, 0, 0, w3d, w3d, w3d, w3d, 0], dtype=np.float64) - w3d / 2
    corners_3d = R.dot(np.array([x_corners, y_corners, z_corners]))
    corners_3d += np.array([x3d, y3d, z3d]).reshape((3, 1))
    corners_2d = p2.dot(np.vstack((corners_3d, np.ones((8,)))))
    corners_2d = corners_2d / corners_2d[2]
    return corners_2d[:2], corners_3d


def test_projection(p2, p2_inv, box_2d, cx, cy, z, w3d, h3d, l3d, rot_y):
    x, y = box_2d[0], box_2d[1]
    x2, y2 = x + box_2d[2] - 1, y + box_2d[3] - 1
    coord3d = p2_inv.dot(np.array([cx * z, cy * z, z, 1]))
    verts, corners_3d = project_3d(p2, coord3d[0], coord3d[1], coord3d[2], w3d, h3d, l3d, rot_y)
    invalid = bool(np.any(corners_3d[2, :] <= 0))
    x_new, y_new, x2_new, y2_new = verts[0].min(), verts[1].min(), verts[0].max(), verts[1].max()
    ol = -(abs(x - x_new) + abs(y - y_new) + abs(x2 - x2_new) + abs(y2 - y2_new))
    return ol, invalid


def hill_climb(p2, p2_inv, box_2d, x2d, y2d, z2d, w3d, h3d, l3d, ry3d, step_z_init=0, step_r_init=0, z_lim=0, r_lim=0,
               min_ol_dif=0.0):
    step_z, step_r = step_z_init, step_r_init
    ol_best, invalid = test_projection(p2, p2_inv, box_2d, x2d, y2d, z2d, w3d, h3d, l3d, ry3d)
    if invalid:
        return z2d, ry3d
    while step_z > z_lim or step_r > r_lim:
        if step_z > z_lim:
            ol_neg, inv_neg = test_projection(p2, p2_inv, box_2d, x2d, y2d, z2d - step_z, w3d, h3d, l3d, ry3d)
            ol_pos, inv_pos = test_projection(p2, p2_inv, box_2d, x2d, y2d, z2d + step_z, w3d, h3d, l3d, ry3d)
            if ((ol_pos - ol_best) <= min_ol_dif) and ((ol_neg - ol_best) <= min_ol_dif):
                step_z = step_z * 0.5
            elif (ol_pos - ol_best) > min_ol_dif and ol_pos > ol_neg and not inv_pos:
                z2d += step_z
                ol_best = ol_pos
            elif (ol_neg - ol_best) > min_ol_dif and not inv_neg:
                z2d -= step_z
                ol_best = ol_neg
            else:
                step_z = step_z * 0.5
        if step_r > r_lim:
            ol_neg, inv_neg = test_projection(p2, p2_inv, box_2d, x2d, y2d, z2d, w3d, h3d, l3d, ry3d - step_r)
            ol_pos, inv_pos = test_projection(p2, p2_inv, box_2d, x2d, y2d, z2d, w3d, h3d, l3d, ry3d + step_r)
            if ((ol_pos - ol_best) <= min_ol_dif) and ((ol_neg - ol_best) <= min_ol_dif):
                step_r = step_r * 0.5
            elif (ol_pos - ol_best) > min_ol_dif and ol_pos > ol_neg and not inv_pos:
                ry3d += step_r
                ol_best = ol_pos
            elif (ol_neg - ol_best) > min_ol_dif and not inv_neg:
                ry3d -= step_r
                ol_best = ol_neg
            else:
                step_r = step_r * 0.5
    while ry3d > math.pi:
        ry3d -= math.pi * 2
    while ry3d < (-math.pi):
        ry3d += math.pi * 2
    return z2d, ry3d


def refine_detections(aboxes, p2, hill_climbing=True, score_thresh=0.75, max_out=40):
    """The per-box loop of test_kitti_3d (lib/rpn_util.py:1801-1852): rows of `aboxes` [n, >=13] (x1, y1, x2, y2,
    score, cls, x3d, y3d, z3d, w3d, h3d, l3d, alpha) -> [m, 14] float64 rows in the order the KITTI line prints
    them: (cls index, alpha, x1, y1, x2, y2, h3d, w3d, l3d, x3d, y3d, z3d, ry3d, score), m = boxes above the
    score cut among the first max_out."""
    p2 = np.asarray(p2, dtype=np.float64)
    p2_inv = np.linalg.inv(p2)
    rows = []
    for i in range(min(max_out, aboxes.shape[0])):
        box = aboxes[i].astype(np.float64)
        score = box[4]
        if not score >= score_thresh:
            continue
        x1, y1, x2, y2 = box[0], box[1], box[2], box[3]
        width, height = (x2 - x1 + 1), (y2 - y1 + 1)
        x3d, y3d, z3d, w3d, h3d, l3d, ry3d = box[6], box[7], box[8], box[9], box[10], box[11], box[12]
        coord3d = p2_inv.dot(np.array([x3d * z3d, y3d * z3d, 1 * z3d, 1]))
        ry3d = convert_alpha2rot(ry3d, coord3d[2], coord3d[0])
        if hill_climbing:
            z3d, ry3d = hill_climb(p2, p2_inv, np.array([x1, y1, width, height]), x3d, y3d, z3d, w3d, h3d, l3d, ry3d,
                                   step_r_init=0.3 * math.pi, r_lim=0.01)
        coord3d = p2_inv.dot(np.array([x3d * z3d, y3d * z3d, 1 * z3d, 1]))
        alpha = convert_rot2alpha(ry3d, coord3d[2], coord3d[0])
        rows.append([box[5] - 1, alpha, x1, y1, x2, y2, h3d, w3d, l3d, coord3d[0], coord3d[1] + h3d / 2, coord3d[2],
                     ry3d, score])
    return np.asarray(rows, dtype=np.float64).reshape(-1, 14)


def synthetic_detections(n, seed, hw=(384, 1280)):
    """Seeded KITTI-like post-NMS rows [n, 14] float32 + a KITTI-like P2 (fx = fy = 721.5) for tests / goldens."""
    rng = np.random.default_rng(seed)
    p2 = np.array([[721.5377, 0.0, 609.5593, 44.85728], [0.0, 721.5377, 172.854, 0.2163791],
                   [0.0, 0.0, 1.0, 0.002745884], [0.0, 0.0, 0.0, 1.0]])
    z = rng.uniform(4.0, 60.0, n)
    x = rng.uniform(-12.0, 12.0, n) * z / 20.0
    y = rng.uniform(1.0, 2.0, n)
    w3d, h3d, l3d = rng.uniform(1.4, 1.9, n), rng.uniform(1.3, 1.8, n), rng.uniform(3.2, 4.6, n)
    ry = rng.uniform(-math.pi, math.pi, n)
    rows = np.zeros((n, 14), dtype=np.float32)
    for i in range(n):
        verts, _ = project_3d(p2, x[i], y[i] - h3d[i] / 2, z[i], w3d[i], h3d[i], l3d[i], ry[i])
        b = np.array([verts[0].min(), verts[1].min(), verts[0].max(), verts[1].max()])
        b += rng.normal(0, 2.0, 4)  # the 2D head disagrees a little with the 3D head: that is what hill_climb fixes
        c = p2.dot(np.array([x[i], y[i] - h3d[i] / 2, z[i], 1.0]))
        alpha = convert_rot2alpha(ry[i] + rng.normal(0, 0.3), z[i], x[i])
        rows[i] = [b[0], b[1], b[2], b[3], rng.uniform(0.5, 1.0), rng.integers(1, 4), c[0] / c[2], c[1] / c[2], c[2],
                   w3d[i], h3d[i], l3d[i], alpha, rng.integers(0, 36)]
    rows = rows[np.argsort(-rows[:, 4], kind="stable")]
    return rows, p2
