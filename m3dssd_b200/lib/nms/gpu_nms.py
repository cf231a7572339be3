"""gpu_nms(dets, thresh, device_id=0) -> list of kept indices (mirror of lib/nms/gpu_nms.pyx:16-31).

Same contract as the reference's Cython wrapper: `dets` is a host float32 array
[N, 5] (x1, y1, x2, y2, score); rows are ordered by score with numpy's
argsort()[::-1] exactly as the reference does, the sorted boxes go through the
C-ABI replacement of `_nms` (m3d_nms), and the kept positions are mapped back to
the caller's indices.
"""
import ctypes as C

import numpy as np

from ..._lib import check, lib


def gpu_nms(dets, thresh, device_id=0):
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.ndim != 2 or dets.shape[1] < 5:
        raise ValueError("dets must be [N, >=5] float32 (x1, y1, x2, y2, score)")
    boxes_num, boxes_dim = dets.shape
    if boxes_num == 0:
        return []
    order = dets[:, 4].argsort()[::-1]
    sorted_dets = np.ascontiguousarray(dets[order, :])
    keep = np.zeros(boxes_num, dtype=np.int32)
    num_out = C.c_int(0)
    check(lib().m3d_nms(keep.ctypes.data_as(C.c_void_p), C.byref(num_out), sorted_dets.ctypes.data_as(C.c_void_p),
                        boxes_num, boxes_dim, C.c_float(thresh), int(device_id)))
    return list(order[keep[:num_out.value]])
