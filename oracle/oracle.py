"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front-end of the CPU oracle (libm3d_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module; the product package (m3dssd_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libm3d_oracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _prep(dtype, *arrs):
    return [None if a is None else np.ascontiguousarray(a, dtype=dtype) for a in arrs]


def dcn_out_shape(H, W, kh, kw, stride, pad, dil):
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(pad), _pair(dil)
    return ((H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1, (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1)


def dcn_v2_forward(input, offset, mask, weight, bias, stride=1, pad=0, dil=1, dg=1, dtype=np.float32):
    """DCNv2Function.forward semantics on NCHW arrays (dcn_v2_cuda.c:10-102)."""
    x, off, m, w, b = _prep(dtype, input, offset, mask, weight, bias)
    B, Cin, H, W = x.shape
    Cout, Ck, kh, kw = w.shape
    if Ck != Cin:
        raise RuntimeError("Input shape and kernel channels wont match: (%d vs %d)." % (Cin, Ck))
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(pad), _pair(dil)
    Ho, Wo = dcn_out_shape(H, W, kh, kw, stride, pad, dil)
    assert off.shape == (B, 2 * dg * kh * kw, Ho, Wo), off.shape
    assert m.shape == (B, dg * kh * kw, Ho, Wo), m.shape
    out = np.empty((B, Cout, Ho, Wo), dtype=dtype)
    fn = lib().m3d_oracle_dcn_forward_f32 if dtype == np.float32 else lib().m3d_oracle_dcn_forward_f64
    rc = fn(_ptr(x), _ptr(off), _ptr(m), _ptr(w), _ptr(b) if b is not None else None, _ptr(out),
            B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, dg)
    if rc != 0:
        raise RuntimeError("oracle dcn_forward failed rc=%d" % rc)
    return out


def dcn_v2_im2col(input1, offset1, mask1, kh, kw, stride=1, pad=0, dil=1, dg=1):
    """columns [C*kh*kw, Ho, Wo] for one sample (fp32), as modulated_deformable_im2col_cuda writes them."""
    x, off, m = _prep(np.float32, input1, offset1, mask1)
    Cin, H, W = x.shape
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(pad), _pair(dil)
    Ho, Wo = dcn_out_shape(H, W, kh, kw, stride, pad, dil)
    col = np.empty((Cin * kh * kw, Ho, Wo), dtype=np.float32)
    lib().m3d_oracle_dcn_im2col_f32(_ptr(x), _ptr(off), _ptr(m), Cin, H, W, Ho, Wo, kh, kw, ph, pw, sh, sw, dh, dw, dg,
                                    _ptr(col))
    return col


def dcn_v2_backward(input, offset, mask, weight, grad_output, stride=1, pad=0, dil=1, dg=1, dtype=np.float32):
    """DCNv2Function.backward: returns (grad_input, grad_offset, grad_mask, grad_weight, grad_bias)."""
    x, off, m, w, gy = _prep(dtype, input, offset, mask, weight, grad_output)
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = w.shape
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(pad), _pair(dil)
    gi, go, gm, gw = np.zeros_like(x), np.zeros_like(off), np.zeros_like(m), np.zeros_like(w)
    gb = np.zeros((Cout,), dtype=dtype)
    fn = lib().m3d_oracle_dcn_backward_f32 if dtype == np.float32 else lib().m3d_oracle_dcn_backward_f64
    rc = fn(_ptr(x), _ptr(off), _ptr(m), _ptr(w), _ptr(gy), _ptr(gi), _ptr(go), _ptr(gm), _ptr(gw), _ptr(gb),
            B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, dg)
    if rc != 0:
        raise RuntimeError("oracle dcn_backward failed rc=%d" % rc)
    return gi, go, gm, gw, gb


def nms_sorted(boxes, thresh):
    """`_nms` (lib/nms/nms_kernel.cu:91-144) on score-sorted boxes [N,>=4] fp32 -> kept row indices."""
    b = np.ascontiguousarray(boxes, dtype=np.float32)
    n, dim = b.shape
    keep = np.zeros((max(n, 1),), dtype=np.int32)
    num = C.c_int(0)
    lib().m3d_oracle_nms(_ptr(keep), C.byref(num), _ptr(b), n, dim, C.c_float(thresh))
    return keep[:num.value].copy()


def nms_mask(boxes, thresh):
    b = np.ascontiguousarray(boxes, dtype=np.float32)
    n, dim = b.shape
    cb = (n + 63) // 64
    mask = np.zeros((n, cb), dtype=np.uint64)
    lib().m3d_oracle_nms_mask(_ptr(mask), _ptr(b), n, dim, C.c_float(thresh))
    return mask


def gpu_nms(dets, thresh, device_id=0):
    """gpu_nms (lib/nms/gpu_nms.pyx:16-31): sort by score (descending), _nms, map back."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.shape[0] == 0:
        return []
    order = dets[:, 4].argsort()[::-1]
    keep = nms_sorted(dets[order, :], thresh)
    return list(order[keep])


def preprocess_u8(im_hwc, mean, std):
    """The reference's test-time input transform restated (lib/augmentations.py:44-57 Normalize; lib/dataloader.py:942-950
    BGR->RGB swap and HWC->CHW): uint8 [H,W,3] or [N,H,W,3] -> float32 [.., 3, H, W].  float32 arithmetic in the
    reference's order; mean / std are applied in the channel order of the INPUT."""
    im = np.asarray(im_hwc)
    assert im.dtype == np.uint8 and im.shape[-1] == 3
    x = im.astype(np.float32)
    x /= 255.0
    x -= np.asarray(mean, dtype=np.float32)
    x /= np.asarray(std, dtype=np.float32)
    x = x[..., ::-1]  # cv2.COLOR_BGR2RGB
    return np.ascontiguousarray(np.moveaxis(x, -1, -3)).astype(np.float32)


def preprocess_pad_u8(images, size, mean, std):
    """The reference's test-time transform restated for a ragged batch: Preprocess(size, mean, stds)
    (lib/augmentations.py:472-492) = ConvertToFloat (:36-41), Padding (:136-160: cv2.copyMakeBorder on the bottom /
    right with the constant 0 -- BEFORE Normalize, so a padded pixel ends up at -mean/std), Normalize (:44-57); then
    BGR->RGB and HWC->CHW (lib/dataloader.py:942-950).  images: list of uint8 [h_i, w_i, 3]; -> float32 [N,3,H,W].
    An image larger than `size` is an error (copyMakeBorder raises on a negative border)."""
    H, W = int(size[0]), int(size[1])
    out = np.empty((len(images), 3, H, W), dtype=np.float32)
    for n, im in enumerate(images):
        im = np.asarray(im)
        assert im.dtype == np.uint8 and im.ndim == 3 and im.shape[2] == 3
        h, w = im.shape[:2]
        if h > H or w > W:
            raise ValueError("image %d is %dx%d, larger than %dx%d" % (n, h, w, H, W))
        padded = np.zeros((H, W, 3), dtype=np.float32)
        padded[:h, :w] = im.astype(np.float32)
        padded /= 255.0
        padded -= np.asarray(mean, dtype=np.float32)
        padded /= np.asarray(std, dtype=np.float32)
        out[n] = np.moveaxis(padded[..., ::-1], -1, 0)
    return out

