"""ctypes binding of libm3dssd_b200.so (the C ABI declared in include/m3dssd_b200.h).

The library is the product: importing this module fails loudly when it has not
been built -- there is no Python/CPU fallback for any op.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("M3D_LIB") or os.path.join(_HERE, "libm3dssd_b200.so")  # M3D_LIB: development A/B builds

M3D_BF16, M3D_F32, M3D_BF16X3 = 0, 1, 2
MAX_CONCAT = 6


class M3DError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("act_dtype", C.c_int), ("out_dtype", C.c_int), ("num_inputs", C.c_int),
        ("in_", C.c_void_p * MAX_CONCAT),
        ("in_c", C.c_int * MAX_CONCAT), ("in_cstride", C.c_int * MAX_CONCAT),
        ("in_coff", C.c_int * MAX_CONCAT), ("in_goff", C.c_int * MAX_CONCAT),
        ("N", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("R", C.c_int), ("S", C.c_int), ("stride", C.c_int), ("pad", C.c_int), ("dil", C.c_int),
        ("out_h", C.c_int), ("out_w", C.c_int),
        ("Cout", C.c_int), ("groups", C.c_int),
        ("weight", C.c_void_p), ("weight_mid", C.c_void_p), ("weight_lo", C.c_void_p), ("weight_f32", C.c_void_p),
        ("weight_rows", C.c_int), ("weight_goff", C.c_int),
        ("bias", C.c_void_p), ("bias_goff", C.c_int),
        ("res", C.c_void_p), ("res_cstride", C.c_int), ("res_coff", C.c_int), ("res_goff", C.c_int),
        ("out", C.c_void_p), ("out_cstride", C.c_int), ("out_coff", C.c_int), ("out_goff", C.c_int),
        ("slope", C.c_float),
        ("om", C.c_void_p), ("om_cstride", C.c_int), ("sigmoid_mask", C.c_int),
        ("force_gather", C.c_int),
        ("k16_zero", C.c_ulonglong * 2),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise M3DError(
                "libm3dssd_b200.so is not built (%s). Run `python -m m3dssd_b200.build`; "
                "there is no CPU fallback." % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    L.m3d_last_error.restype = C.c_char_p
    L.m3d_last_error.argtypes = []
    L.m3d_version.restype = C.c_int
    L.m3d_last_kernel.restype = C.c_char_p
    L.m3d_last_kernel.argtypes = []
    L.m3d_conv2d_nhwc.restype = C.c_int
    L.m3d_conv2d_nhwc.argtypes = [C.POINTER(ConvDesc), C.c_void_p]
    from . import _decl
    _decl.declare(L)
    if L.m3d_conv_desc_size() != C.sizeof(ConvDesc):  # a stale .so next to newer Python (or the reverse)
        raise M3DError("m3d_conv_desc is %d bytes in %s but %d in the Python mirror: rebuild with `python -m m3dssd_b200.build`"
                       % (L.m3d_conv_desc_size(), LIB_PATH, C.sizeof(ConvDesc)))


def check(rc):
    if rc != 0:
        raise M3DError("m3dssd_b200 error %d: %s" % (rc, lib().m3d_last_error().decode()))
