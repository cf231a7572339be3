"""argtypes/restype declarations for the C ABI beyond m3d_conv2d_nhwc."""
import ctypes as C


def declare(L):
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    for name, args in _SIGS.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = args


_SIGS = {}
