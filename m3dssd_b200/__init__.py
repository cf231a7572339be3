"""m3dssd_b200 -- Blackwell-native (sm_100a) dense forward path of M3DSSD.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all
arithmetic runs in the hand-written CUDA kernels of libm3dssd_b200.so, reached
through the C ABI in include/m3dssd_b200.h.
"""
__version__ = "0.1.0"
