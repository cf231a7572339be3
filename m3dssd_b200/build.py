"""Build libm3dssd_b200.so (all CUDA kernels + the C ABI) in-tree with nvcc for sm_100a.

    python -m m3dssd_b200.build [--force]

Objects go to build/obj (git-ignored); the shared library is written next to
this file so it travels with the source snapshot to the GPU box.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# Development: M3D_VARIANT=name builds libm3dssd_b200.name.so (own object cache) with M3D_NVCC_EXTRA flags, and
# M3D_LIB=<path> makes m3dssd_b200._lib load it -- several kernel variants can then be A/B-timed in ONE GPU session.
VARIANT = os.environ.get("M3D_VARIANT", "")
OBJ = os.path.join(ROOT, "build", "obj" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(HERE, "libm3dssd_b200%s.so" % ("." + VARIANT if VARIANT else ""))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "--expt-relaxed-constexpr",
    # devIoU and the bilinear blend must round like the reference's plain fp32
    # expressions: no fast-math anywhere.
]
# Development knob: extra nvcc flags, e.g. M3D_NVCC_EXTRA=-DM3D_SOME_VARIANT (experimental kernel variants that are
# compiled out of the default build).  Part of the object cache key.
FLAGS += os.environ.get("M3D_NVCC_EXTRA", "").split()


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest():
    h = hashlib.sha1()
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cuh", ".h")):
                with open(os.path.join(d, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, digest, force):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    stamp = obj + ".stamp"
    with open(os.path.join(CSRC, src), "rb") as fh:
        key = hashlib.sha1(fh.read() + digest.encode()).hexdigest()
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == key:
        return obj, False
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp, "w") as fh:
        fh.write(key)
    return obj, True


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    digest = _headers_digest()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, digest, force), srcs))
    objs = [o for o, _ in res]
    changed = any(c for _, c in res)
    if changed or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("built", LIB)
    elif verbose:
        print("up to date:", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
