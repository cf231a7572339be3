#!/bin/bash
# Build build/libm3d_probe.so: the library with the timeline probes compiled in (-DM3D_PROBE) for
# tools/probe_heads.py, probe_stem.py, probe_dcn_timeline.py.  Run `python -m m3dssd_b200.build` first.
set -e
cd "$(dirname "$0")/.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr -DM3D_PROBE -I include"
for f in heads igemm dcn_fused; do nvcc $FLAGS -c m3dssd_b200/csrc/$f.cu -o build/${f}_probe.o; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/libm3d_probe.so \
  $(ls build/obj/*.o | grep -v "/heads.o\|/igemm.o\|/dcn_fused.o") build/heads_probe.o build/igemm_probe.o build/dcn_fused_probe.o
echo built build/libm3d_probe.so
