echo "=== new (unified window)"; timeout 300 python tools/probe_conv_perf.py level2 level0 offset
echo "=== old (3 column windows)"; M3D_LIB=$PWD/m3dssd_b200/libm3dssd_b200.old.so timeout 300 python tools/probe_conv_perf.py level2 level0 offset
