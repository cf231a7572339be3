M3D_DCN_HALO=0 timeout 120 python tools/probe_dcn_timeline.py | head -4
