timeout 120 python tools/probe_halo2.py
