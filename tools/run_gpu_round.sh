timeout 300 python tools/km_dim_probe.py 2>&1 | tail -12
