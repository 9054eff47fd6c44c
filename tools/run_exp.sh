PMB_TICA_CLUSTER=0 python tools/tica_bench.py 256
PMB_TICA_CLUSTER=1 python tools/tica_bench.py 256
PMB_TICA_CLUSTER=0 python tools/tica_bench.py 84
PMB_TICA_CLUSTER=1 python tools/tica_bench.py 84
PMB_TICA_CLUSTER=1 python tools/tica_bench.py 512 3
timeout 600 python -m pytest tests -x -q -m gpu -k "tica or vamp or reduce" 2>&1 | tail -3
