PMB_TICA_CHOL=0 python tools/tica_bench.py 256
PMB_TICA_CHOL=1 python tools/tica_bench.py 256
PMB_TICA_CHOL=1 python tools/tica_bench.py 84
PMB_TICA_CHOL=1 python tools/tica_bench.py 400 3
timeout 600 python -m pytest tests -x -q -m gpu -k "tica or vamp or reduce or c1 or c3" 2>&1 | tail -3
