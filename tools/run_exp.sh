python tools/km_bench.py c4 2>&1 | tail -5
timeout 900 python -m pytest tests -x -q -m gpu -k "kmeans or cluster or lloyd or assign" 2>&1 | tail -3
