set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t42.log 2>&1; echo "pytest_exit=$?"; tail -n 4 gpurun_out/t42.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench42_10M.log 2> gpurun_out/bench42_10M.err; echo "bench_10M_exit=$?"
