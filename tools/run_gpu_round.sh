set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t34.log 2>&1; echo "pytest_exit=$?"; tail -n 4 gpurun_out/t34.log
python bench.py --frames-per-gpu 1250000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench34_small.log 2> gpurun_out/bench34_small.err; echo "bench_small_exit=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench34_10M.log 2> gpurun_out/bench34_10M.err; echo "bench_10M_exit=$?"
