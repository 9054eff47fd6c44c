set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "kmeans or assign" > gpurun_out/t39.log 2>&1; echo "pytest_exit=$?"; tail -n 3 gpurun_out/t39.log
python bench.py --frames-per-gpu 1250000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench39_small.log 2> gpurun_out/bench39_small.err; echo "bench_small_exit=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench39_10M.log 2> gpurun_out/bench39_10M.err; echo "bench_10M_exit=$?"
