set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "featur or pipeline or smoke or trig" > gpurun_out/t44.log 2>&1; echo "pytest_exit=$?"; tail -n 4 gpurun_out/t44.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench44_10M.log 2> gpurun_out/bench44_10M.err; echo "bench_10M_exit=$?"
