set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t8.log 2>&1; echo "pytest_exit=$?"
./tools/ubench > gpurun_out/ubench.log 2>&1; echo "ubench_exit=$?"
python bench.py --frames-per-gpu 1250000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench8_small.log 2> gpurun_out/bench8_small.err; echo "bench_small_exit=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench8_10M.log 2> gpurun_out/bench8_10M.err; echo "bench_10M_exit=$?"
python bench.py --frames-per-gpu 1250000 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --profile-range > gpurun_out/plain_r1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1b.csv python bench.py --frames-per-gpu 1250000 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --profile-range > gpurun_out/ncu_r1b.log 2>&1; echo "ncu_exit=$?"
tail -c 600 gpurun_out/t8.log; cat gpurun_out/ubench.log
