set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t46.log 2>&1; echo "pytest_exit=$?"; tail -n 4 gpurun_out/t46.log
python bench.py > gpurun_out/bench46_default.log 2> gpurun_out/bench46_default.err; echo "bench=$?"
python bench.py --frames-per-gpu 1250000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench46_small.log 2> gpurun_out/bench46_small.err; echo "bench_small=$?"
python bench.py --impl reference > gpurun_out/bench46_ref.log 2>&1; echo "ref=$?"
