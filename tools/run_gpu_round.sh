# One GPU round on a fresh box (run from the repo root through gpurun): parity tests, smoke, the three bench lines.
set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest_exit=$?"; tail -n 4 gpurun_out/tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke=$?"
python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench=$?"
python bench.py --frames-per-gpu 1250000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small.log 2> gpurun_out/bench_small.err; echo "bench_small=$?"
python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "ref=$?"
