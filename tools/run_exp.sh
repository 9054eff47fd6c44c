set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu -k "kmeans or assign or cluster or c5_assignment or nan_and_inf or pipeline_small" > gpurun_out/t_km.log 2>&1; echo "pytest=$?"; tail -n 5 gpurun_out/t_km.log
python tools/km_bench.py c4 > gpurun_out/km_plain.log 2>&1; echo "km_plain=$?"; cat gpurun_out/km_plain.log
