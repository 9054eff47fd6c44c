# Experiment round (edit freely)
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu -k "kmeans or gram or assign or cluster or c5_assignment or nan_and_inf" > gpurun_out/t_km.log 2>&1; echo "pytest=$?"; tail -n 6 gpurun_out/t_km.log
python tools/km_bench.py c4 > gpurun_out/km_plain.log 2>&1; echo "km_plain=$?"; cat gpurun_out/km_plain.log
PMB_LIB=build_exp/libpmb200_kmprof.so python tools/km_bench.py c4 > gpurun_out/km_prof.log 2>&1; echo "km_prof=$?"; cat gpurun_out/km_prof.log
python tools/gram_bench.py 10000000 5 > gpurun_out/gram_bench.log 2>&1; echo "gb=$?"; cat gpurun_out/gram_bench.log
