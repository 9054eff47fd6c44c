set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_analysis.py -x -q -m gpu -k "gram or tica or pipeline_small or host_buffer or pcca" > gpurun_out/t_gram.log 2>&1; echo "pytest=$?"; tail -n 5 gpurun_out/t_gram.log
timeout 300 python tools/gram_bench.py 10000000 5 > gpurun_out/gram_bench.log 2>&1; echo "gb=$?"; cat gpurun_out/gram_bench.log
timeout 300 python tools/gram_accuracy.py > gpurun_out/gram_acc.log 2>&1; echo "acc=$?"; grep "impl=5" gpurun_out/gram_acc.log
