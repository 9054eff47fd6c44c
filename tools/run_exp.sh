set -x
bash tools/run_gpu_round.sh
bash tools/run_ncu_round.sh
