# ncu launch list + one --set full capture of the hot kernels, after the same command ran clean without ncu.
# Reports stay on the box (/tmp/ncu); only CSV pages come back through gpurun_out/ (64 MiB limit).
set -x
B="python bench.py --frames-per-gpu 1250000 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --profile-range"
mkdir -p /tmp/ncu
$B > gpurun_out/plain_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1; echo "ncu_list=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"featurize_kernel|col_moments_v4|gram_tc_kernel|tica_solve_grid|project_warp|count_global" -c 8 -f -o /tmp/ncu/hot $B > gpurun_out/ncu_hot.log 2>&1; echo "ncu_hot=$?"
ncu -i /tmp/ncu/hot.ncu-rep --page raw --csv > gpurun_out/ncu_hot_raw.csv 2>/dev/null
ls -la gpurun_out/ncu_hot_raw.csv gpurun_out/launches.csv
