set -x
B="python bench.py --frames-per-gpu 1250000 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --profile-range"
mkdir -p /tmp/ncu
$B > gpurun_out/plain_ncu.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"mle_grid_kernel|lanczos_kernel" -c 2 -f -o /tmp/ncu/lat $B > gpurun_out/ncu_lat2.log 2>&1; echo "ncu_lat=$?"
ncu -i /tmp/ncu/lat.ncu-rep --page raw --csv > gpurun_out/ncu_r1_mle_lanczos_raw.csv 2>/dev/null
ncu -i /tmp/ncu/lat.ncu-rep --page source --csv -k regex:mle_grid_kernel > gpurun_out/ncu_r1_mle_source.csv 2>/dev/null
ncu -i /tmp/ncu/lat.ncu-rep --page source --csv -k regex:lanczos_kernel > gpurun_out/ncu_r1_lanczos_source.csv 2>/dev/null
ls -la gpurun_out/ncu_r1_mle* gpurun_out/ncu_r1_lanczos*
