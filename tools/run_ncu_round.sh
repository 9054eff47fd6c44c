set -x
B="python bench.py --frames-per-gpu 1250000 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --profile-range"
mkdir -p /tmp/ncu
$B > gpurun_out/plain_ncu.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"tica_solve_grid" -c 1 -f -o /tmp/ncu/tica $B > gpurun_out/ncu_tica.log 2>&1; echo "ncu_tica=$?"
ncu -i /tmp/ncu/tica.ncu-rep --page raw --csv > gpurun_out/ncu_r1g_tica_raw.csv 2>/dev/null
ncu -i /tmp/ncu/tica.ncu-rep --page source --csv > gpurun_out/ncu_r1g_tica_source.csv 2>/dev/null
ls -la gpurun_out/ncu_r1g*
