mkdir -p /tmp/ncu
timeout 400 ncu --set full --clock-control none -k regex:"mle_grid_kernel" -c 1 -f -o /tmp/ncu/mle5 python bench.py --config C5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_mle_c5.log 2>&1; echo "ncu=$?"
ncu -i /tmp/ncu/mle5.ncu-rep --page raw --csv > gpurun_out/ncu_mle_c5_raw.csv 2>/dev/null
cat gpurun_out/ncu_mle_c5_raw.csv | python tools/ncu_summary.py
