mkdir -p /tmp/ncu
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"kmeans_tc_kernel" -s 4 -c 1 -f -o /tmp/ncu/km5 python bench.py --config C5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_km_c5.log 2>&1; echo "ncu=$?"
ncu -i /tmp/ncu/km5.ncu-rep --page raw --csv > gpurun_out/ncu_km_c5_raw.csv 2>/dev/null
cat gpurun_out/ncu_km_c5_raw.csv | python tools/ncu_summary.py
tail -3 gpurun_out/ncu_km_c5.log | cut -c1-300
