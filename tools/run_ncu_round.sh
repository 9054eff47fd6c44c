# ncu launch list (10 M frames, the bench size) + --set full captures of the hot kernels, each after the same command
# ran clean without ncu.  Reports stay on the box (/tmp/ncu); only CSV pages come back through gpurun_out/ (64 MiB limit).
set -x
B="python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-extras --profile-range"
mkdir -p /tmp/ncu
$B > gpurun_out/plain_ncu.log 2>&1; echo "plain=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_10M.csv $B > gpurun_out/ncu_launches.log 2>&1; echo "ncu_list=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"featurize_kernel|col_moments_v4|gram_h_kernel|tica_solve_grid|project_warp|count_global|mle_grid|lanczos" -c 12 -f -o /tmp/ncu/hot $B > gpurun_out/ncu_hot.log 2>&1; echo "ncu_hot=$?"
ncu -i /tmp/ncu/hot.ncu-rep --page raw --csv > gpurun_out/ncu_hot_raw.csv 2>/dev/null
# k-means: launch 12 of the tensor kernel = a warm-hint Lloyd iteration with accumulation
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"kmeans_tc_kernel" -s 11 -c 1 -f -o /tmp/ncu/km $B > gpurun_out/ncu_km.log 2>&1; echo "ncu_km=$?"
ncu -i /tmp/ncu/km.ncu-rep --page raw --csv > gpurun_out/ncu_km_raw.csv 2>/dev/null
ncu -i /tmp/ncu/km.ncu-rep --page source --csv > gpurun_out/ncu_km_src.csv 2>/dev/null
python tools/ncu_hotspots.py gpurun_out/ncu_km_src.csv 30 > gpurun_out/ncu_km_hot.txt 2>&1
ncu -i /tmp/ncu/hot.ncu-rep --page source --csv --kernel-name regex:gram_h_kernel > gpurun_out/ncu_gh_src.csv 2>/dev/null
python tools/ncu_hotspots.py gpurun_out/ncu_gh_src.csv 30 > gpurun_out/ncu_gh_hot.txt 2>&1
rm -f gpurun_out/ncu_km_src.csv gpurun_out/ncu_gh_src.csv
ls -la gpurun_out/
