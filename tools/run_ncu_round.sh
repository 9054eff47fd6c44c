set -x
B="python bench.py --frames-per-gpu 1250000 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --profile-range"
$B > gpurun_out/plain_ncu.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:kmeans_tc_kernel -s 10 -c 1 -f -o gpurun_out/prof_kmeans_tc2 $B > gpurun_out/ncu_kmeans.log 2>&1; echo "ncu1=$?"
