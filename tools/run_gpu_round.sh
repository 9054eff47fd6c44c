set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t65.log 2>&1; echo "pytest_exit=$?"; tail -n 4 gpurun_out/t65.log
python bench.py > gpurun_out/bench65_default.log 2> gpurun_out/bench65_default.err; echo "bench=$?"
python bench.py --frames-per-gpu 1250000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench65_small.log 2> gpurun_out/bench65_small.err; echo "bench_small=$?"
python bench.py --impl reference > gpurun_out/bench65_ref.log 2>&1; echo "ref=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke65.log 2>&1; echo "smoke=$?"; tail -n 2 gpurun_out/smoke65.log
python - <<'PY'
import json
for f in ("gpurun_out/bench65_default.log","gpurun_out/bench65_small.log"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["tica_phase_cycles"], {k: round(v,2) for k,v in d["stages_ms"].items()})
PY
B="python bench.py --frames-per-gpu 1250000 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --profile-range"
mkdir -p /tmp/ncu
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1g.csv $B > gpurun_out/ncu_launches.log 2>&1; echo "ncu_list=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"featurize_kernel|col_moments_v4|gram_tc_kernel|tica_solve_grid|project_warp|count_global" -c 8 -f -o /tmp/ncu/hot $B > gpurun_out/ncu_hot.log 2>&1; echo "ncu_hot=$?"
ncu -i /tmp/ncu/hot.ncu-rep --page raw --csv > gpurun_out/ncu_r1g_hot_raw.csv 2>/dev/null
ls -la gpurun_out/ncu_r1g_hot_raw.csv gpurun_out/launches_r1g.csv
