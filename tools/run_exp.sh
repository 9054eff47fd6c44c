python tools/km_bench.py c4 2>&1 | tail -5
python tools/km_bench.py c5 2>&1 | tail -5
timeout 900 python -m pytest tests -x -q -m gpu -k "kmeans or cluster or lloyd or assign or c5 or tica" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_x.log 2> gpurun_out/bench_x.err; tail -c 300 gpurun_out/bench_x.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_x.log") if l.startswith("{")][-1])
print(round(d["value"]/1e6,2), round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["stages_ms"].items()}, d["e2e"]["ms_per_step"])
PY
