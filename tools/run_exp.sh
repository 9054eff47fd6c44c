timeout 1200 python -m pytest tests -x -q -m gpu -k "kmeans or cluster or lloyd or assign or config or c1 or c2 or c3 or c5 or discretize or pipeline or smoke" 2>&1 | tail -3
python bench.py --config C5 --steps 2 --warmup 1 > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err; echo "c5=$?"; tail -c 300 gpurun_out/bench_c5.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_c5.log") if l.startswith("{")][-1])
print(round(d["value"]/1e6,2), round(d["ms_per_step"],1), {k: round(v,1) for k,v in d["stages_ms"].items()}, d["roofline"]["frac"], d["roofline"]["all"].get("mle"), d["properties"])
PY
