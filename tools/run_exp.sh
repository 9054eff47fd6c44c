set -x
for a in "1000 6" "2000 10" "3000 10" "5000 20"; do python tools/eig_bench.py $a 4; done > gpurun_out/eig_diag.log 2>&1
timeout 900 python -m pytest tests/ -x -q -m gpu -k "eig or c5 or implied or lanczos or enhanced" > gpurun_out/t_eig.log 2>&1; echo "pytest=$?"; tail -n 5 gpurun_out/t_eig.log
python bench.py --config C5 --steps 2 --warmup 1 > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err; echo "c5=$?"; tail -c 300 gpurun_out/bench_c5.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_c5.log") if l.startswith("{")][-1])
print(round(d["value"]/1e6,2), round(d["ms_per_step"],1), {k: round(v,1) for k,v in d["stages_ms"].items()}, d["roofline"]["frac"], d["roofline"]["all"].get("mle"), d["properties"])
PY
cat gpurun_out/eig_diag.log
