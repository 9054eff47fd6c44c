timeout 900 python -m pytest tests -m gpu -x -q -k "eig or msm or its or pipeline or lanczos or timescale" > gpurun_out/t70.log 2>&1; echo "pytest_exit=$?"; tail -n 6 gpurun_out/t70.log
python bench.py --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench70.log 2> gpurun_out/bench70.err; echo "bench=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench70.log") if l.startswith("{")][-1])
print(round(d["value"]/1e6,2), round(d["ms_per_step"],2), d["timescales"], {k: round(v,2) for k,v in d["stages_ms"].items()})
PY
