timeout 600 python -m pytest tests -m gpu -x -q -k "kmeans or pipeline_small or smoke or abi" > gpurun_out/t67.log 2>&1; echo "pytest_exit=$?"; tail -n 3 gpurun_out/t67.log
python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench67_default.log 2> gpurun_out/bench67_default.err; echo "bench=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench67_default.log") if l.startswith("{")][-1])
print(round(d["value"]/1e6,2), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]/1e6,2), d["kmeans_role_cycles"][:4], {k: round(v,2) for k,v in d["stages_ms"].items()})
PY
