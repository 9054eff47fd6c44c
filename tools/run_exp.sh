timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest_exit=$?"; tail -n 3 gpurun_out/tests_gpu.log
python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_default.log") if l.startswith("{")][-1])
print(round(d["value"]/1e6,2), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]/1e6,2), d["mle_iters"], {k: round(v,2) for k,v in d["stages_ms"].items()})
PY
