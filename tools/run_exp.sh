timeout 300 python -m pytest tests -m gpu -x -q -k "msm or its or pipeline_small" > gpurun_out/t72.log 2>&1; echo "pytest_exit=$?"; tail -n 3 gpurun_out/t72.log
python bench.py --frames-per-gpu 1250000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench72.log 2> gpurun_out/bench72.err; echo "bench=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench72.log") if l.startswith("{")][-1])
c=d["mle_phase_cycles"]; it=c[0]
print("iters", it, "per iteration: compute+store", c[1]/it, "(row product", c[7]/it, ") gather own", c[2]/it, "rest", c[3]/it, "(wait all + norm", c[4]/it, ") mle ms", d["stages_ms"]["mle"], d["timescales"][:3])
PY
