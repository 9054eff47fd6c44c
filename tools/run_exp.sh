set -x
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -x -q -m gpu -k "c5 or msm or mle or two_state or implied or enhanced or bayes" > gpurun_out/t_mle.log 2>&1; echo "pytest=$?"; tail -n 5 gpurun_out/t_mle.log
python bench.py --config C5 --steps 2 --warmup 1 > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err; echo "c5=$?"; tail -c 300 gpurun_out/bench_c5.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_c5.log") if l.startswith("{")][-1])
print(round(d["value"]/1e6,2), round(d["ms_per_step"],1), {k: round(v,1) for k,v in d["stages_ms"].items()}, d["roofline"]["frac"], d["roofline"]["all"]["mle"], d["properties"])
PY
