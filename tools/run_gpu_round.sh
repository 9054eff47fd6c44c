for i in 1 2 3; do
python bench.py --no-cpu-baseline > gpurun_out/bench63_$i.log 2> gpurun_out/bench63_$i.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench63_$i.log") if l.startswith("{")][-1])
print("run $i", round(d["value"]/1e6,2), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]/1e6,2), {k: round(v,2) for k,v in d["stages_ms"].items()})
PY
done
