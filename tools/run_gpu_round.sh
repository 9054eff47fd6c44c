for m in 0 2 3 1 0 2 3 1; do
PMB_TICA_BARRIER=$m python bench.py --frames-per-gpu 1250000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench51_m$m.log 2> gpurun_out/bench51_m$m.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench51_m$m.log") if l.startswith("{")][-1])
print("mode $m", round(d["ms_per_step"],2), d["tica_phase_cycles"][:6], "tica_solve_ms", round(d["stages_ms"]["tica_solve"],2), "GHz", round(d["tica_phase_cycles"][5]/d["stages_ms"]["tica_solve"]/1e6,2))
PY
done
