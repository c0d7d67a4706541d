#!/bin/bash
# final validation + evidence of the round: GPU tests, smoke, the default bench line, ncu launch list, eager step profile
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/fin_tests.log 2>&1; echo tests rc=$?; tail -n 2 gpurun_out/fin_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -n 1
python bench.py > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; echo bench rc=$?
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/fin_bench.json") if l.startswith("{")][0])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "kernel", d["roofline"]["us_per_launch"], d["roofline"]["frac"], "step", d["step_roofline"]["frac"])
print({k: (round(v.get("value", 0)), round(v.get("roofline_frac", v.get("step_roofline_frac", 0)) or 0, 4)) if isinstance(v, dict) else v for k, v in d.get("configs", {}).items()})
PY
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-configs"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02f_launches.csv $CMD > gpurun_out/r02f_ncu_list.log 2>&1
echo "launch list rc=$?"
GJ_PDL=0 python tools/step_profile.py 30 4096 bf16 > gpurun_out/r02f_step_profile.txt 2>&1
head -3 gpurun_out/r02f_step_profile.txt
