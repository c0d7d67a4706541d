#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/bt_tests.log 2>&1; echo tests rc=$?; tail -n 3 gpurun_out/bt_tests.log
for v in 0 1 0 1; do
GNNJET_BATCHED_LAUNCHES=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('BATCHED=$v', 'jets/s', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'bwd us', round(d['roofline']['us_per_launch'],1), 'launches', d['gpu_launches'])"
done
GJ_PDL=0 python tools/step_profile.py 30 4096 bf16 2>/dev/null | head -40
