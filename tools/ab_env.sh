#!/bin/bash
# A/B on one box: bench.py (config 2, short) with and without an environment switch.   usage: tools/ab_env.sh VAR
for i in 1 2; do
for v in 0 1; do
env $1=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('$1=$v', 'jets/s', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'bwd us', round(d['roofline']['us_per_launch'],1))"
done; done
