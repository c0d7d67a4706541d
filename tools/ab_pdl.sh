#!/bin/bash
# A/B of the programmatic-dependent-launch modes (GJ_PDL bit 0: node-level kernels, bit 1: edge kernels) on one box, config 2
run() { env GJ_PDL=$1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs $2 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('GJ_PDL=$1 $2', 'jets/s', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'bwd us', round(d['roofline']['us_per_launch'],1))"; }
run 0; run 1; run 2; run 3; run 0 --no-graph; run 3 --no-graph
