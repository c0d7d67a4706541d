#!/bin/bash
mkdir -p gpurun_out
a=$(GJ_NODE_FWD_V1=1 python tools/node_fwd_ab.py 2>&1 | tail -n 1); b=$(python tools/node_fwd_ab.py 2>&1 | tail -n 1)
echo "v1  $a"; echo "new $b"; [ "$a" == "$b" ] && echo "BITWISE IDENTICAL" || echo "DIFFERENT"
python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/nw_tests.log 2>&1; echo tests rc=$?; tail -n 3 gpurun_out/nw_tests.log
for v in 1 0 1 0; do
GJ_NODE_FWD_V1=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('GJ_NODE_FWD_V1=$v', 'jets/s', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'bwd us', round(d['roofline']['us_per_launch'],1))"
done
python tools/step_profile.py 30 4096 bf16 2>/dev/null | grep -i "node_p\|us of kernel" | head -14
