#!/bin/bash
# one development iteration on the GPU box: bf16 parity subset, short bench (config 2), stage trace of the backward edge kernel
# usage: tools/gpu_iter.sh <tag>
TAG=${1:-it}
mkdir -p gpurun_out
T="python -m pytest tests/test_gpu_parity.py -q -x -m gpu -p no:cacheprovider"
timeout 600 $T -k "(bf16 and (mp_step or modules or trainer or full_size or n150)) or saved or bench_hooks" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/${TAG}_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - gpurun_out/${TAG}_bench.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
    print("jets/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "bwd us", round(d["roofline"]["us_per_launch"], 1), "frac", round(d["roofline"]["frac"], 4), "step frac", round(d["step_roofline"]["frac"], 4))
except Exception as e:
    print("bench failed", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
python tools/trace_bwd2.py > gpurun_out/${TAG}_trace.txt 2>&1; cat gpurun_out/${TAG}_trace.txt | tail -18
