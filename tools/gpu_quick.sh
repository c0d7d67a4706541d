#!/bin/bash
# quick bf16 check + bench (run on the GPU box)
mkdir -p gpurun_out
T="python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider"
timeout 600 $T -k "bf16 and (mp_step or modules or trainer or full_size or n150 or permutation)" > gpurun_out/quick.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/quick.log
for extra in "$@" ""; do
  env $extra python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench.log 2>&1
  python - "$extra" <<'PY'
import json, sys
for ln in open("gpurun_out/quick_bench.log"):
    if ln.startswith("{"):
        d = json.loads(ln); print(sys.argv[1] or "default", "jets/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "bwd us", round(d["roofline"]["us_per_launch"], 1), "frac", round(d["step_roofline"]["frac"], 4))
        break
else:
    print(open("gpurun_out/quick_bench.log").read()[-800:])
PY
done
