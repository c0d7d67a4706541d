#!/bin/bash
# Runs the GPU checks in separate processes (a faulting kernel poisons only its own process) and leaves logs
# under gpurun_out/.  Usage on the GPU box: bash tools/gpu_check.sh [quick]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
T="python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider"
run umma $T -k umma
run mp_fp32 $T -k "mp_step and fp32"
run mp_bf16 $T -k "mp_step and bf16"
run small $T -k "chamfer or adam or linear or bad_arguments"
run modules_fp32 $T -k "modules and fp32"
run modules_bf16 $T -k "modules and bf16"
run trainer $T -k "trainer"
run props $T -k "permutation or full_size or n150 or module_path"
run smoke python -c "import __graft_entry__ as g; g.smoke()"
if [ "$1" != "quick" ]; then
run bench_fp32 python bench.py --steps 5 --warmup 3 --precision fp32 --no-cpu-baseline
run bench_bf16 python bench.py --steps 10 --warmup 3 --precision bf16
fi
