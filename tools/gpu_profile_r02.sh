#!/bin/bash
# Round-2 ncu / timing evidence (run on the GPU box; summaries are then copied from gpurun_out/ into profiles/).
# Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-configs"
$CMD > gpurun_out/r02_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
echo "launch list rc=$?"
python tools/run_mp.py all > gpurun_out/r02_plain_mp.log 2>&1 || { echo "run_mp failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:edge_bwd2 -s 2 -c 1 -o gpurun_out/r02_bwd2 -f python tools/run_mp.py all > gpurun_out/r02_ncu_b.log 2>&1
echo "bwd2 capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:edge_fwd2 -s 2 -c 1 -o gpurun_out/r02_fwd2 -f python tools/run_mp.py all > gpurun_out/r02_ncu_f.log 2>&1
echo "fwd2 capture rc=$?"
python tools/trace_bwd2.py > gpurun_out/r02_bwd2_trace.txt 2>&1
python tools/step_profile.py 30 4096 bf16 > gpurun_out/r02_step_profile.txt 2>&1
[ "$1" = quick ] || { python tools/sweep_bench.py 30 2048 > gpurun_out/r02_sweep.txt 2>&1; cat gpurun_out/r02_sweep.txt; }
python - > gpurun_out/r02_deterministic.txt 2>&1 <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
from gnn_jet_autoencoder_b200 import ops
N, B, H, edge, node = 30, 4096, 16, [32, 128, 64, 16], [16, 32]
npar = sum(o * i + o for i, o in zip([2 * H + 1] + edge[:-1], edge)) + sum(o * i + o for i, o in zip([edge[-1] + H] + node[:-1], node))
torch.manual_seed(0)
flat = (torch.rand(npar, device="cuda") - 0.5) * 0.3
h = torch.randn(B, N, H, device="cuda") * 0.5
dy = torch.randn(B, N, node[-1], device="cuda")
args = (N, H, edge, node, 0.2, 0, ops.PRECISIONS["bf16"])
y, e = torch.ops.gnnjet.mp_step_fwd(h, flat, *args)
for det in (False, True):
    ops.set_deterministic(det)
    for _ in range(3): torch.ops.gnnjet.mp_step_bwd(h, e, flat, dy, *args)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): torch.ops.gnnjet.mp_step_bwd(h, e, flat, dy, *args)
    b.record(); torch.cuda.synchronize()
    print(f"gj_mp_step_bwd (N=30, B=4096, default widths, all launches), deterministic={det}: {a.elapsed_time(b) * 100:.1f} us per call")
PY
cat gpurun_out/r02_deterministic.txt
ls -la gpurun_out/r02_*.ncu-rep
