"""Per-kernel device-time breakdown of one GNNAETrainer step (torch profiler, no CUDA graph).
Usage on the GPU box: python tools/step_profile.py [N] [B] [bf16|fp32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from gnn_jet_autoencoder_b200 import GNNAETrainer, synthetic_jets
from gnn_jet_autoencoder_b200.config import DEFAULT_ARCH, build_models
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
PREC = sys.argv[3] if len(sys.argv) > 3 else "bf16"
enc, dec = build_models(N, DEFAULT_ARCH, device="cuda:0", precision=PREC, seed=0)
tr = GNNAETrainer(enc, dec, batch_size=B, use_cuda_graph=False)
x = torch.from_numpy(synthetic_jets(B, N, seed=1234)).pin_memory()
for _ in range(3): tr.step(x)
torch.cuda.synchronize()
steps = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps): tr.step(x)
    torch.cuda.synchronize()
tot = sum(e.device_time_total for e in prof.key_averages())
print(f"N={N} B={B} {PREC}: {tot / steps:.1f} us of kernel time per step")
for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if ev.device_time_total <= 0: continue
    print(f"{ev.device_time_total / steps:10.1f} us/step {100 * ev.device_time_total / tot:5.1f}%  x{ev.count / steps:<5g} avg {ev.device_time_total / ev.count:8.1f} us  {ev.key[:80]}")
