"""Every kernel launch of one GNNAETrainer step in order, with its duration and grid (torch profiler).  GPU box:
python tools/launch_trace.py [H] [num_mps] [N] [B]    (H = 0: the default architecture)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from gnn_jet_autoencoder_b200 import GNNAETrainer, synthetic_jets
from gnn_jet_autoencoder_b200.config import DEFAULT_ARCH, build_models
H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
M = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
B = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
arch = DEFAULT_ARCH if H == 0 else dict(DEFAULT_ARCH, edge_sizes=[[H, H]], node_sizes=[[H]], num_mps=M, latent_node_size=8)
enc, dec = build_models(N, arch, device="cuda:0", precision="bf16", seed=0)
tr = GNNAETrainer(enc, dec, batch_size=B, use_cuda_graph=False)
x = torch.from_numpy(synthetic_jets(B, N, seed=1234)).pin_memory()
for _ in range(3): tr.step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(x)
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_time_total > 0], key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    print(f"{e.time_range.start - t0:10.1f} us  {e.device_time_total:8.1f} us  {e.name[:70]}")
