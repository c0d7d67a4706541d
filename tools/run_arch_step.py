"""Two plain training steps of a config-5 style architecture (no CUDA graph, no profiler) -- the command ncu captures.
GPU box: python tools/run_arch_step.py [H] [num_mps] [N] [B]   (H = 0: the default architecture)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_jet_autoencoder_b200 import GNNAETrainer, synthetic_jets
from gnn_jet_autoencoder_b200.config import DEFAULT_ARCH, build_models
H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
M = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
B = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
arch = DEFAULT_ARCH if H == 0 else dict(DEFAULT_ARCH, edge_sizes=[[H, H]], node_sizes=[[H]], num_mps=M, latent_node_size=8)
enc, dec = build_models(N, arch, device="cuda:0", precision="bf16", seed=0)
tr = GNNAETrainer(enc, dec, batch_size=B, use_cuda_graph=False)
x = torch.from_numpy(synthetic_jets(B, N, seed=1234)).pin_memory()
for _ in range(2):
    print("loss", tr.step(x))
torch.cuda.synchronize()
