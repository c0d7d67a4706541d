"""Launches one message-passing step a few times at bench size (for ncu).  Usage: python tools/run_mp.py [fwd|all] [N] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_jet_autoencoder_b200 import ops
what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
H, edge, node = 16, [32, 128, 64, 16], [16, 32]
npar = sum(o * i + o for i, o in zip([2 * H + 1] + edge[:-1], edge)) + sum(o * i + o for i, o in zip([edge[-1] + H] + node[:-1], node))
torch.manual_seed(0)
flat = (torch.rand(npar, device="cuda") - 0.5) * 0.3
h = torch.randn(B, N, H, device="cuda") * 0.5
dy = torch.randn(B, N, node[-1], device="cuda")
args = (N, H, edge, node, 0.2, 0, ops.PRECISIONS["bf16"])
for _ in range(4):
    y, e = torch.ops.gnnjet.mp_step_fwd(h, flat, *args)
    if what == "all":
        torch.ops.gnnjet.mp_step_bwd(h, e, flat, dy, *args)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
