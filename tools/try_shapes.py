import os, sys
sys.path.insert(0, "/root/repo")
import torch
from gnn_jet_autoencoder_b200 import GraphNet
for prec in ("fp32", "bf16"):
    for H in (64, 128, 256):
        for N in (30, 150):
            try:
                g = GraphNet(N, 3, 8, [[H]], [[H, H]], 2, alphas=0.2, device="cuda:0", precision=prec)
                x = torch.randn(4, N, 3, device="cuda:0", requires_grad=True)
                y = g(x); y.sum().backward(); torch.cuda.synchronize()
                print(prec, H, N, "ok", tuple(y.shape))
            except Exception as ex:
                print(prec, H, N, type(ex).__name__, str(ex)[:150])
