#!/usr/bin/env python
"""Anomaly-score chamfer distance: gnn_jet_autoencoder_b200.anomaly.chamfer against the reference formulation
(anomaly_detection.py:482-488 written out with torch broadcasting) on the same GPU.  GPU box: python tools/anomaly_bench.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_jet_autoencoder_b200 import anomaly

dev = torch.device("cuda", 0)
B, N, D = 65536, 30, 3
g = torch.Generator(device="cpu").manual_seed(0)
p, q = torch.randn(B, N, D, generator=g).to(dev), torch.randn(B, N, D, generator=g).to(dev)


def ref(p, q):
    diffs = torch.unsqueeze(p, -2) - torch.unsqueeze(q, -3)
    dist = torch.norm(diffs, dim=-1)
    return torch.min(dist, dim=-1).values + torch.min(dist, dim=-2).values


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


t_ours, a = timeit(lambda: anomaly.chamfer(p, q))
t_ref, b = timeit(lambda: torch.cat([ref(p[i:i + 4096], q[i:i + 4096]) for i in range(0, B, 4096)]))
print(f"B={B} N={N} D={D}: ours {t_ours * 1e3:.3f} ms ({B / t_ours / 1e6:.1f} M jets/s), torch broadcast form {t_ref * 1e3:.3f} ms "
      f"({B / t_ref / 1e6:.1f} M jets/s), max |diff| {float((a - b).abs().max()):.2e}")

# ---- Hungarian scores: device assignment solver against the reference's per-jet scipy loop (host) ----
import numpy as np
from scipy import optimize
Bh = 4096
ph, qh = p[:Bh], q[:Bh]
t_ours, a = timeit(lambda: anomaly.hungarian(ph, qh), reps=5)
t0 = time.perf_counter()
cost = torch.cdist(ph, qh).cpu().numpy()
matching = [optimize.linear_sum_assignment(cost[i])[1] for i in range(len(cost))]
p_shuffle = torch.stack([ph[i, torch.from_numpy(matching[i]).to(dev)] for i in range(Bh)])
b = ((p_shuffle - qh) ** 2).sum(-1)
torch.cuda.synchronize()
t_ref = time.perf_counter() - t0
print(f"hungarian, B={Bh} N={N}: ours {t_ours * 1e3:.2f} ms ({Bh / t_ours / 1e3:.0f} k jets/s), cdist + scipy loop + per-jet gather "
      f"{t_ref * 1e3:.0f} ms ({Bh / t_ref / 1e3:.1f} k jets/s), max |diff| {float((a - b).abs().max()):.2e}")
