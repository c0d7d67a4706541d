#!/usr/bin/env python
"""A few points of BASELINE config 5 (num_mps x hidden width H with node_sizes=[[H]], edge_sizes=[[H,H]]) on one GPU:
jets/s of the full train step and the fraction of the sustained bf16 roofline.  These shapes run the first-generation
generic tensor-core kernels (the fused second-generation kernels are compiled for the default 32-128-64-16 edge network).
GPU box: python tools/sweep_bench.py [N] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_jet_autoencoder_b200 import GNNAETrainer, synthetic_jets
from gnn_jet_autoencoder_b200.config import DEFAULT_ARCH, build_models, train_flops_per_jet

N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
PEAK = 1375.0      # TFLOP/s, sustained bf16 (MEASURED_PEAKS.json on this pod)
x = torch.from_numpy(synthetic_jets(B, N, seed=1234)).pin_memory()
print(f"N={N} B={B}, bf16 mode, one GPU")
for num_mps, H, latent in [(3, 64, 8), (4, 64, 8), (6, 64, 8), (3, 128, 8), (4, 128, 32), (3, 256, 8)]:
    arch = dict(DEFAULT_ARCH, edge_sizes=[[H, H]], node_sizes=[[H]], num_mps=num_mps, latent_node_size=latent)
    try:
        enc, dec = build_models(N, arch, device="cuda:0", precision="bf16", seed=0)
        tr = GNNAETrainer(enc, dec, batch_size=B)
        for _ in range(3):
            tr.step(x)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        steps = 5
        for _ in range(steps):
            tr.load_batch(x); tr.compute_gradients(); tr.apply_gradients()
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / steps
        jets = B / (ms * 1e-3)
        tf = jets * train_flops_per_jet(N, arch) / 1e12
        print(f"  num_mps={num_mps} H={H:3d} latent={latent:2d}: {ms:8.2f} ms/step  {jets:10.0f} jets/s  {tf:7.1f} TFLOP/s = {100 * tf / PEAK:4.1f} % of {PEAK:.0f}")
        del tr, enc, dec
        torch.cuda.empty_cache()
    except Exception as ex:      # widths the kernels do not cover are reported, not hidden
        print(f"  num_mps={num_mps} H={H:3d} latent={latent:2d}: {type(ex).__name__}: {str(ex)[:120]}")
