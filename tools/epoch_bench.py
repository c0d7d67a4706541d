#!/usr/bin/env python
"""Batch loop with per-batch host reads (GNNAETrainer.step + D2H of latent / reconstruction, the shape of reference
utils/train.py:51-120) against GNNAETrainer.run_epoch (one synchronisation per epoch).  GPU box: python tools/epoch_bench.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_jet_autoencoder_b200 import GNNAETrainer, synthetic_jets
from gnn_jet_autoencoder_b200.config import DEFAULT_ARCH, build_models

N, B, NB = 30, 4096, 40
enc, dec = build_models(N, DEFAULT_ARCH, device="cuda:0", precision="bf16", seed=0)
tr = GNNAETrainer(enc, dec, batch_size=B)
batches = [torch.from_numpy(synthetic_jets(B, N, seed=100 + i)).pin_memory() for i in range(NB)]
for x in batches[:3]:
    tr.step(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
tot, keep = 0.0, []
for x in batches:                      # reference-shaped loop: loss .item() and output copies after every batch
    tot += tr.step(x)
    keep.append((tr.latent.cpu(), tr.recon.cpu()))
loop_s = time.perf_counter() - t0
tr.run_epoch(batches)                  # the first epoch allocates the pinned epoch buffers
t0 = time.perf_counter()
avg, rec, tgt, lat = tr.run_epoch(batches)
epoch_s = time.perf_counter() - t0
t0 = time.perf_counter()
vavg, _, _, _ = tr.run_epoch(batches, is_train=False)
val_s = time.perf_counter() - t0
print(f"per-batch sync loop : {NB * B / loop_s:10.0f} jets/s   ({loop_s / NB * 1e3:.2f} ms per batch)")
print(f"run_epoch (train)   : {NB * B / epoch_s:10.0f} jets/s   ({epoch_s / NB * 1e3:.2f} ms per batch), avg loss {avg:.4f}")
print(f"run_epoch (validate): {NB * B / val_s:10.0f} jets/s   ({val_s / NB * 1e3:.2f} ms per batch), avg loss {vavg:.4f}")
