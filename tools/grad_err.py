#!/usr/bin/env python
"""Per-parameter-tensor gradient error of the CUDA modules against the float64 numpy oracle for one golden case (GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import gnnae_oracle as O  # noqa: E402
from test_gpu_parity import CASES, DEV, build, make_input, rel  # noqa: E402
from gnn_jet_autoencoder_b200 import ChamferLoss  # noqa: E402

name, precision = sys.argv[1], sys.argv[2]
case = dict(CASES[name])
if len(sys.argv) > 3:      # other weights / inputs for the same architecture
    case["seed"] = int(sys.argv[3])
quiet = len(sys.argv) > 4
enc, dec, ep, dp = build(case, precision)
x = make_input(case)
loss_ref, z_ref, y_ref, eg, dg = O.loss_and_grads(x, ep, dp, case["enc"], case["dec"], metric=case["metric"],
                                                  loss_norm_choice=case["loss_norm_choice"],
                                                  jet_features_weight=case["jet_features_weight"], l1_lambda=case["l1_lambda"])
xt = torch.from_numpy(x).float().to(DEV)
z = enc(xt, metric=case["metric"])
y = dec(z, metric=case["metric"])
loss = ChamferLoss(case["loss_norm_choice"])(y, xt, jet_features_weight=case["jet_features_weight"])
(loss + case["l1_lambda"] * (enc.l1_norm() + dec.l1_norm())).backward()
print("latent", rel(z.detach().cpu().numpy(), z_ref), "recon", rel(y.detach().cpu().numpy(), y_ref))
tot_n = tot_d = 0.0
for tag, mod, ref in (("enc", enc, eg), ("dec", dec, dg)):
    for k, p in mod.named_parameters():
        g = p.grad.cpu().numpy().astype(np.float64)
        d = np.linalg.norm(g - ref[k]); n = np.linalg.norm(ref[k])
        tot_n += d * d; tot_d += n * n
        if not quiet: print(f"{tag}.{k:40s} |g| {n:10.3e}  abs err {d:10.3e}  rel {d / (n + 1e-300):9.2e}")
print("concatenated", (tot_n / tot_d) ** 0.5)
