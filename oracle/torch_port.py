"""CPU PyTorch restatement of the reference's GNNAE train step.  TEST / BASELINE INFRASTRUCTURE ONLY.

Purpose: the reference is Python and lives only in the build container, so it cannot be timed on the GPU
box.  This file restates its op sequence with stock ``torch`` CPU ops (same materialised (B,N,N,2H+1) pair
tensor, same Linear + leaky_relu chains, autograd backward, two ``torch.optim.Adam``), so that
``bench.py --impl reference`` / ``cpu_baseline`` time the same work the reference does on the host cores
("kind": "port").  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs import it;
the product never does.  ``tests/test_oracle.py`` pins it against the golden vectors produced by the real
reference (oracle/gen_golden.py).

Cited reference lines are relative to the reference checkout.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LOCAL_MIX = ("local", "local_mix", "node", "node_mix")      # models/const.py:1
GLOBAL_MIX = ("global", "global_mix", "graph", "graph_mix")  # models/const.py:2
EPS = 1e-16                                                   # utils/const.py:5


def _adjust(data, num):                                       # graphnet.py:305-311
    if isinstance(data, (list, tuple)):
        data = list(data)
        if len(data) < num:
            data = data + [data[-1]] * (num - len(data))
    else:
        data = [data] * num
    return data[:num]


def _stack(params, prefix, kind, t):
    out, k = [], 0
    while f"{prefix}{kind}.{t}.{k}.weight" in params:
        out.append((params[f"{prefix}{kind}.{t}.{k}.weight"], params[f"{prefix}{kind}.{t}.{k}.bias"]))
        k += 1
    return out


def graphnet_forward(x, params, prefix, num_mps, input_node_size, node_sizes, alphas, metric="euclidean"):
    """graphnet.py:136-171 with the reference's op sequence: repeat/repeat/cat (_getA :211-222),
    Linear+leaky_relu per edge layer (:284-286), sum over j and cat((e, x)) (:243-246), node chain (:266-268)."""
    node_sizes = _adjust(node_sizes, num_mps)
    alphas = _adjust(alphas, num_mps)
    B, N = x.shape[0], x.shape[1]
    x = F.pad(x, (0, node_sizes[0][0] - input_node_size))
    for t in range(num_mps):
        H = x.shape[-1]
        x1 = x.repeat(1, 1, N).view(B, N * N, H)
        x2 = x.repeat(1, N, 1)
        diff = x2 - x1 + EPS
        if H == 4 and str(metric).lower() == "minkowskian":      # graphnet.py:155,320-323
            d = (2 * diff[..., 0] ** 2 - (diff ** 2).sum(-1)).unsqueeze(-1)
        else:
            d = (diff ** 2).sum(-1).unsqueeze(-1)
        A = torch.cat((x1, x2, d), 2).view(B, N, N, 2 * H + 1)
        for w, b in _stack(params, prefix, "edge_net", t):
            A = F.leaky_relu(F.linear(A, w, b), negative_slope=alphas[t])
        x = torch.cat((A.sum(dim=-2), x), dim=-1)
        for w, b in _stack(params, prefix, "node_net", t):
            x = F.leaky_relu(F.linear(x, w, b), negative_slope=alphas[t])
    return x


def encoder_forward(x, params, cfg, metric="euclidean"):
    """encoder.py:133-171."""
    lm_raw = cfg["latent_map"]
    y = graphnet_forward(x, params, "encoder.", cfg["num_mps"], cfg["input_node_size"], cfg["node_sizes"],
                         cfg["alphas"], metric)
    lm = lm_raw.lower().replace(" ", "_")
    B = x.shape[0]
    if lm == "max":
        return torch.amax(y, dim=-2)
    if lm == "min":
        return torch.amin(y, dim=-2)
    if lm in GLOBAL_MIX:
        return F.linear(y.reshape(B, -1), params["mix_layer.weight"])
    if lm in LOCAL_MIX:
        return F.linear(y, params["mix_layer.weight"], params["mix_layer.bias"]).reshape(B, -1)
    return y.mean(dim=-2)


def decoder_forward(z, params, cfg, metric="euclidean"):
    """decoder.py:119-136."""
    N, h0 = cfg["num_nodes"], cfg["node_sizes"][0][0]
    if cfg["latent_map"].lower().replace(" ", "_") in LOCAL_MIX:
        x = F.linear(z.view(-1, N, cfg["latent_node_size"]), params["linear.weight"], params["linear.bias"])
    else:
        x = F.linear(z, params["linear.weight"], params["linear.bias"]).view(-1, N, h0)
    y = graphnet_forward(x, params, "decoder.", cfg["num_mps"], h0, cfg["node_sizes"], cfg["alphas"], metric)
    return torch.tanh(y) if cfg.get("normalize_output", False) else y


def chamfer_intended(p, q, norm_choice="cartesian", jet_features_weight=1.0):
    """chamfer_loss.py:26-41 + distance_sq.py:46-77 (the value the reference computes before discarding it)."""
    B, N, D = p.shape
    M = q.shape[1]
    pr = p.repeat(1, 1, M).view(B, N * M, D)
    qr = q.repeat(1, N, 1)
    diff = pr - qr
    mink_name = str(norm_choice).lower() in ("minkowskian", "polar")
    n_mink, n_cart = (lambda v: 2 * v[..., 0] ** 2 - (v ** 2).sum(-1)), (lambda v: (v ** 2).sum(-1))
    # 3-vectors force cartesian in pairwise_distance_sq only (distance_sq.py:43-44); the jet term keeps the choice (:40)
    dist = (n_mink if (mink_name and D != 3) else n_cart)(diff).view(B, N, M)
    cham = torch.min(dist, dim=-1).values.sum() + torch.min(dist, dim=-2).values.sum()
    jet = (n_mink if mink_name else n_cart)(p.sum(dim=-2) - q.sum(dim=-2)).sum()
    return cham + jet_features_weight * jet, cham, jet


def make_params(np_params, dtype=torch.float32, requires_grad=True):
    return {k: torch.tensor(v, dtype=dtype, requires_grad=requires_grad) for k, v in np_params.items()}


def loss_fn(x, enc_p, dec_p, enc_cfg, dec_cfg, *, metric="euclidean", loss_norm_choice="cartesian",
            jet_features_weight=1.0, l1_lambda=1e-8, l2_lambda=0.0):
    """utils/train.py:52-76,330-385 (Chamfer branch, intended value) incl. the L1/L2 regularisers."""
    z = encoder_forward(x, enc_p, enc_cfg, metric)
    y = decoder_forward(z, dec_p, dec_cfg, metric)
    loss, cham, jet = chamfer_intended(y, x, loss_norm_choice, jet_features_weight)
    if l1_lambda > 0:
        loss = loss + l1_lambda * (sum(p.abs().sum() for p in enc_p.values()) + sum(p.abs().sum() for p in dec_p.values()))
    if l2_lambda > 0:
        loss = loss + l2_lambda * (sum(p.pow(2).sum() for p in enc_p.values()) + sum(p.pow(2).sum() for p in dec_p.values()))
    return loss, z, y


class TorchTrainStep:
    """utils/train.py:81-85 with the optimisers of utils/initialize.py:152-153 (two Adams, lr 1e-5)."""

    def __init__(self, enc_p, dec_p, enc_cfg, dec_cfg, lr=1e-5, **kw):
        self.enc_p, self.dec_p, self.enc_cfg, self.dec_cfg, self.kw = enc_p, dec_p, enc_cfg, dec_cfg, kw
        self.opt_e = torch.optim.Adam(list(enc_p.values()), lr)
        self.opt_d = torch.optim.Adam(list(dec_p.values()), lr)

    def step(self, x):
        loss, _, _ = loss_fn(x, self.enc_p, self.dec_p, self.enc_cfg, self.dec_cfg, **self.kw)
        value = loss.item()                      # the reference syncs on .item() every batch (train.py:77)
        self.opt_e.zero_grad()
        self.opt_d.zero_grad()
        loss.backward()
        self.opt_e.step()
        self.opt_d.step()
        return value
