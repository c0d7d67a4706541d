"""The benchmark architecture (reference CLI defaults, utils/argparse_utils.py:81-119,141-179) and the
algorithmic FLOP count used for the tensor-core roofline (SURVEY.md 8.d / BASELINE.md 3)."""
from __future__ import annotations

DEFAULT_ARCH = dict(
    edge_sizes=[[32, 128, 64, 16]],
    node_sizes=[[16], [32], [8]],
    num_mps=3,
    alphas=0.2,
    latent_node_size=20,
    latent_map="mean",
    vec_dims=3,
)


def _broadcast(data, num):
    data = list(data)
    if len(data) < num:
        data = data + [data[-1]] * (num - len(data))
    return data[:num]


def edge_macs_per_row(node_sizes, edge_sizes, num_mps) -> int:
    """sum_t sum_k in_{t,k} * out_{t,k} with in_{t,0} = 2 H_t + 1: MACs per edge row of one GraphNet forward in the
    reference's dense formulation (graphnet.py:84,102-104)."""
    ns, es = _broadcast(node_sizes, num_mps), _broadcast(edge_sizes, num_mps)
    total = 0
    for t in range(num_mps):
        widths = [2 * ns[t][0] + 1] + list(es[t])
        total += sum(a * b for a, b in zip(widths[:-1], widths[1:]))
    return total


def train_flops_per_jet(num_nodes: int, arch=DEFAULT_ARCH) -> float:
    """Edge-MLP FLOPs of one training step per jet: encoder + decoder, forward + dgrad + wgrad = 3 x forward
    (node MLPs, 0.24 %, are excluded; padding and recomputation do not count)."""
    macs = edge_macs_per_row(arch["node_sizes"], arch["edge_sizes"], arch["num_mps"])
    fwd = 2.0 * num_nodes * num_nodes * macs      # one GraphNet forward
    return 3.0 * 2.0 * fwd


def build_models(num_nodes: int, arch=DEFAULT_ARCH, device=None, precision=None, seed: int = 0):
    """Encoder + Decoder of the benchmark architecture with default nn.Linear init under ``seed``."""
    import torch
    from .models import Decoder, Encoder
    torch.manual_seed(seed)
    enc = Encoder(num_nodes=num_nodes, input_node_size=arch["vec_dims"], latent_node_size=arch["latent_node_size"],
                  node_sizes=arch["node_sizes"], edge_sizes=arch["edge_sizes"], num_mps=arch["num_mps"],
                  alphas=arch["alphas"], latent_map=arch["latent_map"], device=device, precision=precision)
    dec = Decoder(num_nodes=num_nodes, latent_node_size=arch["latent_node_size"], output_node_size=arch["vec_dims"],
                  node_sizes=arch["node_sizes"], edge_sizes=arch["edge_sizes"], num_mps=arch["num_mps"],
                  alphas=arch["alphas"], latent_map=arch["latent_map"], device=device, precision=precision)
    return enc, dec
