"""Decoder: drop-in for reference models/decoder.py (latent -> node features -> GraphNet)."""
from __future__ import annotations

from typing import List, Optional, Union

import torch
import torch.nn as nn

from .. import ops
from .const import LOCAL_MIX
from .graphnet import Float32ParamsMixin, GraphNet, _default_device


class Decoder(Float32ParamsMixin, nn.Module):
    """latent (B, latent) or (B, N*latent) -> (B, N, output_node_size).  Same constructor / attributes /
    ``state_dict`` keys as the reference (decoder.py:12-117); see GraphNet for ``precision``."""

    def __init__(
        self,
        num_nodes: int,
        latent_node_size: int,
        output_node_size: int,
        node_sizes: List[List[int]],
        edge_sizes: List[List[int]],
        num_mps: int,
        alphas: Union[float, List[float]],
        dropout: float = 0.0,
        batch_norm: bool = False,
        latent_map: str = "mix",
        normalize_output: bool = False,
        device: Optional[torch.device] = None,
        dtype: Optional[torch.dtype] = None,
        precision: Optional[str] = None,
    ):
        super().__init__()
        self.device = torch.device(device) if device is not None else _default_device()
        self.dtype = dtype if dtype is not None else torch.float
        self.num_nodes = num_nodes
        self.latent_map = latent_map
        self.latent_node_size = latent_node_size
        self.output_node_size = output_node_size
        self.node_sizes = node_sizes
        self.edge_sizes = edge_sizes
        self.num_mps = num_mps
        self.normalize_output = normalize_output

        h0 = node_sizes[0][0]   # raw list, as decoder.py:90,96,104
        self._per_node = latent_map.lower().replace(" ", "_") in LOCAL_MIX
        self.linear = nn.Linear(latent_node_size, h0 if self._per_node else num_nodes * h0).to(self.device)
        self.decoder = GraphNet(num_nodes=num_nodes, input_node_size=h0, output_node_size=output_node_size,
                                node_sizes=node_sizes, edge_sizes=edge_sizes, num_mps=num_mps, alphas=alphas,
                                dropout=dropout, batch_norm=batch_norm, device=self.device, dtype=self.dtype,
                                precision=precision)

    def forward(self, x: torch.Tensor, metric="euclidean") -> torch.Tensor:
        dev = self.linear.weight.device
        if dev.type != "cuda":
            raise RuntimeError("Decoder.forward runs on sm_100a CUDA kernels only (there is no CPU fallback)")
        z = x.to(device=dev, dtype=torch.float32)
        h0 = self.node_sizes[0][0]
        if self._per_node:      # decoder.py:130-132
            h = ops.linear(z.reshape(-1, self.num_nodes, self.latent_node_size), self.linear.weight, self.linear.bias)
        else:                   # decoder.py:133-135
            h = ops.linear(z, self.linear.weight, self.linear.bias).view(-1, self.num_nodes, h0)
        y = self.decoder(h, metric=metric)
        if self.normalize_output:
            y = torch.tanh(y)
        return y

    def l1_norm(self):
        """Sum of |p| over all parameters (decoder.py:138-140)."""
        return sum(p.abs().sum() for p in self.parameters())

    def l2_norm(self):
        """Sum of p^2 over all parameters (decoder.py:142-144)."""
        return sum(p.pow(2).sum() for p in self.parameters())

    @property
    def num_learnable_params(self):
        return sum(p.nelement() for p in self.parameters() if p.requires_grad)
