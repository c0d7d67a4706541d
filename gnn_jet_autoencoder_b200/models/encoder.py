"""Encoder: drop-in for reference models/encoder.py (GraphNet + latent aggregation)."""
from __future__ import annotations

import logging
from typing import List, Optional, Union

import torch
import torch.nn as nn

from .. import ops
from .const import GLOBAL_MIX, LOCAL_MIX
from .graphnet import Float32ParamsMixin, GraphNet, _default_device


def _canon(latent_map: str) -> str:
    return latent_map.lower().replace(" ", "_")


class Encoder(Float32ParamsMixin, nn.Module):
    """(B, N, input_node_size) -> latent (B, latent_node_size), or (B, N*latent_node_size) for the
    per-node ('local mix') map.  Same constructor / attributes / ``state_dict`` keys as the reference
    (encoder.py:13-131); see GraphNet for ``precision``."""

    def __init__(
        self,
        num_nodes: int,
        input_node_size: int,
        latent_node_size: int,
        node_sizes: List[List[int]],
        edge_sizes: List[List[int]],
        num_mps: int,
        alphas: Union[float, List[float]],
        dropout: float = 0.0,
        batch_norm: bool = False,
        latent_map: str = "global mix",
        device: Optional[torch.device] = None,
        dtype: Optional[torch.dtype] = None,
        precision: Optional[str] = None,
    ):
        super().__init__()
        self.device = torch.device(device) if device is not None else _default_device()
        self.dtype = dtype if dtype is not None else torch.float
        self.num_nodes = num_nodes
        self.input_node_size = input_node_size
        self.latent_node_size = latent_node_size
        self.node_sizes = node_sizes
        self.edge_sizes = edge_sizes
        self.num_mps = num_mps
        self.latent_map = latent_map
        canon = _canon(latent_map)
        self.latent_space_size = latent_node_size * num_nodes if canon in LOCAL_MIX else latent_node_size

        # encoder.py:91 tests the spelling WITHOUT the space->underscore normalisation used everywhere else,
        # and reads the caller's raw node_sizes list: "local mix" and "local_mix" therefore build different
        # GraphNets.  Kept, because checkpoints depend on it.
        out_width = node_sizes[-1][-1] if latent_map.lower() in LOCAL_MIX else latent_node_size
        self.encoder = GraphNet(num_nodes=num_nodes, input_node_size=input_node_size, output_node_size=out_width,
                                node_sizes=node_sizes, edge_sizes=edge_sizes, num_mps=num_mps, alphas=alphas,
                                dropout=dropout, batch_norm=batch_norm, device=self.device, dtype=self.dtype,
                                precision=precision)
        if canon in GLOBAL_MIX:
            self.mix_layer = nn.Linear(latent_node_size * num_nodes, latent_node_size, bias=False).to(self.device)
        elif canon in LOCAL_MIX:
            self.mix_layer = nn.Linear(out_width, latent_node_size).to(self.device)

    def forward(self, x: torch.Tensor, metric: str = "euclidean") -> torch.Tensor:
        batch = x.shape[0]
        y = self.encoder(x, metric=metric)
        z = self._to_latent(y.float() if y.dtype != torch.float32 else y, batch)
        return z if self.dtype == torch.float32 else z.to(self.dtype)

    def _to_latent(self, y: torch.Tensor, batch: int) -> torch.Tensor:
        """encoder.py:144-171."""
        canon = _canon(self.latent_map)
        if canon == "mean":
            return ops.latent_mean(y)
        if canon == "max":
            return torch.amax(y, dim=-2)
        if canon == "min":
            return torch.amin(y, dim=-2)
        if canon in GLOBAL_MIX:
            return ops.linear(y.reshape(batch, -1), self.mix_layer.weight, None)
        if canon in LOCAL_MIX:
            return ops.linear(y, self.mix_layer.weight, self.mix_layer.bias).reshape(batch, -1)
        logging.warning(f"Unknown latent map {self.latent_map} in Encoder. Using mean.")
        self.latent_map = "mean"
        return ops.latent_mean(y)

    def l1_norm(self):
        """Sum of |p| over all parameters (encoder.py:173-175)."""
        return sum(p.abs().sum() for p in self.parameters())

    def l2_norm(self):
        """Sum of p^2 over all parameters (encoder.py:177-179)."""
        return sum(p.pow(2).sum() for p in self.parameters())

    @property
    def num_learnable_params(self):
        return sum(p.nelement() for p in self.parameters() if p.requires_grad)
