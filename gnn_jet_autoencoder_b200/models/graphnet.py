"""GraphNet: drop-in for reference models/graphnet.py backed by fused sm_100a kernels.

Constructor arguments, public attributes, sub-module names (``edge_net.{t}.{k}``, ``node_net.{t}.{k}``,
hence the ``state_dict`` layout) and ``forward(x, metric)`` follow the reference (graphnet.py:13-171,
SURVEY.md 8.b).  The arithmetic does not: one message-passing step is ONE library call
(``gj_mp_step_fwd`` / ``gj_mp_step_bwd``: 4-5 kernel launches forward, 5-9 backward; the trainer's chain calls
save one of each per step), the (B,N,N,2H+1) pair
tensor never exists in HBM for the widths the fused edge kernels cover, and there is no CPU implementation -- calling
``forward`` on a CPU module raises.
"""
from __future__ import annotations

import logging
import os
import warnings
from typing import List, Optional, Union

import torch
import torch.nn as nn

from .. import ops
from .const import EPS


def _broadcast(data, num):
    """graphnet.py:305-311: lists are padded by repeating their last entry, scalars replicated;
    the result is truncated to ``num`` entries."""
    if isinstance(data, (list, tuple)):
        data = list(data)
        if len(data) < num:
            data = data + [data[-1]] * (num - len(data))
    else:
        data = [data] * num
    return data[:num]


def _default_device():
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


_fp32_hint_given = False


def _hint_default_precision(precision: Optional[str], device: torch.device) -> None:
    """One log line per process when a CUDA GraphNet is built in the default precision: the fp32 mode reproduces the reference to
    <=1e-5 but runs on the CUDA cores, ~27x slower than the tensor-core mode -- a reference user who only swaps the
    import should learn about the ``precision`` keyword / the GNNJET_PRECISION environment variable."""
    global _fp32_hint_given
    if _fp32_hint_given or precision is not None or "GNNJET_PRECISION" in os.environ or device.type != "cuda":
        return
    _fp32_hint_given = True
    logging.info("gnn_jet_autoencoder_b200: GraphNet runs in the fp32-parity mode (CUDA cores, <=1e-5 vs the float64 reference); "
                 "pass precision='bf16' or set GNNJET_PRECISION=bf16 for the tcgen05 tensor-core kernels (<=2e-2, ~27x faster).")


def _resolve_precision(precision: Optional[str]) -> str:
    p = (precision or os.environ.get("GNNJET_PRECISION", "fp32")).lower()
    if p not in ops.PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(ops.PRECISIONS)}, got {precision!r}")
    return "bf16" if ops.PRECISIONS[p] == ops.GJ_PREC_BF16 else "fp32"


class Float32ParamsMixin:
    """Parameters of the fused path are ALWAYS stored in float32 (the kernels read them in place from one packed
    buffer).  ``module.to(dtype=torch.float64)`` -- which reference callers issue with the CLI's default dtype, e.g.
    ``PermutationTest`` (utils/permutation.py:19-20 via train.py:73-76) -- therefore leaves floating-point parameters in
    float32; when neither device nor dtype would change, the very same storage is kept, so views into a packed buffer
    (``GraphNet.flatten_parameters``, ``GNNAETrainer.flat``) survive the call.  Inputs of any dtype are accepted and
    outputs are returned in the constructor's ``dtype``."""

    def _apply(self, fn, recurse=True):
        def keep32(t):
            out = fn(t)
            if out.is_floating_point() and out.dtype != torch.float32 and t.dtype == torch.float32:
                out = out.to(torch.float32)
                if out.device == t.device:
                    return t          # fp32 -> wider -> fp32 is the identity: keep the storage (and any packed-buffer view)
            return out
        return super()._apply(keep32, recurse)


class GraphNet(Float32ParamsMixin, nn.Module):
    """Fully connected message-passing network with the pair distance as edge feature.

    Extra keyword (not in the reference): ``precision`` -- ``"fp32"`` (default; SIMT FFMA kernels,
    <=1e-5 relative to the float64 reference) or ``"bf16"`` (edge-MLP hidden layers on tcgen05 tensor
    cores, bf16 operands / fp32 accumulation).  Parameters are always stored in float32; a float64
    ``dtype`` only changes the dtype of the returned tensor (fp64 compute is not a B200 target).
    """

    def __init__(
        self,
        num_nodes: int,
        input_node_size: int,
        output_node_size: int,
        node_sizes: List[List[int]],
        edge_sizes: List[List[int]],
        num_mps: int,
        alphas: Union[List[float], float] = 0.1,
        dropout: float = 0.0,
        batch_norm: bool = False,
        device: Optional[torch.device] = None,
        dtype: Optional[torch.dtype] = None,
        precision: Optional[str] = None,
    ):
        super().__init__()
        node_sizes = _broadcast(node_sizes, num_mps)
        edge_sizes = _broadcast(edge_sizes, num_mps)
        alphas = _broadcast(alphas, num_mps)
        self.device = torch.device(device) if device is not None else _default_device()
        self.dtype = dtype if dtype is not None else torch.float
        self.eps = EPS
        self.precision = _resolve_precision(precision)
        _hint_default_precision(precision, self.device)

        self.num_nodes = num_nodes
        self.input_node_size = input_node_size
        self.output_node_size = output_node_size
        self.node_sizes = node_sizes
        self.edge_sizes = edge_sizes
        self.input_edge_sizes = [2 * s[0] + 1 for s in node_sizes]   # [h_i | h_j | d_ij], graphnet.py:84
        self.num_mps = num_mps
        self.alphas = alphas
        self.dropout_p = dropout
        self.batch_norm = batch_norm
        if dropout > 0:
            warnings.warn("Dropout is going to break the permutation symmetry of the model in training mode.")

        # Parameter creation order matches graphnet.py:100-127 step by step (edge stack, inner node chain,
        # first node layer, output node layer) so that equal seeds give equal initial weights.
        self.node_net = nn.ModuleList()
        self.edge_net = nn.ModuleList()
        if batch_norm:
            self.bn_node = nn.ModuleList()
            self.bn_edge = nn.ModuleList()
        for t in range(num_mps):
            h_t, e_t, n_t = node_sizes[t][0], list(edge_sizes[t]), list(node_sizes[t])
            widths = [self.input_edge_sizes[t]] + e_t
            edge = nn.ModuleList(nn.Linear(a, b) for a, b in zip(widths[:-1], widths[1:]))
            chain = [nn.Linear(a, b) for a, b in zip(n_t[:-1], n_t[1:])]
            first = nn.Linear(e_t[-1] + h_t, h_t)                      # input is [sum_j edge | h], graphnet.py:246
            nxt = node_sizes[t + 1][0] if t + 1 < num_mps else output_node_size
            last = nn.Linear(n_t[-1], nxt)
            node = nn.ModuleList([first, *chain, last])
            self.edge_net.append(edge)
            self.node_net.append(node)
            if batch_norm:   # kept only so that state_dict keys exist; forward refuses
                self.bn_edge.append(nn.ModuleList(nn.BatchNorm1d(w) for w in e_t))
                self.bn_node.append(nn.ModuleList(nn.BatchNorm1d(l.out_features) for l in node))
        self._flat = None
        self._flat_owner = None      # set by GNNAETrainer: the packed buffer then belongs to the trainer
        self.to(device=self.device, dtype=torch.float32)

    # ---- packed parameter storage -----------------------------------------------------------------
    def step_parameters(self, t: int) -> List[nn.Parameter]:
        """Parameters of step ``t`` in the kernel's packing order (include/gnnjet_b200.h): edge stack
        (weight, bias per layer) then node stack."""
        out = []
        for layer in list(self.edge_net[t]) + list(self.node_net[t]):
            out += [layer.weight, layer.bias]
        return out

    def _flat_ok(self) -> bool:
        f = self._flat
        if f is None:
            return False
        off = 0
        for t in range(self.num_mps):
            for p in self.step_parameters(t):
                if p.dtype != torch.float32 or p.device != f.device or p.data_ptr() != f.data_ptr() + 4 * off:
                    return False
                off += p.numel()
        return True

    def flatten_parameters(self, storage: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Re-home every Linear parameter as a view of one flat fp32 buffer (values preserved), step by step
        in packing order, so the kernels read the packed block without a gather.  ``storage`` lets a caller
        (the trainer) provide a slice of a larger buffer."""
        plist = [p for t in range(self.num_mps) for p in self.step_parameters(t)]
        total = sum(p.numel() for p in plist)
        dev = plist[0].device
        if storage is None:
            storage = torch.empty(total, device=dev, dtype=torch.float32)
        assert storage.numel() == total and storage.dtype == torch.float32 and storage.is_contiguous()
        off = 0
        with torch.no_grad():
            for p in plist:
                n = p.numel()
                seg = storage[off:off + n].view(p.shape)
                seg.copy_(p.detach().to(device=storage.device, dtype=torch.float32))
                p.data = seg
                off += n
        self._flat = storage
        self._step_offsets = []
        off = 0
        for t in range(self.num_mps):
            n = sum(p.numel() for p in self.step_parameters(t))
            self._step_offsets.append((off, n))
            off += n
        return storage

    @property
    def num_flat_params(self) -> int:
        return sum(p.numel() for t in range(self.num_mps) for p in self.step_parameters(t))

    def step_config(self, t: int, metric: str):
        """(num_nodes, H_t, edge widths, node widths, alpha, metric id, precision id) of step ``t``."""
        node_w = [l.out_features for l in self.node_net[t]]
        return (self.num_nodes, self.node_sizes[t][0], list(self.edge_sizes[t]), node_w, float(self.alphas[t]),
                ops.metric_id(metric), ops.PRECISIONS[self.precision])

    # ---- forward ----------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, metric: str = "euclidean") -> torch.Tensor:
        """x (B, num_nodes, input_node_size) -> (B, num_nodes, output_node_size).

        ``metric``: 'euclidean' | 'minkowskian' (applied only where the current node width is 4,
        graphnet.py:155)."""
        self.metric = metric.lower()
        if self.metric not in ("euclidean", "cartesian", "minkowskian"):      # graphnet.py:324-326
            logging.warning(f"Metric ({self.metric}) for adjacency matrix is not implemented. Use 'cartesian' instead.")
        if self.batch_norm:
            raise NotImplementedError("batch_norm=True crashes in the reference (BatchNorm1d on a 4-D tensor, "
                                      "graphnet.py:287-288) and is not part of the fused path")
        if self.dropout_p > 0:
            raise NotImplementedError("dropout > 0 is not part of the fused GraphNet path")
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphNet.forward runs on sm_100a CUDA kernels only; move the module to a CUDA device "
                               "(there is no CPU fallback)")
        if not self._flat_ok():
            if self._flat_owner is not None:
                raise RuntimeError("GraphNet parameters were moved or re-typed after a GNNAETrainer packed them into its flat "
                                   "buffer; the trainer would keep updating storage the module no longer reads.  Build a new "
                                   "GNNAETrainer after moving the modules.")
            self.flatten_parameters()
        batch = x.shape[0]
        h = x.to(device=dev, dtype=torch.float32)
        if h.dim() != 3 or h.shape[1] != self.num_nodes:
            h = h.reshape(batch, self.num_nodes, -1)
        # the zero padding / cropping of graphnet.py:152 is folded into the first step (h_ld / h_cols)
        for t in range(self.num_mps):
            off, n = self._step_offsets[t]
            h = ops.mp_step(h, self._flat[off:off + n], self.step_config(t, self.metric), self.step_parameters(t))
        h = h.view(batch, self.num_nodes, self.output_node_size)
        return h if self.dtype == torch.float32 else h.to(self.dtype)
