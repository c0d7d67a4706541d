"""Anomaly-score distances of reference utils/jet_analysis/anomaly_detection.py (:454-510) on one kernel launch:
``mse``, ``chamfer`` (Euclidean norm) and ``chamfer_lorentz`` with the reference's signatures and return shapes.

The reference builds the (B, N, N, D) difference tensor and, in the batched form, goes through a DataLoader with one
host <-> device round trip per batch (:471-481); here a CTA per jet keeps the two particle sets in shared memory and
writes the per-particle minima.  There is no CPU fallback: the inputs are moved to the CUDA device.
"""
from __future__ import annotations

import torch

from . import _lib


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("gnn_jet_autoencoder_b200.anomaly needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def mse(p: torch.Tensor, q: torch.Tensor, dim: int = -1) -> torch.Tensor:
    """anomaly_detection.py:454-455: squared difference summed over ``dim`` (plain elementwise arithmetic)."""
    return ((p - q) ** 2).sum(dim=dim)


def _pair_min(p: torch.Tensor, q: torch.Tensor, lorentz: bool) -> torch.Tensor:
    if p.dim() != 3 or q.dim() != 3 or p.shape[0] != q.shape[0] or p.shape[2] != q.shape[2]:
        raise ValueError(f"expected (B, N, D) jets with equal B and D, got {tuple(p.shape)} and {tuple(q.shape)}")
    if p.shape[1] != q.shape[1]:
        raise RuntimeError("the reference adds the two per-particle minima elementwise: both jets need the same number of "
                           f"particles (got {p.shape[1]} and {q.shape[1]})")
    if lorentz and p.shape[2] != 4:
        raise ValueError("chamfer_lorentz needs 4-vectors (E, px, py, pz)")
    dev = p.device if p.is_cuda else _device()
    pc = p.detach().to(device=dev, dtype=torch.float32).contiguous()
    qc = q.detach().to(device=dev, dtype=torch.float32).contiguous()
    B, N, D = pc.shape
    out = torch.empty((2, B, N), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().gj_pair_min_dist(B, N, N, D, 1 if lorentz else 0, pc.data_ptr(), qc.data_ptr(), out[0].data_ptr(),
                                                out[1].data_ptr(), torch.cuda.current_stream().cuda_stream), "gj_pair_min_dist")
    return (out[0] + out[1]).to(p.dtype if p.dtype.is_floating_point else torch.float32)


def _batched(fn, p, q, batch_size):
    """anomaly_detection.py:471-481: with a positive ``batch_size`` the reference walks the data in batches on DEVICE and
    returns a CPU tensor; otherwise the result stays on the inputs' device."""
    if batch_size is not None and batch_size > 0:
        return torch.cat([fn(p[i:i + batch_size], q[i:i + batch_size]).cpu() for i in range(0, p.shape[0], batch_size)], dim=0)
    out = fn(p, q)
    return out if p.is_cuda else out.cpu()


def chamfer(p: torch.Tensor, q: torch.Tensor, batch_size: int = -1) -> torch.Tensor:
    """anomaly_detection.py:459-488: (B, N) tensor  min_j |p_i - q_j|_2 + min_i' |p_i' - q_i|_2  (per-particle sums of the two
    nearest-neighbour distances; L2 norm, not squared)."""
    return _batched(lambda a, b: _pair_min(a, b, False), p, q, batch_size)


def chamfer_lorentz(p: torch.Tensor, q: torch.Tensor, batch_size: int = -1) -> torch.Tensor:
    """anomaly_detection.py:491-510: the same with the Lorentz norm squared E^2 - px^2 - py^2 - pz^2 of the differences."""
    return _batched(lambda a, b: _pair_min(a, b, True), p, q, batch_size)
