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


def assignment(p: torch.Tensor, q: torch.Tensor, lorentz: bool = False):
    """Optimal particle matching per jet on the device: (col_for_row (B, N) int64, total_cost (B,)) with
    col_for_row[b] == scipy.optimize.linear_sum_assignment(cost[b])[1] for cost = cdist(p, q) (or the Lorentz norm squared of
    the differences) -- the matching step of anomaly_detection.py:537-541 / :579-585 and hungarian_mse.py:51-52."""
    if p.dim() != 3 or p.shape != q.shape:
        raise ValueError(f"expected two (B, N, D) tensors of equal shape, got {tuple(p.shape)} and {tuple(q.shape)}")
    if lorentz and p.shape[2] != 4:
        raise ValueError("the Lorentz norm needs 4-vectors (E, px, py, pz)")
    dev = p.device if p.is_cuda else _device()
    pc = p.detach().to(device=dev, dtype=torch.float32).contiguous()
    qc = q.detach().to(device=dev, dtype=torch.float32).contiguous()
    B, N, D = pc.shape
    match = torch.empty((B, N), device=dev, dtype=torch.int32)
    total = torch.empty((B,), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().gj_assignment(B, N, D, 1 if lorentz else 0, pc.data_ptr(), qc.data_ptr(), match.data_ptr(),
                                             total.data_ptr(), torch.cuda.current_stream().cuda_stream), "gj_assignment")
    return match.long(), total


def _hungarian(p: torch.Tensor, q: torch.Tensor, lorentz: bool) -> torch.Tensor:
    match, _ = assignment(p, q, lorentz)
    pd = p.to(match.device)
    # anomaly_detection.py:543-546: p_shuffle[b] = p[b, matching[b]], then the particle-wise squared error against q
    p_shuffle = torch.gather(pd, 1, match.unsqueeze(-1).expand(-1, -1, pd.shape[-1]))
    return mse(p_shuffle, q.to(match.device))


def hungarian(p: torch.Tensor, q: torch.Tensor, batch_size: int = -1) -> torch.Tensor:
    """anomaly_detection.py:513-547: (B, N) squared errors after the optimal (Euclidean-cost) matching of the particles."""
    return _batched(lambda a, b: _hungarian(a, b, False), p, q, batch_size)


def hungarian_lorentz(p: torch.Tensor, q: torch.Tensor, batch_size: int = -1) -> torch.Tensor:
    """anomaly_detection.py:550-590: the same with the Lorentz norm squared of the differences as the matching cost."""
    return _batched(lambda a, b: _hungarian(a, b, True), p, q, batch_size)
