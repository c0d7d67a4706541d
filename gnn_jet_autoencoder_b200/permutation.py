"""Permutation invariance / equivariance test of reference utils/permutation.py with the per-jet Python loops
(``apply_perm`` :112-114, ``particle_perm_rand`` :117-133: one ``randperm`` and one indexing op per jet) replaced by batched
device ops: the permutations of a whole batch come from one ``argsort`` of uniform keys, and they are applied with one
``gather``.  Same class / function names, arguments and return values.

Quirk (utils/permutation.py:157-161): with ``verbose=True`` the reference executes ``perm["values"] = dev`` on a tensor and
raises; here ``verbose`` adds ``summary["values"]`` and ``summary["perm"]``, which is what those lines are after.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Optional, Tuple, Union

import torch
from torch.utils.data import DataLoader

EPS = 1e-12


def dev(output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Relative deviation of output from target (utils/permutation.py:107-109)."""
    return (output - target).abs() / (target.abs() + EPS)


def apply_perm(perm: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """out[b, i] = x[b, perm[b, i]] for every jet at once (utils/permutation.py:112-114)."""
    idx = perm.to(device=x.device, dtype=torch.long)
    return torch.gather(x, 1, idx.unsqueeze(-1).expand(-1, -1, x.shape[-1]))


def particle_perm_rand(x: torch.Tensor, generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Independent uniformly random permutation of the particles of every jet (utils/permutation.py:117-133); returns
    (x_perm, perm) with perm of shape (batch_size, num_particles) on x's device."""
    batch_size, num_particles, _ = x.shape
    keys = torch.rand((batch_size, num_particles), device=x.device, generator=generator)
    perm = keys.argsort(dim=1)
    return apply_perm(perm, x), perm


def get_model_output(x: torch.Tensor, encoder, decoder) -> torch.Tensor:
    return decoder(encoder(x)).detach()


def get_dev_summary(dev: torch.Tensor, perm: Optional[torch.Tensor], verbose: bool = False,
                    save_path: Optional[Union[str, Path]] = None) -> Dict[str, Union[torch.Tensor, float]]:
    summary = {"mean": dev.mean().item(), "median": dev.median().item(), "max": dev.max().item(), "min": dev.min().item(),
               "std": dev.std().item()}
    if verbose:
        summary["values"] = dev
        if perm is not None:
            summary["perm"] = perm
    if save_path is not None:
        torch.save(summary, save_path)
    return summary


class PermutationTest:
    """``PermutationTest(encoder, decoder, device, dtype)(x)`` -> {"invariance": summary, "equivariance": summary} for a
    (B, N, F) tensor or a DataLoader of such batches (utils/permutation.py:11-74).  The models stay where they are when
    ``device`` / ``dtype`` are omitted (the reference's default ``dtype=DEFAULT_DEVICE`` is a slip; callers always pass both)."""

    def __init__(self, encoder, decoder, device: Optional[torch.device] = None, dtype: Optional[torch.dtype] = None):
        p = next(encoder.parameters())
        self.device = torch.device(device) if device is not None else p.device
        self.dtype = dtype if dtype is not None else p.dtype
        # as utils/permutation.py:19-20; the fused modules keep float32 parameters whatever dtype is requested
        # (Float32ParamsMixin) -- with the CLI default float64 (train.py:73-76) only the inputs / outputs are float64
        self.encoder = encoder.to(device=self.device, dtype=self.dtype)
        self.decoder = decoder.to(device=self.device, dtype=self.dtype)

    def __call__(self, x: Union[torch.Tensor, DataLoader], verbose: bool = False, save_dir: Optional[Union[str, Path]] = None):
        if isinstance(x, DataLoader):
            parts = [[], [], [], []]
            for xb in x:
                xb = xb.to(device=self.device, dtype=self.dtype)
                inv, p_inv = self.invariance_dev(xb)
                eqv, p_eqv = self.equivariance_dev(xb)
                for lst, t in zip(parts, (inv, eqv, p_inv, p_eqv)):
                    lst.append(t.detach())
            inv_dev, eqv_dev, perm_inv, perm_eqv = (torch.cat(lst, dim=0).cpu() for lst in parts)   # one host copy each
        elif isinstance(x, torch.Tensor):
            x = x.to(device=self.device, dtype=self.dtype)
            inv_dev, perm_inv = self.invariance_dev(x)
            eqv_dev, perm_eqv = self.equivariance_dev(x)
            inv_dev, eqv_dev, perm_inv, perm_eqv = (t.detach().cpu() for t in (inv_dev, eqv_dev, perm_inv, perm_eqv))
        else:
            raise TypeError("x must be a DataLoader or a Tensor. " f"Found: {type(x)}")
        path_inv = Path(save_dir) / "invariance.pt" if save_dir is not None else None
        path_eqv = Path(save_dir) / "equivariance.pt" if save_dir is not None else None
        return {"invariance": get_dev_summary(inv_dev, perm=perm_inv, verbose=verbose, save_path=path_inv),
                "equivariance": get_dev_summary(eqv_dev, perm=perm_eqv, verbose=verbose, save_path=path_eqv)}

    def invariance_dev(self, x: torch.Tensor):
        """NN(P(x)) against NN(x) for a random permutation P of every jet's particles."""
        y = get_model_output(x, self.encoder, self.decoder)
        x_perm, perm = particle_perm_rand(x)
        return dev(output=get_model_output(x_perm, self.encoder, self.decoder), target=y), perm

    def equivariance_dev(self, x: torch.Tensor):
        """NN(P(x)) against P(NN(x))."""
        y = get_model_output(x, self.encoder, self.decoder)
        x_perm, perm = particle_perm_rand(x)
        return dev(output=get_model_output(x_perm, self.encoder, self.decoder), target=apply_perm(perm, y)), perm
