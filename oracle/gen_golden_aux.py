"""Golden fixtures for the adjacent components of SURVEY.md 8.f, produced by running the REFERENCE ITSELF (build container only):

  * utils/jet_analysis/anomaly_detection.py: ``mse``, ``chamfer``, ``chamfer_lorentz``, ``hungarian``, ``hungarian_lorentz``
    (:454-590), plain and batched (``batch_size``) forms;
  * utils/losses/hungarian_mse/hungarian_mse.py: ``HungarianMSELoss`` in its four coordinate modes (value and gradient w.r.t. the
    reconstruction);
  * utils/permutation.py: ``apply_perm`` and ``dev`` on a fixed permutation.

    python oracle/gen_golden_aux.py [--ref /root/reference]

Inputs are regenerated from seeds (numpy PCG64) by ``aux_inputs``; the fixture tests/golden/aux_reference.npz holds outputs only.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import add_reference_to_path  # noqa: E402

ANOMALY_SHAPES = [(5, 30, 3), (3, 30, 4), (2, 150, 4), (4, 1, 3), (6, 7, 4), (3, 33, 3)]
HUNGARIAN_MODES = [(True, False), (True, True), (False, False), (False, True)]      # (abs_coord, polar_coord)


def aux_inputs(shape, seed):
    """(p, q): reconstructed / target jets, float64.  4-vectors get a positive energy-like first component."""
    rng = np.random.default_rng(5000 + seed)
    p, q = rng.normal(0.0, 1.0, shape), rng.normal(0.0, 1.0, shape)
    if shape[-1] == 4:
        p[..., 0] = np.abs(p[..., 0]) + 1.0
        q[..., 0] = np.abs(q[..., 0]) + 1.0
    return p, q


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "..", "tests", "golden", "aux_reference.npz"))
    args = ap.parse_args()
    add_reference_to_path(args.ref)
    import torch
    import utils.jet_analysis.anomaly_detection as AD
    from utils.losses.hungarian_mse.hungarian_mse import HungarianMSELoss
    import utils.permutation as PM
    AD.DEVICE = torch.device("cpu")
    out = {}
    for s, shape in enumerate(ANOMALY_SHAPES):
        p, q = aux_inputs(shape, s)
        pt, qt = torch.from_numpy(p), torch.from_numpy(q)
        tag = "x".join(map(str, shape))
        out[f"mse_{tag}"] = AD.mse(pt, qt).numpy()
        out[f"chamfer_{tag}"] = AD.chamfer(pt, qt).numpy()
        out[f"chamfer_b2_{tag}"] = AD.chamfer(pt, qt, batch_size=2).numpy()
        out[f"hungarian_{tag}"] = AD.hungarian(pt, qt).numpy()
        out[f"hungarian_b2_{tag}"] = AD.hungarian(pt, qt, batch_size=2).numpy()
        if shape[-1] == 4:
            out[f"chamfer_lorentz_{tag}"] = AD.chamfer_lorentz(pt, qt).numpy()
            out[f"hungarian_lorentz_{tag}"] = AD.hungarian_lorentz(pt, qt).numpy()
    for s, shape in enumerate([(4, 30, 3), (3, 12, 4)]):
        p, q = aux_inputs(shape, 100 + s)
        for abs_coord, polar in HUNGARIAN_MODES:
            tag = "x".join(map(str, shape)) + f"_abs{int(abs_coord)}_polar{int(polar)}"
            if abs_coord:
                pt = torch.from_numpy(p).clone().requires_grad_(True)
                loss = HungarianMSELoss()(pt, torch.from_numpy(q), abs_coord=abs_coord, polar_coord=polar)
                loss.backward()
                out[f"hmse_grad_{tag}"] = pt.grad.numpy()
            else:
                # the relative-coordinate branch modifies unbind() views in place (hungarian_mse/utils.py:43-45) and raises under
                # autograd when the reconstruction requires grad: the reference can only evaluate it without a graph
                with torch.no_grad():
                    loss = HungarianMSELoss()(torch.from_numpy(p).clone(), torch.from_numpy(q).clone(), abs_coord=abs_coord, polar_coord=polar)
            out[f"hmse_{tag}"] = np.float64(loss.item())
    # permutation helpers
    rng = np.random.default_rng(9)
    x = rng.normal(size=(3, 6, 3))
    perm = np.stack([rng.permutation(6) for _ in range(3)])
    out["perm_apply"] = PM.apply_perm(torch.from_numpy(perm), torch.from_numpy(x)).numpy()
    out["perm_dev"] = PM.dev(torch.from_numpy(x), torch.from_numpy(x[:, ::-1].copy())).numpy()
    out["perm_x"], out["perm_perm"] = x, perm
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, len(out), "arrays")


if __name__ == "__main__":
    main()
