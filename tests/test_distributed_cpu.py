"""world_size-2 gloo test of the data-parallel host logic: batch sharding + ONE all-reduce(sum) of the flat
gradient buffer reproduces the single-process gradient.  Gradients come from the oracle here
(no GPU); the same ``allreduce_flat_`` / ``shard_range`` / ``flat_layout`` run in GNNAETrainer on NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gnnae_oracle as O
    from golden_cases import CASES, make_input, make_params
    from gnn_jet_autoencoder_b200 import Decoder, Encoder
    from gnn_jet_autoencoder_b200.trainer import allreduce_flat_, flat_layout, layout_size, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = CASES["trainsh_n30"]
    ep, dp = make_params(case)
    x = make_input(case)
    lo, hi = shard_range(x.shape[0], rank, world)
    _, _, _, eg, dg = O.loss_and_grads(x[lo:hi], ep, dp, case["enc"], case["dec"], l1_lambda=0.0)
    enc, dec = Encoder(**case["enc"], device="cpu"), Decoder(**case["dec"], device="cpu")
    lay = flat_layout(enc, dec)
    flat = torch.zeros(layout_size(lay), dtype=torch.float64)
    for name, (off, shape) in lay.items():
        side, key = name.split(".", 1)
        g = (eg if side == "encoder" else dg)[key]
        flat[off:off + g.size] = torch.from_numpy(g.ravel())
    allreduce_flat_(flat)
    if rank == 0:
        np.save(os.path.join(out_dir, "flat.npy"), flat.numpy())
    dist.destroy_process_group()


def test_sharded_gradients_allreduce_to_full_batch(tmp_path):
    import gnnae_oracle as O
    from golden_cases import CASES, make_input, make_params
    from gnn_jet_autoencoder_b200 import Decoder, Encoder
    from gnn_jet_autoencoder_b200.trainer import flat_layout
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "flat.npy")
    case = CASES["trainsh_n30"]
    ep, dp = make_params(case)
    x = make_input(case)
    _, _, _, eg, dg = O.loss_and_grads(x, ep, dp, case["enc"], case["dec"], l1_lambda=0.0)
    lay = flat_layout(Encoder(**case["enc"], device="cpu"), Decoder(**case["dec"], device="cpu"))
    want = np.zeros_like(got)
    for name, (off, shape) in lay.items():
        side, key = name.split(".", 1)
        g = (eg if side == "encoder" else dg)[key]
        want[off:off + g.size] = g.ravel()
    assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)
