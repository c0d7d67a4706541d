"""Property tests (hypothesis) of the host-side logic around the CUDA path: batch sharding, list broadcasting
(reference graphnet.py:305-311), descriptor bookkeeping through the C-ABI (no device work), the FLOP model."""
import numpy as np
from hypothesis import given, settings, strategies as st

from gnn_jet_autoencoder_b200 import _lib
from gnn_jet_autoencoder_b200.config import edge_macs_per_row
from gnn_jet_autoencoder_b200.models.graphnet import _broadcast
from gnn_jet_autoencoder_b200.trainer import shard_range, synthetic_jets


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 100000), st.integers(1, 64))
def test_shard_ranges_partition_the_batch(batch, world):
    """Rank shards are contiguous, disjoint, cover [0, B) and differ by at most one jet."""
    spans = [shard_range(batch, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == batch
    assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1


def _adjust_var_list(data, num):
    """The reference's helper, restated (graphnet.py:305-311): lists are padded with their last entry, scalars replicated,
    the result truncated to num entries."""
    if isinstance(data, (list, tuple)):
        data = list(data)
        if len(data) < num:
            data = data + [data[-1]] * (num - len(data))
        return data[:num]
    return [data] * num


@settings(max_examples=100, deadline=None)
@given(st.one_of(st.floats(0, 1), st.lists(st.integers(1, 64), min_size=1, max_size=8)), st.integers(1, 8))
def test_broadcast_follows_the_reference_helper(data, num):
    out = _broadcast(data, num)
    assert out == _adjust_var_list(data, num) and len(out) == num


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 64), st.lists(st.integers(1, 256), min_size=1, max_size=8), st.lists(st.integers(1, 256), min_size=1, max_size=8),
       st.integers(0, 300), st.integers(1, 200))
def test_param_count_matches_the_linear_layers(H, edge, node, batch, nodes):
    """gj_mp_param_count (host only) = the parameters of the nn.Linear stack of one step: edge layers from 2H+1, node layers
    from E_last + H (graphnet.py:84,102-127)."""
    want = sum(o * i + o for i, o in zip([2 * H + 1] + edge[:-1], edge)) + sum(o * i + o for i, o in zip([edge[-1] + H] + node[:-1], node))
    d = _lib.make_desc(batch, nodes, H, edge, node, 0.2, 0, 0)
    assert _lib.load().gj_mp_param_count(d) == want


@settings(max_examples=50, deadline=None)
@given(st.integers(1, 6), st.integers(1, 128), st.lists(st.integers(1, 128), min_size=1, max_size=4))
def test_edge_mac_model_counts_the_dense_formulation(num_mps, H, edge):
    """MACs per edge row = sum over steps and layers of in x out with in_0 = 2H + 1."""
    widths = [2 * H + 1] + edge
    assert edge_macs_per_row([[H]], [edge], num_mps) == num_mps * sum(a * b for a, b in zip(widths[:-1], widths[1:]))


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 40), st.integers(1, 160), st.integers(0, 2 ** 31 - 1))
def test_synthetic_jets_are_jetnet_shaped(batch, n, seed):
    x = synthetic_jets(batch, n, seed=seed)
    assert x.shape == (batch, n, 3) and x.dtype == np.float32 and np.isfinite(x).all()
    pt = x[..., 0]
    assert (pt >= 0).all() and (np.diff(pt, axis=1) <= 1e-7).all() and (pt.sum(1) <= 1 + 1e-5).all()      # pT ordered, relative
    assert (np.abs(x[..., 1:]) <= 0.5).all()
    real = pt > 0
    assert (real.sum(1) >= 1).all() or n < 3      # zero padded behind the real particles
    assert ((x == 0).all(-1) | real | (pt == 0)).all()
