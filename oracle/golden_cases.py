"""Case matrix shared by ``gen_golden.py`` (runs the reference) and the tests (run the oracle
and the CUDA path).  Weights and inputs are regenerated from seeds with numpy's PCG64, which
is bit-reproducible across machines, so fixtures only need to hold outputs.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import numpy as np

import gnnae_oracle as O

DEFAULT_EDGE = [[32, 128, 64, 16]]          # utils/argparse_utils.py:89-96 (CLI default)
DEFAULT_NODE = [[16], [32], [8]]            # utils/argparse_utils.py:97-104


def _case(N, B, *, edge=DEFAULT_EDGE, node=DEFAULT_NODE, num_mps=3, latent=20, latent_map="mean", vec=3,
          alphas=0.2, metric="euclidean", norm="cartesian", seed=0, normalize_output=False, store64=True,
          jet_w=1.0, dec_edge=None, dec_node=None, grad_stride=1, loss="chamfer", polar=False):
    enc = dict(num_nodes=N, input_node_size=vec, latent_node_size=latent, node_sizes=node, edge_sizes=edge,
               num_mps=num_mps, alphas=alphas, latent_map=latent_map)
    dec = dict(num_nodes=N, latent_node_size=latent, output_node_size=vec, node_sizes=dec_node or node,
               edge_sizes=dec_edge or edge, num_mps=num_mps, alphas=alphas, latent_map=latent_map,
               normalize_output=normalize_output)
    return dict(enc=enc, dec=dec, B=B, N=N, vec=vec, metric=metric, loss_norm_choice=norm, seed=seed,
                store64=store64, jet_features_weight=jet_w, l1_lambda=1e-8, grad_stride=grad_stride, loss_choice=loss,
                polar_coord=polar)


def gsub(case, v):
    """Fixtures of the wide architectures keep every ``grad_stride``-th entry of the (name-sorted, concatenated) encoder and
    decoder gradients, so that the committed files stay small; comparisons apply the same subsampling."""
    return np.asarray(v)[::case.get("grad_stride", 1)]


CASES = {
    # CLI-default architecture (BASELINE configs 1-4), small batch
    "default_n30": _case(30, 3, store64=False),
    "default_n33": _case(33, 2, store64=False, seed=1),
    # examples/train.sh widths: edge '16,16,8,8;' node '3;3;3;3;' latent 2
    "trainsh_n30": _case(30, 4, edge=[[16, 16, 8, 8]], node=[[3], [3], [3], [3]], latent=2, seed=2),
    # latent maps (models/const.py:1-2 spellings) incl. the 'local mix' vs 'local_mix' quirk
    "local_mix_us_n8": _case(8, 3, edge=[[16, 16]], node=[[8], [8]], num_mps=2, latent=4, latent_map="local_mix", seed=3),
    "local_mix_sp_n8": _case(8, 3, edge=[[16, 16]], node=[[8], [8]], num_mps=2, latent=4, latent_map="local mix", seed=4),
    "global_mix_n8": _case(8, 3, edge=[[16, 16]], node=[[8], [8]], num_mps=2, latent=4, latent_map="global mix", seed=5),
    "max_n5": _case(5, 5, edge=[[16, 32, 16]], node=[[8, 16], [16]], num_mps=2, latent=6, latent_map="max", seed=6),
    "min_n5": _case(5, 5, edge=[[16, 32, 16]], node=[[8, 16], [16]], num_mps=2, latent=6, latent_map="min", seed=7),
    "bogus_map_n5": _case(5, 2, edge=[[16]], node=[[8]], num_mps=1, latent=3, latent_map="bogus", seed=8),
    # 4-vectors with the minkowskian metric (applies only where the current width is 4) and loss norm
    "mink_n6": _case(6, 3, edge=[[16, 16]], node=[[4], [4]], num_mps=2, latent=4, vec=4, metric="minkowskian",
                     norm="minkowskian", seed=9),
    # num_mps > len(node_sizes) (broadcast) and H_0 < input width (crop), tanh output, jet weight != 1
    "broadcast_crop_n7": _case(7, 2, edge=[[16, 16]], node=[[2, 8]], num_mps=3, latent=5, seed=10,
                               normalize_output=True, jet_w=0.5, alphas=0.1),
    # single particle and two particles
    "n1": _case(1, 4, edge=[[16, 16]], node=[[8]], num_mps=2, latent=3, seed=11),
    "n2": _case(2, 4, edge=[[16, 16]], node=[[8]], num_mps=2, latent=3, seed=12),
    # wide-sweep shaped (BASELINE config 5, smallest member) at small N
    "wide64_n9": _case(9, 2, edge=[[64, 64]], node=[[64]], num_mps=3, latent=8, seed=13, store64=False),
    # j-block boundaries of the fused tensor-core kernels (32 j's per warp) and the N^2 tiling limit (JetNet150 shape)
    "default_n31": _case(31, 2, store64=False, seed=14),
    "default_n32": _case(32, 2, store64=False, seed=15),
    "default_n64": _case(64, 2, store64=False, seed=16),
    "default_n150": _case(150, 2, store64=False, seed=17),
    # a batch that is not a multiple of the four (jet, j block) tasks of a tile group, larger than one wave of groups
    "default_b257": _case(30, 257, store64=False, seed=18),
    # BASELINE config 5 (deep / wide sweep): node_sizes [[H]], edge_sizes [[H, H]], 3-6 steps, latent 1-64
    # (with seed 19 every gradient upstream of the decoder's last edge network differs from float64 by the same 4e-5 in fp32 mode
    # while other seeds give 2e-7: the signature of one pre-activation within fp32 rounding of the LeakyReLU kink, where any
    # fp32 evaluation order picks a slope; tools/grad_err.py wide128_n30 fp32 19 shows it)
    "wide128_n30": _case(30, 2, edge=[[128, 128]], node=[[128]], num_mps=3, latent=8, seed=119, store64=False, grad_stride=4),
    "wide256_n12": _case(12, 2, edge=[[256, 256]], node=[[256]], num_mps=3, latent=16, seed=20, store64=False, grad_stride=16),
    "wide64_mps6_n10": _case(10, 2, edge=[[64, 64]], node=[[64]], num_mps=6, latent=1, seed=21, store64=False, grad_stride=2),
    "wide128_lat64_n33": _case(33, 2, edge=[[128, 128]], node=[[128]], num_mps=4, latent=64, seed=22, store64=False,
                               grad_stride=8),
    # 3-vectors with a Minkowskian loss norm: pairwise distances stay cartesian (distance_sq.py:43-44), the jet term does not
    # (chamfer_loss.py:40)
    "loss_mink3_n6": _case(6, 3, edge=[[16, 16]], node=[[8]], num_mps=2, latent=4, norm="minkowskian", seed=23, jet_w=0.7),
    # the MSE branch of get_loss (utils/train.py:359-361) and the polar-coordinate clamp of the train loop (:55-65) in front of
    # the Chamfer loss with the polar norm, for 3- and 4-vectors; the last one also has the tanh output
    "mse_n6": _case(6, 3, edge=[[16, 16]], node=[[8]], num_mps=2, latent=4, seed=24, loss="mse"),
    "polar3_n6": _case(6, 3, edge=[[16, 16]], node=[[8]], num_mps=2, latent=4, seed=25, norm="polar", polar=True),
    "polar4_tanh_n6": _case(6, 3, edge=[[16, 16]], node=[[4], [4]], num_mps=2, latent=4, vec=4, seed=26, norm="polar", polar=True,
                            normalize_output=True, loss="mse"),
}


def _linear(rng, out_f, in_f, bias=True):
    bound = 1.0 / np.sqrt(in_f)
    w = rng.uniform(-bound, bound, (out_f, in_f))
    b = rng.uniform(-bound, bound, (out_f,)) if bias else None
    return w, b


def make_params(case):
    """Seeded parameters under the reference's state_dict names."""
    rng = np.random.default_rng(1000 + case["seed"])
    enc, dec = case["enc"], case["dec"]
    N = enc["num_nodes"]
    eout = O.encoder_out_width(enc["latent_map"], enc["latent_node_size"], enc["node_sizes"])
    ep = O.init_graphnet_params(rng, enc["input_node_size"], eout, enc["node_sizes"], enc["edge_sizes"],
                                enc["num_mps"], prefix="encoder.")
    lm = enc["latent_map"].lower().replace(" ", "_")
    if lm in O.GLOBAL_MIX:
        ep["mix_layer.weight"], _ = _linear(rng, enc["latent_node_size"], enc["latent_node_size"] * N, bias=False)
    elif lm in O.LOCAL_MIX:
        ep["mix_layer.weight"], ep["mix_layer.bias"] = _linear(rng, enc["latent_node_size"], eout)
    dp = {}
    h0 = dec["node_sizes"][0][0]
    if lm in O.LOCAL_MIX:
        dp["linear.weight"], dp["linear.bias"] = _linear(rng, h0, dec["latent_node_size"])
    else:
        dp["linear.weight"], dp["linear.bias"] = _linear(rng, N * h0, dec["latent_node_size"])
    dp.update(O.init_graphnet_params(rng, h0, dec["output_node_size"], dec["node_sizes"], dec["edge_sizes"],
                                     dec["num_mps"], prefix="decoder."))
    return ep, dp


def make_input(case):
    if case["vec"] == 3:
        return O.synthetic_jets(case["B"], case["N"], seed=1234 + case["seed"], dtype=np.float64)
    rng = np.random.default_rng(1234 + case["seed"])
    return rng.normal(0.0, 0.5, (case["B"], case["N"], case["vec"]))
