"""CPU oracle for the GNNAE GraphNet hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may import it.  The product
path (``gnn_jet_autoencoder_b200``) never imports anything under ``oracle/`` and has no
CPU fallback.

It is an independent numpy restatement (default float64) of the reference's algorithm,
forward AND hand-derived backward, written from the maths in SURVEY.md section 0 and the
reference files cited per function (paths relative to the reference checkout):

  models/graphnet.py, models/encoder.py, models/decoder.py,
  utils/losses/chamfer_loss/{chamfer_loss,distance_sq}.py, utils/train.py:51-85,330-385,
  utils/initialize.py:148-153.

Parity status: PINNED.  The reference ships no golden vectors or tests, so
the oracle is pinned against outputs of the reference itself: ``oracle/gen_golden.py``
imports the reference's own ``Encoder`` / ``Decoder`` / ``ChamferLoss`` from the read-only
reference checkout in the build container, runs them in float64 with torch autograd, and
commits inputs, weights, outputs and gradients under ``tests/golden/``;
``tests/test_oracle.py`` checks this file against every one of those vectors.

Parameters are carried as ``dict[str, np.ndarray]`` keyed by the reference's ``state_dict``
names (``edge_net.{t}.{k}.weight`` ... ), weight shape ``(out, in)`` as in ``nn.Linear``.
"""
from __future__ import annotations

import numpy as np

EPS = 1e-16  # utils/const.py:5

LOCAL_MIX = ("local", "local_mix", "node", "node_mix")  # models/const.py:1
GLOBAL_MIX = ("global", "global_mix", "graph", "graph_mix")  # models/const.py:2


# --------------------------------------------------------------------------------------
# architecture bookkeeping
# --------------------------------------------------------------------------------------
def adjust_var_list(data, num):
    """graphnet.py:305-311: pad by repeating the LAST element, truncate to ``num``;
    scalars become ``[v] * num``."""
    if isinstance(data, (list, tuple)):
        data = list(data)
        if len(data) < num:
            data = data + [data[-1]] * (num - len(data))
    else:
        data = [data] * num
    return data[:num]


def graphnet_layer_shapes(input_node_size, output_node_size, node_sizes, edge_sizes, num_mps):
    """Per-step (out,in) shapes of the edge and node Linear stacks (graphnet.py:59-61,84,100-127).

    edge step t : Linear(2*H_t+1 -> E_t0), Linear(E_t0 -> E_t1), ...
    node step t : Linear(E_t,last + H_t -> H_t), chain over node_sizes[t],
                  Linear(node_sizes[t][-1] -> node_sizes[t+1][0] | output_node_size)
    """
    node_sizes = adjust_var_list(node_sizes, num_mps)
    edge_sizes = adjust_var_list(edge_sizes, num_mps)
    edge_shapes, node_shapes = [], []
    for t in range(num_mps):
        h_t = node_sizes[t][0]
        ins = [2 * h_t + 1] + list(edge_sizes[t][:-1])
        edge_shapes.append([(o, i) for i, o in zip(ins, edge_sizes[t])])
        ns = [(h_t, edge_sizes[t][-1] + h_t)]
        for a, b in zip(node_sizes[t][:-1], node_sizes[t][1:]):
            ns.append((b, a))
        nxt = node_sizes[t + 1][0] if t + 1 < num_mps else output_node_size
        ns.append((nxt, node_sizes[t][-1]))
        node_shapes.append(ns)
    return edge_shapes, node_shapes


def init_graphnet_params(rng, input_node_size, output_node_size, node_sizes, edge_sizes, num_mps,
                         prefix="", dtype=np.float64):
    """Random parameters with nn.Linear-like scale (uniform +-1/sqrt(in)); test helper."""
    es, ns = graphnet_layer_shapes(input_node_size, output_node_size, node_sizes, edge_sizes, num_mps)
    p = {}
    for t in range(num_mps):
        for kind, shapes in (("edge_net", es[t]), ("node_net", ns[t])):
            for k, (o, i) in enumerate(shapes):
                bound = 1.0 / np.sqrt(i)
                p[f"{prefix}{kind}.{t}.{k}.weight"] = rng.uniform(-bound, bound, (o, i)).astype(dtype)
                p[f"{prefix}{kind}.{t}.{k}.bias"] = rng.uniform(-bound, bound, (o,)).astype(dtype)
    return p


def _stack(params, prefix, kind, t):
    ws, bs, k = [], [], 0
    while f"{prefix}{kind}.{t}.{k}.weight" in params:
        ws.append(params[f"{prefix}{kind}.{t}.{k}.weight"])
        bs.append(params[f"{prefix}{kind}.{t}.{k}.bias"])
        k += 1
    return ws, bs


# --------------------------------------------------------------------------------------
# elementary pieces
# --------------------------------------------------------------------------------------
def leaky(z, alpha):
    """F.leaky_relu (graphnet.py:268,286)."""
    return np.where(z > 0, z, alpha * z)


def leaky_grad(z, alpha):
    """torch's LeakyReluBackward: slope 1 where z > 0, alpha elsewhere (incl. z == 0)."""
    return np.where(z > 0, 1.0, alpha).astype(z.dtype)


def metric_signs(width, metric):
    """graphnet.py:155,314-327: 'minkowskian' applies only when the CURRENT width is 4:
    2*x0^2 - sum x^2 = x0^2 - x1^2 - x2^2 - x3^2; everything else is euclidean."""
    s = np.ones(width)
    if width == 4 and str(metric).lower() == "minkowskian":
        s[1:] = -1.0
    return s


def pair_distance(h, metric):
    """graphnet.py:211-218: d[b,i,j] = metric(h[b,j] - h[b,i] + eps) with eps added per
    component BEFORE squaring.  Returns (d, diff) with diff[b,i,j,:] = h_j - h_i + eps."""
    diff = h[:, None, :, :] - h[:, :, None, :] + EPS
    s = metric_signs(h.shape[-1], metric).astype(h.dtype)
    return (diff * diff * s).sum(-1), diff


# --------------------------------------------------------------------------------------
# one message-passing step (graphnet.py:154-168)
# --------------------------------------------------------------------------------------
def mp_step_forward(h, edge_w, edge_b, node_w, node_b, alpha, metric="euclidean"):
    """h (B,N,H) -> (h' (B,N,H'), cache).

    A_ij = [h_i | h_j | d_ij]                      graphnet.py:211-222  (_getA)
    A    = leaky(W_k A + b_k) for every edge layer graphnet.py:284-286  (_edge_conv)
    e_i  = sum_j A_ij  (self pair included)        graphnet.py:243      (_concat)
    u_i  = [e_i | h_i]  (edge aggregate FIRST)     graphnet.py:246
    h_i  = leaky(V_k u + c_k) for every node layer graphnet.py:266-268  (_aggregate)
    """
    B, N, H = h.shape
    d, diff = pair_distance(h, metric)
    hi = np.broadcast_to(h[:, :, None, :], (B, N, N, H))
    hj = np.broadcast_to(h[:, None, :, :], (B, N, N, H))
    a = np.concatenate([hi, hj, d[..., None]], axis=-1)
    edge_in, edge_z = [], []
    for w, b in zip(edge_w, edge_b):
        edge_in.append(a)
        z = a @ w.T + b
        edge_z.append(z)
        a = leaky(z, alpha)
    e = a.sum(axis=2)
    y = np.concatenate([e, h], axis=-1)
    node_in, node_z = [], []
    for w, b in zip(node_w, node_b):
        node_in.append(y)
        z = y @ w.T + b
        node_z.append(z)
        y = leaky(z, alpha)
    cache = dict(h=h, diff=diff, edge_in=edge_in, edge_z=edge_z, node_in=node_in, node_z=node_z,
                 alpha=alpha, metric=metric, e_width=e.shape[-1])
    return y, cache


def mp_step_backward(dy, cache, edge_w, node_w):
    """Hand-derived adjoint of ``mp_step_forward``.  Returns (dh, d_edge_w, d_edge_b, d_node_w, d_node_b)."""
    alpha = cache["alpha"]
    h = cache["h"]
    B, N, H = h.shape
    # node MLP, last layer first
    d_node_w, d_node_b = [None] * len(node_w), [None] * len(node_w)
    g = dy
    for k in reversed(range(len(node_w))):
        gz = g * leaky_grad(cache["node_z"][k], alpha)
        x = cache["node_in"][k]
        d_node_w[k] = np.einsum("bno,bni->oi", gz, x)
        d_node_b[k] = gz.sum(axis=(0, 1))
        g = gz @ node_w[k]
    E = cache["e_width"]
    de, dh = g[..., :E], g[..., E:].copy()
    # edge MLP: d(sum_j A_ij) broadcasts de_i over j
    g = np.broadcast_to(de[:, :, None, :], (B, N, N, E))
    d_edge_w, d_edge_b = [None] * len(edge_w), [None] * len(edge_w)
    for k in reversed(range(len(edge_w))):
        gz = g * leaky_grad(cache["edge_z"][k], alpha)
        x = cache["edge_in"][k]
        d_edge_w[k] = np.einsum("bijo,bijc->oc", gz, x)
        d_edge_b[k] = gz.sum(axis=(0, 1, 2))
        g = gz @ edge_w[k]
    # g is d[h_i | h_j | d_ij]
    dh += g[..., :H].sum(axis=2)            # h_i slot: sum over j
    dh += g[..., H:2 * H].sum(axis=1)       # h_j slot: sum over i
    dd = g[..., 2 * H]                      # (B,N,N)
    s = metric_signs(H, cache["metric"]).astype(h.dtype)
    gd = 2.0 * dd[..., None] * cache["diff"] * s   # d d_ij / d diff
    dh += gd.sum(axis=1)                    # diff = h_j - h_i + eps : + on j
    dh -= gd.sum(axis=2)                    #                          - on i
    return dh, d_edge_w, d_edge_b, d_node_w, d_node_b


# --------------------------------------------------------------------------------------
# GraphNet (graphnet.py:136-171)
# --------------------------------------------------------------------------------------
def _alphas(alphas, num_mps):
    return adjust_var_list(alphas, num_mps)


def graphnet_forward(x, params, prefix, num_mps, input_node_size, first_width, alphas, metric="euclidean"):
    """x (B,N,F) -> (B,N,out).  Zero-pads (or crops: negative pad) the features to H_0
    (graphnet.py:152), then runs ``num_mps`` steps."""
    pad = first_width - input_node_size
    if pad >= 0:
        h = np.concatenate([x, np.zeros(x.shape[:-1] + (pad,), dtype=x.dtype)], axis=-1)
    else:
        h = x[..., :pad]
    al = _alphas(alphas, num_mps)
    caches = []
    for t in range(num_mps):
        ew, eb = _stack(params, prefix, "edge_net", t)
        nw, nb = _stack(params, prefix, "node_net", t)
        h, c = mp_step_forward(h, ew, eb, nw, nb, al[t], metric)
        caches.append(c)
    return h, dict(caches=caches, pad=pad, in_width=x.shape[-1])


def graphnet_backward(dy, cache, params, prefix, num_mps):
    grads = {}
    g = dy
    for t in reversed(range(num_mps)):
        ew, _ = _stack(params, prefix, "edge_net", t)
        nw, _ = _stack(params, prefix, "node_net", t)
        g, dew, deb, dnw, dnb = mp_step_backward(g, cache["caches"][t], ew, nw)
        for k in range(len(ew)):
            grads[f"{prefix}edge_net.{t}.{k}.weight"] = dew[k]
            grads[f"{prefix}edge_net.{t}.{k}.bias"] = deb[k]
        for k in range(len(nw)):
            grads[f"{prefix}node_net.{t}.{k}.weight"] = dnw[k]
            grads[f"{prefix}node_net.{t}.{k}.bias"] = dnb[k]
    pad = cache["pad"]
    if pad >= 0:
        dx = g[..., : cache["in_width"]]
    else:
        dx = np.concatenate([g, np.zeros(g.shape[:-1] + (-pad,), dtype=g.dtype)], axis=-1)
    return dx, grads


# --------------------------------------------------------------------------------------
# Encoder / Decoder (encoder.py:133-171, decoder.py:119-136)
# --------------------------------------------------------------------------------------
def _norm_map(latent_map):
    return latent_map.lower().replace(" ", "_")


def encoder_out_width(latent_map, latent_node_size, node_sizes):
    """encoder.py:91 tests ``latent_map.lower() in LOCAL_MIX`` WITHOUT the space->underscore
    replacement applied everywhere else, so 'local mix' and 'local_mix' build different
    GraphNets; it also indexes the RAW node_sizes list (encoder.py:92)."""
    return node_sizes[-1][-1] if latent_map.lower() in LOCAL_MIX else latent_node_size


def encoder_forward(x, params, cfg, metric="euclidean"):
    """cfg keys: num_nodes,input_node_size,latent_node_size,node_sizes,edge_sizes,num_mps,alphas,latent_map."""
    ns = adjust_var_list(cfg["node_sizes"], cfg["num_mps"])
    y, gc = graphnet_forward(x, params, "encoder.", cfg["num_mps"], cfg["input_node_size"], ns[0][0],
                             cfg["alphas"], metric)
    B = x.shape[0]
    lm = _norm_map(cfg["latent_map"])
    cache = dict(g=gc, y=y, lm=lm)
    if lm == "max":
        idx = y.argmax(axis=-2)
        cache["idx"] = idx
        z = np.take_along_axis(y, idx[:, None, :], axis=-2)[:, 0, :]
    elif lm == "min":
        idx = y.argmin(axis=-2)
        cache["idx"] = idx
        z = np.take_along_axis(y, idx[:, None, :], axis=-2)[:, 0, :]
    elif lm in GLOBAL_MIX:
        z = y.reshape(B, -1) @ params["mix_layer.weight"].T          # bias=False (encoder.py:116-120)
    elif lm in LOCAL_MIX:
        z = (y @ params["mix_layer.weight"].T + params["mix_layer.bias"]).reshape(B, -1)
    else:  # 'mean' and every unknown spelling (encoder.py:147-149,162-168)
        cache["lm"] = "mean"
        z = y.mean(axis=-2)
    return z, cache


def encoder_backward(dz, cache, params, cfg):
    y = cache["y"]
    B, N, W = y.shape
    lm = cache["lm"]
    grads = {}
    if lm in ("max", "min"):
        dy = np.zeros_like(y)
        np.put_along_axis(dy, cache["idx"][:, None, :], dz[:, None, :], axis=-2)
    elif lm in GLOBAL_MIX:
        grads["mix_layer.weight"] = dz.T @ y.reshape(B, -1)
        dy = (dz @ params["mix_layer.weight"]).reshape(B, N, W)
    elif lm in LOCAL_MIX:
        g = dz.reshape(B, N, -1)
        grads["mix_layer.weight"] = np.einsum("bno,bni->oi", g, y)
        grads["mix_layer.bias"] = g.sum(axis=(0, 1))
        dy = g @ params["mix_layer.weight"]
    else:
        dy = np.broadcast_to(dz[:, None, :] / N, y.shape).copy()
    dx, gg = graphnet_backward(dy, cache["g"], params, "encoder.", cfg["num_mps"])
    grads.update(gg)
    return dx, grads


def decoder_forward(z, params, cfg, metric="euclidean"):
    """cfg keys: num_nodes,latent_node_size,output_node_size,node_sizes,edge_sizes,num_mps,alphas,
    latent_map,normalize_output.  decoder.py:127-136 then GraphNet then optional tanh (:123-124).
    Note decoder.py:90,96,104 index the RAW node_sizes list."""
    N = cfg["num_nodes"]
    h0 = cfg["node_sizes"][0][0]
    w, b = params["linear.weight"], params["linear.bias"]
    if _norm_map(cfg["latent_map"]) in LOCAL_MIX:
        zin = z.reshape(-1, N, cfg["latent_node_size"])
        x = zin @ w.T + b
    else:
        zin = z
        x = (z @ w.T + b).reshape(-1, N, h0)
    ns = adjust_var_list(cfg["node_sizes"], cfg["num_mps"])
    y, gc = graphnet_forward(x, params, "decoder.", cfg["num_mps"], h0, ns[0][0], cfg["alphas"], metric)
    out = np.tanh(y) if cfg.get("normalize_output", False) else y
    return out, dict(g=gc, zin=zin, out=out, zshape=z.shape)


def decoder_backward(dout, cache, params, cfg):
    if cfg.get("normalize_output", False):
        dout = dout * (1.0 - cache["out"] ** 2)
    dx, grads = graphnet_backward(dout, cache["g"], params, "decoder.", cfg["num_mps"])
    w = params["linear.weight"]
    zin = cache["zin"]
    if _norm_map(cfg["latent_map"]) in LOCAL_MIX:
        grads["linear.weight"] = np.einsum("bno,bni->oi", dx, zin)
        grads["linear.bias"] = dx.sum(axis=(0, 1))
        dz = (dx @ w).reshape(cache["zshape"])
    else:
        g = dx.reshape(dx.shape[0], -1)
        grads["linear.weight"] = g.T @ zin
        grads["linear.bias"] = g.sum(axis=0)
        dz = g @ w
    return dz, grads


# --------------------------------------------------------------------------------------
# Chamfer loss (chamfer_loss.py:11-42, distance_sq.py:4-77)
# --------------------------------------------------------------------------------------
def normsq_signs(dim, norm_choice):
    """distance_sq.py:43-44 forces 'cartesian' for 3-vectors; 'minkowskian' and 'polar'
    both compute 2*p0^2 - sum p^2 (:66-68,75-77)."""
    s = np.ones(dim)
    if dim != 3 and str(norm_choice).lower() in ("minkowskian", "polar"):
        s[1:] = -1.0
    return s


def jet_term_signs(dim, norm_choice):
    """The jet term calls normsq with the loss's norm choice directly (chamfer_loss.py:40): no 3-vector override, so
    3-vectors with 'minkowskian' / 'polar' get p0^2 - p1^2 - p2^2."""
    s = np.ones(dim)
    if str(norm_choice).lower() in ("minkowskian", "polar"):
        s[1:] = -1.0
    return s


def pairwise_distance_sq(p, q, norm_choice="cartesian"):
    """dist[b,i,j] = normsq(p[b,i] - q[b,j])   (distance_sq.py:46-54)."""
    if p.shape[0] != q.shape[0]:
        raise ValueError("batch sizes differ")
    if p.shape[-1] not in (3, 4) or q.shape[-1] not in (3, 4) or p.shape[-1] != q.shape[-1]:
        raise ValueError("p and q must both be 3- or 4-vectors")
    s = normsq_signs(p.shape[-1], norm_choice).astype(p.dtype)
    diff = p[:, :, None, :] - q[:, None, :, :]
    return (diff * diff * s).sum(-1)


def chamfer_terms(p, q, norm_choice="cartesian"):
    """Returns (chamfer_term, jet_term, dchamfer/dp, djet/dp).

    chamfer_term = sum_b [ sum_i min_j dist + sum_j min_i dist ]     chamfer_loss.py:30-35
    jet_term     = sum_b normsq(sum_i p_i - sum_i q_i)               chamfer_loss.py:37-40
    Gradients flow through the arg-mins only (torch.min backward picks one index)."""
    D = p.shape[-1]
    s = normsq_signs(D, norm_choice).astype(p.dtype)
    dist = pairwise_distance_sq(p, q, norm_choice)
    j_star = dist.argmin(axis=-1)            # (B,Np) nearest target for each recon particle
    i_star = dist.argmin(axis=-2)            # (B,Nq) nearest recon particle for each target
    cham = dist.min(axis=-1).sum() + dist.min(axis=-2).sum()
    B = p.shape[0]
    dp = np.zeros_like(p)
    bi = np.arange(B)[:, None]
    qn = q[bi, j_star]                       # (B,Np,D)
    dp += 2.0 * s * (p - qn)
    pn = p[bi, i_star]                       # (B,Nq,D)
    np.add.at(dp, (np.broadcast_to(bi, i_star.shape), i_star), 2.0 * s * (pn - q))
    jd = p.sum(axis=-2) - q.sum(axis=-2)     # (B,D)
    sj = jet_term_signs(D, norm_choice).astype(p.dtype)
    jet = (jd * jd * sj).sum()
    djet = np.broadcast_to((2.0 * sj * jd)[:, None, :], p.shape).copy()
    return cham, jet, dp, djet


def anomaly_chamfer(p, q, lorentz=False):
    """Per-particle anomaly score of utils/jet_analysis/anomaly_detection.py: ``chamfer`` (:482-488: diffs = p[:, :, None] -
    q[:, None, :], dist = L2 norm over the last axis, min over j plus min over i, added elementwise -> (B, N)) and, with
    ``lorentz``, ``chamfer_lorentz`` (:505-510: dist = E^2 - px^2 - py^2 - pz^2 of the differences, :401-403)."""
    diffs = p[:, :, None, :] - q[:, None, :, :]
    if lorentz:
        dist = diffs[..., 0] ** 2 - diffs[..., 1] ** 2 - diffs[..., 2] ** 2 - diffs[..., 3] ** 2
    else:
        dist = np.sqrt((diffs ** 2).sum(axis=-1))
    return dist.min(axis=-1) + dist.min(axis=-2)


def anomaly_hungarian(p, q, lorentz=False):
    """utils/jet_analysis/anomaly_detection.py hungarian (:537-547) / hungarian_lorentz (:579-590): per jet, cost = |p_i -
    q_j|_2 (torch.cdist) or the Lorentz norm squared of p_i - q_j; matching = scipy.optimize.linear_sum_assignment(cost)[1]
    (the reference's own solver); p_shuffle = p[matching]; score = sum over components of (p_shuffle - q)^2 -> (B, N).
    Returns (scores, matching (B, N), total assignment cost (B,))."""
    from scipy import optimize
    diffs = p[:, :, None, :] - q[:, None, :, :]
    if lorentz:
        cost = diffs[..., 0] ** 2 - diffs[..., 1] ** 2 - diffs[..., 2] ** 2 - diffs[..., 3] ** 2
    else:
        cost = np.sqrt((diffs ** 2).sum(axis=-1))
    matching = np.stack([optimize.linear_sum_assignment(c)[1] for c in cost])
    bi = np.arange(p.shape[0])[:, None]
    p_shuffle = p[bi, matching]
    total = cost[bi, np.arange(p.shape[1])[None, :], matching].sum(axis=1)
    return ((p_shuffle - q) ** 2).sum(axis=-1), matching, total


def hungarian_coords(recons, target, abs_coord=True, polar_coord=False):
    """The coordinate options of utils/losses/hungarian_mse/hungarian_mse.py:60-100 (with utils.py:8-71 of the same package,
    including its ``py = pt * cos(phi)`` line :63), numpy float64."""
    eps = 1e-16

    def polar(p):
        x, y, z = p[..., -3], p[..., -2], p[..., -1]
        pt = np.sqrt(x * x + y * y + eps)
        return np.stack((pt, np.arcsinh(z / (pt + eps)), np.arctan2(y + eps, x + eps)), axis=-1)

    if abs_coord:
        return (polar(recons), polar(target)) if polar_coord else (recons, target)
    jet = polar(target.sum(axis=-2))[:, None, :]

    def rel(p):
        pp = polar(p)
        return np.stack((pp[..., 0] / (jet[..., 0] + eps), pp[..., 1] - jet[..., 1],
                         np.mod(pp[..., 2] - jet[..., 2] + np.pi, 2 * np.pi) - np.pi), axis=-1)

    r, t = rel(recons), rel(target)
    if polar_coord:
        return r, t
    cart = lambda p: np.stack((p[..., 0] * np.cos(p[..., 2]), p[..., 0] * np.cos(p[..., 2]), p[..., 0] * np.sinh(p[..., 1])), axis=-1)
    return cart(r), cart(t)


def hungarian_mse_loss(recons, target, abs_coord=True, polar_coord=False):
    """hungarian_mse.py:22-58: Euclidean-cost optimal matching in the chosen coordinates, then the mean squared error."""
    r, t = hungarian_coords(recons, target, abs_coord, polar_coord)
    _, matching, _ = anomaly_hungarian(r, t)
    bi = np.arange(r.shape[0])[:, None]
    return float(((r[bi, matching] - t) ** 2).mean())


def chamfer_loss(p, q, norm_choice="cartesian", jet_features_weight=1.0, mode="intended"):
    """mode='intended': chamfer + w*jet (what chamfer_loss.py:35-41 computes and then discards);
    mode='reference': the value the reference actually RETURNS, ``jet_loss`` alone
    (chamfer_loss.py:42; w == 0 raises UnboundLocalError there).  Returns (loss, dloss/dp)."""
    cham, jet, dcham, djet = chamfer_terms(p, q, norm_choice)
    if mode == "reference":
        if jet_features_weight == 0:
            raise UnboundLocalError("jet_loss referenced before assignment (chamfer_loss.py:42)")
        return jet, djet
    return cham + jet_features_weight * jet, dcham + jet_features_weight * djet


# --------------------------------------------------------------------------------------
# training step (utils/train.py:51-85, 330-385; utils/initialize.py:152-153)
# --------------------------------------------------------------------------------------
def l1_norm(params):
    """encoder.py:173-175 / decoder.py:138-140."""
    return sum(np.abs(v).sum() for v in params.values())


POLAR_EPS = 1e-16      # utils/const.py:5


def polar_clamp(y):
    """utils/train.py:55-65: (E, pT) of 4-vectors / pT of 3-vectors clamped from below at EPS.  Returns (clamped, d clamped / d y)."""
    k = 2 if y.shape[-1] == 4 else 1
    mask = np.ones_like(y)
    mask[..., :k] = (y[..., :k] >= POLAR_EPS).astype(y.dtype)      # torch.clamp passes the gradient where min <= y
    out = y.copy()
    out[..., :k] = np.maximum(y[..., :k], POLAR_EPS)
    return out, mask


def mse_loss(p, q):
    """nn.MSELoss() (utils/train.py:359-361): mean over all elements."""
    d = p - q
    return float((d * d).mean()), 2.0 * d / d.size


def loss_and_grads(x, enc_params, dec_params, enc_cfg, dec_cfg, *, metric="euclidean",
                   loss_norm_choice="cartesian", jet_features_weight=1.0, chamfer_mode="intended",
                   l1_lambda=1e-8, l2_lambda=0.0, loss_choice="chamfer", polar_coord=False):
    """encoder -> decoder [-> polar clamp] -> Chamfer or MSE (+L1/L2 regularisers, train.py:51-65,338-384) and all gradients."""
    z, ec = encoder_forward(x, enc_params, enc_cfg, metric)
    y, dc = decoder_forward(z, dec_params, dec_cfg, metric)
    y_loss, mask = polar_clamp(y) if polar_coord else (y, None)
    if loss_choice == "mse":
        loss, dy = mse_loss(y_loss, x)
    else:
        loss, dy = chamfer_loss(y_loss, x, loss_norm_choice, jet_features_weight, chamfer_mode)
    if mask is not None:
        dy = dy * mask
    dz, dgrads = decoder_backward(dy, dc, dec_params, dec_cfg)
    _, egrads = encoder_backward(dz, ec, enc_params, enc_cfg)
    for params, grads in ((enc_params, egrads), (dec_params, dgrads)):
        for k, v in params.items():
            g = grads.get(k)
            if g is None:
                g = np.zeros_like(v)
            if l1_lambda > 0:
                g = g + l1_lambda * np.sign(v)
            if l2_lambda > 0:
                g = g + l2_lambda * 2.0 * v
            grads[k] = g
    if l1_lambda > 0:
        loss = loss + l1_lambda * (l1_norm(enc_params) + l1_norm(dec_params))
    if l2_lambda > 0:
        loss = loss + l2_lambda * sum((v * v).sum() for p in (enc_params, dec_params) for v in p.values())
    return loss, z, y_loss, egrads, dgrads      # y_loss: the reconstruction as the loss sees it (after the polar clamp)


def adam_update(params, grads, state, lr=1e-5, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam defaults (initialize.py:152-153): no weight decay, no amsgrad.
    ``state`` maps name -> (step, m, v); updated in place, params updated in place."""
    b1, b2 = betas
    for k, p in params.items():
        g = grads[k]
        step, m, v = state.get(k, (0, np.zeros_like(p), np.zeros_like(p)))
        step += 1
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        bc1 = 1 - b1 ** step
        bc2 = 1 - b2 ** step
        denom = np.sqrt(v) / np.sqrt(bc2) + eps
        p -= (lr / bc1) * m / denom
        state[k] = (step, m, v)


def train_step(x, enc_params, dec_params, enc_cfg, dec_cfg, enc_state, dec_state, lr=1e-5, **kw):
    """One optimisation step as in utils/train.py:51-85 with two independent Adams."""
    loss, z, y, eg, dg = loss_and_grads(x, enc_params, dec_params, enc_cfg, dec_cfg, **kw)
    adam_update(enc_params, eg, enc_state, lr)
    adam_update(dec_params, dg, dec_state, lr)
    return loss, z, y


# --------------------------------------------------------------------------------------
# synthetic JetNet-shaped jets (SURVEY.md 8.d recipe) -- shared by tests and bench
# --------------------------------------------------------------------------------------
def synthetic_jets(batch, num_particles, seed=1234, dtype=np.float32):
    """(B,N,3) [pt_rel, eta_rel, phi_rel]: eta,phi ~ N(0,0.1^2) clipped to +-0.5; pt_rel ~
    Dirichlet(0.5) sorted descending; n ~ UniformInt[N/3, N] real particles, the rest zero."""
    rng = np.random.default_rng(seed)
    eta = np.clip(rng.normal(0.0, 0.1, (batch, num_particles)), -0.5, 0.5)
    phi = np.clip(rng.normal(0.0, 0.1, (batch, num_particles)), -0.5, 0.5)
    pt = rng.dirichlet(np.full(num_particles, 0.5), size=batch)
    pt = -np.sort(-pt, axis=1)
    n = rng.integers(max(1, num_particles // 3), num_particles + 1, size=batch)
    mask = np.arange(num_particles)[None, :] < n[:, None]
    x = np.stack([pt, eta, phi], axis=-1) * mask[..., None]
    return x.astype(dtype)


# --------------------------------------------------------------------------------------
# permutation property (utils/permutation.py:76-109): reported deviation, no threshold
# --------------------------------------------------------------------------------------
def relative_deviation(output, target):
    """utils/permutation.py:8,107-109 (its own EPS = 1e-12, not utils/const.EPS)."""
    return np.abs(output - target) / (np.abs(target) + 1e-12)
