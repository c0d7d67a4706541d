"""Developer check of one message-passing step (bf16 tensor-core kernels) against the numpy oracle + per-kernel timing.
Usage on the GPU box: python tools/dev_mp.py [fwd|all] [--time]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gnn_jet_autoencoder_b200 import ops
from oracle import gnnae_oracle as O

DEV = "cuda:0"
what = sys.argv[1] if len(sys.argv) > 1 else "all"


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30))


def make(N, H, edge, node, B, seed):
    rng = np.random.default_rng(seed)
    shapes_e = [(o, i) for i, o in zip([2 * H + 1] + edge[:-1], edge)]
    shapes_n = [(o, i) for i, o in zip([edge[-1] + H] + node[:-1], node)]
    ew = [rng.uniform(-1, 1, s) / np.sqrt(s[1]) for s in shapes_e]
    eb = [rng.uniform(-1, 1, s[0]) / np.sqrt(s[1]) for s in shapes_e]
    nw = [rng.uniform(-1, 1, s) / np.sqrt(s[1]) for s in shapes_n]
    nb = [rng.uniform(-1, 1, s[0]) / np.sqrt(s[1]) for s in shapes_n]
    h = rng.normal(0, 0.5, (B, N, H))
    dy = rng.normal(0, 1.0, (B, N, node[-1]))
    return ew, eb, nw, nb, h, dy


def check(N, H, B, alpha=0.2, edge=(32, 128, 64, 16), node=(16, 32)):
    edge, node = list(edge), list(node)
    ew, eb, nw, nb, h, dy = make(N, H, edge, node, B, N * 100 + H)
    y_ref, cache = O.mp_step_forward(h, ew, eb, nw, nb, alpha, "euclidean")
    pack = lambda ws, bs: [np.concatenate([w.ravel(), b.ravel()]) for w, b in zip(ws, bs)]
    flat = torch.from_numpy(np.concatenate(pack(ew, eb) + pack(nw, nb))).float().to(DEV)
    ht = torch.from_numpy(h).float().to(DEV)
    args = (N, H, edge, node, alpha, 0, ops.PRECISIONS["bf16"])
    y, e = torch.ops.gnnjet.mp_step_fwd(ht, flat, *args)
    torch.cuda.synchronize()
    e_ref = O.leaky(cache["edge_z"][-1], alpha).sum(axis=2)
    msg = f"N={N} H={H} B={B} alpha={alpha}: e {rel(e.cpu().numpy(), e_ref):.2e} y {rel(y.cpu().numpy(), y_ref):.2e}"
    if what == "all":
        dh_ref, dew, deb, dnw, dnb = O.mp_step_backward(dy, cache, ew, nw)
        gref = np.concatenate(pack(dew, deb) + pack(dnw, dnb))
        dh, dflat = torch.ops.gnnjet.mp_step_bwd(ht, e, flat, torch.from_numpy(dy).float().to(DEV), *args)
        torch.cuda.synchronize()
        g = dflat.cpu().numpy()
        msg += f" dh {rel(dh.cpu().numpy(), dh_ref):.2e} dflat {rel(g, gref):.2e}"
        # per-tensor breakdown of the edge parameter gradients
        off = 0
        parts = []
        for l, (w, b) in enumerate(zip(dew, deb)):
            parts.append(f"W{l} {rel(g[off:off + w.size], w.ravel()):.1e}"); off += w.size
            parts.append(f"b{l} {rel(g[off:off + b.size], b.ravel()):.1e}"); off += b.size
        msg += " [" + " ".join(parts) + "]"
    print(msg, flush=True)


for (N, H, B) in [(30, 16, 3), (30, 3, 9), (5, 4, 7), (33, 8, 5), (150, 32, 2), (1, 3, 4), (64, 6, 3)]:
    check(N, H, B)
check(12, 8, 3, alpha=0.0)
check(12, 8, 3, alpha=1.0)

if "--time" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    for (N, B) in [(30, 4096), (150, 512)]:
        H, edge, node = 16, [32, 128, 64, 16], [16, 32]
        ew, eb, nw, nb, h, dy = make(N, H, edge, node, 4, 1)
        pack = lambda ws, bs: [np.concatenate([w.ravel(), b.ravel()]) for w, b in zip(ws, bs)]
        flat = torch.from_numpy(np.concatenate(pack(ew, eb) + pack(nw, nb))).float().to(DEV)
        ht = (torch.randn(B, N, H, device=DEV) * 0.5)
        dyt = torch.randn(B, N, node[-1], device=DEV)
        args = (N, H, edge, node, 0.2, 0, ops.PRECISIONS["bf16"])
        for _ in range(3):
            y, e = torch.ops.gnnjet.mp_step_fwd(ht, flat, *args)
            if what == "all": torch.ops.gnnjet.mp_step_bwd(ht, e, flat, dyt, *args)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                y, e = torch.ops.gnnjet.mp_step_fwd(ht, flat, *args)
                if what == "all": torch.ops.gnnjet.mp_step_bwd(ht, e, flat, dyt, *args)
            torch.cuda.synchronize()
        print(f"--- N={N} B={B}: kernel times (us, avg of 5)")
        tiles = B * N * ((N + 31) // 32) / 4
        for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
            if ev.device_time_total <= 0: continue
            us = ev.device_time_total / ev.count
            extra = ""
            if "edge" in ev.key:
                extra = f"  {us * 1e-6 * 1.9e9 / (tiles / 148):.0f} clk/tile/SM @1.9GHz"
            print(f"{us:10.1f} x{ev.count:<3d} {ev.key[:90]}{extra}")
