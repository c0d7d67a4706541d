"""Stage timeline of the second-generation forward edge kernel (run with GJ_TRACE=3 on the GPU box)."""
import os, sys, ctypes
os.environ["GJ_TRACE"] = "3"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_jet_autoencoder_b200 import ops, _lib
N, B, H, edge, node = 30, 4096, 16, [32, 128, 64, 16], [16, 32]
npar = sum(o * i + o for i, o in zip([2 * H + 1] + edge[:-1], edge)) + sum(o * i + o for i, o in zip([edge[-1] + H] + node[:-1], node))
flat = (torch.rand(npar, device="cuda") - 0.5) * 0.3
h = torch.randn(B, N, H, device="cuda") * 0.5
for _ in range(2):
    torch.ops.gnnjet.mp_step_fwd(h, flat, N, H, edge, node, 0.2, 0, ops.PRECISIONS["bf16"])
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (8 * 256))()
lib = _lib.load()
assert lib.gj_debug_read_fwd2_trace(buf) == 0
t = np.array(buf[:], dtype=np.int64).reshape(256, 8)
names = ["bar+issue MMA1", "MMA1 wait", "epi1", "bar+issue MMA2+L0(next)", "MMA2 wait (rest)", "epi2+bar+issue MMA3", "MMA3 wait", "epi3+loop"]
rows = [r for r in range(2, 60) if t[r + 1, 0] > 0]
d = np.zeros((len(rows), 8))
for n, r in enumerate(rows):
    for s in range(7): d[n, s] = t[r, s + 1] - t[r, s]
    d[n, 7] = t[r + 1, 0] - t[r, 7]
print("tile stage durations (cycles), tiles", rows[0], "..", rows[-1])
for s in range(8): print(f"  {names[s]:24s} mean {d[:, s].mean():8.0f}  min {d[:, s].min():6.0f}  max {d[:, s].max():6.0f}")
print(f"  total per tile           mean {d.sum(1).mean():8.0f}")
