"""Stage timeline of the second-generation backward edge kernel (GJ_TRACE=4 is set here; run on the GPU box)."""
import os, sys, ctypes
os.environ["GJ_TRACE"] = "4"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_jet_autoencoder_b200 import ops, _lib
N, B, H, edge, node = 30, 4096, 16, [32, 128, 64, 16], [16, 32]
npar = sum(o * i + o for i, o in zip([2 * H + 1] + edge[:-1], edge)) + sum(o * i + o for i, o in zip([edge[-1] + H] + node[:-1], node))
flat = (torch.rand(npar, device="cuda") - 0.5) * 0.3
h = torch.randn(B, N, H, device="cuda") * 0.5
dy = torch.randn(B, N, node[-1], device="cuda")
args = (N, H, edge, node, 0.2, 0, ops.PRECISIONS["bf16"])
for _ in range(2):
    y, e = torch.ops.gnnjet.mp_step_fwd(h, flat, *args)
    torch.ops.gnnjet.mp_step_bwd(h, e, flat, dy, *args)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (16 * 128))()
lib = _lib.load()
assert lib.gj_debug_read_bwd2_trace(buf) == 0
t = np.array(buf[:], dtype=np.int64).reshape(128, 16)
names = ["refill + done2 wait", "L0 + publish", "F1 wait", "epi F1 + publish", "F2 wait", "epi F2 + publish", "F3 wait", "epi F3 + publish",
         "B3 wait", "epi B3 + publish", "B2 wait", "epi B2 + publish", "B1 wait", "epi B1", "loop"]
rows = [r for r in range(2, 60) if t[r + 1, 0] > 0]
d = np.zeros((len(rows), 15))
for n, r in enumerate(rows):
    for s in range(14): d[n, s] = t[r, s + 1] - t[r, s]
    d[n, 14] = t[r + 1, 0] - t[r, 14]
print("tile stage durations (cycles), tiles", rows[0], "..", rows[-1])
for s in range(15): print(f"  {names[s]:24s} mean {d[:, s].mean():8.0f}  min {d[:, s].min():6.0f}  max {d[:, s].max():6.0f}")
print(f"  total per tile           mean {d.sum(1).mean():8.0f}")

