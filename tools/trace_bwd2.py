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
d = _lib.make_desc(B, N, H, edge, node, 0.2, 0, ops.PRECISIONS["bf16"], h_ld=H)
lib = _lib.load()
alloc = lambda nbytes: torch.empty((max(int(nbytes), 4) + 3) // 4, device="cuda")
saved, wsf, wsb = alloc(lib.gj_mp_step_saved_bytes(d)), alloc(lib.gj_mp_step_fwd_workspace(d)), alloc(lib.gj_mp_step_bwd_workspace(d))
y, e = torch.empty(B, N, node[-1], device="cuda"), torch.empty(B, N, edge[-1], device="cuda")
dh, dflat = torch.empty_like(h), torch.empty_like(flat)
st = torch.cuda.current_stream().cuda_stream
use_saved = os.environ.get("GJ_TRACE_UNSAVED", "0") == "0"      # the trainer's path: forward by-products saved for the backward call
for _ in range(2):
    ops.raw_mp_fwd(d, h.data_ptr(), flat.data_ptr(), y.data_ptr(), e.data_ptr(), wsf.data_ptr(), wsf.numel() * 4, st, saved.data_ptr() if use_saved else None)
    ops.raw_mp_bwd(d, h.data_ptr(), e.data_ptr(), flat.data_ptr(), dy.data_ptr(), dh.data_ptr(), dflat.data_ptr(), wsb.data_ptr(), wsb.numel() * 4, st,
                   saved.data_ptr() if use_saved else None)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (24 * 128))()
assert lib.gj_debug_read_bwd2_trace(buf) == 0
t = np.array(buf[:], dtype=np.int64).reshape(128, 24)
names = ["F1 wait", "epi F1 + publish", "F2 wait", "epi F2 + publish", "F3 wait", "epi F3 + publish", "B3 wait", "epi B3 + publish",
         "B2 wait", "epi B2 + publish", "B1 wait", "epi B1 + refill + done2 wait", "L0 (next tile) + publish", "prefetch", "loop"]
rows = [r for r in range(2, 60) if t[r + 1, 0] > 0]
d = np.zeros((len(rows), 15))
for n, r in enumerate(rows):
    for s in range(14): d[n, s] = t[r, s + 1] - t[r, s]
    d[n, 14] = t[r + 1, 0] - t[r, 14]
print("tile stage durations (cycles), tiles", rows[0], "..", rows[-1])
for s in range(15): print(f"  {names[s]:24s} mean {d[:, s].mean():8.0f}  min {d[:, s].min():6.0f}  max {d[:, s].max():6.0f}")
print(f"  total per tile           mean {d.sum(1).mean():8.0f}")


fine = [("epi B1", 11, 15), ("refill", 15, 16), ("done2 wait", 16, 12), ("L0 math + STS", 12, 17), ("fence.proxy.async", 17, 18), ("group barrier", 18, 19), ("issue F1", 19, 13)]
for nm, a, b in fine:
    v = np.array([t[r, b] - t[r, a] for r in rows])
    print(f"    {nm:22s} mean {v.mean():8.0f}  min {v.min():6.0f}  max {v.max():6.0f}")
