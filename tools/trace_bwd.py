"""GJ_TRACE=1 python tools/trace_bwd.py : prints the issuer / warpgroup handshake timeline of one backward launch."""
import ctypes, os, sys
os.environ.setdefault("GJ_TRACE", "1")   # 1: backward kernel, 2: forward kernel
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_jet_autoencoder_b200 import _lib, ops
lib = _lib.load()
N, H, ew, nw, B = 30, 32, [32, 128, 64, 16], [32, 8], 4096
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
n = (2 * H + 1) * 32 + 32 + 32 * 128 + 128 + 128 * 64 + 64 + 64 * 16 + 16 + (16 + H) * 32 + 32 + 32 * 8 + 8
flat = (torch.rand(n, generator=g) - 0.5).mul(0.3).to(dev)
h = torch.randn(B, N, H, generator=g).mul(0.3).to(dev)
args = (N, H, ew, nw, 0.2, 0, 1)
y, e = torch.ops.gnnjet.mp_step_fwd(h, flat, *args)
dy = torch.randn_like(y)
for _ in range(2):
    if os.environ["GJ_TRACE"] == "2":
        torch.ops.gnnjet.mp_step_fwd(h, flat, *args)
    else:
        torch.ops.gnnjet.mp_step_bwd(h, e, flat, dy, *args)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 8192)()
cnt = (ctypes.c_int * 4)()
raw = ctypes.CDLL(_lib.LIB_PATH)
assert raw.gj_debug_read_trace(buf, cnt) == 0
ev = []
for who in range(3):
    for i in range(cnt[who]):
        ev.append((buf[who * 2048 + 2 * i + 1], who, buf[who * 2048 + 2 * i]))
ev.sort()
t0 = ev[0][0]
names = {10: "arrive(L0 done)", 60: "B1 epilogue done"}
limit = int(os.environ.get('GJ_TRACE_LINES', '260'))
for tcl, who, tag in ev[:limit]:
    who_s = ["WG0", "WG1", "ISS"][who]
    if who == 2:
        w, s = (tag % 1000) // 100, tag % 100
        desc = f"{'issued+commit' if tag >= 1000 else 'saw ready'} wg{w} stage{s}"
    else:
        k = tag // 10 * 10
        desc = {10: "arrive after layer0", 20: f"woke for F{tag-20} epilogue", 30: f"arrive after F{tag-30} epi", 40: f"woke for B{tag-40} epilogue",
                50: f"arrive after B{tag-50} epi", 60: "B1 epilogue done"}[k]
    print(f"{tcl - t0:9d}  {who_s}  {desc}")
