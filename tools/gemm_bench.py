"""Timing of the generic dense-layer GEMM (gj_dense_gemm) at node-level shapes.  GPU box: python tools/gemm_bench.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_jet_autoencoder_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
def run(form, M, N, K, prec, acc=0, reps=20):
    A = torch.randn(M, K, device=dev) if form < 2 else torch.randn(K, M, device=dev)
    B = torch.randn(N, K, device=dev) if form == 0 else torch.randn(K, N, device=dev)
    C = torch.zeros(M, N, device=dev)
    wsb = lib.gj_dense_gemm_workspace(form, M, N, K)
    ws = torch.empty(wsb // 4 + 16, device=dev)
    f = lambda: lib.gj_dense_gemm(form, M, N, K, A.data_ptr(), B.data_ptr(), None, 1 if form == 0 else 0, 0.2, None, acc, C.data_ptr(), ws.data_ptr(), wsb, prec, st)
    for _ in range(3): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): f()
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / reps
    gb = (M * K + N * K + M * N * (2 if acc else 1)) * 4 / 1e9 if form < 2 else (K * (M + N)) * 4 / 1e9
    print(f"form {form} M={M:6d} N={N:4d} K={K:6d} prec={prec} acc={acc}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s  {gb / us * 1e6:7.0f} GB/s")
R = 61440
for prec in (1, 0):
    run(0, R, 128, 128, prec); run(0, R, 128, 256, prec); run(0, R, 64, 64, prec); run(0, R, 256, 256, prec); run(0, R, 16, 128, prec)
    run(0, R, 128, 128, prec, acc=1)
    run(1, R, 128, 128, prec); run(1, R, 256, 128, prec)
    run(2, 128, 128, R, prec); run(2, 128, 256, R, prec); run(2, 256, 512, R, prec)
    run(0, 1024, 128, 128, prec); run(0, 8192, 128, 128, prec)
