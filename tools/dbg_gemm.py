import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_jet_autoencoder_b200 import _lib
lib = _lib.load()
form, M, N, K, prec = [int(v) for v in sys.argv[1:6]]
dev = torch.device("cuda", 0)
A = torch.randn(M, K, device=dev) if form < 2 else torch.randn(K, M, device=dev)
B = torch.randn(N, K, device=dev) if form == 0 else torch.randn(K, N, device=dev)
C = torch.zeros(M, N, device=dev)
wsb = lib.gj_dense_gemm_workspace(form, M, N, K)
ws = torch.empty(wsb // 4 + 16, device=dev)
st = torch.cuda.current_stream().cuda_stream
rc = lib.gj_dense_gemm(form, M, N, K, A.data_ptr(), B.data_ptr(), None, 0, 0.2, None, 0, C.data_ptr(), ws.data_ptr(), wsb, prec, st)
print("rc", rc, _lib.last_error())
torch.cuda.synchronize()
if prec == 1:
    A, B = A.to(torch.bfloat16).float(), B.to(torch.bfloat16).float()
ref = (A.double() @ B.double().T) if form == 0 else (A.double() @ B.double()) if form == 1 else (A.double().T @ B.double())
print("err", ((C.double() - ref).norm() / ref.norm()).item())
