"""Run one MLP GEMM shape a few times (ncu target).  python tools/gemm_one.py [backend] [N] [K] [mode] [M]"""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from aimnetcentral_b200 import _capi
lib = _capi.load()
a = [int(x) for x in sys.argv[1:]] + [None] * 5
be, N, K, mode, M = a[0] or 2, a[1] or 512, a[2] or 704, 2 if a[3] is None else a[3], a[4] or 51200
dev = "cuda:0"
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.05; b = torch.randn(N, device=dev)
Y = torch.empty(M, N, device=dev); aux = torch.randn(M, N, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(5):
    rc = lib.aimnet2_gemm_nt(A.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Y.data_ptr(), N, aux.data_ptr(), N, M, N, K, mode, be, st)
    assert rc == 0, lib.aimnet2_last_error()
torch.cuda.synchronize()
