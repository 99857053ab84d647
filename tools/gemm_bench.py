"""Time the MLP GEMM shapes of cfg-2 (M=51200) per backend / epilogue mode through the aimnet2_gemm_nt seam; report TFLOP/s
(1x flops).  `python tools/gemm_bench.py 2,3,18,19` compares the engine's 3xFP16 kernel (2) with the experimental pipelined
epilogue (3); +16 = pre-split output.  For backends >= 2 the seam also splits W and A on the device and (with +16) expands
the result again, the same overhead for 2 and 3: use ncu on this script for kernel-only times."""
import ctypes as C, sys, os
import torch
sys.path.insert(0, ".")
from aimnetcentral_b200 import _capi
lib = _capi.load()
dev = "cuda:0"
M = 51200
shapes = [(512, 704, 2), (384, 512, 2), (288, 384, 1), (512, 384, 3), (736, 512, 0), (256, 384, 2), (128, 256, 2), (704, 512, 3)]
backends = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1").split(",")]
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for N, K, mode in shapes:
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.05; b = torch.randn(N, device=dev)
    Y = torch.empty(M, N, device=dev); aux = torch.randn(M, N, device=dev)
    from aimnetcentral_b200 import _capi as cap
    for be in backends:
        # pre-split weights are re-made inside aimnet2_gemm_nt for backend 1 (small, included in the timing)
        for _ in range(3):
            lib.aimnet2_gemm_nt(A.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Y.data_ptr(), N, aux.data_ptr(), N, M, N, K, mode | (be & 16), be & 15, st)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            rc = lib.aimnet2_gemm_nt(A.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Y.data_ptr(), N, aux.data_ptr(), N, M, N, K, mode | (be & 16), be & 15, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        assert rc == 0, lib.aimnet2_last_error()
        print(f"backend={be:2d} N={N:4d} K={K:4d} mode={mode} {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
