"""Timeline of the two-stream 3xFP16 GEMM (backend 4, CTA 0): per stream, K-loop span, epilogue span and chunk period.

    python tools/gemm_trace2.py [N] [K] [mode] [M] [backend 4|5]
Events per stream: 1 MMA got free TMEM buffer, 2 MMA got operands (chunk start), 5 epilogue got chunk, 6 drained, 7 tile done."""
import ctypes as C, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from aimnetcentral_b200 import _capi
lib = _capi.load()
a = [int(x) for x in sys.argv[1:]] + [None] * 5
N, K, mode, M, BE = a[0] or 512, a[1] or 704, 2 if a[2] is None else a[2], a[3] or 51200, a[4] or 4
dev = "cuda:0"
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.05; b = torch.randn(N, device=dev)
Y = torch.empty(M, N, device=dev); aux = torch.randn(M, N, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
L = 2048
buf = torch.zeros(16 * L, dtype=torch.int64, device=dev)
def run():
    rc = lib.aimnet2_gemm_nt(A.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Y.data_ptr(), N, aux.data_ptr(), N, M, N, K, mode, BE, st)
    assert rc == 0, lib.aimnet2_last_error()
for _ in range(3): run()
torch.cuda.synchronize()
lib.aimnet2_gemm_set_trace(C.c_void_p(buf.data_ptr()))
run(); torch.cuda.synchronize()
lib.aimnet2_gemm_set_trace(None)
t = buf.cpu().numpy().reshape(2, 8, L).astype(np.int64)
nchunk = (K // 32 + 1) // 2
t0 = t[0, 2, 0]
end = max(t[s, 7][t[s, 7] > 0].max() for s in range(2) if (t[s, 7] > 0).any())
print(f"backend {BE} N={N} K={K} mode={mode} M={M}: {nchunk} chunks per tile; CTA 0 total {end - t0} clk")
for s in range(2):
    nc = int((t[s, 2] > 0).sum())
    nt = nc // nchunk
    print(f"stream {s}: {nt} tiles")
    for ti in range(nt):
        c0, c1 = ti * nchunk, (ti + 1) * nchunk - 1
        k_start, k_end = t[s, 2, c0] - t0, t[s, 5, c1] - t0   # first operands -> last chunk complete
        e_end = t[s, 7, c1] - t0
        drained = t[s, 6, c1] - t0
        per = np.diff(t[s, 5, c0:c1 + 1]).mean() if nchunk > 1 else 0
        print(f"   tile {ti}: K loop {k_start:7d} .. {k_end:7d} ({k_end - k_start:6d} clk, {per:5.0f} clk/chunk)   "
              f"epilogue {drained:7d} .. {e_end:7d} ({e_end - drained:6d} clk)")
print("stream 0, first stages: producer got slot | issue time of the 4 TMA ops | issued -> MMA has operands | MMA stage period")
for i in range(0, min(30, int((t[0, 4] > 0).sum()))):
    print(f"   stage {i:3d}: slot {t[0, 0, i] - t0:7d}   issue {t[0, 3, i] - t[0, 0, i]:5d}   load {t[0, 4, i] - t[0, 3, i]:6d}   "
          f"period {t[0, 4, i] - t[0, 4, i - 1] if i else 0:6d}")
