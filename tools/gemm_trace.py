"""Pipeline timeline of the 3xFP16 GEMM (CTA 0): where each warp role waits, per K=32 stage.

    python tools/gemm_trace.py [N] [K] [mode] [M]
Events: 0 TMA got empty slot, 1 MMA got free TMEM buffer, 2 MMA got operands, 5 epilogue got chunk,
6 epilogue drained chunk, 7 tile epilogue done."""
import ctypes as C, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from aimnetcentral_b200 import _capi
lib = _capi.load()
a = [int(x) for x in sys.argv[1:]] + [None] * 4
N, K, mode, M = a[0] or 512, a[1] or 704, 2 if a[2] is None else a[2], a[3] or 51200
dev = "cuda:0"
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.05; b = torch.randn(N, device=dev)
Y = torch.empty(M, N, device=dev); aux = torch.randn(M, N, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
L = 2048
buf = torch.zeros(8 * L, dtype=torch.int64, device=dev)
def run():
    rc = lib.aimnet2_gemm_nt(A.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Y.data_ptr(), N, aux.data_ptr(), N, M, N, K, mode, 2, st)
    assert rc == 0, lib.aimnet2_last_error()
for _ in range(3): run()
torch.cuda.synchronize()
lib.aimnet2_gemm_set_trace(C.c_void_p(buf.data_ptr()))
run(); torch.cuda.synchronize()
lib.aimnet2_gemm_set_trace(None)
t = buf.cpu().numpy().reshape(8, L).astype(np.int64)
nk = K // 32
n = int((t[2] > 0).sum())
t0 = t[0, 0]
print(f"N={N} K={K} mode={mode} M={M}: CTA 0 ran {n} stages ({n // nk} tiles of {nk}); total {(t[7][t[7] > 0].max() - t0)} clk")
names = ["tma:slot", "mma:tmem", "mma:split", "spl:data", "spl:done", "epi:chunk", "epi:drained", "epi:tile"]
# steady-state per-stage period and lags inside tile 1 .. last (skip the first tile)
lo, hi = nk, n
per = np.diff(t[2, lo:hi]).mean()
print(f"period per stage (MMA issue to MMA issue): {per:.0f} clk; MMA busy 768 clk")
def lag(a_, b_, sa=0, sb=0):
    x = t[b_, lo + sb:hi + sb] - t[a_, lo + sa:hi + sa]
    x = x[(t[b_, lo + sb:hi + sb] > 0) & (t[a_, lo + sa:hi + sa] > 0)]
    return x.mean(), np.percentile(x, 10), np.percentile(x, 90)
for name, (a_, b_) in {"TMA issue -> MMA has operands": (0, 2), "MMA got TMEM -> MMA got operands (MMA waits for TMA)": (1, 2),
                        "MMA issue -> epilogue got chunk": (2, 5), "epilogue got chunk -> drained": (5, 6)}.items():
    m, p10, p90 = lag(a_, b_)
    print(f"  {name:58s} mean {m:7.0f}  p10 {p10:7.0f}  p90 {p90:7.0f} clk")
# how long the MMA waits for a free TMEM buffer = (t1[c] - t2[c-1])
w = t[1, lo:hi] - t[2, lo - 1:hi - 1]
print(f"  MMA: previous issue -> got TMEM buffer                      mean {w.mean():7.0f}  p10 {np.percentile(w,10):7.0f}  p90 {np.percentile(w,90):7.0f} clk")
te = t[7][t[7] > 0]
idx = np.nonzero(t[7] > 0)[0]
d = t[7, idx] - t[6, idx]
print(f"  tile epilogue (last drain -> stores issued): mean {d.mean():.0f} clk over {len(d)} tiles")
tile_span = np.diff(t[7, idx])
print(f"  tile to tile: mean {tile_span.mean():.0f} clk")
