"""Error anatomy of the 3xTF32 tcgen05 GEMM: split error (emulated in fp64) vs accumulation error (measured)."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from aimnetcentral_b200 import _capi
lib = _capi.load()
dev = "cuda:0"
torch.manual_seed(0)
M, N, K = 4096, 512, 704
A = torch.randn(M, K, device=dev).abs() * 0.7 + 0.1 * torch.randn(M, K, device=dev)   # mostly positive like GELU outputs
W = torch.randn(N, K, device=dev) * 0.05
ref = A.double() @ W.double().T
def run(backend):
    Y = torch.empty(M, N, device=dev)
    rc = lib.aimnet2_gemm_nt(A.data_ptr(), K, W.data_ptr(), K, None, Y.data_ptr(), N, None, N, M, N, K, 0, backend,
                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.aimnet2_last_error()
    torch.cuda.synchronize()
    return Y.double()
def stats(name, Y):
    e = (Y - ref)
    scale = ref.abs().mean()
    print(f"{name:34s} rms={e.pow(2).mean().sqrt()/scale:.3e} max={e.abs().max()/scale:.3e} bias(mean signed/|ref|mean)={e.mean()/scale:+.3e} corr(e,ref)={(e*ref).mean()/ (ref*ref).mean():+.3e}")
def trunc(x): return (x.view(torch.int32) & ~0x1fff).view(torch.float32)
def rn(x):
    i = x.view(torch.int32); return ((i + 0x1000) & ~0x1fff).view(torch.float32)
for nm, f in (("trunc", trunc), ("rn", rn)):
    Ah, Wh = f(A), f(W); Al, Wl = f(A - Ah), f(W - Wh)
    emu = Ah.double() @ Wh.double().T + Ah.double() @ Wl.double().T + Al.double() @ Wh.double().T
    stats(f"emulated 3xTF32 split={nm} (exact acc)", emu)
stats("fp32 torch matmul (highest)", (A @ W.T).double())
stats("simt backend", run(0))
stats("tcgen05 3xTF32 backend", run(1))
stats("tcgen05 3xFP16 backend", run(2))
