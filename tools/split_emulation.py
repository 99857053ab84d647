"""CPU emulation of operand-split GEMM schemes (exact accumulation in fp64): 3xTF32 vs 3xFP16 with per-(row, K-chunk)
power-of-two scaling.  Justifies the tensor-core precision scheme of gemm_tc.cu before any GPU time is spent.

    python tools/split_emulation.py
"""
import torch

torch.manual_seed(0)
M, N, K, CH = 512, 384, 704, 32


def rn_tf32(x):
    i = x.view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def pow2_scale(m, target_exp=13):
    """2^(target_exp - floor(log2 m)) for m > 0 (row-chunk max), 1 for m == 0."""
    e = torch.floor(torch.log2(m.clamp_min(1e-38)))
    s = torch.exp2(target_exp - e)
    return torch.where(m > 0, s, torch.ones_like(s))


def split_fp16(x, s):
    xs = x * s
    hi = xs.half()
    lo = (xs - hi.float()).half()
    return hi, lo


def emu_tf32(A, W):
    Ah, Wh = rn_tf32(A), rn_tf32(W)
    Al, Wl = rn_tf32(A - Ah), rn_tf32(W - Wh)
    return Ah.double() @ Wh.double().T + Ah.double() @ Wl.double().T + Al.double() @ Wh.double().T


def emu_fp16(A, W):
    sw = pow2_scale(W.abs().max().reshape(1))
    Wh, Wl = split_fp16(W, sw)
    out = torch.zeros(A.shape[0], W.shape[0], dtype=torch.float64)
    for k0 in range(0, A.shape[1], CH):
        a = A[:, k0:k0 + CH]
        sa = pow2_scale(a.abs().amax(dim=1, keepdim=True))
        Ah, Al = split_fp16(a, sa)
        assert torch.isfinite(Ah.float()).all()
        wh, wl = Wh[:, k0:k0 + CH].double(), Wl[:, k0:k0 + CH].double()
        part = Ah.double() @ wh.T + Ah.double() @ wl.T + Al.double() @ wh.T
        out += part / (sa.double() * sw.double())
    return out


def report(name, A, W):
    ref = A.double() @ W.double().T
    scale = ref.abs().mean()
    for nm, f in (("3xTF32", emu_tf32), ("3xFP16 row-chunk scaled", emu_fp16), ("fp32 matmul", lambda a, w: (a @ w.T).double())):
        e = f(A, W) - ref
        print(f"{name:28s} {nm:26s} rms={float(e.pow(2).mean().sqrt() / scale):.3e} max={float(e.abs().max() / scale):.3e}")


A = torch.randn(M, K).abs() * 0.7 + 0.1 * torch.randn(M, K)
W = torch.randn(N, K) * 0.05
report("gelu-like activations", A, W)
# wide dynamic range inside rows and between rows (gradients: 1e-6 .. 1e2)
A2 = torch.randn(M, K) * torch.exp(torch.randn(M, K) * 3.0) * torch.exp(torch.randn(M, 1) * 6.0)
report("wide dynamic range", A2, W)
# tiny gradients
report("tiny values (1e-9)", A * 1e-9, W)
# outlier weights
W3 = W.clone()
W3[::7, ::13] *= 300.0
report("outlier weights", A, W3)
