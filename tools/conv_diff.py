"""Which workspace buffers differ between the list conv kernels (conv_impl 0) and the dense walk (conv_impl 1)?"""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import golden_state_dict, load_golden
from aimnetcentral_b200 import AIMNet2Calculator

inputs, ref, meta = load_golden(sys.argv[1] if len(sys.argv) > 1 else "mols_8x50")
sd, spec = golden_state_dict(meta)
calc = AIMNet2Calculator((sd, spec), device="cuda:0")
calc.engine.set_small_m_rows(512)
snaps = {}
for impl in (0, 1):
    calc.engine.set_conv_impl(impl)
    out = calc(dict(inputs), forces=True)
    snaps[impl] = (calc.engine.debug_snapshot(), {k: v.cpu().numpy() for k, v in out.items()})
N = len(inputs["numbers"])
a, b = snaps[0][0], snaps[1][0]
for name in ("a0", "T_a0", "y0", "a1", "q0", "T_a1", "y1", "a2", "q1", "T_a2", "T_q1", "T_q2", "aim", "e_nn", "dx", "dS_a", "grad_a", "grad_q", "da_tot", "dq"):
    if name not in a or name not in b:
        continue
    dt = np.float64 if name.startswith("e_") else np.float32
    x, y = a[name].view(dt), b[name].view(dt)
    d = np.abs(x - y)
    per = len(x) // N
    rows = np.nonzero((d.reshape(N, per) > 0).any(axis=1))[0] if per * N == len(x) else []
    cols = np.nonzero((d.reshape(N, per) > 0).any(axis=0))[0] if per * N == len(x) else []
    print(f"{name:8s} equal={np.array_equal(x, y)} max|d|={d.max():.3e} rel={d.max() / (np.abs(x).max() + 1e-30):.2e} rows differing {len(rows)} first {list(rows[:6])} cols {len(cols)} first {list(cols[:8])}")
for k in snaps[0][1]:
    print(k, np.abs(snaps[0][1][k] - snaps[1][1][k]).max())
