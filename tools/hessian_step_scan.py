"""Finite-difference step of the Hessian (calculator.hessian_step): error against the reference's float64 Hessian."""
import sys, time
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import golden_state_dict, load_golden
from aimnetcentral_b200 import AIMNet2Calculator
for name in ("hessian_caffeine", "hessian_mols_3x12"):
    inputs, ref, meta = load_golden(name)
    H64 = np.load(f"tests/golden/{name}.npz")["ref64_hessian"]
    calc = AIMNet2Calculator(golden_state_dict(meta), device="cuda:0")
    data = {k: inputs[k] for k in ("coord", "numbers", "charge")}
    for m, h in [(2, 5e-4), (2, 1e-3), (2, 2e-3), (2, 4e-3), (4, 2e-3), (4, 4e-3), (4, 8e-3), (4, 1.6e-2), (4, 3.2e-2),
                 (6, 8e-3), (6, 1.6e-2), (6, 3.2e-2), (6, 6.4e-2)]:
        calc.hessian_step, calc.hessian_stencil = h, m
        calc(data, hessian=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        H = calc(data, hessian=True)["hessian"]
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        H = H.double().cpu().numpy()
        print(f"{name} stencil {m} step {h:8.1e}: max|H - H64| {np.abs(H - H64).max():.2e}  rms {np.sqrt(((H - H64) ** 2).mean()):.2e}  "
              f"vs ref fp32 {np.abs(H - ref['hessian']).max():.2e}   {dt * 1e3:.1f} ms")
