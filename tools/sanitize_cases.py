"""Small evaluations covering every kernel family, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck   python tools/sanitize_cases.py
    compute-sanitizer --tool initcheck  python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck  python tools/sanitize_cases.py
"simt" as argument restricts the run to the small-M SIMT MLP path: initcheck does not see TMA bulk stores as writes, so
the tensor-core path floods it with false positives (everything downstream of a TMA-stored activation).
"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
from aimnetcentral_b200.structures import allose_supercell, random_molecules

which = sys.argv[1] if len(sys.argv) > 1 else "all"
spec = ModelSpec()
calc = AIMNet2Calculator((random_state_dict(0, spec), spec), device="cuda:0")
coord, numbers = random_molecules(6, 23, seed=5)
mol = {"coord": coord, "numbers": numbers, "charge": np.zeros(6, np.float32)}
z, x, cell = allose_supercell((1, 1, 1), jitter=0.02, seed=1)
pbc = {"coord": x, "numbers": z, "charge": np.zeros(1, np.float32), "cell": cell}
for tc in ((False,) if which == "simt" else (True, False)):
    calc.engine.set_small_m_rows(0 if tc else 512)   # tensor-core GEMMs / small-M SIMT GEMMs
    calc.set_lrcoulomb_method("simple")
    out = calc(dict(mol), forces=True)
    calc.set_lrcoulomb_method("dsf")
    out2 = calc(dict(pbc), forces=True, stress=True)
    if which in ("all", "simt"):
        calc.set_lrcoulomb_method("ewald")
        out3 = calc(dict(pbc), forces=True, stress=True)
    torch.cuda.synchronize()
    print("tensor-core" if tc else "small-m", float(out["energy"].sum()), float(out2["energy"].sum()))
spec2 = ModelSpec(num_charge_channels=2)
calc2 = AIMNet2Calculator((random_state_dict(1, spec2), spec2), device="cuda:0")
calc2.engine.set_small_m_rows(512 if which == "simt" else 0)
out4 = calc2({**mol, "mult": np.ones(6, np.float32)}, forces=True)
torch.cuda.synchronize()
print("nse", float(out4["energy"].sum()))
# round 2: the dense (shared-memory / TMA) conv forward needs a batch of >= 64 molecules; GEMM backends 4 and 5 through the
# operator seam (two tile streams; CTA pairs), a Hessian (finite differences over a molecule batch), batched Ewald
if which == "all":
    import ctypes as C
    from aimnetcentral_b200 import _capi
    coord64, numbers64 = random_molecules(64, 12, seed=9)
    calc.engine.set_small_m_rows(0)
    calc.set_lrcoulomb_method("simple")
    out5 = calc({"coord": coord64, "numbers": numbers64, "charge": np.zeros(64, np.float32)}, forces=True)
    print("dense conv batch", float(out5["energy"].sum()), calc.engine.conv_mode())
    H = calc({"coord": coord[0], "numbers": numbers[0], "charge": np.zeros(1, np.float32)}, hessian=True)["hessian"]
    print("hessian", tuple(H.shape), float(H.abs().max()))
    lib = _capi.load()
    M, N, K = 700, 288, 96
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") * 0.05; bias = torch.randn(N, device="cuda")
    Y = torch.empty(M, N, device="cuda"); aux = torch.randn(M, N, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for be in (2, 4, 5):
        for mode in (2, 3, 2 | 16):
            rc = lib.aimnet2_gemm_nt(A.data_ptr(), K, W.data_ptr(), K, bias.data_ptr(), Y.data_ptr(), N, aux.data_ptr(), N, M, N, K, mode, be, st)
            assert rc == 0, lib.aimnet2_last_error()
    torch.cuda.synchronize()
    print("gemm backends 2/4/5", float(Y.abs().sum()))
