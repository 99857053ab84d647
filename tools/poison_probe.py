"""Does any kernel read workspace memory it has not written?  aimnet2_engine_debug_poison fills the engine's device
workspace with a byte pattern before every evaluation; results must stay bit-identical to an un-poisoned evaluation.
0xFF turns any read that matters into NaNs (fp32 and fp16) or index -1; 0x7B is "huge but finite" (fp32 1.3e36, fp16
61 280) and catches reads that would only feed the row-chunk scales of the 3xFP16 GEMM (lost precision, no NaN); 0x00 is
what a fresh cudaMalloc hands out on these boxes.

    python tools/poison_probe.py
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
from aimnetcentral_b200.structures import allose_supercell, random_molecules


def cases():
    coord, numbers = random_molecules(64, 50, seed=99)
    yield "mol 64x50", ModelSpec(), None, {"coord": coord, "numbers": numbers, "charge": np.zeros(64, np.float32)}, {}
    c1, n1 = random_molecules(3, 20, seed=4)
    yield "mol 3x20", ModelSpec(), None, {"coord": c1, "numbers": n1, "charge": np.zeros(3, np.float32)}, {}
    z, x, cell = allose_supercell((2, 2, 2), jitter=0.02, seed=1)
    pbc = {"coord": x, "numbers": z, "charge": np.zeros(1, np.float32), "cell": cell}
    yield "pbc dsf", ModelSpec(), "dsf", pbc, {"stress": True}
    yield "pbc ewald", ModelSpec(), "ewald", pbc, {"stress": True}
    c2, n2 = random_molecules(40, 30, seed=3)
    yield "nse 40x30", ModelSpec(num_charge_channels=2), None, {"coord": c2, "numbers": n2, "charge": np.zeros(40, np.float32),
                                                                 "mult": np.ones(40, np.float32)}, {}


bad = 0
for name, spec, coulomb, inp, kw in cases():
    calc = AIMNet2Calculator((random_state_dict(0, spec), spec), device="cuda:0")
    if coulomb:
        calc.set_lrcoulomb_method(coulomb)
    for rows in (0, 512):   # tensor-core MLPs / small-M SIMT MLPs (the latter only where the case is small enough)
        calc.engine.set_small_m_rows(rows)
        calc.engine.debug_poison(-1)
        ref = {k: v.detach().cpu().numpy().copy() for k, v in calc(dict(inp), forces=True, **kw).items() if torch.is_tensor(v)}
        for byte in (0xFF, 0x7B, 0x00, 0xFF):
            calc.engine.debug_poison(byte)
            out = {k: v.detach().cpu().numpy().copy() for k, v in calc(dict(inp), forces=True, **kw).items() if torch.is_tensor(v)}
            diffs = {k: (float(np.nanmax(np.abs(out[k].astype(np.float64) - ref[k]))) if out[k].size else 0.0,
                         bool(np.isnan(out[k]).any())) for k in ref}
            if any(d[0] != 0.0 or d[1] for d in diffs.values()):
                bad += 1
                print(f"{name} / small_m_rows {rows} / poison 0x{byte:02X}: DIFFERS " +
                      " ".join(f"{k}: {d[0]:.3e}{' NaN' if d[1] else ''}" for k, d in diffs.items()))
    print(f"{name}: done", flush=True)
print(f"poison probe: {bad} evaluations differ from the clean run")

# ---- stress: many poisoned evaluations of the tensor-core path (a consumer that races ahead of its producer reads the
# poison instead of last evaluation's identical values, which is what hides such a race in repeat-evaluation tests)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
spec = ModelSpec()
calc = AIMNet2Calculator((random_state_dict(0, spec), spec), device="cuda:0")
calc.engine.set_small_m_rows(0)
for nmol, nat, seed in ((64, 50, 99), (100, 37, 5), (13, 50, 8)):
    coord, numbers = random_molecules(nmol, nat, seed=seed)
    inp = {"coord": torch.as_tensor(coord, device="cuda:0"), "numbers": torch.as_tensor(numbers, device="cuda:0"),
           "charge": torch.zeros(nmol, device="cuda:0")}
    calc.engine.debug_poison(-1)
    ref = {k: v.clone() for k, v in calc(dict(inp), forces=True).items() if torch.is_tensor(v)}
    nbad = 0
    for r in range(reps):
        calc.engine.debug_poison(0xFF if r % 2 == 0 else 0x00)
        out = calc(dict(inp), forces=True)
        if not all(torch.equal(ref[k], out[k]) for k in ref):
            nbad += 1
            if nbad <= 5:
                print(f"stress {nmol}x{nat} rep {r}: " + " ".join(
                    f"{k}: max diff {float((out[k].double() - ref[k].double()).abs().nan_to_num(nan=1e30).max()):.3e} nan={bool(out[k].isnan().any())}"
                    for k in ref), flush=True)
    print(f"stress {nmol}x{nat}: {reps} poisoned evaluations, {nbad} differ", flush=True)
