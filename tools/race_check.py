"""Repeat identical evaluations and compare bit for bit: every kernel is atomics-free with fixed reduction orders, so any
difference between repetitions is a race.  python tools/race_check.py [reps]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
from aimnetcentral_b200.structures import allose_supercell, random_molecules

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
spec = ModelSpec()
sd = random_state_dict(0, spec)
calc = AIMNet2Calculator((sd, spec), device="cuda:0")
cases = {}
coord, numbers = random_molecules(512, 50, seed=3)
cases["512x50 molecules"] = ({"coord": coord, "numbers": numbers, "charge": np.zeros(512, np.float32)}, dict(forces=True))
coord, numbers = random_molecules(37, 23, seed=5)
cases["37x23 molecules"] = ({"coord": coord, "numbers": numbers, "charge": np.zeros(37, np.float32)}, dict(forces=True))
z, x, cell = allose_supercell((2, 1, 1), jitter=0.02, seed=1)
cases["allose 2x1x1 pbc dsf+stress"] = ({"coord": x, "numbers": z, "charge": np.zeros(1, np.float32), "cell": cell}, dict(forces=True, stress=True))
bad = 0
for name, (inp, kw) in cases.items():
    if "cell" in inp:
        calc.set_lrcoulomb_method("dsf")
    else:
        calc.set_lrcoulomb_method("simple")
    ref = None
    nbad = 0
    for r in range(reps):
        out = calc(dict(inp), **kw)
        cur = {k: v.clone() for k, v in out.items()}
        if ref is None:
            ref = cur
            continue
        for k in ref:
            if not torch.equal(ref[k], cur[k]):
                d = (ref[k].double() - cur[k].double()).abs().max().item()
                nbad += 1
                if nbad <= 5:
                    print(f"  MISMATCH {name} rep {r} key {k} max|diff|={d:.3e}")
    print(f"{name}: {reps} repetitions, {nbad} mismatching outputs")
    bad += nbad
print("RACE-FREE" if bad == 0 else f"NONDETERMINISTIC: {bad}")
