"""Print per-phase device times of a few consecutive evaluations (engine timing level 1/2)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from aimnetcentral_b200 import ModelSpec, random_state_dict
from aimnetcentral_b200.engine import Engine
from bench import make_workload

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
dev = torch.device("cuda:0")
spec = ModelSpec(); sd = random_state_dict(0, spec)
eng = Engine(sd, 1, dev)
w = make_workload(wl, 1234)
pbc = w["cell"] is not None
eng.set_options(coulomb_method="dsf" if pbc else "simple", dispersion=True)
c = torch.from_numpy(w["coord"]).to(dev); z = torch.from_numpy(w["numbers"]).to(dev); q = torch.from_numpy(w["charge"]).to(dev)
m = torch.from_numpy(w["mol_idx"]).to(dev) if w["mol_idx"] is not None else None
cell = torch.from_numpy(w["cell"]).to(dev) if pbc else None
for lvl in (1, 2, 0):
    eng.enable_timing(lvl)
    for i in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = eng.eval(c, z, q, mol_idx=m, cell=cell, forces=True, stress=w["stress"])
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        print(f"timing={lvl} wall={dt:.2f} ms", {k: round(v, 3) for k, v in eng.last_timing().items()}, eng.info())
