"""Where does the translation-dependence of the periodic forces come from?  Percentiles of |dF| per atom between a
cfg-3 sized box and its rigid translate, per Coulomb method / Ewald accuracy / dispersion."""
import sys, warnings
import numpy as np
import torch
sys.path.insert(0, ".")
from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
from aimnetcentral_b200.structures import allose_supercell

spec = ModelSpec()
sd = random_state_dict(0, spec)
z, x, cell = allose_supercell((7, 3, 5), jitter=0.02, seed=3)
shift = np.array([3.217, -1.04, 7.9], np.float32)
warnings.simplefilter("ignore")


def run(calc, xx):
    out = calc({"coord": xx, "numbers": z, "charge": np.zeros(1, np.float32), "cell": cell}, forces=True, stress=True)
    return {k: v.cpu().numpy() for k, v in out.items()}


for label, kw, setup in (("dsf+d3", {}, lambda c: c.set_lrcoulomb_method("dsf")),
                         ("ewald1e-6+d3", {}, lambda c: c.set_lrcoulomb_method("ewald")),
                         ("ewald1e-8+d3", {}, lambda c: c.set_lrcoulomb_method("ewald", ewald_accuracy=1e-8)),
                         ("ewald1e-6 no d3", {"needs_dispersion": False}, lambda c: c.set_lrcoulomb_method("ewald")),
                         ("nn only", {"needs_dispersion": False, "needs_coulomb": False}, lambda c: None)):
    calc = AIMNet2Calculator((sd, spec), device="cuda:0", **kw)
    setup(calc)
    a, b = run(calc, x), run(calc, x + shift)
    d = np.abs(a["forces"] - b["forces"]).max(axis=1)
    i = int(d.argmax())
    print(f"{label:18s} dE {abs(a['energy'][0] - b['energy'][0]):.2e}  |dF| p50 {np.percentile(d, 50):.1e} p99 {np.percentile(d, 99):.1e} "
          f"p99.9 {np.percentile(d, 99.9):.1e} max {d.max():.1e} (atom {i}, Z {z[i]}, |F| {np.abs(a['forces'][i]).max():.2f})  n>2e-4: {(d > 2e-4).sum()}",
          flush=True)
    del calc
