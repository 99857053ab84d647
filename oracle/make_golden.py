"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.
Run in the build container only:   python -m oracle.make_golden

Each fixture stores the inputs (so the GPU box needs no access to the reference tree), the reference's outputs and the
(seed, scale, channels) needed to regenerate the weights with aimnetcentral_b200.model_spec.random_state_dict, plus
a checksum of those weights.  Fixtures are small (<= a few hundred atoms).
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

from aimnetcentral_b200.model_spec import ModelSpec, random_state_dict
from aimnetcentral_b200.structures import (allose_supercell, random_molecules, random_periodic_box, read_xyz_frame)
from oracle import ref_harness as rh

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF = rh.REF_ROOT


def weights_checksum(sd) -> float:
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float(torch.nan_to_num(v, nan=0.0).abs().sum())
    return tot


def run(calc, data, forces=True, stress=False):
    out = rh.run_reference(calc, data, forces=forces, stress=stress)
    return {f"ref_{k}": v for k, v in out.items()}


def save(name, seed, spec, sd, inputs, outputs, **extra):
    path = os.path.join(GOLD, name + ".npz")
    meta = dict(weights_seed=seed, weights_scale=0.5, num_charge_channels=spec.C, weights_checksum=weights_checksum(sd))
    np.savez_compressed(path, **{f"in_{k}": v for k, v in inputs.items()}, **outputs, **meta, **extra)
    print(f"{name}: " + ", ".join(f"{k}{tuple(np.shape(v))}" for k, v in outputs.items()), os.path.getsize(path), "B")


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    spec = ModelSpec()
    sd = random_state_dict(0, spec)
    full = rh.build_reference_calculator(sd, spec)
    # component calculators: NN(-SR) only, and NN + Coulomb without D3
    spec_nolr = ModelSpec()
    nn_only = rh.build_reference_calculator(sd, spec)
    nn_only.external_coulomb = None
    nn_only.external_dftd3 = None
    nod3 = rh.build_reference_calculator(sd, spec)
    nod3.external_dftd3 = None

    # ---- cfg-1: taxol, 113 atoms, charge 0 and +1, Coulomb simple + D3 --------------------------
    z, xyz = read_xyz_frame(os.path.join(REF, "examples", "taxol.xyz"))
    for q in (0.0, 1.0):
        inp = {"coord": xyz, "numbers": z, "charge": np.array([q], np.float32)}
        out = run(full, inp)
        out.update({k + "_nn": v for k, v in run(nn_only, inp).items()})
        out.update({k + "_nod3": v for k, v in run(nod3, inp).items()})
        save(f"taxol_q{int(q)}", 0, spec, sd, inp, out)

    # ---- caffeine (tests/data/caffeine.xyz geometry; real-weight golden unavailable offline) ---------------------
    with open(os.path.join(REF, "tests", "data", "caffeine.xyz")) as fh:
        lines = fh.read().splitlines()
    n = int(lines[0])
    from aimnetcentral_b200.structures import SYMBOLS

    zc = np.array([SYMBOLS[ln.split()[0]] for ln in lines[2:2 + n]], np.int32)
    xc = np.array([[float(t) for t in ln.split()[1:4]] for ln in lines[2:2 + n]], np.float32)
    inp = {"coord": xc, "numbers": zc, "charge": np.array([0.0], np.float32)}
    save("caffeine", 0, spec, sd, inp, run(full, inp))

    # ---- cfg-2 shaped: 8 x 50-atom molecules, flat + mol_idx, per-molecule charges ------------------------------
    coord, numbers = random_molecules(8, 50, seed=1234)
    charge = np.array([0, 0, 1, -1, 0, 0, 2, 0], np.float32)
    inp = {"coord": coord.reshape(-1, 3), "numbers": numbers.reshape(-1), "charge": charge,
           "mol_idx": np.repeat(np.arange(8), 50).astype(np.int64)}
    out = run(full, inp)
    out.update({k + "_nn": v for k, v in run(nn_only, inp).items()})
    save("mols_8x50", 0, spec, sd, inp, out)
    # ragged batch: molecules of different sizes
    sizes = [3, 17, 50, 1, 29]
    cs, zs, mi = [], [], []
    for k, s in enumerate(sizes):
        c, zz = random_molecules(1, s, seed=100 + k)
        cs.append(c[0]); zs.append(zz[0]); mi.append(np.full(s, k))
    inp = {"coord": np.concatenate(cs), "numbers": np.concatenate(zs), "charge": np.zeros(len(sizes), np.float32),
           "mol_idx": np.concatenate(mi).astype(np.int64)}
    save("mols_ragged", 0, spec, sd, inp, run(full, inp))

    # ---- PBC: 60-atom triclinic cell (multi-image lists), DSF + D3, E+F+stress ----------------------------------
    zb, xb, cell = random_periodic_box(60, seed=7)
    inp = {"coord": xb, "numbers": zb, "charge": np.array([0.0], np.float32), "cell": cell}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        full.set_lrcoulomb_method("dsf")
        nod3.set_lrcoulomb_method("dsf")
    out = run(full, inp, stress=True)
    out.update({k + "_nn": v for k, v in run(nn_only, inp, stress=True).items()})
    out.update({k + "_nod3": v for k, v in run(nod3, inp, stress=True).items()})
    save("pbc_box60_dsf", 0, spec, sd, inp, out)
    # partially periodic slab (pbc = T,T,F)
    inp2 = dict(inp, pbc=np.array([True, True, False]))
    save("pbc_slab60_dsf", 0, spec, sd, inp2, run(full, inp2, stress=False))

    # ---- PBC: allose crystal, 1x1x1 (96 atoms) and 2x1x1 (192), DSF + D3, stress (cfg-3 shaped) -----------------
    for reps in ((1, 1, 1), (2, 1, 1)):
        za, xa, ca = allose_supercell(reps, jitter=0.02, seed=3)
        inp = {"coord": xa, "numbers": za, "charge": np.array([0.0], np.float32), "cell": ca}
        out = run(full, inp, stress=True)
        save("allose_%dx%dx%d_dsf" % reps, 0, spec, sd, inp, out)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        full.set_lrcoulomb_method("simple")

    # ---- NSE (cfg-4 shaped): 2 charge channels, 4 x 20 atoms, charge / mult ------------------------------------
    spec2 = ModelSpec(num_charge_channels=2)
    sd2 = random_state_dict(1, spec2)
    nse = rh.build_reference_calculator(sd2, spec2)
    coord, numbers = random_molecules(4, 20, seed=4321)
    inp = {"coord": coord.reshape(-1, 3), "numbers": numbers.reshape(-1),
           "charge": np.array([0, 1, -1, 0], np.float32), "mult": np.array([1, 2, 2, 3], np.float32),
           "mol_idx": np.repeat(np.arange(4), 20).astype(np.int64)}
    save("nse_4x20", 1, spec2, sd2, inp, run(nse, inp))

    # ---- op-level: conv_sv einsum reference (tests/test_conv_sv_2d_sp.py:73-101) on random data ----------------
    sys.path.insert(0, os.path.join(REF, "tests"))
    g = torch.Generator().manual_seed(5)
    B, A, G, M = 40, 16, 16, 12
    a = torch.randn(B, A, G, generator=g)
    idx = torch.full((B, M), B - 1, dtype=torch.int64)
    for b in range(B - 1):
        k = int(torch.randint(0, M + 1, (1,), generator=g))
        idx[b, :k] = torch.randint(0, B - 1, (k,), generator=g)
    gg = torch.randn(B, M, G, 4, generator=g)
    valid = (idx < B - 1)[..., None, None]
    sel = a.index_select(0, idx.clamp(0, B - 1).flatten()).unflatten(0, (B, M))
    out = torch.einsum("bmag,bmgd->bagd", sel, gg * valid)
    out[-1] = 0
    go = torch.randn(B, A, G, 4, generator=g)
    a_ = a.clone().requires_grad_(True)
    g_ = gg.clone().requires_grad_(True)
    sel = a_.index_select(0, idx.clamp(0, B - 1).flatten()).unflatten(0, (B, M))
    o2 = torch.einsum("bmag,bmgd->bagd", sel, g_ * valid)
    go[-1] = 0
    ga, g_g = torch.autograd.grad(o2, [a_, g_], go)
    np.savez_compressed(os.path.join(GOLD, "conv_sv_op.npz"), a=a.numpy(), idx=idx.numpy().astype(np.int32),
                        g=gg.numpy(), out=out.numpy(), grad_out=go.numpy(), grad_a=ga.numpy(), grad_g=g_g.numpy())
    print("conv_sv_op written")


if __name__ == "__main__":
    main()
