"""Hessian fixtures from the UNMODIFIED reference (double backward through its autograd graph, calculator.py:904-947,
derivatives.py:149-192), fp32 and a float64 twin as arbiter.  Build container only:  python -m oracle.make_golden_hessian

tests/golden/hessian_caffeine.npz   24 atoms, Coulomb simple + DFT-D3: E, F, H (N,3,N,3), H @ v for two directions
tests/golden/hessian_mols_3x12.npz  three 12-atom molecules as a (B,N,3) batch: per-structure Hessians
"""
from __future__ import annotations

import os
import warnings

import numpy as np
import torch

from aimnetcentral_b200.model_spec import ModelSpec, random_state_dict
from aimnetcentral_b200.structures import SYMBOLS, random_molecules
from oracle import ref_harness as rh
from oracle.make_golden import GOLD, REF, weights_checksum


def hess(calc, inp):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = calc(dict(inp), forces=True, hessian=True)
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.stack([t.detach().cpu().numpy() for t in v]))
            for k, v in out.items()}


def main():
    torch.set_num_threads(8)
    spec = ModelSpec()
    sd = random_state_dict(0, spec)
    c32 = rh.build_reference_calculator(sd, spec)
    c64 = rh.build_reference_calculator(sd, spec, double=True)
    meta = dict(weights_seed=0, weights_scale=0.5, num_charge_channels=spec.C, weights_checksum=weights_checksum(sd))

    with open(os.path.join(REF, "tests", "data", "caffeine.xyz")) as fh:
        lines = fh.read().splitlines()
    n = int(lines[0])
    z = np.array([SYMBOLS[ln.split()[0]] for ln in lines[2:2 + n]], np.int32)
    x = np.array([[float(t) for t in ln.split()[1:4]] for ln in lines[2:2 + n]], np.float32)
    inp = {"coord": x, "numbers": z, "charge": np.array([0.0], np.float32)}
    o32 = hess(c32, inp)
    o64 = hess(c64, {**inp, "coord": x.astype(np.float64)})
    rng = np.random.default_rng(11)
    v = rng.standard_normal((2, n, 3)).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hv32 = c32.hessian_vector_product(dict(inp), torch.tensor(v)).detach().cpu().numpy()
    H64 = o64["hessian"].reshape(3 * n, 3 * n)
    print("caffeine: |H|max", np.abs(H64).max(), " ref fp32 vs fp64", np.abs(o32["hessian"].reshape(3 * n, 3 * n) - H64).max(),
          " asym fp32", np.abs(o32["hessian"].reshape(3 * n, 3 * n) - o32["hessian"].reshape(3 * n, 3 * n).T).max(),
          " hvp vs H64@v", np.abs(hv32.reshape(2, -1) - v.reshape(2, -1).astype(np.float64) @ H64.T).max())
    np.savez_compressed(os.path.join(GOLD, "hessian_caffeine.npz"), **{f"in_{k}": a for k, a in inp.items()},
                        ref_energy=o32["energy"], ref_forces=o32["forces"], ref_hessian=o32["hessian"],
                        ref64_energy=o64["energy"], ref64_forces=o64["forces"], ref64_hessian=o64["hessian"],
                        in_vectors=v, ref_hvp=hv32, **meta)

    coord, numbers = random_molecules(3, 12, seed=77)
    inp = {"coord": coord, "numbers": numbers, "charge": np.array([0.0, 1.0, -1.0], np.float32)}
    o32 = hess(c32, inp)
    o64 = hess(c64, {**inp, "coord": coord.astype(np.float64)})
    print("mols_3x12:", o32["hessian"].shape, " ref fp32 vs fp64", np.abs(o32["hessian"] - o64["hessian"]).max())
    np.savez_compressed(os.path.join(GOLD, "hessian_mols_3x12.npz"), **{f"in_{k}": a for k, a in inp.items()},
                        ref_energy=o32["energy"], ref_forces=o32["forces"], ref_hessian=o32["hessian"],
                        ref64_hessian=o64["hessian"], **meta)


if __name__ == "__main__":
    main()
