"""Full-size golden outputs: the UNMODIFIED reference (CPU, deterministic=True) on the BASELINE.json configs at their own
sizes.  Build container only:   python -m oracle.make_golden_full [cfg2 cfg3 cfg4]

Inputs are regenerated from seeds by aimnetcentral_b200.structures.benchmark_workload, so a fixture holds only the
reference outputs, the input checksum and the weight recipe:
  full_cfg2.npz  1024 x 50 atoms, aimnet2, Coulomb simple + DFT-D3, charges {0,+1,-1}; reference run in 16 chunks of 64
                 molecules (molecules are independent; the mode-1 all-pairs list needs an N_total^2 scratch otherwise)
                 (+ the float64 twin of the same modules, `*64` arrays: the arbiter where |F| is large)
  full_cfg4.npz  512 x 80 atoms, aimnet2-nse (2 charge channels), charge / mult per molecule; 16 chunks of 32
  full_cfg3.npz  10 080-atom allose supercell, PBC, NN + DSF Coulomb (no D3: the reference's torch D3 path materialises
                 (N, M, 5, 5) temporaries, ~2 GB each at this size), E + F + stress in ONE reference call
  full_d3_1152.npz  3x2x2 allose supercell (1 152 atoms), DSF + DFT-D3, E + F + stress
"""
from __future__ import annotations

import os
import sys
import time
import warnings

import numpy as np
import torch

from aimnetcentral_b200.model_spec import ModelSpec, random_state_dict
from aimnetcentral_b200.structures import allose_supercell, benchmark_workload
from oracle import ref_harness as rh
from oracle.make_golden import GOLD, weights_checksum


def checksum(*arrays) -> float:
    return float(sum(np.abs(np.asarray(a, np.float64)).sum() for a in arrays))


def reference_wrapped(coord, cell):
    """The coordinates the reference actually evaluates for a periodic input: its own move_coord_to_cell
    (aimnet/calculators/neighbors.py:331-381: ((x @ inv(cell)) mod 1) @ cell in fp32), which moves even in-cell atoms by
    up to an ulp of the cell length.  Stored with the periodic fixtures so that the parity test compares both
    implementations at the SAME coordinates (the engine's wrap leaves in-cell atoms bit-for-bit unchanged)."""
    rh._bootstrap()
    from aimnet.calculators.neighbors import move_coord_to_cell

    return move_coord_to_cell(torch.as_tensor(coord, dtype=torch.float32), torch.as_tensor(cell, dtype=torch.float32)).numpy()


def chunked(calc, w, n_chunks, extra=(), calc64=None):
    B = len(w["charge"])
    per = B // n_chunks
    outs = []
    mi = w["mol_idx"]
    for c in range(n_chunks):
        lo, hi = c * per, (c + 1) * per if c < n_chunks - 1 else B
        a0, a1 = np.searchsorted(mi, lo), np.searchsorted(mi, hi)
        inp = dict(coord=w["coord"][a0:a1], numbers=w["numbers"][a0:a1], charge=w["charge"][lo:hi], mol_idx=(mi[a0:a1] - lo).astype(np.int64))
        for k in extra:
            inp[k] = w[k][lo:hi]
        o = rh.run_reference(calc, inp, forces=True)
        if calc64 is not None:   # the float64 twin of the same reference modules: the arbiter of SURVEY.md Appendix C.2
            # per-atom arrays of the twin are stored rounded once to fp32 (6e-8 relative: far below what they arbitrate)
            o.update({k + "64": (v if k == "energy" else v.astype(np.float32))
                      for k, v in rh.run_reference(calc64, inp, forces=True).items()})
        outs.append(o)
        print(f"  chunk {c + 1}/{n_chunks}", flush=True)
    return {k: np.concatenate([o[k] for o in outs]) for k in outs[0]}


def save(name, seed, spec, sd, w, out, **extra):
    meta = dict(weights_seed=seed, weights_scale=0.5, num_charge_channels=spec.C, weights_checksum=weights_checksum(sd),
                input_checksum=checksum(w["coord"], w["numbers"], w["charge"]))
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **{f"ref_{k}": v for k, v in out.items()}, **meta, **extra)
    print(name, {k: v.shape for k, v in out.items()}, os.path.getsize(path), "B", flush=True)


def main(which):
    torch.set_num_threads(os.cpu_count() or 8)
    spec = ModelSpec()
    sd = random_state_dict(0, spec)
    if "cfg2" in which:
        w = benchmark_workload("cfg2", 1234)
        w["charge"] = w["charge"].copy()
        w["charge"][::7] = 1.0
        w["charge"][3::11] = -1.0
        calc = rh.build_reference_calculator(sd, spec)
        t = time.time()
        out = chunked(calc, w, 16, calc64=rh.build_reference_calculator(sd, spec, double=True))
        save("full_cfg2", 0, spec, sd, w, out, charge=w["charge"])
        print("cfg2", time.time() - t, "s")
    if "cfg4" in which:
        spec2 = ModelSpec(num_charge_channels=2)
        sd2 = random_state_dict(1, spec2)
        w = benchmark_workload("cfg4", 1234)
        calc = rh.build_reference_calculator(sd2, spec2)
        t = time.time()
        out = chunked(calc, w, 16, extra=("mult",), calc64=rh.build_reference_calculator(sd2, spec2, double=True))
        save("full_cfg4", 1, spec2, sd2, w, out)
        print("cfg4", time.time() - t, "s")
    if "d3" in which:
        z, x, cell = allose_supercell((3, 2, 2), jitter=0.02, seed=11)
        w = dict(coord=x, numbers=z, charge=np.zeros(1, np.float32))
        calc = rh.build_reference_calculator(sd, spec)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            calc.set_lrcoulomb_method("dsf")
        t = time.time()
        out = rh.run_reference(calc, dict(w, cell=cell), forces=True, stress=True)
        save("full_d3_1152", 0, spec, sd, w, out, coord_wrapped=reference_wrapped(x, cell))
        print("d3_1152", time.time() - t, "s")
    if "cfg3" in which:
        w = benchmark_workload("cfg3", 1234)
        calc = rh.build_reference_calculator(sd, spec)
        calc.external_dftd3 = None
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            calc.set_lrcoulomb_method("dsf")
        t = time.time()
        out = rh.run_reference(calc, dict(coord=w["coord"], numbers=w["numbers"], charge=w["charge"], cell=w["cell"]), forces=True, stress=True)
        save("full_cfg3", 0, spec, sd, w, out, coord_wrapped=reference_wrapped(w["coord"], w["cell"]))
        print("cfg3", time.time() - t, "s")


if __name__ == "__main__":
    main(sys.argv[1:] or ["cfg2", "cfg4", "d3", "cfg3"])
