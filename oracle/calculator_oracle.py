"""CPU oracle of the calculator-level flow (TEST INFRASTRUCTURE ONLY): wrap -> neighbor matrices -> padding row ->
model + external Coulomb / DFT-D3 -> autograd forces/stress.  Mirrors AIMNet2Calculator.prepare_input / make_nbmat /
eval of the reference (aimnet/calculators/calculator.py:879-947, 1036-1090, 1521-1702) for flat (mode-1) inputs.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from .aimnet2_oracle import D3Tables, OracleModel, evaluate
from .nblist_oracle import neighbor_matrix, wrap_positions

_D3 = None


def d3_tables() -> D3Tables:
    global _D3
    if _D3 is None:
        here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        _D3 = D3Tables.load(os.path.join(here, "aimnetcentral_b200", "data", "dftd3_tables.npz"))
    return _D3


def _pad_rows(nbmat, shifts, N):
    """aimnet/calculators/neighbors.py:253-284."""
    nbmat = np.concatenate([nbmat, np.full((1, nbmat.shape[1]), N, np.int32)])
    if shifts is not None:
        shifts = np.concatenate([shifts, np.zeros((1, shifts.shape[1], 3), np.int32)])
    return nbmat, shifts


def oracle_calculate(state_dict, inputs: dict, *, num_charge_channels=1, dtype=torch.float32, coulomb="auto",
                     dsf_alpha=0.2, dsf_rc=15.0, dispersion=True, d3_params=None, d3_cutoff=15.0, forces=True,
                     stress=False, cutoff=5.0, timings: dict | None = None, ewald_accuracy=1e-6):
    """inputs: coord (N,3) | (B,n,3), numbers, charge, [mol_idx], [mult], [cell (3,3)], [pbc (3,)].
    coulomb: "auto" -> "simple" without cell, "dsf" with cell (calculator.py:1044-1062); None disables."""
    coord = np.asarray(inputs["coord"], np.float32)
    numbers = np.asarray(inputs["numbers"])
    batch_shape = None
    if coord.ndim == 3:  # mol_flatten (calculator.py:1475-1511), CPU => always flattened
        B, n = coord.shape[:2]
        batch_shape = (B, n)
        mol_idx = np.repeat(np.arange(B), n)
        coord = coord.reshape(-1, 3)
        numbers = numbers.reshape(-1)
        # the reference keeps Z=0 padded atoms of a dense batch as (padding) atoms; we do not support that here
        assert (numbers > 0).all(), "dense batches with Z=0 padding are not supported by the oracle wrapper"
    else:
        mol_idx = inputs.get("mol_idx")
        mol_idx = np.zeros(len(coord), np.int64) if mol_idx is None else np.asarray(mol_idx)
    charge = np.atleast_1d(np.asarray(inputs["charge"], np.float32))
    cell = inputs.get("cell")
    pbc = inputs.get("pbc")
    N = coord.shape[0]
    t0 = time.perf_counter()
    if cell is not None:
        cell = np.asarray(cell, np.float32)
        coord = wrap_positions(coord, cell, pbc)
    if coulomb == "auto":
        coulomb = "dsf" if cell is not None else "simple"
    nb, _, sh = neighbor_matrix(coord, cutoff, cell=cell, pbc=pbc, batch_idx=mol_idx, fill_value=N)
    nb, sh = _pad_rows(nb, sh, N)
    nbl = shl = None
    if coulomb in ("simple", "dsf") or dispersion:
        lr_cut = 1e6 if coulomb == "simple" else max(dsf_rc if coulomb == "dsf" else 0.0, d3_cutoff if dispersion else 0.0)
        nbl, _, shl = neighbor_matrix(coord, lr_cut, cell=cell, pbc=pbc, batch_idx=mol_idx, fill_value=N)
        nbl, shl = _pad_rows(nbl, shl, N)
    t1 = time.perf_counter()
    model = OracleModel(state_dict, num_charge_channels, dtype)
    res = evaluate(model, coord, numbers, charge, mol_idx=mol_idx, mult=inputs.get("mult"), cell=cell, nbmat=nb,
                   shifts=sh, nbmat_lr=nbl, shifts_lr=shl, coulomb=coulomb, dsf_alpha=dsf_alpha, dsf_rc=dsf_rc,
                   ewald_accuracy=ewald_accuracy,
                   d3=d3_tables() if dispersion else None,
                   d3_params=d3_params or {"s8": 0.3908, "a1": 0.5660, "a2": 3.1280, "s6": 1.0}, d3_cutoff=d3_cutoff,
                   forces=forces, stress=stress)
    t2 = time.perf_counter()
    if timings is not None:
        timings["neighbors_s"] = t1 - t0
        timings["model_s"] = t2 - t1
    res["nbmat"], res["shifts"], res["nbmat_lr"], res["shifts_lr"], res["coord_wrapped"] = nb, sh, nbl, shl, coord
    if batch_shape is not None:
        for k in ("forces", "charges", "spin_charges"):
            if k in res:
                res[k] = res[k].reshape(*batch_shape, *res[k].shape[1:])
    return res
