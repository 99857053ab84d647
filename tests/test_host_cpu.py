"""CPU-side checks: the C-ABI library builds/loads and exports every symbol the header declares; host logic that
needs no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from aimnetcentral_b200 import _capi, build

    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "aimnet2_b200.h")).read()
    declared = set(re.findall(r"\b(aimnet2_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_capi.EXPORTS) == declared
    lib.aimnet2_abi_version.restype = ctypes.c_int
    assert lib.aimnet2_abi_version() == 1


def test_calculator_refuses_cpu():
    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict

    spec = ModelSpec()
    with pytest.raises(RuntimeError):
        AIMNet2Calculator((random_state_dict(0, spec), spec), device="cpu")


def test_model_spec_shapes_match_reference_layout():
    """state_dict key names / shapes of SURVEY.md §8b B2."""
    from aimnetcentral_b200 import ModelSpec, random_state_dict

    sd = random_state_dict(0, ModelSpec())
    assert sd["afv.weight"].shape == (64, 256) and sd["conv_a.agh"].shape == (16, 16, 12)
    assert sd["conv_q.agh"].shape == (1, 16, 12)
    assert sd["mlps.0.0.weight"].shape == (512, 704) and sd["mlps.0.4.weight"].shape == (258, 380)
    assert sd["mlps.1.0.weight"].shape == (512, 733) and sd["mlps.2.6.weight"].shape == (256, 380)
    assert str(sd["outputs.atomic_shift.shifts.weight"].dtype) == "torch.float64"
    sd2 = random_state_dict(0, ModelSpec(num_charge_channels=2))
    assert sd2["mlps.1.0.weight"].shape == (512, 762) and sd2["mlps.0.4.weight"].shape == (260, 380)


def test_structures():
    from aimnetcentral_b200.structures import allose_supercell, allose_unit_cell, random_molecules

    z, frac, cell = allose_unit_cell()
    assert len(z) == 96 and abs(abs(np.linalg.det(cell)) - 739.36) < 0.05  # _cell_volume of 2019828.cif
    z, x, big = allose_supercell((7, 3, 5), jitter=0.0)
    assert len(z) == 10080
    c, n = random_molecules(4, 50, seed=1)
    d = np.linalg.norm(c[:, :, None] - c[:, None], axis=-1) + np.eye(50) * 10
    assert d.min() >= 0.9 - 1e-5


def test_nblist_oracle_properties():
    """Oracle neighbor matrix: symmetric (j in row i with shift s <=> i in row j with shift -s), sorted rows."""
    from aimnetcentral_b200.structures import random_periodic_box
    from oracle.nblist_oracle import neighbor_matrix, wrap_positions

    z, x, cell = random_periodic_box(40, seed=3)
    x = wrap_positions(x, cell)
    nb, nnb, sh = neighbor_matrix(x, 6.0, cell=cell)
    N = len(x)
    pairs = set()
    for i in range(N):
        keys = []
        for m in range(nnb[i]):
            pairs.add((i, int(nb[i, m]), *map(int, sh[i, m])))
            keys.append((int(nb[i, m]), *map(int, sh[i, m])))
        assert keys == sorted(keys)
        assert (nb[i, nnb[i]:] == N).all()
    for (i, j, a, b, c) in pairs:
        assert (j, i, -a, -b, -c) in pairs
