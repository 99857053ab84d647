"""CPU-side checks: the C-ABI library builds/loads and exports every symbol the header declares; host logic that
needs no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

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
    assert lib.aimnet2_abi_version() == _capi.ABI_VERSION == 2
    # struct layouts of the ctypes binding == the compiled header (checked inside _capi.load(), repeated here)
    sizes = [ctypes.c_int() for _ in range(4)]
    lib.aimnet2_abi_struct_sizes(*[ctypes.byref(x) for x in sizes])
    assert [x.value for x in sizes] == [ctypes.sizeof(t) for t in (_capi.Weights, _capi.Options, _capi.System, _capi.Result)]


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


def test_model_sources_v2_artifact_and_module(tmp_path):
    """B2: the weights ABI — a v2 `.pt` dict (docs/model_format.md:205-222) and an nn.Module carrying `_metadata`
    resolve to the same (state_dict, metadata, channels) triple the engine consumes."""
    import torch
    import yaml

    from aimnetcentral_b200 import ModelSpec, random_state_dict
    from aimnetcentral_b200.calculator import _load_model_source

    spec = ModelSpec(num_charge_channels=2)
    sd = random_state_dict(3, spec)
    artifact = {"format_version": 2, "model_yaml": yaml.safe_dump({"class": "aimnet.models.AIMNet2",
                                                                  "kwargs": {"num_charge_channels": 2}}),
                **{k: v for k, v in spec.metadata().items() if k != "format_version"}, "state_dict": sd}
    path = tmp_path / "model.pt"
    torch.save(artifact, path)
    sd2, meta, C = _load_model_source(str(path))
    assert C == 2 and meta["coulomb_mode"] == "sr_embedded" and meta["d3_params"]["s8"] == 0.3908
    assert set(sd2) == set(sd) and torch.equal(sd2["mlps.1.0.weight"], sd["mlps.1.0.weight"])

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.num_charge_channels = 1
            self.lin = torch.nn.Linear(2, 2)
            self.__dict__["_metadata"] = {"cutoff": 5.0}

    sd3, meta3, C3 = _load_model_source(Tiny())
    assert C3 == 1 and meta3["cutoff"] == 5.0 and "lin.weight" in sd3
    with pytest.raises(FileNotFoundError):
        _load_model_source("aimnet2")  # registry names need a download: outside scope, reported clearly
    with pytest.raises(TypeError):
        _load_model_source(42)


def test_estimate_ewald_parameters_seam_matches_oracle():
    """Host-only operator seam (no GPU work): splitting parameters as the reference reads them from
    estimate_ewald_parameters (aimnet/calculators/calculator.py:1566-1587, formulas :663-666), per system."""
    import math

    import torch

    from aimnetcentral_b200 import ops
    from oracle.aimnet2_oracle import ewald_parameters

    cells = torch.tensor([[[10.0, 0, 0], [1.0, 12.0, 0], [0.5, 0.3, 9.0]], [[20.0, 0, 0], [0, 21.0, 0], [0, 0, 19.0]]])
    batch_idx = torch.tensor([0] * 57 + [1] * 300, dtype=torch.int32)
    p = ops.estimate_ewald_parameters(torch.zeros(357, 3), cells, batch_idx=batch_idx, accuracy=1e-6)
    for s, n in enumerate((57, 300)):
        vol = abs(np.linalg.det(cells[s].numpy().astype(np.float64)))
        eta, rc, kc = ewald_parameters(vol, n, 1e-6)
        assert p.real_space_cutoff[s].item() == pytest.approx(rc, rel=1e-6)
        assert p.reciprocal_space_cutoff[s].item() == pytest.approx(kc, rel=1e-6)
        assert p.alpha[s].item() == pytest.approx(1.0 / (math.sqrt(2.0) * eta), rel=1e-6)
    with pytest.raises(ValueError):
        ops.estimate_ewald_parameters(torch.zeros(3, 3), torch.zeros(3, 3), accuracy=1e-6)   # singular cell


def test_ase_adapter_host_logic(monkeypatch):
    """AIMNet2ASE is pure host marshalling around `calc(dict, forces=, stress=, validate_species=)`: with a stand-in for
    ASE's Calculator base class and a recording stand-in for the calculator (CPU tensors) every branch of it runs here —
    batch-of-one for isolated systems, flat + cell + pbc for periodic ones, atoms.info over constructor values, result
    keys and shapes, residency of the per-system inputs, the error paths (surface of aimnet/calculators/aimnet2ase.py:35-274)."""
    import sys
    import types

    import torch

    class _Base:
        def __init__(self, *a, **k):
            self.results, self.atoms = {}, None

        def reset(self):
            self.results = {}

        def check_state(self, atoms, tol=1e-15):
            return []

        def calculate(self, atoms=None, properties=None, system_changes=None):
            if atoms is not None:
                self.atoms = atoms

        def get_charges(self):
            return self.results["charges"]

    mod = types.ModuleType("ase.calculators.calculator")
    mod.Calculator, mod.PropertyNotImplementedError, mod.all_changes = _Base, RuntimeError, ["positions"]
    for name, m in (("ase", types.ModuleType("ase")), ("ase.calculators", types.ModuleType("ase.calculators")),
                    ("ase.calculators.calculator", mod)):
        monkeypatch.setitem(sys.modules, name, m)
    monkeypatch.delitem(sys.modules, "aimnetcentral_b200.aimnet2ase", raising=False)
    from aimnetcentral_b200.aimnet2ase import AIMNet2ASE

    class FakeCalc:
        device, metadata = torch.device("cpu"), {"implemented_species": [1, 6, 7, 8]}

        def __init__(self, nse=False):
            self.is_nse, self.calls = nse, []

        def __call__(self, data, forces=False, stress=False, hessian=False, validate_species=True):
            self.calls.append((data, forces, stress, validate_species))
            if hessian:
                n = data["coord"].shape[-2]
                return {"energy": torch.tensor([-7.25], dtype=torch.float64), "forces": torch.ones(n, 3),
                        "hessian": torch.arange(9 * n * n, dtype=torch.float32).reshape(n, 3, n, 3)}
            batched = data["coord"].ndim == 3
            n = data["coord"].shape[-2]
            lead = (1,) if batched else ()
            out = {"energy": torch.tensor([-7.25], dtype=torch.float64), "charges": torch.arange(n, dtype=torch.float32).reshape(*lead, n)}
            if forces:
                out["forces"] = torch.ones(*lead, n, 3)
            if stress:
                out["stress"] = torch.eye(3)
            if self.is_nse:
                out["spin_charges"] = torch.zeros(*lead, n)
            return out

    class Atoms:
        def __init__(self, numbers, positions, cell=None, pbc=False, info=None):
            self.numbers, self.positions = np.asarray(numbers), np.asarray(positions, dtype=np.float64)
            self.cell = None if cell is None else np.asarray(cell, dtype=np.float64)
            self.pbc = np.array([pbc] * 3) if np.isscalar(pbc) else np.asarray(pbc)
            self.info = info or {}

        def get_positions(self):
            return self.positions

    pos = np.arange(12, dtype=np.float64).reshape(4, 3)
    fake = FakeCalc()
    ase_calc = AIMNet2ASE(fake, charge=0, mult=1, validate_species=False)
    assert "spin_charges" not in ase_calc.implemented_properties and list(ase_calc.implemented_species) == [1, 6, 7, 8]
    # isolated system: a batch of one, atoms.info["charge"] wins, numbers/charge/mult are device tensors
    ase_calc.calculate(Atoms([6, 1, 1, 8], pos, info={"charge": 1}), properties=["energy", "forces"])
    data, forces, stress, validate = fake.calls[-1]
    assert data["coord"].shape == (1, 4, 3) and data["coord"].dtype == torch.float32
    assert data["numbers"].shape == (1, 4) and data["numbers"].dtype == torch.int32
    assert data["charge"].shape == (1,) and data["charge"].item() == 1.0 and data["mult"].item() == 1.0
    assert forces and not stress and validate is False and "cell" not in data
    r = ase_calc.results
    assert r["energy"] == -7.25 and r["free_energy"] == -7.25 and r["charges"].shape == (4,) and r["forces"].shape == (4, 3)
    assert np.allclose(r["dipole_moment"], (np.arange(4)[:, None] * pos).sum(axis=0)) and "stress" not in r
    assert np.allclose(ase_calc.get_dipole_moment(ase_calc.atoms), r["dipole_moment"])
    # the species tensor stays resident while the numbers do not change, and follows them when they do
    first = data["numbers"]
    ase_calc.calculate(Atoms([6, 1, 1, 8], pos + 0.1, info={"charge": 1}), properties=["energy"])
    assert fake.calls[-1][0]["numbers"].data_ptr() == first.data_ptr()
    ase_calc.calculate(Atoms([6, 1, 1, 7], pos, info={"charge": 1}), properties=["energy"])
    assert fake.calls[-1][0]["numbers"].flatten().tolist() == [6, 1, 1, 7]
    # periodic system: flat coordinates + cell + pbc, stress through
    cell = np.diag([9.0, 10.0, 11.0])
    ase_calc.calculate(Atoms([6, 1, 1, 8], pos, cell=cell, pbc=True), properties=["energy", "forces", "stress"])
    data, forces, stress, _ = fake.calls[-1]
    assert data["coord"].shape == (4, 3) and np.allclose(data["cell"], cell) and data["pbc"].all() and forces and stress
    assert ase_calc.results["stress"].shape == (3, 3) and ase_calc.results["charges"].shape == (4,)
    # setters, info-driven invalidation, error paths
    ase_calc.set_charge(-1)
    ase_calc.calculate(Atoms([6, 1, 1, 8], pos), properties=["energy"])
    assert fake.calls[-1][0]["charge"].item() == -1.0
    a0, a1 = Atoms([6], pos[:1], info={"charge": 0}), Atoms([6], pos[:1], info={"charge": 1})
    ase_calc.atoms = a0
    assert ase_calc.check_state(a1) == ["info"] and ase_calc.check_state(a0) == []
    with pytest.raises(ValueError):
        ase_calc.set_atoms(Atoms([6, 26], pos[:2]))
    with pytest.raises(RuntimeError):
        ase_calc.get_spin_charges()
    # Hessian: Sella's callback contract (atoms) -> (3N, 3N) ndarray; flat single-structure input; periodic / no atoms raise
    H = ase_calc.get_hessian(Atoms([6, 1, 1, 8], pos))
    assert H.shape == (12, 12) and H[1, 0] == 12.0 and fake.calls[-1][0]["coord"].shape == (4, 3)
    with pytest.raises(RuntimeError):
        ase_calc.get_hessian(Atoms([6, 1, 1, 8], pos, cell=cell, pbc=True))
    ase_calc.atoms = None
    with pytest.raises(RuntimeError):
        ase_calc.get_hessian()
    ase_calc.atoms = a0
    # open-shell model: multiplicity from atoms.info ("mult" or "spin"), spin populations in the results
    nse = AIMNet2ASE(FakeCalc(nse=True), charge=0, mult=1)
    assert "spin_charges" in nse.implemented_properties
    nse.calculate(Atoms([6, 1, 1, 8], pos, info={"spin": 3}), properties=["energy"])
    assert nse.base_calc.calls[-1][0]["mult"].item() == 3.0 and nse.get_spin_charges().shape == (4,)
    assert nse.check_state(Atoms([6, 1, 1, 8], pos, info={"spin": 1})) == ["info"]
    monkeypatch.delitem(sys.modules, "aimnetcentral_b200.aimnet2ase", raising=False)


def test_pair_term_seams_reject_bad_arguments_before_touching_the_gpu():
    """Argument validation of the pair-term entry points happens before any CUDA call, so it can be exercised here:
    status -1 (invalid argument, raised as ValueError by the Python layer) and a message in aimnet2_last_error()."""
    from aimnetcentral_b200 import _capi

    lib = _capi.load()
    one = ctypes.c_void_p(1)   # never dereferenced: the calls below fail validation first
    # no neighbor matrix
    rc = lib.aimnet2_dsf_coulomb(one, one, 4, 9.0, 0.2, None, 0, None, 1, None, None, 0, 4, one, None, None, None, None)
    assert rc == -1 and b"neighbor matrix" in lib.aimnet2_last_error()
    # a cell without shifts
    rc = lib.aimnet2_dsf_coulomb(one, one, 4, 9.0, 0.2, one, 1, None, 1, one, None, 8, 4, one, None, None, None, None)
    assert rc == -1 and b"shifts" in lib.aimnet2_last_error()
    # cell count that is neither 1 nor the number of systems
    rc = lib.aimnet2_dsf_coulomb(one, one, 4, 9.0, 0.2, one, 2, None, 3, one, one, 8, 4, one, None, None, None, None)
    assert rc == -1 and b"n_cells" in lib.aimnet2_last_error()
    # non-positive cutoff
    rc = lib.aimnet2_dsf_coulomb(one, one, 4, 0.0, 0.2, None, 0, None, 1, one, None, 8, 4, one, None, None, None, None)
    assert rc == -1 and b"cutoff" in lib.aimnet2_last_error()
    # D3: switching window upside down, missing tables
    rc = lib.aimnet2_dftd3(one, one, 4, 1.0, 0.39, 0.57, 3.1, 30.0, 20.0, one, one, one, one, None, 0, None, 1, one, None, 8, 4,
                           one, None, None, None, None)
    assert rc == -1 and b"r_on" in lib.aimnet2_last_error()
    rc = lib.aimnet2_dftd3(one, one, 4, 1.0, 0.39, 0.57, 3.1, 20.0, 30.0, None, one, one, one, None, 0, None, 1, one, None, 8, 4,
                           one, None, None, None, None)
    assert rc == -1 and b"null argument" in lib.aimnet2_last_error()
    with pytest.raises(ValueError):
        _capi.check(rc, "dftd3")


# ---- TorchSim / PySisyphus adapters: host logic with stand-ins (neither package is in the image) ----------------------
class _FakeCalc:
    """Duck-typed calculator, the contract of tests/test_torchsim.py:20-50 of the reference."""

    def __init__(self, is_nse=False):
        self.device, self.is_nse = "cpu", is_nse
        self.calls = []

    @property
    def metadata(self):
        return {"family": "fake"}

    def __call__(self, data, forces=False, stress=False, hessian=False, validate_species=True):
        self.calls.append(dict(data=data, forces=forces, stress=stress, validate_species=validate_species))
        n_sys = int(torch.as_tensor(data["mol_idx"]).max().item()) + 1 if "mol_idx" in data else 1
        out = {"energy": torch.arange(n_sys, dtype=torch.float64) + 1.0, "charges": torch.zeros(data["coord"].shape[0])}
        if forces:
            out["forces"] = torch.ones_like(data["coord"])
        if stress:
            out["stress"] = torch.zeros(n_sys, 3, 3)
        if self.is_nse:
            out["spin_charges"] = torch.zeros(data["coord"].shape[0])
        return out


class _State:
    def __init__(self, n_sys=2, n_per=3, periodic=False, **extras):
        self.positions = torch.arange(n_sys * n_per * 3, dtype=torch.float32).reshape(-1, 3)
        self.atomic_numbers = torch.full((n_sys * n_per,), 6)
        self.system_idx = torch.arange(n_sys).repeat_interleave(n_per)
        self.n_systems = n_sys
        self.pbc = torch.tensor([periodic] * 3)
        self.row_vector_cell = (torch.eye(3) * 10.0).repeat(n_sys, 1, 1) if periodic else torch.zeros(n_sys, 3, 3)
        self.device, self.dtype = torch.device("cpu"), torch.float32
        for k, v in extras.items():
            setattr(self, k, v)

    def to(self, device, dtype):
        return self


def test_torchsim_adapter_host_logic(monkeypatch):
    from aimnetcentral_b200 import aimnet2torchsim as m

    with pytest.raises(ImportError):
        m.AIMNet2TorchSim(_FakeCalc())          # TorchSim is not installed here: same failure as the reference
    monkeypatch.setattr(m, "_TORCHSIM_IMPORT_ERROR", None)
    calc = _FakeCalc()
    w = m.AIMNet2TorchSim(calc)
    assert w.implemented_properties == ["energy", "forces", "charges", "partial_charges"]
    st = _State()
    pos0 = st.positions.clone()
    res = w(st)
    d = calc.calls[-1]["data"]
    assert d["coord"].data_ptr() != st.positions.data_ptr() and torch.equal(st.positions, pos0)
    assert torch.equal(d["charge"], torch.zeros(2)) and "cell" not in d and "mult" not in d
    assert calc.calls[-1]["forces"] is True and calc.calls[-1]["stress"] is False
    assert res["partial_charges"].data_ptr() == res["charges"].data_ptr()
    w.compute_stress = True
    assert "stress" in w.implemented_properties
    with pytest.raises(ValueError, match="periodic TorchSim state"):
        w(st)
    res = w(_State(periodic=True))
    d = calc.calls[-1]["data"]
    assert d["cell"].shape == (2, 3, 3) and "pbc" in d and res["stress"].shape == (2, 3, 3)
    # per-system extras: scalar broadcast, one value per system, wrong count
    nse = m.AIMNet2TorchSim(_FakeCalc(is_nse=True), compute_forces=False)
    assert nse.implemented_properties == ["energy", "charges", "partial_charges", "spin_charges"]
    nse(_State(charge=torch.tensor([1.0, -1.0]), spin=2.0))
    d = nse.base_calc.calls[-1]["data"]
    assert d["charge"].tolist() == [1.0, -1.0] and d["mult"].tolist() == [2.0, 2.0]
    with pytest.raises(ValueError, match="one value per system"):
        nse(_State(charge=torch.tensor([1.0, 0.0, 0.0])))


def test_pysis_adapter_host_logic(monkeypatch):
    from aimnetcentral_b200 import aimnet2pysis as m

    with pytest.raises(ImportError):
        m.AIMNet2Pysis(_FakeCalc())
    monkeypatch.setattr(m, "_PYSIS_IMPORT_ERROR", None)
    monkeypatch.setattr(m, "ATOMIC_NUMBERS", {"c": 6, "h": 1})
    monkeypatch.setattr(m, "BOHR2ANG", 0.5)
    monkeypatch.setattr(m, "ANG2BOHR", 2.0)
    monkeypatch.setattr(m, "AU2EV", 4.0)
    calc = _FakeCalc()
    p = m.AIMNet2Pysis(calc, charge=1, mult=2)
    atoms, coords = ("C", "H"), np.arange(6, dtype=np.float64)
    f = p.get_forces(atoms, coords)
    d = calc.calls[-1]["data"]
    assert d["numbers"].tolist() == [6, 1] and d["coord"].dtype == torch.float32
    assert np.allclose(d["coord"].numpy().reshape(-1), coords * 0.5)          # Bohr -> Angstrom
    assert d["charge"].tolist() == [1.0] and d["mult"].tolist() == [2.0]
    assert f["energy"] == pytest.approx(1.0 / 4.0) and np.allclose(f["forces"], 1.0 / 4.0 / 2.0) and f["forces"].dtype == np.float64
    n = len(calc.calls)
    assert p.get_energy(atoms, coords)["energy"] == pytest.approx(0.25) and len(calc.calls) == n   # served from the cache
    p.get_energy(atoms, coords + 1.0)
    assert len(calc.calls) == n + 1 and calc.calls[-1]["forces"] is False
    p.get_forces(atoms, coords + 1.0)                                         # an energy-only entry is not a forces hit
    assert len(calc.calls) == n + 2


def test_hessian_host_logic_on_a_quadratic_surface():
    """The finite-difference Hessian / HVP code of the calculator (displaced molecule batches, chunking, stencil weights
    from the realised fp32 displacements, symmetrisation, batched and padded inputs) on a fake engine whose energy is the
    quadratic form E = x^T K x / 2 per structure: every stencil is exact there, so the Hessian must come back as K."""
    import torch

    from aimnetcentral_b200.calculator import AIMNet2Calculator

    n = 5
    g = torch.Generator().manual_seed(0)
    A = torch.randn(3 * n, 3 * n, generator=g, dtype=torch.float64)
    K = (A + A.T) * 0.5

    class FakeEngine:
        calls = []

        def eval(self, coord, numbers, charge, mol_idx=None, forces=True, **kw):
            B = int(charge.shape[0])
            per = coord.shape[0] // B
            self.calls.append(B)
            out_e, out_f = [], []
            for b in range(B):
                x = coord[b * per:(b + 1) * per].double().reshape(-1)
                m = 3 * per
                out_e.append(0.5 * x @ K[:m, :m] @ x)
                out_f.append((-(K[:m, :m] @ x)).reshape(per, 3).float())
            return {"energy": torch.stack(out_e), "forces": torch.cat(out_f), "charges": torch.zeros(coord.shape[0])}

    calc = AIMNet2Calculator.__new__(AIMNet2Calculator)
    calc.device, calc.engine, calc._coulomb_method = "cpu", FakeEngine(), "simple"
    calc._num_charge_channels, calc._mult_ignored_checked = 1, True
    calc._upload_cache, calc._validated_numbers, calc._host_cell_cache = {}, None, None
    x = torch.randn(n, 3, generator=g).numpy().astype(np.float32) * 2.0
    z = np.array([6, 1, 1, 8, 7])
    data = {"coord": x, "numbers": z, "charge": 0.0}
    for stencil in (2, 4, 6):
        calc.hessian_stencil = stencil
        out = calc.eval(dict(data), hessian=True, validate_species=False)
        H = out["hessian"].double().reshape(3 * n, 3 * n)
        assert out["hessian"].shape == (n, 3, n, 3) and torch.equal(H, H.T)
        assert float((H - K).abs().max()) < 2e-3 * float(K.abs().max()), stencil   # fp32 force rounding / step
    del calc.hessian_stencil
    # stencil weights: classical coefficients on nominal nodes, exact first derivative of polynomials on perturbed ones
    w = calc._fd_weights(torch.tensor([[1.0, -1.0, 2.0, -2.0, 3.0, -3.0]]))[0]
    assert torch.allclose(w, torch.tensor([3 / 4, -3 / 4, -3 / 20, 3 / 20, 1 / 60, -1 / 60], dtype=torch.float64), atol=1e-12)
    nodes = torch.tensor([[1.0003, -0.9998, 2.0001, -1.9996]], dtype=torch.float64)
    w = calc._fd_weights(nodes)[0]
    poly = lambda u: 0.3 + 1.7 * u - 0.4 * u ** 2 + 0.9 * u ** 3   # noqa: E731  (derivative at 0: 1.7)
    assert abs(float((w * poly(nodes[0])).sum()) - 1.7) < 1e-9
    # chunks of the displaced batch (here 2 components per engine call), a stacked batch, a mol_idx batch, a padded structure
    calc.hessian_batch_atoms = 6 * n * 2
    calc.engine.calls.clear()
    H1 = calc.eval(dict(data), hessian=True, validate_species=False)["hessian"]
    assert max(calc.engine.calls) == 12 and float((H1.double().reshape(3 * n, 3 * n) - K).abs().max()) < 2e-3 * float(K.abs().max())
    del calc.hessian_batch_atoms
    xb = np.stack([x, x * 0.5])
    outb = calc.eval({"coord": xb, "numbers": np.stack([z, z]), "charge": np.zeros(2, np.float32)}, hessian=True, validate_species=False)
    assert outb["hessian"].shape == (2, n, 3, n, 3) and outb["energy"].shape == (2, 1)
    assert float((outb["hessian"][1].double().reshape(3 * n, 3 * n) - K).abs().max()) < 2e-3 * float(K.abs().max())
    outl = calc.eval({"coord": xb.reshape(-1, 3), "numbers": np.concatenate([z, z]), "charge": np.zeros(2, np.float32),
                      "mol_idx": np.repeat([0, 1], n)}, hessian=True, validate_species=False)
    assert isinstance(outl["hessian"], list) and len(outl["hessian"]) == 2 and outl["hessian"][0].shape == (n, 3, n, 3)
    pad = {"coord": np.concatenate([x[:4], np.zeros((1, 3), np.float32)])[None], "numbers": np.array([[6, 1, 1, 8, 0]]), "charge": np.zeros(1, np.float32)}
    Hp = calc.eval(pad, hessian=True, validate_species=False)["hessian"]
    assert Hp.shape == (5, 3, 5, 3) and not Hp[4].any() and not Hp[:, :, 4].any()
    assert float((Hp[:4, :, :4].double().reshape(12, 12) - K[:12, :12]).abs().max()) < 2e-3 * float(K.abs().max())
    # Hessian-vector products: one and several directions, shape errors
    v = torch.randn(3, n, 3, generator=g)
    hv = calc.hessian_vector_product(dict(data), v, validate_species=False)
    want = (K @ v.double().reshape(3, -1).T).T.reshape(3, n, 3)
    assert hv.shape == (3, n, 3) and float((hv.double() - want).abs().max()) < 5e-3 * float(want.abs().max())
    assert calc.hessian_vector_product(dict(data), v[0], validate_species=False).shape == (n, 3)
    with pytest.raises(ValueError):
        calc.hessian_vector_product(dict(data), v[:, :4], validate_species=False)
    with pytest.raises(NotImplementedError):
        calc.hessian_vector_product({"coord": xb, "numbers": np.stack([z, z]), "charge": np.zeros(2, np.float32)}, v[0], validate_species=False)
    with pytest.raises(NotImplementedError):
        calc.hessian_vector_product(dict(data), v[0], create_graph=True, validate_species=False)
