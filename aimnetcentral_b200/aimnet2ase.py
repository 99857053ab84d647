"""ASE front end for the engine: an `ase.calculators.calculator.Calculator` whose `calculate()` is one call of
`AIMNet2Calculator` (the surface of aimnet/calculators/aimnet2ase.py:35-274: constructor arguments, `set_atoms`,
`set_charge`, `set_mult`, `update_tensors`, `get_dipole_moment`, `get_spin_charges`, `get_hessian`, the result keys).

How a structure is handed over:
  * isolated system  -> a batch of one, `coord (1, N, 3)`, `numbers (1, N)`, `charge (1,)`, `mult (1,)`;
  * periodic system  -> flat `coord (N, 3)` plus `cell (3, 3)` and `pbc (3,)`.
Total charge and multiplicity are taken from `atoms.info` ("charge", "mult" or "spin") when present, otherwise from the
values given to the constructor / setters.  The small per-system inputs (species, charge, multiplicity) live on the
device between steps and are re-uploaded only when their host values change, so an MD or optimizer loop pays for the
coordinates only.  Second derivatives are outside the engine's hot path and raise.
"""
from __future__ import annotations

from typing import ClassVar

import numpy as np
import torch

try:
    from ase.calculators.calculator import Calculator, PropertyNotImplementedError, all_changes  # type: ignore
    _ASE_MISSING: ImportError | None = None
except ImportError as _exc:   # importable without ASE (like the reference); constructing the adapter then fails
    _ASE_MISSING = _exc
    all_changes = []  # type: ignore[assignment]

    class PropertyNotImplementedError(RuntimeError):  # type: ignore[no-redef]
        """Stand-in raised for properties the model cannot provide."""

    class Calculator:  # type: ignore[no-redef]
        """Just enough of ASE's base class for this module to import."""

        def __init__(self, *_, **__):
            self.results = {}

        def reset(self):
            self.results = {}

        def check_state(self, *_, **__):
            return []

        def calculate(self, *_, **__):
            return None

from .calculator import AIMNet2Calculator

_BASE_PROPERTIES = ("energy", "forces", "free_energy", "charges", "stress", "dipole_moment")


def _info_of(atoms) -> dict:
    return getattr(atoms, "info", None) or {}


def _multiplicity_in(info: dict):
    return info.get("mult", info.get("spin"))


def _lattice(atoms):
    """(cell (3,3) float array, pbc (3,) bool array) of a periodic structure, or (None, None)."""
    cell = getattr(atoms, "cell", None)
    if cell is None or not np.asarray(atoms.pbc).any():
        return None, None
    return np.asarray(getattr(cell, "array", cell)), np.asarray(atoms.pbc)


class _DeviceScalars:
    """Species / charge / multiplicity tensors kept on the device; each is re-uploaded only when the host value it was
    made from has changed (so ASE's per-step `reset()` costs nothing here)."""

    def __init__(self, device):
        self.device = device
        self.numbers = self.charge = self.mult = None
        self._host = {}

    def species(self, numbers) -> torch.Tensor:
        host = np.asarray(numbers)
        last = self._host.get("numbers")
        if last is None or host.shape != last.shape or (host != last).any():
            self._host["numbers"] = host.copy()
            self.numbers = torch.as_tensor(host, dtype=torch.int32, device=self.device)
        return self.numbers

    def scalar(self, which: str, value) -> torch.Tensor:
        if which not in self._host or self._host[which] != value:
            self._host[which] = value
            setattr(self, which, torch.tensor(value, dtype=torch.float32, device=self.device))
        return getattr(self, which)


class AIMNet2ASE(Calculator):
    implemented_properties: ClassVar[list[str]] = list(_BASE_PROPERTIES)

    def __init__(self, base_calc: AIMNet2Calculator | str = "aimnet2", charge=0, mult=1, validate_species: bool = True):
        if _ASE_MISSING is not None:
            raise ImportError("AIMNet2ASE requires ASE.") from _ASE_MISSING
        super().__init__()
        self.base_calc = AIMNet2Calculator(base_calc) if isinstance(base_calc, str) else base_calc
        self.validate_species = validate_species
        self.charge, self.mult = charge, mult
        if self.base_calc.is_nse:   # open-shell models also report per-atom spin populations
            self.__dict__["implemented_properties"] = [*_BASE_PROPERTIES, "spin_charges"]
        known = (self.base_calc.metadata or {}).get("implemented_species")
        self.implemented_species = np.array(known, dtype=np.int64) if known else None
        self._dev = _DeviceScalars(self.base_calc.device)
        self.update_tensors()

    # ---- state ------------------------------------------------------------------------------------------------
    def set_atoms(self, atoms):
        allowed = self.implemented_species
        if allowed is not None and not np.isin(atoms.numbers, allowed).all():
            raise ValueError("Some species are not implemented in the AIMNet2Calculator")
        self.reset()
        self.atoms = atoms

    def check_state(self, atoms, tol=1e-15):
        """ASE's comparison does not look at `atoms.info`; a changed charge (or multiplicity for an open-shell model)
        there must invalidate cached results too."""
        changed = super().check_state(atoms, tol=tol)
        mine = getattr(self, "atoms", None)
        if changed or mine is None:
            return changed
        before, now = _info_of(mine), _info_of(atoms)
        differs = before.get("charge") != now.get("charge")
        if not differs and self.base_calc.is_nse:
            differs = _multiplicity_in(before) != _multiplicity_in(now)
        return ["info"] if differs else changed

    def set_charge(self, charge):
        self.charge = charge
        self.update_tensors()

    def set_mult(self, mult):
        self.mult = mult
        self.update_tensors()

    def _adopt_info(self, atoms):
        """`atoms.info` wins over the values set on the calculator."""
        info = _info_of(atoms)
        q = info.get("charge")
        if q is not None:
            self.charge = q
        m = _multiplicity_in(info) if self.base_calc.is_nse else None
        if m is not None:
            self.mult = m

    def update_tensors(self, atoms=None):
        """Bring the device-resident species / charge / multiplicity up to date with the host values."""
        atoms = atoms if atoms is not None else getattr(self, "atoms", None)
        if atoms is not None:
            self._dev.species(atoms.numbers)
        self._dev.scalar("charge", self.charge)
        self._dev.scalar("mult", self.mult)

    # ---- derived properties -----------------------------------------------------------------------------------
    def get_dipole_moment(self, atoms):
        return (self.get_charges()[:, None] * atoms.get_positions()).sum(axis=0)

    def get_spin_charges(self, atoms=None):
        if "spin_charges" not in self.results:
            raise PropertyNotImplementedError("spin_charges is not available. Use an NSE model.")
        return self.results["spin_charges"]

    def get_hessian(self, atoms=None):
        """Cartesian Hessian as a (3N, 3N) ndarray in eV/A^2, usable as Sella's `hessian_function` (aimnet2ase.py:163-226)."""
        if atoms is None:
            atoms = getattr(self, "atoms", None)
            if atoms is None:
                raise PropertyNotImplementedError("get_hessian() requires an attached Atoms object or an explicit argument.")
        if np.asarray(atoms.pbc).any():
            raise PropertyNotImplementedError(
                "Hessian for periodic systems is not supported by AIMNet2ASE.get_hessian(). "
                "For periodic transition states, use pysisyphus dimer or climbing-image NEB.")
        self._adopt_info(atoms)
        self.update_tensors(atoms)
        device = self.base_calc.device
        system = {"coord": torch.tensor(np.asarray(atoms.positions), dtype=torch.float32, device=device),
                  "numbers": self._dev.numbers, "charge": self._dev.charge, "mult": self._dev.mult}
        H = self.base_calc(system, forces=True, hessian=True, validate_species=self.validate_species)["hessian"].detach()
        n = H.shape[0]
        return H.reshape(3 * n, 3 * n).cpu().numpy()

    # ---- the step ---------------------------------------------------------------------------------------------
    def _inputs_for(self, atoms) -> tuple[dict, bool]:
        """The calculator's input dict for `atoms` and whether it was wrapped into a batch of one."""
        device = self.base_calc.device
        system = {"coord": torch.tensor(np.asarray(atoms.positions), dtype=torch.float32, device=device),
                  "numbers": self._dev.numbers, "charge": self._dev.charge, "mult": self._dev.mult}
        cell, pbc = _lattice(atoms)
        if cell is None:
            return {key: t.unsqueeze(0) for key, t in system.items()}, True
        system["cell"], system["pbc"] = cell, pbc
        return system, False

    def calculate(self, atoms=None, properties=None, system_changes=all_changes):
        wanted = ["energy"] if properties is None else properties
        super().calculate(atoms, wanted, system_changes)
        self._adopt_info(self.atoms)
        self.update_tensors()
        system, batched = self._inputs_for(self.atoms)
        raw = self.base_calc(system, forces="forces" in wanted, stress="stress" in wanted,
                             validate_species=self.validate_species)
        host = {key: (t.squeeze(0) if batched and key != "energy" else t).detach().cpu().numpy() for key, t in raw.items()}
        energy = host["energy"].item()
        self.results.update(energy=energy, free_energy=energy, charges=host["charges"],
                            dipole_moment=(host["charges"][:, None] * np.asarray(self.atoms.positions)).sum(axis=0))
        for key in ("forces", "stress"):
            if key in wanted:
                self.results[key] = host[key]
        if "spin_charges" in host:
            self.results["spin_charges"] = host["spin_charges"]
