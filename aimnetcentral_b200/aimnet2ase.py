"""ASE adapter with the reference's surface (aimnet/calculators/aimnet2ase.py:35-274): pure host marshalling around
`AIMNet2Calculator.__call__`.  Non-periodic Atoms go in as a `(1, N, 3)` batch, periodic ones as flat `(N, 3)` +
`cell` + `pbc`; charge / multiplicity come from `atoms.info` first, then from the calculator.  Hessians are outside
the engine's hot path and raise.
"""
from __future__ import annotations

from typing import ClassVar

import numpy as np
import torch

try:
    from ase.calculators.calculator import Calculator, PropertyNotImplementedError, all_changes  # type: ignore
except ImportError as exc:  # same behaviour as the reference: importable without ASE, unusable until it is installed
    _ASE_IMPORT_ERROR: ImportError | None = exc

    class Calculator:  # type: ignore[no-redef]
        def __init__(self, *args, **kwargs):
            self.results = {}

        def reset(self):
            self.results = {}

        def check_state(self, *args, **kwargs):
            return []

        def calculate(self, *args, **kwargs):
            return None

    class PropertyNotImplementedError(RuntimeError):  # type: ignore[no-redef]
        pass

    all_changes = []  # type: ignore[assignment]
else:
    _ASE_IMPORT_ERROR = None

from .calculator import AIMNet2Calculator


class AIMNet2ASE(Calculator):
    implemented_properties: ClassVar[list[str]] = ["energy", "forces", "free_energy", "charges", "stress", "dipole_moment"]

    def __init__(self, base_calc: AIMNet2Calculator | str = "aimnet2", charge=0, mult=1, validate_species: bool = True):
        if _ASE_IMPORT_ERROR is not None:
            raise ImportError("AIMNet2ASE requires ASE.") from _ASE_IMPORT_ERROR
        super().__init__()
        if isinstance(base_calc, str):
            base_calc = AIMNet2Calculator(base_calc)
        self.base_calc = base_calc
        self.validate_species = validate_species
        if self.base_calc.is_nse:
            self.__dict__["implemented_properties"] = [*self.__class__.implemented_properties, "spin_charges"]
        self.reset()
        self.charge = charge
        self.mult = mult
        self.update_tensors()
        meta = self.base_calc.metadata
        species = meta.get("implemented_species") if meta is not None else None
        self.implemented_species = np.array(species, dtype=np.int64) if species else None

    def reset(self):
        super().reset()
        self._t_numbers = None
        self._t_charge = None
        self._t_mult = None

    def set_atoms(self, atoms):
        if self.implemented_species is not None and not np.isin(atoms.numbers, self.implemented_species).all():
            raise ValueError("Some species are not implemented in the AIMNet2Calculator")
        self.reset()
        self.atoms = atoms

    def check_state(self, atoms, tol=1e-15):
        state = super().check_state(atoms, tol=tol)
        if (not state) and getattr(self, "atoms", None) is not None:
            old, new = getattr(self.atoms, "info", {}), getattr(atoms, "info", {})
            if old.get("charge") != new.get("charge"):
                state.append("info")
            elif self.base_calc.is_nse and old.get("spin", old.get("mult")) != new.get("spin", new.get("mult")):
                state.append("info")
        return state

    def set_charge(self, charge):
        self.charge = charge
        self._t_charge = None
        self.update_tensors()

    def set_mult(self, mult):
        self.mult = mult
        self._t_mult = None
        self.update_tensors()

    def _update_charge_spin_from_info(self, atoms=None):
        atoms = atoms if atoms is not None else getattr(self, "atoms", None)
        if atoms is None:
            return
        info = getattr(atoms, "info", {})
        charge = info.get("charge")
        if charge is not None and charge != self.charge:
            self.charge = charge
            self._t_charge = None
        if self.base_calc.is_nse:
            mult = info.get("mult", info.get("spin"))
            if mult is not None and mult != self.mult:
                self.mult = mult
                self._t_mult = None

    def update_tensors(self, atoms=None):
        atoms = atoms if atoms is not None else getattr(self, "atoms", None)
        dev = self.base_calc.device
        if atoms is not None:
            new = torch.as_tensor(np.asarray(atoms.numbers), dtype=torch.int32, device=dev)
            if self._t_numbers is None or self._t_numbers.shape != new.shape or not torch.equal(self._t_numbers, new):
                self._t_numbers = new
        if self._t_charge is None:
            self._t_charge = torch.tensor(self.charge, dtype=torch.float32, device=dev)
        if self._t_mult is None:
            self._t_mult = torch.tensor(self.mult, dtype=torch.float32, device=dev)

    def get_dipole_moment(self, atoms):
        return np.sum(self.get_charges()[:, np.newaxis] * atoms.get_positions(), axis=0)

    def get_spin_charges(self, atoms=None):
        if "spin_charges" not in self.results:
            raise PropertyNotImplementedError("spin_charges is not available. Use an NSE model.")
        return self.results["spin_charges"]

    def get_hessian(self, atoms=None):
        raise PropertyNotImplementedError("Hessians are outside the B200 engine's hot path (SURVEY.md §8f f4)")

    def calculate(self, atoms=None, properties=None, system_changes=all_changes):
        if properties is None:
            properties = ["energy"]
        super().calculate(atoms, properties, system_changes)
        self._update_charge_spin_from_info()
        self.update_tensors()
        periodic = self.atoms.cell is not None and np.asarray(self.atoms.pbc).any()
        cell = np.asarray(self.atoms.cell.array if hasattr(self.atoms.cell, "array") else self.atoms.cell) if periodic else None
        dev = self.base_calc.device
        _in = {"coord": torch.tensor(np.asarray(self.atoms.positions), dtype=torch.float32, device=dev),
               "numbers": self._t_numbers, "charge": self._t_charge, "mult": self._t_mult}
        unsqueezed = False
        if cell is not None:
            _in["cell"] = cell
            _in["pbc"] = np.asarray(self.atoms.pbc)
        else:
            _in = {k: v.unsqueeze(0) for k, v in _in.items()}
            unsqueezed = True
        results = self.base_calc(_in, forces="forces" in properties, stress="stress" in properties,
                                 validate_species=self.validate_species)
        out = {}
        for k, v in results.items():
            if unsqueezed and k != "energy":
                v = v.squeeze(0)
            out[k] = v.detach().cpu().numpy()
        self.results["energy"] = out["energy"].item()
        self.results["free_energy"] = self.results["energy"]
        self.results["charges"] = out["charges"]
        self.results["dipole_moment"] = np.sum(out["charges"][:, None] * np.asarray(self.atoms.positions), axis=0)
        if "forces" in properties:
            self.results["forces"] = out["forces"]
        if "stress" in properties:
            self.results["stress"] = out["stress"]
        if "spin_charges" in out:
            self.results["spin_charges"] = out["spin_charges"]
