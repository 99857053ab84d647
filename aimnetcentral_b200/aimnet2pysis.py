"""`AIMNet2Pysis` — PySisyphus `Calculator` around the B200 `AIMNet2Calculator`, same surface as the reference's adapter
(aimnet/calculators/aimnet2pysis.py:28-108): Bohr / Hartree on the PySisyphus side, Angstrom / eV on the model side,
`get_energy` / `get_forces` / `get_hessian`, and the one-entry result cache that serves `get_energy` right after
`get_forces` at the same geometry (AFIR / IRC call them back to back).  Hessians are outside this engine's hot path
(SURVEY.md section 8f f4): `get_hessian` raises the calculator's NotImplementedError."""
from __future__ import annotations

import numpy as np
import torch

try:  # PySisyphus is optional (not in this image)
    import pysisyphus.run  # type: ignore
    from pysisyphus.calculators.Calculator import Calculator  # type: ignore
    from pysisyphus.constants import ANG2BOHR, AU2EV, BOHR2ANG  # type: ignore
    from pysisyphus.elem_data import ATOMIC_NUMBERS  # type: ignore
except ImportError as exc:
    _PYSIS_IMPORT_ERROR: ImportError | None = exc

    class Calculator:  # type: ignore[no-redef]
        def __init__(self, *args, charge=0, mult=1, **kwargs):
            self.charge, self.mult = charge, mult

    ANG2BOHR = 1.0
    AU2EV = 1.0
    BOHR2ANG = 1.0
    ATOMIC_NUMBERS: dict[str, int] = {}
    pysisyphus = None  # type: ignore[assignment]
else:
    _PYSIS_IMPORT_ERROR = None


class AIMNet2Pysis(Calculator):
    def __init__(self, model="aimnet2", charge=0, mult=1, validate_species: bool = True, **kwargs):
        if _PYSIS_IMPORT_ERROR is not None:
            raise ImportError("AIMNet2Pysis requires PySisyphus. Install it with `pip install pysisyphus`.") from _PYSIS_IMPORT_ERROR
        super().__init__(charge=charge, mult=mult, **kwargs)
        if isinstance(model, str):
            from .calculator import AIMNet2Calculator

            model = AIMNet2Calculator(model)
        self.model = model
        self.validate_species = validate_species
        self._cache_key = None
        self._cache_results = None

    # unit factors are looked up at call time so that they follow the PySisyphus constants
    @staticmethod
    def _ev2au() -> float:
        return 1.0 / AU2EV

    def _prepare_input(self, atoms, coord):
        dev = self.model.device
        numbers = torch.as_tensor([ATOMIC_NUMBERS[a.lower()] for a in atoms], device=dev)
        xyz = (np.asarray(coord, dtype=np.float32) * BOHR2ANG).reshape(-1, 3)   # cast + scale on the host: half the H2D bytes
        return {"coord": torch.from_numpy(xyz).to(dev), "numbers": numbers,
                "charge": torch.as_tensor([self.charge], dtype=torch.float, device=dev),
                "mult": torch.as_tensor([self.mult], dtype=torch.float, device=dev)}

    def _energy(self, res) -> float:
        return res["energy"].item() * self._ev2au()

    def _forces(self, res) -> np.ndarray:
        return (res["forces"].detach() * (self._ev2au() / ANG2BOHR)).flatten().to(torch.double).cpu().numpy()

    @staticmethod
    def _key(atoms, coords):
        return (tuple(atoms), np.asarray(coords).tobytes())

    def get_energy(self, atoms, coords):
        key = self._key(atoms, coords)
        if self._cache_key == key and self._cache_results is not None:
            return {"energy": self._energy(self._cache_results)}
        res = self.model(self._prepare_input(atoms, coords), validate_species=self.validate_species)
        self._cache_key, self._cache_results = key, res
        return {"energy": self._energy(res)}

    def get_forces(self, atoms, coords):
        key = self._key(atoms, coords)
        if self._cache_key == key and self._cache_results is not None and "forces" in self._cache_results:
            return {"energy": self._energy(self._cache_results), "forces": self._forces(self._cache_results)}
        res = self.model(self._prepare_input(atoms, coords), forces=True, validate_species=self.validate_species)
        self._cache_key, self._cache_results = key, res
        return {"energy": self._energy(res), "forces": self._forces(res)}

    def get_hessian(self, atoms, coords):
        res = self.model(self._prepare_input(atoms, coords), forces=True, hessian=True, validate_species=self.validate_species)
        self._cache_key, self._cache_results = self._key(atoms, coords), res
        scale = self._ev2au() / ANG2BOHR / ANG2BOHR
        return {"energy": self._energy(res), "forces": self._forces(res),
                "hessian": (res["hessian"].detach().flatten(0, 1).flatten(-2, -1) * scale).to(torch.double).cpu().numpy()}


def run_pysis():
    if _PYSIS_IMPORT_ERROR is not None:
        raise ImportError("AIMNet2Pysis requires PySisyphus. Install it with `pip install pysisyphus`.") from _PYSIS_IMPORT_ERROR
    pysisyphus.run.CALC_DICT["aimnet"] = AIMNet2Pysis
    pysisyphus.run.run()
