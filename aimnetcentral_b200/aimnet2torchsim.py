"""`AIMNet2TorchSim` — TorchSim `ModelInterface` around the B200 `AIMNet2Calculator`, same surface as the reference's adapter
(aimnet/calculators/aimnet2torchsim.py:39-191): constructor keywords, `compute_forces` / `compute_stress` setters that keep
`implemented_properties` in step, `forward(state)` -> dict of detached tensors with `partial_charges` aliasing `charges`,
per-system `charge` / `mult` (or `spin`) extras, periodic states passed on as batched row-vector cells + `pbc`.

The engine side needs nothing special: a TorchSim state is a flat atom list with `system_idx`, i.e. the calculator's flat
input form; batched periodic systems go through the engine's per-system cells (`(S, 3, 3)`), which is also how replicas
shard across GPUs."""
from __future__ import annotations

from collections.abc import Mapping
from typing import Any

import torch
from torch import Tensor

try:  # TorchSim is optional (not in this image): the adapter is importable without it and says so when constructed
    from torch_sim.models.interface import ModelInterface
    from torch_sim.state import SimState
except ImportError as exc:
    _TORCHSIM_IMPORT_ERROR: ImportError | None = exc

    class ModelInterface(torch.nn.Module):  # type: ignore[no-redef]
        """Minimal stand-in with the two properties the adapter overrides."""

        @property
        def compute_forces(self) -> bool:
            return self._compute_forces

        @compute_forces.setter
        def compute_forces(self, value: bool) -> None:
            self._compute_forces = bool(value)

        @property
        def compute_stress(self) -> bool:
            return self._compute_stress

        @compute_stress.setter
        def compute_stress(self, value: bool) -> None:
            self._compute_stress = bool(value)

        @property
        def device(self):
            return self._device

        @property
        def dtype(self):
            return self._dtype

    SimState = Any  # type: ignore[misc, assignment]
else:
    _TORCHSIM_IMPORT_ERROR = None


class AIMNet2TorchSim(ModelInterface):
    def __init__(self, base_calc, *, compute_forces: bool = True, compute_stress: bool = False,
                 validate_species: bool = True) -> None:
        if _TORCHSIM_IMPORT_ERROR is not None:
            raise ImportError("AIMNet2TorchSim requires TorchSim (Python 3.12+ only). Install it with "
                              "`pip install torch-sim-atomistic`.") from _TORCHSIM_IMPORT_ERROR
        super().__init__()
        self._base_calc = base_calc
        self._device = torch.device(base_calc.device)
        self._dtype = torch.float32   # the engine computes in float32 whatever the state's dtype is
        self._compute_forces = bool(compute_forces)
        self._compute_stress = bool(compute_stress)
        self._validate_species = validate_species
        self._memory_scales_with = "n_atoms_x_density"
        self._update_implemented_properties()

    @property
    def base_calc(self):
        return self._base_calc

    @property
    def metadata(self) -> Mapping[str, Any] | None:
        return self._base_calc.metadata

    @ModelInterface.compute_forces.setter
    def compute_forces(self, value: bool) -> None:
        self._compute_forces = bool(value)
        self._update_implemented_properties()

    @ModelInterface.compute_stress.setter
    def compute_stress(self, value: bool) -> None:
        self._compute_stress = bool(value)
        self._update_implemented_properties()

    def forward(self, state: SimState, **kwargs: Any) -> dict[str, Tensor]:
        if state.device != self._device or state.dtype != self._dtype:
            state = state.to(self._device, self._dtype)
        results = self._base_calc(self._state_to_aimnet2_data(state), forces=self._compute_forces,
                                  stress=self._compute_stress, validate_species=self._validate_species)
        if "charges" in results:
            results["partial_charges"] = results["charges"]
        return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in results.items()}

    def _state_to_aimnet2_data(self, state: SimState) -> dict[str, Tensor]:
        data: dict[str, Tensor] = {
            "coord": state.positions.clone(),   # the state's own tensor is never handed to the calculator
            "numbers": state.atomic_numbers.to(torch.int64),
            "mol_idx": state.system_idx.to(torch.int64),
            "charge": self._system_tensor(state, "charge", default=0.0),
        }
        if self._base_calc.is_nse:
            data["mult"] = self._system_tensor(state, "mult", "spin", default=1.0)
        pbc, cell = state.pbc, state.row_vector_cell
        periodic = bool(torch.as_tensor(pbc, device=self._device, dtype=torch.bool).any()) and not torch.allclose(
            cell, torch.zeros_like(cell))
        if periodic:
            data["cell"] = cell.contiguous()
            data["pbc"] = pbc
        elif self._compute_stress:
            raise ValueError("AIMNet2 stress calculation requires a periodic TorchSim state with a non-zero cell.")
        return data

    def _system_tensor(self, state: SimState, *names: str, default: float) -> Tensor:
        value = None
        for name in names:
            value = getattr(state, name, None)
            if value is not None:
                break
        if value is None:
            return torch.full((state.n_systems,), default, dtype=torch.float32, device=self._device)
        t = torch.as_tensor(value, dtype=torch.float32, device=self._device).reshape(-1)
        if t.numel() == 1:
            return t.expand(state.n_systems)
        if t.numel() != state.n_systems:
            raise ValueError(f"TorchSim system extra '{'/'.join(names)}' must be scalar or have one value per system "
                             f"({state.n_systems}); got {t.numel()} values.")
        return t

    def _update_implemented_properties(self) -> None:
        props = ["energy"]
        if self._compute_forces:
            props.append("forces")
        if self._compute_stress:
            props.append("stress")
        props += ["charges", "partial_charges"]
        if self._base_calc.is_nse:
            props.append("spin_charges")
        self.implemented_properties = props
