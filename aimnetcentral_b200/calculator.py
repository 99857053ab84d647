"""`AIMNet2Calculator` — drop-in for the reference's calculator facade on the E / F / charges / stress path
(aimnet/calculators/calculator.py:130-165, 377-515, 638-783, 879-947), backed by the sm_100a engine through the C ABI.

Same constructor signature, `keys_in` / `keys_in_optional` / `keys_out`, input conventions (flat `(N,3)` + `mol_idx`, or
dense `(B,N,3)` with `numbers == 0` padding; numpy / lists / tensors accepted), output dtypes/shapes (energy `(B,)` f64,
charges / forces in the input's atom layout, stress `(3,3)` or `(B,3,3)`), errors and warnings (PBC auto-switch
simple -> DSF scoped to one eval, Ewald without a cell, unsupported species).  Out of this path's scope and rejected
loudly: Hessians, training mode, torch.compile.
"""
from __future__ import annotations

import math
import os
import warnings
import weakref
from types import MappingProxyType, SimpleNamespace
from typing import Any, ClassVar, Mapping

import numpy as np
import torch
from torch import Tensor, nn

from .engine import Engine
from .model_spec import ModelSpec


def _load_model_source(model) -> tuple[dict, dict | None, int]:
    """Return (state_dict, metadata, num_charge_channels) from the accepted model sources."""
    if isinstance(model, tuple) and len(model) == 2 and isinstance(model[1], ModelSpec):
        sd, spec = model
        return dict(sd), spec.metadata(), spec.num_charge_channels
    if isinstance(model, nn.Module):
        meta = getattr(model, "metadata", None)
        if meta is None or callable(meta):
            meta = getattr(model, "_metadata", None)
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        return sd, (dict(meta) if meta is not None else None), int(getattr(model, "num_charge_channels", 1))
    if isinstance(model, str):
        path = model
        if not os.path.isfile(path):
            cache = os.environ.get("AIMNET_CACHE_DIR", os.path.expanduser("~/.cache/aimnet"))
            for cand in (os.path.join(cache, model), os.path.join(cache, model + ".pt")):
                if os.path.isfile(cand):
                    path = cand
                    break
            else:
                raise FileNotFoundError(
                    f"model '{model}' is neither a file nor present in {cache}; registry / Hugging Face downloads are "
                    "outside this engine's scope (SURVEY.md §2 row 13) — pass a v2 .pt path, a (state_dict, ModelSpec) "
                    "pair or an nn.Module")
        model = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(model, Mapping):
        if "state_dict" not in model:
            raise TypeError("model mapping must be a v2 artifact with a 'state_dict' entry (docs/model_format.md)")
        meta = {k: v for k, v in model.items() if k not in ("state_dict", "model_yaml")}
        C = 1
        yml = model.get("model_yaml")
        if isinstance(yml, str):
            import yaml

            yml = yaml.safe_load(yml)
        if isinstance(yml, Mapping):
            C = int(yml.get("kwargs", {}).get("num_charge_channels", 1))
        elif "conv_q.agh" in model["state_dict"]:
            C = int(model["state_dict"]["conv_q.agh"].shape[0])
        return dict(model["state_dict"]), meta, C
    raise TypeError("Invalid model type/name.")


class AIMNet2Calculator:
    keys_in: ClassVar[dict[str, torch.dtype]] = {"coord": torch.float, "numbers": torch.int, "charge": torch.float}
    keys_in_optional: ClassVar[dict[str, torch.dtype]] = {
        "mult": torch.float, "mol_idx": torch.int, "nbmat": torch.int, "nbmat_lr": torch.int,
        "nb_pad_mask": torch.bool, "nb_pad_mask_lr": torch.bool, "shifts": torch.float, "shifts_lr": torch.float,
        "cell": torch.float, "pbc": torch.bool,
    }
    keys_out: ClassVar[list[str]] = ["energy", "charges", "spin_charges", "forces", "hessian", "stress"]
    atom_feature_keys: ClassVar[list[str]] = ["coord", "numbers", "charges", "spin_charges", "forces"]

    def __init__(self, model: Any = "aimnet2", nb_threshold: int = 120, needs_coulomb: bool | None = None,
                 needs_dispersion: bool | None = None, device: str | None = None, compile_model: bool = False,
                 compile_kwargs: dict | None = None, cache_static: bool = False, train: bool = False,
                 deterministic: bool = False, ensemble_member: int = 0, revision: str | None = None,
                 token: str | None = None, *, model_import_paths=None, model_import_mode: str = "extend",
                 neighbor_skin: float = 0.0, cuda_graph: bool = False):
        if device is None:
            device = "cuda"
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("aimnetcentral_b200.AIMNet2Calculator needs a CUDA device: the hot path is hand-written "
                               "sm_100a kernels with no CPU fallback")
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA is not available")
        if train:
            raise NotImplementedError("training mode is outside the inference engine's scope (SURVEY.md §2 row 18)")
        if compile_model:
            warnings.warn("compile_model is ignored: the engine does not use torch.compile", stacklevel=2)
        self.device = str(torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device()))
        sd, metadata, C = _load_model_source(model)
        self._metadata = metadata
        self._num_charge_channels = C
        self.cutoff = float(metadata.get("cutoff", 5.0)) if metadata else 5.0
        final_needs_coulomb = needs_coulomb if needs_coulomb is not None else bool(metadata and metadata.get("needs_coulomb", False))
        final_needs_dispersion = needs_dispersion if needs_dispersion is not None else bool(metadata and metadata.get("needs_dispersion", False))
        sr_embedded = bool(metadata) and metadata.get("coulomb_mode") == "sr_embedded"
        if not sr_embedded and "outputs.srcoulomb.rc" not in sd and metadata is not None and metadata.get("coulomb_mode") not in (None, "sr_embedded"):
            raise NotImplementedError("only v2 models with embedded SRCoulomb (coulomb_mode='sr_embedded') are supported")
        sr_rc = (metadata or {}).get("coulomb_sr_rc") or 4.6
        sr_env = (metadata or {}).get("coulomb_sr_envelope") or "exp"
        self._d3_params = None
        if final_needs_dispersion:
            self._d3_params = (metadata or {}).get("d3_params")
            if self._d3_params is None:
                raise ValueError("needs_dispersion=True but d3_params not found in metadata. "
                                 "Provide d3_params in model metadata or set needs_dispersion=False.")
        self._has_coulomb = bool(final_needs_coulomb)
        self._has_dftd3 = bool(final_needs_dispersion)
        self.nb_threshold = nb_threshold
        self.cache_static = bool(cache_static)
        self._deterministic = bool(deterministic)
        self._coulomb_method: str | None = "simple" if self._has_coulomb else None
        self._dsf_alpha, self._dsf_rc, self._ewald_accuracy = 0.2, 15.0, 1e-6
        self._default_dsf_cutoff = 15.0
        self._default_dftd3_cutoff = 15.0
        self._default_dftd3_smoothing = 0.2
        self._coulomb_cutoff: float | None = float("inf") if self._has_coulomb else None
        self._dftd3_cutoff = self._default_dftd3_cutoff
        self._dftd3_smoothing = self._default_dftd3_smoothing
        self.cutoff_lr = float("inf") if self._has_coulomb else (self._dftd3_cutoff if self._has_dftd3 else None)
        self._mult_ignored_checked = False
        self._species_validation_cache = None   # (key, weakref(numbers tensor), impl) — calculator.py:808-826
        self._impl_lut = None
        self._host_cell_cache = None            # (key, weakref(cell tensor), host copy)
        self._upload_cache: dict = {}           # key -> (host copy, device tensor) of static per-system inputs
        self._validated_numbers = None          # the cached host copy of `numbers` that passed species validation
        self._batch: int | None = None
        # extension over the reference API: Verlet skin (A) for neighbor-list reuse across MD steps; 0 = rebuild every call
        self._neighbor_skin = float(neighbor_skin)
        if self.cache_static and self._neighbor_skin <= 0.0:
            # cache_static (calculator.py:1091-1238 of the reference: reuse the neighbor matrices while the caller keeps
            # passing the same, unmodified coordinates) maps onto the engine's list reuse with a vanishing skin: the
            # device-side displacement check replaces the reference's tensor-identity bookkeeping
            self._neighbor_skin = 1.0e-3
        self.engine = Engine(sd, C, self.device, sr_rc=float(sr_rc), sr_envelope=sr_env, load_d3=self._has_dftd3)
        # duck-typed `model` handle for the adapters that read `base_calc.model._metadata` / `.num_charge_channels`
        # (aimnet/calculators/aimnet2ase.py:69, aimnet2torchsim.py:78-125)
        self.model = SimpleNamespace(_metadata=metadata, metadata=metadata, num_charge_channels=C)
        self.engine.set_deterministic(self._deterministic)
        # extension over the reference API: replay repeated fixed-shape evaluations as one CUDA graph (MD / optimizer loops
        # on small systems are launch-bound); see aimnet2_engine_enable_cuda_graph
        self.cuda_graph = bool(cuda_graph)
        if self.cuda_graph:
            self.engine.enable_cuda_graph(True)
        self._push_options()

    # ---- properties (calculator.py:380-515) ------------------------------------------------------------------
    @property
    def metadata(self):
        return MappingProxyType(self._metadata) if self._metadata is not None else None

    @property
    def is_nse(self) -> bool:
        return self._num_charge_channels == 2

    @property
    def has_external_coulomb(self) -> bool:
        return self._has_coulomb

    @property
    def has_external_dftd3(self) -> bool:
        return self._has_dftd3

    @property
    def coulomb_method(self) -> str | None:
        return self._coulomb_method if self._has_coulomb else None

    @property
    def coulomb_cutoff(self) -> float | None:
        return self._coulomb_cutoff

    @property
    def dftd3_cutoff(self) -> float:
        return self._dftd3_cutoff

    # ---- LR configuration (calculator.py:638-783) ------------------------------------------------------------
    def _push_options(self, coulomb_override: str | None = None):
        method = coulomb_override or self._coulomb_method
        d3 = self._d3_params or {}
        self.engine.set_options(
            coulomb_method=method if self._has_coulomb else None, dsf_alpha=float(self._dsf_alpha),
            dsf_rc=float(self._dsf_rc), ewald_accuracy=float(self._ewald_accuracy), dispersion=self._has_dftd3,
            d3_s6=float(d3.get("s6", 1.0)), d3_s8=float(d3.get("s8", 0.0)), d3_a1=float(d3.get("a1", 0.0)),
            d3_a2=float(d3.get("a2", 0.0)), d3_cutoff=float(self._dftd3_cutoff), d3_smoothing=float(self._dftd3_smoothing),
            sr_cutoff=float(self.cutoff), neighbor_skin=float(getattr(self, "_neighbor_skin", 0.0)))

    def set_lrcoulomb_method(self, method: str, cutoff: float = 15.0, dsf_alpha: float = 0.2, ewald_accuracy: float = 1e-6):
        if method not in ("simple", "dsf", "ewald", "pme"):
            raise ValueError(f"Invalid method: {method}")
        if method == "pme":
            raise NotImplementedError("PME is outside the configured hot path (SURVEY.md §2 row 9)")
        if not self._has_coulomb:
            return
        self._coulomb_method = method
        if method == "dsf":
            self._dsf_alpha, self._dsf_rc = dsf_alpha, cutoff
            self._coulomb_cutoff = cutoff
        elif method == "simple":
            self._coulomb_cutoff = float("inf")
        else:
            self._ewald_accuracy = ewald_accuracy
            self._coulomb_cutoff = None
        self.cutoff_lr = self._coulomb_cutoff if self._coulomb_cutoff is not None else (
            self._dftd3_cutoff if self._has_dftd3 else None)
        self._push_options()

    def set_lr_cutoff(self, cutoff: float) -> None:
        if self._coulomb_method not in ("ewald", "pme"):
            self._coulomb_cutoff = cutoff
            if self._coulomb_method == "dsf":
                self._dsf_rc = cutoff
        self._dftd3_cutoff = cutoff
        self.cutoff_lr = cutoff
        self._push_options()

    def set_dftd3_cutoff(self, cutoff: float | None = None, smoothing_fraction: float | None = None) -> None:
        self._dftd3_cutoff = self._default_dftd3_cutoff if cutoff is None else cutoff
        self._dftd3_smoothing = self._default_dftd3_smoothing if smoothing_fraction is None else smoothing_fraction
        self._push_options()

    # ---- validation (calculator.py:785-851) ------------------------------------------------------------------
    def _validate_species_and_charge(self, data: dict) -> None:
        """calculator.py:785-851.  The species part is skipped when the same `numbers` tensor was validated before
        (identity + `_version` cache, calculator.py:808-826); other inputs (numpy arrays, lists) are checked with one
        vectorised table lookup instead of a Python set over every atom."""
        if "numbers" not in data:
            return
        impl = (self._metadata or {}).get("implemented_species") or []
        if impl:
            numbers = data["numbers"]
            key = self._numbers_validation_key(numbers)
            cached = self._species_validation_cache
            hit = (key is not None and cached is not None and cached[0] == key and cached[1]() is numbers
                   and cached[2] is impl)
            if not hit and key is None:
                # host input (numpy / list): unchanged VALUES since the last validated call count as validated (the same
                # comparison decides whether the device copy is reused, see _upload)
                c = self._upload_cache.get("numbers")
                host = np.asarray(numbers)
                hit = (c is not None and self._validated_numbers is c[0] and c[0].shape == host.shape
                       and c[0].dtype == host.dtype and np.array_equal(c[0], host))
            if not hit:
                self._validate_numbers(numbers, impl)
                if key is not None:
                    self._species_validation_cache = (key, weakref.ref(numbers), impl)
        if (self._metadata or {}).get("supports_charged_systems") is False:
            charge_t = torch.as_tensor(data.get("charge", 0.0))
            if charge_t.numel() > 0 and float(charge_t.abs().max().item()) > 1e-6:
                raise ValueError("This model does not support net-charged systems. Pass validate_species=False to bypass.")

    @staticmethod
    def _numbers_validation_key(value):
        """calculator.py:853-870: identity key of a tensor (`_version` guards in-place mutation); None for inputs whose
        mutation cannot be detected (numpy arrays, lists) — those are re-validated every call."""
        if isinstance(value, Tensor) and value.layout == torch.strided:
            return (id(value), value.data_ptr(), value._version, tuple(value.shape), value.dtype, value.device)
        return None

    def _validate_numbers(self, numbers, impl) -> None:
        lut = self._impl_lut
        if lut is None or lut[0] is not impl:
            ok = np.zeros(256, bool)
            ok[[int(z) for z in impl if 0 <= int(z) < 256]] = True
            lut = self._impl_lut = (impl, ok)
        if isinstance(numbers, Tensor):
            z = numbers.detach().flatten()
            if z.is_cuda:
                # one small D2H (the histogram) instead of every atom
                z = torch.bincount(z.clamp(0, 255).to(torch.int64), minlength=256).cpu().numpy()
                present = np.nonzero(z)[0]
            else:
                present = self._present_species(z.numpy())
        else:
            present = self._present_species(np.asarray(numbers))
        present = present[present > 0]
        bad = [int(v) for v in present if v > 255 or not lut[1][int(v)]]
        if bad:
            raise ValueError(f"Atomic numbers {sorted(bad)} are not in this model's implemented_species "
                             f"{sorted(impl)}. Pass validate_species=False to bypass.")

    @staticmethod
    def _present_species(z: np.ndarray) -> np.ndarray:
        z = z.reshape(-1)
        if z.size and z.dtype.kind in "iu" and int(z.min()) >= 0 and int(z.max()) < 4096:
            return np.nonzero(np.bincount(z, minlength=1))[0]   # histogram: no sort of every atom
        return np.unique(z)

    def _host_cell(self, raw, cell_dev: Tensor) -> np.ndarray:
        """Host copy of the cell for the engine's grid sizing without a per-call device sync: taken from the caller's
        host data when there is one, else cached per CUDA tensor (identity + `_version`)."""
        if not (isinstance(raw, Tensor) and raw.is_cuda):
            return np.ascontiguousarray(np.asarray(raw.detach().numpy() if isinstance(raw, Tensor) else raw, dtype=np.float32))
        key = (id(raw), raw.data_ptr(), raw._version, tuple(raw.shape))
        c = self._host_cell_cache
        if c is not None and c[0] == key and c[1]() is raw:
            return c[2]
        host = np.ascontiguousarray(cell_dev.detach().cpu().numpy().astype(np.float32))
        self._host_cell_cache = (key, weakref.ref(raw), host)
        return host

    def _maybe_warn_mult_ignored(self, data: dict) -> None:
        if self._mult_ignored_checked or self.is_nse or data.get("mult") is None:
            return
        self._mult_ignored_checked = True
        mult_t = torch.as_tensor(data["mult"]).detach().cpu()
        if bool((mult_t != 1).any()):
            warnings.warn(f"Input mult={mult_t.flatten().tolist()} is ignored: this model is closed-shell "
                          "(num_charge_channels=1). For radicals/open-shell systems use an NSE model.",
                          UserWarning, stacklevel=3)

    # ---- evaluation ------------------------------------------------------------------------------------------
    def __call__(self, *args, **kwargs) -> dict[str, Any]:
        return self.eval(*args, **kwargs)

    _STATIC_KEYS = ("numbers", "mol_idx", "charge", "mult", "cell", "pbc")

    def _upload(self, k: str, v, dt) -> Tensor:
        """Host -> device.  The per-system inputs that stay the same from call to call in a screening / MD loop (species,
        molecule index, charges, cell) arrive as numpy arrays or lists; their device copies are kept and reused while the
        host VALUES are unchanged (one memcmp of a few hundred KB instead of a pageable H2D copy + sync per key)."""
        if k in self._STATIC_KEYS and not isinstance(v, Tensor):
            host = np.asarray(v)
            c = self._upload_cache.get(k)
            if c is not None and c[0].shape == host.shape and c[0].dtype == host.dtype and np.array_equal(c[0], host):
                return c[1]
            t = torch.as_tensor(host, device=self.device, dtype=dt)
            self._upload_cache[k] = (host.copy(), t)
            return t
        return torch.as_tensor(v, device=self.device, dtype=dt).detach()

    # ---- second derivatives (calculator.py:904-910, 1247-1450, 1755-1985; derivatives.py:149-192) ----------
    # The reference differentiates its autograd force graph a second time (3N reverse passes; its periodic long-range block is
    # itself a float64 central difference, lr.py `_coul_nvalchemi_fd_hessian`).  The engine has analytic first derivatives and
    # is at its best on batches of molecules, so the Hessian is the central difference of the analytic forces over the 6N
    # displaced copies of the structure, evaluated as molecule batches (taxol: 678 x 113 atoms, two engine calls).
    # fp32 forces carry ~1e-6 eV/A of rounding noise, which a difference over 2h amplifies by 1/h, while the truncation error
    # of the two-point formula grows as h^2 (tools/hessian_step_scan.py: no step gets both below 5e-3 eV/A^2 on the fixtures).
    # The six-point stencil (error h^6) allows a step of 8e-3 A: max error 1-3e-3 eV/A^2 (2-5e-5 of |H|max), rms 2-4e-4.  The
    # stencil weights are solved per displaced coordinate from the displacements actually realised in fp32 (x + h is
    # rounded to the grid of x), and the result is symmetrised (the noise of H[ia,jb] and H[jb,ia] is independent).
    hessian_step = 8.0e-3          # Angstrom
    hessian_stencil = 6            # points per displaced coordinate: 2, 4 or 6
    hessian_batch_atoms = 65_536   # displaced copies are evaluated in chunks of at most this many atoms

    def _fd_weights(self, nodes: Tensor) -> Tensor:
        """First-derivative weights for the (K, m) stencil nodes (in units of the step): solve sum_t w_t u_t^p = [p == 1]."""
        m = nodes.shape[1]
        u = nodes.double()
        V = torch.stack([u ** p for p in range(m)], dim=1)                 # (K, m, m): row p, column t
        rhs = torch.zeros((u.shape[0], m, 1), dtype=torch.float64, device=u.device)
        rhs[:, 1, 0] = 1.0
        return torch.linalg.solve(V, rhs).squeeze(2)                       # (K, m)

    def _single_structures(self, data: dict) -> list[dict] | None:
        """Per-structure inputs of a batched Hessian request (calculator.py:1247-1330), None for a single structure."""
        coord = torch.as_tensor(data["coord"])
        pick = lambda v, b, n: (lambda t: t[b] if t.ndim >= 1 and t.shape[0] == n else t)(torch.as_tensor(v))
        subs = None
        if coord.ndim == 3 and coord.shape[0] > 1:
            B = int(coord.shape[0])
            subs = []
            for b in range(B):
                sub = {}
                for k, v in data.items():
                    if v is None or k.startswith(("nbmat", "shifts")) or k == "mol_idx":
                        continue
                    if k in ("coord", "numbers"):
                        sub[k] = torch.as_tensor(v)[b]
                    elif k in ("charge", "mult"):
                        sub[k] = pick(v, b, B)
                    elif k == "cell":
                        t = torch.as_tensor(v)
                        sub[k] = t[b] if t.ndim == 3 else t
                    else:
                        sub[k] = v
                subs.append(sub)
        elif coord.ndim == 2 and data.get("mol_idx") is not None:
            mi = torch.as_tensor(data["mol_idx"]).to("cpu")
            if mi.numel() and int(mi.max()) > 0:
                n_mol = int(mi.max()) + 1
                subs = []
                for m in range(n_mol):
                    sel = (mi == m).nonzero().squeeze(1)
                    sub = {}
                    for k, v in data.items():
                        if v is None or k.startswith(("nbmat", "shifts")) or k == "mol_idx":
                            continue
                        if k in ("coord", "numbers"):
                            t = torch.as_tensor(v)
                            sub[k] = t[sel.to(t.device)]
                        elif k in ("charge", "mult"):
                            sub[k] = pick(v, m, n_mol)
                        elif k == "cell":
                            t = torch.as_tensor(v)
                            sub[k] = t[m] if t.ndim == 3 else t
                        else:
                            sub[k] = v
                    subs.append(sub)
        return subs

    def _eval_hessian(self, data: dict, *, forces: bool, stress: bool, validate_species: bool) -> dict:
        subs = self._single_structures(data)
        if subs is not None:
            stack = torch.as_tensor(data["coord"]).ndim == 3
            results = [self._eval_hessian(sub, forces=forces, stress=stress, validate_species=validate_species) for sub in subs]
            out = {}
            for k in results[0]:
                vals = [r[k] for r in results]
                if stack and all(v.shape == vals[0].shape for v in vals):
                    out[k] = torch.stack(vals, dim=0)
                else:
                    out[k] = vals
            return out
        d = self.to_input_tensors(data)
        coord = d["coord"]
        single = {k: v for k, v in data.items() if v is not None and not k.startswith(("nbmat", "shifts")) and k != "mol_idx"}
        if coord.ndim == 3:   # a batch of one
            coord = coord[0]
            single["coord"], single["numbers"] = coord, d["numbers"].reshape(coord.shape[0])
        numbers = d["numbers"].reshape(-1)
        real = (numbers > 0).nonzero().squeeze(1)
        base = self.eval({**single, "coord": coord, "numbers": numbers}, forces=True, stress=stress, validate_species=validate_species)
        x = coord.index_select(0, real).contiguous()          # real atoms only: padding rows get zero blocks
        z = numbers.index_select(0, real)
        n = int(x.shape[0])
        per_struct = {k: d[k].reshape(-1)[:1] for k in ("charge", "mult") if k in d}
        cell, pbc = d.get("cell"), d.get("pbc")
        if cell is not None and cell.ndim == 3:
            cell = cell[0]
        H = torch.zeros((3 * n, 3 * n), dtype=torch.float32, device=self.device)
        m = int(self.hessian_stencil)
        if m not in (2, 4, 6):
            raise ValueError("hessian_stencil must be 2, 4 or 6")
        offsets = torch.tensor([o for q in range(1, m // 2 + 1) for o in (q, -q)], dtype=torch.float32, device=self.device)
        h = float(self.hessian_step)
        comps_per_call = max(1, self.hessian_batch_atoms // (m * max(n, 1)))
        for c0 in range(0, 3 * n, comps_per_call):
            comps = torch.arange(c0, min(3 * n, c0 + comps_per_call), device=self.device)
            k = int(comps.numel())
            ia, ax = comps // 3, comps % 3
            xb = x.unsqueeze(0).repeat(m * k, 1, 1).view(m, k, n, 3)
            rows = torch.arange(k, device=self.device)
            xb[:, rows, ia, ax] += offsets.unsqueeze(1) * h
            nodes = ((xb[:, rows, ia, ax] - x[ia, ax].unsqueeze(0)) / h).t()          # (k, m): realised, exact in fp32
            batch = {"coord": xb.view(m * k, n, 3), "numbers": z.unsqueeze(0).expand(m * k, n)}
            for key, v in per_struct.items():
                batch[key] = v.expand(m * k)
            if cell is not None:
                batch["cell"] = cell.unsqueeze(0).expand(m * k, 3, 3).contiguous()
                if pbc is not None:
                    batch["pbc"] = pbc
            f = self.eval(batch, forces=True, validate_species=False)["forces"].view(m, k, 3 * n)
            w = self._fd_weights(nodes)                                                # (k, m)
            H[comps] = (-(w.t().unsqueeze(2) * f.double()).sum(dim=0) / h).float()
        n_all = int(numbers.shape[0])
        H = 0.5 * (H + H.t())
        Hn = H.view(n, 3, n, 3)
        if n != n_all:
            full = torch.zeros((n_all, 3, n_all, 3), dtype=H.dtype, device=H.device)
            full[real.unsqueeze(1), :, real.unsqueeze(0), :] = Hn.permute(0, 2, 1, 3)
            Hn = full
        out = dict(base)
        out["hessian"] = Hn
        return {k: v for k, v in out.items() if k in self.keys_out}

    def hessian_vector_product(self, data: dict, vectors, *, eps: float = 5e-4, validate_species: bool = True,
                               create_graph: bool = False) -> Tensor:
        """`H @ v` for one structure without forming H (calculator.py:1755-1985): directional central difference of the
        analytic forces (the `hessian_stencil`-point formula), all K directions in one molecule batch.  The step along a
        direction is `max(eps, hessian_step) / max|v|`.  The result is a detached value (`create_graph` is not available: there is no
        autograd graph behind the engine)."""
        if create_graph:
            raise NotImplementedError("hessian_vector_product(create_graph=True): the engine has no autograd graph")
        if validate_species:
            self._validate_species_and_charge(data)
        self._maybe_warn_mult_ignored(data)
        coord_in = torch.as_tensor(data["coord"])
        if coord_in.ndim == 3 and coord_in.shape[0] > 1:
            raise NotImplementedError("hessian_vector_product supports a single structure only (got 3D batch).")
        if coord_in.ndim == 2 and data.get("mol_idx") is not None:
            mi = torch.as_tensor(data["mol_idx"])
            if mi.numel() and int(mi.max()) > 0:
                raise NotImplementedError("hessian_vector_product supports a single structure only (got mol_idx batch).")
        d = self.to_input_tensors(data)
        coord = d["coord"][0] if d["coord"].ndim == 3 else d["coord"]
        numbers = d["numbers"].reshape(-1)
        n = int(coord.shape[0])
        vecs = torch.as_tensor(vectors, device=self.device, dtype=torch.float32)
        one = vecs.ndim == 2
        if one:
            vecs = vecs.unsqueeze(0)
        if tuple(vecs.shape[-2:]) != (n, 3):
            raise ValueError(f"vectors must have trailing shape ({n}, 3); got {tuple(vecs.shape)}")
        K = int(vecs.shape[0])
        m = int(self.hessian_stencil)
        offsets = torch.tensor([o for q in range(1, m // 2 + 1) for o in (q, -q)], dtype=torch.float32, device=self.device)
        step = max(float(eps), self.hessian_step) / vecs.abs().amax(dim=(1, 2)).clamp(min=1e-12)      # (K,): max |dx| = step
        xb = coord.view(1, 1, n, 3) + offsets.view(m, 1, 1, 1) * (step.view(1, K, 1, 1) * vecs.unsqueeze(0))
        batch = {"coord": xb.reshape(m * K, n, 3), "numbers": numbers.unsqueeze(0).expand(m * K, n)}
        for key in ("charge", "mult"):
            if key in d:
                batch[key] = d[key].reshape(-1)[:1].expand(m * K)
        if d.get("cell") is not None:
            cell = d["cell"][0] if d["cell"].ndim == 3 else d["cell"]
            batch["cell"] = cell.unsqueeze(0).expand(m * K, 3, 3).contiguous()
            if d.get("pbc") is not None:
                batch["pbc"] = d["pbc"]
        f = self.eval(batch, forces=True, validate_species=False)["forces"].view(m, K, n, 3)
        w = self._fd_weights(offsets.unsqueeze(0))[0]                                                  # nominal nodes
        hv = (-(w.view(m, 1, 1, 1) * f.double()).sum(dim=0) / step.view(K, 1, 1).double()).float()
        return hv[0] if one else hv

    def to_input_tensors(self, data: dict) -> dict[str, Tensor]:
        """calculator.py:1452-1473."""
        ret = {}
        for k, dt in self.keys_in.items():
            if k not in data:
                raise KeyError(f"Missing key {k} in the input data")
            ret[k] = self._upload(k, data[k], dt)
        for k, dt in self.keys_in_optional.items():
            if k in data and data[k] is not None:
                if k == "shifts":
                    dt = torch.int  # kernels take the integer lattice shifts the neighbor builder produces
                ret[k] = self._upload(k, data[k], dt)
        for k, v in ret.items():
            if v.ndim == 0:
                ret[k] = v.unsqueeze(0)
        return ret

    def eval(self, data: dict, forces=False, stress=False, hessian=False, *, validate_species: bool = True) -> dict:
        if validate_species:
            self._validate_species_and_charge(data)
        self._maybe_warn_mult_ignored(data)
        if hessian:
            return self._eval_hessian(data, forces=forces, stress=stress, validate_species=False)
        d = self.to_input_tensors(data)
        if validate_species and not isinstance(data.get("numbers"), Tensor):
            c = self._upload_cache.get("numbers")
            self._validated_numbers = c[0] if c is not None else None
        coord, numbers, charge = d["coord"], d["numbers"], d["charge"]
        cell = d.get("cell")
        method = self._coulomb_method
        if cell is not None and method == "simple":
            warnings.warn("Switching to DSF Coulomb for PBC for this evaluation; call set_lrcoulomb_method() to select "
                          "a periodic method persistently.", stacklevel=2)
            method = "dsf"  # scoped to this evaluation (calculator.py:1044-1062, 939-947)
        if method in ("ewald", "pme") and cell is None:
            raise ValueError(f"Coulomb method '{method}' requires a periodic 'cell' in the input data. Provide a (3,3) "
                             "or (B,3,3) cell tensor, or switch to a non-periodic method via "
                             "set_lrcoulomb_method('simple' | 'dsf').")
        if method == "ewald" and cell.ndim == 2 and charge.shape[0] != 1:
            raise ValueError("Ewald Coulomb with several systems needs one cell per system: pass cell with shape (B, 3, 3)")
        if stress and cell is None:
            raise AssertionError("Stress calculation requires cell")
        # ---- flatten (mol_flatten, calculator.py:1475-1511): the engine always runs the sparse layout ----
        batch_shape = None
        keep = None
        mult = d.get("mult")
        if coord.ndim == 3:
            Bn, Nn = coord.shape[:2]
            batch_shape = (Bn, Nn)
            numbers2 = numbers.reshape(Bn, Nn)
            real = numbers2 > 0
            mol_idx = torch.arange(Bn, device=self.device, dtype=torch.int32).unsqueeze(1).expand(Bn, Nn)
            if bool(real.all()):
                coord_f, numbers_f, mol_f = coord.reshape(-1, 3), numbers2.reshape(-1), mol_idx.reshape(-1)
            else:
                keep = real.reshape(-1).nonzero().squeeze(1)
                coord_f = coord.reshape(-1, 3).index_select(0, keep)
                numbers_f = numbers2.reshape(-1).index_select(0, keep)
                mol_f = mol_idx.reshape(-1).index_select(0, keep)
            if charge.shape[0] != Bn:
                charge = charge.expand(Bn)
        else:
            coord_f, numbers_f = coord, numbers
            mol_f = d.get("mol_idx")
            if mol_f is None and charge.shape[0] != 1:
                raise ValueError("mol_idx is required when charge has more than one entry")
        coord_f = coord_f.contiguous()
        numbers_f = numbers_f.to(torch.int32).contiguous()
        mol_f = None if mol_f is None else mol_f.to(torch.int32).contiguous()
        charge = charge.to(torch.float32).contiguous()
        if self.is_nse:
            if mult is None:
                raise ValueError("mult key is required for NSE if two channels for charge are not provided")
            mult = mult.to(torch.float32).expand(charge.shape[0]).contiguous()
        else:
            mult = None
        host_cell = None
        if cell is not None:
            cell = cell.to(torch.float32).contiguous()
            host_cell = self._host_cell(data["cell"], cell)
        nbmat, shifts = d.get("nbmat"), d.get("shifts")
        if nbmat is not None:
            nbmat = nbmat.to(torch.int32).contiguous()
            shifts = None if shifts is None else shifts.to(torch.int32).contiguous()
        if method != self._coulomb_method:
            self._push_options(coulomb_override=method)
        try:
            out = self.engine.eval(coord_f, numbers_f, charge, mol_idx=mol_f, mult=mult, cell=cell, pbc=d.get("pbc"),
                                   host_cell=host_cell, nbmat=nbmat, shifts=shifts, forces=bool(forces), stress=bool(stress))
        finally:
            if method != self._coulomb_method:
                self._push_options()
        # ---- un-flatten (process_output, calculator.py:1240-1245) ----
        if batch_shape is not None:
            Bn, Nn = batch_shape
            for k in ("charges", "spin_charges", "forces"):
                if k in out:
                    v = out[k]
                    if keep is not None:
                        full = torch.zeros((Bn * Nn, *v.shape[1:]), dtype=v.dtype, device=v.device)
                        full.index_copy_(0, keep, v)
                        v = full
                    out[k] = v.view(Bn, Nn, *v.shape[1:])
        return {k: v for k, v in out.items() if k in self.keys_out}
