"""Python handle on the C-ABI engine: packs a reference-layout `state_dict` into the weights struct, owns the engine
object, and marshals torch device tensors / host buffers into `aimnet2_engine_eval` / `aimnet2_engine_eval_host`."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _capi
from .model_spec import ModelSpec

_D3_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "dftd3_tables.npz")


def _np32(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(t, dtype=np.float32))


def _pbc_bytes(pbc, n_cells: int) -> np.ndarray:
    if isinstance(pbc, torch.Tensor):
        pbc = pbc.detach().cpu().numpy()
    p = np.asarray(pbc).astype(np.uint8).reshape(-1, 3)
    return np.ascontiguousarray(np.broadcast_to(p, (n_cells, 3)))


def _ptr(a) -> int | None:
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    return a.ctypes.data


class Engine:
    """One engine = one weight set resident on one GPU."""

    def __init__(self, state_dict: dict, num_charge_channels: int = 1, device: int | str | torch.device = 0,
                 sr_rc: float = 4.6, sr_envelope: str = "exp", load_d3: bool = True):
        self._lib = _capi.load()
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise RuntimeError("aimnetcentral_b200 runs on CUDA devices only (no CPU fallback)")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.C = int(num_charge_channels)
        self._keep = []  # host arrays referenced by the struct during create
        w = _capi.Weights()
        w.num_charge_channels = self.C

        def hold(arr):
            self._keep.append(arr)
            return arr.ctypes.data

        sd = state_dict
        w.afv = hold(_np32(sd["afv.weight"]))
        w.agh_a = hold(_np32(sd["conv_a.agh"]))
        w.agh_q = hold(_np32(sd["conv_q.agh"]))
        w.shifts_s = hold(_np32(sd["aev.shifts_s"]))
        w.eta_s = float(sd["aev.eta_s"])
        w.rc_s = float(sd["aev.rc_s"])
        for p in range(3):
            keys = sorted((k for k in sd if k.startswith(f"mlps.{p}.") and k.endswith(".weight")),
                          key=lambda k: int(k.split(".")[2]))
            ws = [_np32(sd[k]) for k in keys]
            bs = [_np32(sd[k.replace(".weight", ".bias")]) for k in keys]
            dims = np.array([ws[0].shape[1]] + [x.shape[0] for x in ws], dtype=np.int32)
            w.n_layers[p] = len(ws)
            w.layer_dims[p] = hold(dims)
            wp = (C.c_void_p * len(ws))(*[hold(x) for x in ws])
            bp = (C.c_void_p * len(bs))(*[hold(x) for x in bs])
            self._keep += [wp, bp]
            w.mlp_w[p] = C.cast(wp, C.c_void_p)
            w.mlp_b[p] = C.cast(bp, C.c_void_p)
        for li in range(3):
            w.head_w[li] = hold(_np32(sd[f"outputs.energy_mlp.mlp.{2 * li}.weight"]))
            w.head_b[li] = hold(_np32(sd[f"outputs.energy_mlp.mlp.{2 * li}.bias"]))
        sae = sd["outputs.atomic_shift.shifts.weight"]
        sae = sae.detach().cpu().numpy() if isinstance(sae, torch.Tensor) else np.asarray(sae)
        w.sae = hold(np.ascontiguousarray(sae.reshape(-1).astype(np.float64)))
        w.sr_rc = float(sd.get("outputs.srcoulomb.rc", sr_rc))
        w.sr_envelope = 0 if sr_envelope == "exp" else 1
        if load_d3:
            z = np.load(_D3_PATH)
            w.d3_c6ref = hold(_np32(z["c6ref"]))
            w.d3_cnref = hold(_np32(z["cnref"]))
            w.d3_rcov = hold(_np32(z["rcov"]))
            w.d3_r4r2 = hold(_np32(z["r4r2"]))
        h = C.c_void_p()
        _capi.check(self._lib.aimnet2_engine_create(C.byref(h), C.byref(w), self.device.index), "engine_create")
        self._h = h
        self._keep = []
        # engine_create selects the tcgen05 backend when it was built in; mirror that choice here
        self.gemm_backend = 2 if self._lib.aimnet2_engine_set_gemm_backend(h, 2) == 0 else 0
        self.options = dict(coulomb_method="simple", dsf_alpha=0.2, dsf_rc=15.0, ewald_accuracy=1e-6, dispersion=False,
                            d3_s6=1.0, d3_s8=0.3908, d3_a1=0.566, d3_a2=3.128, d3_cutoff=15.0, d3_smoothing=0.2,
                            sr_cutoff=5.0, neighbor_skin=0.0)

    # ------------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.aimnet2_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_options(self, **kw):
        self.options.update(kw)
        o = _capi.Options()
        d = self.options
        o.coulomb_method = _capi.COULOMB[d["coulomb_method"]]
        o.dsf_alpha, o.dsf_rc, o.ewald_accuracy = d["dsf_alpha"], d["dsf_rc"], d["ewald_accuracy"]
        o.dispersion = 1 if d["dispersion"] else 0
        o.d3_s6, o.d3_s8, o.d3_a1, o.d3_a2 = d["d3_s6"], d["d3_s8"], d["d3_a1"], d["d3_a2"]
        o.d3_cutoff, o.d3_smoothing, o.sr_cutoff = d["d3_cutoff"], d["d3_smoothing"], d["sr_cutoff"]
        o.neighbor_skin = float(d.get("neighbor_skin", 0.0))
        _capi.check(self._lib.aimnet2_engine_set_options(self._h, C.byref(o)), "set_options")

    def set_small_m_rows(self, rows: int):
        """Evaluations with at most `rows` atoms use the small-M fp32 SIMT GEMM (default 512, 0 = never)."""
        _capi.check(self._lib.aimnet2_engine_set_small_m_rows(self._h, int(rows)), "set_small_m_rows")

    def set_conv_impl(self, impl: int):
        """0 = list kernels always; batches of small molecules: 1 = dense shared-memory forward + list backward (default),
        2 = dense forward and backward."""
        _capi.check(self._lib.aimnet2_engine_set_conv_impl(self._h, int(impl)), "set_conv_impl")

    def set_dense_min_molecules(self, n_mol: int):
        """Batches with at least `n_mol` molecules take the dense conv walk (default 64)."""
        _capi.check(self._lib.aimnet2_engine_set_dense_min_molecules(self._h, int(n_mol)), "set_dense_min_molecules")

    def conv_mode(self) -> dict:
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._lib.aimnet2_engine_conv_mode(self._h, C.byref(a), C.byref(b), C.byref(c))
        return {"impl": a.value, "dense_last": bool(b.value), "max_molecule_last": c.value}

    def debug_poison(self, byte: int):
        """Test seam: fill the device workspace with `byte` before every evaluation (-1 = off)."""
        _capi.check(self._lib.aimnet2_engine_debug_poison(self._h, int(byte)), "debug_poison")

    def debug_layout(self) -> dict:
        """Test seam: {buffer name: (byte offset, bytes)} of the workspace as carved for the last evaluation."""
        buf = C.create_string_buffer(1 << 16)
        _capi.check(self._lib.aimnet2_engine_debug_layout(self._h, buf, len(buf)), "debug_layout")
        out = {}
        for line in buf.value.decode().splitlines():
            name, off, nbytes = line.split()
            out[name] = (int(off), int(nbytes))
        return out

    def debug_snapshot(self, names=None) -> dict:
        """Test seam: host copies (uint8 arrays) of the named workspace buffers (all when None) after a device sync."""
        snap = {}
        for name, (off, nbytes) in self.debug_layout().items():
            if names is not None and name not in names:
                continue
            arr = np.empty(nbytes, np.uint8)
            _capi.check(self._lib.aimnet2_engine_debug_read_workspace(self._h, arr.ctypes.data, off, nbytes), "debug_read_workspace")
            snap[name] = arr
        return snap

    def set_species_first_pass(self, on: bool):
        """Backward of the first convolution through per-species tables (default) or through the generic pair kernel."""
        _capi.check(self._lib.aimnet2_engine_set_species_first_pass(self._h, int(bool(on))), "set_species_first_pass")

    def set_gemm_backend(self, backend: int):
        _capi.check(self._lib.aimnet2_engine_set_gemm_backend(self._h, int(backend)), "set_gemm_backend")
        self.gemm_backend = int(backend)

    def set_deterministic(self, on: bool = True):
        _capi.check(self._lib.aimnet2_engine_set_deterministic(self._h, 1 if on else 0), "set_deterministic")

    def enable_cuda_graph(self, on: bool = True):
        """Replay repeated fixed-shape evaluations as one CUDA graph (small systems are launch-bound: taxol 0.53 -> ~0.2 ms)."""
        _capi.check(self._lib.aimnet2_engine_enable_cuda_graph(self._h, 1 if on else 0), "enable_cuda_graph")

    def graph_stats(self) -> dict:
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._lib.aimnet2_engine_graph_stats(self._h, C.byref(a), C.byref(b), C.byref(c))
        return {"captures": a.value, "launches": b.value, "fallbacks": c.value}

    def enable_timing(self, level: int = 1):
        """0 off, 1 phase events, 2 additionally one CUDA-event pair around every GEMM launch."""
        _capi.check(self._lib.aimnet2_engine_enable_timing(self._h, int(level)))

    def last_timing(self) -> dict:
        buf = (C.c_float * 9)()
        self._lib.aimnet2_engine_last_timing(self._h, buf, 9)
        return dict(zip(("neighbors_ms", "forward_ms", "pair_terms_ms", "backward_ms", "total_ms", "gemm_ms",
                         "gemm_launches", "conv_ms", "conv_calls"), list(buf)))

    def last_launches(self) -> int:
        return int(self._lib.aimnet2_engine_last_launches(self._h))

    def skin_stats(self) -> tuple[int, int]:
        """(list builds, list reuses) since the engine was created; only counted with neighbor_skin > 0."""
        a, b = C.c_int(), C.c_int()
        self._lib.aimnet2_engine_skin_stats(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def info(self) -> dict:
        a, b, c = C.c_int(), C.c_int(), C.c_int64()
        self._lib.aimnet2_engine_info(self._h, C.byref(a), C.byref(b), C.byref(c))
        sr, lr = C.c_int(), C.c_int()
        self._lib.aimnet2_engine_neighbor_caps(self._h, C.byref(sr), C.byref(lr))
        return {"sr_width": a.value, "lr_width": b.value, "workspace_bytes": c.value, "sr_cap": sr.value, "lr_cap": lr.value}

    # ------------------------------------------------------------------------------------------------------
    def eval(self, coord: torch.Tensor, numbers: torch.Tensor, charge: torch.Tensor, mol_idx: torch.Tensor | None = None,
             mult: torch.Tensor | None = None, cell: torch.Tensor | None = None, pbc=None,
             nbmat: torch.Tensor | None = None, shifts: torch.Tensor | None = None, forces: bool = True,
             stress: bool = False, return_nbmat: bool = False, host_cell: np.ndarray | None = None) -> dict:
        """Device-resident evaluation. coord (N,3) f32, numbers (N) i32, charge (B) f32, mol_idx (N) i32 sorted,
        cell (3,3)|(B,3,3) f32 — all on self.device, contiguous."""
        dev = self.device
        N, B = int(coord.shape[0]), int(charge.shape[0])

        def chk(t, dt, name):
            if t is None:
                return None
            if t.device != dev or t.dtype != dt or not t.is_contiguous():
                raise TypeError(f"{name} must be a contiguous {dt} tensor on {dev}")
            return t

        coord = chk(coord, torch.float32, "coord")
        numbers = chk(numbers, torch.int32, "numbers")
        charge = chk(charge, torch.float32, "charge")
        mol_idx = chk(mol_idx, torch.int32, "mol_idx")
        mult = chk(mult, torch.float32, "mult")
        cell = chk(cell, torch.float32, "cell")
        nbmat = chk(nbmat, torch.int32, "nbmat")
        shifts = chk(shifts, torch.int32, "shifts")
        sys_ = _capi.System()
        sys_.n_atoms, sys_.n_mol = N, B
        sys_.coord, sys_.numbers, sys_.charge = coord.data_ptr(), numbers.data_ptr(), charge.data_ptr()
        sys_.mol_idx = _ptr(mol_idx)
        sys_.mult = _ptr(mult)
        pbc_arr = None
        n_cells = 0
        if cell is not None:
            n_cells = 1 if cell.ndim == 2 else int(cell.shape[0])
            if host_cell is None:   # device sync; callers that know the cell on the host pass it (calculator.py does)
                host_cell = cell.detach().cpu().numpy()
            host_cell = np.ascontiguousarray(np.asarray(host_cell, dtype=np.float32).reshape(n_cells, 3, 3))
            sys_.cell, sys_.host_cell = cell.data_ptr(), host_cell.ctypes.data
            if pbc is not None:
                pbc_arr = _pbc_bytes(pbc, n_cells)
                sys_.pbc_host = pbc_arr.ctypes.data
        sys_.n_cells = n_cells
        if nbmat is not None:
            sys_.nbmat, sys_.nb_width = nbmat.data_ptr(), int(nbmat.shape[1])
            sys_.shifts = _ptr(shifts)
        out = {"energy": torch.empty(B, dtype=torch.float64, device=dev),
               "charges": torch.empty(N, dtype=torch.float32, device=dev)}
        res = _capi.Result()
        res.energy, res.charges = out["energy"].data_ptr(), out["charges"].data_ptr()
        if self.C == 2:
            out["spin_charges"] = torch.empty(N, dtype=torch.float32, device=dev)
            res.spin_charges = out["spin_charges"].data_ptr()
        flags = 0
        if forces:
            out["forces"] = torch.empty(N, 3, dtype=torch.float32, device=dev)
            res.forces = out["forces"].data_ptr()
            flags |= _capi.WANT_FORCES
        if stress:
            out["stress"] = torch.empty((3, 3) if cell.ndim == 2 else (n_cells, 3, 3), dtype=torch.float32, device=dev)
            res.stress = out["stress"].data_ptr()
            flags |= _capi.WANT_STRESS
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = self._lib.aimnet2_engine_eval(self._h, C.byref(sys_), C.byref(res), flags, C.c_void_p(stream))
        _capi.check(rc, "engine_eval")
        del host_cell, pbc_arr
        return out

    def eval_host(self, coord: np.ndarray, numbers: np.ndarray, charge: np.ndarray, mol_idx: np.ndarray | None = None,
                  mult: np.ndarray | None = None, cell: np.ndarray | None = None, pbc=None, forces: bool = True,
                  stress: bool = False, out: dict | None = None) -> dict:
        """Host-buffer evaluation through aimnet2_engine_eval_host (H2D + compute + D2H inside the call).
        Arrays may be numpy arrays or CPU (ideally pinned) torch tensors; dtypes f32 / i32."""

        def as_np(a, dt):
            if a is None:
                return None
            if isinstance(a, torch.Tensor):
                a = a.numpy()
            if a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                a = np.ascontiguousarray(a, dtype=dt)
            return a

        coord, numbers, charge = as_np(coord, np.float32), as_np(numbers, np.int32), as_np(charge, np.float32)
        mol_idx, mult, cell = as_np(mol_idx, np.int32), as_np(mult, np.float32), as_np(cell, np.float32)
        N, B = coord.shape[0], charge.shape[0]
        sys_ = _capi.System()
        sys_.n_atoms, sys_.n_mol = N, B
        sys_.coord, sys_.numbers, sys_.charge = coord.ctypes.data, numbers.ctypes.data, charge.ctypes.data
        sys_.mol_idx, sys_.mult = _ptr(mol_idx), _ptr(mult)
        n_cells = 0
        pbc_arr = None
        if cell is not None:
            n_cells = 1 if cell.ndim == 2 else cell.shape[0]
            sys_.cell = sys_.host_cell = cell.ctypes.data
            if pbc is not None:
                pbc_arr = _pbc_bytes(pbc, n_cells)
                sys_.pbc_host = pbc_arr.ctypes.data
        sys_.n_cells = n_cells
        if out is None:
            out = {"energy": np.empty(B, np.float64), "charges": np.empty(N, np.float32)}
            if self.C == 2:
                out["spin_charges"] = np.empty(N, np.float32)
            if forces:
                out["forces"] = np.empty((N, 3), np.float32)
            if stress:
                out["stress"] = np.empty((3, 3) if cell.ndim == 2 else (n_cells, 3, 3), np.float32)
        res = _capi.Result()
        res.energy, res.charges = _ptr(out["energy"]), _ptr(out["charges"])
        res.spin_charges = _ptr(out.get("spin_charges"))
        flags = 0
        if forces:
            res.forces = _ptr(out["forces"])
            flags |= _capi.WANT_FORCES
        if stress:
            res.stress = _ptr(out["stress"])
            flags |= _capi.WANT_STRESS
        rc = self._lib.aimnet2_engine_eval_host(self._h, C.byref(sys_), C.byref(res), flags)
        _capi.check(rc, "engine_eval_host")
        return out


def engine_from_spec(state_dict: dict, spec: ModelSpec | None = None, device=0) -> Engine:
    spec = spec or ModelSpec()
    e = Engine(state_dict, spec.num_charge_channels, device, sr_rc=spec.coulomb_sr_rc, sr_envelope=spec.coulomb_sr_envelope)
    return e
