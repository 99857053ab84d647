"""Operator seams of the reference served by the C-ABI kernels.

* `neighbor_list`  — same call signature / return convention as `nvalchemiops.torch.neighbors.neighbor_list` at the
  reference's call sites (aimnet/calculators/neighbors.py:106-125, aimnet/modules/lr.py:388-396), raising
  `NeighborOverflowError` when `max_neighbors` is too small.
* `conv_sv_2d_sp`  — same contract as `aimnet.kernels.conv_sv_2d_sp` (aimnet/kernels/conv_sv_2d_sp_wp.py:616-665),
  autograd-enabled (first order).
* `AdaptiveNeighborList` — the reference's auto-sizing wrapper (aimnet/calculators/neighbors.py:21-147) over our op.
* `dsf_coulomb`, `dftd3`, `estimate_ewald_parameters` — the pair-term kernels with the keyword arguments and return
  tuples of the nvalchemiops calls at aimnet/modules/lr.py:526-540, :1204-1228 and
  aimnet/calculators/calculator.py:1566-1587 (GPU parity tests pending: tests/test_gpu_unverified.py).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
from torch import Tensor

from . import _capi
from ._capi import NeighborOverflowError  # noqa: F401  (re-export)


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _require_cuda(t: Tensor, name: str):
    if not isinstance(t, Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if t.device.type != "cuda":
        raise ValueError(f"{name} must be a CUDA tensor (the B200 kernels have no CPU fallback)")


def neighbor_list(positions: Tensor, cutoff: float, cell: Tensor | None = None, pbc: Tensor | None = None,
                  batch_idx: Tensor | None = None, max_neighbors: int | None = None, half_fill: bool = False,
                  fill_value: int | None = None, method: str | None = None, sorted_rows: bool = True):
    """Full neighbor matrix. Returns (nbmat (N,max_nb) i32, num_neighbors (N,) i32[, shifts (N,max_nb,3) i32])."""
    lib = _capi.load()
    _require_cuda(positions, "positions")
    if half_fill:
        raise ValueError("half_fill=True is not supported (the AIMNet2 path always requests full lists)")
    if positions.ndim != 2 or positions.shape[1] != 3:
        raise ValueError("positions must have shape (N, 3)")
    dev = positions.device
    pos = positions.detach().to(torch.float32).contiguous()
    N = pos.shape[0]
    if fill_value is None:
        fill_value = N
    if max_neighbors is None:
        max_neighbors = max(16, int(0.2 * 4 / 3 * math.pi * min(cutoff, 30.0) ** 3))
    n_sys = 1
    bidx = None
    if batch_idx is not None:
        bidx = batch_idx.detach().to(device=dev, dtype=torch.int32).contiguous()
        n_sys = int(bidx[-1].item()) + 1 if N > 0 else 1
    cell_t = host_cell = pbc_arr = None
    n_cells = 0
    if cell is not None:
        cell_t = cell.detach().to(device=dev, dtype=torch.float32).contiguous()
        if cell_t.ndim == 2:
            cell_t = cell_t.unsqueeze(0)
        n_cells = cell_t.shape[0]
        if n_cells not in (1, n_sys):
            raise ValueError("cell must have shape (3,3), (1,3,3) or (num_systems,3,3)")
        host_cell = np.ascontiguousarray(cell_t.cpu().numpy())
        if pbc is not None:
            p = pbc.detach().cpu().numpy() if isinstance(pbc, Tensor) else np.asarray(pbc)
            pbc_arr = np.ascontiguousarray(np.broadcast_to(p.astype(np.uint8).reshape(-1, 3), (n_cells, 3)))
    nbmat = torch.empty((N, max_neighbors), dtype=torch.int32, device=dev)
    nnb = torch.zeros((N,), dtype=torch.int32, device=dev)
    shifts = torch.empty((N, max_neighbors, 3), dtype=torch.int32, device=dev) if cell is not None else None
    maxc = C.c_int(0)
    with torch.cuda.device(dev):
        rc = lib.aimnet2_neighbor_matrix(
            pos.data_ptr(), N, float(cutoff), cell_t.data_ptr() if cell_t is not None else None,
            host_cell.ctypes.data if host_cell is not None else None,
            pbc_arr.ctypes.data if pbc_arr is not None else None, n_cells,
            bidx.data_ptr() if bidx is not None else None, n_sys, int(max_neighbors), int(fill_value),
            1 if sorted_rows else 0, nbmat.data_ptr(), shifts.data_ptr() if shifts is not None else None,
            nnb.data_ptr(), C.byref(maxc), _stream(dev))
    _capi.check(rc, f"neighbor_list: a row needs {maxc.value} slots, max_neighbors={max_neighbors}")
    if cell is not None:
        return nbmat, nnb, shifts
    return nbmat, nnb


def wrap_positions(positions: Tensor, cell: Tensor, pbc=None, batch_idx: Tensor | None = None) -> Tensor:
    """move_coord_to_cell (aimnet/calculators/neighbors.py:331-381) on the GPU."""
    lib = _capi.load()
    _require_cuda(positions, "positions")
    dev = positions.device
    pos = positions.detach().to(torch.float32).contiguous()
    cell_t = cell.detach().to(device=dev, dtype=torch.float32).contiguous()
    if cell_t.ndim == 2:
        cell_t = cell_t.unsqueeze(0)
    n_cells = cell_t.shape[0]
    pbc_arr = None
    if pbc is not None:
        p = pbc.detach().cpu().numpy() if isinstance(pbc, Tensor) else np.asarray(pbc)
        pbc_arr = np.ascontiguousarray(np.broadcast_to(p.astype(np.uint8).reshape(-1, 3), (n_cells, 3)))
    bidx = None if batch_idx is None else batch_idx.detach().to(device=dev, dtype=torch.int32).contiguous()
    out = torch.empty_like(pos)
    with torch.cuda.device(dev):
        rc = lib.aimnet2_wrap_positions(pos.data_ptr(), out.data_ptr(), pos.shape[0], cell_t.data_ptr(), n_cells,
                                        pbc_arr.ctypes.data if pbc_arr is not None else None,
                                        bidx.data_ptr() if bidx is not None else None, _stream(dev))
    _capi.check(rc, "wrap_positions")
    return out


class AdaptiveNeighborList:
    """Auto-sizing wrapper, same policy as the reference (aimnet/calculators/neighbors.py:21-147): start from a density
    estimate, grow x1.5 on overflow, shrink with hysteresis, trim columns to the actual maximum."""

    def __init__(self, cutoff: float, density: float = 0.2, target_utilization: float = 0.75):
        self.cutoff = cutoff
        self.target_utilization = target_utilization
        self.max_neighbors = self._round_to_16(int(density * 4 / 3 * math.pi * min(cutoff, 30.0) ** 3))

    @staticmethod
    def _round_to_16(n: int) -> int:
        return ((n + 15) // 16) * 16

    def __call__(self, positions, cell=None, pbc=None, batch_idx=None, fill_value=None):
        N = positions.shape[0]
        if fill_value is None:
            fill_value = N
        while True:
            try:
                out = neighbor_list(positions, self.cutoff, cell=cell, pbc=pbc, batch_idx=batch_idx,
                                    max_neighbors=self.max_neighbors, fill_value=fill_value)
            except NeighborOverflowError:
                self.max_neighbors = self._round_to_16(int(self.max_neighbors * 1.5))
                continue
            nbmat, nnb = out[0], out[1]
            shifts = out[2] if cell is not None else None
            actual_max = int(nnb.max().item()) if N else 0
            if actual_max < (2 / 3) * self.target_utilization * self.max_neighbors:
                self.max_neighbors = max(self._round_to_16(int(actual_max / self.target_utilization)), 16)
            w = max(1, actual_max)
            return nbmat[:, :w], nnb, (shifts[:, :w] if shifts is not None else None)


# ---------------------------------------------------------------------------------------------------------------
class _ConvSV2dSP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, idx, g):
        lib = _capi.load()
        B, A, G = a.shape
        M = idx.shape[1]
        out = torch.empty((B, A, G, 4), dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            rc = lib.aimnet2_conv_sv_2d_sp_fwd(a.data_ptr(), idx.data_ptr(), g.data_ptr(), out.data_ptr(), B, A, G, M,
                                               _stream(a.device))
        _capi.check(rc, "conv_sv_2d_sp_fwd")
        ctx.save_for_backward(a, idx, g)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _capi.load()
        a, idx, g = ctx.saved_tensors
        B, A, G = a.shape
        M = idx.shape[1]
        go = grad_out.contiguous()
        grad_a = torch.empty_like(a)
        grad_g = torch.empty_like(g)
        with torch.cuda.device(a.device):
            rc = lib.aimnet2_conv_sv_2d_sp_bwd(go.data_ptr(), a.data_ptr(), idx.data_ptr(), g.data_ptr(),
                                               grad_a.data_ptr(), grad_g.data_ptr(), B, A, G, M, _stream(a.device))
        _capi.check(rc, "conv_sv_2d_sp_bwd")
        return grad_a, None, grad_g


def conv_sv_2d_sp(a: Tensor, idx: Tensor, g: Tensor) -> Tensor:
    """out[b,a,g,:] = sum_m a[idx[b,m],a,g] * g[b,m,g,:] with padding value B-1 (validation as
    aimnet/kernels/conv_sv_2d_sp_wp.py:649-665)."""
    for name, t in (("a", a), ("idx", idx), ("g", g)):
        _require_cuda(t, name)
    if a.dtype != torch.float32 or g.dtype != torch.float32:
        raise TypeError("conv_sv_2d_sp: a and g must be float32")
    if a.ndim != 3 or idx.ndim != 2 or g.ndim != 4 or g.shape[-1] != 4:
        raise ValueError("conv_sv_2d_sp: expected a (B,A,G), idx (B,M), g (B,M,G,4)")
    if idx.shape[0] != a.shape[0] or g.shape[0] != a.shape[0] or g.shape[1] != idx.shape[1] or g.shape[2] != a.shape[2]:
        raise ValueError("conv_sv_2d_sp: inconsistent shapes")
    if not (a.is_contiguous() and idx.is_contiguous() and g.is_contiguous()):
        raise ValueError("conv_sv_2d_sp: inputs must be contiguous")
    return _ConvSV2dSP.apply(a, idx.to(torch.int32), g)


# ----------------------------------------------------------------------------------------------------------------
# Pair-term seams (SURVEY.md §8b B3 iii): drop-ins for the nvalchemiops interaction kernels at the reference's call
# sites.  Same argument names and return conventions as those calls; torch owns every buffer.
# ----------------------------------------------------------------------------------------------------------------
HARTREE = 27.211386024367243   # aimnet/constants.py:6
BOHR = 0.5291772105638411      # aimnet/constants.py:8


def _pair_inputs(positions: Tensor, cell, batch_idx, neighbor_matrix: Tensor, neighbor_matrix_shifts, num_systems: int):
    _require_cuda(positions, "positions")
    _require_cuda(neighbor_matrix, "neighbor_matrix")
    if positions.ndim != 2 or positions.shape[1] != 3:
        raise ValueError("positions must have shape (N, 3)")
    dev = positions.device
    n = positions.shape[0]
    if neighbor_matrix.ndim != 2 or neighbor_matrix.shape[0] < n:
        raise ValueError("neighbor_matrix must have shape (N, max_neighbors)")
    nb = neighbor_matrix[:n].detach().to(torch.int32).contiguous()
    bidx = None if batch_idx is None else batch_idx.detach().to(device=dev, dtype=torch.int32).contiguous()
    cell_t = sh = None
    n_cells = 0
    if cell is not None:
        cell_t = cell.detach().to(device=dev, dtype=torch.float32).reshape(-1, 3, 3).contiguous()
        n_cells = cell_t.shape[0]
        if n_cells not in (1, num_systems):
            raise ValueError("cell must have shape (3,3), (1,3,3) or (num_systems,3,3)")
        if neighbor_matrix_shifts is None:
            raise ValueError("neighbor_matrix_shifts is required with a cell")
        sh = neighbor_matrix_shifts[:n].detach().to(torch.int32).contiguous()
        if sh.shape != (*nb.shape, 3):
            raise ValueError("neighbor_matrix_shifts must have shape (N, max_neighbors, 3)")
    return dev, n, nb, bidx, cell_t, n_cells, sh


class _DSFEnergy(torch.autograd.Function):
    """Energy with the charge-response term attached: d E_s / d q_i comes from the kernel (fixed geometry), as the
    reference gets it from autograd through the graph-attached charges (aimnet/modules/lr.py:516-521)."""

    @staticmethod
    def forward(ctx, charges, energy, charge_grad, batch_idx):
        ctx.save_for_backward(charge_grad, batch_idx)
        return energy.clone()

    @staticmethod
    def backward(ctx, grad_energy):
        charge_grad, batch_idx = ctx.saved_tensors
        g = grad_energy[batch_idx.long()] if batch_idx.numel() else grad_energy.expand(charge_grad.shape[0])
        return (g.to(charge_grad.dtype) * charge_grad), None, None, None


def dsf_coulomb(positions: Tensor, charges: Tensor, cutoff: float, alpha: float, cell: Tensor | None = None,
                batch_idx: Tensor | None = None, neighbor_matrix: Tensor | None = None,
                neighbor_matrix_shifts: Tensor | None = None, fill_value: int | None = None, compute_forces: bool = False,
                compute_virial: bool = False, num_systems: int = 1, device: str | None = None):
    """Damped-shifted-force Coulomb with the call signature of `nvalchemiops…dsf_coulomb` at
    aimnet/modules/lr.py:526-540.  Returns `(energy (S,) f64 [e^2/A][, forces (N,3) f32 [e^2/A^2][, virial (S,3,3) f32]])`;
    the caller multiplies by Hartree*Bohr (lr.py:542-547).  `energy` is differentiable with respect to `charges`."""
    if neighbor_matrix is None:
        raise ValueError("dsf_coulomb needs a neighbor_matrix")
    lib = _capi.load()
    dev, n, nb, bidx, cell_t, n_cells, sh = _pair_inputs(positions, cell, batch_idx, neighbor_matrix,
                                                         neighbor_matrix_shifts, int(num_systems))
    pos = positions.detach().to(torch.float32).contiguous()
    q = charges.detach().to(device=dev, dtype=torch.float32).reshape(-1)[:n].contiguous()
    S = int(num_systems)
    energy = torch.empty(S, dtype=torch.float64, device=dev)
    want_f = bool(compute_forces or compute_virial)
    forces = torch.empty(n, 3, dtype=torch.float32, device=dev) if want_f else None
    gq = torch.empty(n, dtype=torch.float32, device=dev)
    virial = torch.empty(S, 3, 3, dtype=torch.float64, device=dev) if compute_virial else None
    p = lambda t: None if t is None else t.data_ptr()   # noqa: E731
    with torch.cuda.device(dev):
        rc = lib.aimnet2_dsf_coulomb(pos.data_ptr(), q.data_ptr(), n, float(cutoff), float(alpha), p(cell_t), n_cells,
                                     p(bidx), S, nb.data_ptr(), p(sh), nb.shape[1],
                                     int(n if fill_value is None else fill_value), energy.data_ptr(), p(forces),
                                     gq.data_ptr(), p(virial), _stream(dev))
    _capi.check(rc, "dsf_coulomb")
    if charges.requires_grad:
        full_gq = gq if charges.reshape(-1).shape[0] == n else torch.cat([gq, gq.new_zeros(charges.numel() - n)])
        idx = bidx if bidx is not None else torch.zeros(0, dtype=torch.int32, device=dev)
        if bidx is not None and charges.numel() != n:
            idx = torch.cat([bidx, bidx.new_zeros(charges.numel() - n)])
        energy = _DSFEnergy.apply(charges.reshape(-1), energy, full_gq.to(charges.dtype), idx)
    out = [energy]
    if want_f:
        out.append(forces)
    if compute_virial:
        out.append(virial.to(torch.float32))
    return tuple(out)


_D3_PACKED: dict = {}


def _pack_d3_tables(c6_reference: Tensor, coord_num_ref: Tensor, dev):
    """(95,95,5,5) C6 reference + reference coordination numbers -> the packed tables of the kernels: C6 rows padded to 28
    floats, reference CN of element z's a-th reference system as a (95,5) table with -1 where c6 == 0 for every partner
    (the same derivation as oracle/make_d3_tables.py; aimnet/modules/lr.py:1405-1422 for the unpacking)."""
    key = (c6_reference.data_ptr(), coord_num_ref.data_ptr(), str(dev))
    hit = _D3_PACKED.get(key)
    if hit is not None:
        return hit[0], hit[1]
    c6 = c6_reference.detach().to(device=dev, dtype=torch.float32)
    cn = coord_num_ref.detach().to(device=dev, dtype=torch.float32)
    if c6.shape != (95, 95, 5, 5):
        raise ValueError("c6_reference must have shape (95, 95, 5, 5)")
    if cn.shape == (95, 5):
        cnref = cn.contiguous()
    elif cn.shape == (95, 95, 5, 5):
        valid = c6 != 0
        cnref = torch.where(valid, cn, torch.full_like(cn, -float("inf"))).amax(dim=(1, 3))
        cnref = torch.where(torch.isfinite(cnref), cnref, torch.full_like(cnref, -1.0)).contiguous()
    else:
        raise ValueError("coord_num_ref must have shape (95, 95, 5, 5) or (95, 5)")
    c6p = torch.nn.functional.pad(c6.reshape(95, 95, 25), (0, 3)).contiguous()
    _D3_PACKED.clear()   # one entry: the tables are process-wide constants in practice
    # the entry keeps the source tensors alive, so a matching data_ptr cannot be a recycled allocation
    _D3_PACKED[key] = (c6p, cnref, c6_reference, coord_num_ref)
    return c6p, cnref


def dftd3(positions: Tensor, numbers: Tensor, a1: float, a2: float, s8: float, s6: float = 1.0,
          covalent_radii: Tensor | None = None, r4r2: Tensor | None = None, c6_reference: Tensor | None = None,
          coord_num_ref: Tensor | None = None, batch_idx: Tensor | None = None, cell: Tensor | None = None,
          neighbor_matrix: Tensor | None = None, neighbor_matrix_shifts: Tensor | None = None,
          fill_value: int | None = None, num_systems: int = 1, compute_virial: bool = False, device: str | None = None,
          s5_smoothing_on: float = 1e10, s5_smoothing_off: float = 1e10):
    """DFT-D3(BJ) with the call signature of `nvalchemiops…dftd3` at aimnet/modules/lr.py:1204-1228: positions, cell and
    the smoothing radii in Bohr; returns `(energy (S,) [Hartree], forces (N,3) [Hartree/Bohr], coord_num (N,)[, virial
    (S,3,3) [Hartree]])`.  The kernels work in Angstrom / eV; the conversions happen here."""
    if neighbor_matrix is None or covalent_radii is None or r4r2 is None or c6_reference is None or coord_num_ref is None:
        raise ValueError("dftd3 needs neighbor_matrix, covalent_radii, r4r2, c6_reference and coord_num_ref")
    lib = _capi.load()
    dev, n, nb, bidx, cell_b, n_cells, sh = _pair_inputs(positions, cell, batch_idx, neighbor_matrix,
                                                         neighbor_matrix_shifts, int(num_systems))
    pos = (positions.detach().to(torch.float32) * BOHR).contiguous()
    cell_t = None if cell_b is None else (cell_b * BOHR).contiguous()
    z = numbers.detach().to(device=dev, dtype=torch.int32).reshape(-1)[:n].contiguous()
    c6p, cnref = _pack_d3_tables(c6_reference, coord_num_ref, dev)
    rcov = covalent_radii.detach().to(device=dev, dtype=torch.float32).contiguous()
    rr = r4r2.detach().to(device=dev, dtype=torch.float32).contiguous()
    S = int(num_systems)
    energy = torch.empty(S, dtype=torch.float64, device=dev)
    forces = torch.empty(n, 3, dtype=torch.float32, device=dev)
    cn = torch.empty(n, dtype=torch.float32, device=dev)
    virial = torch.empty(S, 3, 3, dtype=torch.float64, device=dev) if compute_virial else None
    r_off = float(min(s5_smoothing_off, 1e10))
    r_on = float(min(s5_smoothing_on, r_off))
    if not r_on < r_off:     # no smoothing window requested: a hard cutoff far outside any list
        r_on, r_off = 0.999e10, 1e10
    p = lambda t: None if t is None else t.data_ptr()   # noqa: E731
    with torch.cuda.device(dev):
        rc = lib.aimnet2_dftd3(pos.data_ptr(), z.data_ptr(), n, float(s6), float(s8), float(a1), float(a2), r_on, r_off,
                               c6p.data_ptr(), cnref.data_ptr(), rcov.data_ptr(), rr.data_ptr(), p(cell_t), n_cells, p(bidx),
                               S, nb.data_ptr(), p(sh), nb.shape[1], int(n if fill_value is None else fill_value),
                               energy.data_ptr(), forces.data_ptr(), cn.data_ptr(), p(virial), _stream(dev))
    _capi.check(rc, "dftd3")
    out = [energy / HARTREE, forces * (BOHR / HARTREE), cn]
    if compute_virial:
        out.append((virial / HARTREE).to(torch.float32))
    return tuple(out)


class _EwaldEnergy(torch.autograd.Function):
    """Per-atom energies with the charge response attached (d E / d q_i at fixed geometry from the kernel)."""

    @staticmethod
    def forward(ctx, charges, energies, charge_grad):
        ctx.save_for_backward(charge_grad)
        return energies.clone()

    @staticmethod
    def backward(ctx, grad_e):
        (charge_grad,) = ctx.saved_tensors
        # E_total = sum_i e_i and dE_total/dq_i = charge_grad_i; a caller that weights atoms of one system equally (the
        # reference scatter-adds per system, lr.py:698-703) gets the exact chain rule
        return (grad_e.to(charge_grad.dtype) * charge_grad), None, None


def ewald_summation(positions: Tensor, charges: Tensor, cell: Tensor, batch_idx: Tensor | None = None,
                    neighbor_matrix: Tensor | None = None, neighbor_matrix_shifts: Tensor | None = None,
                    mask_value: int | None = None, accuracy: float = 1e-6, compute_forces: bool = False,
                    compute_virial: bool = False):
    """Ewald summation with the call signature of `nvalchemiops…ewald_summation` at aimnet/modules/lr.py:687-696.
    Returns `energies_per_atom (N,) f64 [e^2/A]` (differentiable with respect to `charges`); with `compute_forces` /
    `compute_virial` additionally `forces (N,3) f32 [e^2/A^2]` / `virial (S,3,3) f32` — the reference takes those from
    autograd through the energies, this library returns them explicitly.  The neighbor matrix must reach every system's
    real-space cutoff (`estimate_ewald_parameters`)."""
    if neighbor_matrix is None or cell is None:
        raise ValueError("ewald_summation needs a cell and a neighbor_matrix")
    lib = _capi.load()
    cells = cell.detach().reshape(-1, 3, 3)
    S = int(cells.shape[0])
    dev, n, nb, bidx, cell_t, n_cells, sh = _pair_inputs(positions, cells, batch_idx, neighbor_matrix, neighbor_matrix_shifts, S)
    if S > 1 and bidx is None:
        raise ValueError("batch_idx is required with more than one cell")
    host_cell = np.ascontiguousarray(cell_t.cpu().numpy(), dtype=np.float32)
    offs = None
    if S > 1:
        counts = torch.bincount(bidx.to(torch.int64), minlength=S).cpu().numpy()
        offs = np.ascontiguousarray(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32))
    pos = positions.detach().to(torch.float32).contiguous()
    q = charges.detach().to(device=dev, dtype=torch.float32).reshape(-1)[:n].contiguous()
    e_atom = torch.empty(n, dtype=torch.float64, device=dev)
    want_f = bool(compute_forces or compute_virial)
    forces = torch.empty(n, 3, dtype=torch.float32, device=dev) if want_f else None
    gq = torch.empty(n, dtype=torch.float32, device=dev)
    virial = torch.empty(S, 3, 3, dtype=torch.float64, device=dev) if compute_virial else None
    p = lambda t: None if t is None else t.data_ptr()   # noqa: E731
    with torch.cuda.device(dev):
        rc = lib.aimnet2_ewald_summation(pos.data_ptr(), q.data_ptr(), n, cell_t.data_ptr(), host_cell.ctypes.data, p(bidx),
                                         None if offs is None else offs.ctypes.data, S, nb.data_ptr(), p(sh), nb.shape[1],
                                         int(n if mask_value is None else mask_value), float(accuracy), e_atom.data_ptr(),
                                         p(forces), gq.data_ptr(), p(virial), _stream(dev))
    _capi.check(rc, "ewald_summation")
    if charges.requires_grad:
        full_gq = gq if charges.reshape(-1).shape[0] == n else torch.cat([gq, gq.new_zeros(charges.numel() - n)])
        full_e = e_atom if charges.reshape(-1).shape[0] == n else torch.cat([e_atom, e_atom.new_zeros(charges.numel() - n)])
        e_atom = _EwaldEnergy.apply(charges.reshape(-1), full_e, full_gq.to(charges.dtype))[:n]
    out = [e_atom]
    if want_f:
        out.append(forces)
    if compute_virial:
        out.append(virial.to(torch.float32))
    return out[0] if len(out) == 1 else tuple(out)


class EwaldParameters:
    """What the reference reads from `estimate_ewald_parameters` (aimnet/calculators/calculator.py:1566-1587)."""

    def __init__(self, alpha: Tensor, real_space_cutoff: Tensor, reciprocal_space_cutoff: Tensor):
        self.alpha = alpha
        self.real_space_cutoff = real_space_cutoff
        self.reciprocal_space_cutoff = reciprocal_space_cutoff


def estimate_ewald_parameters(positions: Tensor, cell: Tensor, batch_idx: Tensor | None = None,
                              accuracy: float = 1e-6) -> EwaldParameters:
    """Per-system Ewald splitting parameters for a target accuracy; host arithmetic only (works on CPU tensors too)."""
    lib = _capi.load()
    cells = cell.detach().to("cpu", torch.float32).reshape(-1, 3, 3)
    n = positions.shape[0]
    if batch_idx is None:
        counts = [n] * cells.shape[0]
    else:
        counts = torch.bincount(batch_idx.detach().to("cpu", torch.int64), minlength=cells.shape[0]).tolist()
    vals = []
    for c, cnt in zip(cells, counts):
        host = np.ascontiguousarray(c.numpy().reshape(-1), dtype=np.float32)
        a, r, k = C.c_double(), C.c_double(), C.c_double()
        _capi.check(lib.aimnet2_estimate_ewald_parameters(host.ctypes.data_as(_capi.c_float_p), int(max(cnt, 1)),
                                                          float(accuracy), C.byref(a), C.byref(r), C.byref(k)),
                    "estimate_ewald_parameters")
        vals.append((a.value, r.value, k.value))
    t = torch.tensor(vals, dtype=torch.float64, device=positions.device)
    return EwaldParameters(t[:, 0], t[:, 1], t[:, 2])
