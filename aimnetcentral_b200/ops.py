"""Operator seams of the reference served by the C-ABI kernels.

* `neighbor_list`  — same call signature / return convention as `nvalchemiops.torch.neighbors.neighbor_list` at the
  reference's call sites (aimnet/calculators/neighbors.py:106-125, aimnet/modules/lr.py:388-396), raising
  `NeighborOverflowError` when `max_neighbors` is too small.
* `conv_sv_2d_sp`  — same contract as `aimnet.kernels.conv_sv_2d_sp` (aimnet/kernels/conv_sv_2d_sp_wp.py:616-665),
  autograd-enabled (first order).
* `AdaptiveNeighborList` — the reference's auto-sizing wrapper (aimnet/calculators/neighbors.py:21-147) over our op.
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
