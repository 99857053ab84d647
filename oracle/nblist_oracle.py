"""CPU oracle for the neighbor-matrix build.  TEST INFRASTRUCTURE ONLY — never imported by the product path.

Restates, in numpy, the contract of the third-party `nvalchemiops.torch.neighbors.neighbor_list`
(nvalchemi-toolkit-ops 0.4.0, un-vendored; pinned in /root/reference/pyproject.toml:37, uv.lock:2499) as it is
consumed by the reference:

* call sites: aimnet/calculators/neighbors.py:106-125 (full list, `half_fill=False`, `fill_value=N`),
  aimnet/modules/lr.py:388-396;
* geometry convention: r_ij = x_j + s @ cell - x_i  (aimnet/ops.py:37-66), cell rows = lattice vectors;
* image range per periodic axis: |n_k| <= ceil(rc * ||column k of inv(cell)||)  (rule in aimnet/ops.py:171-193);
* padding: unused slots = fill_value, zero shifts (aimnet/calculators/neighbors.py:253-284).

The upstream kernel's row order and the `<` vs `<=` edge rule are unpinned (the reference's own tests compare rows
as sets, tests/test_calculator_gpu.py:143-184); we fix the canonical form: `d2 < rc2` evaluated in float32 with the
exact operation order below, rows sorted by (j, sx, sy, sz).  The CUDA builder must reproduce this bit for bit.

float32 recipe (every operation individually rounded, no FMA):
    sv_k = ((sx*cell[0,k]) + (sy*cell[1,k])) + (sz*cell[2,k])
    r_k  = (x_j[k] + sv_k) - x_i[k]
    d2   = ((r_x*r_x) + (r_y*r_y)) + (r_z*r_z)
    keep if d2 < float32(rc)*float32(rc)   and not (j == i and s == 0)
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def image_ranges(cell: np.ndarray, pbc: np.ndarray, cutoff: float) -> np.ndarray:
    """Number of lattice images to scan along each axis (int, shape (3,))."""
    inv = np.linalg.inv(cell.astype(np.float64))
    # column k of inv(cell) is the reciprocal vector b_k (a_l . b_k = delta_lk); plane spacing = 1/|b_k|
    bnorm = np.sqrt((inv**2).sum(axis=0))
    n = np.ceil(cutoff * bnorm - 1e-9).astype(np.int64)
    n = np.maximum(n, 1)
    return np.where(pbc, n, 0)


def _system_rows(pos: np.ndarray, idx: np.ndarray, cutoff: float, cell, pbc):
    """Neighbor rows for one system. Returns list of (j_global, shifts(k,3)) per atom in idx."""
    rc = F32(cutoff)
    rc2 = F32(rc * rc)
    n = len(idx)
    x = pos[idx].astype(F32)
    if cell is None:
        img = np.zeros((1, 3), np.int32)
        sv = np.zeros((1, 3), F32)
    else:
        nr = image_ranges(cell, pbc, cutoff)
        gx, gy, gz = np.meshgrid(
            np.arange(-nr[0], nr[0] + 1), np.arange(-nr[1], nr[1] + 1), np.arange(-nr[2], nr[2] + 1), indexing="ij"
        )
        img = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], 1).astype(np.int32)
        c = cell.astype(F32)
        s = img.astype(F32)
        sv = np.empty((len(img), 3), F32)
        for k in range(3):
            sv[:, k] = ((s[:, 0] * c[0, k]) + (s[:, 1] * c[1, k])) + (s[:, 2] * c[2, k])
    zero_img = np.flatnonzero((img == 0).all(1))[0]
    rows = []
    # chunk over centre atoms to bound memory: (chunk, n, n_img)
    chunk = max(1, int(4e6 // max(1, n * len(img))))
    for a0 in range(0, n, chunk):
        a1 = min(n, a0 + chunk)
        # xj_img[j, m, k] = x_j[k] + sv[m, k]
        xj = x[None, :, None, :] + sv[None, None, :, :]  # (1, n, n_img, 3)
        r = xj - x[a0:a1, None, None, :]  # (c, n, n_img, 3)
        d2 = ((r[..., 0] * r[..., 0]) + (r[..., 1] * r[..., 1])) + (r[..., 2] * r[..., 2])
        keep = d2 < rc2
        ar = np.arange(a0, a1)
        keep[ar - a0, ar, zero_img] = False
        for a in range(a0, a1):
            jj, mm = np.nonzero(keep[a - a0])
            sh = img[mm]
            order = np.lexsort((sh[:, 2], sh[:, 1], sh[:, 0], jj))
            rows.append((idx[jj[order]], sh[order]))
    return rows


def neighbor_matrix(
    positions: np.ndarray,
    cutoff: float,
    cell: np.ndarray | None = None,
    pbc: np.ndarray | None = None,
    batch_idx: np.ndarray | None = None,
    fill_value: int | None = None,
):
    """Full neighbor matrix in canonical form.

    positions (N,3) float32; cell None | (3,3) | (S,3,3); pbc None | (3,) | (S,3); batch_idx None | (N,) ints.
    Returns nbmat (N, max(1, max_count)) int32, num_neighbors (N,) int32, shifts (N, W, 3) int32 or None.
    """
    pos = np.asarray(positions, dtype=F32)
    N = pos.shape[0]
    if fill_value is None:
        fill_value = N
    if batch_idx is None:
        batch_idx = np.zeros(N, np.int64)
    batch_idx = np.asarray(batch_idx).astype(np.int64)
    if cell is not None:
        cell = np.asarray(cell, dtype=F32)
        if cell.ndim == 2:
            cell = cell[None]
        nsys = cell.shape[0]
        if pbc is None:
            pbc = np.ones((nsys, 3), bool)
        pbc = np.asarray(pbc, dtype=bool)
        if pbc.ndim == 1:
            pbc = np.broadcast_to(pbc[None], (nsys, 3))
    rows: list = [None] * N
    for s in np.unique(batch_idx):
        idx = np.flatnonzero(batch_idx == s)
        c = None if cell is None else cell[min(int(s), cell.shape[0] - 1)]
        p = None if cell is None else pbc[min(int(s), pbc.shape[0] - 1)]
        for a, row in zip(idx, _system_rows(pos, idx, cutoff, c, p)):
            rows[a] = row
    nnb = np.array([len(r[0]) for r in rows], np.int32) if N else np.zeros(0, np.int32)
    W = max(1, int(nnb.max(initial=0)))
    nbmat = np.full((N, W), fill_value, np.int32)
    shifts = np.zeros((N, W, 3), np.int32) if cell is not None else None
    for a, (jj, sh) in enumerate(rows):
        nbmat[a, : len(jj)] = jj
        if shifts is not None:
            shifts[a, : len(jj)] = sh
    return nbmat, nnb, shifts


def wrap_positions(positions: np.ndarray, cell: np.ndarray, pbc: np.ndarray | None = None) -> np.ndarray:
    """move_coord_to_cell (aimnet/calculators/neighbors.py:331-381), single cell, float32 like the reference."""
    import torch

    c = torch.as_tensor(np.asarray(cell, dtype=F32))
    x = torch.as_tensor(np.asarray(positions, dtype=F32))
    f = x @ torch.linalg.inv(c)
    if pbc is None:
        pbc = np.ones(3, bool)
    f = torch.where(torch.as_tensor(np.asarray(pbc, dtype=bool)), f % 1, f)
    return (f @ c).numpy()
