"""Brute-force stand-in for nvalchemiops.torch.neighbors.neighbor_list (call-site signature:
aimnet/calculators/neighbors.py:106-125, aimnet/modules/lr.py:388-396). Delegates to
oracle.nblist_oracle (numpy)."""
import numpy as np
import torch

from nvalchemiops.neighbors import NeighborOverflowError
from oracle.nblist_oracle import neighbor_matrix


def neighbor_list(positions, cutoff, cell=None, pbc=None, batch_idx=None, max_neighbors=None,
                  half_fill=False, fill_value=None, method=None, **_):
    assert not half_fill
    N = positions.shape[0]
    if fill_value is None:
        fill_value = N
    pos = positions.detach().cpu().numpy().astype(np.float32)
    cell_np = None if cell is None else cell.detach().cpu().numpy().astype(np.float32)
    pbc_np = None if pbc is None else np.asarray(pbc.detach().cpu().numpy(), dtype=bool)
    bidx = None if batch_idx is None else batch_idx.detach().cpu().numpy().astype(np.int64)
    nbmat, nnb, shifts = neighbor_matrix(pos, float(cutoff), cell=cell_np, pbc=pbc_np, batch_idx=bidx,
                                         fill_value=int(fill_value))
    width = nbmat.shape[1]
    if max_neighbors is not None:
        if int(nnb.max(initial=0)) > max_neighbors:
            raise NeighborOverflowError(f"max_neighbors={max_neighbors} < {int(nnb.max())}")
        pad = max_neighbors - width
        if pad > 0:
            nbmat = np.concatenate([nbmat, np.full((N, pad), fill_value, np.int32)], 1)
            if shifts is not None:
                shifts = np.concatenate([shifts, np.zeros((N, pad, 3), np.int32)], 1)
        elif pad < 0:
            nbmat = nbmat[:, :max_neighbors]
            if shifts is not None:
                shifts = shifts[:, :max_neighbors]
    dev = positions.device
    out = (torch.from_numpy(nbmat).to(dev), torch.from_numpy(nnb.astype(np.int32)).to(dev))
    if cell is not None:
        out = (*out, torch.from_numpy(shifts).to(dev))
    return out
