"""Multi-GPU batch split: independent molecules / periodic replicas are sharded across ranks with no
data-path collective; the only communication is the rank-ordered result gather (NCCL over NVLink on GPUs, gloo in
the CPU tests).  One process per GPU, launched with torchrun.

The reference has no multi-GPU inference path ("run independent processes per GPU", docs/tutorials/performance.md:
275-288); this module is that advice with the bookkeeping done for the caller.
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import torch
import torch.distributed as dist


def split_molecules(n_mol: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, balanced molecule ranges: rank r owns [n_mol*r/world, n_mol*(r+1)/world)."""
    return [(n_mol * r // world, n_mol * (r + 1) // world) for r in range(world)]


def shard_batch(data: dict, rank: int, world: int) -> tuple[dict, dict]:
    """Slice a batch for `rank`.  Accepts the calculator's two batched input forms:
    dense coord (B,N,3) / numbers (B,N) / charge (B,), or flat coord (Ntot,3) + sorted mol_idx + charge (B,).
    Returns (local data, bookkeeping for the gather)."""
    coord = np.asarray(data["coord"]) if not isinstance(data["coord"], torch.Tensor) else data["coord"]
    charge = data["charge"]
    n_mol = int(np.shape(charge)[0]) if np.ndim(charge) else 1
    lo, hi = split_molecules(n_mol, world)[rank]
    out = {}
    if coord.ndim == 3:
        for k, v in data.items():
            if v is None:
                continue
            if k in ("coord", "numbers", "charge", "mult") or (k == "cell" and np.ndim(v) == 3):
                out[k] = v[lo:hi]
            else:
                out[k] = v
        info = {"form": "dense", "n_mol": n_mol, "atoms_per_mol": coord.shape[1]}
    else:
        mol_idx = data["mol_idx"]
        mi = mol_idx.cpu().numpy() if isinstance(mol_idx, torch.Tensor) else np.asarray(mol_idx)
        a0, a1 = int(np.searchsorted(mi, lo, "left")), int(np.searchsorted(mi, hi, "left"))
        for k, v in data.items():
            if v is None:
                continue
            if k in ("coord", "numbers"):
                out[k] = v[a0:a1]
            elif k == "mol_idx":
                out[k] = v[a0:a1] - lo
            elif k in ("charge", "mult") or (k == "cell" and np.ndim(v) == 3):
                out[k] = v[lo:hi]
            else:
                out[k] = v
        counts = np.bincount(mi, minlength=n_mol)
        info = {"form": "flat", "n_mol": n_mol,
                "atoms_per_rank": [int(counts[a:b].sum()) for a, b in split_molecules(n_mol, world)]}
    info["mol_ranges"] = split_molecules(n_mol, world)
    return out, info


def _all_gather_var(x: torch.Tensor, sizes: list[int], group=None) -> torch.Tensor:
    """all_gather of tensors whose first dimension differs per rank (padded to the max)."""
    world = dist.get_world_size(group)
    mx = max(sizes)
    pad = torch.zeros((mx, *x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[: x.shape[0]] = x
    if len(set(sizes)) == 1:
        out = torch.empty((world * mx, *x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, pad, group=group)
        return out
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)


def gather_results(local: dict, info: dict, group=None) -> dict:
    """Rank-ordered gather of energy (per molecule) and per-atom outputs; every rank gets the full batch."""
    world = dist.get_world_size(group)
    mol_sizes = [b - a for a, b in info["mol_ranges"]]
    out = {}
    for k, v in local.items():
        if k in ("energy", "stress"):
            out[k] = _all_gather_var(v, mol_sizes, group)
        elif info["form"] == "dense":
            out[k] = _all_gather_var(v, mol_sizes, group)
        else:
            out[k] = _all_gather_var(v, info["atoms_per_rank"], group)
    assert world == len(mol_sizes)
    return out


class ShardedCalculator:
    """Wrap a per-rank calculator (any callable `calc(data, forces=..., stress=...) -> dict of tensors`).

    >>> dist.init_process_group("nccl"); torch.cuda.set_device(local_rank)
    >>> calc = ShardedCalculator(AIMNet2Calculator(model, device=f"cuda:{local_rank}"))
    >>> out = calc(batch, forces=True)        # same on every rank: the whole batch
    """

    def __init__(self, calc: Callable, group=None):
        self.calc = calc
        self.group = group

    def __call__(self, data: dict, forces: bool = False, stress: bool = False, gather: bool = True) -> dict:
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return self.calc(data, forces=forces, stress=stress)
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        local_in, info = shard_batch(data, rank, world)
        local = self.calc(local_in, forces=forces, stress=stress)
        return gather_results(local, info, self.group) if gather else local
