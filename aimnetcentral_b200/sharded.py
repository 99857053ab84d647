"""Multi-GPU batch split: independent molecules / periodic replicas are sharded across ranks with no
data-path collective; the only communication is ONE rank-ordered result gather per evaluation (NCCL over NVLink on
GPUs, gloo in the CPU tests).  One process per GPU, launched with torchrun.

The reference has no multi-GPU inference path ("run independent processes per GPU", docs/tutorials/performance.md:
275-288); this module is that advice with the bookkeeping done for the caller:

* contiguous molecule ranges balanced by ATOM count (a rank's work is proportional to its atoms, not its molecules);
* every output of a rank (energy f64, forces / charges / spin charges f32, stress f32) packed into one byte buffer and
  exchanged with a single `all_gather_into_tensor` (payloads are ~1 MB: latency-bound, so one launch, not one per key);
* the shard plan is cached per batch layout (`mol_idx` tensor identity + version, or the dense shape), so a step of an MD /
  screening loop does no host round trip for the split.
"""
from __future__ import annotations

import weakref
from typing import Callable

import numpy as np
import torch
import torch.distributed as dist


def split_molecules(n_mol: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, count-balanced molecule ranges: rank r owns [n_mol*r/world, n_mol*(r+1)/world)."""
    return [(n_mol * r // world, n_mol * (r + 1) // world) for r in range(world)]


def split_molecules_balanced(atom_counts, world: int) -> list[tuple[int, int]]:
    """Contiguous molecule ranges with (nearly) equal ATOM totals: boundary r sits at the molecule edge closest to
    total * r / world.  Equal-size molecules give the count-balanced split."""
    counts = np.asarray(atom_counts, dtype=np.int64)
    n_mol = len(counts)
    edges = np.concatenate([[0], np.cumsum(counts)])
    total = int(edges[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(edges, target, "left"))
        if k > 0 and (k > n_mol or abs(edges[k - 1] - target) <= abs(edges[min(k, n_mol)] - target)):
            k -= 1
        bounds.append(min(max(k, bounds[-1]), n_mol))
    bounds.append(n_mol)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def make_plan(data: dict, world: int) -> dict:
    """Shard plan of a batch in one of the calculator's two batched input forms: dense coord (B,N,3) / numbers (B,N) /
    charge (B,), or flat coord (Ntot,3) + sorted mol_idx + charge (B,)."""
    coord = data["coord"]
    ndim = coord.ndim if isinstance(coord, torch.Tensor) else np.ndim(coord)
    charge = data["charge"]
    n_mol = int(np.shape(charge)[0]) if np.ndim(charge) else 1
    if ndim == 3:
        n_per = int(coord.shape[1])
        ranges = split_molecules(n_mol, world)
        return {"form": "dense", "n_mol": n_mol, "atoms_per_mol": n_per, "mol_ranges": ranges,
                "atom_ranges": [(a * n_per, b * n_per) for a, b in ranges]}
    mol_idx = data["mol_idx"]
    mi = mol_idx.detach().cpu().numpy() if isinstance(mol_idx, torch.Tensor) else np.asarray(mol_idx)
    counts = np.bincount(mi, minlength=n_mol)
    ranges = split_molecules_balanced(counts, world)
    edges = np.concatenate([[0], np.cumsum(counts)])
    return {"form": "flat", "n_mol": n_mol, "mol_ranges": ranges,
            "atom_ranges": [(int(edges[a]), int(edges[b])) for a, b in ranges],
            "atoms_per_rank": [int(edges[b] - edges[a]) for a, b in ranges]}


def shard_batch(data: dict, rank: int, world: int, plan: dict | None = None, cache: dict | None = None) -> tuple[dict, dict]:
    """Slice a batch for `rank` (views for tensors / arrays; only `mol_idx` is rebased).  Returns (local data, plan).
    `cache` (optional, owned by the caller): slices of tensors that do not change from call to call (species, molecule
    index, charges ...) are handed out as the SAME tensor objects while the source tensor is unchanged (identity +
    `_version`), so that the calculator's per-tensor caches (species validation) keep hitting."""
    plan = plan or make_plan(data, world)
    lo, hi = plan["mol_ranges"][rank]
    a0, a1 = plan["atom_ranges"][rank]

    def cut(k, v):
        per_mol = k in ("charge", "mult") or (k == "cell" and np.ndim(v) == 3) or (k == "pbc" and np.ndim(v) == 2)
        if plan["form"] == "dense":
            return v[lo:hi] if (per_mol or k in ("coord", "numbers")) else v
        if k in ("coord", "numbers"):
            return v[a0:a1]
        if k == "mol_idx":
            return v[a0:a1] - lo
        return v[lo:hi] if per_mol else v

    out = {}
    for k, v in data.items():
        if v is None:
            continue
        if cache is not None and k != "coord" and isinstance(v, torch.Tensor):
            key = (id(v), v.data_ptr(), v._version, tuple(v.shape), lo, hi, a0, a1)
            hit = cache.get(k)
            if hit is None or hit[0] != key or hit[1]() is not v:
                hit = (key, weakref.ref(v), cut(k, v))
                cache[k] = hit
            out[k] = hit[2]
        else:
            out[k] = cut(k, v)
    return out, plan


_PER_MOL = ("energy", "stress")


def gather_results(local: dict, plan: dict, group=None) -> dict:
    """Rank-ordered gather of every output with ONE collective: each rank packs its tensors into a byte buffer (padded to
    the largest rank), `all_gather_into_tensor` moves it, and the per-key tensors are rebuilt in rank order.  Every rank
    gets the full batch."""
    world = dist.get_world_size(group)
    keys = sorted(local)
    dense = plan["form"] == "dense"
    rows = {}   # key -> rows per rank
    for k in keys:
        if k in _PER_MOL or dense:
            rows[k] = [b - a for a, b in plan["mol_ranges"]]
        else:
            rows[k] = [b - a for a, b in plan["atom_ranges"]]
    row_bytes = {k: (local[k][0].numel() if local[k].shape[0] else int(np.prod(local[k].shape[1:]))) * local[k].element_size()
                 for k in keys}
    # 16-byte aligned segments so that every dtype can be viewed in place
    seg = {k: (max(rows[k]) * row_bytes[k] + 15) // 16 * 16 for k in keys}
    offs, tot = {}, 0
    for k in keys:
        offs[k] = tot
        tot += seg[k]
    dev = local[keys[0]].device
    send = torch.zeros(tot, dtype=torch.uint8, device=dev)
    for k in keys:
        v = local[k].contiguous()
        n = v.numel() * v.element_size()
        if n:
            send[offs[k]: offs[k] + n] = v.view(-1).view(torch.uint8)
    recv = torch.empty(world * tot, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    out = {}
    for k in keys:
        tail = tuple(local[k].shape[1:])
        parts = []
        for r in range(world):
            n = rows[k][r] * row_bytes[k]
            b0 = r * tot + offs[k]
            parts.append(recv[b0: b0 + n].view(local[k].dtype).view(rows[k][r], *tail))
        out[k] = torch.cat(parts, dim=0)
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
        self._plan_cache = None   # (key, weakref or None, plan)
        self._slice_cache: dict = {}

    def _plan(self, data: dict, world: int) -> dict:
        coord, mi = data["coord"], data.get("mol_idx")
        nd = coord.ndim if isinstance(coord, torch.Tensor) else np.ndim(coord)
        if nd == 3:
            key, ref = ("dense", tuple(coord.shape), world), None
        elif isinstance(mi, torch.Tensor):
            key, ref = ("flat", id(mi), mi.data_ptr(), mi._version, tuple(mi.shape), world), weakref.ref(mi)
        else:
            return make_plan(data, world)   # numpy / list mol_idx: cheap host arithmetic, mutation undetectable
        c = self._plan_cache
        if c is not None and c[0] == key and (c[1] is None or c[1]() is mi):
            return c[2]
        plan = make_plan(data, world)
        self._plan_cache = (key, ref, plan)
        return plan

    def __call__(self, data: dict, forces: bool = False, stress: bool = False, gather: bool = True) -> dict:
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return self.calc(data, forces=forces, stress=stress)
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        local_in, plan = shard_batch(data, rank, world, self._plan(data, world), self._slice_cache)
        local = self.calc(local_in, forces=forces, stress=stress)
        if not gather:
            return local
        return gather_results(local, plan, self.group)
