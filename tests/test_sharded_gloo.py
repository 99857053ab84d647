"""Multi-rank host logic on CPU: world_size-2 gloo, a stand-in per-rank compute function (the real one needs a GPU).
Checks shard boundaries, index rebasing and the rank-ordered variable-size gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aimnetcentral_b200.sharded import ShardedCalculator, make_plan, shard_batch, split_molecules, split_molecules_balanced


def fake_calc(data, forces=False, stress=False):
    """Deterministic per-molecule / per-atom function of the inputs, same contract as the calculator."""
    coord = torch.as_tensor(np.asarray(data["coord"]), dtype=torch.float32)
    charge = torch.as_tensor(np.asarray(data["charge"]), dtype=torch.float32)
    if coord.ndim == 3:
        energy = coord.double().sum(dim=(1, 2)) + charge.double()
        return {"energy": energy, "forces": coord * 2.0, "charges": coord[..., 0] + charge[:, None]}
    mi = torch.as_tensor(np.asarray(data["mol_idx"]), dtype=torch.long)
    energy = torch.zeros(charge.shape[0], dtype=torch.float64).index_add_(0, mi, coord.double().sum(1)) + charge.double()
    return {"energy": energy, "forces": coord * 2.0, "charges": coord[:, 0] + charge[mi]}


def _worker(rank, world, port, form, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    if form == "dense":
        data = {"coord": rng.normal(size=(5, 7, 3)).astype(np.float32), "numbers": np.ones((5, 7), np.int32),
                "charge": np.arange(5, dtype=np.float32)}
    else:
        sizes = [3, 9, 1, 6, 4]
        data = {"coord": rng.normal(size=(sum(sizes), 3)).astype(np.float32), "numbers": np.ones(sum(sizes), np.int32),
                "charge": np.arange(5, dtype=np.float32), "mol_idx": np.repeat(np.arange(5), sizes)}
    full = fake_calc(data)
    sharded = ShardedCalculator(fake_calc)
    out = sharded(data, forces=True)
    ok = all(torch.equal(out[k], full[k]) and out[k].shape == full[k].shape and out[k].dtype == full[k].dtype for k in full)
    # tensor mol_idx: the shard plan is cached per tensor identity / version and reused on the next call
    if form == "flat":
        tdata = dict(data, mol_idx=torch.as_tensor(data["mol_idx"]))
        out2 = sharded(tdata, forces=True)
        plan_a = sharded._plan_cache[2]
        out3 = sharded(tdata, forces=True)
        ok = ok and sharded._plan_cache[2] is plan_a and all(torch.equal(out2[k], full[k]) and torch.equal(out3[k], full[k]) for k in full)
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("form", ["dense", "flat"])
def test_sharded_gather_world2(form):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, form, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_split_and_shard_bookkeeping():
    assert split_molecules(10, 4) == [(0, 2), (2, 5), (5, 7), (7, 10)]
    sizes = [3, 9, 1, 6, 4]
    data = {"coord": np.zeros((sum(sizes), 3), np.float32), "numbers": np.ones(sum(sizes), np.int32),
            "charge": np.zeros(5, np.float32), "mol_idx": np.repeat(np.arange(5), sizes)}
    # atom-balanced contiguous split: 23 atoms -> boundary at the molecule edge nearest 11.5
    assert split_molecules_balanced(sizes, 2) == [(0, 2), (2, 5)]
    loc, info = shard_batch(data, 1, 2)
    assert info["atoms_per_rank"] == [12, 11] and loc["coord"].shape[0] == 11
    assert loc["mol_idx"].min() == 0 and loc["mol_idx"].max() == 2 and loc["charge"].shape[0] == 3
    # one giant molecule next to small ones: count-balanced would give rank 0 all the atoms
    sizes = [100, 2, 2, 2, 2, 2]
    assert split_molecules_balanced(sizes, 2) == [(0, 1), (1, 6)]
    assert split_molecules_balanced([5] * 8, 4) == split_molecules(8, 4)
    # more ranks than molecules: empty trailing / leading shards are legal
    r = split_molecules_balanced([4, 4], 4)
    assert r[0][0] == 0 and r[-1][1] == 2 and all(a <= b for a, b in r) and all(r[k][1] == r[k + 1][0] for k in range(3))
    plan = make_plan({"coord": np.zeros((6, 5, 3), np.float32), "charge": np.zeros(6, np.float32)}, 3)
    assert plan["form"] == "dense" and plan["atom_ranges"] == [(0, 10), (10, 20), (20, 30)]
