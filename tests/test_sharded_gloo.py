"""Multi-rank host logic on CPU: world_size-2 gloo, a stand-in per-rank compute function (the real one needs a GPU).
Checks shard boundaries, index rebasing and the rank-ordered variable-size gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aimnetcentral_b200.sharded import ShardedCalculator, shard_batch, split_molecules


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
    out = ShardedCalculator(fake_calc)(data, forces=True)
    ok = all(torch.allclose(out[k], full[k]) and out[k].shape == full[k].shape for k in full)
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
    loc, info = shard_batch(data, 1, 2)
    assert info["atoms_per_rank"] == [12, 11] and loc["coord"].shape[0] == 11
    assert loc["mol_idx"].min() == 0 and loc["mol_idx"].max() == 2 and loc["charge"].shape[0] == 3
