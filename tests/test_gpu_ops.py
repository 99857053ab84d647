"""GPU parity of the operator seams: neighbor matrix (bit-exact vs the CPU oracle), conv_sv_2d_sp, GEMM epilogues."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _cuda_nb(pos, cutoff, cell=None, pbc=None, batch_idx=None, width=None):
    from aimnetcentral_b200.ops import neighbor_list

    dev = "cuda:0"
    out = neighbor_list(torch.as_tensor(pos, device=dev), cutoff,
                        cell=None if cell is None else torch.as_tensor(cell, device=dev),
                        pbc=None if pbc is None else torch.as_tensor(pbc),
                        batch_idx=None if batch_idx is None else torch.as_tensor(batch_idx, device=dev, dtype=torch.int32),
                        max_neighbors=width)
    return [o.cpu().numpy() for o in out]


def _compare(pos, cutoff, cell=None, pbc=None, batch_idx=None):
    from oracle.nblist_oracle import neighbor_matrix

    nb_o, nnb_o, sh_o = neighbor_matrix(pos, cutoff, cell=cell, pbc=pbc, batch_idx=batch_idx)
    w = nb_o.shape[1]
    out = _cuda_nb(pos, cutoff, cell, pbc, batch_idx, width=w)
    assert np.array_equal(out[1], nnb_o), "neighbor counts differ"
    assert np.array_equal(out[0], nb_o), "neighbor matrix not bit-exact"
    if cell is not None:
        assert np.array_equal(out[2], sh_o), "shifts not bit-exact"


def test_nblist_molecule_batch_bit_exact():
    from aimnetcentral_b200.structures import random_molecules

    coord, _ = random_molecules(16, 50, seed=5)
    _compare(coord.reshape(-1, 3), 5.0, batch_idx=np.repeat(np.arange(16), 50))
    _compare(coord.reshape(-1, 3)[:50], 1e6)  # all pairs


def test_nblist_empty_and_single():
    _compare(np.zeros((1, 3), np.float32), 5.0)
    pos = np.array([[0, 0, 0], [10, 0, 0]], np.float32)
    _compare(pos, 5.0)  # no neighbours at all -> width 1, all fill


def test_nblist_small_triclinic_cell_multi_image():
    from aimnetcentral_b200.structures import random_periodic_box
    from oracle.nblist_oracle import wrap_positions

    z, x, cell = random_periodic_box(60, seed=7)
    x = wrap_positions(x, cell)
    _compare(x, 5.0, cell=cell)
    _compare(x, 15.0, cell=cell)
    _compare(x, 5.0, cell=cell, pbc=np.array([True, True, False]))


def test_nblist_cell_list_builder_bit_exact():
    """>= 512 atoms with a cell -> cell-list builder + per-row canonical sort."""
    from aimnetcentral_b200.structures import allose_supercell
    from oracle.nblist_oracle import wrap_positions

    z, x, cell = allose_supercell((3, 1, 2), jitter=0.02, seed=3)  # 576 atoms
    x = wrap_positions(x, cell)
    _compare(x, 5.0, cell=cell)
    _compare(x, 9.0, cell=cell)
    z, x, cell = allose_supercell((4, 2, 2), jitter=0.02, seed=3)  # 1536 atoms, several bins per axis
    x = wrap_positions(x, cell)
    _compare(x, 5.0, cell=cell)


def test_nblist_overflow_raises():
    from aimnetcentral_b200.ops import NeighborOverflowError, neighbor_list
    from aimnetcentral_b200.structures import random_molecules

    coord, _ = random_molecules(1, 50, seed=5)
    with pytest.raises(NeighborOverflowError):
        neighbor_list(torch.as_tensor(coord[0], device="cuda:0"), 5.0, max_neighbors=4)


def test_adaptive_neighbor_list_wrapper():
    from aimnetcentral_b200.ops import AdaptiveNeighborList
    from aimnetcentral_b200.structures import random_molecules
    from oracle.nblist_oracle import neighbor_matrix

    coord, _ = random_molecules(1, 50, seed=5)
    nl = AdaptiveNeighborList(cutoff=5.0)
    nl.max_neighbors = 16  # force the overflow -> grow path (aimnet/calculators/neighbors.py:127-130)
    nbmat, nnb, shifts = nl(torch.as_tensor(coord[0], device="cuda:0"))
    ref, nnb_ref, _ = neighbor_matrix(coord[0], 5.0)
    assert shifts is None and np.array_equal(nbmat.cpu().numpy(), ref) and nl.max_neighbors >= nnb_ref.max()


def test_wrap_positions_matches_reference_formula():
    from aimnetcentral_b200.ops import wrap_positions
    from aimnetcentral_b200.structures import random_periodic_box
    from oracle.nblist_oracle import wrap_positions as wrap_o

    z, x, cell = random_periodic_box(60, seed=11)
    x = x + np.random.default_rng(0).normal(0, 8.0, x.shape).astype(np.float32)
    w = wrap_positions(torch.as_tensor(x, device="cuda:0"), torch.as_tensor(cell, device="cuda:0")).cpu().numpy()
    ref = wrap_o(x, cell)
    frac = (w - ref) @ np.linalg.inv(cell)
    assert np.abs(frac - np.round(frac)).max() < 1e-4  # equal up to a lattice vector at the wrap boundary
    assert (np.abs(np.round(frac)) > 0).mean() < 0.02
    # atoms already inside the cell are not touched: wrapping is idempotent bit for bit
    w2 = wrap_positions(torch.as_tensor(w, device="cuda:0"), torch.as_tensor(cell, device="cuda:0")).cpu().numpy()
    assert np.array_equal(w2, w)


def test_conv_sv_op_against_reference_einsum():
    """tests/test_conv_sv_2d_sp.py:147-192 of the reference: fwd atol 1e-5/rtol 1e-4, bwd atol 1e-4/rtol 1e-3."""
    from aimnetcentral_b200.ops import conv_sv_2d_sp

    z = np.load(f"{GOLDEN}/conv_sv_op.npz")
    dev = "cuda:0"
    a = torch.tensor(z["a"], device=dev, requires_grad=True)
    g = torch.tensor(z["g"], device=dev, requires_grad=True)
    idx = torch.tensor(z["idx"], device=dev)
    out = conv_sv_2d_sp(a, idx, g)
    assert torch.allclose(out.detach().cpu(), torch.tensor(z["out"]), atol=1e-5, rtol=1e-4)
    ga, gg = torch.autograd.grad(out, [a, g], torch.tensor(z["grad_out"], device=dev))
    assert torch.allclose(ga.cpu(), torch.tensor(z["grad_a"]), atol=1e-4, rtol=1e-3)
    assert torch.allclose(gg.cpu(), torch.tensor(z["grad_g"]), atol=1e-4, rtol=1e-3)
    with pytest.raises(TypeError):
        conv_sv_2d_sp(a.double(), idx, g.double())
    with pytest.raises(ValueError):
        conv_sv_2d_sp(a.cpu(), idx.cpu(), g.cpu())


@pytest.mark.parametrize("backend", [0, 1, 2, 18, 4, 20, 5, 21])   # +16 = the backend writing its output pre-split (mode | 16)
@pytest.mark.parametrize("M,N,K,mode", [(300, 512, 704, 2), (1000, 288, 384, 1), (77, 736, 512, 0), (513, 384, 512, 3),
                                        (1, 128, 256, 2)])
def test_gemm_epilogues(M, N, K, mode, backend):
    from aimnetcentral_b200 import _capi

    lib = _capi.load()
    dev = "cuda:0"
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(dev)
    W = (torch.randn(N, K, generator=g) * 0.05).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    aux_in = torch.randn(M, N, generator=g).to(dev)
    Y = torch.empty(M, N, device=dev)
    aux = aux_in.clone() if mode == 3 else torch.empty(M, N, device=dev)
    flag = 0
    if backend >= 16:
        backend, flag = backend - 16, 16
    rc = lib.aimnet2_gemm_nt(A.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Y.data_ptr(), N, aux.data_ptr(), N, M, N, K,
                             mode | flag, backend, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if backend >= 1 and rc != 0 and "not available" in lib.aimnet2_last_error().decode():
        pytest.skip("tcgen05 backend not built")
    assert rc == 0, lib.aimnet2_last_error()
    z = A.double() @ W.double().T
    if mode in (1, 2):
        z = z + b.double()
    if mode == 2:
        ref = torch.nn.functional.gelu(z)
        zz = z.clone().requires_grad_(True)
        gp = torch.autograd.grad(torch.nn.functional.gelu(zz).sum(), zz)[0]
        assert torch.allclose(aux.double(), gp, atol=2e-5, rtol=1e-5)
    elif mode == 3:
        ref = z * aux_in.double()
    else:
        ref = z
    assert torch.allclose(Y.double(), ref, atol=5e-5, rtol=1e-5), float((Y.double() - ref).abs().max())
