"""GPU tests of the pair-term operator seams, the multi-tile GEMM path and backend 3 (first run on a B200 in round 2,
all green: profiles/r2a_gpu_verify.txt).

* Pair-term operator seams (SURVEY.md section 8b, B3 iii): `ops.dsf_coulomb` and `ops.dftd3` against independent float64
  torch restatements of the reference's closed forms (aimnet/modules/lr.py:559-615 and :1580-1657) on the very neighbor
  matrices the seam receives.
* The MLP GEMM with more output tiles than SMs (the persistent CTA runs several tiles back to back: all of cfg-2).
* Backend 3 (pipelined tile epilogue) bitwise against backend 2."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

HARTREE, BOHR = 27.211386024367243, 0.5291772105638411


def _system(name, cutoff):
    from aimnetcentral_b200 import ops

    inputs, _, _ = load_golden(name)
    x = torch.as_tensor(inputs["coord"], dtype=torch.float32, device="cuda")
    cell = torch.as_tensor(inputs["cell"], dtype=torch.float32, device="cuda").reshape(3, 3)
    z = torch.as_tensor(inputs["numbers"], dtype=torch.int32, device="cuda")
    xw = ops.wrap_positions(x, cell)
    nb, cnt, sh = ops.neighbor_list(xw, cutoff, cell=cell.reshape(1, 3, 3), pbc=torch.ones(1, 3, dtype=torch.bool, device="cuda"),
                                    max_neighbors=int(1.3 * 4.19 * cutoff**3 * len(z) / abs(torch.det(cell).item())) + 32)
    return xw, cell, z, nb, sh


def _pairs(x, cell, nb, sh, strain):
    """r_ij = (x_j + s @ cell - x_i) @ (1 + strain), distances, validity mask (float64, CPU, differentiable)."""
    n = x.shape[0]
    valid = nb < n
    j = nb.clamp(max=n - 1).long()
    r = x[j] + sh.to(x.dtype) @ cell - x.unsqueeze(1)
    r = r @ (torch.eye(3, dtype=x.dtype) + strain)
    d = torch.where(valid, r.norm(dim=-1), torch.ones_like(r[..., 0]))
    return j, d, valid


def test_dsf_coulomb_against_closed_form():
    from aimnetcentral_b200 import ops

    alpha, R = 0.2, 9.0
    xw, cell, z, nb, sh = _system("pbc_box60_dsf", R)
    n = xw.shape[0]
    q = torch.as_tensor(np.random.default_rng(3).normal(0, 0.4, n), dtype=torch.float32, device="cuda").requires_grad_(True)
    e, f, w = ops.dsf_coulomb(positions=xw, charges=q, cutoff=R, alpha=alpha, cell=cell.reshape(1, 3, 3), batch_idx=None,
                              neighbor_matrix=nb, neighbor_matrix_shifts=sh, fill_value=n, compute_forces=True,
                              compute_virial=True, num_systems=1, device="cuda")
    (gq,) = torch.autograd.grad(e.sum(), q)
    # float64 restatement (lr.py:594-611)
    x64 = xw.detach().cpu().double().requires_grad_(True)
    q64 = q.detach().cpu().double().requires_grad_(True)
    eps = torch.zeros(3, 3, dtype=torch.float64, requires_grad=True)
    j, d, valid = _pairs(x64, cell.cpu().double(), nb.cpu(), sh.cpu(), eps)
    a, Rt = torch.tensor(alpha, dtype=torch.float64), torch.tensor(R, dtype=torch.float64)
    shift_val = torch.erfc(a * Rt) / Rt
    slope = torch.erfc(a * Rt) / Rt**2 + 2 * a / np.sqrt(np.pi) * torch.exp(-(a * Rt) ** 2) / Rt
    e_pair = torch.erfc(a * d) / d - shift_val + (d - Rt) * slope
    e_pair = torch.where(valid & (d < Rt), e_pair, torch.zeros_like(e_pair))
    e_ref = 0.5 * (q64.unsqueeze(1) * q64[j] * e_pair).sum() - (shift_val / 2 + a / np.sqrt(np.pi)) * (q64**2).sum()
    gx, gq_ref, geps = torch.autograd.grad(e_ref, (x64, q64, eps))
    print(f"[seam] dsf: dE={abs(e.item() - e_ref.item()):.2e} dF={(f.cpu().double() + gx).abs().max():.2e} "
          f"dgq={(gq.cpu().double() - gq_ref).abs().max():.2e} dW={(w[0].cpu().double() + geps).abs().max():.2e}")
    assert abs(e.item() - e_ref.item()) < 1e-5 * max(1.0, abs(e_ref.item()))
    assert (f.cpu().double() + gx).abs().max() < 2e-5            # forces = -dE/dx
    assert (gq.cpu().double() - gq_ref).abs().max() < 2e-5
    assert (w[0].cpu().double() + geps).abs().max() < 1e-4       # W = -dE/d(strain)


def test_dftd3_against_closed_form():
    from aimnetcentral_b200 import ops
    from oracle.aimnet2_oracle import D3Tables, dftd3_energy
    from oracle.calculator_oracle import d3_tables

    r_on, r_off = 7.5, 9.0      # Angstrom
    xw, cell, z, nb, sh = _system("allose_1x1x1_dsf", r_off)
    n = xw.shape[0]
    tab = d3_tables()
    # the call site's tables: (95,95,5,5) C6 and a (95,95,5,5) reference-CN table constant over the partner (lr.py:1405-1422)
    cn4 = tab.cnref[:, None, :, None].expand(95, 95, 5, 5).contiguous()
    out = ops.dftd3(positions=xw / BOHR, numbers=z, a1=0.566, a2=3.128, s8=0.3908, s6=1.0, covalent_radii=tab.rcov.cuda(),
                    r4r2=tab.r4r2.cuda(), c6_reference=tab.c6ref.cuda(), coord_num_ref=cn4.cuda(), batch_idx=None,
                    cell=cell.reshape(1, 3, 3) / BOHR, neighbor_matrix=nb, neighbor_matrix_shifts=sh, fill_value=n,
                    num_systems=1, compute_virial=True, device="cuda", s5_smoothing_on=r_on / BOHR,
                    s5_smoothing_off=r_off / BOHR)
    e_h, f_hb, cn, w_h = out
    x64 = xw.detach().cpu().double().requires_grad_(True)
    eps = torch.zeros(3, 3, dtype=torch.float64, requires_grad=True)
    j, d, valid = _pairs(x64, cell.cpu().double(), nb.cpu(), sh.cpu(), eps)
    # oracle convention: rows padded with index n, one extra (padding) atom at the end of the per-atom tensors
    nbp = torch.cat([torch.where(valid, j, torch.full_like(j, n)), torch.full((1, nb.shape[1]), n, dtype=torch.long)])
    dp = torch.cat([d, torch.ones(1, d.shape[1], dtype=d.dtype)])
    mask = nbp == n
    zp = torch.cat([z.cpu().long(), torch.zeros(1, dtype=torch.long)])
    tab64 = D3Tables(tab.c6ref.numpy(), tab.cnref.numpy(), tab.rcov.numpy(), tab.r4r2.numpy())
    mol_idx = torch.zeros(n + 1, dtype=torch.long)
    e_ref = dftd3_energy(dp, mask, zp, nbp, mol_idx, 1, tab64, 1.0, 0.3908, 0.566, 3.128, r_on=r_on, r_off=r_off)[0]
    gx, geps = torch.autograd.grad(e_ref, (x64, eps))
    e_ev, f_ev, w_ev = e_h.item() * HARTREE, f_hb.cpu().double() * HARTREE / BOHR, w_h[0].cpu().double() * HARTREE
    print(f"[seam] d3: dE={abs(e_ev - e_ref.item()):.2e} dF={(f_ev + gx).abs().max():.2e} dW={(w_ev + geps).abs().max():.2e}")
    assert abs(e_ev - e_ref.item()) < 1e-5
    assert (f_ev + gx).abs().max() < 2e-5
    assert (w_ev + geps).abs().max() < 1e-4
    assert cn.min() > 0


def _gemm_seam(A, W, b, aux_in, mode, backend):
    """Y, aux of aimnet2_gemm_nt; backend + 16 = the same backend writing its output pre-split (mode | 16)."""
    from aimnetcentral_b200 import _capi

    lib = _capi.load()
    M, K = A.shape
    N = W.shape[0]
    Y = torch.empty(M, N, device=A.device)
    aux = aux_in.clone() if mode == 3 else torch.empty(M, N, device=A.device)
    flag = 16 if backend >= 16 else 0
    rc = lib.aimnet2_gemm_nt(A.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Y.data_ptr(), N, aux.data_ptr(), N, M, N, K,
                             mode | flag, backend & 15, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.aimnet2_last_error()
    torch.cuda.synchronize()
    return Y, aux


def _gemm_operands(M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) * 0.05).cuda()
    b = torch.randn(N, generator=g).cuda()
    aux_in = torch.randn(M, N, generator=g).cuda()
    return A, W, b, aux_in


@pytest.mark.parametrize("split_out", [0, 16])
@pytest.mark.parametrize("M,N,K,mode", [(300, 512, 704, 2), (1000, 288, 384, 1), (77, 736, 512, 0), (513, 384, 512, 3),
                                        (1, 128, 256, 2), (640, 128, 128, 2), (20000, 512, 704, 2), (20000, 704, 512, 3),
                                        (20000, 288, 384, 1), (19999, 256, 512, 0)])
def test_gemm_pipelined_epilogue_equals_backend2_bitwise(M, N, K, mode, split_out):
    """Backend 3 (gemm_tc16p.cu: 128x128 tiles, two accumulator sets, the epilogue of tile t sliced under the K loop of
    tile t+1) performs the same operations in the same order per output element as backend 2, so its results must be
    bit-identical — including short K loops (fewer chunks than epilogue slices), ragged last tiles and CTAs that run
    many tiles.  RUN UNDER `timeout`: the kernel has never executed."""
    A, W, b, aux_in = _gemm_operands(M, N, K)
    Y2, aux2 = _gemm_seam(A, W, b, aux_in, mode, 2 + split_out)
    Y3, aux3 = _gemm_seam(A, W, b, aux_in, mode, 3 + split_out)
    assert torch.equal(Y2, Y3), float((Y2 - Y3).abs().max())
    if mode == 2:
        assert torch.equal(aux2, aux3)


@pytest.mark.parametrize("split_out", [0, 16])
@pytest.mark.parametrize("M,N,K,mode", [(300, 512, 704, 2), (1000, 288, 384, 1), (77, 736, 512, 0), (513, 384, 512, 3),
                                        (1, 128, 256, 2), (640, 128, 128, 2), (51200, 512, 704, 2), (51200, 704, 512, 3),
                                        (40000, 288, 384, 1), (39999, 256, 512, 0), (30000, 96, 96, 2), (51200, 512, 32, 3)])
@pytest.mark.parametrize("backend", [4, 5])
def test_gemm_two_streams_equals_backend2_bitwise(M, N, K, mode, split_out, backend):
    """Backend 4 (gemm_tc16d.cu: two 128x128 tile streams per SM, each with its own operand ring, TMEM half, MMA issuer and
    four epilogue warps) does per output element exactly what backend 2 does, in the same order: bit-identical results,
    including CTAs whose second stream has no tile, short K loops (one chunk), ragged last tiles in M and N, and
    many tiles per stream (51 200 x 704: 2 400 tiles, eight per stream).  Backend 5 (gemm_tc16c.cu) is the same kernel on
    CTA pairs (cta_group::2: 256 x 128 pair tiles, the W tile split between the two CTAs, accumulator rows split between
    their TMEMs); M = 1 / 77 / 300 leave the second CTA of the pair with no or few rows."""
    A, W, b, aux_in = _gemm_operands(M, N, K)
    Y2, aux2 = _gemm_seam(A, W, b, aux_in, mode, 2 + split_out)
    Y4, aux4 = _gemm_seam(A, W, b, aux_in, mode, backend + split_out)
    assert torch.equal(Y2, Y4), float((Y2 - Y4).abs().max())
    if mode == 2:
        assert torch.equal(aux2, aux4)


_MANY_TILES = [(20000, 512, 704, 2), (20000, 288, 384, 1), (20000, 704, 512, 3), (19999, 512, 512, 0)]


@pytest.mark.parametrize("backend", [3, 19])
@pytest.mark.parametrize("M,N,K,mode", _MANY_TILES)
def test_gemm_more_tiles_than_sms_pipelined(M, N, K, mode, backend):
    """The same check for the experimental backend 3 (942 tiles for the 704-wide layer: six or seven tiles per CTA)."""
    test_gemm_more_tiles_than_sms(M, N, K, mode, backend)


@pytest.mark.parametrize("backend", [2, 18, 4, 20, 5, 21])   # +16 = the backend writing its output pre-split (mode | 16)
@pytest.mark.parametrize("M,N,K,mode", _MANY_TILES)
def test_gemm_more_tiles_than_sms(M, N, K, mode, backend):
    """157 row tiles x 2-3 column tiles = 314-471 tiles on 148 persistent CTAs: each CTA runs 2-4 tiles back to back
    (TMEM / box-ring / aux-ring state carried across tiles), checked against float64 like tests/test_gpu_ops.py does for
    the single-wave shapes."""
    A, W, b, aux_in = _gemm_operands(M, N, K)
    Y, aux = _gemm_seam(A, W, b, aux_in, mode, backend)
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
    err = (Y.double() - ref).abs()
    worst = int(err.max(dim=1).values.argmax())
    assert torch.allclose(Y.double(), ref, atol=5e-5, rtol=1e-5), (float(err.max()), "row", worst, "row tile", worst // 128)


def test_rotation_invariance_molecules():
    """Rigid rotation of every molecule of a batch (tests/test_calculator.py:976-1014 of the reference): energies and
    charges unchanged, forces rotate with the frame.  The rotated coordinates are different fp32 numbers (~5e-7 A), so
    the bounds are a few times the parity tolerance."""
    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
    from aimnetcentral_b200.structures import random_molecules

    spec = ModelSpec()
    calc = AIMNet2Calculator((random_state_dict(0, spec), spec), device="cuda:0")
    coord, numbers = random_molecules(96, 50, seed=17)
    charge = np.zeros(96, np.float32)
    charge[::5] = -1.0
    rng = np.random.default_rng(2)
    q = rng.normal(size=(96, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                  np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                  np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)   # (96,3,3)
    rotated = np.einsum("bij,bnj->bni", R, coord.astype(np.float64)).astype(np.float32)
    for rows in (0, 512):
        calc.engine.set_small_m_rows(rows)
        r1 = {k: v.cpu().numpy() for k, v in calc({"coord": coord, "numbers": numbers, "charge": charge}, forces=True).items()}
        r2 = {k: v.cpu().numpy() for k, v in calc({"coord": rotated, "numbers": numbers, "charge": charge}, forces=True).items()}
        f1_rot = np.einsum("bij,bnj->bni", R, r1["forces"].astype(np.float64))
        de = np.abs(r2["energy"] - r1["energy"]).max()
        df = np.abs(r2["forces"] - f1_rot).max()
        dq = np.abs(r2["charges"] - r1["charges"]).max()
        print(f"[invariance] rotation (small_m_rows {rows}): dE {de:.2e} dF {df:.2e} dq {dq:.2e}")
        assert de < 2e-4 and df < 5e-4 and dq < 1e-4


def test_engine_with_pipelined_gemm_equals_backend2():
    """The whole evaluation with the per-atom MLPs on the experimental backend 3: bit-identical to backend 2 (same
    arithmetic per element), on a batch large enough for several tiles per CTA."""
    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
    from aimnetcentral_b200.structures import random_molecules

    spec = ModelSpec()
    calc = AIMNet2Calculator((random_state_dict(0, spec), spec), device="cuda:0")
    coord, numbers = random_molecules(600, 50, seed=21)
    inp = {"coord": coord, "numbers": numbers, "charge": np.zeros(600, np.float32)}
    calc.engine.set_gemm_backend(2)
    ref = {k: v.clone() for k, v in calc(dict(inp), forces=True).items()}
    calc.engine.set_gemm_backend(3)
    try:
        out = calc(dict(inp), forces=True)
        torch.cuda.synchronize()
    finally:
        calc.engine.set_gemm_backend(2)
    for k in ref:
        assert torch.equal(ref[k], out[k]), (k, float((ref[k].double() - out[k].double()).abs().max()))


def test_ewald_summation_seam_against_oracle():
    """`ops.ewald_summation` (call-site signature of aimnet/modules/lr.py:687-696) against the float64 textbook Ewald of the
    oracle (oracle/aimnet2_oracle.py: coulomb_ewald, checked against the rock-salt Madelung constant): per-system energy,
    forces and charge response via autograd on the oracle, for one system and for a batch of two different cells."""
    from aimnetcentral_b200 import ops
    from aimnetcentral_b200.structures import allose_supercell, random_periodic_box
    from oracle.aimnet2_oracle import coulomb_ewald

    ke = HARTREE * BOHR
    systems = []
    z1, x1, c1 = random_periodic_box(60, seed=7)
    z2, x2, c2 = allose_supercell((1, 1, 1), jitter=0.02, seed=3)
    rng = np.random.default_rng(4)
    for x, c in ((x1, c1), (x2, c2)):
        q = rng.normal(0, 0.4, len(x)).astype(np.float32)
        q[0] += 0.3   # net charge: the neutralising background term is exercised
        systems.append((x, c, q))

    def oracle(x, c, q):
        xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
        qt = torch.tensor(q, dtype=torch.float64, requires_grad=True)
        e = coulomb_ewald(xt, qt, torch.tensor(c, dtype=torch.float64), accuracy=1e-6)
        gx, gq = torch.autograd.grad(e, [xt, qt])
        return float(e) / ke, (-gx / ke).numpy(), (gq / ke).numpy()

    def seam(xs, cs, qs):
        dev = "cuda"
        x = torch.as_tensor(np.concatenate(xs), dtype=torch.float32, device=dev)
        q = torch.as_tensor(np.concatenate(qs), dtype=torch.float32, device=dev).requires_grad_(True)
        cells = torch.as_tensor(np.stack(cs), dtype=torch.float32, device=dev)
        bidx = torch.as_tensor(np.concatenate([np.full(len(a), k) for k, a in enumerate(xs)]), dtype=torch.int32, device=dev)
        par = ops.estimate_ewald_parameters(x, cells, bidx, accuracy=1e-6)
        rc = float(par.real_space_cutoff.max())
        xw = torch.cat([ops.wrap_positions(x[bidx == k], cells[k]) for k in range(len(xs))])
        nb, cnt, sh = ops.neighbor_list(xw, rc, cell=cells, pbc=torch.ones(len(xs), 3, dtype=torch.bool, device=dev), batch_idx=bidx,
                                        max_neighbors=int(1.5 * 4.19 * rc**3 * 0.12) + 64)
        e_atom, f = ops.ewald_summation(positions=xw, charges=q, cell=cells, batch_idx=bidx, neighbor_matrix=nb,
                                        neighbor_matrix_shifts=sh, mask_value=xw.shape[0], accuracy=1e-6, compute_forces=True)
        e_sys = torch.zeros(len(xs), dtype=torch.float64, device=dev).scatter_add(0, bidx.long(), e_atom)
        (gq,) = torch.autograd.grad(e_sys.sum(), q)
        return e_sys.detach().cpu().numpy(), f.cpu().numpy(), gq.cpu().numpy(), bidx.cpu().numpy()

    for label, sel in (("one system", [0]), ("batch of two cells", [0, 1])):
        xs, cs, qs = zip(*[systems[k] for k in sel])
        e, f, gq, b = seam(xs, cs, qs)
        for k, s in enumerate(sel):
            eo, fo, gqo = oracle(*systems[s])
            m = b == k
            print(f"[seam] ewald {label} system {s}: dE={abs(e[k] - eo):.2e} (E {eo:.3f} e^2/A) dF={np.abs(f[m] - fo).max():.2e} dgq={np.abs(gq[m] - gqo).max():.2e}")
            assert abs(e[k] - eo) < 2e-5 * max(1.0, abs(eo))
            assert np.abs(f[m] - fo).max() < 2e-5
            assert np.abs(gq[m] - gqo).max() < 2e-5
