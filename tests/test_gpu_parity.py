"""GPU parity: the CUDA path (through the C ABI / AIMNet2Calculator) against the committed golden outputs of the
unmodified reference and against the CPU oracle on the same seeded inputs."""
import os
import warnings

import numpy as np
import pytest
import torch

from conftest import CHARGE_ATOL, ENERGY_ATOL, FORCE_ATOL, golden_state_dict, load_golden

# Hessians are finite differences of the fp32 analytic forces (calculator.py `_eval_hessian`; measured 1-3e-3 eV/A^2 max,
# 2-4e-4 rms on values up to 57 eV/A^2, tools/hessian_step_scan.py).  The bound is the one the reference's own tests use for
# its finite-difference Hessian blocks (tests/test_calculator.py:432-452: 5e-3; test_ase.py:468-482: 1e-3 of |H|max),
# checked against the reference's float64 twin; the reference's own fp32 Hessian is 0.8-1.8e-4 from that twin.
HESSIAN_ATOL = 5.0e-3

pytestmark = pytest.mark.gpu

_CALCS = {}


def get_calc(meta):
    from aimnetcentral_b200 import AIMNet2Calculator

    key = (meta["weights_seed"], meta["num_charge_channels"])
    if key not in _CALCS:
        sd, spec = golden_state_dict(meta)
        _CALCS[key] = AIMNet2Calculator((sd, spec), device="cuda:0")
    return _CALCS[key]


CASES = [
    ("taxol_q0", {}),
    ("taxol_q1", {}),
    ("caffeine", {}),
    ("mols_8x50", {}),
    ("mols_ragged", {}),
    ("pbc_box60_dsf", {"stress": True}),
    ("pbc_slab60_dsf", {}),
    ("allose_1x1x1_dsf", {"stress": True}),
    ("allose_2x1x1_dsf", {"stress": True}),
    ("nse_4x20", {}),
]


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            return next(line.split(":", 1)[1].strip() for line in f if line.startswith("model name"))
    except Exception:
        import platform
        return platform.processor() or platform.machine()


def _report(name, out, ref, n):
    de = np.abs(out["energy"] - ref["energy"]).max()
    df = np.abs(out["forces"] - ref["forces"]).max()
    dq = np.abs(out["charges"] - ref["charges"]).max()
    print(f"[parity] {name}: N={n} max|dE|={de:.3e} eV  max|dF|={df:.3e} eV/A  max|dq|={dq:.3e}")
    return de, df, dq


@pytest.mark.parametrize("mlp", ["tensor-core", "small-m-simt"])
@pytest.mark.parametrize("name,kw", CASES, ids=[c[0] for c in CASES])
def test_calculator_matches_reference_golden(name, kw, mlp):
    """Every fixture through both MLP paths: the tcgen05 3xFP16 GEMMs (forced by switching the small-system shortcut
    off) and the small-M fp32 SIMT kernel that systems of <= 512 atoms take by default."""
    inputs, ref, meta = load_golden(name)
    calc = get_calc(meta)
    calc.engine.set_small_m_rows(0 if mlp == "tensor-core" else 512)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = calc(dict(inputs), forces=True, stress=bool(kw.get("stress")))
    finally:
        calc.engine.set_small_m_rows(512)
    out = {k: v.detach().cpu().numpy() for k, v in res.items()}
    assert out["energy"].dtype == np.float64 and out["forces"].dtype == np.float32
    de, df, dq = _report(name, out, ref, len(inputs["numbers"]))
    assert de < ENERGY_ATOL, f"energy off by {de}"
    assert df < FORCE_ATOL, f"forces off by {df}"
    assert dq < CHARGE_ATOL, f"charges off by {dq}"
    if "spin_charges" in ref:
        assert np.abs(out["spin_charges"] - ref["spin_charges"]).max() < CHARGE_ATOL
    if kw.get("stress"):
        ds = np.abs(out["stress"] - ref["stress"]).max()
        print(f"[parity] {name}: max|dstress|={ds:.3e} eV/A^3")
        assert ds < 1e-5


@pytest.mark.parametrize("name,kw", [("mols_8x50", {}), ("mols_ragged", {}), ("taxol_q1", {}), ("pbc_box60_dsf", {"stress": True}),
                                     ("nse_4x20", {})], ids=lambda v: v if isinstance(v, str) else "")
def test_conv_list_and_dense_walks_agree(name, kw):
    """conv.cu (matrix rows, neighbour rows gathered through L1 / L2) vs conv_dense.cu (the molecule's tables staged into
    shared memory with TMA, every centre walks all atoms of its molecule): the sums run over the same neighbours in the same
    order (pairs beyond the cutoff add exactly zero), so the two must agree to fp32 round-off.  The dense walk must be what
    batches of small molecules take, and must not be taken with a cell."""
    inputs, ref, meta = load_golden(name)
    calc = get_calc(meta)
    res = {}
    try:
        calc.engine.set_dense_min_molecules(1)   # the fixtures are small: take the dense kernels whatever the batch size
        for impl in (0, 1, 2):
            calc.engine.set_conv_impl(impl)
            for rows in (0, 512):   # both MLP paths
                calc.engine.set_small_m_rows(rows)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    out = calc(dict(inputs), forces=True, stress=bool(kw.get("stress")))
                res[impl, rows] = {k: v.cpu().numpy() for k, v in out.items()}
            mode = calc.engine.conv_mode()
            assert mode["impl"] == impl
            assert mode["dense_last"] == (impl >= 1 and "cell" not in inputs), mode
    finally:
        calc.engine.set_conv_impl(1)
        calc.engine.set_dense_min_molecules(64)
        calc.engine.set_small_m_rows(512)
    for rows, impl in ((0, 1), (512, 1), (0, 2), (512, 2)):
        base, r = res[0, rows], res[impl, rows]
        df = np.abs(r["forces"] - base["forces"]).max()
        print(f"[conv_dense] {name} impl {impl} small_m_rows {rows}: bitwise E {np.array_equal(r['energy'], base['energy'])} "
              f"bitwise q {np.array_equal(r['charges'], base['charges'])} max|dF| vs list kernels {df:.2e}")
        # same neighbours, same order, same formulas; the two kernels are compiled separately, so fused-multiply-add contraction
        # of the pair geometry may differ in the last bit (measured: 3e-7 relative on the conv outputs)
        assert np.abs(r["energy"] - base["energy"]).max() < 2e-5 and np.abs(r["charges"] - base["charges"]).max() < 5e-6
        assert df < 5e-5
        assert np.abs(r["forces"] - ref["forces"]).max() < FORCE_ATOL
        if kw.get("stress"):
            assert np.abs(r["stress"] - ref["stress"]).max() < 1e-5


def test_components_against_reference():
    """NN-only and NN+Coulomb (no D3) goldens isolate the terms."""
    from aimnetcentral_b200 import AIMNet2Calculator

    inputs, ref, meta = load_golden("taxol_q1")
    sd, spec = golden_state_dict(meta)
    nn_only = AIMNet2Calculator((sd, spec), device="cuda:0", needs_coulomb=False, needs_dispersion=False)
    out = {k: v.cpu().numpy() for k, v in nn_only(dict(inputs), forces=True).items()}
    print("[parity] nn-only dE", abs(out["energy"][0] - ref["energy_nn"][0]), "dF", np.abs(out["forces"] - ref["forces_nn"]).max())
    assert abs(out["energy"][0] - ref["energy_nn"][0]) < ENERGY_ATOL
    assert np.abs(out["forces"] - ref["forces_nn"]).max() < FORCE_ATOL
    nod3 = AIMNet2Calculator((sd, spec), device="cuda:0", needs_dispersion=False)
    out = {k: v.cpu().numpy() for k, v in nod3(dict(inputs), forces=True).items()}
    print("[parity] no-d3 dE", abs(out["energy"][0] - ref["energy_nod3"][0]), "dF", np.abs(out["forces"] - ref["forces_nod3"]).max())
    assert abs(out["energy"][0] - ref["energy_nod3"][0]) < ENERGY_ATOL
    assert np.abs(out["forces"] - ref["forces_nod3"]).max() < FORCE_ATOL


def test_dense_batch_equals_flat():
    """(B,N,3) input == flat + mol_idx input (tests/test_calculator.py:1017-1218 of the reference)."""
    inputs, ref, meta = load_golden("mols_8x50")
    calc = get_calc(meta)
    dense = {"coord": inputs["coord"].reshape(8, 50, 3), "numbers": inputs["numbers"].reshape(8, 50),
             "charge": inputs["charge"]}
    res = calc(dense, forces=True)
    assert res["forces"].shape == (8, 50, 3) and res["charges"].shape == (8, 50) and res["energy"].shape == (8,)
    assert np.abs(res["forces"].cpu().numpy().reshape(-1, 3) - ref["forces"]).max() < FORCE_ATOL
    assert np.abs(res["energy"].cpu().numpy() - ref["energy"]).max() < ENERGY_ATOL


def test_padded_dense_batch():
    """numbers == 0 padding in a dense batch: padded atoms get zero outputs, real atoms unchanged."""
    inputs, ref, meta = load_golden("mols_ragged")
    calc = get_calc(meta)
    mi = inputs["mol_idx"]
    B, nmax = int(mi.max()) + 1, int(np.bincount(mi).max())
    coord = np.zeros((B, nmax, 3), np.float32)
    numbers = np.zeros((B, nmax), np.int32)
    for b in range(B):
        sel = mi == b
        coord[b, : sel.sum()] = inputs["coord"][sel]
        numbers[b, : sel.sum()] = inputs["numbers"][sel]
    res = calc({"coord": coord, "numbers": numbers, "charge": inputs["charge"]}, forces=True)
    f = res["forces"].cpu().numpy()
    flat = np.concatenate([f[b, : (mi == b).sum()] for b in range(B)])
    assert np.abs(flat - ref["forces"]).max() < FORCE_ATOL
    assert np.abs(res["energy"].cpu().numpy() - ref["energy"]).max() < ENERGY_ATOL
    assert float(np.abs(f[numbers == 0]).max()) == 0.0


def test_oracle_parity_random_batch():
    """Seeded batch bigger than the fixtures: CUDA vs the CPU oracle run here (64 x 50 atoms)."""
    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
    from aimnetcentral_b200.structures import random_molecules
    from oracle.calculator_oracle import oracle_calculate

    spec = ModelSpec()
    sd = random_state_dict(0, spec)
    coord, numbers = random_molecules(64, 50, seed=99)
    inp = {"coord": coord, "numbers": numbers, "charge": np.zeros(64, np.float32)}
    ref = oracle_calculate(sd, inp)
    calc = AIMNet2Calculator((sd, spec), device="cuda:0")
    res = calc(inp, forces=True)
    df = np.abs(res["forces"].cpu().numpy() - ref["forces"]).max()
    de = np.abs(res["energy"].cpu().numpy() - ref["energy"]).max()
    dq = np.abs(res["charges"].cpu().numpy() - ref["charges"]).max()
    print(f"[parity] random 64x50: max|dE|={de:.3e} max|dF|={df:.3e} max|dq|={dq:.3e}")
    if not (de < ENERGY_ATOL and df < FORCE_ATOL and dq < CHARGE_ATOL):   # localise a failure before reporting it
        dE = np.abs(res["energy"].cpu().numpy() - ref["energy"])
        dQ = np.abs(res["charges"].cpu().numpy() - ref["charges"]).reshape(-1)
        dF = np.abs(res["forces"].cpu().numpy() - ref["forces"]).reshape(-1, 3).max(axis=1)
        res2 = calc(inp, forces=True)
        again = float((res2["forces"] - res["forces"]).abs().max())
        print(f"[parity] molecules off: {np.nonzero(dE > 1e-5)[0].tolist()}  atoms with dq > 1e-5: {np.nonzero(dQ > 1e-5)[0].tolist()[:40]}"
              f" (n={int((dQ > 1e-5).sum())})  atoms with dF > 5e-5: n={int((dF > 5e-5).sum())} first {np.nonzero(dF > 5e-5)[0].tolist()[:20]}"
              f"  second evaluation differs from the first by {again:.3e}")
        # which side moved?  arbitrate with the oracle in float64 and a second run of the float32 oracle
        ref2 = oracle_calculate(sd, inp)
        ref64 = oracle_calculate(sd, inp, dtype=torch.float64)
        print(f"[parity] fp32 oracle rerun differs by {np.abs(ref2['energy'] - ref['energy']).max():.3e};"
              f" vs fp64 oracle: cuda max|dE|={np.abs(res['energy'].cpu().numpy() - ref64['energy']).max():.3e},"
              f" fp32 oracle max|dE|={np.abs(ref['energy'] - ref64['energy']).max():.3e};"
              f" host {_cpu_model()}, torch threads {torch.get_num_threads()},"
              f" device {torch.cuda.get_device_name(0)}")
    assert de < ENERGY_ATOL and df < FORCE_ATOL and dq < CHARGE_ATOL


def test_invariances_full_size():
    """cfg-2 sized batch (1024 x 50): size-independent properties — total force on each isolated molecule vanishes,
    rigid translation + permutation of molecules leaves per-molecule results unchanged, charges sum to the target."""
    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
    from aimnetcentral_b200.structures import random_molecules

    spec = ModelSpec()
    sd = random_state_dict(0, spec)
    calc = AIMNet2Calculator((sd, spec), device="cuda:0")
    coord, numbers = random_molecules(1024, 50, seed=1234)
    charge = np.zeros(1024, np.float32)
    charge[::7] = 1.0
    r1 = calc({"coord": coord, "numbers": numbers, "charge": charge}, forces=True)
    f1 = r1["forces"].cpu().numpy()
    assert np.abs(f1.sum(axis=1)).max() < 2e-4
    assert np.abs(r1["charges"].cpu().numpy().sum(axis=1) - charge).max() < 1e-4
    perm = np.random.default_rng(0).permutation(1024)
    shift = np.random.default_rng(1).normal(0, 3.0, (1024, 1, 3)).astype(np.float32)
    r2 = calc({"coord": (coord + shift)[perm], "numbers": numbers[perm], "charge": charge[perm]}, forces=True)
    assert np.abs(r2["energy"].cpu().numpy() - r1["energy"].cpu().numpy()[perm]).max() < 2e-4
    assert np.abs(r2["forces"].cpu().numpy() - f1[perm]).max() < 2e-4


def test_periodic_invariances_full_size():
    """cfg-3 sized periodic system (7x3x5 allose supercell, 10 080 atoms, E+F+stress): size-independent properties of
    the periodic path — rigid translation (atoms wrap through different faces, Ewald phases all change), atom
    permutation, vanishing net force, and for DSF a doubled cell (2x1x1 images of the same jittered box: energy doubles,
    forces and stress repeat).  The moved / replicated coordinates are different fp32 numbers (4e-6 A at 60 A), so
    those bounds sit above the north-star tolerance; the permutation is held to it."""
    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
    from aimnetcentral_b200.structures import allose_supercell

    spec = ModelSpec()
    calc = AIMNet2Calculator((random_state_dict(0, spec), spec), device="cuda:0")
    z, x, cell = allose_supercell((7, 3, 5), jitter=0.02, seed=3)
    N = len(z)
    assert N == 10080

    def run(zz, xx, cc):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = calc({"coord": xx, "numbers": zz, "charge": np.zeros(1, np.float32), "cell": cc}, forces=True, stress=True)
        return {k: v.cpu().numpy() for k, v in out.items()}

    perm = np.random.default_rng(5).permutation(N)
    shift = np.array([3.217, -1.04, 7.9], np.float32)
    for method in ("dsf", "ewald"):
        calc.set_lrcoulomb_method(method)
        r1 = run(z, x, cell)
        assert np.isfinite(r1["forces"]).all() and np.isfinite(r1["stress"]).all()
        net = np.abs(r1["forces"].astype(np.float64).sum(axis=0)).max()
        r2 = run(z, x + shift, cell)
        r3 = run(z[perm], x[perm], cell)
        d = {"translate": (abs(r2["energy"][0] - r1["energy"][0]) / N, np.abs(r2["forces"] - r1["forces"]).max(),
                           np.abs(r2["stress"] - r1["stress"]).max(), np.abs(r2["charges"] - r1["charges"]).max()),
             "permute": (abs(r3["energy"][0] - r1["energy"][0]) / N, np.abs(r3["forces"] - r1["forces"][perm]).max(),
                         np.abs(r3["stress"] - r1["stress"]).max(), np.abs(r3["charges"] - r1["charges"][perm]).max())}
        if method == "dsf":
            x2 = np.concatenate([x.astype(np.float64), x.astype(np.float64) + cell[0].astype(np.float64)]).astype(np.float32)
            cell2 = cell.copy()
            cell2[0] *= 2.0
            r4 = run(np.concatenate([z, z]), x2, cell2)
            f2 = np.concatenate([r1["forces"], r1["forces"]])
            d["double"] = (abs(r4["energy"][0] - 2.0 * r1["energy"][0]) / (2 * N), np.abs(r4["forces"] - f2).max(),
                           np.abs(r4["stress"] - r1["stress"]).max(),
                           np.abs(r4["charges"] - np.concatenate([r1["charges"], r1["charges"]])).max())
        print(f"[invariance] {method}: net force {net:.2e} eV/A; " +
              "; ".join(f"{k}: dE/N {v[0]:.1e} dF {v[1]:.1e} dstress {v[2]:.1e} dq {v[3]:.1e}" for k, v in d.items()))
        assert net < 2e-2
        # (dE/N, dF, dstress, dq).  Translated / replicated inputs are different fp32 numbers (x + shift rounds at 4e-6 A
        # in a 60 A box) and the random-weight network turns that into ~1e-4 eV/A.  Measured with the wrap that leaves in-cell
        # atoms untouched (round 2): DSF 6.6e-5, Ewald 2.0e-4, doubled cell 1.4e-4 (with the reference's fractional round
        # trip in the wrap, round 1: 2.0e-4 / 6.8e-4 / 2.3e-4); a permutation keeps the numbers and only changes summation
        # orders (measured 6.5e-6 / 2.6e-5)
        tol = {"translate": (1e-6, 5e-4, 5e-6, 1e-4), "permute": (1e-6, 1e-4, 1e-6, 1e-5), "double": (1e-6, 5e-4, 5e-6, 1e-4)}
        for k, v in d.items():
            assert all(a < b for a, b in zip(v, tol[k])), (method, k, v)


def test_host_buffer_entry_matches_device_entry():
    """aimnet2_engine_eval_host (H2D + compute + D2H inside the C call) == device-resident call."""
    inputs, ref, meta = load_golden("mols_8x50")
    calc = get_calc(meta)
    calc.engine.set_options(coulomb_method="simple", dispersion=True)
    out = calc.engine.eval_host(inputs["coord"], inputs["numbers"].astype(np.int32), inputs["charge"],
                                mol_idx=inputs["mol_idx"].astype(np.int32), forces=True)
    assert np.abs(out["forces"] - ref["forces"]).max() < FORCE_ATOL
    assert np.abs(out["energy"] - ref["energy"]).max() < ENERGY_ATOL
    assert calc.engine.last_launches() > 20


def test_ewald_against_oracle():
    """Row a16: Ewald Coulomb (real-space pair walker + reciprocal structure factors) vs the fp32 oracle with its
    fp64 textbook Ewald.  Upstream parity is unpinned (third-party kernel); tolerance = the north-star 1e-4."""
    from aimnetcentral_b200 import AIMNet2Calculator
    from oracle.calculator_oracle import oracle_calculate

    for name in ("allose_1x1x1_dsf", "pbc_box60_dsf"):
        inputs, _, meta = load_golden(name)
        sd, spec = golden_state_dict(meta)
        inputs = dict(inputs)
        inputs["charge"] = np.array([1.0], np.float32) if name.startswith("pbc") else inputs["charge"]  # background term
        ref = oracle_calculate(sd, inputs, coulomb="ewald", stress=True)
        calc = AIMNet2Calculator((sd, spec), device="cuda:0")
        calc.set_lrcoulomb_method("ewald", ewald_accuracy=1e-6)
        out = {k: v.cpu().numpy() for k, v in calc(inputs, forces=True, stress=True).items()}
        de = abs(out["energy"][0] - ref["energy"][0])
        df = np.abs(out["forces"] - ref["forces"]).max()
        ds = np.abs(out["stress"] - ref["stress"]).max()
        dq = np.abs(out["charges"] - ref["charges"]).max()
        print(f"[parity] ewald {name}: dE={de:.3e} dF={df:.3e} dstress={ds:.3e} dq={dq:.3e}")
        assert de < ENERGY_ATOL and df < FORCE_ATOL and dq < CHARGE_ATOL and ds < 1e-5


def test_periodic_components_against_oracle():
    """Periodic systems with terms switched off (the goldens only isolate terms for an isolated molecule): NN only, NN +
    DSF without D3, NN + D3 without Coulomb, each with stress, against the oracle run the same way."""
    from aimnetcentral_b200 import AIMNet2Calculator
    from oracle.calculator_oracle import oracle_calculate

    variants = [("nn only", dict(needs_coulomb=False, needs_dispersion=False), dict(coulomb=None, dispersion=False)),
                ("nn + dsf", dict(needs_dispersion=False), dict(coulomb="dsf", dispersion=False)),
                ("nn + d3", dict(needs_coulomb=False), dict(coulomb=None, dispersion=True))]
    for name in ("allose_1x1x1_dsf", "pbc_box60_dsf"):
        inputs, _, meta = load_golden(name)
        sd, spec = golden_state_dict(meta)
        for label, ckw, okw in variants:
            ref = oracle_calculate(sd, dict(inputs), stress=True, **okw)
            calc = AIMNet2Calculator((sd, spec), device="cuda:0", **ckw)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                out = {k: v.cpu().numpy() for k, v in calc(dict(inputs), forces=True, stress=True).items()}
            de = abs(out["energy"][0] - ref["energy"][0])
            df = np.abs(out["forces"] - ref["forces"]).max()
            ds = np.abs(out["stress"] - ref["stress"]).max()
            fmax = np.abs(ref["forces"]).max()
            print(f"[parity] {name} / {label}: dE={de:.3e} dF={df:.3e} dstress={ds:.3e} max|F|={fmax:.1f}")
            # without the Coulomb term the random-weight model pushes the random box's atoms with up to 19 eV/A, and
            # the fp32 oracle itself is 7.6e-5 eV/A away from its float64 twin there: absolute 1e-4, or 1e-5 relative
            # (the reference's own CPU<->GPU force check is rtol 1e-4)
            assert de < ENERGY_ATOL and df < max(FORCE_ATOL, 1e-5 * fmax) and ds < 1e-5, (name, label)


def test_batched_ewald_equals_individual_systems():
    """Ewald with batch_idx (aimnet/modules/lr.py:687-696): two different periodic systems in one call, each with its own
    cell, splitting parameters and k vectors, reproduce the two single-system evaluations (E, F, stress, charges)."""
    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
    from aimnetcentral_b200.structures import allose_supercell, random_periodic_box

    spec = ModelSpec()
    calc = AIMNet2Calculator((random_state_dict(0, spec), spec), device="cuda:0")
    calc.set_lrcoulomb_method("ewald", ewald_accuracy=1e-6)
    z1, x1, c1 = random_periodic_box(60, seed=7)
    z2, x2, c2 = allose_supercell((1, 1, 1), jitter=0.02, seed=3)
    singles = []
    for z, x, c, qq in ((z1, x1, c1, 1.0), (z2, x2, c2, 0.0)):
        out = calc({"coord": x, "numbers": z, "charge": np.array([qq], np.float32), "cell": c}, forces=True, stress=True)
        singles.append({k: v.cpu().numpy() for k, v in out.items()})
    both = calc({"coord": np.concatenate([x1, x2]), "numbers": np.concatenate([z1, z2]), "charge": np.array([1.0, 0.0], np.float32),
                 "mol_idx": np.concatenate([np.zeros(len(z1)), np.ones(len(z2))]).astype(np.int32), "cell": np.stack([c1, c2])},
                forces=True, stress=True)
    both = {k: v.cpu().numpy() for k, v in both.items()}
    n1 = len(z1)
    de = max(abs(both["energy"][0] - singles[0]["energy"][0]), abs(both["energy"][1] - singles[1]["energy"][0]))
    df = max(np.abs(both["forces"][:n1] - singles[0]["forces"]).max(), np.abs(both["forces"][n1:] - singles[1]["forces"]).max())
    ds = max(np.abs(both["stress"][0] - singles[0]["stress"]).max(), np.abs(both["stress"][1] - singles[1]["stress"]).max())
    dq = max(np.abs(both["charges"][:n1] - singles[0]["charges"]).max(), np.abs(both["charges"][n1:] - singles[1]["charges"]).max())
    print(f"[parity] batched Ewald vs individual: dE {de:.2e} dF {df:.2e} dstress {ds:.2e} dq {dq:.2e}")
    assert de < 2e-5 and df < 2e-5 and ds < 1e-6 and dq < 1e-6
    with pytest.raises(ValueError, match="one cell per system"):
        calc({"coord": np.concatenate([x1, x2]), "numbers": np.concatenate([z1, z2]), "charge": np.zeros(2, np.float32),
              "mol_idx": np.concatenate([np.zeros(len(z1)), np.ones(len(z2))]).astype(np.int32), "cell": c1}, forces=True)


def test_errors_and_warnings():
    inputs, ref, meta = load_golden("caffeine")
    calc = get_calc(meta)
    with pytest.raises(KeyError):
        calc({"coord": inputs["coord"], "numbers": inputs["numbers"]})
    bad = dict(inputs)
    bad["numbers"] = inputs["numbers"].copy()
    bad["numbers"][0] = 26
    with pytest.raises(ValueError):
        calc(bad)
    with pytest.raises(NotImplementedError):
        calc.hessian_vector_product({**inputs, "coord": np.stack([inputs["coord"]] * 2), "numbers": np.stack([inputs["numbers"]] * 2),
                                     "charge": np.zeros(2, np.float32)}, np.zeros((24, 3), np.float32))
    with pytest.raises(ValueError):
        calc.hessian_vector_product(dict(inputs), np.zeros((5, 3), np.float32))
    inputs_p, _, _ = load_golden("pbc_box60_dsf")
    with pytest.warns(UserWarning, match="Switching to DSF"):
        calc(dict(inputs_p), forces=True)
    assert calc.coulomb_method == "simple"  # the auto-switch is scoped to one evaluation
    calc.set_lrcoulomb_method("ewald")
    with pytest.raises(ValueError):
        calc(dict(inputs))
    calc.set_lrcoulomb_method("simple")


def test_ase_adapter_with_stand_in_atoms(monkeypatch):
    """AIMNet2ASE marshalling (aimnet/calculators/aimnet2ase.py:228-274).  ASE is not in this image, so a minimal
    stand-in for ase.calculators.calculator.Calculator / Atoms is injected; the adapter code is unchanged."""
    import sys
    import types

    class _Calc:
        def __init__(self, *a, **k):
            self.results, self.atoms = {}, None

        def reset(self):
            self.results = {}

        def check_state(self, atoms, tol=1e-15):
            return []

        def calculate(self, atoms=None, properties=None, system_changes=None):
            if atoms is not None:
                self.atoms = atoms

        def get_charges(self):
            return self.results["charges"]

    ase = types.ModuleType("ase")
    calcs = types.ModuleType("ase.calculators")
    calc_mod = types.ModuleType("ase.calculators.calculator")
    calc_mod.Calculator, calc_mod.PropertyNotImplementedError, calc_mod.all_changes = _Calc, RuntimeError, ["positions"]
    for name, mod in (("ase", ase), ("ase.calculators", calcs), ("ase.calculators.calculator", calc_mod)):
        monkeypatch.setitem(sys.modules, name, mod)
    sys.modules.pop("aimnetcentral_b200.aimnet2ase", None)
    from aimnetcentral_b200.aimnet2ase import AIMNet2ASE

    class Atoms:
        def __init__(self, numbers, positions, cell=None, pbc=False, info=None):
            self.numbers, self.positions = np.asarray(numbers), np.asarray(positions, dtype=np.float64)
            self.cell = None if cell is None else np.asarray(cell, dtype=np.float64)
            self.pbc = np.array([pbc] * 3) if np.isscalar(pbc) else np.asarray(pbc)
            self.info = info or {}

        def get_positions(self):
            return self.positions

    inputs, ref, meta = load_golden("taxol_q1")
    calc = get_calc(meta)
    ase_calc = AIMNet2ASE(calc, charge=0)
    atoms = Atoms(inputs["numbers"], inputs["coord"], info={"charge": 1})
    ase_calc.calculate(atoms, properties=["energy", "forces"])
    assert abs(ase_calc.results["energy"] - ref["energy"][0]) < ENERGY_ATOL
    assert np.abs(ase_calc.results["forces"] - ref["forces"]).max() < FORCE_ATOL
    assert ase_calc.results["charges"].shape == (len(inputs["numbers"]),)
    inputs, ref, meta = load_golden("allose_1x1x1_dsf")
    atoms = Atoms(inputs["numbers"], inputs["coord"], cell=inputs["cell"], pbc=True, info={"charge": 0})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ase_calc.calculate(atoms, properties=["energy", "forces", "stress"])
    assert np.abs(ase_calc.results["forces"] - ref["forces"]).max() < FORCE_ATOL
    assert np.abs(ase_calc.results["stress"] - ref["stress"]).max() < 1e-5
    sys.modules.pop("aimnetcentral_b200.aimnet2ase", None)


def test_deterministic_mode_is_bitwise_reproducible():
    """deterministic=True: repeated identical evaluations agree bit for bit (tests/test_calculator_gpu.py:620-636
    of the reference).  All kernels are atomics-free; the flag pins the GEMM's K-chunking."""
    from aimnetcentral_b200 import AIMNet2Calculator

    inputs, ref, meta = load_golden("mols_8x50")
    sd, spec = golden_state_dict(meta)
    calc = AIMNet2Calculator((sd, spec), device="cuda:0", deterministic=True)
    a = calc(dict(inputs), forces=True)
    b = calc(dict(inputs), forces=True)
    for k in ("energy", "forces", "charges"):
        assert (a[k] == b[k]).all(), k
    assert np.abs(a["forces"].cpu().numpy() - ref["forces"]).max() < FORCE_ATOL
    calc.engine.set_deterministic(False)


def test_single_atom_and_tiny_systems():
    """Edge cases: one atom (no neighbours at all), two atoms beyond the cutoff, a diatomic."""
    from oracle.calculator_oracle import oracle_calculate

    inputs, _, meta = load_golden("caffeine")
    sd, spec = golden_state_dict(meta)
    calc = get_calc(meta)
    for coord, numbers in (([[0.0, 0.0, 0.0]], [8]), ([[0.0, 0.0, 0.0], [9.0, 0.0, 0.0]], [6, 1]),
                           ([[0.0, 0.0, 0.0], [1.1, 0.0, 0.0]], [7, 7])):
        inp = {"coord": np.array(coord, np.float32), "numbers": np.array(numbers, np.int32), "charge": np.array([0.0], np.float32)}
        ref = oracle_calculate(sd, inp)
        out = {k: v.cpu().numpy() for k, v in calc(inp, forces=True).items()}
        assert abs(out["energy"][0] - ref["energy"][0]) < ENERGY_ATOL
        assert np.abs(out["forces"] - ref["forces"]).max() < FORCE_ATOL
        assert np.abs(out["charges"] - ref["charges"]).max() < CHARGE_ATOL


def test_batched_periodic_cells_equal_individual():
    """Flat coordinates + mol_idx + per-system cells (B,3,3): batch == individual evaluations, incl. per-system stress
    (tests/test_pbc.py:551-745 of the reference)."""
    from aimnetcentral_b200.structures import random_periodic_box

    inputs, _, meta = load_golden("caffeine")
    calc = get_calc(meta)
    systems = [random_periodic_box(40, seed=21), random_periodic_box(56, seed=22, triclinic=False)]
    singles = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for z, x, cell in systems:
            singles.append({k: v.cpu().numpy() for k, v in calc({"coord": x, "numbers": z, "charge": 0.0, "cell": cell},
                                                                 forces=True, stress=True).items()})
        batch = {"coord": np.concatenate([s[1] for s in systems]), "numbers": np.concatenate([s[0] for s in systems]),
                 "charge": np.zeros(2, np.float32), "mol_idx": np.repeat([0, 1], [40, 56]).astype(np.int64),
                 "cell": np.stack([s[2] for s in systems])}
        out = {k: v.cpu().numpy() for k, v in calc(batch, forces=True, stress=True).items()}
    assert out["stress"].shape == (2, 3, 3) and out["energy"].shape == (2,)
    f = np.concatenate([s["forces"] for s in singles])
    assert np.abs(out["forces"] - f).max() < 2e-5
    for b in range(2):
        assert abs(out["energy"][b] - singles[b]["energy"][0]) < 2e-5
        assert np.abs(out["stress"][b] - singles[b]["stress"]).max() < 1e-6


def test_large_nonperiodic_molecule_vs_oracle():
    """One 300-atom non-periodic system: naive all-pairs builder path, long rows, Coulomb over the whole molecule."""
    from aimnetcentral_b200.structures import random_molecules
    from oracle.calculator_oracle import oracle_calculate

    inputs, _, meta = load_golden("caffeine")
    sd, spec = golden_state_dict(meta)
    calc = get_calc(meta)
    coord, numbers = random_molecules(1, 300, seed=77, box=16.0)
    inp = {"coord": coord[0], "numbers": numbers[0], "charge": np.array([-1.0], np.float32)}
    ref = oracle_calculate(sd, inp)
    out = {k: v.cpu().numpy() for k, v in calc(inp, forces=True).items()}
    print(f"[parity] 300-atom molecule: dE={abs(out['energy'][0] - ref['energy'][0]):.3e} "
          f"dF={np.abs(out['forces'] - ref['forces']).max():.3e}")
    assert abs(out["energy"][0] - ref["energy"][0]) < 3e-4   # total energy of 300 atoms: fp32 round-off of the reference itself
    # dense 300-atom blob: |F| reaches 36 eV/A; reference CPU<->GPU practice is rtol 1e-4 / atol 1e-5 (tests/test_calculator_gpu.py:106-137)
    assert np.abs(out["forces"] - ref["forces"]).max() < FORCE_ATOL + 1e-5 * np.abs(ref["forces"]).max()
    assert np.abs(out["charges"] - ref["charges"]).max() < CHARGE_ATOL


def test_user_supplied_neighbor_matrix():
    """Caller-provided nbmat (+ padding row, sentinel N) as in keys_in_optional (calculator.py:130-142)."""
    from oracle.nblist_oracle import neighbor_matrix

    inputs, ref, meta = load_golden("caffeine")
    calc = get_calc(meta)
    N = len(inputs["numbers"])
    nb, _, _ = neighbor_matrix(inputs["coord"], 5.0)
    nb = np.concatenate([nb, np.full((1, nb.shape[1]), N, np.int32)])
    out = {k: v.cpu().numpy() for k, v in calc(dict(inputs, nbmat=nb), forces=True).items()}
    assert np.abs(out["forces"] - ref["forces"]).max() < FORCE_ATOL
    assert abs(out["energy"][0] - ref["energy"][0]) < ENERGY_ATOL


def test_user_supplied_periodic_neighbor_matrix_with_atoms_outside_the_cell():
    """Caller-provided nbmat + shifts with a cell refer to the caller's positions AS GIVEN (the reference wraps only
    inside make_nbmat, skipped when 'nbmat' is in the data: calculator.py:1071, 1521-1529).  Half of the atoms are moved
    out of the cell by lattice vectors n_i and the shifts corrected to s' = s + n_i - n_j; results must not change."""
    from aimnetcentral_b200 import ops

    inputs, ref, meta = load_golden("pbc_box60_dsf")
    calc = get_calc(meta)
    N = len(inputs["numbers"])
    cell = torch.as_tensor(inputs["cell"], dtype=torch.float32, device="cuda").reshape(3, 3)
    xw = ops.wrap_positions(torch.as_tensor(inputs["coord"], dtype=torch.float32, device="cuda"), cell)
    nb, cnt, sh = ops.neighbor_list(xw, 5.0, cell=cell.reshape(1, 3, 3), pbc=torch.ones(1, 3, dtype=torch.bool, device="cuda"),
                                    max_neighbors=160)
    w = int(cnt.max().item())
    nb, sh = nb[:, :w].contiguous(), sh[:, :w].contiguous()
    n_i = torch.as_tensor(np.random.default_rng(8).integers(-2, 3, (N, 3)), dtype=torch.int32, device="cuda")
    n_i[::2] = 0
    x_out = xw + n_i.to(torch.float32) @ cell
    j = nb.clamp(max=N - 1).long()
    sh_out = torch.where((nb < N).unsqueeze(-1), sh + n_i.unsqueeze(1) - n_i[j], torch.zeros_like(sh))
    pad = lambda t, v: torch.cat([t, torch.full((1, *t.shape[1:]), v, dtype=t.dtype, device=t.device)])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = calc(dict(inputs, coord=x_out, nbmat=pad(nb, N), shifts=pad(sh_out, 0)), forces=True, stress=True)
    out = {k: v.cpu().numpy() for k, v in res.items()}
    de, df, dq = _report("pbc_box60 caller nbmat, atoms outside the cell", out, ref, N)
    assert de < ENERGY_ATOL and df < FORCE_ATOL and dq < CHARGE_ATOL
    assert np.abs(out["stress"] - ref["stress"]).max() < 1e-5


@pytest.mark.gpu
def test_repeated_evaluations_are_bitwise_identical():
    """Every kernel is atomics-free with fixed reduction orders, so repeating an evaluation must reproduce it bit for
    bit in the default mode too; any difference is a race (tools/race_check.py is the longer version)."""
    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
    from aimnetcentral_b200.structures import allose_supercell, random_molecules

    spec = ModelSpec()
    calc = AIMNet2Calculator((random_state_dict(0, spec), spec), device="cuda:0")
    coord, numbers = random_molecules(300, 50, seed=3)
    z, x, cell = allose_supercell((2, 1, 1), jitter=0.02, seed=1)
    cases = [({"coord": coord, "numbers": numbers, "charge": np.zeros(300, np.float32)}, dict(forces=True), "simple"),
             ({"coord": x, "numbers": z, "charge": np.zeros(1, np.float32), "cell": cell}, dict(forces=True, stress=True), "dsf")]
    for inp, kw, method in cases:
        calc.set_lrcoulomb_method(method)
        ref = {k: v.clone() for k, v in calc(dict(inp), **kw).items()}
        for _ in range(20):
            out = calc(dict(inp), **kw)
            for k in ref:
                assert torch.equal(ref[k], out[k]), k


def test_poisoned_workspace_does_not_change_results():
    """No kernel may read scratch memory it has not written in the same evaluation: with the device workspace filled
    with 0xFF (NaN as fp32 and as fp16, -1 as an index) or 0x7B (huge finite values, which would wreck the row-chunk
    scales of the 3xFP16 GEMM) before the evaluation, results stay bit-identical (tools/poison_probe.py is the longer
    version; compute-sanitizer's initcheck cannot do this for the tensor-core path, it does not see TMA stores)."""
    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
    from aimnetcentral_b200.structures import allose_supercell, random_molecules

    spec = ModelSpec()
    calc = AIMNet2Calculator((random_state_dict(0, spec), spec), device="cuda:0")
    coord, numbers = random_molecules(20, 40, seed=11)
    z, x, cell = allose_supercell((1, 1, 1), jitter=0.02, seed=1)
    cases = [({"coord": coord, "numbers": numbers, "charge": np.zeros(20, np.float32)}, dict(forces=True), "simple"),
             ({"coord": x, "numbers": z, "charge": np.zeros(1, np.float32), "cell": cell}, dict(forces=True, stress=True), "ewald")]
    for rows in (0, 512):
        calc.engine.set_small_m_rows(rows)
        for inp, kw, method in cases:
            calc.set_lrcoulomb_method(method)
            calc.engine.debug_poison(-1)
            ref = {k: v.clone() for k, v in calc(dict(inp), **kw).items() if torch.is_tensor(v)}
            for byte in (0xFF, 0x7B):
                calc.engine.debug_poison(byte)
                out = calc(dict(inp), **kw)
                for k in ref:
                    assert torch.equal(ref[k], out[k]), (rows, method, hex(byte), k)
    calc.engine.debug_poison(-1)


def test_verlet_skin_reuse_matches_rebuild():
    """neighbor_skin > 0: lists built at cutoff + skin are reused while no atom has moved by more than skin / 2, with the
    lattice offsets of the build; the results must equal those of a calculator that rebuilds every call (every pair
    kernel applies its own cutoff, extra listed pairs contribute exactly zero)."""
    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
    from aimnetcentral_b200.structures import allose_supercell, random_molecules

    spec = ModelSpec()
    sd = random_state_dict(0, spec)
    ref_calc = AIMNet2Calculator((sd, spec), device="cuda:0")
    skin_calc = AIMNet2Calculator((sd, spec), device="cuda:0", neighbor_skin=1.0)
    rng = np.random.default_rng(7)
    z, x0, cell = allose_supercell((2, 1, 1), jitter=0.02, seed=1)
    x0 = x0.astype(np.float32)
    # put one atom right at a cell face so that the trajectory carries it across the periodic boundary
    frac = x0 @ np.linalg.inv(cell)
    frac[0, 0] = 0.999
    x0 = (frac @ cell).astype(np.float32)
    for calc in (ref_calc, skin_calc):
        calc.set_lrcoulomb_method("dsf")
    x = x0.copy()
    for step in range(8):
        inp = {"coord": x, "numbers": z, "charge": np.zeros(1, np.float32), "cell": cell}
        a = ref_calc(dict(inp), forces=True, stress=True)
        b = skin_calc(dict(inp), forces=True, stress=True)
        # not bitwise: (i) the long-range rows are unsorted, a list built at another cutoff walks its pairs in another
        # order (fp32 force sums round differently, ~1e-6 eV/A); (ii) a reused step adds the stored lattice offset to
        # x instead of sending x through the fp32 fractional-coordinate round trip of the wrap, which moves atoms by
        # ~1e-6 A (a few 1e-5 eV/A on the forces -- the reference's own sensitivity to a lattice translation).  A wrong
        # image or a missed pair would show up at 1e-2 .. 1 eV/A.
        assert abs(float(a["energy"] - b["energy"])) < ENERGY_ATOL, step
        assert float((a["forces"] - b["forces"]).abs().max()) < FORCE_ATOL, step
        assert float((a["stress"] - b["stress"]).abs().max()) < 1e-6, step
        x = x + rng.normal(0, 0.04, x.shape).astype(np.float32)
        x[0, 0] += 0.15   # atom 0 crosses the cell face at once and passes skin / 2 after four steps: rebuild
    builds, reuses = skin_calc.engine.skin_stats()
    assert reuses >= 2 and builds >= 2, (builds, reuses)
    # isolated molecules: only the short-range list exists
    coord, numbers = random_molecules(8, 40, seed=2)
    ref_calc.set_lrcoulomb_method("simple")
    skin_calc.set_lrcoulomb_method("simple")
    for step in range(4):
        inp = {"coord": coord, "numbers": numbers, "charge": np.zeros(8, np.float32)}
        a = ref_calc(dict(inp), forces=True)
        b = skin_calc(dict(inp), forces=True)
        assert float((a["forces"] - b["forces"]).abs().max()) < 2e-5
        coord = coord + rng.normal(0, 0.02, coord.shape).astype(np.float32)
    # the molecule batch reused its list too; its first list was built in the crystal's wide rows, the row capacity shrank
    # one evaluation later (neighbors.py:135-140 hysteresis), which costs one extra build
    assert skin_calc.engine.skin_stats()[1] >= reuses + 2


def test_cache_static_reuses_lists_for_unchanged_geometry():
    """cache_static=True (reference: calculator.py:1091-1238): repeated evaluation of an unchanged geometry reuses the
    neighbor matrices; a changed geometry rebuilds them; results are those of a plain calculator."""
    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
    from aimnetcentral_b200.structures import random_molecules

    spec = ModelSpec()
    sd = random_state_dict(0, spec)
    plain = AIMNet2Calculator((sd, spec), device="cuda:0")
    cached = AIMNet2Calculator((sd, spec), device="cuda:0", cache_static=True)
    coord, numbers = random_molecules(4, 30, seed=21)
    inp = {"coord": coord, "numbers": numbers, "charge": np.zeros(4, np.float32)}
    ref = plain(dict(inp), forces=True)
    for _ in range(3):
        out = cached(dict(inp), forces=True)
        assert float((out["forces"] - ref["forces"]).abs().max()) < 5e-6   # fp32 round-off of |F| ~ 5 eV/A (lists at cutoff + 1e-3 A)
    builds, reuses = cached.engine.skin_stats()
    assert (builds, reuses) == (1, 2)
    moved = dict(inp, coord=coord + 0.05)
    out = cached(moved, forces=True)
    assert float((out["forces"] - plain(moved, forces=True)["forces"]).abs().max()) < 5e-6
    assert cached.engine.skin_stats()[0] == 2


def test_torchsim_and_pysis_adapters_on_the_engine(monkeypatch):
    """The TorchSim / PySisyphus adapters (aimnet2torchsim.py:110-151, aimnet2pysis.py:44-108 of the reference) drive the real
    engine: a flat multi-system state reproduces the golden batch, and the PySisyphus unit round trip reproduces taxol."""
    from types import SimpleNamespace

    from aimnetcentral_b200 import aimnet2pysis, aimnet2torchsim

    inputs, ref, meta = load_golden("mols_8x50")
    calc = get_calc(meta)
    monkeypatch.setattr(aimnet2torchsim, "_TORCHSIM_IMPORT_ERROR", None)
    dev = torch.device("cuda:0")
    state = SimpleNamespace(positions=torch.as_tensor(inputs["coord"], device=dev), atomic_numbers=torch.as_tensor(inputs["numbers"], device=dev),
                            system_idx=torch.as_tensor(inputs["mol_idx"], device=dev), n_systems=8, pbc=torch.zeros(3, dtype=torch.bool),
                            row_vector_cell=torch.zeros(8, 3, 3, device=dev), device=dev, dtype=torch.float32,
                            charge=torch.as_tensor(inputs["charge"], device=dev))
    out = aimnet2torchsim.AIMNet2TorchSim(calc)(state)
    assert np.abs(out["forces"].cpu().numpy() - ref["forces"]).max() < FORCE_ATOL
    assert np.abs(out["energy"].cpu().numpy() - ref["energy"]).max() < ENERGY_ATOL
    assert out["partial_charges"].data_ptr() == out["charges"].data_ptr()

    inputs, ref, meta = load_golden("taxol_q0")
    monkeypatch.setattr(aimnet2pysis, "_PYSIS_IMPORT_ERROR", None)
    sym = {1: "H", 6: "C", 7: "N", 8: "O"}
    monkeypatch.setattr(aimnet2pysis, "ATOMIC_NUMBERS", {v.lower(): k for k, v in sym.items()})
    bohr, ha = 0.5291772105638411, 27.211386024367243
    monkeypatch.setattr(aimnet2pysis, "BOHR2ANG", bohr)
    monkeypatch.setattr(aimnet2pysis, "ANG2BOHR", 1.0 / bohr)
    monkeypatch.setattr(aimnet2pysis, "AU2EV", ha)
    p = aimnet2pysis.AIMNet2Pysis(get_calc(meta), charge=0, mult=1)
    atoms = [sym[int(z)] for z in inputs["numbers"]]
    r = p.get_forces(atoms, (inputs["coord"].astype(np.float64) / bohr).reshape(-1))
    assert abs(r["energy"] * ha - ref["energy"][0]) < ENERGY_ATOL
    assert np.abs(r["forces"].reshape(-1, 3) * ha / bohr - ref["forces"]).max() < FORCE_ATOL
    small, ref_h, meta_h = load_golden("hessian_caffeine")
    ph = aimnet2pysis.AIMNet2Pysis(get_calc(meta_h), charge=0, mult=1)
    rh_ = ph.get_hessian([sym[int(z)] for z in small["numbers"]], (small["coord"].astype(np.float64) / bohr).reshape(-1))
    assert rh_["hessian"].shape == (72, 72) and rh_["hessian"].dtype == np.float64
    assert np.abs(rh_["hessian"] * ha / bohr / bohr - ref_h["hessian"].reshape(72, 72)).max() < HESSIAN_ATOL


def test_cuda_graph_replay_equals_eager():
    """cuda_graph=True: the second evaluation of a repeating shape captures the step, later ones replay it as one graph
    launch; results are BITWISE those of the eager path (same kernels, same order), for an isolated molecule, a small
    periodic cell with stress, and through the host-buffer entry.  A geometry whose rows outgrow the recorded neighbor
    buffers is detected after the replay and redone eagerly."""
    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
    from aimnetcentral_b200.structures import allose_supercell

    spec = ModelSpec()
    sd = random_state_dict(0, spec)
    eager = AIMNet2Calculator((sd, spec), device="cuda:0")
    graphed = AIMNet2Calculator((sd, spec), device="cuda:0", cuda_graph=True)
    inputs, _, _ = load_golden("taxol_q0")
    rng = np.random.default_rng(3)
    for step in range(6):
        x = (inputs["coord"] + rng.normal(0, 0.02, inputs["coord"].shape)).astype(np.float32)
        d = {"coord": x, "numbers": inputs["numbers"], "charge": np.zeros(1, np.float32)}
        a, b = eager(dict(d), forces=True), graphed(dict(d), forces=True)
        for k in a:
            assert torch.equal(a[k], b[k]), (step, k)
    st = graphed.engine.graph_stats()
    assert st["captures"] == 1 and st["launches"] == 4 and st["fallbacks"] == 0, st
    # squeezed geometry: more neighbors than the recorded buffers hold -> replay detects it, the step is redone eagerly
    xs = (inputs["coord"] * 0.55).astype(np.float32)
    d = {"coord": xs, "numbers": inputs["numbers"], "charge": np.zeros(1, np.float32)}
    a, b = eager(dict(d), forces=True), graphed(dict(d), forces=True)
    assert torch.equal(a["forces"], b["forces"]) and torch.equal(a["energy"], b["energy"])
    assert graphed.engine.graph_stats()["fallbacks"] == 1
    # periodic cell + stress
    z, x0, cell = allose_supercell((1, 1, 1), jitter=0.02, seed=3)
    for calc in (eager, graphed):
        calc.set_lrcoulomb_method("dsf")
    for step in range(4):
        x = (x0 + rng.normal(0, 0.01, x0.shape)).astype(np.float32)
        d = {"coord": x, "numbers": z, "charge": np.zeros(1, np.float32), "cell": cell}
        a, b = eager(dict(d), forces=True, stress=True), graphed(dict(d), forces=True, stress=True)
        for k in a:
            assert torch.equal(a[k], b[k]), (step, k)
    assert graphed.engine.graph_stats()["captures"] >= 2
    # host-buffer entry
    n0 = graphed.engine.graph_stats()["launches"]
    for step in range(4):
        x = (x0 + rng.normal(0, 0.01, x0.shape)).astype(np.float32)
        oa = eager.engine.eval_host(x, z.astype(np.int32), np.zeros(1, np.float32), cell=cell, forces=True, stress=True)
        ob = graphed.engine.eval_host(x, z.astype(np.int32), np.zeros(1, np.float32), cell=cell, forces=True, stress=True)
        for k in oa:
            assert np.array_equal(oa[k], ob[k]), (step, k)
    assert graphed.engine.graph_stats()["launches"] >= n0 + 2


@pytest.mark.gpu
def test_hessian_matches_reference_double_backward():
    """`calc(data, hessian=True)` (calculator.py:904-947; derivatives.py:149-192): (N,3,N,3), eV/A^2, against the reference's
    autograd Hessian (fp32) and its float64 twin; energy / forces of the same call unchanged; symmetric within the same
    bound; acoustic sum rule (tests/test_calculator.py:432-452)."""
    inputs, ref, meta = load_golden("hessian_caffeine")
    calc = get_calc(meta)
    out = calc({k: inputs[k] for k in ("coord", "numbers", "charge")}, forces=True, hessian=True)
    H = out["hessian"].double().cpu().numpy()
    n = len(inputs["numbers"])
    assert H.shape == (n, 3, n, 3) and np.isfinite(H).all()
    H64 = np.load(os.path.join(os.path.dirname(__file__), "golden", "hessian_caffeine.npz"))["ref64_hessian"]
    d64, d32 = np.abs(H - H64).max(), np.abs(H - ref["hessian"]).max()
    Hf = H.reshape(3 * n, 3 * n)
    print(f"[hessian] caffeine |H|max {np.abs(H64).max():.1f}: vs float64 twin {d64:.2e}, vs reference fp32 {d32:.2e}, "
          f"asymmetry {np.abs(Hf - Hf.T).max():.2e}, sum rule {np.abs(H.sum(axis=2)).max():.2e}")
    assert d64 < HESSIAN_ATOL and d32 < HESSIAN_ATOL
    assert np.abs(Hf - Hf.T).max() < HESSIAN_ATOL and np.abs(H.sum(axis=2)).max() < 5e-3
    assert abs(out["energy"].item() - ref["energy"][0]) < ENERGY_ATOL
    assert np.abs(out["forces"].cpu().numpy() - ref["forces"]).max() < FORCE_ATOL
    # matrix-free products (calculator.py:1755-1985), one and several directions
    v = inputs["vectors"]
    hv = calc.hessian_vector_product({k: inputs[k] for k in ("coord", "numbers", "charge")}, v).cpu().numpy()
    want = np.einsum("iajb,kjb->kia", H64, v.astype(np.float64))
    print(f"[hessian] H@v: vs float64 twin {np.abs(hv - want).max():.2e}, vs reference fp32 {np.abs(hv - ref['hvp']).max():.2e}")
    assert hv.shape == v.shape and np.abs(hv - want).max() < 2 * HESSIAN_ATOL   # |v| ~ 1 per component: sums of ~70 noisy terms
    hv1 = calc.hessian_vector_product({k: inputs[k] for k in ("coord", "numbers", "charge")}, v[1]).cpu().numpy()
    assert hv1.shape == (n, 3) and np.abs(hv1 - hv[1]).max() < 1e-4


@pytest.mark.gpu
def test_hessian_batched_inputs():
    """A (B,N,3) batch gives stacked per-structure Hessians, a flat `mol_idx` batch a list (calculator.py:1247-1450); small
    chunks of the displaced batch give the same answer as one chunk."""
    inputs, ref, meta = load_golden("hessian_mols_3x12")
    calc = get_calc(meta)
    out = calc({k: inputs[k] for k in ("coord", "numbers", "charge")}, hessian=True)
    H = out["hessian"].cpu().numpy()
    assert H.shape == (3, 12, 3, 12, 3) and tuple(out["energy"].shape) == ref["energy"].shape    # stacked per-structure results
    assert np.abs(out["energy"].cpu().numpy() - ref["energy"]).max() < ENERGY_ATOL
    assert np.abs(H - ref["hessian"]).max() < HESSIAN_ATOL
    flat = {"coord": inputs["coord"].reshape(-1, 3), "numbers": inputs["numbers"].reshape(-1), "charge": inputs["charge"],
            "mol_idx": np.repeat(np.arange(3), 12)}
    calc.hessian_batch_atoms = 100          # 4 displaced pairs per engine call
    try:
        out_l = calc(flat, hessian=True)
    finally:
        del calc.hessian_batch_atoms
    assert isinstance(out_l["hessian"], list) and len(out_l["hessian"]) == 3
    for b in range(3):
        assert np.abs(out_l["hessian"][b].cpu().numpy() - H[b]).max() < HESSIAN_ATOL   # 6-molecule calls walk the list kernels
    # padded structure: zero blocks for the padding atom
    pad = {"coord": np.concatenate([inputs["coord"][0], np.zeros((1, 3), np.float32)])[None],
           "numbers": np.concatenate([inputs["numbers"][0], [0]])[None].astype(inputs["numbers"].dtype), "charge": inputs["charge"][:1]}
    Hp = calc(pad, hessian=True)["hessian"].cpu().numpy()
    assert Hp.shape == (13, 3, 13, 3) and np.abs(Hp[:12, :, :12] - H[0]).max() < 1e-4 and not Hp[12].any() and not Hp[:, :, 12].any()


@pytest.mark.gpu
def test_neighbor_capacity_grows_and_shrinks():
    """Row capacities follow the geometry both ways (aimnet/calculators/neighbors.py:118-140): a compressed geometry overflows
    the short-range rows and grows them, the relaxed geometry brings them back down one evaluation later (shrink below half
    of the capacity, to widest / 0.75), results stay those of a fresh engine, and the device workspace follows after 32
    evaluations that needed less than half of it."""
    from aimnetcentral_b200 import AIMNet2Calculator

    inputs, ref, meta = load_golden("taxol_q0")
    calc = AIMNet2Calculator(golden_state_dict(meta), device="cuda:0")   # a fresh engine: capacities at their initial values
    eng = calc.engine
    dev = "cuda:0"
    z = torch.tensor(inputs["numbers"], dtype=torch.int32, device=dev)
    q = torch.tensor(inputs["charge"], dtype=torch.float32, device=dev)
    x = torch.tensor(inputs["coord"], dtype=torch.float32, device=dev)
    base = eng.eval(x, z, q, forces=True)
    cap0 = eng.info()["sr_cap"]
    big_n = 40
    xs = (x * 0.45).contiguous()                       # everything within the 5 A cutoff of everything: rows of ~112 entries
    xb = xs.repeat(big_n, 1) + torch.arange(big_n, device=dev).repeat_interleave(x.shape[0]).unsqueeze(1) * 100.0
    mol = torch.arange(big_n, dtype=torch.int32, device=dev).repeat_interleave(x.shape[0])
    eng.eval(xb.contiguous(), z.repeat(big_n), q.repeat(big_n), mol_idx=mol, forces=True)
    grown = eng.info()
    assert grown["sr_cap"] > cap0 and grown["sr_width"] >= 100
    out = eng.eval(x, z, q, forces=True)              # built in the wide layout; the shrink is adopted by the next evaluation
    out2 = eng.eval(x, z, q, forces=True)
    shrunk = eng.info()
    assert shrunk["sr_cap"] < grown["sr_cap"] and shrunk["sr_cap"] >= shrunk["sr_width"]
    for o in (out, out2):
        assert torch.equal(o["energy"], base["energy"]) and torch.equal(o["forces"], base["forces"])
    assert shrunk["workspace_bytes"] == grown["workspace_bytes"]
    for _ in range(34):
        out3 = eng.eval(x, z, q, forces=True)
    assert eng.info()["workspace_bytes"] < grown["workspace_bytes"] // 2
    assert torch.equal(out3["forces"], base["forces"])


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["dsf", "ewald"])
def test_hessian_periodic_finite_symmetric_sumrule(method):
    """The reference's own periodic Hessian checks (tests/test_calculator.py:354-372, 432-452): water in an 8 A box, DSF and
    Ewald Coulomb: shape, finite, non-zero, symmetric and acoustic sum rule within 5e-3 eV/A^2; and against a plain two-point
    difference of the forces at the same geometry."""
    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict

    spec = ModelSpec()
    calc = AIMNet2Calculator((random_state_dict(0, spec), spec), device="cuda:0")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        calc.set_lrcoulomb_method(method, cutoff=8.0) if method == "dsf" else calc.set_lrcoulomb_method(method)
    data = {"coord": np.array([[0.0, 0.0, 0.0], [0.96, 0.0, 0.0], [-0.24, 0.93, 0.0]], np.float32) + 3.0,
            "numbers": np.array([8, 1, 1]), "charge": 0.0, "cell": np.eye(3, dtype=np.float32) * 8.0}
    H = calc(data, hessian=True)["hessian"].double().cpu().numpy()
    assert H.shape == (3, 3, 3, 3) and np.isfinite(H).all() and np.abs(H).sum() > 0
    Hf = H.reshape(9, 9)
    assert np.abs(Hf - Hf.T).max() < 5e-3 and np.abs(H.sum(axis=2)).max() < 5e-3
    h = 4e-3
    for comp in (0, 4, 8):
        xp, xm = data["coord"].copy(), data["coord"].copy()
        xp[comp // 3, comp % 3] += h
        xm[comp // 3, comp % 3] -= h
        fp = calc({**data, "coord": xp}, forces=True)["forces"].double().cpu().numpy()
        fm = calc({**data, "coord": xm}, forces=True)["forces"].double().cpu().numpy()
        row = -(fp - fm).reshape(9) / float(xp[comp // 3, comp % 3] - xm[comp // 3, comp % 3])
        assert np.abs(row - Hf[comp]).max() < 2e-2, (method, comp, np.abs(row - Hf[comp]).max())


@pytest.mark.gpu
def test_first_pass_backward_by_species_equals_generic_kernel():
    """The backward of the first convolution through per-species tables (conv.cu: species_scan / conv0_table / conv0_force)
    against the generic pair kernel on the same engine: molecules, a periodic cell with stress, a batch with many species
    present, and more species than slots (the generic kernel then runs, decided on the device: bitwise equal)."""
    from aimnetcentral_b200 import AIMNet2Calculator, ModelSpec, random_state_dict
    from aimnetcentral_b200.structures import random_molecules

    spec = ModelSpec()
    calc = AIMNet2Calculator((random_state_dict(0, spec), spec), device="cuda:0")
    eng = calc.engine

    def both(data, **kw):
        eng.set_species_first_pass(True)
        a = calc(dict(data), forces=True, **kw)
        eng.set_species_first_pass(False)
        b = calc(dict(data), forces=True, **kw)
        eng.set_species_first_pass(True)
        return a, b

    coord, numbers = random_molecules(70, 40, seed=5)
    a, b = both({"coord": coord, "numbers": numbers, "charge": np.zeros(70, np.float32)})
    assert torch.equal(a["energy"], b["energy"]) and float((a["forces"] - b["forces"]).abs().max()) < 2e-5
    inputs, ref, meta = load_golden("pbc_box60_dsf")
    calc.set_lrcoulomb_method("dsf")
    a, b = both(inputs, stress=True)
    calc.set_lrcoulomb_method("simple")
    assert float((a["forces"] - b["forces"]).abs().max()) < 2e-5 and float((a["stress"] - b["stress"]).abs().max()) < 1e-6
    assert np.abs(a["forces"].cpu().numpy() - ref["forces"]).max() < FORCE_ATOL
    # 14 species in one evaluation (all of aimnet2's elements), and 20 on a model that implements them: more than the 16
    # slots, so the by-species kernels return at once and the generic kernel does the pass
    spec20 = ModelSpec(implemented_species=tuple(range(1, 21)))
    calc20 = AIMNet2Calculator((random_state_dict(0, spec20), spec20), device="cuda:0")
    for zs, exact in (([1, 5, 6, 7, 8, 9, 14, 15, 16, 17, 33, 34, 35, 53], False), (list(range(1, 21)), True)):
        eng = (calc20 if exact else calc).engine
        nat = 60
        x = random_molecules(1, nat, seed=17)[0][0].astype(np.float32)   # a sane geometry; only the species are replaced
        z = np.array([zs[k % len(zs)] for k in range(nat)], np.int64)
        eng.set_species_first_pass(True)
        fa = eng.eval(torch.tensor(x, device="cuda:0"), torch.tensor(z, dtype=torch.int32, device="cuda:0"),
                      torch.zeros(1, device="cuda:0"), forces=True)["forces"]
        eng.set_species_first_pass(False)
        fb = eng.eval(torch.tensor(x, device="cuda:0"), torch.tensor(z, dtype=torch.int32, device="cuda:0"),
                      torch.zeros(1, device="cuda:0"), forces=True)["forces"]
        eng.set_species_first_pass(True)
        scale = float(fb.abs().max())
        assert torch.isfinite(fb).all() and torch.isfinite(fa).all()
        if exact:
            assert torch.equal(fa, fb)
        else:
            assert float((fa - fb).abs().max()) < 1e-5 * max(1.0, scale), (float((fa - fb).abs().max()), scale)
