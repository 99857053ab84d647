"""Input structures for the BASELINE.json configs: small readers + seeded synthetic generators.

No ASE in this image, so the crystal used by cfg-3/cfg-5 (COD 2019828, beta-D,L-allose, the file the
reference ships as examples/2019828.cif) is carried here as its published asymmetric unit + the four P2_1/c
operations, and expanded/super-celled numerically.  Molecule batches for cfg-2/cfg-4 are generated, seeded.
"""
from __future__ import annotations

import math

import numpy as np

SYMBOLS = {"H": 1, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9, "Si": 14, "P": 15, "S": 16, "Cl": 17, "As": 33, "Se": 34,
           "Br": 35, "I": 53}

# COD 2019828: a, b, c (Angstrom), beta (deg); alpha = gamma = 90; space group P 1 21/c 1
_ALLOSE_ABC_BETA = (4.98211, 12.5624, 11.8156, 91.1262)
# asymmetric unit: (Z, x, y, z) fractional
_ALLOSE_SITES = (
    (8, 0.1621, 0.79703, 0.67894), (8, 0.1676, 0.85269, 0.91826), (8, -0.0213, 1.05723, 0.88005),
    (8, 0.3426, 1.21319, 0.80258), (8, 0.2673, 0.96785, 0.63757), (8, 0.6115, 1.14807, 0.56175),
    (6, 0.1454, 0.90015, 0.71937), (6, 0.2848, 0.91736, 0.83377), (6, 0.2552, 1.03366, 0.86734),
    (6, 0.3680, 1.10357, 0.77418), (6, 0.2263, 1.07910, 0.66196), (6, 0.3259, 1.14427, 0.56431),
    (1, 0.3191, 0.7774, 0.6812), (1, -0.0442, 0.9196, 0.7254), (1, 0.2377, 0.7936, 0.9180),
    (1, 0.4758, 0.8998, 0.8280), (1, -0.0440, 1.0847, 0.9421), (1, 0.3529, 1.0470, 0.9387),
    (1, 0.5592, 1.0872, 0.7668), (1, 0.1849, 1.2268, 0.8151), (1, 0.0336, 1.0920, 0.6699),
    (1, 0.6640, 1.1146, 0.5069), (1, 0.2577, 1.1140, 0.4939), (1, 0.2565, 1.2162, 0.5702),
)


def allose_unit_cell():
    """96-atom unit cell of 2019828.cif: numbers (96,), fractional coords (96,3) in [0,1), cell (3,3) rows=a,b,c."""
    a, b, c, beta = _ALLOSE_ABC_BETA
    br = math.radians(beta)
    cell = np.array([[a, 0.0, 0.0], [0.0, b, 0.0], [c * math.cos(br), 0.0, c * math.sin(br)]], dtype=np.float64)
    z = np.array([s[0] for s in _ALLOSE_SITES], dtype=np.int32)
    f = np.array([s[1:] for s in _ALLOSE_SITES], dtype=np.float64)
    x, y, w = f[:, 0], f[:, 1], f[:, 2]
    ops = [
        np.stack([x, y, w], 1),  # +X,+Y,+Z
        np.stack([-x, 0.5 + y, 0.5 - w], 1),  # -X,1/2+Y,1/2-Z
        np.stack([-x, -y, -w], 1),  # -X,-Y,-Z
        np.stack([x, 0.5 - y, 0.5 + w], 1),  # +X,1/2-Y,1/2+Z
    ]
    frac = np.concatenate(ops, 0) % 1.0
    numbers = np.tile(z, 4)
    return numbers, frac, cell


def supercell(numbers, frac, cell, reps):
    """Replicate a unit cell reps=(na,nb,nc) times. Returns numbers (N,), cartesian coord (N,3) f32, cell (3,3) f32."""
    na, nb, nc = reps
    gx, gy, gz = np.meshgrid(np.arange(na), np.arange(nb), np.arange(nc), indexing="ij")
    offs = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], 1).astype(np.float64)
    f = (frac[None, :, :] + offs[:, None, :]).reshape(-1, 3)
    coord = f @ cell
    big = cell * np.array(reps, dtype=np.float64)[:, None]
    return np.tile(numbers, len(offs)).astype(np.int32), coord.astype(np.float32), big.astype(np.float32)


def allose_supercell(reps=(7, 3, 5), jitter: float = 0.02, seed: int = 3):
    """cfg-3 (7,3,5) -> 10 080 atoms; cfg-5 (14,6,10) -> 80 640 atoms."""
    numbers, frac, cell = allose_unit_cell()
    z, coord, big = supercell(numbers, frac, cell, reps)
    if jitter > 0:
        rng = np.random.Generator(np.random.PCG64(seed))
        coord = (coord + rng.normal(0.0, jitter, coord.shape)).astype(np.float32)
    return z, coord, big


def random_molecules(n_mol: int, n_atoms: int, seed: int = 1234, species=((1, 0.5), (6, 0.3), (7, 0.1), (8, 0.1)),
                     dmin: float = 0.9, bond=(1.1, 1.6), box: float = 7.0):
    """Batch of blob-like organic-ish molecules: random growth, each new atom bonded-ish
    (1.1-1.6 A) to a random earlier atom, no pair closer than `dmin`, confined to a ~box^3 blob.

    Returns coord (n_mol, n_atoms, 3) float32, numbers (n_mol, n_atoms) int32.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    zs = np.array([s[0] for s in species], dtype=np.int32)
    ps = np.array([s[1] for s in species], dtype=np.float64)
    numbers = rng.choice(zs, size=(n_mol, n_atoms), p=ps / ps.sum()).astype(np.int32)
    coord = np.zeros((n_mol, n_atoms, 3), dtype=np.float64)
    half = box / 2.0
    for k in range(1, n_atoms):
        todo = np.arange(n_mol)
        for _ in range(200):
            if todo.size == 0:
                break
            parent = rng.integers(0, k, size=todo.size)
            v = rng.standard_normal((todo.size, 3))
            v /= np.linalg.norm(v, axis=1, keepdims=True)
            r = rng.uniform(bond[0], bond[1], size=(todo.size, 1))
            cand = coord[todo, parent] + v * r
            d = np.linalg.norm(coord[todo, :k] - cand[:, None, :], axis=-1).min(axis=1)
            ok = (d >= dmin) & (np.abs(cand).max(axis=1) <= half)
            coord[todo[ok], k] = cand[ok]
            todo = todo[~ok]
        if todo.size:  # extremely unlikely: relax the box constraint
            for m in todo:
                while True:
                    parent = rng.integers(0, k)
                    v = rng.standard_normal(3)
                    cand = coord[m, parent] + v / np.linalg.norm(v) * rng.uniform(*bond)
                    if np.linalg.norm(coord[m, :k] - cand, axis=-1).min() >= dmin:
                        coord[m, k] = cand
                        break
    return coord.astype(np.float32), numbers


def random_periodic_box(n_atoms: int, seed: int = 7, density: float = 0.09, triclinic: bool = True,
                        species=((1, 0.5), (6, 0.3), (7, 0.1), (8, 0.1)), dmin: float = 1.0):
    """Small random triclinic periodic cell for parity fixtures. Returns numbers, coord f32, cell f32."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = (n_atoms / density) ** (1.0 / 3.0)
    cell = np.diag([L, L * 1.1, L * 0.95])
    if triclinic:
        cell[1, 0] = 0.15 * L
        cell[2, 0] = -0.1 * L
        cell[2, 1] = 0.2 * L
    zs = np.array([s[0] for s in species], dtype=np.int32)
    ps = np.array([s[1] for s in species], dtype=np.float64)
    numbers = rng.choice(zs, size=n_atoms, p=ps / ps.sum()).astype(np.int32)
    frac = np.zeros((n_atoms, 3))
    k = 0
    while k < n_atoms:
        f = rng.random(3)
        if k:
            df = frac[:k] - f
            df -= np.round(df)
            if np.linalg.norm(df @ cell, axis=1).min() < dmin:
                continue
        frac[k] = f
        k += 1
    return numbers, (frac @ cell).astype(np.float32), cell.astype(np.float32)


def read_xyz_frame(path: str, frame: int = 0):
    """Minimal multi-frame .xyz reader -> numbers (N,) int32, coord (N,3) float32."""
    with open(path) as fh:
        lines = fh.read().splitlines()
    pos = 0
    for _ in range(frame + 1):
        n = int(lines[pos].split()[0])
        block = lines[pos + 2: pos + 2 + n]
        pos += 2 + n
    z = np.array([SYMBOLS[ln.split()[0]] for ln in block], dtype=np.int32)
    xyz = np.array([[float(t) for t in ln.split()[1:4]] for ln in block], dtype=np.float32)
    return z, xyz


def benchmark_workload(name: str, seed: int = 1234):
    """The synthetic inputs of BASELINE.json configs 1-5 as flat arrays: what bench.py times and what the
    full-size parity fixtures (oracle/make_golden_full.py, tests/test_gpu_fullsize.py) are generated from."""
    import os

    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if name == "cfg2":
        coord, numbers = random_molecules(1024, 50, seed=seed)
        B, n = coord.shape[:2]
        return dict(coord=coord.reshape(-1, 3), numbers=numbers.reshape(-1).astype(np.int32),
                    charge=np.zeros(B, np.float32), mol_idx=np.repeat(np.arange(B), n).astype(np.int32), cell=None,
                    desc="cfg-2: 1024 x 50-atom random organic molecules, aimnet2, E+F, Coulomb simple + DFT-D3",
                    stress=False)
    if name == "cfg3":
        z, x, cell = allose_supercell((7, 3, 5), jitter=0.02, seed=seed)
        return dict(coord=x, numbers=z.astype(np.int32), charge=np.zeros(1, np.float32), mol_idx=None, cell=cell,
                    desc="cfg-3: 10 080-atom allose supercell, PBC, DSF Coulomb + DFT-D3, E+F+stress", stress=True)
    if name == "cfg5":
        z, x, cell = allose_supercell((14, 6, 10), jitter=0.02, seed=seed)
        return dict(coord=x, numbers=z.astype(np.int32), charge=np.zeros(1, np.float32), mol_idx=None, cell=cell,
                    desc="cfg-5: 80 640-atom allose supercell, PBC, Ewald Coulomb (1e-6) + DFT-D3, E+F+stress (one "
                         "replica per GPU)", stress=True, coulomb="ewald")
    if name == "cfg1":
        g = np.load(os.path.join(ROOT, "tests", "golden", "taxol_q0.npz"))
        return dict(coord=g["in_coord"].astype(np.float32), numbers=g["in_numbers"].astype(np.int32),
                    charge=np.zeros(1, np.float32), mol_idx=None, cell=None,
                    desc="cfg-1: taxol, 113 atoms, single molecule, E+F, Coulomb simple + DFT-D3", stress=False)
    if name == "cfg4":
        coord, numbers = random_molecules(512, 80, seed=4321 + seed, box=8.5)
        B, n = coord.shape[:2]
        rng = np.random.default_rng(seed)
        charge = rng.integers(-1, 2, size=B).astype(np.float32)
        nelec = numbers.sum(axis=1) - charge.astype(np.int64)
        mult = np.where(nelec % 2 == 0, rng.choice([1, 3], size=B), 2).astype(np.float32)
        return dict(coord=coord.reshape(-1, 3), numbers=numbers.reshape(-1).astype(np.int32), charge=charge, mult=mult,
                    mol_idx=np.repeat(np.arange(B), n).astype(np.int32), cell=None, channels=2,
                    desc="cfg-4: aimnet2-nse graph (2 charge channels), 512 x 80-atom molecules, E+F+charges+spin charges",
                    stress=False)
    raise ValueError(name)
