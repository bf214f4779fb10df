"""Synthetic workloads of BASELINE.json (numpy only -- no oracle, no torch): shared by bench.py, the tools and the
tests so that everybody measures and checks the same systems."""
from types import SimpleNamespace

import numpy as np


def _bundle(x, q, typ, mol, alpha, L, ntypes):
    n = x.shape[0]
    return SimpleNamespace(x=np.ascontiguousarray(x, dtype=np.float64), q=np.ascontiguousarray(q, dtype=np.float64),
                           type=np.ascontiguousarray(typ, dtype=np.int32), molecule=np.ascontiguousarray(mol, dtype=np.int32),
                           alpha=np.ascontiguousarray(alpha, dtype=np.float64), tag=np.arange(1, n + 1, dtype=np.int32),
                           boxlo=np.zeros(3), boxhi=np.asarray(L, dtype=np.float64), ntypes=ntypes, n=n,
                           periodic=np.ones(3, dtype=np.int32), nspecial=None, special=None)


def lj_charge_fluid(ncell, seed=12345, rho=0.1, jitter=0.3):
    """BASELINE config 2: fcc sites jittered by U(-jitter,jitter) A, density rho atoms/A^3, two types with
    q = +-0.4 e alternating, alpha 1.0 / 0.5 A^3, molecule 0.  N = 4*ncell^3; ncell may be a triple (nx,ny,nz) of fcc
    cells for the brick-shaped boxes of the multi-GPU runs."""
    rng = np.random.default_rng(seed)
    nc = np.array([ncell] * 3 if np.isscalar(ncell) else list(ncell), dtype=np.int64)
    n = 4 * int(nc.prod())
    a = (4.0 / rho) ** (1.0 / 3.0)
    if np.isscalar(ncell):
        a = ((n / rho) ** (1.0 / 3.0)) / ncell  # the historical expression (bit-identical fixtures)
    L = a * nc
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    g = np.stack(np.meshgrid(np.arange(nc[0]), np.arange(nc[1]), np.arange(nc[2]), indexing="ij"), -1).reshape(-1, 3)
    x = ((g[:, None, :] + base[None, :, :]) * a).reshape(-1, 3)
    x = x + rng.uniform(-jitter, jitter, size=x.shape)
    x = np.mod(x, L)
    typ = (np.arange(n) % 2 + 1).astype(np.int32)
    q = np.where(typ == 1, 0.4, -0.4)
    alpha = np.where(typ == 1, 1.0, 0.5)
    return _bundle(x, q, typ, np.zeros(n, dtype=np.int32), alpha, L, 2)


def water_box(nmol_side, seed=2, rho=0.1):
    """BASELINE config 3: rigid 3-site water-like molecules on a jittered cubic lattice with random
    orientations; r_OH = 0.9572 A, HOH = 104.52 deg; O: q -0.8 e, alpha 0.837 A^3 (type 1), H: q +0.4 e, alpha 0.496 A^3
    (type 2); molecule = molecule id; rho atoms/A^3.  N = 3 * nmol_side^3 atoms."""
    rng = np.random.default_rng(seed)
    nmol = nmol_side ** 3
    n = 3 * nmol
    L = (n / rho) ** (1.0 / 3.0)
    a = L / nmol_side
    g = np.stack(np.meshgrid(*[np.arange(nmol_side)] * 3, indexing="ij"), -1).reshape(-1, 3)
    centre = (g + 0.5) * a + rng.uniform(-0.25, 0.25, size=(nmol, 3))
    rot, _ = np.linalg.qr(rng.normal(size=(nmol, 3, 3)))
    rot = rot * np.sign(np.linalg.det(rot))[:, None, None]
    roh, half = 0.9572, np.deg2rad(104.52) / 2.0
    local = np.array([[0.0, 0.0, 0.0], [roh * np.sin(half), roh * np.cos(half), 0.0],
                      [-roh * np.sin(half), roh * np.cos(half), 0.0]])
    x = centre[:, None, :] + np.einsum("mij,aj->mai", rot, local)
    x = np.mod(x.reshape(-1, 3), L)
    typ = np.tile(np.array([1, 2, 2], dtype=np.int32), nmol)
    q = np.where(typ == 1, -0.8, 0.4)
    alpha = np.where(typ == 1, 0.837, 0.496)
    mol = np.repeat(np.arange(1, nmol + 1, dtype=np.int32), 3)
    return _bundle(x, q, typ, mol, alpha, [L, L, L], 2)


def mof_supercell(cell_npz, R):
    """BASELINE config 4 shape: the reference's MOF-5 + CO2 example cell (924 atoms, 10 atom types, bond topology as special
    lists, 101 molecules; the arrays of `cell_npz` as dumped from the reference run, tests/golden/co2_singlepoint_step0.npz)
    replicated R x R x R like LAMMPS `replicate` would (new atom ids and molecule ids per image).  Returns (system,
    cut_coul, pair_coeff lines)."""
    fx = dict(np.load(cell_npz, allow_pickle=False))
    n0 = fx["x"].shape[0]
    prd = fx["boxhi"] - fx["boxlo"]
    nmol0 = int(fx["molecule"].max())
    xs, tags, mols, specs = [], [], [], []
    r = 0
    for ix in range(R):
        for iy in range(R):
            for iz in range(R):
                xs.append(fx["x"] - fx["boxlo"] + np.array([ix, iy, iz]) * prd)
                tags.append(fx["tag"] + r * n0)
                mols.append(np.where(fx["molecule"] > 0, fx["molecule"] + r * nmol0, 0))
                specs.append(np.where(fx["special"] > 0, fx["special"] + r * n0, 0))
                r += 1
    rep = R ** 3
    sysm = SimpleNamespace(x=np.ascontiguousarray(np.concatenate(xs)), q=np.tile(fx["q"], rep),
                           type=np.tile(fx["type"], rep).astype(np.int32), alpha=np.tile(fx["alpha"], rep),
                           tag=np.concatenate(tags).astype(np.int32), molecule=np.concatenate(mols).astype(np.int32),
                           nspecial=np.ascontiguousarray(np.tile(fx["nspecial"], (rep, 1)).astype(np.int32)),
                           special=np.ascontiguousarray(np.concatenate(specs).astype(np.int32)), boxlo=np.zeros(3),
                           boxhi=prd * R, ntypes=int(fx["ntypes"]), n=n0 * rep, periodic=np.ones(3, dtype=np.int32))
    return sysm, 12.8345, str(fx["pair_coeff"]).splitlines()
