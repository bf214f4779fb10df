"""Shared by the rigid-body oracle test (CPU) and the device parity test (GPU): fixtures, the trajectory check and the
two drivers it accepts (oracle / device) with the same four calls."""
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"


def load(name):
    z = np.load(GOLD / f"{name}.npz")
    return {k: z[k] for k in z.files}


def minimg(d, L):
    return d - L * np.rint(d / L)


def fix_args(fx):
    """the keyword part of the fixture's fix line: (style, temp triple or None, tparam triple)"""
    t = str(fx["fix_line"]).split()
    style, temp, tparam = t[3], None, (10, 1, 3)
    if "temp" in t:
        i = t.index("temp")
        temp = tuple(float(v) for v in t[i + 1:i + 4])
    if "tparam" in t:
        i = t.index("tparam")
        tparam = tuple(int(v) for v in t[i + 1:i + 4])
    return style, temp, tparam


def ingroup(fx):
    g = str(fx["group_moving"]) if "group_moving" in fx else "all"
    return np.ones(fx["molecule"].shape[0], bool) if g == "all" else fx["molecule"] > 1


def make_oracle(fx):
    import sys
    if str(ROOT) not in sys.path:
        sys.path.insert(0, str(ROOT))
    from oracle import rigidref as RR
    _, temp, tparam = fix_args(fx)
    return RR.RigidRef(fx["x"][0], fx["v_init"], fx["image"][0], fx["mass"], fx["molecule"], ingroup(fx), fx["boxlo"],
                       fx["boxhi"], float(fx["dt"]), float(fx["ftm2v"]), float(fx["mvv2e"]), float(fx["boltz"]),
                       temp=temp, tparam=tparam)


class OracleDriver:
    def __init__(self, R):
        self.R = R

    def setup(self, f):
        self.R.setup(f, vflag=1)

    def initial(self, f, frac):
        self.R.initial_integrate(f, vflag=1, run_fraction=frac)

    def final(self, f):
        self.R.final_integrate(f)

    x = property(lambda s: s.R.x)
    v = property(lambda s: s.R.v)
    virial = property(lambda s: s.R.virial)

    def scalar(self):
        return self.R.compute_scalar()


def check_trajectory(fx, D, tol_x, tol_v, tol_vir):
    """setup + every step of the fixture: positions (minimum image), velocities, the fix's virial and -- for
    rigid/nvt -- its scalar against the reference's dump and thermo output."""
    L = fx["boxhi"] - fx["boxlo"]
    vol = float(np.prod(L))
    cols = list(fx["thermo_cols"])
    pcol = [cols.index(f"c_pfix[{k}]") for k in range(1, 7)]
    scol = [i for i, c in enumerate(cols) if c.startswith("f_")]
    nktv2p = float(fx["nktv2p"])
    nfr = fx["x"].shape[0]
    vscale = np.abs(fx["v"][0]).max()

    def check_scalars(n):
        vir_ref = fx["thermo"][n, pcol] * vol / nktv2p
        assert np.abs(D.virial - vir_ref).max() < tol_vir * max(np.abs(vir_ref).max(), 1.0), (n, D.virial, vir_ref)
        if scol:
            ref = fx["thermo"][n, scol[0]]
            assert abs(D.scalar() - ref) < 1e-9 * max(abs(ref), 1.0), (n, D.scalar(), ref)

    D.setup(fx["f"][0])
    assert np.abs(D.v - fx["v"][0]).max() < tol_v * vscale
    check_scalars(0)
    for n in range(nfr - 1):
        D.initial(fx["f"][n], (n + 1) / float(fx["nrun"]))
        assert np.abs(minimg(D.x - fx["x"][n + 1], L)).max() < tol_x * L.max(), n
        D.final(fx["f"][n + 1])
        assert np.abs(D.v - fx["v"][n + 1]).max() < tol_v * vscale, n
        check_scalars(n + 1)


CO2_MASS = {1: 65.39, 2: 15.999, 3: 15.999, 4: 1.0079, 5: 12.011, 6: 12.011, 7: 12.011, 8: 12.01, 9: 16.0, 10: 0.000001}


def shipped_co2_system():
    """the reference's MOF5+CO2 example (polarization/examples/MOF5+CO2, masses of its input script) from the golden
    fixture co2_singlepoint_step0: positions, per-atom masses, molecule ids, image flags that make every molecule whole,
    and the rigid group `molecule > 1` of the script.  The reference aborts on this input with
    "Fix rigid: Bad principal moments" (fix_rigid.cpp:2099): the CO2 model's two 1e-6 amu off-axis sites leave a smallest
    principal moment of 2.4e-6, which is zeroed (< 1e-7 of the largest) and then fails the 1e-6 consistency check."""
    import polhelpers as H
    fx = H.load_fixture("co2_singlepoint_step0")
    x, mol, tag = fx["x"], fx["molecule"], fx["tag"]
    L = fx["boxhi"] - fx["boxlo"]
    mass = np.array([CO2_MASS[int(t)] for t in fx["type"]])
    image = np.zeros((len(tag), 3), dtype=np.int64)
    first = {}
    for i in np.argsort(tag):
        m = int(mol[i])
        if m not in first:
            first[m] = x[i]
        image[i] = -np.rint((x[i] - first[m]) / L).astype(np.int64)
    return dict(x=np.ascontiguousarray(x), tag=np.ascontiguousarray(tag, dtype=np.int32), molecule=np.ascontiguousarray(mol, dtype=np.int32),
                mass=mass, image=image, ingroup=mol > 1, boxlo=fx["boxlo"], boxhi=fx["boxhi"])
