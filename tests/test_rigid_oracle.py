"""Pins the rigid-body oracle (oracle/rigidref.py) against golden vectors dumped from the reference binary
(tests/golden/rigid_*.npz, generator oracle/make_golden_rigid.py): `fix rigid/nve molecule` and `fix rigid/nvt
molecule` of src/RIGID/fix_rigid_nh.cpp, fed with the reference's own per-step forces."""
import numpy as np
import pytest

import rigid_common as RC

CASES = ["rigid_water_nve", "rigid_water_nvt", "rigid_water_nvt5", "rigid_h2_nve", "rigid_methane_nve"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_trajectory(name):
    fx = RC.load(name)
    R = RC.make_oracle(fx)
    RC.check_trajectory(fx, RC.OracleDriver(R), tol_x=1e-11, tol_v=1e-10, tol_vir=1e-7)


@pytest.mark.parametrize("name", ["rigid_water_nve", "rigid_h2_nve", "rigid_methane_nve"])
def test_oracle_dof_matches_thermo_temperature(name):
    """compute temp divides by 3N - 3 - fix->dof(): recover the dof from the reference's temp and ke columns"""
    fx = RC.load(name)
    R = RC.make_oracle(fx)
    cols = list(fx["thermo_cols"])
    temp, ke = fx["thermo"][0, cols.index("Temp")], fx["thermo"][0, cols.index("KinEng")]
    dof_ref = 2.0 * ke / (float(fx["boltz"]) * temp)
    n = fx["x"].shape[1]
    assert abs((3 * n - 3 - R.dof()) - dof_ref) < 1e-6 * dof_ref


def test_linear_bodies_have_a_zero_moment():
    fx = RC.load("rigid_h2_nve")
    R = RC.make_oracle(fx)
    assert ((R.inertia == 0.0).sum(1) == 1).all()
    R.dof()
    assert R.nlinear == R.nbody


def test_wrapping_atoms_and_pre_neighbor_do_not_change_the_trajectory():
    """Domain::pbc + FixRigid::pre_neighbor (remap xcm, image_shift) between steps, as on a reneighbor step"""
    fx = RC.load("rigid_water_nve")
    A, B = RC.make_oracle(fx), RC.make_oracle(fx)
    L = fx["boxhi"] - fx["boxlo"]
    A.setup(fx["f"][0]), B.setup(fx["f"][0])
    image = fx["image"][0].astype(np.int64)
    for n in range(fx["x"].shape[0] - 1):
        for R in (A, B):
            R.initial_integrate(fx["f"][n])
        # B: wrap every atom into the box (update its true image flags), then pre_neighbor
        true_unw = B.x + (B.xcmimage + B.imagebody[B.body]) * L
        shift = np.floor((B.x - fx["boxlo"]) / L).astype(np.int64)
        B.x = B.x - shift * L
        image = np.rint((true_unw - B.x) / L).astype(np.int64)
        B.pre_neighbor(image)
        assert np.abs(RC.minimg(A.x - B.x, L)).max() < 1e-11
        for R in (A, B):
            R.final_integrate(fx["f"][n + 1])
        assert np.abs(A.v - B.v).max() < 1e-12 * np.abs(A.v).max()


def test_shipped_co2_example_aborts_like_the_reference():
    """MOF5+CO2 as shipped stops in the reference with "Fix rigid: Bad principal moments" (SURVEY §4); the restated
    setup_bodies_static reaches the same verdict"""
    import sys
    sys.path.insert(0, str(RC.ROOT))
    from oracle import rigidref as RR
    s = RC.shipped_co2_system()
    with pytest.raises(RuntimeError, match="Fix rigid: Bad principal moments"):
        RR.RigidRef(s["x"], np.zeros_like(s["x"]), s["image"], s["mass"], s["molecule"], s["ingroup"], s["boxlo"], s["boxhi"],
                    1.0, 1.0, 1.0, 1.0)


def test_reneighboring_every_step_with_the_reference_image_flags():
    """A hot 50-step reference run that re-neighbors on every step (21 face crossings): after each initial_integrate
    the atoms are wrapped the way Domain::pbc does it, the resulting TRUE image flags must equal the reference's dump,
    the wrapped coordinates must equal the dumped ones (not just modulo the box), and pre_neighbor gets the flags --
    the bookkeeping of xcmimage / imagebody exactly as LAMMPS drives it."""
    fx = RC.load("rigid_water_nve_wrap")
    assert int((np.abs(np.diff(fx["image"], axis=0)) > 0).any(axis=2).sum()) >= 10
    R = RC.make_oracle(fx)
    lo, hi = fx["boxlo"], fx["boxhi"]
    L = hi - lo
    image = fx["image"][0].astype(np.int64).copy()
    R.setup(fx["f"][0])
    assert np.abs(R.v - fx["v"][0]).max() < 1e-10 * np.abs(fx["v"][0]).max()
    for n in range(fx["x"].shape[0] - 1):
        R.initial_integrate(fx["f"][n])
        # Domain::pbc (domain.cpp:520-600): one period per call is all a step can need
        below, above = R.x < lo, R.x >= hi
        R.x = np.where(below, R.x + L, np.where(above, R.x - L, R.x))
        image = image - below.astype(np.int64) + above.astype(np.int64)
        assert np.array_equal(image, fx["image"][n + 1]), n
        assert np.abs(R.x - fx["x"][n + 1]).max() < 1e-11 * L.max(), n
        R.pre_neighbor(image)
        R.final_integrate(fx["f"][n + 1])
        assert np.abs(R.v - fx["v"][n + 1]).max() < 1e-10 * np.abs(fx["v"][0]).max(), n
