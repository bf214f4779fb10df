"""GPU parity of the device rigid-body integrator (polb200_rigid_*) -- through the C ABI, against
the reference's own trajectories (golden vectors dumped from the reference binary) and against the oracle, plus
size-independent properties at the BASELINE config-3 size (256k-atom rigid water box)."""
import numpy as np
import pytest

import rigid_common as RC
from gpu_common import pb

pytestmark = pytest.mark.gpu

CASES = ["rigid_water_nve", "rigid_water_nvt", "rigid_water_nvt5", "rigid_h2_nve", "rigid_methane_nve"]


class DeviceDriver:
    """the four calls of rigid_common.check_trajectory on the device integrator, host buffers through the C ABI"""

    def __init__(self, fx, order=None):
        n = fx["x"].shape[1]
        self.order = np.arange(n) if order is None else order          # caller order = a permutation of the id order
        self.inv = np.argsort(self.order)
        o = self.order
        _, temp, tparam = RC.fix_args(fx)
        self.tag = np.ascontiguousarray(fx["tag"][o], dtype=np.int32)
        self._x = np.ascontiguousarray(fx["x"][0][o])
        self._v = np.ascontiguousarray(fx["v_init"][o])
        self.R = pb.Rigid(device=0)
        self.info = self.R.init(self.tag, fx["molecule"][o], fx["mass"][o], fx["image"][0][o], self._x, self._v, fx["boxlo"],
                                fx["boxhi"], float(fx["dt"]), ingroup=RC.ingroup(fx)[o].astype(np.int32), temp=temp,
                                tparam=tparam, ftm2v=float(fx["ftm2v"]), mvv2e=float(fx["mvv2e"]), boltz=float(fx["boltz"]))

    def _f(self, f):
        return np.ascontiguousarray(f[self.order])

    def setup(self, f):
        self.R.setup(self.tag, self._x, self._v, self._f(f), vflag=1)

    def initial(self, f, frac):
        self.R.initial_integrate(self.tag, self._x, self._v, self._f(f), vflag=1, run_fraction=frac)

    def final(self, f):
        self.R.final_integrate(self.tag, self._x, self._v, self._f(f))

    x = property(lambda s: s._x[s.inv])
    v = property(lambda s: s._v[s.inv])
    virial = property(lambda s: s.R.virial())

    def scalar(self):
        return self.R.scalars()[0]


@pytest.mark.parametrize("name", CASES)
def test_device_reproduces_reference_trajectory(name):
    fx = RC.load(name)
    D = DeviceDriver(fx)
    RC.check_trajectory(fx, D, tol_x=1e-11, tol_v=1e-10, tol_vir=1e-7)
    assert D.R.launch_count() > 0
    D.R.close()


@pytest.mark.parametrize("name", ["rigid_water_nvt", "rigid_h2_nve", "rigid_methane_nve"])
def test_device_body_state_matches_oracle(name):
    """every per-body array after setup and after two steps, against the oracle (1e-11 relative)"""
    fx = RC.load(name)
    D, O = DeviceDriver(fx), RC.OracleDriver(RC.make_oracle(fx))
    assert D.info.nbody == O.R.nbody and D.info.nf_r == O.R.nf_r and D.info.nf_t == O.R.nf_t
    assert D.R.dof(D.tag) == O.R.dof()

    def compare():
        for name_, ref in (("xcm", O.R.xcm), ("vcm", O.R.vcm), ("angmom", O.R.angmom), ("omega", O.R.omega), ("quat", O.R.quat),
                           ("conjqm", O.R.conjqm), ("inertia", O.R.inertia), ("ex", O.R.ex), ("ey", O.R.ey), ("ez", O.R.ez),
                           ("fcm", O.R.fcm), ("torque", O.R.torque)):
            got = D.R.fetch(name_)
            assert np.abs(got - ref).max() <= 1e-11 * max(np.abs(ref).max(), 1e-300), name_
        assert np.abs(D.R.fetch("masstotal")[:, 0] - O.R.masstotal).max() < 1e-13

    for drv in (D, O):
        drv.setup(fx["f"][0])
    compare()
    for n in range(2):
        for drv in (D, O):
            drv.initial(fx["f"][n], (n + 1) / float(fx["nrun"]))
            drv.final(fx["f"][n + 1])
        compare()
    D.R.close()


def test_caller_order_does_not_matter():
    """per-atom body data is keyed by atom id: a shuffled (and re-shuffled between steps) caller order gives the same
    trajectory bit for bit"""
    fx = RC.load("rigid_water_nve")
    n = fx["x"].shape[1]
    rng = np.random.default_rng(5)
    A, B = DeviceDriver(fx), DeviceDriver(fx, order=rng.permutation(n))
    for drv in (A, B):
        drv.setup(fx["f"][0])
    for s in range(3):
        if s == 2:  # re-sort B's atoms between steps, as LAMMPS' atom sorting does
            xs, vs = B.x, B.v
            B.order = rng.permutation(n)
            B.inv = np.argsort(B.order)
            B.tag = np.ascontiguousarray(fx["tag"][B.order], dtype=np.int32)
            B._x, B._v = np.ascontiguousarray(xs[B.order]), np.ascontiguousarray(vs[B.order])
        for drv in (A, B):
            drv.initial(fx["f"][s], 0.0)
            drv.final(fx["f"][s + 1])
        assert np.array_equal(A.x, B.x) and np.array_equal(A.v, B.v)
    A.R.close(), B.R.close()


def test_pre_neighbor_after_wrapping():
    """wrap all atoms into the box between steps (Domain::pbc), hand the new image flags to pre_neighbor: same
    trajectory as the unwrapped run modulo the box"""
    fx = RC.load("rigid_water_nve")
    L = fx["boxhi"] - fx["boxlo"]
    A, B = DeviceDriver(fx), DeviceDriver(fx)
    image = fx["image"][0].astype(np.int64)
    for drv in (A, B):
        drv.setup(fx["f"][0])
    for s in range(fx["x"].shape[0] - 1):
        for drv in (A, B):
            drv.initial(fx["f"][s], 0.0)
        # B: true unwrapped position = x + image*L with the flags B last declared; re-wrap, re-declare
        unw = B._x + image * L
        shift = np.floor((B._x - fx["boxlo"]) / L).astype(np.int64)
        B._x = np.ascontiguousarray(B._x - shift * L)
        image = np.rint((unw - B._x) / L).astype(np.int64)
        B.R.pre_neighbor(B.tag, image)
        assert np.abs(RC.minimg(A.x - B.x, L)).max() < 1e-11
        assert (B._x >= fx["boxlo"]).all() and (B._x < fx["boxhi"]).all()
        for drv in (A, B):
            drv.final(fx["f"][s + 1])
        assert np.abs(A.v - B.v).max() < 1e-12 * np.abs(A.v).max()
    A.R.close(), B.R.close()


def test_large_bodies_use_the_warp_path():
    """one 192-atom body (all water atoms declared one molecule would not be rigid-consistent, so build a random rigid
    cluster): warp-per-body force/torque sums agree with the oracle"""
    import sys
    sys.path.insert(0, str(RC.ROOT))
    from oracle import rigidref as RR
    rng = np.random.default_rng(11)
    nb, per = 40, 50
    n = nb * per
    L = np.array([60.0, 60.0, 60.0])
    centre = rng.uniform(0, 60, size=(nb, 1, 3))
    x = (centre + rng.normal(scale=2.0, size=(nb, per, 3))).reshape(n, 3)
    image = np.floor(x / L).astype(np.int64)
    x = x - image * L
    mol = np.repeat(np.arange(1, nb + 1), per).astype(np.int32)
    mass = rng.uniform(1.0, 16.0, size=n)
    v = rng.normal(scale=0.01, size=(n, 3))
    tag = np.arange(1, n + 1, dtype=np.int32)
    f0, f1 = rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
    O = RR.RigidRef(x, v, image, mass, mol, np.ones(n, bool), np.zeros(3), L, 1.0, pb.REAL_FTM2V, pb.REAL_MVV2E, pb.REAL_BOLTZ)
    R = pb.Rigid(device=0)
    info = R.init(tag, mol, mass, image, x, v, np.zeros(3), L, 1.0)
    assert info.maxmembers == per and info.nbody == nb
    xd, vd = np.ascontiguousarray(x), np.ascontiguousarray(v)
    O.setup(f0)
    R.setup(tag, xd, vd, np.ascontiguousarray(f0))
    O.initial_integrate(f0)
    R.initial_integrate(tag, xd, vd, np.ascontiguousarray(f0))
    O.final_integrate(f1)
    R.final_integrate(tag, xd, vd, np.ascontiguousarray(f1))
    assert np.abs(xd - O.x).max() < 1e-11 * 60 and np.abs(vd - O.v).max() < 1e-10 * np.abs(O.v).max()
    assert np.abs(R.virial() - O.virial).max() < 1e-9 * np.abs(O.virial).max()
    for name_, ref in (("fcm", O.fcm), ("torque", O.torque)):
        assert np.abs(R.fetch(name_) - ref).max() < 1e-11 * np.abs(ref).max()
    R.close()


def test_config3_water_box_properties():
    """BASELINE config 3 size (255 552 atoms, 85 184 rigid molecules): properties that need no reference run --
    (1) bond lengths and the HOH angle are preserved to rounding over 20 steps of free flight plus random forces,
    (2) with zero forces the total linear momentum and every body's angular momentum are conserved,
    (3) device-resident (on_device = 1) and host-buffer calls give identical results."""
    import sys
    import torch
    sys.path.insert(0, str(RC.ROOT / "lammps-induced-dipole-polarization-pair-style_b200"))
    import workloads as W
    s = W.water_box(44)
    n = s.n
    rng = np.random.default_rng(3)
    mass = np.where(s.type == 1, 15.9994, 1.008)
    v0 = rng.normal(scale=0.005, size=(n, 3))
    image = np.zeros((n, 3), dtype=np.int64)
    # molecules were wrapped atom by atom: recover consistent image flags from the first atom of each molecule
    L = s.boxhi - s.boxlo
    first = s.x[0::3].repeat(3, axis=0)
    image = -np.rint((s.x - first) / L).astype(np.int64)
    R = pb.Rigid(device=0)
    info = R.init(s.tag, s.molecule, mass, image, s.x, v0, s.boxlo, s.boxhi, 1.0)
    assert info.nbody == n // 3 and info.nlinear == 0 and info.nf_r == n

    def geometry(x):
        u = x + image * L
        o, h1, h2 = u[0::3], u[1::3], u[2::3]
        a, b = h1 - o, h2 - o
        cosang = (a * b).sum(1) / np.linalg.norm(a, axis=1) / np.linalg.norm(b, axis=1)
        return np.linalg.norm(a, axis=1), np.linalg.norm(b, axis=1), np.degrees(np.arccos(cosang))

    x, v = s.x.copy(), v0.copy()      # the integrator works in place: keep the initial state for the second handle
    zero = np.zeros((n, 3))
    R.setup(s.tag, x, v, zero, vflag=0)
    p0 = (mass[:, None] * v).sum(0)
    L0 = R.fetch("angmom").copy()
    for _ in range(10):
        R.initial_integrate(s.tag, x, v, zero, vflag=0)
        R.final_integrate(s.tag, x, v, zero)
    assert np.abs((mass[:, None] * v).sum(0) - p0).max() < 1e-9 * np.abs(mass[:, None] * v).sum()
    assert np.abs(R.fetch("angmom") - L0).max() < 1e-12 * np.abs(L0).max()
    # device-resident continuation vs host-buffer continuation with the same random forces
    xt, vt = torch.from_numpy(x).cuda(), torch.from_numpy(v).cuda()
    tagt = torch.from_numpy(s.tag).cuda()
    torch.cuda.synchronize()
    R2 = pb.Rigid(device=0)
    R2.init(s.tag, s.molecule, mass, image, s.x, v0, s.boxlo, s.boxhi, 1.0)
    x2, v2 = s.x.copy(), v0.copy()
    R2.setup(s.tag, x2, v2, zero, vflag=0)
    for _ in range(10):
        R2.initial_integrate(s.tag, x2, v2, zero, vflag=0)
        R2.final_integrate(s.tag, x2, v2, zero)
    assert np.array_equal(x2, x) and np.array_equal(v2, v)
    fprev = zero
    for _ in range(10):
        f = rng.normal(scale=2.0, size=(n, 3))
        ft_prev, ft = torch.from_numpy(fprev).cuda(), torch.from_numpy(f).cuda()
        torch.cuda.synchronize()   # the library runs on its own stream: torch's copies must have landed
        R.initial_integrate_device(n, tagt.data_ptr(), xt.data_ptr(), vt.data_ptr(), ft_prev.data_ptr(), vflag=1)
        R.final_integrate_device(n, tagt.data_ptr(), xt.data_ptr(), vt.data_ptr(), ft.data_ptr())
        R2.initial_integrate(s.tag, x2, v2, np.ascontiguousarray(fprev), vflag=1)
        R2.final_integrate(s.tag, x2, v2, f)
        fprev = f
    torch.cuda.synchronize()
    assert np.array_equal(xt.cpu().numpy(), x2) and np.array_equal(vt.cpu().numpy(), v2)
    assert np.array_equal(R.virial(), R2.virial())
    r1, r2, ang = geometry(x2)
    assert np.abs(r1 - 0.9572).max() < 1e-11 and np.abs(r2 - 0.9572).max() < 1e-11 and np.abs(ang - 104.52).max() < 1e-9
    print(f"rigid step at {n} atoms: initial+final {R.last_ms():.3f} ms (final half), launches {R.launch_count()}")
    R.close(), R2.close()


def test_thermostat_chain_round_trip():
    """FixRigidNH::write_restart / restart: stop a rigid/nvt run after 2 steps, carry the chain state (and x, v) into a
    fresh handle, continue.  The restarted device run is compared with the equally restarted ORACLE run (1e-11): a
    restart rebuilds the principal axes from the current positions, and the NO_SQUISH splitting is not invariant under
    relabelling the axes, so neither the reference nor we continue the uninterrupted trajectory bit for bit (the
    oracle shows 5e-5 A after one 2 fs step) -- what must hold exactly is the thermostat state and the fix's scalar."""
    import sys
    sys.path.insert(0, str(RC.ROOT))
    from oracle import rigidref as RR
    fx = RC.load("rigid_water_nvt5")
    nrun, half = float(fx["nrun"]), 2
    A = DeviceDriver(fx)
    A.setup(fx["f"][0])
    for s in range(half):
        A.initial(fx["f"][s], (s + 1) / nrun)
        A.final(fx["f"][s + 1])
    chain = A.R.get_chain()
    assert chain.shape == (4, 4) and np.abs(chain[:, 2]).max() > 0.0
    L = fx["boxhi"] - fx["boxlo"]
    # true image flags of A's positions (A never wraps its atoms): x_h + image_h * L = A.x + img * L
    img = fx["image"][half] - np.rint((A.x - fx["x"][half]) / L).astype(np.int64)
    fx2 = dict(fx)
    fx2["x"] = np.stack([A.x] * fx["x"].shape[0])
    fx2["image"] = np.stack([img] * fx["x"].shape[0]).astype(np.int32)
    fx2["v_init"] = A.v
    B = DeviceDriver(fx2)
    B.R.set_chain(chain)
    assert np.array_equal(B.R.get_chain(), chain)
    _, temp, tparam = RC.fix_args(fx)
    O = RR.RigidRef(A.x, A.v, img, fx["mass"], fx["molecule"], RC.ingroup(fx), fx["boxlo"], fx["boxhi"], float(fx["dt"]),
                    float(fx["ftm2v"]), float(fx["mvv2e"]), float(fx["boltz"]), temp=temp, tparam=tparam)
    O.eta_t[:], O.eta_r[:], O.eta_dot_t[:], O.eta_dot_r[:] = chain[:, 0], chain[:, 1], chain[:, 2], chain[:, 3]
    scalar_before = A.scalar()
    B.setup(fx["f"][half])
    O.setup(fx["f"][half])
    assert abs(B.scalar() - scalar_before) < 1e-10 * abs(scalar_before)      # kinetic + chain energy carried over
    for s in range(half, int(nrun)):
        B.initial(fx["f"][s], (s + 1) / nrun)
        O.initial_integrate(fx["f"][s], 1, (s + 1) / nrun)
        B.final(fx["f"][s + 1])
        O.final_integrate(fx["f"][s + 1])
        assert np.abs(RC.minimg(B.x - O.x, L)).max() < 1e-11 * L.max() and np.abs(B.v - O.v).max() < 1e-10 * np.abs(O.v).max()
    got = B.R.get_chain()
    ref = np.stack([O.eta_t, O.eta_r, O.eta_dot_t, O.eta_dot_r], 1)
    assert np.abs(got - ref).max() < 1e-10 * np.abs(ref).max()
    # and the restarted run stays within the integrator's truncation error of the reference's uninterrupted trajectory
    assert np.abs(RC.minimg(B.x - fx["x"][int(nrun)], L)).max() < 1e-3
    A.R.close(), B.R.close()


def test_shipped_co2_example_aborts_like_the_reference():
    """error-behaviour parity: the reference stops on its MOF5+CO2 example with "Fix rigid: Bad principal moments"
    (fix_rigid.cpp:2099); polb200_rigid_init returns the same message"""
    s = RC.shipped_co2_system()
    R = pb.Rigid(device=0)
    with pytest.raises(pb.Polb200Error, match="Fix rigid: Bad principal moments") as e:
        R.init(s["tag"], s["molecule"], s["mass"], s["image"], s["x"], np.zeros_like(s["x"]), s["boxlo"], s["boxhi"], 1.0,
               ingroup=s["ingroup"].astype(np.int32))
    assert e.value.code == pb.ERR_ARG
    R.close()


def test_reneighboring_every_step_with_the_reference_image_flags():
    """the device integrator driven exactly as LAMMPS drives its fix through a hot 50-step reference run that
    re-neighbors on every step (21 face crossings): wrap as Domain::pbc, the true image flags must equal the dumped
    ones, the wrapped coordinates the dumped coordinates, pre_neighbor gets the flags"""
    fx = RC.load("rigid_water_nve_wrap")
    D = DeviceDriver(fx)
    lo, hi = fx["boxlo"], fx["boxhi"]
    L = hi - lo
    image = fx["image"][0].astype(np.int64).copy()
    D.setup(fx["f"][0])
    vs = np.abs(fx["v"][0]).max()
    assert np.abs(D.v - fx["v"][0]).max() < 1e-10 * vs
    for n in range(fx["x"].shape[0] - 1):
        D.initial(fx["f"][n], 0.0)
        below, above = D._x < lo, D._x >= hi
        D._x = np.ascontiguousarray(np.where(below, D._x + L, np.where(above, D._x - L, D._x)))
        image = image - below.astype(np.int64) + above.astype(np.int64)
        same = image == fx["image"][n + 1]
        if not same.all():   # an atom within rounding of a face may wrap one step apart from the reference: adopt its choice
            assert np.abs(RC.minimg(D._x - fx["x"][n + 1], L))[~same].max() < 1e-9, n
            D._x = np.ascontiguousarray(np.where(same, D._x, D._x + (image - fx["image"][n + 1]) * L))
            image = fx["image"][n + 1].astype(np.int64).copy()
        assert np.abs(D.x - fx["x"][n + 1]).max() < 1e-10 * L.max(), n
        D.R.pre_neighbor(D.tag, image)
        D.final(fx["f"][n + 1])
        assert np.abs(D.v - fx["v"][n + 1]).max() < 1e-9 * vs, n
    D.R.close()
