"""ORACLE (test infrastructure, never imported by the product): numpy restatement of the reference's rigid-body
integrator for `fix rigid/nve molecule` and `fix rigid/nvt molecule` -- point particles, orthogonal periodic box.

Follows, under /root/reference/src:
  RIGID/fix_rigid.cpp:130-220     molecule -> body numbering (ascending molecule id among the group's atoms)
  RIGID/fix_rigid.cpp:701-765     init(): dtv / dtf / dtq, tfactor
  RIGID/fix_rigid.cpp:782-889     setup(): fcm, torque, omega, set_v, doubled virial
  RIGID/fix_rigid.cpp:1137-1175   pre_neighbor(), image_shift()
  RIGID/fix_rigid.cpp:1181-1262   dof()
  RIGID/fix_rigid.cpp:1289-1392   set_xv()      RIGID/fix_rigid.cpp:1465-1560  set_v()
  RIGID/fix_rigid.cpp:1605-2112   setup_bodies_static()   :2120-2211  setup_bodies_dynamic()
  RIGID/fix_rigid.cpp:2595-2622   compute_scalar()
  RIGID/fix_rigid_nh.cpp:208-262  init(): nf_t, nf_r, Yoshida-Suzuki weights
  RIGID/fix_rigid_nh.cpp:323-421  setup(): conjqm, thermostat masses and forces
  RIGID/fix_rigid_nh.cpp:428-603  initial_integrate()     :607-790  final_integrate()
  RIGID/fix_rigid_nh.cpp:794-885  nhc_temp_integrate()    :991-1016 compute_scalar()  :1109-1115 compute_temp_target()
  math_extra.cpp:101-175 jacobi/rotate, :234-277 no_squish_rotate, :290-305 angmom_to_omega, :359-415 exyz_to_q /
  q_to_exyz, :422-446 quat_to_mat; math_extra.h:571-641 qnormalize / quatvec / invquatvec; domain.cpp:1329-1410 remap

Parity: PINNED against tests/golden/rigid_*.npz (dumped from the reference binary by oracle/make_golden_rigid.py):
tests/test_rigid_oracle.py.  Per-body arithmetic is vectorised over bodies; per-atom sums into bodies use
np.add.at, which accumulates in atom order like the reference's loops.
"""
import numpy as np

EPS_STATIC = 1.0e-7   # fix_rigid.cpp:52 (scaled by the largest moment)
EPS_NH = 1.0e-7       # fix_rigid_nh.cpp:45
TOLERANCE = 1.0e-6    # fix_rigid.cpp:51
MAXJACOBI = 50        # math_extra.cpp:26


def jacobi3(a):
    """math_extra.cpp:101-161 on one symmetric 3x3 matrix; returns (evalues, evectors[columns])."""
    m = np.array(a, dtype=np.float64)
    ev = np.eye(3)
    b = np.array([m[0, 0], m[1, 1], m[2, 2]])
    d = b.copy()
    z = np.zeros(3)

    def rot(mat, i, j, k, l, s, tau):
        g, h = mat[i, j], mat[k, l]
        mat[i, j] = g - s * (h + g * tau)
        mat[k, l] = h + s * (g - h * tau)

    for it in range(1, MAXJACOBI + 1):
        sm = abs(m[0, 1]) + abs(m[0, 2]) + abs(m[1, 2])
        if sm == 0.0:
            return d, ev
        tresh = 0.2 * sm / 9 if it < 4 else 0.0
        for i in range(2):
            for j in range(i + 1, 3):
                g = 100.0 * abs(m[i, j])
                if it > 4 and abs(d[i]) + g == abs(d[i]) and abs(d[j]) + g == abs(d[j]):
                    m[i, j] = 0.0
                elif abs(m[i, j]) > tresh:
                    h = d[j] - d[i]
                    if abs(h) + g == abs(h):
                        t = m[i, j] / h
                    else:
                        theta = 0.5 * h / m[i, j]
                        t = 1.0 / (abs(theta) + np.sqrt(1.0 + theta * theta))
                        if theta < 0.0:
                            t = -t
                    c = 1.0 / np.sqrt(1.0 + t * t)
                    s = t * c
                    tau = s / (1.0 + c)
                    h = t * m[i, j]
                    z[i] -= h
                    z[j] += h
                    d[i] -= h
                    d[j] += h
                    m[i, j] = 0.0
                    for k in range(0, i):
                        rot(m, k, i, k, j, s, tau)
                    for k in range(i + 1, j):
                        rot(m, i, k, k, j, s, tau)
                    for k in range(j + 1, 3):
                        rot(m, i, k, j, k, s, tau)
                    for k in range(3):
                        rot(ev, k, i, k, j, s, tau)
        b = b + z
        d = b.copy()
        z[:] = 0.0
    raise RuntimeError("Insufficient Jacobi rotations for rigid body")


def exyz_to_q(ex, ey, ez):
    q = np.zeros(4)
    q0sq = 0.25 * (ex[0] + ey[1] + ez[2] + 1.0)
    q1sq = q0sq - 0.5 * (ey[1] + ez[2])
    q2sq = q0sq - 0.5 * (ex[0] + ez[2])
    q3sq = q0sq - 0.5 * (ex[0] + ey[1])
    if q0sq >= 0.25:
        q[0] = np.sqrt(q0sq)
        q[1] = (ey[2] - ez[1]) / (4.0 * q[0])
        q[2] = (ez[0] - ex[2]) / (4.0 * q[0])
        q[3] = (ex[1] - ey[0]) / (4.0 * q[0])
    elif q1sq >= 0.25:
        q[1] = np.sqrt(q1sq)
        q[0] = (ey[2] - ez[1]) / (4.0 * q[1])
        q[2] = (ey[0] + ex[1]) / (4.0 * q[1])
        q[3] = (ex[2] + ez[0]) / (4.0 * q[1])
    elif q2sq >= 0.25:
        q[2] = np.sqrt(q2sq)
        q[0] = (ez[0] - ex[2]) / (4.0 * q[2])
        q[1] = (ey[0] + ex[1]) / (4.0 * q[2])
        q[3] = (ez[1] + ey[2]) / (4.0 * q[2])
    elif q3sq >= 0.25:
        q[3] = np.sqrt(q3sq)
        q[0] = (ex[1] - ey[0]) / (4.0 * q[3])
        q[1] = (ez[0] + ex[2]) / (4.0 * q[3])
        q[2] = (ez[1] + ey[2]) / (4.0 * q[3])
    return q * (1.0 / np.sqrt(np.dot(q, q)))


def q_to_exyz(q):
    """vectorised over bodies: q[nb,4] -> ex, ey, ez [nb,3]"""
    q0, q1, q2, q3 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    ex = np.stack([q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2.0 * (q1 * q2 + q0 * q3), 2.0 * (q1 * q3 - q0 * q2)], 1)
    ey = np.stack([2.0 * (q1 * q2 - q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2.0 * (q2 * q3 + q0 * q1)], 1)
    ez = np.stack([2.0 * (q1 * q3 + q0 * q2), 2.0 * (q2 * q3 - q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3], 1)
    return ex, ey, ez


def quatvec(a, b):
    return np.stack([-a[:, 1] * b[:, 0] - a[:, 2] * b[:, 1] - a[:, 3] * b[:, 2],
                     a[:, 0] * b[:, 0] + a[:, 2] * b[:, 2] - a[:, 3] * b[:, 1],
                     a[:, 0] * b[:, 1] + a[:, 3] * b[:, 0] - a[:, 1] * b[:, 2],
                     a[:, 0] * b[:, 2] + a[:, 1] * b[:, 1] - a[:, 2] * b[:, 0]], 1)


def invquatvec(a, b):
    return np.stack([-a[:, 1] * b[:, 0] + a[:, 0] * b[:, 1] + a[:, 3] * b[:, 2] - a[:, 2] * b[:, 3],
                     -a[:, 2] * b[:, 0] - a[:, 3] * b[:, 1] + a[:, 0] * b[:, 2] + a[:, 1] * b[:, 3],
                     -a[:, 3] * b[:, 0] + a[:, 2] * b[:, 1] - a[:, 1] * b[:, 2] + a[:, 0] * b[:, 3]], 1)


def matvec(ex, ey, ez, v):
    """MathExtra::matvec(ex,ey,ez,v): columns ex,ey,ez"""
    return ex * v[:, 0:1] + ey * v[:, 1:2] + ez * v[:, 2:3]


def transpose_matvec(ex, ey, ez, v):
    return np.stack([(ex * v).sum(1), (ey * v).sum(1), (ez * v).sum(1)], 1)


def tmv_ordered(ex, ey, ez, v):
    """transpose_matvec with the reference's left-to-right sums (np.sum may pair differently)"""
    return np.stack([ex[:, 0] * v[:, 0] + ex[:, 1] * v[:, 1] + ex[:, 2] * v[:, 2],
                     ey[:, 0] * v[:, 0] + ey[:, 1] * v[:, 1] + ey[:, 2] * v[:, 2],
                     ez[:, 0] * v[:, 0] + ez[:, 1] * v[:, 1] + ez[:, 2] * v[:, 2]], 1)


def mv_ordered(ex, ey, ez, v):
    return np.stack([ex[:, k] * v[:, 0] + ey[:, k] * v[:, 1] + ez[:, k] * v[:, 2] for k in range(3)], 1)


def angmom_to_omega(m, ex, ey, ez, idiag):
    wb = tmv_ordered(ex, ey, ez, m)
    with np.errstate(divide="ignore", invalid="ignore"):
        wb = np.where(idiag == 0.0, 0.0, wb / np.where(idiag == 0.0, 1.0, idiag))
    return mv_ordered(ex, ey, ez, wb)


def no_squish_rotate(k, p, q, inertia, dt):
    if k == 1:
        kq = np.stack([-q[:, 1], q[:, 0], q[:, 3], -q[:, 2]], 1)
        kp = np.stack([-p[:, 1], p[:, 0], p[:, 3], -p[:, 2]], 1)
    elif k == 2:
        kq = np.stack([-q[:, 2], -q[:, 3], q[:, 0], q[:, 1]], 1)
        kp = np.stack([-p[:, 2], -p[:, 3], p[:, 0], p[:, 1]], 1)
    else:
        kq = np.stack([-q[:, 3], q[:, 2], -q[:, 1], q[:, 0]], 1)
        kp = np.stack([-p[:, 3], p[:, 2], -p[:, 1], p[:, 0]], 1)
    phi = p[:, 0] * kq[:, 0] + p[:, 1] * kq[:, 1] + p[:, 2] * kq[:, 2] + p[:, 3] * kq[:, 3]
    ik = inertia[:, k - 1]
    small = np.abs(ik) < 1e-6
    phi = np.where(small, phi * 0.0, phi / (4.0 * np.where(small, 1.0, ik)))
    c, s = np.cos(dt * phi)[:, None], np.sin(dt * phi)[:, None]
    return c * p + s * kp, c * q + s * kq


def maclaurin(x):
    x2 = x * x
    x4 = x2 * x2
    return 1.0 + (1.0 / 6.0) * x2 + (1.0 / 120.0) * x4 + (1.0 / 5040.0) * x2 * x4 + (1.0 / 362880.0) * x4 * x4


class RigidRef:
    """State of one `fix rigid/nve|nvt molecule` instance.  Per-atom arrays are in the caller's order and stay there
    (tests key everything by atom id)."""

    def __init__(self, x, v, image, mass, molecule, ingroup, boxlo, boxhi, dt, ftm2v, mvv2e, boltz,
                 temp=None, tparam=(10, 1, 3), natoms_dof_extra=3):
        self.n = n = x.shape[0]
        self.x = np.array(x, dtype=np.float64)
        self.v = np.array(v, dtype=np.float64)
        self.mass = np.asarray(mass, dtype=np.float64)
        self.lo, self.hi = np.asarray(boxlo, float), np.asarray(boxhi, float)
        self.prd = self.hi - self.lo
        ingroup = np.asarray(ingroup, dtype=bool)
        mol = np.asarray(molecule)
        ids = np.unique(mol[ingroup])              # ascending molecule ids = body order (fix_rigid.cpp:203-215)
        self.nbody = nb = len(ids)
        self.body = np.full(n, -1, dtype=np.int64)
        self.body[ingroup] = np.searchsorted(ids, mol[ingroup])
        self.inb = self.body >= 0
        self.nrigid = np.bincount(self.body[self.inb], minlength=nb)
        self.dtv, self.dtf, self.dtq = dt, 0.5 * dt * ftm2v, 0.5 * dt
        self.mvv2e, self.boltz = mvv2e, boltz
        self.nlinear = 0
        self.virial = np.zeros(6)
        self.evflag = 0
        self.tstat = temp is not None
        if self.tstat:
            self.t_start, self.t_stop, self.t_period = temp
            self.t_freq = 1.0 / self.t_period
            self.t_chain, self.t_iter, self.t_order = tparam
        self._setup_bodies_static(np.asarray(image, dtype=np.int64))
        self._setup_bodies_dynamic()
        ndof = 6.0 * nb - self.nlinear             # fix_rigid.cpp:757-764 (nlinear is still 0 on the first init)
        self.tfactor = mvv2e / (ndof * boltz) if ndof > 0 else 0.0
        # FixRigidNH::init, fix_rigid_nh.cpp:232-262
        self.nf_t = 3 * nb
        self.nf_r = 3 * nb - int((np.abs(self.inertia) < EPS_NH).sum())
        if self.tstat:
            if self.t_order == 3:
                w0 = 1.0 / (2.0 - 2.0 ** (1.0 / 3.0))
                self.w = np.array([w0, 1.0 - 2.0 * w0, w0])
            else:
                w0 = 1.0 / (4.0 - 4.0 ** (1.0 / 3.0))
                self.w = np.array([w0, w0, 1.0 - 4.0 * w0, w0, w0])
            c = self.t_chain
            self.eta_t, self.eta_r = np.zeros(c), np.zeros(c)
            self.eta_dot_t, self.eta_dot_r = np.zeros(c), np.zeros(c)
            self.f_eta_t, self.f_eta_r = np.zeros(c), np.zeros(c)
            self.q_t, self.q_r = np.zeros(c), np.zeros(c)

    # ---- helpers ----------------------------------------------------------------------------------------------
    def _unwrap(self):
        return self.x + self.xcmimage * self.prd

    def _bsum(self, vals):
        out = np.zeros((self.nbody,) + vals.shape[1:])
        np.add.at(out, self.body[self.inb], vals[self.inb])
        return out

    def _remap_bodies(self):
        """Domain::remap on every body's xcm (domain.cpp:1329-1410), imagebody updated alongside"""
        for d in range(3):
            for b in range(self.nbody):
                while self.xcm[b, d] < self.lo[d]:
                    self.xcm[b, d] += self.prd[d]
                    self.imagebody[b, d] -= 1
                while self.xcm[b, d] >= self.hi[d]:
                    self.xcm[b, d] -= self.prd[d]
                    self.imagebody[b, d] += 1
                self.xcm[b, d] = max(self.xcm[b, d], self.lo[d])

    def pre_neighbor(self, image):
        """fix_rigid.cpp:1137-1175; `image` = the atoms' current true image flags [n,3]"""
        self._remap_bodies()
        self.xcmimage = np.where(self.inb[:, None], np.asarray(image, dtype=np.int64) - self.imagebody[self.body], 0)

    # ---- setup ------------------------------------------------------------------------------------------------
    def _setup_bodies_static(self, image):
        nb, m = self.nbody, self.mass[:, None]
        self.xcmimage = np.where(self.inb[:, None], image, 0)
        unw = self._unwrap()
        s = self._bsum(np.concatenate([unw * m, m], 1))
        self.masstotal = s[:, 3].copy()
        self.xcm = s[:, :3] / self.masstotal[:, None]
        self.vcm = np.zeros((nb, 3))
        self.angmom = np.zeros((nb, 3))
        self.imagebody = np.zeros((nb, 3), dtype=np.int64)
        self.pre_neighbor(image)
        d = self._unwrap() - self.xcm[self.body]
        dx, dy, dz = d[:, 0:1], d[:, 1:2], d[:, 2:3]
        s = self._bsum(np.concatenate([m * (dy * dy + dz * dz), m * (dx * dx + dz * dz), m * (dx * dx + dy * dy),
                                       -(m * dy * dz), -(m * dx * dz), -(m * dx * dy)], 1))
        self.inertia = np.zeros((nb, 3))
        self.ex, self.ey, self.ez = np.zeros((nb, 3)), np.zeros((nb, 3)), np.zeros((nb, 3))
        self.quat = np.zeros((nb, 4))
        for b in range(nb):
            t = np.array([[s[b, 0], s[b, 5], s[b, 4]], [s[b, 5], s[b, 1], s[b, 3]], [s[b, 4], s[b, 3], s[b, 2]]])
            ival, evec = jacobi3(t)
            ex, ey, ez = evec[:, 0].copy(), evec[:, 1].copy(), evec[:, 2].copy()
            mx = max(ival[0], ival[1], ival[2])
            ival = np.where(ival < EPS_STATIC * mx, 0.0, ival)
            if np.dot(np.cross(ex, ey), ez) < 0.0:
                ez = -ez
            self.inertia[b], self.ex[b], self.ey[b], self.ez[b] = ival, ex, ey, ez
            self.quat[b] = exyz_to_q(ex, ey, ez)
        delta = self._unwrap() - self.xcm[self.body]
        self.displace = np.where(self.inb[:, None],
                                 tmv_ordered(self.ex[self.body], self.ey[self.body], self.ez[self.body], delta), 0.0)
        dp = self.displace
        s = self._bsum(np.concatenate([m * (dp[:, 1:2] ** 2 + dp[:, 2:3] ** 2), m * (dp[:, 0:1] ** 2 + dp[:, 2:3] ** 2),
                                       m * (dp[:, 0:1] ** 2 + dp[:, 1:2] ** 2), -(m * dp[:, 1:2] * dp[:, 2:3]),
                                       -(m * dp[:, 0:1] * dp[:, 2:3]), -(m * dp[:, 0:1] * dp[:, 1:2])], 1))
        for b in range(nb):
            for k in range(3):
                if self.inertia[b, k] == 0.0:
                    bad = abs(s[b, k]) > TOLERANCE
                else:
                    bad = abs((s[b, k] - self.inertia[b, k]) / self.inertia[b, k]) > TOLERANCE
                if bad:
                    raise RuntimeError("Fix rigid: Bad principal moments")
            norm = self.inertia[b].sum() / 3.0
            if (np.abs(s[b, 3:6] / norm) > TOLERANCE).any():
                raise RuntimeError("Fix rigid: Bad principal moments")

    def _setup_bodies_dynamic(self):
        m = self.mass[:, None]
        d = self._unwrap() - self.xcm[self.body]
        mv = m * self.v
        s = self._bsum(np.concatenate([self.v * m,
                                       d[:, 1:2] * mv[:, 2:3] - d[:, 2:3] * mv[:, 1:2],
                                       d[:, 2:3] * mv[:, 0:1] - d[:, 0:1] * mv[:, 2:3],
                                       d[:, 0:1] * mv[:, 1:2] - d[:, 1:2] * mv[:, 0:1]], 1))
        self.vcm = s[:, :3] / self.masstotal[:, None]
        self.angmom = s[:, 3:6].copy()

    def dof(self, tgroup=None):
        """fix_rigid.cpp:1181-1262 for point particles; tgroup = boolean mask of the temperature group"""
        tg = np.ones(self.n, bool) if tgroup is None else np.asarray(tgroup, bool)
        nall = np.bincount(self.body[self.inb & tg], minlength=self.nbody)
        whole = nall == self.nrigid
        lin = (self.inertia == 0.0).any(1)
        self.nlinear = int((whole & lin).sum())
        return int((3 * nall[whole] - 6).sum() + self.nlinear)

    def _force_torque(self, f):
        d = self._unwrap() - self.xcm[self.body]
        s = self._bsum(np.concatenate([f, d[:, 1:2] * f[:, 2:3] - d[:, 2:3] * f[:, 1:2],
                                       d[:, 2:3] * f[:, 0:1] - d[:, 0:1] * f[:, 2:3],
                                       d[:, 0:1] * f[:, 1:2] - d[:, 1:2] * f[:, 0:1]], 1))
        self.fcm, self.torque = s[:, :3].copy(), s[:, 3:6].copy()

    def _t_target(self, frac):
        self.t_target = self.t_start + frac * (self.t_stop - self.t_start)

    def setup(self, f, vflag=1):
        """FixRigid::setup + FixRigidNH::setup"""
        self.f = np.asarray(f, dtype=np.float64)
        self._force_torque(self.f)
        self.evflag = vflag
        self.virial[:] = 0.0
        self.omega = angmom_to_omega(self.angmom, self.ex, self.ey, self.ez, self.inertia)
        self._set_v()
        self.virial *= 2.0
        mbody = tmv_ordered(self.ex, self.ey, self.ez, self.angmom)
        self.conjqm = 2.0 * quatvec(self.quat, mbody)
        if self.tstat:
            self.akin_t = float((self.masstotal * (self.vcm ** 2).sum(1)).sum())
            self.akin_r = float((self.angmom * self.omega).sum())
            self._t_target(0.0)
            kt = self.boltz * self.t_target
            t_mass = kt / (self.t_freq * self.t_freq)
            self.q_t[:] = t_mass
            self.q_r[:] = t_mass
            self.q_t[0], self.q_r[0] = self.nf_t * t_mass, self.nf_r * t_mass
            for i in range(1, self.t_chain):
                self.f_eta_t[i] = (self.q_t[i - 1] * self.eta_dot_t[i - 1] ** 2 - kt) / self.q_t[i]
                self.f_eta_r[i] = (self.q_r[i - 1] * self.eta_dot_r[i - 1] ** 2 - kt) / self.q_r[i]
            self.wdti1 = self.w * self.dtv / self.t_iter
            self.wdti2 = self.wdti1 / 2.0
            self.wdti4 = self.wdti1 / 4.0

    # ---- per-step ---------------------------------------------------------------------------------------------
    def _tally(self, x0, vold):
        fc = self.mass[:, None] * (self.v - vold) / self.dtf - self.f
        vr = np.stack([0.5 * x0[:, 0] * fc[:, 0], 0.5 * x0[:, 1] * fc[:, 1], 0.5 * x0[:, 2] * fc[:, 2],
                       0.5 * x0[:, 0] * fc[:, 1], 0.5 * x0[:, 0] * fc[:, 2], 0.5 * x0[:, 1] * fc[:, 2]], 1)
        for i in np.nonzero(self.inb)[0]:          # v_tally in atom order
            self.virial += vr[i]

    def _set_xv(self):
        b = self.body
        x0, vold = self._unwrap(), self.v.copy()
        xr = mv_ordered(self.ex[b], self.ey[b], self.ez[b], self.displace)
        om, vc = self.omega[b], self.vcm[b]
        vnew = np.stack([om[:, 1] * xr[:, 2] - om[:, 2] * xr[:, 1] + vc[:, 0],
                         om[:, 2] * xr[:, 0] - om[:, 0] * xr[:, 2] + vc[:, 1],
                         om[:, 0] * xr[:, 1] - om[:, 1] * xr[:, 0] + vc[:, 2]], 1)
        xnew = xr + (self.xcm[b] - self.xcmimage * self.prd)
        self.v = np.where(self.inb[:, None], vnew, self.v)
        self.x = np.where(self.inb[:, None], xnew, self.x)
        if self.evflag:
            self._tally(x0, vold)

    def _set_v(self):
        b = self.body
        vold = self.v.copy()
        dl = mv_ordered(self.ex[b], self.ey[b], self.ez[b], self.displace)
        om, vc = self.omega[b], self.vcm[b]
        vnew = np.stack([om[:, 1] * dl[:, 2] - om[:, 2] * dl[:, 1] + vc[:, 0],
                         om[:, 2] * dl[:, 0] - om[:, 0] * dl[:, 2] + vc[:, 1],
                         om[:, 0] * dl[:, 1] - om[:, 1] * dl[:, 0] + vc[:, 2]], 1)
        self.v = np.where(self.inb[:, None], vnew, self.v)
        if self.evflag:
            self._tally(self._unwrap(), vold)

    def _angmom_from_conjqm(self):
        mbody = invquatvec(self.quat, self.conjqm)
        self.angmom = 0.5 * mv_ordered(self.ex, self.ey, self.ez, mbody)
        self.omega = angmom_to_omega(self.angmom, self.ex, self.ey, self.ez, self.inertia)

    def initial_integrate(self, f, vflag=1, run_fraction=0.0):
        """fix_rigid_nh.cpp:428-603.  f = forces at the current positions (the ones the previous force call left in
        atom->f: set_xv's virial uses them); run_fraction = (ntimestep - beginstep)/(endstep - beginstep)."""
        self.f = np.asarray(f, dtype=np.float64)
        dtf2 = self.dtf * 2.0
        scale_t = scale_r = 1.0
        if self.tstat:
            scale_t = np.exp(-self.dtq * self.eta_dot_t[0])
            scale_r = np.exp(-self.dtq * self.eta_dot_r[0])
        dtfm = (self.dtf / self.masstotal)[:, None]
        self.vcm = self.vcm + dtfm * self.fcm
        if self.tstat:
            self.vcm = self.vcm * scale_t
            tmp = self.vcm[:, 0] ** 2 + self.vcm[:, 1] ** 2 + self.vcm[:, 2] ** 2
            self.akin_t = 0.0
            for b in range(self.nbody):
                self.akin_t += self.masstotal[b] * tmp[b]
        self.xcm = self.xcm + self.dtv * self.vcm
        tbody = tmv_ordered(self.ex, self.ey, self.ez, self.torque)
        fquat = quatvec(self.quat, tbody)
        self.conjqm = self.conjqm + dtf2 * fquat
        if self.tstat:
            self.conjqm = self.conjqm * scale_r
        p, q = self.conjqm, self.quat
        p, q = no_squish_rotate(3, p, q, self.inertia, self.dtq)
        p, q = no_squish_rotate(2, p, q, self.inertia, self.dtq)
        p, q = no_squish_rotate(1, p, q, self.inertia, self.dtv)
        p, q = no_squish_rotate(2, p, q, self.inertia, self.dtq)
        p, q = no_squish_rotate(3, p, q, self.inertia, self.dtq)
        self.conjqm, self.quat = p, q
        self.ex, self.ey, self.ez = q_to_exyz(q)
        self._angmom_from_conjqm()
        if self.tstat:
            ak = self.angmom[:, 0] * self.omega[:, 0] + self.angmom[:, 1] * self.omega[:, 1] + \
                self.angmom[:, 2] * self.omega[:, 2]
            self.akin_r = 0.0
            for b in range(self.nbody):
                self.akin_r += ak[b]
            self._t_target(run_fraction)
            self._nhc_temp_integrate()
        self.evflag = vflag
        if vflag:
            self.virial[:] = 0.0
        self._set_xv()

    def final_integrate(self, f):
        """fix_rigid_nh.cpp:607-790.  f = forces at the new positions"""
        self.f = np.asarray(f, dtype=np.float64)
        dtf2 = self.dtf * 2.0
        scale_t = scale_r = 1.0
        if self.tstat:
            scale_t = np.exp(-1.0 * self.dtq * self.eta_dot_t[0])
            scale_r = np.exp(-1.0 * self.dtq * self.eta_dot_r[0])
        self._force_torque(self.f)
        dtfm = (self.dtf / self.masstotal)[:, None]
        if self.tstat:
            self.vcm = self.vcm * scale_t
        self.vcm = self.vcm + dtfm * self.fcm
        tbody = tmv_ordered(self.ex, self.ey, self.ez, self.torque)
        fquat = quatvec(self.quat, tbody)
        if self.tstat:
            self.conjqm = scale_r * self.conjqm + dtf2 * fquat
        else:
            self.conjqm = self.conjqm + dtf2 * fquat
        self._angmom_from_conjqm()
        self._set_v()

    def _nhc_temp_integrate(self):
        kt = self.boltz * self.t_target
        c = self.t_chain
        t_mass = self.boltz * self.t_target / (self.t_freq * self.t_freq)
        self.q_t[:] = t_mass
        self.q_r[:] = t_mass
        self.q_t[0], self.q_r[0] = self.nf_t * t_mass, self.nf_r * t_mass
        self.f_eta_t[0] = (self.akin_t * self.mvv2e - self.nf_t * kt) / self.q_t[0]
        self.f_eta_r[0] = (self.akin_r * self.mvv2e - self.nf_r * kt) / self.q_r[0]
        for _ in range(self.t_iter):
            for j in range(self.t_order):
                w1, w2, w4 = self.wdti1[j], self.wdti2[j], self.wdti4[j]
                for ed, fe, q in ((self.eta_dot_t, self.f_eta_t, self.q_t), (self.eta_dot_r, self.f_eta_r, self.q_r)):
                    ed[c - 1] += w2 * fe[c - 1]
                    for k in range(1, c):
                        tmp = w4 * ed[c - k]
                        ms, s = maclaurin(tmp), np.exp(-1.0 * tmp)
                        ed[c - k - 1] = ed[c - k - 1] * s * s + w2 * fe[c - k - 1] * s * ms
                for e, ed in ((self.eta_t, self.eta_dot_t), (self.eta_r, self.eta_dot_r)):
                    e += w1 * ed
                for ed, fe, q in ((self.eta_dot_t, self.f_eta_t, self.q_t), (self.eta_dot_r, self.f_eta_r, self.q_r)):
                    for k in range(1, c):
                        fe[k] = (q[k - 1] * ed[k - 1] * ed[k - 1] - kt) / q[k]
                    for k in range(0, c - 1):
                        tmp = w4 * ed[k + 1]
                        ms, s = maclaurin(tmp), np.exp(-1.0 * tmp)
                        ed[k] = ed[k] * s * s + w2 * fe[k] * s * ms
                        fe[k + 1] = (q[k] * ed[k] * ed[k] - kt) / q[k + 1]
                    ed[c - 1] += w2 * fe[c - 1]

    # ---- scalars ----------------------------------------------------------------------------------------------
    def rigid_temperature(self):
        """FixRigid::compute_scalar, fix_rigid.cpp:2595-2622"""
        wb = tmv_ordered(self.ex, self.ey, self.ez, self.angmom)   # quat_to_mat columns = ex,ey,ez
        wb = np.where(self.inertia == 0.0, 0.0, wb / np.where(self.inertia == 0.0, 1.0, self.inertia))
        t = (self.masstotal * (self.vcm ** 2).sum(1)).sum() + (self.inertia * wb * wb).sum()
        return t * self.tfactor

    def compute_scalar(self):
        """FixRigidNH::compute_scalar, fix_rigid_nh.cpp:991-1016 (thermostat part)"""
        e = self.rigid_temperature()
        if self.tstat:
            kt = self.boltz * self.t_target
            e += kt * (self.nf_t * self.eta_t[0] + self.nf_r * self.eta_r[0])
            e += kt * (self.eta_t[1:] + self.eta_r[1:]).sum()
            e += (0.5 * self.q_t * self.eta_dot_t ** 2).sum() + (0.5 * self.q_r * self.eta_dot_r ** 2).sum()
        return e
