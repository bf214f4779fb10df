"""ORACLE (test infrastructure, never imported by the product): numpy restatement of the reference's
`kspace_style pppm <accuracy>` (src/KSPACE/pppm.cpp) -- ik differentiation, no stagger, orthogonal periodic box, one
process.  Stage 1 of the PPPM widening (second half): the device Ewald sum is O(N^1.5) and already
costs more than the pair style at 256k atoms; PPPM is what the reference itself would use there.

Follows, under /root/reference/src/KSPACE/pppm.cpp:
  :985-1135   set_grid_global    g_ewald estimate, grid from estimate_ik_error, factorable (2, 3, 5)
  :1161-1181  compute_df_kspace  :1270-1281 estimate_ik_error (acons table :129-161)
  :1287-1340  adjust_gewald / newton_raphson_f / derivf
  :1370-1395  set_grid_local     nlower, nupper, shift, shiftone
  :400-495    setup              fkx/fky/fkz, virial coefficients vg
  :1526-1544  compute_gf_denom   pppm.h:185-196 gf_denom
  :1549-1627  compute_gf_ik      Hockney-Eastwood optimal influence function
  :2908-2952  compute_rho_coeff  :2844-2863 compute_rho1d
  :1907-1945  particle_map       :1951-1995 make_rho
  :2032-2157  poisson_ik         :2453-2505 fieldforce_ik
  :622-765    compute            energy (self and neutralising terms), virial
  ../math_special.h:82-93 powsinxx

Parity: PINNED against tests/golden/pppm_*.npz (E_long, per-atom KSpace forces and virial dumped from the reference
binary by oracle/make_golden.py pppm; the reference uses its bundled KISS FFT, numpy's pocketfft agrees to rounding):
tests/test_pppm_oracle.py.
"""
import math

import numpy as np

OFFSET = 16384
EPS_HOC = 1.0e-7
MY_PIS = 1.77245385090551602729
MY_PI2 = 1.57079632679489661923

ACONS = {1: [2.0 / 3.0],
         2: [1.0 / 50.0, 5.0 / 294.0],
         3: [1.0 / 588.0, 7.0 / 1440.0, 21.0 / 3872.0],
         4: [1.0 / 4320.0, 3.0 / 1936.0, 7601.0 / 2271360.0, 143.0 / 28800.0],
         5: [1.0 / 23232.0, 7601.0 / 13628160.0, 143.0 / 69120.0, 517231.0 / 106536960.0, 106640677.0 / 11737571328.0],
         6: [691.0 / 68140800.0, 13.0 / 57600.0, 47021.0 / 35512320.0, 9694607.0 / 2095994880.0,
             733191589.0 / 59609088000.0, 326190917.0 / 11700633600.0],
         7: [1.0 / 345600.0, 3617.0 / 35512320.0, 745739.0 / 838397952.0, 56399353.0 / 12773376000.0,
             25091609.0 / 1560084480.0, 1755948832039.0 / 36229939200000.0, 4887769399.0 / 37838389248.0]}


def factorable(n):
    while n > 1:
        for f in (2, 3, 5):
            if n % f == 0:
                n //= f
                break
        else:
            return False
    return True


def powsinxx(x, n):
    """(sin x / x)^n by repeated squaring, math_special.h:82-93 (vectorised)"""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        ww = np.where(x == 0.0, 1.0, np.sin(x) / np.where(x == 0.0, 1.0, x))
    yy = np.ones_like(ww)
    while n:
        if n & 1:
            yy = yy * ww
        n >>= 1
        ww = ww * ww
    return np.where(x == 0.0, 1.0, yy)


class PPPMPlan:
    """PPPM::init + setup for one box / charge set"""

    def __init__(self, accuracy_relative, q, cutoff, prd, order=5, qqrd2e=332.06371, two_charge_force=332.06371,
                 g_ewald=None, mesh=None):
        q = np.asarray(q, dtype=np.float64)
        self.order, self.qqrd2e, self.cutoff = order, qqrd2e, cutoff
        self.prd = np.asarray(prd, dtype=np.float64)
        self.natoms = len(q)
        self.qsum = float(np.cumsum(q)[-1])            # KSpace::qsum_qsq sums in atom order
        self.qsqsum = float(np.cumsum(q * q)[-1])
        self.q2 = self.qsqsum * qqrd2e
        self.accuracy = accuracy_relative * two_charge_force
        xprd, yprd, zprd = self.prd
        # ---- set_grid_global
        if g_ewald is None:
            g = self.accuracy * math.sqrt(self.natoms * cutoff * xprd * yprd * zprd) / (2.0 * self.q2)
            g = (1.35 - 0.15 * math.log(self.accuracy)) / cutoff if g >= 1.0 else math.sqrt(-math.log(g)) / cutoff
            self.g_ewald = g
        else:
            self.g_ewald = g_ewald
        if mesh is None:
            n = []
            for d in range(3):
                h = 1.0 / self.g_ewald
                nd = int(self.prd[d] / h) + 1
                err = self.estimate_ik_error(h, self.prd[d])
                while err > self.accuracy:
                    err = self.estimate_ik_error(h, self.prd[d])
                    nd += 1
                    h = self.prd[d] / nd
                n.append(nd)
        else:
            n = list(mesh)
        for d in range(3):
            while not factorable(n[d]):
                n[d] += 1
        self.n = np.array(n, dtype=np.int64)
        self.h = self.prd / self.n
        # ---- set_grid_local
        self.nlower, self.nupper = -((order - 1) // 2), order // 2
        self.shift = OFFSET + 0.5 if order % 2 else float(OFFSET)
        self.shiftone = 0.0 if order % 2 else 0.5
        if g_ewald is None:
            self.adjust_gewald()
        self.setup()
        self.compute_rho_coeff()

    # -- error estimates ----------------------------------------------------------------------------------------
    def estimate_ik_error(self, h, prd):
        s = 0.0
        for m in range(self.order):
            s += ACONS[self.order][m] * (h * self.g_ewald) ** (2.0 * m)
        return self.q2 * (h * self.g_ewald) ** float(self.order) * \
            math.sqrt(self.g_ewald * prd * math.sqrt(2.0 * math.pi) * s / self.natoms) / (prd * prd)

    def compute_df_kspace(self):
        l = [self.estimate_ik_error(self.h[d], self.prd[d]) for d in range(3)]
        return math.sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]) / math.sqrt(3.0)

    def newton_raphson_f(self):
        xprd, yprd, zprd = self.prd
        df_r = 2.0 * self.q2 * math.exp(-self.g_ewald * self.g_ewald * self.cutoff * self.cutoff) / \
            math.sqrt(self.natoms * self.cutoff * xprd * yprd * zprd)
        return df_r - self.compute_df_kspace()

    def adjust_gewald(self):
        for _ in range(10000):
            f1 = self.newton_raphson_f()
            g_old = self.g_ewald
            self.g_ewald = g_old + 0.000001
            f2 = self.newton_raphson_f()
            self.g_ewald = g_old
            df = (f2 - f1) / 0.000001
            self.g_ewald -= self.newton_raphson_f() / df
            if abs(self.newton_raphson_f()) < 0.00001:
                return
        raise RuntimeError("Could not compute g_ewald")

    # -- setup --------------------------------------------------------------------------------------------------
    def setup(self):
        order, g = self.order, self.g_ewald
        self.volume = float(self.prd[0] * self.prd[1] * self.prd[2])
        self.delinv = self.n / self.prd
        self.delvolinv = float(self.delinv[0] * self.delinv[1] * self.delinv[2])
        unitk = 2.0 * math.pi / self.prd
        per = [np.arange(nd) - nd * (2 * np.arange(nd) // nd) for nd in self.n]
        self.fk = [unitk[d] * per[d] for d in range(3)]
        # compute_gf_denom
        b = np.zeros(order)
        b[0] = 1.0
        for m in range(1, order):
            for l in range(m, 0, -1):
                b[l] = 4.0 * (b[l] * (l - m) * (l - m - 0.5) - b[l - 1] * (l - m - 1) * (l - m - 1))
            b[0] = 4.0 * (b[0] * (0 - m) * (0 - m - 0.5))    # the loop above leaves l = 0
        ifact = 1
        for k in range(1, 2 * order):
            ifact *= k
        self.gf_b = b * (1.0 / ifact)
        # compute_gf_ik: arrays indexed [m(z), l(y), k(x)] like the reference's flat n = (m*ny + l)*nx + k
        nb = [int((g * self.prd[d] / (math.pi * self.n[d])) * (-math.log(EPS_HOC)) ** 0.25) for d in range(3)]
        kx, ly, mz = per[0][None, None, :], per[1][None, :, None], per[2][:, None, None]
        sn = [np.sin(0.5 * unitk[d] * per[d] * self.prd[d] / self.n[d]) ** 2 for d in range(3)]

        def poly(x):
            s = np.zeros_like(x)
            for l in range(order - 1, -1, -1):
                s = self.gf_b[l] + s * x
            return s
        den = (poly(sn[0])[None, None, :] * poly(sn[1])[None, :, None] * poly(sn[2])[:, None, None]) ** 2
        sqk = (unitk[0] * kx) ** 2 + (unitk[1] * ly) ** 2 + (unitk[2] * mz) ** 2
        sum1 = np.zeros(sqk.shape)
        twoorder = 2 * order
        for nx in range(-nb[0], nb[0] + 1):
            qx = unitk[0] * (kx + self.n[0] * nx)
            sx = np.exp(-0.25 * (qx / g) ** 2)
            wx = powsinxx(0.5 * qx * self.prd[0] / self.n[0], twoorder)
            for ny in range(-nb[1], nb[1] + 1):
                qy = unitk[1] * (ly + self.n[1] * ny)
                sy = np.exp(-0.25 * (qy / g) ** 2)
                wy = powsinxx(0.5 * qy * self.prd[1] / self.n[1], twoorder)
                for nz in range(-nb[2], nb[2] + 1):
                    qz = unitk[2] * (mz + self.n[2] * nz)
                    sz = np.exp(-0.25 * (qz / g) ** 2)
                    wz = powsinxx(0.5 * qz * self.prd[2] / self.n[2], twoorder)
                    dot1 = unitk[0] * kx * qx + unitk[1] * ly * qy + unitk[2] * mz * qz
                    dot2 = qx * qx + qy * qy + qz * qz
                    with np.errstate(invalid="ignore", divide="ignore"):
                        sum1 = sum1 + (dot1 / dot2) * sx * sy * sz * wx * wy * wz
        with np.errstate(invalid="ignore", divide="ignore"):
            self.greensfn = np.where(sqk != 0.0, (12.5663706 / sqk) * sum1 / den, 0.0)
        # virial coefficients (setup, :455-480)
        fx, fy, fz = self.fk[0][None, None, :], self.fk[1][None, :, None], self.fk[2][:, None, None]
        with np.errstate(invalid="ignore", divide="ignore"):
            vterm = np.where(sqk != 0.0, -2.0 * (1.0 / sqk + 0.25 / (g * g)), 0.0)
        nz0 = sqk != 0.0
        one = np.where(nz0, 1.0, 0.0)
        self.vg = [one + vterm * fx * fx, one + vterm * fy * fy, one + vterm * fz * fz,
                   vterm * fx * fy * np.ones_like(sqk), vterm * fx * fz * np.ones_like(sqk), vterm * fy * fz * np.ones_like(sqk)]

    def compute_rho_coeff(self):
        order = self.order
        a = np.zeros((order, 2 * order + 1))          # a[l][k + order]
        a[0][order] = 1.0
        for j in range(1, order):
            for k in range(-j, j + 1, 2):
                s = 0.0
                for l in range(j):
                    a[l + 1][k + order] = (a[l][k + 1 + order] - a[l][k - 1 + order]) / (l + 1)
                    s += 0.5 ** (l + 1) * (a[l][k - 1 + order] + (-1.0) ** l * a[l][k + 1 + order]) / (l + 1)
                a[0][k + order] = s
        self.rho_coeff = np.zeros((order, order))     # [l][m - nlower]
        m = 0
        for k in range(-(order - 1), order, 2):
            self.rho_coeff[:, m] = a[:, k + order]
            m += 1

    # -- compute ------------------------------------------------------------------------------------------------
    def rho1d(self, d):
        """d[n] -> weights [n, order] (Horner in the reference's order)"""
        r = np.zeros((len(d), self.order))
        for l in range(self.order - 1, -1, -1):
            r = self.rho_coeff[l][None, :] + r * d[:, None]
        return r

    def compute(self, x, q, boxlo):
        """PPPM::compute: returns dict(energy, f[n,3], virial[6])"""
        x = np.asarray(x, dtype=np.float64)
        q = np.asarray(q, dtype=np.float64)
        n = len(q)
        nx, ny, nz = (int(v) for v in self.n)
        part = np.zeros((n, 3), dtype=np.int64)
        w = []
        for d in range(3):
            part[:, d] = ((x[:, d] - boxlo[d]) * self.delinv[d] + self.shift).astype(np.int64) - OFFSET
            w.append(self.rho1d(part[:, d] + self.shiftone - (x[:, d] - boxlo[d]) * self.delinv[d]))
        offs = np.arange(self.nlower, self.nupper + 1)
        ix = (part[:, 0:1] + offs[None, :]) % nx      # periodic wrap = what the ghost-cell reverse_comm folds
        iy = (part[:, 1:2] + offs[None, :]) % ny
        iz = (part[:, 2:3] + offs[None, :]) % nz
        # make_rho: density[mz][my][mx] += z0*w2[n] * w1[m] * w0[l], atoms in order
        dens = np.zeros((nz, ny, nx))
        z0 = self.delvolinv * q
        for i in range(n):
            contrib = (z0[i] * w[2][i])[:, None, None] * w[1][i][None, :, None] * w[0][i][None, None, :]
            np.add.at(dens, (iz[i][:, None, None], iy[i][None, :, None], ix[i][None, None, :]), contrib)
        # poisson_ik
        # fft1->compute(work1,work1,1): flag 1 selects the e^{+ikr} transform, unscaled (fft3d.cpp:103-123, :606-607)
        work1 = np.fft.ifftn(dens) * (nx * ny * nz)
        scaleinv = 1.0 / (nx * ny * nz)
        eng = scaleinv * scaleinv * self.greensfn * (work1.real ** 2 + work1.imag ** 2)
        energy = float(eng.sum())
        virial = np.array([float((eng * self.vg[j]).sum()) for j in range(6)])
        work1 = work1 * (scaleinv * self.greensfn)
        fkx, fky, fkz = self.fk[0][None, None, :], self.fk[1][None, :, None], self.fk[2][:, None, None]
        # work2 = (fk*Im, -fk*Re) = -i*fk*work1; fft2->compute(work2,work2,-1): the e^{-ikr} transform, unscaled
        vd = []
        for fk in (fkx, fky, fkz):
            w2 = (fk * work1.imag) + 1j * (-fk * work1.real)
            vd.append(np.fft.fftn(w2).real)
        # fieldforce_ik
        f = np.zeros((n, 3))
        for i in range(n):
            x0 = w[2][i][:, None, None] * w[1][i][None, :, None] * w[0][i][None, None, :]
            idx = (iz[i][:, None, None], iy[i][None, :, None], ix[i][None, None, :])
            qf = self.qqrd2e * q[i]
            for d in range(3):
                f[i, d] = qf * (-(x0 * vd[d][idx]).sum())
        energy *= 0.5 * self.volume
        energy -= self.g_ewald * self.qsqsum / MY_PIS + MY_PI2 * self.qsum * self.qsum / (self.g_ewald * self.g_ewald * self.volume)
        energy *= self.qqrd2e
        virial = 0.5 * self.qqrd2e * self.volume * virial
        return dict(energy=energy, f=f, virial=virial)
