// pair_math.cuh -- per-pair arithmetic of the hot path, shared by every kernel (list mode and
// all-pairs mode) and compiled for host too, so tests/cpu_pair_math can check it without a GPU.
//
// Each function states the reference lines whose VALUES it reproduces ("pol.cpp" =
// src/pair_lj_cut_coul_long_polarization.cpp under /root/reference).  The code is organised around
// what a GPU thread needs (one neighbour, results in registers), not around the reference's loops.
#pragma once
#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD inline
#endif

namespace polb200 {

struct Box {
  double lo[3], hi[3], prd[3], half[3];
  int periodic[3];
};

// Constant parameters of one compute() call that the pair functions read.
struct PairConsts {
  double cut_coulsq;
  double f_shift;      // -1/cut_coul^2                          pol.cpp:324
  double kq;           // sqrt(qqrd2e)                           pol.cpp:367
  double qqrd2e;
  double g_ewald;
  double polar_damp;   // a
  double polar_cutsq;  // <= 0: no dipole-dipole cutoff
  double tabinnersq;
  int damping_exponential;
  int ncoultablebits, ncoulmask, ncoulshiftbits;
  int ntypes;
  int has_molecules;   // 0: no owned atom carries a molecule id, every pair is inter-molecular
  double special_lj[4], special_coul[4];
};

// 1/sqrt(x): one rsqrt on the device instead of a square root and a division
PB_HD double pb_rsqrt(double x)
{
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}

// Domain::closest_image for one coordinate (src/domain.cpp:1220-1318, orthogonal branch).
// d = xj - xi on entry; returns the wrapped displacement.  The while-loops run at most a couple of
// times because callers keep atoms within about one box length of the box.
PB_HD double wrap_delta(double d, double prd, double half, int periodic)
{
  if (periodic) {
    if (d < 0.0) {
      while (d < 0.0) d += prd;
      if (d > half) d -= prd;
    } else {
      while (d > 0.0) d -= prd;
      if (d < -half) d += prd;
    }
  }
  return d;
}

// del = xa - closest_image(xa, xb), evaluated exactly as the reference does for the pair whose
// LOWER index atom is a (pol.cpp:336-339, 437-440, 1279-1282): image = xa + wrapped(xb - xa).
PB_HD void min_image_del(const Box &b, double ax, double ay, double az, double bx, double by, double bz,
                         double &dx, double &dy, double &dz)
{
  double wx = wrap_delta(bx - ax, b.prd[0], b.half[0], b.periodic[0]);
  double wy = wrap_delta(by - ay, b.prd[1], b.half[1], b.periodic[1]);
  double wz = wrap_delta(bz - az, b.prd[2], b.half[2], b.periodic[2]);
  dx = ax - (ax + wx);
  dy = ay - (ay + wy);
  dz = az - (az + wz);
}

// rsq with the reference's operation order and NO fused multiply-add, for every comparison against
// a cutoff (neighbor build, rsq <= cut_coulsq ...): keeps pair-set decisions bit-identical.
PB_HD double rsq_nofma(double dx, double dy, double dz)
{
#if defined(__CUDA_ARCH__)
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
#else
  volatile double a = dx * dx, b = dy * dy, c = dz * dz;
  volatile double s = a + b;
  return s + c;
#endif
}

// ---- stage 2a: LJ + real-space Ewald Coulomb for one neighbour (pol.cpp:254-315) ---------------------
// One 64-byte record per table entry, {r, dr, f, df | e, de, c, dc}: the lookup of a pair is one (force only)
// or two (energy / special-bond correction) 32-byte loads from ONE line instead of six scattered 8-byte
// gathers from six arrays -- the table gathers dominated the L1 traffic of the LJ+Coulomb kernel.
struct alignas(32) CoulTabHalf {
  double a, b, c, d;
};
struct CoulTablesDev {
  const CoulTabHalf *rec;  // 2 halves per entry
};
PB_HD CoulTabHalf load_tab(const CoulTabHalf *p)
{
#if defined(__CUDA_ARCH__)
  CoulTabHalf v;  // one 256-bit read-only load
  asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p));
  return v;
#else
  return *p;
#endif
}

struct LJCoeffs {  // (ntypes+1)^2 row-major
  const double *cutsq, *cut_ljsq, *lj1, *lj2, *lj3, *lj4, *offset;
};

// returns fpair (force/r); evdwl/ecoul only meaningful when want_e
PB_HD double lj_coul_pair(const PairConsts &pc, const LJCoeffs &lj, const CoulTablesDev &tb, int ij,
                          double rsq, double qi, double qj, int sb, bool want_e, double &evdwl,
                          double &ecoul)
{
  const double EWALD_F = 1.12837917, EWALD_P = 0.3275911;
  const double A1 = 0.254829592, A2 = -0.284496736, A3 = 1.421413741, A4 = -1.453152027, A5 = 1.061405429;
  const double factor_lj = pc.special_lj[sb], factor_coul = pc.special_coul[sb];
  // one reciprocal square root feeds 1/r^2 and r (the reference divides and takes the root separately, pol.cpp:254-262;
  // the results agree to a few ulp, far inside the 1e-10 parity bar) -- a division and a square root cost ~4x more
  // FP64 issue slots each than the rsqrt sequence
  const double rinv = pb_rsqrt(rsq);
  const double r2inv = rinv * rinv;
  double forcecoul = 0.0, forcelj = 0.0, r6inv = 0.0, prefactor = 0.0;
  evdwl = ecoul = 0.0;
  if (rsq < pc.cut_coulsq) {
    if (!pc.ncoultablebits || rsq <= pc.tabinnersq) {
      const double r = rsq * rinv;
      const double grij = pc.g_ewald * r;
      const double expm2 = exp(-grij * grij);
      const double t = 1.0 / (1.0 + EWALD_P * grij);
      const double erfcv = t * (A1 + t * (A2 + t * (A3 + t * (A4 + t * A5)))) * expm2;
      prefactor = pc.qqrd2e * qi * qj * rinv;
      forcecoul = prefactor * (erfcv + EWALD_F * grij * expm2);
      if (want_e) ecoul = prefactor * erfcv;
    } else {
      // index = bits of the float32 rounding of rsq (pol.cpp:268-273)
      const float rsqf = (float)rsq;
#if defined(__CUDA_ARCH__)
      const int bits = __float_as_int(rsqf);
#else
      union { float f; int i; } u;
      u.f = rsqf;
      const int bits = u.i;
#endif
      const int itable = (bits & pc.ncoulmask) >> pc.ncoulshiftbits;
      const CoulTabHalf t0 = load_tab(tb.rec + 2 * itable);  // r, dr, f, df
      const double fraction = ((double)rsqf - t0.a) * t0.b;
      forcecoul = qi * qj * (t0.c + fraction * t0.d);
      if (factor_coul < 1.0 || want_e) {
        const CoulTabHalf t1 = load_tab(tb.rec + 2 * itable + 1);  // e, de, c, dc
        if (factor_coul < 1.0) prefactor = qi * qj * (t1.c + fraction * t1.d);
        if (want_e) ecoul = qi * qj * (t1.a + fraction * t1.b);
      }
    }
    if (factor_coul < 1.0) {
      forcecoul -= (1.0 - factor_coul) * prefactor;
      if (want_e) ecoul -= (1.0 - factor_coul) * prefactor;
    }
  }
  if (rsq < lj.cut_ljsq[ij]) {
    r6inv = r2inv * r2inv * r2inv;
    forcelj = r6inv * (lj.lj1[ij] * r6inv - lj.lj2[ij]);
    if (want_e) evdwl = factor_lj * (r6inv * (lj.lj3[ij] * r6inv - lj.lj4[ij]) - lj.offset[ij]);
  }
  return (forcecoul + factor_lj * forcelj) * r2inv;
}

// ---- stage 2b: shifted-force Coulomb field of charge qj at distance del (pol.cpp:342-357) --------------
// returns the scalar s such that E_i += s*qj*del (caller applies the sign of the pair orientation)
PB_HD double static_field_scalar(const PairConsts &pc, double rsq)
{
  const double rinv = pb_rsqrt(rsq);
  const double dvdrr = rinv * rinv + pc.f_shift;
  return dvdrr * rinv;
}

// ---- stage 3: T_ij . mu_j for one neighbour (pol.cpp:1282-1306 + 1161-1168), matrix-free ----------------
// del = xlo - image(xhi).  Accumulates  e -= T mu  component-wise.
// radial part of T in the reference's own operation order (pol.cpp:1282-1306): s1 = d1/r^3, s2 = -3 d2/r^5
PB_HD void induced_field_scalars(const PairConsts &pc, double r2, double &s1, double &s2)
{
  const double r = sqrt(r2);
  double r3, r5;
  if (r == 0.0) r3 = r5 = DBL_MAX;
  else {
    r3 = 1.0 / (r * r * r);
    r5 = 1.0 / (r * r * r * r * r);
  }
  double d1 = 1.0, d2 = 1.0;
  if (pc.damping_exponential) {
    const double a = pc.polar_damp;
    const double ex_ = exp(-a * r);
    d1 = 1.0 - ex_ * (0.5 * a * a * r2 + a * r + 1.0);
    d2 = 1.0 - ex_ * (a * a * a * r2 * r / 6.0 + 0.5 * a * a * r2 + a * r + 1.0);
  }
  // T = d1*r3*I - 3*d2*r5*(del x del);  T mu = d1*r3*mu - 3*d2*r5*(del.mu)*del
  s1 = d1 * r3;
  s2 = -3.0 * d2 * r5;
}

PB_HD void induced_field_pair(const PairConsts &pc, double dx, double dy, double dz, double r2, double mx,
                              double my, double mz, double &ex, double &ey, double &ez)
{
  double s1, s2;
  induced_field_scalars(pc, r2, s1, s2);
  const double dm = dx * mx + dy * my + dz * mz;
  ex -= s1 * mx + s2 * dm * dx;
  ey -= s1 * my + s2 * dm * dy;
  ez -= s1 * mz + s2 * dm * dz;
}

// ---- stage 4: polarization force on the LOWER-index atom of a pair (pol.cpp:441-602) -----------------
// a = lower index atom ("i" of the reference loop), b = higher ("j").  del = xa - image(xb).
// Outputs the force on a (the force on b is its negative) and the pair's energy terms.
struct PolPairIn {
  double dx, dy, dz;
  double qa, qb, alpha_a, alpha_b;
  double max_, may, maz, mbx, mby, mbz;  // dipoles of a and b
  bool intermolecular;                   // (molecule[a]!=molecule[b]) || molecule[a]==0
};

PB_HD void pol_force_pair(const PairConsts &pc, const PolPairIn &in, bool want_e, double &fx, double &fy,
                          double &fz, double &u_ef, double &u_dd)
{
  const double delx = in.dx, dely = in.dy, delz = in.dz;
  const double xsq = delx * delx, ysq = dely * dely, zsq = delz * delz;
  const double rsq = xsq + ysq + zsq;
  // one reciprocal square root feeds every inverse power (the reference divides and takes roots separately,
  // pol.cpp:441-452; the results agree to a few ulp, far inside the 1e-10 parity bar)
  const double rinv = pb_rsqrt(rsq);
  const double r2inv = rinv * rinv;
  const double r = rsq * rinv;
  const double r3inv = r2inv * rinv;
  const double f_shift = pc.f_shift, kq = pc.kq;
  fx = fy = fz = 0.0;
  u_ef = u_dd = 0.0;

  if (rsq < pc.cut_coulsq && in.intermolecular) {
    const double dvdrr = r2inv + f_shift;
    const double ef_temp = dvdrr * rinv * kq;
    // M (symmetric) applied to a dipole: the bracketed factors of pol.cpp:467-475
    const double mxx = (-2.0 * xsq + ysq + zsq) * r2inv + f_shift * (ysq + zsq);
    const double myy = (-2.0 * ysq + xsq + zsq) * r2inv + f_shift * (xsq + zsq);
    const double mzz = (-2.0 * zsq + xsq + ysq) * r2inv + f_shift * (xsq + ysq);
    const double mxy = -3.0 * delx * dely * r2inv - f_shift * delx * dely;
    const double mxz = -3.0 * delx * delz * r2inv - f_shift * delx * delz;
    const double myz = -3.0 * dely * delz * r2inv - f_shift * dely * delz;
    if (in.alpha_a != 0.0 && in.qb != 0.0) {  // dipole on a, charge on b (pol.cpp:464-484)
      const double cf = in.qb * kq * r3inv;
      fx += cf * (in.max_ * mxx + in.may * mxy + in.maz * mxz);
      fy += cf * (in.max_ * mxy + in.may * myy + in.maz * myz);
      fz += cf * (in.max_ * mxz + in.may * myz + in.maz * mzz);
      if (want_e)
        u_ef -= in.max_ * (ef_temp * in.qb * delx) + in.may * (ef_temp * in.qb * dely) +
                in.maz * (ef_temp * in.qb * delz);
    }
    if (in.alpha_b != 0.0 && in.qa != 0.0) {  // dipole on b, charge on a (pol.cpp:487-507)
      const double cf = in.qa * kq * r3inv;
      fx -= cf * (in.mbx * mxx + in.mby * mxy + in.mbz * mxz);
      fy -= cf * (in.mbx * mxy + in.mby * myy + in.mbz * myz);
      fz -= cf * (in.mbx * mxz + in.mby * myz + in.mbz * mzz);
      if (want_e)
        u_ef += in.mbx * (ef_temp * in.qa * delx) + in.mby * (ef_temp * in.qa * dely) +
                in.mbz * (ef_temp * in.qa * delz);
    }
  }

  const bool in_polar_cut = !(pc.polar_cutsq > 0.0) || rsq < pc.polar_cutsq;
  if (in.alpha_a != 0.0 && in.alpha_b != 0.0 && in_polar_cut) {  // pol.cpp:512-602
    const double r5inv = r3inv * r2inv;
    const double r7inv = r5inv * r2inv;
    const double pdotp = in.max_ * in.mbx + in.may * in.mby + in.maz * in.mbz;
    const double pidotr = in.max_ * delx + in.may * dely + in.maz * delz;
    const double pjdotr = in.mbx * delx + in.mby * dely + in.mbz * delz;
    double pre1, pre2, pre3, pre45 = 0.0;
    if (pc.damping_exponential) {
      const double a = pc.polar_damp;
      const double t1 = exp(-a * r);
      const double t2 = 1.0 + a * r + 0.5 * a * a * r * r;
      const double t3 = t2 + 1.0 / 6.0 * a * a * a * r * r * r;
      const double w2 = 1.0 - t1 * t2, w3 = 1.0 - t1 * t3;
      pre1 = 3.0 * r5inv * pdotp * w2 - 15.0 * r7inv * pidotr * pjdotr * w3;
      pre2 = 3.0 * r5inv * pjdotr * w3;
      pre3 = 3.0 * r5inv * pidotr * w3;
      const double pre4 = -pdotp * r3inv * (-t1 * (a * rinv + a * a) + t1 * a * t2 * rinv);
      const double pre5 = 3.0 * pidotr * pjdotr * r5inv *
                          (-t1 * (a * rinv + a * a + 0.5 * r * a * a * a) + t1 * a * t3 * rinv);
      pre45 = pre4 + pre5;
      if (want_e) u_dd = r3inv * pdotp * w2 - 3.0 * r5inv * pidotr * pjdotr * w3;
    } else {
      pre1 = 3.0 * r5inv * pdotp - 15.0 * r7inv * pidotr * pjdotr;
      pre2 = 3.0 * r5inv * pjdotr;
      pre3 = 3.0 * r5inv * pidotr;
      if (want_e) u_dd = r3inv * pdotp - 3.0 * r5inv * pidotr * pjdotr;
    }
    fx += (pre1 + pre45) * delx + pre2 * in.max_ + pre3 * in.mbx;
    fy += (pre1 + pre45) * dely + pre2 * in.may + pre3 * in.mby;
    fz += (pre1 + pre45) * delz + pre2 * in.maz + pre3 * in.mbz;
  }
}

// Same pair, regrouped for the list-mode force kernel on the pair-group rows (the row atom is a, del = xa - xb).
// Identities used (exact algebra, different rounding -- agreement with pol_force_pair to a few ulp per term):
//   * the charge-dipole matrix of pol.cpp:467-475 is  M = c1 I - c2 del (x) del  with  c1 = 1 + f_shift r^2,
//     c2 = 3/r^2 + f_shift, so both charge-dipole forces are  kq/r^3 [c1 v - c2 s del]  with
//     v = qb mu_a - qa mu_b,  s = qb (del.mu_a) - qa (del.mu_b), and  u_ef = -ef_temp s;
//   * the derivative terms of the exponential damping collapse:  pre4 + pre5 (pol.cpp:528-533)
//     = (a^3/2) e^{-ar} [a (del.mu_a)(del.mu_b)/r^3 - (mu_a.mu_b)/r^2].
// ~100 FP64 operations per pair instead of ~190: the force kernel is bound by the FP64 pipe.
PB_HD void pol_force_pair_fast(const PairConsts &pc, bool damp, double dx, double dy, double dz, double qa, double qb,
                               double alpha_a, double alpha_b, double max_, double may, double maz, double mbx, double mby,
                               double mbz, bool intermolecular, bool want_e, double &fx, double &fy, double &fz,
                               double &u_ef, double &u_dd)
{
  const double rsq = dx * dx + dy * dy + dz * dz;
  const double rinv = pb_rsqrt(rsq);
  const double r2inv = rinv * rinv;
  const double r3inv = r2inv * rinv;
  const double pa = max_ * dx + may * dy + maz * dz;
  const double pb = mbx * dx + mby * dy + mbz * dz;
  fx = fy = fz = 0.0;
  u_ef = u_dd = 0.0;
  if (rsq < pc.cut_coulsq && intermolecular) {
    const double wa = alpha_a != 0.0 ? qb : 0.0;  // dipole on a feels the charge on b (pol.cpp:464)
    const double wb = alpha_b != 0.0 ? qa : 0.0;  // dipole on b feels the charge on a (pol.cpp:487)
    const double c1 = 1.0 + pc.f_shift * rsq, c2 = 3.0 * r2inv + pc.f_shift;
    const double s = wa * pa - wb * pb;
    const double k3 = pc.kq * r3inv, k3s = k3 * c2 * s, k3c = k3 * c1;
    fx = k3c * (wa * max_ - wb * mbx) - k3s * dx;
    fy = k3c * (wa * may - wb * mby) - k3s * dy;
    fz = k3c * (wa * maz - wb * mbz) - k3s * dz;
    if (want_e) u_ef = -(r2inv + pc.f_shift) * rinv * pc.kq * s;
  }
  const bool in_polar_cut = !(pc.polar_cutsq > 0.0) || rsq < pc.polar_cutsq;
  if (alpha_a != 0.0 && alpha_b != 0.0 && in_polar_cut) {  // pol.cpp:512-602
    const double pdotp = max_ * mbx + may * mby + maz * mbz;
    const double r5inv3 = 3.0 * r3inv * r2inv;
    const double papb = pa * pb;
    double A, B;
    if (damp) {
      const double a = pc.polar_damp;
      const double ar = a * (rsq * rinv);
      const double t1 = exp(-ar);
      const double t2 = 1.0 + ar + 0.5 * ar * ar;
      const double t3 = t2 + (1.0 / 6.0) * ar * ar * ar;
      const double w2 = 1.0 - t1 * t2, w3 = 1.0 - t1 * t3;
      const double g = 0.5 * t1 * a * a * a;
      A = r5inv3 * (w2 * pdotp - 5.0 * r2inv * w3 * papb) + g * (a * papb * r3inv - pdotp * r2inv);
      B = r5inv3 * w3;
      if (want_e) u_dd = r3inv * w2 * pdotp - B * papb;
    } else {
      A = r5inv3 * (pdotp - 5.0 * r2inv * papb);
      B = r5inv3;
      if (want_e) u_dd = r3inv * pdotp - B * papb;
    }
    fx += A * dx + B * (pb * max_ + pa * mbx);
    fy += A * dy + B * (pb * may + pa * mby);
    fz += A * dz + B * (pb * maz + pa * mbz);
  }
}

}  // namespace polb200
