// rigid.cu -- rigid-body integrator on the device behind polb200_rigid_* of include/polb200.h:
// `fix rigid/nve molecule` and `fix rigid/nvt molecule` of the reference (src/RIGID/fix_rigid_nh.cpp on top of
// fix_rigid.cpp), point particles in an orthogonal periodic box.
//
// File map
//   records        BodyFrame (what the per-atom kernels gather: xcm, vcm, omega, principal axes; 160 B) and BodyDyn
//                  (what only the per-body kernels touch: fcm, torque, angmom, quat, conjqm, inertia, mass; 192 B);
//                  per-atom data keyed by atom id: AtomRec {displace, mass}, body index, body-relative image flags
//   k_rigid_index      inverse of the caller's tag array (per call: the caller may have re-sorted its atoms)
//   k_rigid_bodies<>   per body: sum force and torque over the members in id order (thread or warp per body), then
//                      MODE_SETUP (FixRigid::setup + conjqm), MODE_FINAL (second half kick) -- one kernel, no atomics
//   k_rigid_initial    per body: half kick, drift, five no_squish rotations, new axes / angmom / omega
//   k_rigid_nhc        one block: fixed-order sum of the kinetic terms + Nose-Hoover chain update (thermostat state
//                      never leaves the device, so a step needs no host synchronisation)
//   k_rigid_atoms<>    per atom: set_xv / set_v incl. the constraint virial (block partials, fixed-order final sum)
//   k_rigid_remap, k_rigid_image_shift   pre_neighbor
//   host: bodies_static / bodies_dynamic (once per init, Jacobi diagonalisation), the C ABI
// Everything is deterministic: sums run in a fixed order, there are no atomics.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "comm_types.h"
#include "devbuf.h"
#include "host_style.h"
#include "polb200.h"

namespace polb200 {

// LAMMPS image flags (src/lmptype.h:96-103, default 32-bit imageint)
constexpr int IMGMASK = 1023, IMGMAX = 512, IMGBITS = 10, IMG2BITS = 20;

struct BodyFrame {  // 20 doubles
  double xcm[3], vcm[3], omega[3], ex[3], ey[3], ez[3], pad[2];
};
struct BodyDyn {  // 24 doubles
  double fcm[3], torque[3], angmom[3], quat[4], conjqm[4], inertia[3], masstotal, pad[3];
};
struct Chain {  // Nose-Hoover chains of FixRigidNH (translation / rotation); MAXCHAIN bounds t_chain
  static constexpr int MAXCHAIN = 64;
  double eta_t[MAXCHAIN], eta_r[MAXCHAIN], eta_dot_t[MAXCHAIN], eta_dot_r[MAXCHAIN];
  double f_eta_t[MAXCHAIN], f_eta_r[MAXCHAIN], q_t[MAXCHAIN], q_r[MAXCHAIN];
  double akin_t, akin_r, t_target;
};
struct RigidConst {
  double dtv, dtf, dtq, mvv2e, boltz, t_freq;
  double prd[3], lo[3], hi[3];
  double wdti1[5], wdti2[5], wdti4[5];
  int nf_t, nf_r, t_chain, t_iter, t_order, tstat;
};

__device__ __forceinline__ void unpack_image(int img, int &xb, int &yb, int &zb)
{
  xb = (img & IMGMASK) - IMGMAX;
  yb = ((img >> IMGBITS) & IMGMASK) - IMGMAX;
  zb = (img >> IMG2BITS) - IMGMAX;  // arithmetic shift of a non-negative 30-bit value
}

__host__ __device__ inline void tmatvec(const double *ex, const double *ey, const double *ez, const double *v, double *o)
{
  o[0] = ex[0] * v[0] + ex[1] * v[1] + ex[2] * v[2];
  o[1] = ey[0] * v[0] + ey[1] * v[1] + ey[2] * v[2];
  o[2] = ez[0] * v[0] + ez[1] * v[1] + ez[2] * v[2];
}
__host__ __device__ inline void matvec3(const double *ex, const double *ey, const double *ez, const double *v, double *o)
{
  o[0] = ex[0] * v[0] + ey[0] * v[1] + ez[0] * v[2];
  o[1] = ex[1] * v[0] + ey[1] * v[1] + ez[1] * v[2];
  o[2] = ex[2] * v[0] + ey[2] * v[1] + ez[2] * v[2];
}
__host__ __device__ inline void quatvec(const double *a, const double *b, double *c)  // math_extra.h:609-615
{
  c[0] = -a[1] * b[0] - a[2] * b[1] - a[3] * b[2];
  c[1] = a[0] * b[0] + a[2] * b[2] - a[3] * b[1];
  c[2] = a[0] * b[1] + a[3] * b[0] - a[1] * b[2];
  c[3] = a[0] * b[2] + a[1] * b[1] - a[2] * b[0];
}
__host__ __device__ inline void invquatvec(const double *a, const double *b, double *c)  // math_extra.h:636-641
{
  c[0] = -a[1] * b[0] + a[0] * b[1] + a[3] * b[2] - a[2] * b[3];
  c[1] = -a[2] * b[0] - a[3] * b[1] + a[0] * b[2] + a[1] * b[3];
  c[2] = -a[3] * b[0] + a[2] * b[1] - a[1] * b[2] + a[0] * b[3];
}
__host__ __device__ inline void angmom_to_omega(const double *m, const double *ex, const double *ey, const double *ez,
                                                const double *idiag, double *w)  // math_extra.cpp:290-305
{
  double wb[3];
  tmatvec(ex, ey, ez, m, wb);
  for (int k = 0; k < 3; k++) wb[k] = (idiag[k] == 0.0) ? 0.0 : wb[k] / idiag[k];
  matvec3(ex, ey, ez, wb, w);
}
__host__ __device__ inline void q_to_exyz(const double *q, double *ex, double *ey, double *ez)  // math_extra.cpp:402-415
{
  ex[0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3];
  ex[1] = 2.0 * (q[1] * q[2] + q[0] * q[3]);
  ex[2] = 2.0 * (q[1] * q[3] - q[0] * q[2]);
  ey[0] = 2.0 * (q[1] * q[2] - q[0] * q[3]);
  ey[1] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3];
  ey[2] = 2.0 * (q[2] * q[3] + q[0] * q[1]);
  ez[0] = 2.0 * (q[1] * q[3] + q[0] * q[2]);
  ez[1] = 2.0 * (q[2] * q[3] - q[0] * q[1]);
  ez[2] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
}

// MathExtra::no_squish_rotate (math_extra.cpp:234-277): one symplectic sub-rotation about principal axis k
__device__ inline void no_squish_rotate(int k, double *p, double *q, const double *inertia, double dt)
{
  double kq[4], kp[4];
  if (k == 1) {
    kq[0] = -q[1]; kp[0] = -p[1]; kq[1] = q[0]; kp[1] = p[0]; kq[2] = q[3]; kp[2] = p[3]; kq[3] = -q[2]; kp[3] = -p[2];
  } else if (k == 2) {
    kq[0] = -q[2]; kp[0] = -p[2]; kq[1] = -q[3]; kp[1] = -p[3]; kq[2] = q[0]; kp[2] = p[0]; kq[3] = q[1]; kp[3] = p[1];
  } else {
    kq[0] = -q[3]; kp[0] = -p[3]; kq[1] = q[2]; kp[1] = p[2]; kq[2] = -q[1]; kp[2] = -p[1]; kq[3] = q[0]; kp[3] = p[0];
  }
  double phi = p[0] * kq[0] + p[1] * kq[1] + p[2] * kq[2] + p[3] * kq[3];
  if (fabs(inertia[k - 1]) < 1e-6) phi *= 0.0;
  else phi /= 4.0 * inertia[k - 1];
  double s, c;
  sincos(dt * phi, &s, &c);
  for (int i = 0; i < 4; i++) {
    p[i] = c * p[i] + s * kp[i];
    q[i] = c * q[i] + s * kq[i];
  }
}

__device__ inline double maclaurin_series(double x)  // fix_rigid_nh.h:91-98
{
  const double x2 = x * x, x4 = x2 * x2;
  return (1.0 + (1.0 / 6.0) * x2 + (1.0 / 120.0) * x4 + (1.0 / 5040.0) * x2 * x4 + (1.0 / 362880.0) * x4 * x4);
}

__global__ void k_rigid_index(int n, const int *__restrict__ tag, int *__restrict__ idx_of_tag)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx_of_tag[tag[i] - 1] = i;
}

enum { MODE_SETUP = 0, MODE_FINAL = 1 };

// Force and torque on every body from its members (fix_rigid.cpp:782-855, fix_rigid_nh.cpp:646-700), then
//   MODE_SETUP: omega from angmom, conjqm = 2 q (P^T angmom)   (fix_rigid.cpp:871-875, fix_rigid_nh.cpp:327-337)
//   MODE_FINAL: second half kick of vcm and conjqm, angmom, omega   (fix_rigid_nh.cpp:706-764)
// LANES = 1: a thread per body (molecules of a few atoms); LANES = 32: a warp per body, members strided over the
// lanes and combined by an xor butterfly -- the same order on every run.
// STAGE = 0: sum and update in one launch (every member is on this process).
// More than one process (polb200_rigid_comm_init; the reference's MPI_Allreduce of sum[nbody][6], fix_rigid.cpp:826-827,
// fix_rigid_nh.cpp:668-669): STAGE = 1 writes this process's partial sums over the members it owns to `sum6` and stops;
// after the all-reduce STAGE = 2 (LANES = 1) reads the complete sums and updates the replicated body.
template <int LANES, int MODE, int STAGE = 0>
__global__ void __launch_bounds__(256) k_rigid_bodies(int nbody, RigidConst rc, const Chain *__restrict__ chain,
                                                     const int *__restrict__ member_first, const int *__restrict__ member_tag,
                                                     const int *__restrict__ idx_of_tag, const int *__restrict__ xcmimage,
                                                     const double *__restrict__ x, const double *__restrict__ f,
                                                     BodyFrame *__restrict__ frame, BodyDyn *__restrict__ dyn,
                                                     double2 *__restrict__ akin, double *__restrict__ sum6)
{
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = gid / LANES, lane = gid % LANES;
  if (b >= nbody) return;
  BodyFrame F = frame[b];
  double s[6] = {0, 0, 0, 0, 0, 0};
  if (STAGE == 2)
    for (int k = 0; k < 6; k++) s[k] = sum6[(size_t)6 * b + k];
  for (int m = member_first[b] + lane; STAGE != 2 && m < member_first[b + 1]; m += LANES) {
    const int t = member_tag[m], i = idx_of_tag[t];
    if (STAGE == 1 && i < 0) continue;  // owned by another process
    int xb, yb, zb;
    unpack_image(xcmimage[t], xb, yb, zb);
    const double fx = f[3 * i], fy = f[3 * i + 1], fz = f[3 * i + 2];
    const double dx = (x[3 * i] + xb * rc.prd[0]) - F.xcm[0];
    const double dy = (x[3 * i + 1] + yb * rc.prd[1]) - F.xcm[1];
    const double dz = (x[3 * i + 2] + zb * rc.prd[2]) - F.xcm[2];
    s[0] += fx; s[1] += fy; s[2] += fz;
    s[3] += dy * fz - dz * fy;
    s[4] += dz * fx - dx * fz;
    s[5] += dx * fy - dy * fx;
  }
  if (LANES > 1) {
#pragma unroll
    for (int off = LANES / 2; off > 0; off >>= 1)
      for (int k = 0; k < 6; k++) s[k] += __shfl_xor_sync(0xffffffffu, s[k], off);
    if (lane != 0) return;
  }
  if (STAGE == 1) {
    for (int k = 0; k < 6; k++) sum6[(size_t)6 * b + k] = s[k];
    return;
  }
  BodyDyn D = dyn[b];
  for (int k = 0; k < 3; k++) { D.fcm[k] = s[k]; D.torque[k] = s[3 + k]; }
  if (MODE == MODE_SETUP) {
    angmom_to_omega(D.angmom, F.ex, F.ey, F.ez, D.inertia, F.omega);
    double mbody[3];
    tmatvec(F.ex, F.ey, F.ez, D.angmom, mbody);
    quatvec(D.quat, mbody, D.conjqm);
    for (int k = 0; k < 4; k++) D.conjqm[k] *= 2.0;
  } else {
    double scale_t = 1.0, scale_r = 1.0;
    if (rc.tstat) {
      scale_t = exp(-1.0 * rc.dtq * chain->eta_dot_t[0]);
      scale_r = exp(-1.0 * rc.dtq * chain->eta_dot_r[0]);
    }
    const double dtfm = rc.dtf / D.masstotal, dtf2 = rc.dtf * 2.0;
    for (int k = 0; k < 3; k++) {
      if (rc.tstat) F.vcm[k] *= scale_t;
      F.vcm[k] += dtfm * D.fcm[k];
    }
    double tbody[3], fquat[4], mbody[3];
    tmatvec(F.ex, F.ey, F.ez, D.torque, tbody);
    quatvec(D.quat, tbody, fquat);
    for (int k = 0; k < 4; k++) D.conjqm[k] = (rc.tstat ? scale_r * D.conjqm[k] : D.conjqm[k]) + dtf2 * fquat[k];
    invquatvec(D.quat, D.conjqm, mbody);
    matvec3(F.ex, F.ey, F.ez, mbody, D.angmom);
    for (int k = 0; k < 3; k++) D.angmom[k] *= 0.5;
    angmom_to_omega(D.angmom, F.ex, F.ey, F.ez, D.inertia, F.omega);
  }
  if (MODE == MODE_SETUP && rc.tstat)  // akin_t / akin_r of FixRigidNH::setup (fix_rigid_nh.cpp:339-344)
    akin[b] = make_double2(D.masstotal * (F.vcm[0] * F.vcm[0] + F.vcm[1] * F.vcm[1] + F.vcm[2] * F.vcm[2]),
                           D.angmom[0] * F.omega[0] + D.angmom[1] * F.omega[1] + D.angmom[2] * F.omega[2]);
  frame[b] = F;
  dyn[b] = D;
}

// FixRigidNH::initial_integrate, the body loop (fix_rigid_nh.cpp:470-545)
__global__ void __launch_bounds__(128) k_rigid_initial(int nbody, RigidConst rc, const Chain *__restrict__ chain,
                                                      BodyFrame *__restrict__ frame, BodyDyn *__restrict__ dyn,
                                                      double2 *__restrict__ akin)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbody) return;
  BodyFrame F = frame[b];
  BodyDyn D = dyn[b];
  double scale_t = 1.0, scale_r = 1.0;
  if (rc.tstat) {
    scale_t = exp(-rc.dtq * chain->eta_dot_t[0]);
    scale_r = exp(-rc.dtq * chain->eta_dot_r[0]);
  }
  const double dtfm = rc.dtf / D.masstotal, dtf2 = rc.dtf * 2.0;
  double akt = 0.0;
  for (int k = 0; k < 3; k++) {
    F.vcm[k] += dtfm * D.fcm[k];
    if (rc.tstat) F.vcm[k] *= scale_t;
  }
  if (rc.tstat) akt = D.masstotal * (F.vcm[0] * F.vcm[0] + F.vcm[1] * F.vcm[1] + F.vcm[2] * F.vcm[2]);
  for (int k = 0; k < 3; k++) F.xcm[k] += rc.dtv * F.vcm[k];
  double tbody[3], fquat[4], mbody[3];
  tmatvec(F.ex, F.ey, F.ez, D.torque, tbody);
  quatvec(D.quat, tbody, fquat);
  for (int k = 0; k < 4; k++) {
    D.conjqm[k] += dtf2 * fquat[k];
    if (rc.tstat) D.conjqm[k] *= scale_r;
  }
  no_squish_rotate(3, D.conjqm, D.quat, D.inertia, rc.dtq);
  no_squish_rotate(2, D.conjqm, D.quat, D.inertia, rc.dtq);
  no_squish_rotate(1, D.conjqm, D.quat, D.inertia, rc.dtv);
  no_squish_rotate(2, D.conjqm, D.quat, D.inertia, rc.dtq);
  no_squish_rotate(3, D.conjqm, D.quat, D.inertia, rc.dtq);
  q_to_exyz(D.quat, F.ex, F.ey, F.ez);
  invquatvec(D.quat, D.conjqm, mbody);
  matvec3(F.ex, F.ey, F.ez, mbody, D.angmom);
  for (int k = 0; k < 3; k++) D.angmom[k] *= 0.5;
  angmom_to_omega(D.angmom, F.ex, F.ey, F.ez, D.inertia, F.omega);
  if (rc.tstat) akin[b] = make_double2(akt, D.angmom[0] * F.omega[0] + D.angmom[1] * F.omega[1] + D.angmom[2] * F.omega[2]);
  frame[b] = F;
  dyn[b] = D;
}

// One block.  Sums akin[] in a fixed order; then thread 0 runs
//   SETUP = 1: thermostat masses and chain forces of FixRigidNH::setup (fix_rigid_nh.cpp:371-388)
//   SETUP = 0: nhc_temp_integrate (fix_rigid_nh.cpp:794-885)
template <int SETUP>
__global__ void __launch_bounds__(1024) k_rigid_nhc(int nbody, RigidConst rc, double t_target, const double2 *__restrict__ akin,
                                                   Chain *__restrict__ chain)
{
  __shared__ double sh[2][1024];
  double a = 0.0, r = 0.0;
  for (int b = threadIdx.x; b < nbody; b += blockDim.x) {
    a += akin[b].x;
    r += akin[b].y;
  }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = r;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + off];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  Chain &c = *chain;
  const int nc = rc.t_chain;
  c.akin_t = sh[0][0];
  c.akin_r = sh[1][0];
  c.t_target = t_target;
  const double kt = rc.boltz * t_target;
  const double t_mass = kt / (rc.t_freq * rc.t_freq);
  c.q_t[0] = rc.nf_t * t_mass;
  c.q_r[0] = rc.nf_r * t_mass;
  for (int i = 1; i < nc; i++) c.q_t[i] = c.q_r[i] = t_mass;
  if (SETUP) {
    for (int i = 1; i < nc; i++) {
      c.f_eta_t[i] = (c.q_t[i - 1] * c.eta_dot_t[i - 1] * c.eta_dot_t[i - 1] - kt) / c.q_t[i];
      c.f_eta_r[i] = (c.q_r[i - 1] * c.eta_dot_r[i - 1] * c.eta_dot_r[i - 1] - kt) / c.q_r[i];
    }
    return;
  }
  c.f_eta_t[0] = (c.akin_t * rc.mvv2e - rc.nf_t * kt) / c.q_t[0];
  c.f_eta_r[0] = (c.akin_r * rc.mvv2e - rc.nf_r * kt) / c.q_r[0];
  for (int it = 0; it < rc.t_iter; it++)
    for (int j = 0; j < rc.t_order; j++) {
      const double w1 = rc.wdti1[j], w2 = rc.wdti2[j], w4 = rc.wdti4[j];
      for (int which = 0; which < 2; which++) {
        double *ed = which ? c.eta_dot_r : c.eta_dot_t, *fe = which ? c.f_eta_r : c.f_eta_t;
        ed[nc - 1] += w2 * fe[nc - 1];
        for (int k = 1; k < nc; k++) {
          const double tmp = w4 * ed[nc - k], ms = maclaurin_series(tmp), s = exp(-1.0 * tmp);
          ed[nc - k - 1] = ed[nc - k - 1] * (s * s) + w2 * fe[nc - k - 1] * s * ms;
        }
      }
      for (int k = 0; k < nc; k++) {
        c.eta_t[k] += w1 * c.eta_dot_t[k];
        c.eta_r[k] += w1 * c.eta_dot_r[k];
      }
      for (int which = 0; which < 2; which++) {
        double *ed = which ? c.eta_dot_r : c.eta_dot_t, *fe = which ? c.f_eta_r : c.f_eta_t;
        const double *q = which ? c.q_r : c.q_t;
        for (int k = 1; k < nc; k++) fe[k] = (q[k - 1] * ed[k - 1] * ed[k - 1] - kt) / q[k];
        for (int k = 0; k < nc - 1; k++) {
          const double tmp = w4 * ed[k + 1], ms = maclaurin_series(tmp), s = exp(-1.0 * tmp);
          ed[k] = ed[k] * (s * s) + w2 * fe[k] * s * ms;
          fe[k + 1] = (q[k] * ed[k] * ed[k] - kt) / q[k + 1];
        }
        ed[nc - 1] += w2 * fe[nc - 1];
      }
    }
}

struct AtomRec {
  double d[3], mass;  // displace (body frame), mass
};

// FixRigid::set_xv (XV = 1, fix_rigid.cpp:1289-1392) / set_v (XV = 0, :1465-1560) for point particles; VIR = 1 adds the
// constraint-force virial of this half step as per-block partial sums (vpart[6][nblocks]).
template <int XV, int VIR>
__global__ void __launch_bounds__(256) k_rigid_atoms(int n, RigidConst rc, const int *__restrict__ tag, const int *__restrict__ abody,
                                                    const AtomRec *__restrict__ arec, const int *__restrict__ xcmimage,
                                                    const BodyFrame *__restrict__ frame, double *__restrict__ x,
                                                    double *__restrict__ v, const double *__restrict__ f,
                                                    double *__restrict__ vpart)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double vr[6] = {0, 0, 0, 0, 0, 0};
  if (i < n) {
    const int t = tag[i] - 1, b = abody[t];
    if (b >= 0) {
      int xb, yb, zb;
      unpack_image(xcmimage[t], xb, yb, zb);
      const AtomRec a = arec[t];
      const BodyFrame *F = frame + b;
      double ex[3], ey[3], ez[3], om[3], vc[3], xc[3];
      for (int k = 0; k < 3; k++) { ex[k] = F->ex[k]; ey[k] = F->ey[k]; ez[k] = F->ez[k]; om[k] = F->omega[k]; vc[k] = F->vcm[k]; xc[k] = F->xcm[k]; }
      const double xo[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
      const double x0[3] = {xo[0] + xb * rc.prd[0], xo[1] + yb * rc.prd[1], xo[2] + zb * rc.prd[2]};
      const double v0[3] = {v[3 * i], v[3 * i + 1], v[3 * i + 2]};
      double xr[3], vn[3];
      matvec3(ex, ey, ez, a.d, xr);
      vn[0] = om[1] * xr[2] - om[2] * xr[1] + vc[0];
      vn[1] = om[2] * xr[0] - om[0] * xr[2] + vc[1];
      vn[2] = om[0] * xr[1] - om[1] * xr[0] + vc[2];
      v[3 * i] = vn[0]; v[3 * i + 1] = vn[1]; v[3 * i + 2] = vn[2];
      if (XV) {
        x[3 * i] = xr[0] + (xc[0] - xb * rc.prd[0]);
        x[3 * i + 1] = xr[1] + (xc[1] - yb * rc.prd[1]);
        x[3 * i + 2] = xr[2] + (xc[2] - zb * rc.prd[2]);
      }
      if (VIR) {
        const double fc0 = a.mass * (vn[0] - v0[0]) / rc.dtf - f[3 * i];
        const double fc1 = a.mass * (vn[1] - v0[1]) / rc.dtf - f[3 * i + 1];
        const double fc2 = a.mass * (vn[2] - v0[2]) / rc.dtf - f[3 * i + 2];
        vr[0] = 0.5 * x0[0] * fc0; vr[1] = 0.5 * x0[1] * fc1; vr[2] = 0.5 * x0[2] * fc2;
        vr[3] = 0.5 * x0[0] * fc1; vr[4] = 0.5 * x0[0] * fc2; vr[5] = 0.5 * x0[1] * fc2;
      }
    }
  }
  if (VIR) {
    __shared__ double sh[6][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      double s = vr[k];
      for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
      if (lane == 0) sh[k][w] = s;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      double s = 0.0;
      for (int j = 0; j < (int)(blockDim.x >> 5); j++) s += sh[threadIdx.x][j];
      vpart[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
    }
  }
}

// virial[k] = scale * (virial[k] * keep + sum of the block partials), fixed order (one block, 6 warps)
__global__ void k_rigid_virial_sum(int nblocks, const double *__restrict__ vpart, double keep, double scale, double *__restrict__ virial)
{
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0.0;
  for (int j = lane; j < nblocks; j += 32) s += vpart[(size_t)k * nblocks + j];
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if (lane == 0) virial[k] = scale * (virial[k] * keep + s);
}

// Domain::remap on every body's centre of mass (domain.cpp:1329-1410), body image flags alongside
__global__ void k_rigid_remap(int nbody, RigidConst rc, BodyFrame *__restrict__ frame, int *__restrict__ imagebody)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbody) return;
  for (int d = 0; d < 3; d++) {
    double c = frame[b].xcm[d];
    int im = imagebody[3 * b + d];
    if (!isfinite(c)) continue;  // a blown-up trajectory must not hang the device; the caller sees it in x
    // the reference loops `while (c < lo) c += prd` / `while (c >= hi) c -= prd`; one period (the only case a sane
    // trajectory produces between two rebuilds) is done the same way, bit for bit; more periods in one multiply
    const double nper = floor((c - rc.lo[d]) / rc.prd[d]);
    if (nper == -1.0) { c += rc.prd[d]; im -= 1; }
    else if (nper == 1.0) { c -= rc.prd[d]; im += 1; }
    else if (nper != 0.0 && fabs(nper) < 1.0e6) { c -= nper * rc.prd[d]; im += (int)nper; }
    if (c < rc.lo[d]) { c += rc.prd[d]; im -= 1; }
    if (c >= rc.hi[d]) { c -= rc.prd[d]; im += 1; }
    c = fmax(c, rc.lo[d]);
    frame[b].xcm[d] = c;
    imagebody[3 * b + d] = im & IMGMASK;
  }
}

// FixRigid::image_shift (fix_rigid.cpp:1150-1175)
__global__ void k_rigid_image_shift(int n, const int *__restrict__ tag, const int *__restrict__ image, const int *__restrict__ abody,
                                    const int *__restrict__ imagebody, int *__restrict__ xcmimage)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = tag[i] - 1, b = abody[t];
  if (b < 0) return;
  const int im = image[i];
  const int xd = IMGMAX + (im & IMGMASK) - imagebody[3 * b];
  const int yd = IMGMAX + ((im >> IMGBITS) & IMGMASK) - imagebody[3 * b + 1];
  const int zd = IMGMAX + (im >> IMG2BITS) - imagebody[3 * b + 2];
  xcmimage[t] = (zd << IMG2BITS) | (yd << IMGBITS) | xd;
}

// kinetic terms of FixRigid::compute_scalar / extract_ke / extract_erotational per body: {M vcm^2, sum_k I_k wbody_k^2}
__global__ void k_rigid_kinetic(int nbody, const BodyFrame *__restrict__ frame, const BodyDyn *__restrict__ dyn, double2 *__restrict__ out)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbody) return;
  const BodyFrame &F = frame[b];
  const BodyDyn &D = dyn[b];
  double ex[3], ey[3], ez[3], wb[3];
  q_to_exyz(D.quat, ex, ey, ez);  // = the columns of quat_to_mat (math_extra.cpp:422-446)
  tmatvec(ex, ey, ez, D.angmom, wb);
  double rot = 0.0;
  for (int k = 0; k < 3; k++) {
    const double w = (D.inertia[k] == 0.0) ? 0.0 : wb[k] / D.inertia[k];
    rot += D.inertia[k] * w * w;
  }
  out[b] = make_double2(D.masstotal * (F.vcm[0] * F.vcm[0] + F.vcm[1] * F.vcm[1] + F.vcm[2] * F.vcm[2]), rot);
}

__global__ void __launch_bounds__(1024) k_rigid_sum2(int n, const double2 *__restrict__ in, double *__restrict__ out)
{
  __shared__ double sh[2][1024];
  double a = 0.0, r = 0.0;
  for (int b = threadIdx.x; b < n; b += blockDim.x) { a += in[b].x; r += in[b].y; }
  sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = r;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) { sh[0][threadIdx.x] += sh[0][threadIdx.x + off]; sh[1][threadIdx.x] += sh[1][threadIdx.x + off]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = sh[0][0]; out[1] = sh[1][0]; }
}

// ---- host: one-time body setup ----------------------------------------------------------------------------------

// MathExtra::jacobi (math_extra.cpp:101-175) for one symmetric 3x3 matrix; evec columns = eigenvectors
static bool jacobi3(double m[3][3], double eval[3], double evec[3][3])
{
  auto rot = [](double a[3][3], int i, int j, int k, int l, double s, double tau) {
    const double g = a[i][j], h = a[k][l];
    a[i][j] = g - s * (h + g * tau);
    a[k][l] = h + s * (g - h * tau);
  };
  double b[3], z[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) evec[i][j] = (i == j) ? 1.0 : 0.0;
    b[i] = eval[i] = m[i][i];
  }
  for (int iter = 1; iter <= 50; iter++) {
    const double sm = fabs(m[0][1]) + fabs(m[0][2]) + fabs(m[1][2]);
    if (sm == 0.0) return true;
    const double tresh = (iter < 4) ? 0.2 * sm / 9 : 0.0;
    for (int i = 0; i < 2; i++)
      for (int j = i + 1; j < 3; j++) {
        const double g = 100.0 * fabs(m[i][j]);
        if (iter > 4 && fabs(eval[i]) + g == fabs(eval[i]) && fabs(eval[j]) + g == fabs(eval[j])) m[i][j] = 0.0;
        else if (fabs(m[i][j]) > tresh) {
          double h = eval[j] - eval[i], t;
          if (fabs(h) + g == fabs(h)) t = m[i][j] / h;
          else {
            const double theta = 0.5 * h / m[i][j];
            t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
            if (theta < 0.0) t = -t;
          }
          const double c = 1.0 / sqrt(1.0 + t * t), s = t * c, tau = s / (1.0 + c);
          h = t * m[i][j];
          z[i] -= h; z[j] += h; eval[i] -= h; eval[j] += h;
          m[i][j] = 0.0;
          for (int k = 0; k < i; k++) rot(m, k, i, k, j, s, tau);
          for (int k = i + 1; k < j; k++) rot(m, i, k, k, j, s, tau);
          for (int k = j + 1; k < 3; k++) rot(m, i, k, j, k, s, tau);
          for (int k = 0; k < 3; k++) rot(evec, k, i, k, j, s, tau);
        }
      }
    for (int i = 0; i < 3; i++) {
      eval[i] = b[i] += z[i];
      z[i] = 0.0;
    }
  }
  return false;
}

static void exyz_to_q(const double *ex, const double *ey, const double *ez, double *q)  // math_extra.cpp:359-394
{
  const double q0sq = 0.25 * (ex[0] + ey[1] + ez[2] + 1.0), q1sq = q0sq - 0.5 * (ey[1] + ez[2]);
  const double q2sq = q0sq - 0.5 * (ex[0] + ez[2]), q3sq = q0sq - 0.5 * (ex[0] + ey[1]);
  q[0] = q[1] = q[2] = q[3] = 0.0;
  if (q0sq >= 0.25) {
    q[0] = sqrt(q0sq); q[1] = (ey[2] - ez[1]) / (4.0 * q[0]); q[2] = (ez[0] - ex[2]) / (4.0 * q[0]); q[3] = (ex[1] - ey[0]) / (4.0 * q[0]);
  } else if (q1sq >= 0.25) {
    q[1] = sqrt(q1sq); q[0] = (ey[2] - ez[1]) / (4.0 * q[1]); q[2] = (ey[0] + ex[1]) / (4.0 * q[1]); q[3] = (ex[2] + ez[0]) / (4.0 * q[1]);
  } else if (q2sq >= 0.25) {
    q[2] = sqrt(q2sq); q[0] = (ez[0] - ex[2]) / (4.0 * q[2]); q[1] = (ey[0] + ex[1]) / (4.0 * q[2]); q[3] = (ez[1] + ey[2]) / (4.0 * q[2]);
  } else if (q3sq >= 0.25) {
    q[3] = sqrt(q3sq); q[0] = (ex[1] - ey[0]) / (4.0 * q[3]); q[1] = (ez[0] + ex[2]) / (4.0 * q[3]); q[2] = (ez[1] + ey[2]) / (4.0 * q[3]);
  }
  const double norm = 1.0 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; k++) q[k] *= norm;
}

}  // namespace polb200

using namespace polb200;

struct polb200_rigid {
  std::string err;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {};
  long launches = 0;
  float ms_last = 0.f;
  bool ready = false, setup_done = false;
  int evflag = 0;
  RigidConst rc{};
  polb200_rigid_params par{};
  int nbody = 0, maxtag = 0, nlinear = 0, maxmembers = 0, natoms_body = 0;
  double tfactor = 0.0;
  // host mirrors needed by dof()
  std::vector<int> h_abody, h_nrigid;
  std::vector<char> h_linear;
  DBuf<BodyFrame> frame;
  DBuf<BodyDyn> dyn;
  DBuf<Chain> chain;
  DBuf<int> abody, xcmimage, member_first, member_tag, idx_of_tag, imagebody, c_tag, c_image;
  DBuf<AtomRec> arec;
  DBuf<double2> akin;
  DBuf<double> vpart, virial, c_x, c_v, c_f, sums, sum6;
  // more than one process (polb200_rigid_comm_init): bodies replicated, atoms owned by exactly one process
  ncclComm_t nccl = nullptr;
  int rank = 0, nranks = 1;
};

static NcclApi g_rigid_nccl;

namespace polb200 {

template <class F>
static int rigid_guarded(polb200_rigid *r, F &&fn)
{
  try {
    fn();
    return POLB200_OK;
  } catch (const StyleError &x) {
    r->err = x.msg;
    return x.code;
  } catch (const CudaError &x) {
    r->err = x.msg;
    return POLB200_ERR_CUDA;
  } catch (const std::exception &x) {
    r->err = x.what();
    return POLB200_ERR_ARG;
  }
}

#define RIGID_LAUNCHED(r)          \
  do {                             \
    CUDA_CHECK(cudaGetLastError()); \
    (r)->launches++;               \
  } while (0)

static void set_timestep(polb200_rigid *r, double dt)
{
  r->rc.dtv = dt;
  r->rc.dtf = 0.5 * dt * r->par.ftm2v;
  r->rc.dtq = 0.5 * dt;
  if (r->rc.tstat) {
    double w[5] = {0, 0, 0, 0, 0};
    if (r->rc.t_order == 3) {  // Table 1 of Kamberaj et al., fix_rigid_nh.cpp:248-262
      w[0] = 1.0 / (2.0 - pow(2.0, 1.0 / 3.0));
      w[1] = 1.0 - 2.0 * w[0];
      w[2] = w[0];
    } else {
      w[0] = 1.0 / (4.0 - pow(4.0, 1.0 / 3.0));
      w[1] = w[0];
      w[2] = 1.0 - 4.0 * w[0];
      w[3] = w[0];
      w[4] = w[0];
    }
    for (int i = 0; i < r->rc.t_order; i++) {  // fix_rigid_nh.cpp:405-411
      r->rc.wdti1[i] = w[i] * r->rc.dtv / r->rc.t_iter;
      r->rc.wdti2[i] = r->rc.wdti1[i] / 2.0;
      r->rc.wdti4[i] = r->rc.wdti1[i] / 4.0;
    }
  }
}

// the caller's per-step arrays on the device (copied in when they are host memory)
struct StepArrays {
  const int *tag;
  double *x, *v;
  const double *f;
};

static StepArrays stage_in(polb200_rigid *r, const polb200_rigid_atoms *a, bool need_x, bool need_f)
{
  const int n = a->nlocal;
  if (n < 0 || (n > 0 && (!a->tag || !a->v || (need_x && !a->x) || (need_f && !a->f))))
    throw StyleError{POLB200_ERR_ARG, "polb200_rigid: missing per-atom array"};
  StepArrays s{a->tag, a->x, a->v, a->f};
  if (!a->on_device) {
    r->c_tag.ensure(n); r->c_x.ensure((size_t)3 * n); r->c_v.ensure((size_t)3 * n); r->c_f.ensure((size_t)3 * n);
    CUDA_CHECK(cudaMemcpyAsync(r->c_tag.p, a->tag, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, r->stream));
    if (a->x) CUDA_CHECK(cudaMemcpyAsync(r->c_x.p, a->x, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice, r->stream));
    CUDA_CHECK(cudaMemcpyAsync(r->c_v.p, a->v, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice, r->stream));
    if (a->f) CUDA_CHECK(cudaMemcpyAsync(r->c_f.p, a->f, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice, r->stream));
    s = StepArrays{r->c_tag.p, r->c_x.p, r->c_v.p, r->c_f.p};
  }
  return s;
}

static void stage_out(polb200_rigid *r, const polb200_rigid_atoms *a, bool x_written)
{
  const int n = a->nlocal;
  if (!a->on_device && n > 0) {
    if (x_written) CUDA_CHECK(cudaMemcpyAsync(a->x, r->c_x.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    CUDA_CHECK(cudaMemcpyAsync(a->v, r->c_v.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
  }
  CUDA_CHECK(cudaEventRecord(r->ev[1], r->stream));
  CUDA_CHECK(cudaStreamSynchronize(r->stream));
  cudaEventElapsedTime(&r->ms_last, r->ev[0], r->ev[1]);
}

static void check_atom_count(polb200_rigid *r, int n)
{
  if (!r->nccl && n < r->natoms_body)
    throw StyleError{POLB200_ERR_ARG, "polb200_rigid: fewer atoms than the rigid bodies hold (atoms of a body must stay on this process)"};
}

template <int MODE>
static void launch_bodies(polb200_rigid *r, const StepArrays &s, int n)
{
  if (r->nccl) {
    // partial sums over the members this process owns -> all-reduce -> every process updates every body identically
    CUDA_CHECK(cudaMemsetAsync(r->idx_of_tag.p, 0xff, (size_t)r->maxtag * sizeof(int), r->stream));
    if (n > 0) {
      k_rigid_index<<<cdiv(n, 256), 256, 0, r->stream>>>(n, s.tag, r->idx_of_tag.p);
      RIGID_LAUNCHED(r);
    }
    r->sum6.ensure((size_t)6 * r->nbody);
    if (r->maxmembers <= 16)
      k_rigid_bodies<1, MODE, 1><<<cdiv(r->nbody, 256), 256, 0, r->stream>>>(r->nbody, r->rc, r->chain.p, r->member_first.p, r->member_tag.p,
                                                                            r->idx_of_tag.p, r->xcmimage.p, s.x, s.f, r->frame.p, r->dyn.p,
                                                                            r->akin.p, r->sum6.p);
    else
      k_rigid_bodies<32, MODE, 1><<<cdiv((long)r->nbody * 32, 256), 256, 0, r->stream>>>(r->nbody, r->rc, r->chain.p, r->member_first.p,
                                                                                       r->member_tag.p, r->idx_of_tag.p, r->xcmimage.p, s.x,
                                                                                       s.f, r->frame.p, r->dyn.p, r->akin.p, r->sum6.p);
    RIGID_LAUNCHED(r);
    const ncclResult_t rc = g_rigid_nccl.AllReduce(r->sum6.p, r->sum6.p, (size_t)6 * r->nbody, ncclDouble, ncclSum, r->nccl, r->stream);
    if (rc != ncclSuccess) throw CudaError{std::string("ncclAllReduce of the body forces and torques: ") + g_rigid_nccl.GetErrorString(rc)};
    k_rigid_bodies<1, MODE, 2><<<cdiv(r->nbody, 256), 256, 0, r->stream>>>(r->nbody, r->rc, r->chain.p, r->member_first.p, r->member_tag.p,
                                                                          r->idx_of_tag.p, r->xcmimage.p, s.x, s.f, r->frame.p, r->dyn.p,
                                                                          r->akin.p, r->sum6.p);
    RIGID_LAUNCHED(r);
    return;
  }
  k_rigid_index<<<cdiv(n, 256), 256, 0, r->stream>>>(n, s.tag, r->idx_of_tag.p);
  RIGID_LAUNCHED(r);
  if (r->maxmembers <= 16)
    k_rigid_bodies<1, MODE><<<cdiv(r->nbody, 256), 256, 0, r->stream>>>(r->nbody, r->rc, r->chain.p, r->member_first.p, r->member_tag.p,
                                                                       r->idx_of_tag.p, r->xcmimage.p, s.x, s.f, r->frame.p, r->dyn.p, r->akin.p, nullptr);
  else
    k_rigid_bodies<32, MODE><<<cdiv((long)r->nbody * 32, 256), 256, 0, r->stream>>>(r->nbody, r->rc, r->chain.p, r->member_first.p,
                                                                                  r->member_tag.p, r->idx_of_tag.p, r->xcmimage.p, s.x, s.f,
                                                                                  r->frame.p, r->dyn.p, r->akin.p, nullptr);
  RIGID_LAUNCHED(r);
}

template <int XV>
static void launch_atoms(polb200_rigid *r, const StepArrays &s, int n, double keep, double scale)
{
  const int nb = cdiv(n, 256);
  if (n == 0) {  // a process that owns no atom (more than one process): only the virial bookkeeping
    if (r->evflag) {
      r->vpart.ensure(6);
      k_rigid_virial_sum<<<1, 192, 0, r->stream>>>(0, r->vpart.p, keep, scale, r->virial.p);
      RIGID_LAUNCHED(r);
    }
    return;
  }
  if (r->evflag) {
    r->vpart.ensure((size_t)6 * nb);
    k_rigid_atoms<XV, 1><<<nb, 256, 0, r->stream>>>(n, r->rc, s.tag, r->abody.p, r->arec.p, r->xcmimage.p, r->frame.p, s.x, s.v, s.f, r->vpart.p);
    RIGID_LAUNCHED(r);
    k_rigid_virial_sum<<<1, 192, 0, r->stream>>>(nb, r->vpart.p, keep, scale, r->virial.p);
    RIGID_LAUNCHED(r);
  } else {
    k_rigid_atoms<XV, 0><<<nb, 256, 0, r->stream>>>(n, r->rc, s.tag, r->abody.p, r->arec.p, r->xcmimage.p, r->frame.p, s.x, s.v, s.f, nullptr);
    RIGID_LAUNCHED(r);
  }
}

}  // namespace polb200

extern "C" {

int polb200_rigid_create(polb200_rigid_t **out, int device)
{
  if (!out) return POLB200_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    fprintf(stderr, "polb200_rigid_create: no usable CUDA device %d (found %d); there is no CPU fallback\n", device, count);
    return POLB200_ERR_CUDA;
  }
  polb200_rigid *r = new polb200_rigid();
  r->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&r->ev[0]) != cudaSuccess || cudaEventCreate(&r->ev[1]) != cudaSuccess) {
    delete r;
    return POLB200_ERR_CUDA;
  }
  *out = r;
  return POLB200_OK;
}

void polb200_rigid_destroy(polb200_rigid_t *r)
{
  if (!r) return;
  cudaSetDevice(r->device);
  cudaStreamSynchronize(r->stream);
  r->frame.release(); r->dyn.release(); r->chain.release(); r->abody.release(); r->xcmimage.release();
  r->member_first.release(); r->member_tag.release(); r->idx_of_tag.release(); r->imagebody.release();
  r->c_tag.release(); r->c_image.release(); r->arec.release(); r->akin.release(); r->vpart.release();
  r->virial.release(); r->c_x.release(); r->c_v.release(); r->c_f.release(); r->sums.release(); r->sum6.release();
  if (r->nccl) g_rigid_nccl.CommDestroy(r->nccl);
  cudaEventDestroy(r->ev[0]); cudaEventDestroy(r->ev[1]);
  cudaStreamDestroy(r->stream);
  delete r;
}

const char *polb200_rigid_last_error(const polb200_rigid_t *r) { return r ? r->err.c_str() : "null handle"; }

// one atom as it travels at init when the bodies are built on more than one process
struct InitRec {
  double mass, x[3], v[3];
  int tag, molecule, ingroup, image;
};

static int rigid_init_impl(polb200_rigid_t *r, const polb200_rigid_params *p, int nlocal, const int *tag, const int *molecule,
                           const int *ingroup, const double *mass, const int *image, const double *x, const double *v,
                           polb200_rigid_info *info);

int polb200_rigid_init(polb200_rigid_t *r, const polb200_rigid_params *p, int nlocal, const int *tag, const int *molecule,
                       const int *ingroup, const double *mass, const int *image, const double *x, const double *v,
                       polb200_rigid_info *info)
{
  if (!r || !p) return POLB200_ERR_ARG;
  if (!r->nccl) return rigid_init_impl(r, p, nlocal, tag, molecule, ingroup, mass, image, x, v, info);
  // More than one process: the reference builds its bodies from per-process partial sums that it all-reduces
  // (fix_rigid.cpp:1605-2211, six MPI_Allreduce).  Here every process gathers the atom records of all processes once
  // and runs the single-process construction on them: its sums run in atom-id order, so every process ends up with
  // bit-identical bodies whatever the decomposition.
  std::vector<InitRec> all;
  int ntotal = 0;
  const int rc = rigid_guarded(r, [&] {
    const int n = nlocal;
    if (n < 0 || (n > 0 && (!tag || !molecule || !mass || !image || !x || !v))) throw StyleError{POLB200_ERR_ARG, "Illegal fix rigid command"};
    CUDA_CHECK(cudaSetDevice(r->device));
    DBuf<int> counts;
    counts.ensure((size_t)r->nranks + 1);
    CUDA_CHECK(cudaMemcpyAsync(counts.p + r->nranks, &n, sizeof(int), cudaMemcpyHostToDevice, r->stream));
    ncclResult_t nr = g_rigid_nccl.AllGather(counts.p + r->nranks, counts.p, 1, ncclInt, r->nccl, r->stream);
    if (nr != ncclSuccess) throw CudaError{std::string("ncclAllGather (rigid init): ") + g_rigid_nccl.GetErrorString(nr)};
    std::vector<int> hc(r->nranks);
    CUDA_CHECK(cudaMemcpyAsync(hc.data(), counts.p, (size_t)r->nranks * sizeof(int), cudaMemcpyDeviceToHost, r->stream));
    CUDA_CHECK(cudaStreamSynchronize(r->stream));
    counts.release();
    const size_t nmax = (size_t)std::max(1, *std::max_element(hc.begin(), hc.end()));
    std::vector<InitRec> mine(nmax);
    memset(mine.data(), 0, nmax * sizeof(InitRec));
    for (int i = 0; i < n; i++) {
      InitRec &a = mine[i];
      a.mass = mass[i];
      for (int k = 0; k < 3; k++) { a.x[k] = x[3 * i + k]; a.v[k] = v[3 * i + k]; }
      a.tag = tag[i]; a.molecule = molecule[i]; a.ingroup = ingroup ? ingroup[i] : 1; a.image = image[i];
    }
    DBuf<InitRec> send, recv;
    send.ensure(nmax);
    recv.ensure(nmax * r->nranks);
    CUDA_CHECK(cudaMemcpyAsync(send.p, mine.data(), nmax * sizeof(InitRec), cudaMemcpyHostToDevice, r->stream));
    nr = g_rigid_nccl.AllGather(send.p, recv.p, nmax * sizeof(InitRec), ncclChar, r->nccl, r->stream);
    if (nr != ncclSuccess) throw CudaError{std::string("ncclAllGather (rigid init): ") + g_rigid_nccl.GetErrorString(nr)};
    std::vector<InitRec> padded(nmax * r->nranks);
    CUDA_CHECK(cudaMemcpyAsync(padded.data(), recv.p, padded.size() * sizeof(InitRec), cudaMemcpyDeviceToHost, r->stream));
    CUDA_CHECK(cudaStreamSynchronize(r->stream));
    send.release();
    recv.release();
    for (int k = 0; k < r->nranks; k++) all.insert(all.end(), padded.begin() + k * nmax, padded.begin() + k * nmax + hc[k]);
    ntotal = (int)all.size();
  });
  if (rc != POLB200_OK) return rc;
  std::vector<int> g_tag(ntotal), g_mol(ntotal), g_in(ntotal), g_img(ntotal);
  std::vector<double> g_mass(ntotal), g_x((size_t)3 * ntotal), g_v((size_t)3 * ntotal);
  for (int i = 0; i < ntotal; i++) {
    const InitRec &a = all[i];
    g_tag[i] = a.tag; g_mol[i] = a.molecule; g_in[i] = a.ingroup; g_img[i] = a.image; g_mass[i] = a.mass;
    for (int k = 0; k < 3; k++) { g_x[3 * i + k] = a.x[k]; g_v[3 * i + k] = a.v[k]; }
  }
  return rigid_init_impl(r, p, ntotal, g_tag.data(), g_mol.data(), g_in.data(), g_mass.data(), g_img.data(), g_x.data(), g_v.data(), info);
}

static int rigid_init_impl(polb200_rigid_t *r, const polb200_rigid_params *p, int nlocal, const int *tag, const int *molecule,
                           const int *ingroup, const double *mass, const int *image, const double *x, const double *v,
                           polb200_rigid_info *info)
{
  if (!r || !p) return POLB200_ERR_ARG;
  return rigid_guarded(r, [&] {
    const int n = nlocal;
    if (n <= 0 || !tag || !molecule || !mass || !image || !x || !v) throw StyleError{POLB200_ERR_ARG, "Illegal fix rigid command"};
    for (int d = 0; d < 3; d++)
      if (!p->periodic[d]) throw StyleError{POLB200_ERR_UNSUPPORTED, "polb200_rigid: non-periodic dimensions are not supported"};
    if (p->thermostat) {  // fix_rigid_nvt.cpp:36-49
      if (p->t_start < 0.0 || p->t_stop <= 0.0) throw StyleError{POLB200_ERR_ARG, "Target temperature for fix rigid/nvt cannot be 0.0"};
      if (p->t_period <= 0.0) throw StyleError{POLB200_ERR_ARG, "Fix rigid/nvt period must be > 0.0"};
      if (p->t_chain < 1 || p->t_iter < 1) throw StyleError{POLB200_ERR_ARG, "Illegal fix rigid/nvt command"};
      if (p->t_order != 3 && p->t_order != 5) throw StyleError{POLB200_ERR_ARG, "Fix rigid/nvt temperature order must be 3 or 5"};
      if (p->t_chain > Chain::MAXCHAIN) throw StyleError{POLB200_ERR_UNSUPPORTED, "polb200_rigid: tparam Tchain > 64"};
    }
    CUDA_CHECK(cudaSetDevice(r->device));
    r->par = *p;
    RigidConst &rc = r->rc;
    rc = RigidConst{};
    rc.tstat = p->thermostat ? 1 : 0;
    rc.t_chain = p->thermostat ? p->t_chain : 1;
    rc.t_iter = p->t_iter;
    rc.t_order = p->t_order;
    rc.t_freq = p->thermostat ? 1.0 / p->t_period : 0.0;
    rc.mvv2e = p->mvv2e;
    rc.boltz = p->boltz;
    for (int d = 0; d < 3; d++) {
      rc.lo[d] = p->boxlo[d];
      rc.hi[d] = p->boxhi[d];
      rc.prd[d] = p->boxhi[d] - p->boxlo[d];
    }
    set_timestep(r, p->dt);

    // ---- body numbering: molecules of the group in ascending id (fix_rigid.cpp:184-220)
    int maxmol = -1, maxtag = 0;
    for (int i = 0; i < n; i++) {
      for (int k = 0; k < 3; k++)
        if (!std::isfinite(x[3 * i + k]) || !std::isfinite(v[3 * i + k]))
          throw StyleError{POLB200_ERR_NAN, "Non-numeric positions - simulation unstable"};
      if (tag[i] <= 0) throw StyleError{POLB200_ERR_ARG, "polb200_rigid: atom ids must be positive"};
      maxtag = std::max(maxtag, tag[i]);
      if (!ingroup || ingroup[i]) {
        if (molecule[i] < 0) throw StyleError{POLB200_ERR_ARG, "polb200_rigid: negative molecule id"};
        maxmol = std::max(maxmol, molecule[i]);
      }
    }
    if (maxmol < 0) throw StyleError{POLB200_ERR_ARG, "polb200_rigid: the fix group holds no atoms"};
    std::vector<int> mol2body((size_t)maxmol + 1, 0);
    for (int i = 0; i < n; i++)
      if (!ingroup || ingroup[i]) mol2body[molecule[i]]++;
    int nbody = 0;
    for (int m = 0; m <= maxmol; m++) mol2body[m] = mol2body[m] ? nbody++ : -1;
    std::vector<int> body(n, -1);
    for (int i = 0; i < n; i++)
      if (!ingroup || ingroup[i]) body[i] = mol2body[molecule[i]];
    r->nbody = nbody;
    r->maxtag = maxtag;

    // ---- setup_bodies_static (fix_rigid.cpp:1605-2112), point particles.  The reference sums over atoms in local order;
    // here every sum runs in atom-id order, so the bodies do not depend on how the caller has its atoms sorted.
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int p, int q) { return tag[p] < tag[q]; });
    const double *prd = rc.prd;
    auto unwrap = [&](int i, int img, double *u) {
      u[0] = x[3 * i] + ((img & IMGMASK) - IMGMAX) * prd[0];
      u[1] = x[3 * i + 1] + (((img >> IMGBITS) & IMGMASK) - IMGMAX) * prd[1];
      u[2] = x[3 * i + 2] + ((img >> IMG2BITS) - IMGMAX) * prd[2];
    };
    std::vector<int> xcmimage(n, 0), imagebody((size_t)3 * nbody, IMGMAX), nrigid(nbody, 0);
    std::vector<BodyFrame> frame(nbody);
    std::vector<BodyDyn> dyn(nbody);
    memset(frame.data(), 0, sizeof(BodyFrame) * nbody);
    memset(dyn.data(), 0, sizeof(BodyDyn) * nbody);
    std::vector<double> sum((size_t)6 * nbody, 0.0);
    for (int ii = 0; ii < n; ii++) {
      const int i = order[ii];
      if (body[i] < 0) continue;
      xcmimage[i] = image[i];
      double u[3];
      unwrap(i, xcmimage[i], u);
      double *s = &sum[(size_t)6 * body[i]];
      s[0] += u[0] * mass[i]; s[1] += u[1] * mass[i]; s[2] += u[2] * mass[i]; s[3] += mass[i];
      nrigid[body[i]]++;
    }
    for (int b = 0; b < nbody; b++) {
      dyn[b].masstotal = sum[6 * b + 3];
      for (int k = 0; k < 3; k++) frame[b].xcm[k] = sum[6 * b + k] / dyn[b].masstotal;
    }
    // pre_neighbor(): remap the centres of mass into the box, body-relative image flags (fix_rigid.cpp:1137-1175)
    for (int b = 0; b < nbody; b++)
      for (int d = 0; d < 3; d++) {
        double c = frame[b].xcm[d];
        int im = imagebody[3 * b + d];
        while (c < rc.lo[d]) { c += prd[d]; im = (im - 1) & IMGMASK; }
        while (c >= rc.hi[d]) { c -= prd[d]; im = (im + 1) & IMGMASK; }
        frame[b].xcm[d] = std::max(c, rc.lo[d]);
        imagebody[3 * b + d] = im;
      }
    for (int ii = 0; ii < n; ii++) {
      const int i = order[ii];
      if (body[i] < 0) continue;
      const int b = body[i], im = image[i];
      const int xd = IMGMAX + (im & IMGMASK) - imagebody[3 * b], yd = IMGMAX + ((im >> IMGBITS) & IMGMASK) - imagebody[3 * b + 1];
      const int zd = IMGMAX + (im >> IMG2BITS) - imagebody[3 * b + 2];
      xcmimage[i] = (zd << IMG2BITS) | (yd << IMGBITS) | xd;
    }
    std::fill(sum.begin(), sum.end(), 0.0);
    for (int ii = 0; ii < n; ii++) {
      const int i = order[ii];
      if (body[i] < 0) continue;
      double u[3];
      unwrap(i, xcmimage[i], u);
      const double *c = frame[body[i]].xcm;
      const double dx = u[0] - c[0], dy = u[1] - c[1], dz = u[2] - c[2], m = mass[i];
      double *s = &sum[(size_t)6 * body[i]];
      s[0] += m * (dy * dy + dz * dz); s[1] += m * (dx * dx + dz * dz); s[2] += m * (dx * dx + dy * dy);
      s[3] -= m * dy * dz; s[4] -= m * dx * dz; s[5] -= m * dx * dy;
    }
    std::vector<char> linear(nbody, 0);
    for (int b = 0; b < nbody; b++) {
      const double *s = &sum[(size_t)6 * b];
      double tensor[3][3] = {{s[0], s[5], s[4]}, {s[5], s[1], s[3]}, {s[4], s[3], s[2]}}, evec[3][3];
      BodyDyn &D = dyn[b];
      BodyFrame &F = frame[b];
      if (!jacobi3(tensor, D.inertia, evec)) throw StyleError{POLB200_ERR_ARG, "Insufficient Jacobi rotations for rigid body"};
      for (int k = 0; k < 3; k++) { F.ex[k] = evec[k][0]; F.ey[k] = evec[k][1]; F.ez[k] = evec[k][2]; }
      const double mx = std::max(std::max(D.inertia[0], D.inertia[1]), D.inertia[2]);
      for (int k = 0; k < 3; k++)
        if (D.inertia[k] < 1.0e-7 * mx) D.inertia[k] = 0.0;  // EPSILON, fix_rigid.cpp:52,1912-1920
      const double cr[3] = {F.ex[1] * F.ey[2] - F.ex[2] * F.ey[1], F.ex[2] * F.ey[0] - F.ex[0] * F.ey[2], F.ex[0] * F.ey[1] - F.ex[1] * F.ey[0]};
      if (cr[0] * F.ez[0] + cr[1] * F.ez[1] + cr[2] * F.ez[2] < 0.0)
        for (int k = 0; k < 3; k++) F.ez[k] = -F.ez[k];
      exyz_to_q(F.ex, F.ey, F.ez, D.quat);
      if (D.inertia[0] == 0.0 || D.inertia[1] == 0.0 || D.inertia[2] == 0.0) linear[b] = 1;
    }
    std::vector<AtomRec> arec_local(n);
    std::fill(sum.begin(), sum.end(), 0.0);
    for (int ii = 0; ii < n; ii++) {
      const int i = order[ii];
      AtomRec &a = arec_local[i];
      a.d[0] = a.d[1] = a.d[2] = 0.0;
      a.mass = mass[i];
      if (body[i] < 0) continue;
      double u[3], delta[3];
      unwrap(i, xcmimage[i], u);
      const BodyFrame &F = frame[body[i]];
      for (int k = 0; k < 3; k++) delta[k] = u[k] - F.xcm[k];
      tmatvec(F.ex, F.ey, F.ez, delta, a.d);
      double *s = &sum[(size_t)6 * body[i]];
      const double m = mass[i];
      s[0] += m * (a.d[1] * a.d[1] + a.d[2] * a.d[2]); s[1] += m * (a.d[0] * a.d[0] + a.d[2] * a.d[2]);
      s[2] += m * (a.d[0] * a.d[0] + a.d[1] * a.d[1]);
      s[3] -= m * a.d[1] * a.d[2]; s[4] -= m * a.d[0] * a.d[2]; s[5] -= m * a.d[0] * a.d[1];
    }
    for (int b = 0; b < nbody; b++) {  // fix_rigid.cpp:2078-2110
      const double *s = &sum[(size_t)6 * b], *I = dyn[b].inertia;
      const double TOL = 1.0e-6;
      bool bad = false;
      for (int k = 0; k < 3; k++) bad |= (I[k] == 0.0) ? (fabs(s[k]) > TOL) : (fabs((s[k] - I[k]) / I[k]) > TOL);
      const double norm = (I[0] + I[1] + I[2]) / 3.0;
      bad |= fabs(s[3] / norm) > TOL || fabs(s[4] / norm) > TOL || fabs(s[5] / norm) > TOL;
      if (bad) throw StyleError{POLB200_ERR_ARG, "Fix rigid: Bad principal moments"};
    }
    // ---- setup_bodies_dynamic (fix_rigid.cpp:2120-2211)
    std::fill(sum.begin(), sum.end(), 0.0);
    for (int ii = 0; ii < n; ii++) {
      const int i = order[ii];
      if (body[i] < 0) continue;
      double u[3];
      unwrap(i, xcmimage[i], u);
      const double *c = frame[body[i]].xcm;
      const double dx = u[0] - c[0], dy = u[1] - c[1], dz = u[2] - c[2], m = mass[i];
      const double *vi = v + 3 * i;
      double *s = &sum[(size_t)6 * body[i]];
      s[0] += vi[0] * m; s[1] += vi[1] * m; s[2] += vi[2] * m;
      s[3] += dy * m * vi[2] - dz * m * vi[1];
      s[4] += dz * m * vi[0] - dx * m * vi[2];
      s[5] += dx * m * vi[1] - dy * m * vi[0];
    }
    for (int b = 0; b < nbody; b++)
      for (int k = 0; k < 3; k++) {
        frame[b].vcm[k] = sum[6 * b + k] / dyn[b].masstotal;
        dyn[b].angmom[k] = sum[6 * b + 3 + k];
      }
    // tfactor (fix_rigid.cpp:757-764): nlinear is whatever the last dof() call left -- 0 on a first init, as in the reference
    {
      const double ndof = 6.0 * nbody - r->nlinear;
      r->tfactor = ndof > 0.0 ? p->mvv2e / (ndof * p->boltz) : 0.0;
    }
    // FixRigidNH::init (fix_rigid_nh.cpp:232-244)
    rc.nf_t = 3 * nbody;
    rc.nf_r = 3 * nbody;
    for (int b = 0; b < nbody; b++)
      for (int k = 0; k < 3; k++)
        if (fabs(dyn[b].inertia[k]) < 1.0e-7) rc.nf_r--;

    // ---- per-id tables and member lists (ids ascending inside a body)
    std::vector<int> abody_t(maxtag, -1), ximg_t(maxtag, 0);
    std::vector<AtomRec> arec_t(maxtag);
    memset(arec_t.data(), 0, sizeof(AtomRec) * maxtag);
    for (int i = 0; i < n; i++) {
      const int t = tag[i] - 1;
      abody_t[t] = body[i];
      ximg_t[t] = xcmimage[i];
      arec_t[t] = arec_local[i];
    }
    std::vector<int> first(nbody + 1, 0), members;
    for (int b = 0; b < nbody; b++) first[b + 1] = first[b] + nrigid[b];
    members.resize(first[nbody]);
    {
      std::vector<int> fill(first.begin(), first.end() - 1);
      for (int t = 0; t < maxtag; t++)
        if (abody_t[t] >= 0) members[fill[abody_t[t]]++] = t;
    }
    r->maxmembers = nbody ? *std::max_element(nrigid.begin(), nrigid.end()) : 0;
    r->natoms_body = first[nbody];
    r->h_abody = abody_t;
    r->h_nrigid = nrigid;
    r->h_linear = linear;

    auto up = [&](auto &dbuf, const auto &vec) {
      dbuf.ensure(vec.size());
      CUDA_CHECK(cudaMemcpyAsync(dbuf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice, r->stream));
    };
    up(r->frame, frame); up(r->dyn, dyn); up(r->abody, abody_t); up(r->xcmimage, ximg_t); up(r->arec, arec_t);
    up(r->member_first, first); up(r->member_tag, members); up(r->imagebody, imagebody);
    r->idx_of_tag.ensure(maxtag);
    r->akin.ensure(nbody);
    r->virial.ensure(8);
    r->sums.ensure(8);
    CUDA_CHECK(cudaMemsetAsync(r->virial.p, 0, 8 * sizeof(double), r->stream));
    CUDA_CHECK(cudaMemsetAsync(r->akin.p, 0, (size_t)nbody * sizeof(double2), r->stream));
    if (!r->ready) {  // thermostat chains survive a re-init (second `run`), like the fix's member arrays
      r->chain.ensure(1);
      CUDA_CHECK(cudaMemsetAsync(r->chain.p, 0, sizeof(Chain), r->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(r->stream));
    r->ready = true;
    r->setup_done = false;
    if (info) {
      info->nbody = nbody;
      int nl = 0;
      for (char c : linear) nl += c;
      info->nlinear = nl;
      info->nf_t = rc.nf_t;
      info->nf_r = rc.nf_r;
      info->maxmembers = r->maxmembers;
    }
  });
}

int polb200_rigid_dof(polb200_rigid_t *r, int nlocal, const int *tag, const int *tgroup, int *dof)
{
  if (!r || !dof || !tag) return POLB200_ERR_ARG;
  return rigid_guarded(r, [&] {
    if (!r->ready) throw StyleError{POLB200_ERR_STATE, "Cannot count rigid body degrees-of-freedom before bodies are initialized"};
    std::vector<int> nall(r->nbody, 0);
    for (int i = 0; i < nlocal; i++) {
      const int t = tag[i] - 1;
      if (t < 0 || t >= r->maxtag) continue;
      const int b = r->h_abody[t];
      if (b >= 0 && (!tgroup || tgroup[i])) nall[b]++;
    }
    if (r->nccl) {  // MPI_Allreduce(ncount, nall) of FixRigid::dof (fix_rigid.cpp:1221)
      CUDA_CHECK(cudaSetDevice(r->device));
      DBuf<int> d;
      d.ensure(r->nbody);
      CUDA_CHECK(cudaMemcpyAsync(d.p, nall.data(), (size_t)r->nbody * sizeof(int), cudaMemcpyHostToDevice, r->stream));
      const ncclResult_t nr = g_rigid_nccl.AllReduce(d.p, d.p, r->nbody, ncclInt, ncclSum, r->nccl, r->stream);
      if (nr != ncclSuccess) throw CudaError{std::string("ncclAllReduce (rigid dof): ") + g_rigid_nccl.GetErrorString(nr)};
      CUDA_CHECK(cudaMemcpyAsync(nall.data(), d.p, (size_t)r->nbody * sizeof(int), cudaMemcpyDeviceToHost, r->stream));
      CUDA_CHECK(cudaStreamSynchronize(r->stream));
      d.release();
    }
    int n = 0, nlinear = 0;
    for (int b = 0; b < r->nbody; b++)
      if (nall[b] == r->h_nrigid[b]) {
        n += 3 * nall[b] - 6;
        if (r->h_linear[b]) { n++; nlinear++; }
      }
    r->nlinear = nlinear;
    *dof = n;
  });
}

int polb200_rigid_setup(polb200_rigid_t *r, const polb200_rigid_atoms *a, int vflag)
{
  if (!r || !a) return POLB200_ERR_ARG;
  return rigid_guarded(r, [&] {
    if (!r->ready) throw StyleError{POLB200_ERR_STATE, "polb200_rigid_init has not been called"};
    CUDA_CHECK(cudaSetDevice(r->device));
    const int n = a->nlocal;
    check_atom_count(r, n);
    CUDA_CHECK(cudaEventRecord(r->ev[0], r->stream));
    StepArrays s = stage_in(r, a, true, true);
    launch_bodies<MODE_SETUP>(r, s, n);
    r->evflag = vflag ? 1 : 0;
    launch_atoms<0>(r, s, n, 0.0, 2.0);  // set_v; "guesstimate virial as 2x the set_v contribution" (fix_rigid.cpp:882-888)
    if (r->rc.tstat) {
      k_rigid_nhc<1><<<1, 1024, 0, r->stream>>>(r->nbody, r->rc, r->par.t_start, r->akin.p, r->chain.p);
      RIGID_LAUNCHED(r);
    }
    stage_out(r, a, false);
    r->setup_done = true;
  });
}

int polb200_rigid_initial_integrate(polb200_rigid_t *r, const polb200_rigid_atoms *a, int vflag, double run_fraction)
{
  if (!r || !a) return POLB200_ERR_ARG;
  return rigid_guarded(r, [&] {
    if (!r->setup_done) throw StyleError{POLB200_ERR_STATE, "polb200_rigid_setup has not been called"};
    CUDA_CHECK(cudaSetDevice(r->device));
    const int n = a->nlocal;
    check_atom_count(r, n);
    CUDA_CHECK(cudaEventRecord(r->ev[0], r->stream));
    r->evflag = vflag ? 1 : 0;
    StepArrays s = stage_in(r, a, true, r->evflag != 0);
    k_rigid_initial<<<cdiv(r->nbody, 128), 128, 0, r->stream>>>(r->nbody, r->rc, r->chain.p, r->frame.p, r->dyn.p, r->akin.p);
    RIGID_LAUNCHED(r);
    if (r->rc.tstat) {
      const double t_target = r->par.t_start + run_fraction * (r->par.t_stop - r->par.t_start);  // fix_rigid_nh.cpp:1109-1115
      k_rigid_nhc<0><<<1, 1024, 0, r->stream>>>(r->nbody, r->rc, t_target, r->akin.p, r->chain.p);
      RIGID_LAUNCHED(r);
    }
    launch_atoms<1>(r, s, n, 0.0, 1.0);
    stage_out(r, a, true);
  });
}

int polb200_rigid_final_integrate(polb200_rigid_t *r, const polb200_rigid_atoms *a)
{
  if (!r || !a) return POLB200_ERR_ARG;
  return rigid_guarded(r, [&] {
    if (!r->setup_done) throw StyleError{POLB200_ERR_STATE, "polb200_rigid_setup has not been called"};
    CUDA_CHECK(cudaSetDevice(r->device));
    const int n = a->nlocal;
    check_atom_count(r, n);
    CUDA_CHECK(cudaEventRecord(r->ev[0], r->stream));
    StepArrays s = stage_in(r, a, true, true);
    launch_bodies<MODE_FINAL>(r, s, n);
    launch_atoms<0>(r, s, n, 1.0, 1.0);
    stage_out(r, a, false);
  });
}

int polb200_rigid_pre_neighbor(polb200_rigid_t *r, int nlocal, const int *tag, const int *image, int on_device)
{
  if (!r || nlocal < 0 || (nlocal > 0 && (!tag || !image))) return POLB200_ERR_ARG;
  return rigid_guarded(r, [&] {
    if (!r->ready) throw StyleError{POLB200_ERR_STATE, "polb200_rigid_init has not been called"};
    CUDA_CHECK(cudaSetDevice(r->device));
    const int n = nlocal;
    const int *dt = tag, *di = image;
    if (!on_device) {
      r->c_tag.ensure(n); r->c_image.ensure(n);
      CUDA_CHECK(cudaMemcpyAsync(r->c_tag.p, tag, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, r->stream));
      CUDA_CHECK(cudaMemcpyAsync(r->c_image.p, image, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, r->stream));
      dt = r->c_tag.p; di = r->c_image.p;
    }
    k_rigid_remap<<<cdiv(r->nbody, 256), 256, 0, r->stream>>>(r->nbody, r->rc, r->frame.p, r->imagebody.p);
    RIGID_LAUNCHED(r);
    if (n > 0) {
      k_rigid_image_shift<<<cdiv(n, 256), 256, 0, r->stream>>>(n, dt, di, r->abody.p, r->imagebody.p, r->xcmimage.p);
      RIGID_LAUNCHED(r);
    }
    CUDA_CHECK(cudaStreamSynchronize(r->stream));
  });
}

int polb200_rigid_virial(polb200_rigid_t *r, double virial[6])
{
  if (!r || !virial) return POLB200_ERR_ARG;
  return rigid_guarded(r, [&] {
    if (!r->ready) throw StyleError{POLB200_ERR_STATE, "polb200_rigid_init has not been called"};
    CUDA_CHECK(cudaSetDevice(r->device));
    CUDA_CHECK(cudaMemcpyAsync(virial, r->virial.p, 6 * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    CUDA_CHECK(cudaStreamSynchronize(r->stream));
  });
}

int polb200_rigid_scalar(polb200_rigid_t *r, double *scalar, double *ke_t, double *ke_r)
{
  if (!r) return POLB200_ERR_ARG;
  return rigid_guarded(r, [&] {
    if (!r->ready) throw StyleError{POLB200_ERR_STATE, "polb200_rigid_init has not been called"};
    CUDA_CHECK(cudaSetDevice(r->device));
    DBuf<double2> &tmp = r->akin;  // free between steps: every step overwrites it before reading
    k_rigid_kinetic<<<cdiv(r->nbody, 256), 256, 0, r->stream>>>(r->nbody, r->frame.p, r->dyn.p, tmp.p);
    RIGID_LAUNCHED(r);
    k_rigid_sum2<<<1, 1024, 0, r->stream>>>(r->nbody, tmp.p, r->sums.p);
    RIGID_LAUNCHED(r);
    double h[2];
    Chain c;
    CUDA_CHECK(cudaMemcpyAsync(h, r->sums.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    CUDA_CHECK(cudaMemcpyAsync(&c, r->chain.p, sizeof(Chain), cudaMemcpyDeviceToHost, r->stream));
    CUDA_CHECK(cudaStreamSynchronize(r->stream));
    if (ke_t) *ke_t = 0.5 * h[0];
    if (ke_r) *ke_r = 0.5 * h[1];
    if (scalar) {
      double e = (h[0] + h[1]) * r->tfactor;  // FixRigid::compute_scalar (a temperature), fix_rigid.cpp:2595-2622
      if (r->rc.tstat) {                      // FixRigidNH::compute_scalar adds the chain energies to it, fix_rigid_nh.cpp:991-1016
        const double kt = r->rc.boltz * c.t_target;
        const int nc = r->rc.t_chain;
        e += kt * (r->rc.nf_t * c.eta_t[0] + r->rc.nf_r * c.eta_r[0]);
        for (int i = 1; i < nc; i++) e += kt * (c.eta_t[i] + c.eta_r[i]);
        for (int i = 0; i < nc; i++) {
          e += 0.5 * c.q_t[i] * (c.eta_dot_t[i] * c.eta_dot_t[i]);
          e += 0.5 * c.q_r[i] * (c.eta_dot_r[i] * c.eta_dot_r[i]);
        }
      }
      *scalar = e;
    }
  });
}

int polb200_rigid_get_chain(polb200_rigid_t *r, double *state, int capacity, int *t_chain)
{
  if (!r || !state || !t_chain) return POLB200_ERR_ARG;
  return rigid_guarded(r, [&] {
    if (!r->ready) throw StyleError{POLB200_ERR_STATE, "polb200_rigid_init has not been called"};
    const int nc = r->rc.tstat ? r->rc.t_chain : 0;
    *t_chain = nc;
    if (capacity < 4 * nc) throw StyleError{POLB200_ERR_ARG, "polb200_rigid_get_chain: buffer too small"};
    if (!nc) return;
    CUDA_CHECK(cudaSetDevice(r->device));
    Chain c;
    CUDA_CHECK(cudaMemcpyAsync(&c, r->chain.p, sizeof(Chain), cudaMemcpyDeviceToHost, r->stream));
    CUDA_CHECK(cudaStreamSynchronize(r->stream));
    for (int i = 0; i < nc; i++) {
      state[4 * i] = c.eta_t[i];
      state[4 * i + 1] = c.eta_r[i];
      state[4 * i + 2] = c.eta_dot_t[i];
      state[4 * i + 3] = c.eta_dot_r[i];
    }
  });
}

int polb200_rigid_set_chain(polb200_rigid_t *r, const double *state, int t_chain)
{
  if (!r || !state) return POLB200_ERR_ARG;
  return rigid_guarded(r, [&] {
    if (!r->ready) throw StyleError{POLB200_ERR_STATE, "polb200_rigid_init has not been called"};
    if (!r->rc.tstat || t_chain != r->rc.t_chain) return;  // the reference skips a record of another chain length too
    CUDA_CHECK(cudaSetDevice(r->device));
    Chain c;
    CUDA_CHECK(cudaMemcpyAsync(&c, r->chain.p, sizeof(Chain), cudaMemcpyDeviceToHost, r->stream));
    CUDA_CHECK(cudaStreamSynchronize(r->stream));
    for (int i = 0; i < t_chain; i++) {
      c.eta_t[i] = state[4 * i];
      c.eta_r[i] = state[4 * i + 1];
      c.eta_dot_t[i] = state[4 * i + 2];
      c.eta_dot_r[i] = state[4 * i + 3];
    }
    CUDA_CHECK(cudaMemcpyAsync(r->chain.p, &c, sizeof(Chain), cudaMemcpyHostToDevice, r->stream));
    CUDA_CHECK(cudaStreamSynchronize(r->stream));
  });
}

int polb200_rigid_comm_init(polb200_rigid_t *r, int rank, int nranks, const void *id_bytes)
{
  if (!r || !id_bytes || nranks < 1 || rank < 0 || rank >= nranks) return POLB200_ERR_ARG;
  return rigid_guarded(r, [&] {
    if (r->nccl) throw StyleError{POLB200_ERR_STATE, "polb200_rigid_comm_init was already called"};
    if (r->ready) throw StyleError{POLB200_ERR_STATE, "polb200_rigid_comm_init must precede polb200_rigid_init"};
    std::string err;
    if (!g_rigid_nccl.load(err)) throw StyleError{POLB200_ERR_UNSUPPORTED, err};
    CUDA_CHECK(cudaSetDevice(r->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    const ncclResult_t rc = g_rigid_nccl.CommInitRank(&r->nccl, nranks, id, rank);
    if (rc != ncclSuccess) throw CudaError{std::string("ncclCommInitRank: ") + g_rigid_nccl.GetErrorString(rc)};
    r->rank = rank;
    r->nranks = nranks;
  });
}

int polb200_rigid_reset_dt(polb200_rigid_t *r, double dt)
{
  if (!r) return POLB200_ERR_ARG;
  return rigid_guarded(r, [&] {
    r->par.dt = dt;
    set_timestep(r, dt);
  });
}

long polb200_rigid_fetch(polb200_rigid_t *r, const char *name, double *dst, long capacity)
{
  if (!r || !name || !dst || !r->ready) return -1;
  try {
    CUDA_CHECK(cudaSetDevice(r->device));
    const int nb = r->nbody;
    std::vector<BodyFrame> F(nb);
    std::vector<BodyDyn> D(nb);
    CUDA_CHECK(cudaMemcpyAsync(F.data(), r->frame.p, sizeof(BodyFrame) * nb, cudaMemcpyDeviceToHost, r->stream));
    CUDA_CHECK(cudaMemcpyAsync(D.data(), r->dyn.p, sizeof(BodyDyn) * nb, cudaMemcpyDeviceToHost, r->stream));
    CUDA_CHECK(cudaStreamSynchronize(r->stream));
    struct Field { const char *name; int frame, off, len; };
    const Field fields[] = {{"xcm", 1, 0, 3}, {"vcm", 1, 3, 3}, {"omega", 1, 6, 3}, {"ex", 1, 9, 3}, {"ey", 1, 12, 3}, {"ez", 1, 15, 3},
                            {"fcm", 0, 0, 3}, {"torque", 0, 3, 3}, {"angmom", 0, 6, 3}, {"quat", 0, 9, 4}, {"conjqm", 0, 13, 4},
                            {"inertia", 0, 17, 3}, {"masstotal", 0, 20, 1}};
    for (const Field &fd : fields)
      if (!strcmp(fd.name, name)) {
        if ((long)nb * fd.len > capacity) return -2;
        for (int b = 0; b < nb; b++) {
          const double *src = fd.frame ? reinterpret_cast<const double *>(&F[b]) : reinterpret_cast<const double *>(&D[b]);
          for (int k = 0; k < fd.len; k++) dst[(size_t)b * fd.len + k] = src[fd.off + k];
        }
        return (long)nb * fd.len;
      }
    return -3;
  } catch (const CudaError &x) {
    r->err = x.msg;
    return -4;
  }
}

long polb200_rigid_launch_count(polb200_rigid_t *r, int reset)
{
  if (!r) return -1;
  const long n = r->launches;
  if (reset) r->launches = 0;
  return n;
}

double polb200_rigid_last_ms(const polb200_rigid_t *r) { return r ? (double)r->ms_last : 0.0; }

}  // extern "C"
