// ewald.cuh -- reciprocal-space Ewald on the device (the KSpace style every input of the
// pair style needs, src/KSPACE/ewald.cpp of the reference).  Included by engine.cu; C ABI: polb200_ewald_*.
//
// Same sums as Ewald::compute (ewald.cpp:357-497) over the same half-space k set (coeffs, :760-1026):
//   S(k)   = sum_i q_i exp(i k.r_i)                                   structure factors
//   E      = qqrd2e [ sum_k ug(k) |S(k)|^2 - g sum q^2/sqrt(pi) - pi (sum q)^2 / (2 g^2 V) ],  ug = 4 pi/V exp(-k^2/4g^2)/k^2
//   f_i   += qqrd2e q_i sum_k 2 ug(k) k Im( exp(i k.r_i) conj S(k) )
//   virial = qqrd2e sum_k ug |S|^2 (delta_ab - 2 (1/k^2 + 1/4g^2) k_a k_b)
// organised for the GPU instead of the reference's cs/sn tables walked k by k:
//   k_ewald_phase : per atom and dimension the powers exp(i m u_d x_d), m = 0..kmax_d (one sincos + a recurrence)
//   k_ewald_sfac  : one THREAD per QUAD of k-vectors (four consecutive kx, same ky,kz: the factor q Ey Ez is formed
//                   once per atom and feeds four complex FMAs), atoms streamed through shared-memory tiles of phase
//                   rows; a thread keeps its four S(k) in registers over its slice of atoms (no per-k reduction
//                   tree), slices are combined with atomicAdd.  Quads are ordered kx-fastest so that the lanes of a
//                   warp read neighbouring shared-memory slots (Ex) or broadcast (Ey, Ez).
//   k_ewald_force : one THREAD per atom with its three phase rows in shared memory (row stride = 1 mod 8 slots:
//                   conflict-free LDS.128), k-vectors and S(k) streamed through a broadcast tile.
//   k_ewald_energy: ug |S|^2 and the six virial sums, one CTA, fixed-order tree.
#pragma once

namespace polb200 {

struct EwaldK {   // FOUR half-space k-vectors (kx0 .. kx0+3, ky, kz): they share the (ky,kz) phase factor
  int kx0, ky, kz;
  int pad;
  double ug[4];   // 4 pi / V exp(-k^2 / 4 g^2) / k^2, or 0 for a slot outside the k set (padding of a row's last quad)
};

constexpr int EW_TILE = 32;      // atoms per shared-memory tile of the structure-factor kernel
constexpr int EW_KTHREADS = 128; // quads per CTA of the structure-factor kernel
constexpr int EW_ATHREADS = 128; // atoms per CTA of the force kernel
constexpr int EW_KTILE = 128;    // quads per broadcast tile of the force kernel

__host__ __device__ inline int ew_row_slots(int kmax)
{
  int s = kmax + 4;        // a padded quad may reach 3 slots past the largest kx
  while (s % 8 != 1) s++;  // row stride = 16 B mod 128 B: the 8 lanes of an LDS.128 wavefront hit distinct banks
  return s;
}

// phase[(d * n + i) * slots + m] = exp(i m u_d x_{i,d}),  m = 0 .. slots-1
__global__ void k_ewald_phase(int n, const double *__restrict__ x, double ux, double uy, double uz, int slots,
                              double2 *__restrict__ phase)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * n) return;
  const int d = t / n, i = t - d * n;
  const double u = d == 0 ? ux : (d == 1 ? uy : uz);
  double s, c;
  sincos(u * x[3 * i + d], &s, &c);
  double2 *row = phase + (size_t)t * slots;
  double2 cur = make_double2(1.0, 0.0);
  row[0] = cur;
  for (int m = 1; m < slots; m++) {
    cur = make_double2(cur.x * c - cur.y * s, cur.x * s + cur.y * c);
    row[m] = cur;
  }
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 conj_if(double2 a, bool neg) { return make_double2(a.x, neg ? -a.y : a.y); }

// grid = (ceil(nquads / EW_KTHREADS), nslices); S here = per-slice partial sums [nslices][4 * nquads].
// Register tile: per atom one product P = q Ey Ez serves the four kx of the quad.
__global__ void __launch_bounds__(EW_KTHREADS)
k_ewald_sfac(int n, int nquads, const EwaldK *__restrict__ kv, const double *__restrict__ q, const double2 *__restrict__ phase,
             int slots, double2 *__restrict__ S)
{
  extern __shared__ double2 sh[];  // [3][EW_TILE][slots] phase rows, then EW_TILE charges (as double2.x)
  double2 *sq = sh + (size_t)3 * EW_TILE * slots;
  const int k = blockIdx.x * EW_KTHREADS + threadIdx.x;
  int kx0 = 0, ky = 0, kz = 0;
  if (k < nquads) {
    kx0 = kv[k].kx0; ky = kv[k].ky; kz = kv[k].kz;
  }
  const bool ny = ky < 0, nz = kz < 0;
  const int ay = ny ? -ky : ky, az = nz ? -kz : kz;
  const int per = (n + gridDim.y - 1) / gridDim.y;
  const int i0 = blockIdx.y * per, i1 = min(n, i0 + per);
  double2 acc[4] = {make_double2(0, 0), make_double2(0, 0), make_double2(0, 0), make_double2(0, 0)};
  for (int base = i0; base < i1; base += EW_TILE) {
    const int cnt = min(EW_TILE, i1 - base);
    __syncthreads();
    for (int e = threadIdx.x; e < 3 * EW_TILE * slots; e += EW_KTHREADS) {
      const int d = e / (EW_TILE * slots), r = e - d * EW_TILE * slots, t = r / slots, m = r - t * slots;
      sh[e] = t < cnt ? phase[((size_t)d * n + base + t) * slots + m] : make_double2(0, 0);
    }
    if (threadIdx.x < EW_TILE) sq[threadIdx.x] = make_double2(threadIdx.x < cnt ? q[base + threadIdx.x] : 0.0, 0.0);
    __syncthreads();
    for (int t = 0; t < cnt; t++) {
      const double2 ey = conj_if(sh[(size_t)(1 * EW_TILE + t) * slots + ay], ny);
      const double2 ez = conj_if(sh[(size_t)(2 * EW_TILE + t) * slots + az], nz);
      double2 p = cmul(ey, ez);
      const double qi = sq[t].x;
      p.x *= qi;
      p.y *= qi;
      const double2 *ex = sh + (size_t)t * slots + kx0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const double2 e = ex[j];
        acc[j].x = fma(e.x, p.x, fma(-e.y, p.y, acc[j].x));
        acc[j].y = fma(e.x, p.y, fma(e.y, p.x, acc[j].y));
      }
    }
  }
  // every (slice, quad) owns its slot: no atomics, and k_ewald_sum_slices adds the slices in a fixed order, so the
  // structure factors (and with them energy and forces) are bit-reproducible run to run
  if (k < nquads)
#pragma unroll
    for (int j = 0; j < 4; j++) S[((size_t)blockIdx.y * nquads + k) * 4 + j] = acc[j];
}

// S[k] = sum over slices of Spart[slice][k], slices in ascending order
__global__ void k_ewald_sum_slices(int nslots, int nslices, const double2 *__restrict__ Spart, double2 *__restrict__ S)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nslots) return;
  double2 a = make_double2(0.0, 0.0);
  for (int s = 0; s < nslices; s++) {
    const double2 v = Spart[(size_t)s * nslots + t];
    a.x += v.x;
    a.y += v.y;
  }
  S[t] = a;
}

// ek_i = sum_k 2 ug k Im(exp(i k.r_i) conj S(k));  f_i += qscale q_i ek_i   (ewald.cpp:417-449)
__global__ void __launch_bounds__(EW_ATHREADS)  // launched with up to EW_ATHREADS threads (fewer when kmax is large)
k_ewald_force(int n, int nquads, const EwaldK *__restrict__ kv, const double2 *__restrict__ S, const double *__restrict__ q,
              const double2 *__restrict__ phase, int slots, double ux, double uy, double uz, double qscale,
              double *__restrict__ f)
{
  extern __shared__ double2 sh[];  // [blockDim.x][3][slots] own phase rows, then the quad tile: EW_KTILE x {quad | 4 S}
  double2 *rows = sh;
  EwaldK *tk = reinterpret_cast<EwaldK *>(sh + (size_t)blockDim.x * 3 * slots);
  double2 *ts = reinterpret_cast<double2 *>(tk + EW_KTILE);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double2 *mine = rows + (size_t)threadIdx.x * 3 * slots;
  if (i < n)
    for (int d = 0; d < 3; d++)
      for (int m = 0; m < slots; m++) mine[d * slots + m] = phase[((size_t)d * n + i) * slots + m];
  double ekx = 0, eky = 0, ekz = 0;
  for (int k0 = 0; k0 < nquads; k0 += EW_KTILE) {
    const int cnt = min(EW_KTILE, nquads - k0);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) tk[e] = kv[k0 + e];
    for (int e = threadIdx.x; e < 4 * cnt; e += blockDim.x) ts[e] = S[4 * (size_t)k0 + e];
    __syncthreads();
    if (i < n)
      for (int e = 0; e < cnt; e++) {
        const EwaldK kk = tk[e];
        const double2 ey = conj_if(mine[slots + (kk.ky < 0 ? -kk.ky : kk.ky)], kk.ky < 0);
        const double2 ez = conj_if(mine[2 * slots + (kk.kz < 0 ? -kk.kz : kk.kz)], kk.kz < 0);
        const double2 p = cmul(ey, ez);
        double sx = 0.0, sall = 0.0;  // sum_j partial_j * kx_j  and  sum_j partial_j
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const double2 e3 = cmul(mine[kk.kx0 + j], p);
          const double2 s = ts[4 * e + j];
          const double partial = 2.0 * kk.ug[j] * (e3.y * s.x - e3.x * s.y);
          sx = fma(partial, (double)(kk.kx0 + j), sx);
          sall += partial;
        }
        ekx += sx;
        eky = fma(sall, (double)kk.ky, eky);
        ekz = fma(sall, (double)kk.kz, ekz);
      }
  }
  if (i < n) {
    const double c = qscale * q[i];
    f[3 * i] += c * ux * ekx;
    f[3 * i + 1] += c * uy * eky;
    f[3 * i + 2] += c * uz * ekz;
  }
}

// out[0] = sum ug |S|^2, out[1..6] = virial sums (xx yy zz xy xz yz), one CTA of 256 threads over 4 * nquads slots
__global__ void k_ewald_energy(int nquads, const EwaldK *__restrict__ kv, const double2 *__restrict__ S, double ux, double uy,
                               double uz, double ginv2, double *__restrict__ out)
{
  __shared__ double sm[256];
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int t = threadIdx.x; t < 4 * nquads; t += 256) {
    const EwaldK kk = kv[t >> 2];
    const int j = t & 3;
    if (kk.ug[j] == 0.0) continue;
    const double2 s = S[t];
    const double kxv = (kk.kx0 + j) * ux, kyv = kk.ky * uy, kzv = kk.kz * uz;
    const double sqk = kxv * kxv + kyv * kyv + kzv * kzv;
    const double uk = kk.ug[j] * (s.x * s.x + s.y * s.y);
    const double vterm = -2.0 * (1.0 / sqk + 0.25 * ginv2);
    acc[0] += uk;
    acc[1] += uk * (1.0 + vterm * kxv * kxv);
    acc[2] += uk * (1.0 + vterm * kyv * kyv);
    acc[3] += uk * (1.0 + vterm * kzv * kzv);
    acc[4] += uk * vterm * kxv * kyv;
    acc[5] += uk * vterm * kxv * kzv;
    acc[6] += uk * vterm * kyv * kzv;
  }
  for (int v = 0; v < 7; v++) {
    sm[threadIdx.x] = acc[v];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[v] = sm[0];
    __syncthreads();
  }
}

}  // namespace polb200
