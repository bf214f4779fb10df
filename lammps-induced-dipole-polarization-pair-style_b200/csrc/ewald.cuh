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

// ---------------------------------------------------------------------------------------------------
// Round 2: column form.  ncu had the two kernels above at 19 % / 28 % of the FP64 pipe (shared-memory data pipe and, in the
// force kernel, 4 resident warps per SM because every thread kept three phase rows in shared memory).  The k set is now
// stored by COLUMNS: a column = one (ky,kz) with its contiguous run of valid kx, kx = lo..hi (the half-space rule only
// decides whether lo is 0 or 1, the sphere |k|^2 <= gsqmx bounds hi).  S, ug and the packed (kx,ky,kz) of the nk real
// k-vectors live in flat arrays, column after column (ky outer, kz inner), no padding slots.
//   k_ewald_sfac_col : one THREAD per COLS columns, all kx of the column(s) in registers (2*NKX doubles per column);
//                      per atom the thread forms q Ey Ez once (two per-lane LDS.128) and then walks kx with the atom's
//                      Ex[kx] read as a warp BROADCAST (one wavefront) -> 4 FMA per k-vector, FP64-bound.  Blocks of
//                      4 kx above the largest `hi` of the warp's columns are skipped (the sphere, not the box).
//   k_ewald_force_col: one THREAD per atom, NOTHING per thread in shared memory: exp(i kx ux x), exp(i ky uy y) and
//                      exp(i kz uz z) advance by complex recurrences in registers (the same recurrence k_ewald_phase
//                      builds its table with) while W(k) = 2 ug(k) S(k) streams through shared memory as a broadcast.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

struct EwaldCol {   // one (ky,kz)
  short ky, kz;
  short lo, hi;     // valid kx = lo..hi (hi < lo: no k-vector in this column)
  int off;          // index of (lo,ky,kz) in the flat arrays
  int pad;
};

constexpr int EWC_THREADS = 128;   // threads per CTA of both column kernels
constexpr int EWC_TILE = 32;       // atoms per shared-memory tile of k_ewald_sfac_col

// Spart[slice][k]; grid = (ceil(nvalid / (EWC_THREADS * COLS)), nslices); vcol = the columns that hold k-vectors
template <int NKX, int COLS>
__global__ void __launch_bounds__(EWC_THREADS)
k_ewald_sfac_col(int n, int nvalid, int nk, const EwaldCol *__restrict__ vcol, const double *__restrict__ q,
                 const double2 *__restrict__ phase, int slots, int kxbase, double2 *__restrict__ Spart)
{
  // two tile buffers, each [EWC_TILE][NKX] Ex (kxbase..), [EWC_TILE][slots] Ey, [EWC_TILE][slots] Ez, EWC_TILE charges (as double2.x);
  // the next tile is fetched with cp.async while the current one is consumed (fetching a tile is ~96 dependent-latency row
  // reads: without the overlap the kernel spent most of its time waiting for them)
  extern __shared__ double2 sh[];
  const int tile_elems = EWC_TILE * NKX + 2 * EWC_TILE * slots + EWC_TILE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int ay[COLS], az[COLS], hi[COLS], lo[COLS], off[COLS];
  bool ny[COLS], nz[COLS];
  int whi = -1;
#pragma unroll
  for (int c = 0; c < COLS; c++) {
    const int v = (blockIdx.x * EWC_THREADS + threadIdx.x) * COLS + c;
    ay[c] = az[c] = 0; lo[c] = 0; hi[c] = -1; off[c] = 0; ny[c] = nz[c] = false;
    if (v < nvalid) {
      const EwaldCol col = vcol[v];
      ny[c] = col.ky < 0; nz[c] = col.kz < 0;
      ay[c] = ny[c] ? -col.ky : col.ky; az[c] = nz[c] ? -col.kz : col.kz;
      lo[c] = col.lo; hi[c] = col.hi; off[c] = col.off;
    }
    whi = max(whi, hi[c]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) whi = max(whi, __shfl_xor_sync(0xffffffffu, whi, o));
  whi -= kxbase;   // largest kx (relative to this pass) any column of the warp needs
  const int per = (n + gridDim.y - 1) / gridDim.y;
  const int i0 = blockIdx.y * per, i1 = min(n, i0 + per);
  double2 acc[COLS][NKX];
#pragma unroll
  for (int c = 0; c < COLS; c++)
#pragma unroll
    for (int k = 0; k < NKX; k++) acc[c][k] = make_double2(0.0, 0.0);
  auto fetch = [&](int buf, int base) {   // asynchronous copy of the tile of atoms base.. into buffer `buf`
    const int cnt = min(EWC_TILE, i1 - base);
    double2 *bx = sh + (size_t)buf * tile_elems, *by = bx + EWC_TILE * NKX, *bz = by + (size_t)EWC_TILE * slots;
    double2 *bq = bz + (size_t)EWC_TILE * slots;
    for (int r = warp; r < 3 * EWC_TILE; r += EWC_THREADS / 32) {   // one phase row per warp trip
      const int d = r / EWC_TILE, t = r - d * EWC_TILE;
      const double2 *src = phase + ((size_t)d * n + base + t) * slots;
      if (d == 0) {
        for (int m = lane; m < NKX; m += 32) {
          if (t < cnt && kxbase + m < slots) cp_async16(bx + t * NKX + m, src + kxbase + m);
          else bx[t * NKX + m] = make_double2(0.0, 0.0);
        }
      } else {
        double2 *dst = (d == 1 ? by : bz) + (size_t)t * slots;
        for (int m = lane; m < slots; m += 32) {
          if (t < cnt) cp_async16(dst + m, src + m);
          else dst[m] = make_double2(0.0, 0.0);
        }
      }
    }
    if (threadIdx.x < EWC_TILE) bq[threadIdx.x] = make_double2(threadIdx.x < cnt ? q[base + threadIdx.x] : 0.0, 0.0);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (i0 < i1) fetch(0, i0);
  int buf = 0;
  for (int base = i0; base < i1; base += EWC_TILE, buf ^= 1) {
    const int cnt = min(EWC_TILE, i1 - base);
    if (base + EWC_TILE < i1) {
      fetch(buf ^ 1, base + EWC_TILE);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const double2 *sx = sh + (size_t)buf * tile_elems, *sy = sx + EWC_TILE * NKX, *sz = sy + (size_t)EWC_TILE * slots;
    const double2 *sq = sz + (size_t)EWC_TILE * slots;
    for (int t = 0; t < cnt; t++) {
      const double qi = sq[t].x;
      double2 p[COLS];
#pragma unroll
      for (int c = 0; c < COLS; c++) {
        const double2 ey = conj_if(sy[(size_t)t * slots + ay[c]], ny[c]);
        const double2 ez = conj_if(sz[(size_t)t * slots + az[c]], nz[c]);
        p[c] = cmul(ey, ez);
        p[c].x *= qi;
        p[c].y *= qi;
      }
      const double2 *ex = sx + t * NKX;
      // the four Ex of a block are read one block ahead of their use (same address in every lane: broadcasts); without
      // that every block began by waiting for its shared-memory reads
      double2 ecur[4];
#pragma unroll
      for (int k = 0; k < 4; k++) ecur[k] = ex[k];
#pragma unroll
      for (int kb = 0; kb < NKX; kb += 4) {
        if (kb > whi) break;   // warp-uniform
        double2 enext[4];
#pragma unroll
        for (int k = 0; k < 4; k++) enext[k] = (kb + 4 < NKX && NKX * COLS <= 32) ? ex[kb + 4 + k] : make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          // (with more than 32 complex accumulators per thread there is no register left for the look-ahead)
          const double2 e = NKX * COLS <= 32 ? ecur[k] : ex[kb + k];
#pragma unroll
          for (int c = 0; c < COLS; c++) {
            acc[c][kb + k].x = fma(e.x, p[c].x, fma(-e.y, p[c].y, acc[c][kb + k].x));
            acc[c][kb + k].y = fma(e.x, p[c].y, fma(e.y, p[c].x, acc[c][kb + k].y));
          }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) ecur[k] = enext[k];
      }
    }
    __syncthreads();   // the next iteration's prefetch overwrites this buffer's sibling, the one after it this buffer
  }
  // every (slice, k) owns its slot: no atomics; k_ewald_sum_slices adds the slices in a fixed order
#pragma unroll
  for (int c = 0; c < COLS; c++)
#pragma unroll
    for (int k = 0; k < NKX; k++) {
      const int kx = kxbase + k;
      if (kx >= lo[c] && kx <= hi[c]) Spart[(size_t)blockIdx.y * nk + off[c] + (kx - lo[c])] = acc[c][k];
    }
}

// out[g][t] = sum of in[s][t] over the slices s of group g (ascending); grid.y = number of groups.  Two levels of this
// kernel add a few hundred slices with every SM busy and the order of the additions fixed.
__global__ void k_ewald_sum_groups(int nslots, int nslices, int per_group, const double2 *__restrict__ in, double2 *__restrict__ out)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nslots) return;
  const int s0 = blockIdx.y * per_group, s1 = min(nslices, s0 + per_group);
  double2 a = make_double2(0.0, 0.0);
#pragma unroll 8
  for (int sidx = s0; sidx < s1; sidx++) {
    const double2 v = in[(size_t)sidx * nslots + t];
    a.x += v.x;
    a.y += v.y;
  }
  out[(size_t)blockIdx.y * nslots + t] = a;
}

// W(k) = 2 ug(k) S(k)
__global__ void k_ewald_w(int nk, const double *__restrict__ ug, const double2 *__restrict__ S, double2 *__restrict__ W)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nk) return;
  const double u = 2.0 * ug[t];
  const double2 s = S[t];
  W[t] = make_double2(u * s.x, u * s.y);
}

// ek_i = sum_k k Im(exp(i k.r_i) conj W(k));  f_i += qscale q_i ek_i   (ewald.cpp:417-449).  cols = the full
// (2 kmax + 1)^2 table, ky outer, kz inner; one ky row of columns and its W values per shared-memory tile.
__global__ void __launch_bounds__(EWC_THREADS)
k_ewald_force_col(int n, int kmax, const EwaldCol *__restrict__ cols, const double2 *__restrict__ W, const double *__restrict__ q,
                  const double2 *__restrict__ phase, int slots, double ux, double uy, double uz, double qscale,
                  double *__restrict__ f, double *__restrict__ fpart)
{
  extern __shared__ double2 sh[];   // [(2 kmax + 1) * (kmax + 1)] W of one ky row, then its 2 kmax + 1 columns
  const int side = 2 * kmax + 1;
  double2 *sw = sh;
  EwaldCol *sc = reinterpret_cast<EwaldCol *>(sh + (size_t)side * (kmax + 1));
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  const int ii = live ? i : 0;
  const double2 ex1 = phase[((size_t)0 * n + ii) * slots + 1];
  const double2 ey1 = phase[((size_t)1 * n + ii) * slots + 1];
  const double2 ez1 = phase[((size_t)2 * n + ii) * slots + 1];
  const double2 ez_first = conj_if(phase[((size_t)2 * n + ii) * slots + kmax], true);   // exp(-i kmax uz z)
  // small systems: the ky rows are split over gridDim.y so that every SM has work; slice s writes fpart[s][3 n] and
  // k_ewald_force_sum adds the slices in order (gridDim.y == 1: straight into f)
  const int rows_per = (side + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(side, r0 + rows_per);
  const int ky0 = r0 - kmax;
  double2 ey = conj_if(phase[((size_t)1 * n + ii) * slots + (ky0 < 0 ? -ky0 : ky0)], ky0 < 0);   // exp(i ky0 uy y)
  double ekx = 0.0, eky = 0.0, ekz = 0.0;
  for (int r = r0; r < r1; r++) {
    const EwaldCol *row = cols + (size_t)r * side;
    // the W values of this ky row are contiguous: from the first column's offset to the end of the last one
    const int w0 = row[0].off, w1 = row[side - 1].off + max(0, row[side - 1].hi - row[side - 1].lo + 1);
    __syncthreads();
    for (int e = threadIdx.x; e < side; e += blockDim.x) sc[e] = row[e];
    for (int e = threadIdx.x; e < w1 - w0; e += blockDim.x) sw[e] = W[w0 + e];
    __syncthreads();
    if (w1 > w0) {
      double2 ez = ez_first;
      for (int c = 0; c < side; c++) {
        const EwaldCol col = sc[c];   // broadcast
        if (col.hi >= col.lo) {
          const double2 p = cmul(ey, ez);
          double2 e3 = col.lo == 0 ? p : cmul(p, ex1);
          const double2 *w = sw + (col.off - w0);
          double sx = 0.0, sall = 0.0, kxd = (double)col.lo;
          for (int kx = col.lo; kx <= col.hi; kx++) {
            const double2 wk = w[kx - col.lo];   // broadcast
            const double partial = fma(e3.y, wk.x, -e3.x * wk.y);
            sx = fma(partial, kxd, sx);
            sall += partial;
            kxd += 1.0;
            e3 = cmul(e3, ex1);
          }
          ekx += sx;
          eky = fma(sall, (double)col.ky, eky);
          ekz = fma(sall, (double)col.kz, ekz);
        }
        ez = cmul(ez, ez1);
      }
    }
    ey = cmul(ey, ey1);
  }
  if (live) {
    const double c = qscale * q[i];
    if (gridDim.y == 1) {
      f[3 * i] += c * ux * ekx;
      f[3 * i + 1] += c * uy * eky;
      f[3 * i + 2] += c * uz * ekz;
    } else {
      double *o = fpart + ((size_t)blockIdx.y * n + i) * 3;
      o[0] = c * ux * ekx;
      o[1] = c * uy * eky;
      o[2] = c * uz * ekz;
    }
  }
}

// f[t] += sum over the slices (ascending) of fpart[s][t], t over the 3 n force components
__global__ void k_ewald_force_sum(long n3, int nslices, const double *__restrict__ fpart, double *__restrict__ f)
{
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n3) return;
  double a = 0.0;
  for (int s = 0; s < nslices; s++) a += fpart[(size_t)s * n3 + t];
  f[t] += a;
}

// out[0] = sum ug |S|^2, out[1..6] = virial sums (xx yy zz xy xz yz), one CTA of 256 threads over the nk k-vectors
// (kxyz = kx | (ky + 512) << 10 | (kz + 512) << 20)
__global__ void k_ewald_energy_col(int nk, const double *__restrict__ ugv, const int *__restrict__ kxyz, const double2 *__restrict__ S,
                                   double ux, double uy, double uz, double ginv2, double *__restrict__ out)
{
  __shared__ double sm[256];
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int t = threadIdx.x; t < nk; t += 256) {
    const int code = kxyz[t];
    const double2 s = S[t];
    const double kxv = (code & 1023) * ux, kyv = (((code >> 10) & 1023) - 512) * uy, kzv = (((code >> 20) & 1023) - 512) * uz;
    const double sqk = kxv * kxv + kyv * kyv + kzv * kzv;
    const double uk = ugv[t] * (s.x * s.x + s.y * s.y);
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
