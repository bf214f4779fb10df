// pppm.cu -- particle-particle particle-mesh reciprocal space on the device behind polb200_pppm_* (include/polb200.h):
// `kspace_style pppm <accuracy>` of the reference (src/KSPACE/pppm.cpp), ik differentiation, no stagger, orthogonal
// periodic box, one GPU.  Why: the device Ewald sum (ewald.cuh) is O(N^1.5) and costs more than the pair style at
// 256k atoms; PPPM is O(N log N) and what the reference itself offers for such sizes.
//
// File map
//   host plan      g_ewald estimate, grid from estimate_ik_error, factorable(2,3,5), Newton refinement of g_ewald,
//                  gf_b, assignment polynomials rho_coeff -- scalar arithmetic of PPPM::init, done once
//   k_pppm_gf      Hockney-Eastwood optimal influence function (compute_gf_ik), one thread per k-point
//   k_pppm_rho     charge assignment (particle_map + make_rho): one thread per atom, order^3 FP64 atomicAdd into the
//                  grid -- the only place in the library where the summation order is not fixed (order of the atomics)
//   k_pppm_poisson per k-point: energy / virial terms (block partials, fixed-order final sum), scale by the influence
//                  function, the three -i k phi fields (poisson_ik)
//   k_pppm_force   per atom: interpolate the three field grids with the same weights (fieldforce_ik)
//   FFTs           cuFFT Z2Z (a plain library FFT, dlopen'ed so that libpolb200.so keeps no link-time dependency): one
//                  e^{+ikr} transform of the charge grid, three e^{-ikr} transforms of the field grids, in place
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <cufft.h>
#include <dlfcn.h>

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

constexpr int PPPM_MAXORDER = 7;
constexpr int PPPM_OFFSET = 16384;  // pppm.cpp:49

struct CufftApi {
  void *lib = nullptr;
  cufftResult (*Plan3d)(cufftHandle *, int, int, int, cufftType) = nullptr;
  cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
  cufftResult (*ExecZ2Z)(cufftHandle, cufftDoubleComplex *, cufftDoubleComplex *, int) = nullptr;
  cufftResult (*Destroy)(cufftHandle) = nullptr;
  bool load(std::string &err)
  {
    if (lib) return true;
    const char *names[] = {"libcufft.so.11", "libcufft.so", "libcufft.so.12", "libcufft.so.10"};
    for (const char *n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      err = "cannot load libcufft.so (needed by kspace_style pppm on the device)";
      return false;
    }
    Plan3d = reinterpret_cast<decltype(Plan3d)>(dlsym(lib, "cufftPlan3d"));
    SetStream = reinterpret_cast<decltype(SetStream)>(dlsym(lib, "cufftSetStream"));
    ExecZ2Z = reinterpret_cast<decltype(ExecZ2Z)>(dlsym(lib, "cufftExecZ2Z"));
    Destroy = reinterpret_cast<decltype(Destroy)>(dlsym(lib, "cufftDestroy"));
    if (!Plan3d || !SetStream || !ExecZ2Z || !Destroy) {
      err = "libcufft.so lacks cufftPlan3d / cufftExecZ2Z";
      return false;
    }
    return true;
  }
};
static CufftApi g_cufft;
static NcclApi g_pppm_nccl;   // multi-GPU (polb200_pppm_comm_init): the charge grid is all-reduced, the rest is replicated

struct PppmConst {
  int order, nx, ny, nz, nlower, nupper;
  double shift, shiftone, delinv[3], delvolinv, boxlo[3], prd[3], unitk[3], g_ewald;
  double rho_coeff[PPPM_MAXORDER][PPPM_MAXORDER];  // [l][m - nlower]
  double gf_b[PPPM_MAXORDER];
  int nb[3];
};

__host__ __device__ inline double powsinxx(double x, int n)  // math_special.h:82-93
{
  if (x == 0.0) return 1.0;
  double ww = sin(x) / x, yy = 1.0;
  for (; n != 0; n >>= 1, ww *= ww)
    if (n & 1) yy *= ww;
  return yy;
}

__device__ __forceinline__ int per_of(int i, int n) { return i - n * (2 * i / n); }

// compute_gf_ik (pppm.cpp:1549-1627) + gf_denom (pppm.h:185-196); flat index n = (m*ny + l)*nx + k
__global__ void k_pppm_gf(PppmConst C, double *__restrict__ greensfn)
{
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)C.nx * C.ny * C.nz;
  if (idx >= total) return;
  const int k = (int)(idx % C.nx), l = (int)((idx / C.nx) % C.ny), m = (int)(idx / ((long)C.nx * C.ny));
  const int kper = per_of(k, C.nx), lper = per_of(l, C.ny), mper = per_of(m, C.nz);
  const double ux = C.unitk[0], uy = C.unitk[1], uz = C.unitk[2];
  const double sqk = (ux * kper) * (ux * kper) + (uy * lper) * (uy * lper) + (uz * mper) * (uz * mper);
  if (sqk == 0.0) {
    greensfn[idx] = 0.0;
    return;
  }
  const double s0 = sin(0.5 * ux * kper * C.prd[0] / C.nx), s1 = sin(0.5 * uy * lper * C.prd[1] / C.ny),
               s2 = sin(0.5 * uz * mper * C.prd[2] / C.nz);
  const double snx = s0 * s0, sny = s1 * s1, snz = s2 * s2;
  double sx = 0.0, sy = 0.0, sz = 0.0;
  for (int q = C.order - 1; q >= 0; q--) {
    sx = C.gf_b[q] + sx * snx;
    sy = C.gf_b[q] + sy * sny;
    sz = C.gf_b[q] + sz * snz;
  }
  const double sden = sx * sy * sz, denominator = sden * sden;
  const double numerator = 12.5663706 / sqk;
  const int twoorder = 2 * C.order;
  const double g = C.g_ewald;
  double sum1 = 0.0;
  for (int nx = -C.nb[0]; nx <= C.nb[0]; nx++) {
    const double qx = ux * (kper + C.nx * nx);
    const double ex = exp(-0.25 * (qx / g) * (qx / g));
    const double wx = powsinxx(0.5 * qx * C.prd[0] / C.nx, twoorder);
    for (int ny = -C.nb[1]; ny <= C.nb[1]; ny++) {
      const double qy = uy * (lper + C.ny * ny);
      const double ey = exp(-0.25 * (qy / g) * (qy / g));
      const double wy = powsinxx(0.5 * qy * C.prd[1] / C.ny, twoorder);
      for (int nz = -C.nb[2]; nz <= C.nb[2]; nz++) {
        const double qz = uz * (mper + C.nz * nz);
        const double ez = exp(-0.25 * (qz / g) * (qz / g));
        const double wz = powsinxx(0.5 * qz * C.prd[2] / C.nz, twoorder);
        const double dot1 = ux * kper * qx + uy * lper * qy + uz * mper * qz;
        const double dot2 = qx * qx + qy * qy + qz * qz;
        sum1 += (dot1 / dot2) * ex * ey * ez * wx * wy * wz;
      }
    }
  }
  greensfn[idx] = numerator * sum1 / denominator;
}

// particle_map + compute_rho1d for one atom (pppm.cpp:1907-1945, 2844-2863): grid origin and the order weights per dim
__device__ __forceinline__ void pppm_weights(const PppmConst &C, double x, double y, double z, int part[3],
                                             double w[3][PPPM_MAXORDER])
{
  const double xs[3] = {x, y, z};
  for (int d = 0; d < 3; d++) {
    const double u = (xs[d] - C.boxlo[d]) * C.delinv[d];
    part[d] = (int)(u + C.shift) - PPPM_OFFSET;
    const double dd = part[d] + C.shiftone - u;
    for (int k = 0; k < C.order; k++) {
      double r = 0.0;
      for (int l = C.order - 1; l >= 0; l--) r = C.rho_coeff[l][k] + r * dd;
      w[d][k] = r;
    }
  }
}

__device__ __forceinline__ int wrap(int i, int n)
{
  i %= n;
  return i < 0 ? i + n : i;
}

// make_rho (pppm.cpp:1951-1995); the ghost-cell fold of the reference's reverse_comm is the periodic wrap here
__global__ void k_pppm_rho(int n, PppmConst C, const double *__restrict__ x, const double *__restrict__ q, double2 *__restrict__ grid)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int part[3];
  double w[3][PPPM_MAXORDER];
  pppm_weights(C, x[3 * i], x[3 * i + 1], x[3 * i + 2], part, w);
  const double z0 = C.delvolinv * q[i];
  for (int c = 0; c < C.order; c++) {
    const int mz = wrap(part[2] + C.nlower + c, C.nz);
    const double y0 = z0 * w[2][c];
    for (int b = 0; b < C.order; b++) {
      const int my = wrap(part[1] + C.nlower + b, C.ny);
      const double x0 = y0 * w[1][b];
      for (int a = 0; a < C.order; a++) {
        const int mx = wrap(part[0] + C.nlower + a, C.nx);
        atomicAdd(&grid[((size_t)mz * C.ny + my) * C.nx + mx].x, x0 * w[0][a]);
      }
    }
  }
}

// ---- the same charge assignment without atomics: bit-reproducible (every sum has a fixed order) --------------------
// The atoms are sorted by the first grid point their stencil touches (stable radix sort: atoms of a cell stay in caller
// order), their order weights are evaluated once, and ONE THREAD PER GRID POINT gathers what the atoms of the order^3
// cells around it deposit there: cells in (z, y, x) order, atoms in sorted order.
constexpr int PPPM_WSTRIDE = 3 * PPPM_MAXORDER + 2;  // weights of the three dimensions, q * delvolinv, x index of the cell

__global__ void k_pppm_cellkey(int n, PppmConst C, const double *__restrict__ x, int *__restrict__ key, int *__restrict__ iota)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c[3];
  const double xs[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
  const int nn[3] = {C.nx, C.ny, C.nz};
  for (int d = 0; d < 3; d++) {
    const double u = (xs[d] - C.boxlo[d]) * C.delinv[d];
    c[d] = wrap((int)(u + C.shift) - PPPM_OFFSET + C.nlower, nn[d]);
  }
  key[i] = (c[2] * C.ny + c[1]) * C.nx + c[0];
  iota[i] = i;
}

__global__ void k_pppm_cellstart(int ncell, int n, const int *__restrict__ sorted_key, int *__restrict__ start)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > ncell) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (sorted_key[mid] < c) lo = mid + 1;
    else hi = mid;
  }
  start[c] = lo;
}

__global__ void k_pppm_weights_sorted(int n, PppmConst C, const int *__restrict__ order, const double *__restrict__ x,
                                      const double *__restrict__ q, double *__restrict__ wts)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int i = order[s];
  int part[3];
  double w[3][PPPM_MAXORDER];
  pppm_weights(C, x[3 * i], x[3 * i + 1], x[3 * i + 2], part, w);
  double *o = wts + (size_t)s * PPPM_WSTRIDE;
  for (int d = 0; d < 3; d++)
    for (int k = 0; k < PPPM_MAXORDER; k++) o[d * PPPM_MAXORDER + k] = k < C.order ? w[d][k] : 0.0;
  o[3 * PPPM_MAXORDER] = C.delvolinv * q[i];
  o[3 * PPPM_MAXORDER + 1] = (double)wrap(part[0] + C.nlower, C.nx);
}

__global__ void __launch_bounds__(256)
k_pppm_rho_gather(PppmConst C, const int *__restrict__ start, const double *__restrict__ wts, double2 *__restrict__ grid)
{
  // one WARP per grid point: lane = one (z, y) cell row of the order x order rows that reach the point; the partial sums
  // are combined by a fixed shuffle tree, so the result does not depend on anything but the data
  const size_t total = (size_t)C.nx * C.ny * C.nz;
  const size_t m = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= total) return;
  const int mx = (int)(m % C.nx), my = (int)((m / C.nx) % C.ny), mz = (int)(m / ((size_t)C.nx * C.ny));
  double sum = 0.0;
  // the cells of one (z, y) row that reach this point, cx = mx - order + 1 .. mx, are consecutive in the sorted order: one
  // (or, across the periodic wrap, two) contiguous run(s) of atoms per row instead of `order` cell lookups
  const int xlo = mx - C.order + 1;
  for (int r = lane; r < C.order * C.order; r += 32) {
    const int c = r / C.order, b = r - c * C.order;
    const int cz = wrap(mz - c, C.nz), cy = wrap(my - b, C.ny);
    const int row = (cz * C.ny + cy) * C.nx;
    for (int part = 0; part < 2; part++) {
      // part 0: cells [max(xlo,0), mx]; part 1 (only when xlo < 0): the wrapped cells [nx + xlo, nx - 1]
      if (part == 1 && xlo >= 0) break;
      const int c0 = part == 0 ? max(xlo, 0) : C.nx + xlo, c1 = part == 0 ? mx : C.nx - 1;
      const int s1 = start[row + c1 + 1];
      for (int s = start[row + c0]; s < s1; s++) {
        const double *o = wts + (size_t)s * PPPM_WSTRIDE;
        const int cx = (int)o[3 * PPPM_MAXORDER + 1];
        const int a = part == 0 ? mx - cx : mx + C.nx - cx;
        // the reference's product order (pppm.cpp:1981-1990): ((q/vol * wz) * wy) * wx
        sum += ((o[3 * PPPM_MAXORDER] * o[2 * PPPM_MAXORDER + c]) * o[PPPM_MAXORDER + b]) * o[a];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
  if (lane == 0) grid[m] = make_double2(sum, 0.0);
}

constexpr int NPPPM_PART = 7;  // energy + 6 virial terms

// poisson_ik (pppm.cpp:2032-2157) for one k-point, and the virial coefficients of setup (:455-480) on the fly
__global__ void __launch_bounds__(256) k_pppm_poisson(PppmConst C, const double *__restrict__ greensfn, double2 *__restrict__ work1,
                                                     double2 *__restrict__ wx, double2 *__restrict__ wy, double2 *__restrict__ wz,
                                                     int ev, double *__restrict__ partial)
{
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)C.nx * C.ny * C.nz;
  double acc[NPPPM_PART] = {0, 0, 0, 0, 0, 0, 0};
  if (idx < total) {
    const int k = (int)(idx % C.nx), l = (int)((idx / C.nx) % C.ny), m = (int)(idx / ((long)C.nx * C.ny));
    const double fkx = C.unitk[0] * per_of(k, C.nx), fky = C.unitk[1] * per_of(l, C.ny), fkz = C.unitk[2] * per_of(m, C.nz);
    const double scaleinv = 1.0 / ((double)C.nx * C.ny * C.nz);
    const double gf = greensfn[idx];
    double2 w = work1[idx];
    if (ev) {
      const double eng = scaleinv * scaleinv * gf * (w.x * w.x + w.y * w.y);
      const double sqk = fkx * fkx + fky * fky + fkz * fkz;
      acc[0] = eng;
      if (sqk != 0.0) {
        const double vterm = -2.0 * (1.0 / sqk + 0.25 / (C.g_ewald * C.g_ewald));
        acc[1] = eng * (1.0 + vterm * fkx * fkx);
        acc[2] = eng * (1.0 + vterm * fky * fky);
        acc[3] = eng * (1.0 + vterm * fkz * fkz);
        acc[4] = eng * vterm * fkx * fky;
        acc[5] = eng * vterm * fkx * fkz;
        acc[6] = eng * vterm * fky * fkz;
      }
    }
    w.x *= scaleinv * gf;
    w.y *= scaleinv * gf;
    wx[idx] = make_double2(fkx * w.y, -fkx * w.x);
    wy[idx] = make_double2(fky * w.y, -fky * w.x);
    wz[idx] = make_double2(fkz * w.y, -fkz * w.x);
  }
  if (ev) {
    __shared__ double sh[NPPPM_PART][8];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < NPPPM_PART; t++) {
      double s = acc[t];
      for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
      if (lane == 0) sh[t][wp] = s;
    }
    __syncthreads();
    if (threadIdx.x < NPPPM_PART) {
      double s = 0.0;
      for (int j = 0; j < 8; j++) s += sh[threadIdx.x][j];
      partial[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
    }
  }
}

// out[t] = sum of the block partials of term t, fixed order (one block, 7 warps)
__global__ void k_pppm_sum(int nblocks, const double *__restrict__ partial, double *__restrict__ out)
{
  const int t = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0.0;
  for (int j = lane; j < nblocks; j += 32) s += partial[(size_t)t * nblocks + j];
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if (lane == 0) out[t] = s;
}

// fieldforce_ik (pppm.cpp:2453-2505): f += qqrd2e * q * (-sum w * vd)
__global__ void k_pppm_force(int n, PppmConst C, double qqrd2e, const double *__restrict__ x, const double *__restrict__ q,
                             const double2 *__restrict__ vx, const double2 *__restrict__ vy, const double2 *__restrict__ vz,
                             double *__restrict__ f)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int part[3];
  double w[3][PPPM_MAXORDER];
  pppm_weights(C, x[3 * i], x[3 * i + 1], x[3 * i + 2], part, w);
  double ekx = 0.0, eky = 0.0, ekz = 0.0;
  for (int c = 0; c < C.order; c++) {
    const int mz = wrap(part[2] + C.nlower + c, C.nz);
    const double z0 = w[2][c];
    for (int b = 0; b < C.order; b++) {
      const int my = wrap(part[1] + C.nlower + b, C.ny);
      const double y0 = z0 * w[1][b];
      for (int a = 0; a < C.order; a++) {
        const int mx = wrap(part[0] + C.nlower + a, C.nx);
        const double x0 = y0 * w[0][a];
        const size_t g = ((size_t)mz * C.ny + my) * C.nx + mx;
        ekx -= x0 * vx[g].x;
        eky -= x0 * vy[g].x;
        ekz -= x0 * vz[g].x;
      }
    }
  }
  const double qf = qqrd2e * q[i];
  f[3 * i] += qf * ekx;
  f[3 * i + 1] += qf * eky;
  f[3 * i + 2] += qf * ekz;
}

// ---- host plan: the scalar part of PPPM::init ----------------------------------------------------------------

static const double PPPM_ACONS[8][7] = {
    {0, 0, 0, 0, 0, 0, 0},
    {2.0 / 3.0, 0, 0, 0, 0, 0, 0},
    {1.0 / 50.0, 5.0 / 294.0, 0, 0, 0, 0, 0},
    {1.0 / 588.0, 7.0 / 1440.0, 21.0 / 3872.0, 0, 0, 0, 0},
    {1.0 / 4320.0, 3.0 / 1936.0, 7601.0 / 2271360.0, 143.0 / 28800.0, 0, 0, 0},
    {1.0 / 23232.0, 7601.0 / 13628160.0, 143.0 / 69120.0, 517231.0 / 106536960.0, 106640677.0 / 11737571328.0, 0, 0},
    {691.0 / 68140800.0, 13.0 / 57600.0, 47021.0 / 35512320.0, 9694607.0 / 2095994880.0, 733191589.0 / 59609088000.0,
     326190917.0 / 11700633600.0, 0},
    {1.0 / 345600.0, 3617.0 / 35512320.0, 745739.0 / 838397952.0, 56399353.0 / 12773376000.0, 25091609.0 / 1560084480.0,
     1755948832039.0 / 36229939200000.0, 4887769399.0 / 37838389248.0}};  // pppm.cpp:129-161

static bool pppm_factorable(int n)  // pppm.cpp:1140-1156
{
  while (n > 1) {
    if (n % 2 == 0) n /= 2;
    else if (n % 3 == 0) n /= 3;
    else if (n % 5 == 0) n /= 5;
    else return false;
  }
  return true;
}

}  // namespace polb200

using namespace polb200;

struct polb200_pppm {
  std::string err;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {};
  long launches = 0;
  bool ready = false, have_plan = false;
  PppmConst C{};
  double accuracy = 0.0, qqrd2e = 0.0, qsum = 0.0, qsqsum = 0.0, q2 = 0.0, cutoff = 0.0, volume = 0.0;
  long natoms = 0;
  cufftHandle plan = 0;
  DBuf<double> greensfn, partial, out, c_x, c_q, c_f;
  DBuf<double2> work1, wx, wy, wz;
  DBuf<int> key, key2, idx, idx2, cstart;   // atomics-free charge assignment: atoms sorted by grid cell
  DBuf<double> wts;
  DBuf<char> cub_tmp;
  ncclComm_t nccl = nullptr;
  int rank = 0, nranks = 1;
  bool rho_atomics = false;                 // true: the first version (order^3 FP64 atomicAdd per atom; sums not reproducible)
  HPinned<double> h_out, h_f;
  float ms_last = 0.f;
};

namespace polb200 {

template <class F>
static int pppm_guarded(polb200_pppm *p, F &&fn)
{
  try {
    fn();
    return POLB200_OK;
  } catch (const StyleError &x) {
    p->err = x.msg;
    return x.code;
  } catch (const CudaError &x) {
    p->err = x.msg;
    return POLB200_ERR_CUDA;
  } catch (const std::exception &x) {
    p->err = x.what();
    return POLB200_ERR_ARG;
  }
}

static double pppm_ik_error(const polb200_pppm *p, double h, double prd)  // estimate_ik_error, pppm.cpp:1270-1281
{
  if (p->natoms == 0) return 0.0;
  const double g = p->C.g_ewald;
  double sum = 0.0;
  for (int m = 0; m < p->C.order; m++) sum += PPPM_ACONS[p->C.order][m] * pow(h * g, 2.0 * m);
  return p->q2 * pow(h * g, (double)p->C.order) * sqrt(g * prd * sqrt(2.0 * M_PI) * sum / p->natoms) / (prd * prd);
}

static double pppm_nr_f(const polb200_pppm *p)  // newton_raphson_f with compute_df_kspace (ik), pppm.cpp:1161-1181,1306-1320
{
  const double *prd = p->C.prd, g = p->C.g_ewald;
  const double df_r = 2.0 * p->q2 * exp(-g * g * p->cutoff * p->cutoff) / sqrt(p->natoms * p->cutoff * prd[0] * prd[1] * prd[2]);
  const int n[3] = {p->C.nx, p->C.ny, p->C.nz};
  double l2 = 0.0;
  for (int d = 0; d < 3; d++) {
    const double e = pppm_ik_error(p, prd[d] / n[d], prd[d]);
    l2 += e * e;
  }
  return df_r - sqrt(l2) / sqrt(3.0);
}

#define PPPM_LAUNCHED(p)           \
  do {                             \
    CUDA_CHECK(cudaGetLastError()); \
    (p)->launches++;               \
  } while (0)

#define CUFFT_CHECK(expr)                                                                                    \
  do {                                                                                                       \
    cufftResult _r = (expr);                                                                                 \
    if (_r != CUFFT_SUCCESS) throw CudaError{std::string(#expr) + ": cuFFT error " + std::to_string((int)_r)}; \
  } while (0)

}  // namespace polb200

extern "C" {

int polb200_pppm_create(polb200_pppm_t **out, int device)
{
  if (!out) return POLB200_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    fprintf(stderr, "polb200_pppm_create: no usable CUDA device %d (found %d); there is no CPU fallback\n", device, count);
    return POLB200_ERR_CUDA;
  }
  polb200_pppm *p = new polb200_pppm();
  p->device = device;
  if (const char *v = getenv("POLB200_PPPM_ATOMICS")) p->rho_atomics = atoi(v) != 0;  // A/B against the first version
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&p->ev[0]) != cudaSuccess || cudaEventCreate(&p->ev[1]) != cudaSuccess) {
    delete p;
    return POLB200_ERR_CUDA;
  }
  *out = p;
  return POLB200_OK;
}

void polb200_pppm_destroy(polb200_pppm_t *p)
{
  if (!p) return;
  cudaSetDevice(p->device);
  cudaStreamSynchronize(p->stream);
  if (p->have_plan && g_cufft.Destroy) g_cufft.Destroy(p->plan);
  if (p->nccl) g_pppm_nccl.CommDestroy(p->nccl);
  p->greensfn.release(); p->partial.release(); p->out.release(); p->c_x.release(); p->c_q.release(); p->c_f.release();
  p->work1.release(); p->wx.release(); p->wy.release(); p->wz.release(); p->h_out.release(); p->h_f.release();
  cudaEventDestroy(p->ev[0]); cudaEventDestroy(p->ev[1]);
  cudaStreamDestroy(p->stream);
  delete p;
}

const char *polb200_pppm_last_error(const polb200_pppm_t *p) { return p ? p->err.c_str() : "null handle"; }

int polb200_pppm_init(polb200_pppm_t *p, const polb200_pppm_setup *in, polb200_pppm_info *info)
{
  if (!p || !in) return POLB200_ERR_ARG;
  return pppm_guarded(p, [&] {
    CUDA_CHECK(cudaSetDevice(p->device));
    for (int d = 0; d < 3; d++)
      if (!in->periodic[d]) throw StyleError{POLB200_ERR_UNSUPPORTED, "Cannot use nonperiodic boundaries with PPPM"};
    PppmConst &C = p->C;
    C = PppmConst{};
    C.order = in->order > 0 ? in->order : 5;
    if (C.order < 2 || C.order > PPPM_MAXORDER) throw StyleError{POLB200_ERR_ARG, "PPPM order cannot be < 2 or > than 7"};
    p->qqrd2e = in->qqrd2e;
    p->qsum = in->qsum;
    p->qsqsum = in->qsqsum;
    p->q2 = in->qsqsum * in->qqrd2e;
    p->natoms = in->natoms;
    p->cutoff = in->cutoff;
    p->accuracy = in->accuracy_relative * in->two_charge_force;
    for (int d = 0; d < 3; d++) {
      C.boxlo[d] = in->boxlo[d];
      C.prd[d] = in->boxhi[d] - in->boxlo[d];
      C.unitk[d] = 2.0 * M_PI / C.prd[d];
    }
    p->volume = C.prd[0] * C.prd[1] * C.prd[2];
    // set_grid_global (pppm.cpp:985-1135)
    const bool gewaldflag = in->g_ewald > 0.0;
    if (!gewaldflag) {
      if (p->accuracy <= 0.0) throw StyleError{POLB200_ERR_ARG, "KSpace accuracy must be > 0"};
      if (p->q2 == 0.0) throw StyleError{POLB200_ERR_ARG, "Must use kspace_modify gewald for uncharged system"};
      double g = p->accuracy * sqrt(p->natoms * p->cutoff * C.prd[0] * C.prd[1] * C.prd[2]) / (2.0 * p->q2);
      if (g >= 1.0) g = (1.35 - 0.15 * log(p->accuracy)) / p->cutoff;
      else g = sqrt(-log(g)) / p->cutoff;
      C.g_ewald = g;
    } else C.g_ewald = in->g_ewald;
    int n[3];
    if (in->mesh[0] > 0 && in->mesh[1] > 0 && in->mesh[2] > 0) {
      for (int d = 0; d < 3; d++) n[d] = in->mesh[d];
    } else {
      for (int d = 0; d < 3; d++) {
        double h = 1.0 / C.g_ewald;
        n[d] = static_cast<int>(C.prd[d] / h) + 1;
        double err = pppm_ik_error(p, h, C.prd[d]);
        while (err > p->accuracy) {
          err = pppm_ik_error(p, h, C.prd[d]);
          n[d]++;
          h = C.prd[d] / n[d];
        }
      }
    }
    for (int d = 0; d < 3; d++)
      while (!pppm_factorable(n[d])) n[d]++;
    if (n[0] >= PPPM_OFFSET || n[1] >= PPPM_OFFSET || n[2] >= PPPM_OFFSET) throw StyleError{POLB200_ERR_ARG, "PPPM grid is too large"};
    C.nx = n[0]; C.ny = n[1]; C.nz = n[2];
    // set_grid_local (:1379-1385)
    C.nlower = -(C.order - 1) / 2;
    C.nupper = C.order / 2;
    C.shift = (C.order % 2) ? PPPM_OFFSET + 0.5 : (double)PPPM_OFFSET;
    C.shiftone = (C.order % 2) ? 0.0 : 0.5;
    // adjust_gewald (:1287-1340)
    if (!gewaldflag) {
      bool converged = false;
      for (int it = 0; it < 10000 && !converged; it++) {
        const double f1 = pppm_nr_f(p), g_old = C.g_ewald;
        C.g_ewald = g_old + 0.000001;
        const double f2 = pppm_nr_f(p);
        C.g_ewald = g_old;
        const double df = (f2 - f1) / 0.000001;
        C.g_ewald -= pppm_nr_f(p) / df;
        converged = fabs(pppm_nr_f(p)) < 0.00001;
      }
      if (!converged) throw StyleError{POLB200_ERR_ARG, "Could not compute g_ewald"};
    }
    // setup (:400-495)
    for (int d = 0; d < 3; d++) C.delinv[d] = n[d] / C.prd[d];
    C.delvolinv = C.delinv[0] * C.delinv[1] * C.delinv[2];
    // compute_gf_denom (:1526-1544)
    {
      double *b = C.gf_b;
      for (int l = 1; l < C.order; l++) b[l] = 0.0;
      b[0] = 1.0;
      for (int m = 1; m < C.order; m++) {
        int l;
        for (l = m; l > 0; l--) b[l] = 4.0 * (b[l] * (l - m) * (l - m - 0.5) - b[l - 1] * (l - m - 1) * (l - m - 1));
        b[0] = 4.0 * (b[0] * (l - m) * (l - m - 0.5));
      }
      long long ifact = 1;
      for (int k = 1; k < 2 * C.order; k++) ifact *= k;
      const double gaminv = 1.0 / (double)ifact;
      for (int l = 0; l < C.order; l++) b[l] *= gaminv;
    }
    // compute_rho_coeff (:2908-2952)
    {
      const int o = C.order;
      std::vector<double> a((size_t)o * (2 * o + 1), 0.0);
      auto A = [&](int l, int k) -> double & { return a[(size_t)l * (2 * o + 1) + (k + o)]; };
      A(0, 0) = 1.0;
      for (int j = 1; j < o; j++)
        for (int k = -j; k <= j; k += 2) {
          double s = 0.0;
          for (int l = 0; l < j; l++) {
            A(l + 1, k) = (A(l, k + 1) - A(l, k - 1)) / (l + 1);
            s += pow(0.5, (double)l + 1) * (A(l, k - 1) + pow(-1.0, (double)l) * A(l, k + 1)) / (l + 1);
          }
          A(0, k) = s;
        }
      int m = 0;
      for (int k = -(o - 1); k < o; k += 2) {
        for (int l = 0; l < o; l++) C.rho_coeff[l][m] = A(l, k);
        m++;
      }
    }
    for (int d = 0; d < 3; d++)
      C.nb[d] = static_cast<int>((C.g_ewald * C.prd[d] / (M_PI * n[d])) * pow(-log(1.0e-7), 0.25));  // EPS_HOC
    // device side: grids, FFT plan, influence function
    std::string e;
    if (!g_cufft.load(e)) throw StyleError{POLB200_ERR_UNSUPPORTED, e};
    const size_t total = (size_t)n[0] * n[1] * n[2];
    p->greensfn.ensure(total);
    p->work1.ensure(total); p->wx.ensure(total); p->wy.ensure(total); p->wz.ensure(total);
    p->partial.ensure((size_t)NPPPM_PART * cdiv((long)total, 256));
    p->out.ensure(8);
    p->h_out.ensure(8);
    if (p->have_plan) {
      g_cufft.Destroy(p->plan);
      p->have_plan = false;
    }
    CUFFT_CHECK(g_cufft.Plan3d(&p->plan, n[2], n[1], n[0], CUFFT_Z2Z));  // slowest dimension first: index (m*ny + l)*nx + k
    p->have_plan = true;
    CUFFT_CHECK(g_cufft.SetStream(p->plan, p->stream));
    k_pppm_gf<<<cdiv((long)total, 128), 128, 0, p->stream>>>(C, p->greensfn.p);
    PPPM_LAUNCHED(p);
    CUDA_CHECK(cudaStreamSynchronize(p->stream));
    p->ready = true;
    if (info) {
      info->g_ewald = C.g_ewald;
      info->nx = n[0]; info->ny = n[1]; info->nz = n[2];
      info->order = C.order;
    }
  });
}

int polb200_pppm_compute(polb200_pppm_t *p, int nlocal, const double *x, const double *q, double *f, int eflag, int vflag,
                         int on_device, double *energy, double virial[6])
{
  if (!p || nlocal < 0 || (nlocal > 0 && (!x || !q || !f))) return POLB200_ERR_ARG;
  return pppm_guarded(p, [&] {
    if (!p->ready) throw StyleError{POLB200_ERR_STATE, "polb200_pppm_init has not been called"};
    if ((eflag / 2) || (vflag / 4)) throw StyleError{POLB200_ERR_UNSUPPORTED, "per-atom KSpace tallies are not implemented on the device"};
    CUDA_CHECK(cudaSetDevice(p->device));
    if (energy) *energy = 0.0;
    if (virial) for (int k = 0; k < 6; k++) virial[k] = 0.0;
    const int n = nlocal;
    if (p->qsqsum == 0.0 || (n == 0 && !p->nccl)) return;  // pppm.cpp:650 (a rank without atoms still joins the grid sum)
    const PppmConst &C = p->C;
    const size_t total = (size_t)C.nx * C.ny * C.nz;
    CUDA_CHECK(cudaEventRecord(p->ev[0], p->stream));
    const double *dx = x, *dq = q;
    double *df = f;
    if (!on_device) {
      p->c_x.ensure((size_t)3 * n); p->c_q.ensure(n); p->c_f.ensure((size_t)3 * n);
      CUDA_CHECK(cudaMemcpyAsync(p->c_x.p, x, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
      CUDA_CHECK(cudaMemcpyAsync(p->c_q.p, q, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
      CUDA_CHECK(cudaMemsetAsync(p->c_f.p, 0, (size_t)3 * n * sizeof(double), p->stream));
      dx = p->c_x.p; dq = p->c_q.p; df = p->c_f.p;
    }
    const int ev = ((eflag & 1) || (vflag % 4)) ? 1 : 0;
    if (n == 0) {
      CUDA_CHECK(cudaMemsetAsync(p->work1.p, 0, total * sizeof(double2), p->stream));
    } else if (p->rho_atomics) {
      CUDA_CHECK(cudaMemsetAsync(p->work1.p, 0, total * sizeof(double2), p->stream));
      k_pppm_rho<<<cdiv(n, 128), 128, 0, p->stream>>>(n, C, dx, dq, p->work1.p);
      PPPM_LAUNCHED(p);
    } else {
      p->key.ensure(n); p->key2.ensure(n); p->idx.ensure(n); p->idx2.ensure(n); p->cstart.ensure(total + 2);
      p->wts.ensure((size_t)n * PPPM_WSTRIDE);
      k_pppm_cellkey<<<cdiv(n, 256), 256, 0, p->stream>>>(n, C, dx, p->key.p, p->idx.p);
      PPPM_LAUNCHED(p);
      int bits = 1;
      while ((1ul << bits) < total + 1) bits++;
      size_t bytes = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, bytes, p->key.p, p->key2.p, p->idx.p, p->idx2.p, n, 0, bits, p->stream);
      p->cub_tmp.ensure(bytes);
      CUDA_CHECK(cub::DeviceRadixSort::SortPairs(p->cub_tmp.p, bytes, p->key.p, p->key2.p, p->idx.p, p->idx2.p, n, 0, bits, p->stream));
      p->launches += 3;
      k_pppm_cellstart<<<cdiv((long)total + 1, 256), 256, 0, p->stream>>>((int)total, n, p->key2.p, p->cstart.p);
      PPPM_LAUNCHED(p);
      k_pppm_weights_sorted<<<cdiv(n, 128), 128, 0, p->stream>>>(n, C, p->idx2.p, dx, dq, p->wts.p);
      PPPM_LAUNCHED(p);
      k_pppm_rho_gather<<<cdiv((long)total * 32, 256), 256, 0, p->stream>>>(C, p->cstart.p, p->wts.p, p->work1.p);
      PPPM_LAUNCHED(p);
    }
    // decomposed run: every rank assigned the charges of its own atoms; the grid is the sum (the reference folds its ghost
    // cells with GridComm::reverse_comm, pppm.cpp:660-661 -- here the bricks share ONE grid and the FFTs are replicated)
    if (p->nccl) {
      const ncclResult_t rc = g_pppm_nccl.AllReduce(p->work1.p, p->work1.p, 2 * total, ncclDouble, ncclSum, p->nccl, p->stream);
      if (rc != ncclSuccess) throw CudaError{std::string("ncclAllReduce of the PPPM charge grid: ") + g_pppm_nccl.GetErrorString(rc)};
    }
    // fft1->compute(work1,work1,1): flag 1 = the e^{+ikr} transform, unscaled (fft3d.cpp:103-123) = CUFFT_INVERSE
    CUFFT_CHECK(g_cufft.ExecZ2Z(p->plan, reinterpret_cast<cufftDoubleComplex *>(p->work1.p),
                                reinterpret_cast<cufftDoubleComplex *>(p->work1.p), CUFFT_INVERSE));
    const int nblocks = cdiv((long)total, 256);
    k_pppm_poisson<<<nblocks, 256, 0, p->stream>>>(C, p->greensfn.p, p->work1.p, p->wx.p, p->wy.p, p->wz.p, ev, p->partial.p);
    PPPM_LAUNCHED(p);
    if (ev) {
      k_pppm_sum<<<1, 32 * NPPPM_PART, 0, p->stream>>>(nblocks, p->partial.p, p->out.p);
      PPPM_LAUNCHED(p);
      CUDA_CHECK(cudaMemcpyAsync(p->h_out.p, p->out.p, NPPPM_PART * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    }
    // fft2->compute(work2,work2,-1): the e^{-ikr} transform = CUFFT_FORWARD
    for (DBuf<double2> *g : {&p->wx, &p->wy, &p->wz})
      CUFFT_CHECK(g_cufft.ExecZ2Z(p->plan, reinterpret_cast<cufftDoubleComplex *>(g->p), reinterpret_cast<cufftDoubleComplex *>(g->p),
                                  CUFFT_FORWARD));
    if (n > 0) {
      k_pppm_force<<<cdiv(n, 128), 128, 0, p->stream>>>(n, C, p->qqrd2e, dx, dq, p->wx.p, p->wy.p, p->wz.p, df);
      PPPM_LAUNCHED(p);
    }
    if (!on_device && n > 0) {
      p->h_f.ensure((size_t)3 * n);
      CUDA_CHECK(cudaMemcpyAsync(p->h_f.p, p->c_f.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    }
    CUDA_CHECK(cudaEventRecord(p->ev[1], p->stream));
    CUDA_CHECK(cudaStreamSynchronize(p->stream));
    cudaEventElapsedTime(&p->ms_last, p->ev[0], p->ev[1]);
    if (!on_device)
      for (size_t k = 0; k < (size_t)3 * n; k++) f[k] += p->h_f.p[k];
    if (ev) {
      const double qscale = p->qqrd2e;
      if ((eflag & 1) && energy) {  // pppm.cpp:700-708
        double en = p->h_out.p[0] * 0.5 * p->volume;
        en -= C.g_ewald * p->qsqsum / 1.77245385090551602729 +
              1.57079632679489661923 * p->qsum * p->qsum / (C.g_ewald * C.g_ewald * p->volume);
        *energy = en * qscale / p->nranks;   // every rank holds the global sums: per-rank partials that add up
      }
      if ((vflag % 4) && virial)
        for (int k = 0; k < 6; k++) virial[k] = 0.5 * qscale * p->volume * p->h_out.p[1 + k] / p->nranks;  // pppm.cpp:712-716
    }
  });
}

double polb200_pppm_last_ms(const polb200_pppm_t *p) { return p ? (double)p->ms_last : 0.0; }

int polb200_pppm_comm_init(polb200_pppm_t *p, int rank, int nranks, const void *id_bytes)
{
  if (!p || !id_bytes || nranks < 1 || rank < 0 || rank >= nranks) return POLB200_ERR_ARG;
  return pppm_guarded(p, [&] {
    if (p->nccl) throw StyleError{POLB200_ERR_STATE, "polb200_pppm_comm_init was already called"};
    std::string err;
    if (!g_pppm_nccl.load(err)) throw StyleError{POLB200_ERR_UNSUPPORTED, err};
    CUDA_CHECK(cudaSetDevice(p->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    const ncclResult_t rc = g_pppm_nccl.CommInitRank(&p->nccl, nranks, id, rank);
    if (rc != ncclSuccess) throw CudaError{std::string("ncclCommInitRank: ") + g_pppm_nccl.GetErrorString(rc)};
    p->rank = rank;
    p->nranks = nranks;
  });
}

}  // extern "C"
