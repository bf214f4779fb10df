// kernels.cuh -- hand-written sm_100a kernels of the five stages (see DESIGN.md §3 for the map).
//
// Common shape: one warp per owned atom ("row"), the 32 lanes stride over that atom's neighbours
// (list mode) or over all owned atoms (all-pairs minimum-image mode = the reference's semantics),
// per-lane FP64 accumulators, warp-shuffle reduction, fixed-order block/grid reductions so results
// are bit-reproducible run to run.  Tensor cores are not used: the work is a sparse pairwise
// gather with ~60-80 FP64 operations per pair, bounded by the FP64 pipe / L2 gathers.
// The dominant kernel (one Jacobi dipole iteration in list mode) departs from that shape where ncu said so:
// two cell-row neighbours per warp (pair groups) and a TMA bulk-copy / mbarrier shared-memory ring for the
// per-pair stream -- see "pair-group form" and "TMA-fed pair-group sweep" below and DESIGN.md §4.
// File map: stage 1 binning/ghosts/lists -> stage 2 LJ+Coulomb+field -> stage 3 sweeps (first version,
// matrix-free, cached, sequential GS, rank metric) -> stage 4 forces -> stage 5 reductions -> multi-GPU halo
// (send lists, pack/unpack, peer push, signal/wait) -> pair groups -> TMA sweep -> blocked exact-mode GS.
//
// HBM layout (cell-sorted SoA of 32-byte records, ghosts after owned atoms):
//   xq [next] double4 {x,y,z,q}            mua[next] double4 {mu_x,mu_y,mu_z,alpha}
//   tm [next] int2    {type,molecule}      tag[next] int
//   neigh: CSR rows (64-bit row offsets) of 32-bit entries  j | special<<30   (src/lmptype.h:58-59)
#pragma once
#include <cuda_runtime.h>

#include "comm_types.h"
#include "pair_math.cuh"

namespace polb200 {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int BLOCK = WARPS_PER_BLOCK * 32;
constexpr unsigned FULL = 0xffffffffu;
constexpr int SBBITS = 30;
constexpr int NEIGHMASK = 0x3FFFFFFF;

struct Grid {
  double lo[3];     // origin of the cell grid (box lo - ghost cutoff)
  double inv[3];    // 1/cell size
  int nc[3];        // cells per dimension
  int ncell;
  int xbits;        // sort key = cell << xbits | position along x inside the cell (atoms of a cell row end up x-sorted)
};

struct DevParams {
  PairConsts pc;
  LJCoeffs lj;
  CoulTablesDev tb;
  const double *cutneighsq;
  Box box;
  Grid grid;
};

// ---------------------------------------------------------------------------------------------------
// row enumerators: how a warp walks the partners of owned atom s
// ---------------------------------------------------------------------------------------------------
// A partner is described from the point of view of the reference's pair loop: `a` is the atom the
// pair expressions treat as "i" (lower index), del = x_a - image(x_b), and `self_is_a` tells whether
// the row atom is a.  List mode: the row atom is always a (del = x_i - x_j over ghosts).  All-pairs
// mode: orientation follows the caller's atom indices so that every pair is evaluated with exactly
// the operands the reference uses.
struct Partner {
  int j;
  double dx, dy, dz, rsq;
  bool self_is_a;
};

struct ListRows {
  const unsigned long long *rowstart;
  const int *neigh;
  const int *rowcount;  // non-null: row s holds rowcount[s] entries from rowstart[s] (per-step tight list)
  __device__ __forceinline__ unsigned long long begin(int s) const { return rowstart[s]; }
  __device__ __forceinline__ unsigned long long end(int s) const
  {
    return rowcount ? rowstart[s] + (unsigned long long)rowcount[s] : rowstart[s + 1];
  }
};
struct AllPairRows {
  int nloc;
  const int *perm;  // sorted -> caller index
};

// ---------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
  return v;
}

// Device-side stop flag of the precision-mode SCF loop (pol.cpp:1193-1237): the host enqueues iterations ahead of
// the convergence test; once the test kernel (k_scf_check) has raised the flag, every kernel of the iterations
// that were enqueued speculatively returns at once, so the dipoles are exactly those of the converged iteration.
__device__ __forceinline__ bool scf_stopped(const int *stop) { return stop != nullptr && *(const volatile int *)stop != 0; }

// what the fused sweep needs to store a new dipole into the ghost copies of its row atom: periodic images
// in this GPU's own arrays and/or ghost slots of neighbour bricks (peer memory over NVLink)
struct PushArgs {
  const unsigned long long *off;  // CSR over owned atoms (cell-sorted index) into ptr
  double4 *const *ptr;            // absolute address of every copy in the OUTPUT dipole array of this sweep
};

// The address is fetched at the START of the row (push_prefetch) so that the two dependent loads
// (offset -> address) are hidden behind the row's pair loop instead of sitting at the end of the warp's life.
// An atom has at most 26 copies, so one address per lane covers every case.
__device__ __forceinline__ double4 *push_prefetch(const PushArgs &Q, int s, int lane)
{
  const unsigned long long b = Q.off[s], e = Q.off[s + 1];
  return b + lane < e ? Q.ptr[b + lane] : nullptr;
}

__device__ __forceinline__ void push_row(double4 *dst, double nx, double ny, double nz, double a)
{
  nx = __shfl_sync(FULL, nx, 0);
  ny = __shfl_sync(FULL, ny, 0);
  nz = __shfl_sync(FULL, nz, 0);
  if (dst) *dst = make_double4(nx, ny, nz, a);
}

// 256-bit loads of the 32-byte records (sm_100a: LDG.E.ENL2.256 via the aligned double4 types)
__device__ __forceinline__ double4 ld4(const double4 *p)
{
#if __CUDACC_VER_MAJOR__ >= 13
  double4_32a v = *reinterpret_cast<const double4_32a *>(p);
  return make_double4(v.x, v.y, v.z, v.w);
#else
  double4 v;
  asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
               : "l"(p));
  return v;
#endif
}

// same record, but through L2 only (ld.global.cg): for arrays updated in place inside a kernel
__device__ __forceinline__ double4 ld4_cg(const double4 *p)
{
  const double2 *q = reinterpret_cast<const double2 *>(p);
  double2 a = __ldcg(q), b = __ldcg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ int cell_of(const Grid &g, double x, double y, double z, int &cx, int &cy, int &cz)
{
  cx = (int)floor((x - g.lo[0]) * g.inv[0]);
  cy = (int)floor((y - g.lo[1]) * g.inv[1]);
  cz = (int)floor((z - g.lo[2]) * g.inv[2]);
  cx = min(max(cx, 0), g.nc[0] - 1);
  cy = min(max(cy, 0), g.nc[1] - 1);
  cz = min(max(cz, 0), g.nc[2] - 1);
  return (cz * g.nc[1] + cy) * g.nc[0] + cx;
}

// sort key: cell index (major) and the quantised x position inside the cell (minor).  Atoms of one (y,z)
// cell row are therefore stored in ascending x across the whole row, and the partners of an atom inside
// that row -- an x interval -- are one contiguous index run: neighbour gathers of a warp touch consecutive
// 32-byte records (fewer L1 wavefronts per load instruction).
__device__ __forceinline__ int sort_key_of(const Grid &g, double x, double y, double z)
{
  int cx, cy, cz;
  const int cell = cell_of(g, x, y, z, cx, cy, cz);
  double fx = (x - g.lo[0]) * g.inv[0] - (double)cx;
  fx = fmin(fmax(fx, 0.0), 0.999999);
  return (cell << g.xbits) | (int)(fx * (double)(1 << g.xbits));
}

// block-level fixed-order reduction of NV per-warp values; lane 0 of each warp holds its value.
// Thread 0 returns with the block total in out[0..NV).
template <int NV, int NWARPS = WARPS_PER_BLOCK>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double *block_out)
{
  __shared__ double sm[NWARPS][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int k = 0; k < NV; k++) sm[warp][k] = v[k];
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int w = 0; w < NWARPS; w++) s += sm[w][threadIdx.x];
    block_out[(size_t)blockIdx.x * NV + threadIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------------------------------
// stage 1: binning, ghosts, neighbor list
// ---------------------------------------------------------------------------------------------------

// cell key of every owned atom (caller order) + NaN check (src/nbin.cpp:120-121)
__global__ void k_local_keys(int n, const double *__restrict__ x, Grid g, int *__restrict__ key,
                             int *__restrict__ iota, int *__restrict__ errflag)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = x[3 * i], b = x[3 * i + 1], c = x[3 * i + 2];
  if (!isfinite(a) || !isfinite(b) || !isfinite(c)) {
    atomicOr(errflag, 1);
    a = b = c = 0.0;
  }
  key[i] = sort_key_of(g, a, b, c);
  iota[i] = i;
}

// gather caller-order attributes into cell-sorted records (owned part of the ext arrays)
__global__ void k_gather_local(int n, const int *__restrict__ perm, const double *__restrict__ x,
                               const double *__restrict__ q, const int *__restrict__ type,
                               const int *__restrict__ mol, const int *__restrict__ tag,
                               const double *__restrict__ alpha, const double *__restrict__ mu,
                               double4 *__restrict__ xq, double4 *__restrict__ mua, int2 *__restrict__ tm,
                               int *__restrict__ tagout, int *__restrict__ invperm, int *__restrict__ molflag)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int c = perm[s];
  if (mol && mol[c] != 0 && *molflag == 0) atomicOr(molflag, 1);
  xq[s] = make_double4(x[3 * c], x[3 * c + 1], x[3 * c + 2], q[c]);
  mua[s] = make_double4(mu[3 * c], mu[3 * c + 1], mu[3 * c + 2], alpha[c]);
  tm[s] = make_int2(type[c], mol ? mol[c] : 0);
  tagout[s] = tag ? tag[c] : c + 1;
  invperm[c] = s;
}

// per-step refresh (no rebuild): positions and dipoles of owned atoms from caller order
// (charges and polarizabilities are re-read every step as well: a fix may change them between two rebuilds)
__global__ void k_refresh_local(int n, const int *__restrict__ perm, const double *__restrict__ x,
                                const double *__restrict__ mu, const double *__restrict__ q,
                                const double *__restrict__ alpha, double4 *__restrict__ xq,
                                double4 *__restrict__ mua)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int c = perm[s];
  xq[s] = make_double4(x[3 * c], x[3 * c + 1], x[3 * c + 2], q[c]);
  mua[s] = make_double4(mu[3 * c], mu[3 * c + 1], mu[3 * c + 2], alpha[c]);
}

// which of the 26 periodic images of an owned atom fall inside the ghost shell
// (same membership rule as CommBrick::borders, src/comm_brick.cpp:768-772: x <= lo+cut sends the
// +prd image, x >= hi-cut sends the -prd image, independently per dimension)
__device__ __forceinline__ int image_mask(const Box &b, double cut, double x, double y, double z, int *opt)
{
  // opt[d] bit0: shift +1 allowed, bit1: shift -1 allowed
  const double p[3] = {x, y, z};
  int n = 1;
  for (int d = 0; d < 3; d++) {
    int o = 0;
    if (b.periodic[d]) {
      if (p[d] <= b.lo[d] + cut) o |= 1;
      if (p[d] >= b.hi[d] - cut) o |= 2;
    }
    opt[d] = o;
    n *= 1 + (o & 1) + ((o >> 1) & 1);
  }
  return n - 1;  // number of ghost images
}

__global__ void k_ghost_count(int n, const double4 *__restrict__ xq, Box box, double cut,
                              unsigned long long *__restrict__ cnt)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int opt[3];
  double4 v = xq[s];
  cnt[s] = (unsigned long long)image_mask(box, cut, v.x, v.y, v.z, opt);
}

// emits ghosts in (owner, z-shift, y-shift, x-shift) order into unsorted staging arrays
__global__ void k_ghost_fill(int n, const double4 *__restrict__ xq, Box box, double cut, Grid g,
                             const unsigned long long *__restrict__ off, int *__restrict__ g_owner,
                             int *__restrict__ g_shift, int *__restrict__ g_key, int *__restrict__ g_iota)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int opt[3];
  double4 v = xq[s];
  if (image_mask(box, cut, v.x, v.y, v.z, opt) == 0) return;
  unsigned long long o = off[s];
  const int sh[3] = {0, 1, -1};
  for (int kz = 0; kz < 3; kz++) {
    if (kz && !((opt[2] >> (kz - 1)) & 1)) continue;
    for (int ky = 0; ky < 3; ky++) {
      if (ky && !((opt[1] >> (ky - 1)) & 1)) continue;
      for (int kx = 0; kx < 3; kx++) {
        if (kx && !((opt[0] >> (kx - 1)) & 1)) continue;
        if (!kx && !ky && !kz) continue;
        // ghost coordinate = owner + shift*prd, one rounding per shifted dimension
        // (AtomVecFull::pack_border, src/MOLECULE/atom_vec_full.cpp:403-421)
        double gx = sh[kx] ? v.x + sh[kx] * box.prd[0] : v.x;
        double gy = sh[ky] ? v.y + sh[ky] * box.prd[1] : v.y;
        double gz = sh[kz] ? v.z + sh[kz] * box.prd[2] : v.z;
        g_owner[o] = s;
        g_shift[o] = (sh[kx] + 1) | ((sh[ky] + 1) << 2) | ((sh[kz] + 1) << 4);
        g_key[o] = sort_key_of(g, gx, gy, gz);
        g_iota[o] = (int)o;
        o++;
      }
    }
  }
}

// ghost records in cell-sorted order behind the owned atoms
__global__ void k_ghost_gather(int nghost, int nloc, const int *__restrict__ order,
                               const int *__restrict__ owner_in, const int *__restrict__ shift_in, Box box,
                               double4 *__restrict__ xq, double4 *__restrict__ mua, int2 *__restrict__ tm,
                               int *__restrict__ tag, int *__restrict__ owner, int *__restrict__ shift)
{
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  int src = order[g];
  int s = owner_in[src], code = shift_in[src];
  int sx = (code & 3) - 1, sy = ((code >> 2) & 3) - 1, sz = ((code >> 4) & 3) - 1;
  double4 v = xq[s];
  xq[nloc + g] = make_double4(sx ? v.x + sx * box.prd[0] : v.x, sy ? v.y + sy * box.prd[1] : v.y,
                              sz ? v.z + sz * box.prd[2] : v.z, v.w);
  mua[nloc + g] = mua[s];
  tm[nloc + g] = tm[s];
  tag[nloc + g] = tag[s];
  owner[g] = s;
  shift[g] = code;
}

// ghost positions (and optionally dipoles) follow their owners: the device-side equivalent of
// CommBrick::forward_comm (src/comm_brick.cpp:463-524) for a single brick
template <bool POS, bool MU>
__global__ void k_ghost_refresh(const int *stop, int nghost, int nloc, const int *__restrict__ owner,
                                const int *__restrict__ shift, Box box, double4 *__restrict__ xq,
                                double4 *__restrict__ mua)
{
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost || scf_stopped(stop)) return;
  int s = owner[g];
  if (POS) {
    int code = shift[g];
    int sx = (code & 3) - 1, sy = ((code >> 2) & 3) - 1, sz = ((code >> 4) & 3) - 1;
    double4 v = xq[s];
    xq[nloc + g] = make_double4(sx ? v.x + sx * box.prd[0] : v.x, sy ? v.y + sy * box.prd[1] : v.y,
                                sz ? v.z + sz * box.prd[2] : v.z, v.w);
  }
  if (MU) mua[nloc + g] = mua[s];
}

// first index of every cell in a key-sorted array (keys ascending): lower_bound per cell
__global__ void k_cell_starts(int ncell, int n, const int *__restrict__ sorted_keys, int *__restrict__ start, int shift)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > ncell) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if ((sorted_keys[mid] >> shift) < c) lo = mid + 1;
    else hi = mid;
  }
  start[c] = lo;
}

// NPair::find_special (src/npair.h:111-137) with special_flag = {.,2,2,2} (a KSpace style is present,
// src/neighbor.cpp:380-382): returns 0 or the 1-2/1-3/1-4 level
__device__ __forceinline__ int find_special(const int *__restrict__ list, const int *__restrict__ ns, int tagj)
{
  const int n1 = ns[0], n2 = ns[1], n3 = ns[2];
  for (int k = 0; k < n3; k++)
    if (list[k] == tagj) return k < n1 ? 1 : (k < n2 ? 2 : 3);
  return 0;
}

// Full neighbor list of every owned atom over owned+ghost atoms; pair accepted iff
// rsq <= cutneighsq[itype][jtype] with the reference's FP64 expression (src/npair_half_bin_newton.cpp:96-101).
// FILL=false counts, FILL=true writes entries (ext index | special level << 30).
template <bool FILL>
__global__ void k_neigh_build(int nloc, DevParams P, const double4 *__restrict__ xq,
                              const int2 *__restrict__ tm, const int *__restrict__ tag,
                              const int *__restrict__ cl_start, const int *__restrict__ cg_start,
                              int nstencil, const int *__restrict__ stencil,  // rows: (dy+16) | (dz+16)<<6 | xext<<12
                              const int *__restrict__ nspecial, const int *__restrict__ special, int maxspecial,
                              unsigned long long *__restrict__ count,
                              const unsigned long long *__restrict__ rowstart, int *__restrict__ neigh)
{
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (s >= nloc) return;
  const double4 xi = xq[s];
  const int ti = tm[s].x;
  const int n1 = P.pc.ntypes + 1;
  int cx, cy, cz;
  cell_of(P.grid, xi.x, xi.y, xi.z, cx, cy, cz);
  unsigned long long n = 0;
  const unsigned long long base = FILL ? rowstart[s] : 0ull;
  const bool has_special = nspecial != nullptr && nspecial[3 * s + 2] > 0;

  // The atoms are sorted by fine cell with x fastest, so each stencil row (dy,dz) covers ONE contiguous
  // index range per partition (owned / ghost): lanes walk long runs of consecutive records and the
  // rows of the list come out ascending in index with long contiguous stretches (gather locality).
  for (int k = 0; k < nstencil; k++) {
    const int code = stencil[k];
    const int oy = cy + ((code & 63) - 16), oz = cz + (((code >> 6) & 63) - 16), xext = code >> 12;
    if (oy < 0 || oz < 0 || oy >= P.grid.nc[1] || oz >= P.grid.nc[2]) continue;
    const int xlo = max(cx - xext, 0), xhi = min(cx + xext, P.grid.nc[0] - 1);
    const int rowbase = (oz * P.grid.nc[1] + oy) * P.grid.nc[0];
    for (int part = 0; part < 2; part++) {
      const int beg = part ? nloc + cg_start[rowbase + xlo] : cl_start[rowbase + xlo];
      const int end = part ? nloc + cg_start[rowbase + xhi + 1] : cl_start[rowbase + xhi + 1];
      for (int j0 = beg; j0 < end; j0 += 32) {
        const int j = j0 + lane;
        bool ok = false;
        int entry = j;
        if (j < end && j != s) {
          const double4 xj = xq[j];
          const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
          const double rsq = rsq_nofma(dx, dy, dz);
          ok = rsq <= P.cutneighsq[ti * n1 + tm[j].x];
          if (FILL && ok && has_special) {
            int which = find_special(special + (size_t)s * maxspecial, nspecial + 3 * s, tag[j]);
            // Domain::minimum_image_check (src/domain.h:155-160): a far image of a bonded partner
            // is an ordinary neighbour
            if (which && !((P.box.periodic[0] && fabs(dx) > P.box.half[0]) ||
                           (P.box.periodic[1] && fabs(dy) > P.box.half[1]) ||
                           (P.box.periodic[2] && fabs(dz) > P.box.half[2])))
              entry = j | (which << SBBITS);
          }
        }
        const unsigned m = __ballot_sync(FULL, ok);
        if (FILL && ok) neigh[base + n + __popc(m & ((1u << lane) - 1))] = entry;
        n += __popc(m);
      }
    }
  }
  if (!FILL && lane == 0) count[s] = n;
}

// Per-step "tight" list: the entries of the skin list that are inside the interaction range at the
// CURRENT positions (rsq <= reach2, same no-FMA rsq the pair kernels test), compacted in place order
// into a second array with the same row offsets.  One pass per step; every per-pair kernel of the
// step (stage 2, the 30+ sweeps of stage 3, stage 4) then runs without cutoff divergence.
__global__ void __launch_bounds__(BLOCK)
k_tighten(int nloc, double reach2, const unsigned long long *__restrict__ rowstart, const int *__restrict__ neigh,
          const double4 *__restrict__ xq, int *__restrict__ tneigh, int *__restrict__ tcount)
{
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (s >= nloc) return;
  const double4 xi = xq[s];
  const unsigned long long beg = rowstart[s], end = rowstart[s + 1];
  int n = 0;
  for (unsigned long long k0 = beg; k0 < end; k0 += 32) {
    const unsigned long long k = k0 + lane;
    bool ok = false;
    int raw = 0;
    if (k < end) {
      raw = neigh[k];
      const double4 xj = ld4(xq + (raw & NEIGHMASK));
      ok = rsq_nofma(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z) <= reach2;
    }
    const unsigned m = __ballot_sync(FULL, ok);
    if (ok) tneigh[beg + n + __popc(m & ((1u << lane) - 1))] = raw;
    n += __popc(m);
  }
  if (lane == 0) tcount[s] = n;
}

// ---------------------------------------------------------------------------------------------------
// stage 2: LJ + real-space Ewald Coulomb (+ static field in list mode) over the full list
// ---------------------------------------------------------------------------------------------------
// per-block partial sums: evdwl, ecoul, vxx, vyy, vzz, vxy, vxz, vyz  (each pair seen twice => 1/2)
constexpr int NPAIR_PART = 8;

// `neigh_modify exclude` (NPair::exclusion, src/npair.cpp:173-203).  The reference drops excluded pairs while it builds
// the list its LJ/Coulomb loop walks; its polarization loops never look at that list.  Here the device list stays
// complete (rank metric, list-mode polarization need every pair) and the LJ/Coulomb part of k_pair skips the pair.
// Per-atom group membership is pre-digested into exb[atom] (owned atoms and ghosts): .x bit r = mask & rule[r].a,
// .y bit r = mask & rule[r].b (group rules).  Up to 32 rules.
constexpr int MAX_EXCL = 32;
enum { EXCL_TYPE = 0, EXCL_GROUP = 1, EXCL_MOL_INTRA = 2, EXCL_MOL_INTER = 3, EXCL_INCLUDE = 4 };
struct ExclRules {
  int n;
  int kind[MAX_EXCL], a[MAX_EXCL], b[MAX_EXCL];
};

// exb[atom] = {bits of the rules whose first mask the atom is in, bits of the group rules whose second mask it is in}
__device__ __forceinline__ bool excl_pair(const ExclRules &X, int2 tmi, int2 tmj, int2 ei, int2 ej)
{
  for (int r = 0; r < X.n; r++) {
    const int ai = (ei.x >> r) & 1, aj = (ej.x >> r) & 1;
    switch (X.kind[r]) {
      case EXCL_TYPE:
        if ((tmi.x == X.a[r] && tmj.x == X.b[r]) || (tmi.x == X.b[r] && tmj.x == X.a[r])) return true;
        break;
      case EXCL_GROUP: {
        const int bi = (ei.y >> r) & 1, bj = (ej.y >> r) & 1;
        if ((ai && bj) || (bi && aj)) return true;
        break;
      }
      case EXCL_MOL_INTRA:
        if (ai && aj && tmi.y == tmj.y) return true;
        break;
      case EXCL_MOL_INTER:
        if (ai && aj && tmi.y != tmj.y) return true;
        break;
      default:  // EXCL_INCLUDE (neigh_modify include g): the list holds pairs of two atoms of the group only
        if (!(ai && aj)) return true;
    }
  }
  return false;
}

__global__ void k_exbits(int nloc, const int *__restrict__ perm, const int *__restrict__ mask, ExclRules X, int2 *__restrict__ exb)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nloc) return;
  const int m = mask ? mask[perm[s]] : 0;
  unsigned ea = 0, eb = 0;
  for (int r = 0; r < X.n; r++) {
    if (X.kind[r] == EXCL_TYPE) continue;
    if (m & X.a[r]) ea |= 1u << r;
    if (X.kind[r] == EXCL_GROUP && (m & X.b[r])) eb |= 1u << r;
  }
  exb[s] = make_int2((int)ea, (int)eb);
}

// periodic images inherit the bits of their owners (single GPU); bricks receive them with the halo (comm.cuh)
__global__ void k_exbits_ghost(int ng, int nloc, const int *__restrict__ g_owner, int2 *__restrict__ exb)
{
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < ng) exb[nloc + g] = exb[g_owner[g]];
}

__global__ void k_pack_int2(int ns, const int *__restrict__ owner, const int2 *__restrict__ src, int2 *__restrict__ sbuf)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < ns) sbuf[t] = src[owner[t]];
}

__global__ void k_unpack_int2(int ng, const int *__restrict__ gslot, const int2 *__restrict__ rbuf, int2 *__restrict__ dst)
{
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < ng) dst[g] = rbuf[gslot[g]];
}

template <bool EVFLAG, bool FIELD, bool EXCL>
__global__ void __launch_bounds__(BLOCK, 3)
k_pair(int nloc, DevParams P, const double4 *__restrict__ xq, const int2 *__restrict__ tm, ListRows L,
       double4 *__restrict__ f_pair, double4 *__restrict__ ef, double *__restrict__ partial,
       double *__restrict__ eatom_row, double *__restrict__ vatom_row, ExclRules X, const int2 *__restrict__ exb,
       const int *__restrict__ g_owner)
{
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  double acc[NPAIR_PART] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (s < nloc) {
    const double4 xi = xq[s];
    const int2 tmi = tm[s];
    const int n1 = P.pc.ntypes + 1;
    const int2 exi = EXCL ? exb[s] : make_int2(0, 0);
    double fx = 0, fy = 0, fz = 0, ex = 0, ey = 0, ez = 0;
    // neighbour entries are fetched two trips ahead so that the dependent index -> gather chain of the next
    // trip overlaps this trip's arithmetic
    const int *__restrict__ row = L.neigh + L.begin(s);
    const int cnt = (int)(L.end(s) - L.begin(s));
    int rawA = lane < cnt ? __ldcs(row + lane) : 0;
    int rawB = lane + 32 < cnt ? __ldcs(row + lane + 32) : 0;
    for (int k = lane; k < cnt; k += 32) {
      const int rawC = k + 64 < cnt ? __ldcs(row + k + 64) : 0;
      const int raw = rawA;
      rawA = rawB;
      rawB = rawC;
      const int j = raw & NEIGHMASK, sb = (raw >> SBBITS) & 3;
      const double4 xj = ld4(xq + j);
      const int2 tmj = tm[j];
      const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
      const double rsq = rsq_nofma(dx, dy, dz);
      const int ij = tmi.x * n1 + tmj.x;
      bool skip = false;
      if (EXCL) skip = excl_pair(X, tmi, tmj, exi, exb[j]);
      if (!skip && rsq < P.lj.cutsq[ij]) {
        double evdwl, ecoul;
        const double fpair = lj_coul_pair(P.pc, P.lj, P.tb, ij, rsq, xi.w, xj.w, sb, EVFLAG, evdwl, ecoul);
        fx += dx * fpair;
        fy += dy * fpair;
        fz += dz * fpair;
        if (EVFLAG) {
          acc[0] += evdwl;
          acc[1] += ecoul;
          acc[2] += dx * dx * fpair;
          acc[3] += dy * dy * fpair;
          acc[4] += dz * dz * fpair;
          acc[5] += dx * dy * fpair;
          acc[6] += dx * dz * fpair;
          acc[7] += dy * dz * fpair;
        }
      }
      if (FIELD) {
        if (rsq <= P.pc.cut_coulsq && (tmi.y != tmj.y || tmi.y == 0)) {
          const double sc = static_field_scalar(P.pc, rsq) * xj.w;
          ex += sc * dx;
          ey += sc * dy;
          ez += sc * dz;
        }
      }
    }
    fx = warp_sum(fx);
    fy = warp_sum(fy);
    fz = warp_sum(fz);
    if (FIELD) {
      ex = warp_sum(ex);
      ey = warp_sum(ey);
      ez = warp_sum(ez);
    }
    if (lane == 0) {
      f_pair[s] = make_double4(fx, fy, fz, 0.0);
      if (FIELD) ef[s] = make_double4(ex * P.pc.kq, ey * P.pc.kq, ez * P.pc.kq, 0.0);  // pol.cpp:372-374
    }
  }
  if (EVFLAG) {
#pragma unroll
    for (int k = 0; k < NPAIR_PART; k++) acc[k] = 0.5 * warp_sum(acc[k]);
    // per-atom tallies (Pair::ev_tally, src/pair.cpp:888-947): the row atom's half of every pair it is in
    if (lane == 0 && s < nloc) {
      if (eatom_row) eatom_row[s] = acc[0] + acc[1];
      if (vatom_row)
        for (int k = 0; k < 6; k++) vatom_row[6 * (size_t)s + k] = acc[2 + k];
    }
    block_reduce_store<NPAIR_PART>(acc, partial);
  }
}

// ---------------------------------------------------------------------------------------------------
// stage 2 (all-pairs mode): static field over all owned pairs, minimum image (pol.cpp:329-361)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK)
k_static_allpairs(int nloc, DevParams P, const double4 *__restrict__ xq, const int2 *__restrict__ tm,
                  const int *__restrict__ perm, double4 *__restrict__ ef, int row0, int row_end)
{
  // rows [row0, row_end) of the nloc atoms (the whole system unless the all-pairs work is shared by several GPUs)
  const int lane = threadIdx.x & 31;
  const int s = row0 + blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (s >= row_end) return;
  const double4 xi = xq[s];
  const int moli = tm[s].y, ci = perm[s];
  double ex = 0, ey = 0, ez = 0;
  for (int j = lane; j < nloc; j += 32) {
    if (j == s) continue;
    const double4 xj = ld4(xq + j);
    const int molj = tm[j].y;
    const bool i_is_a = ci < perm[j];
    double dx, dy, dz;
    if (i_is_a) min_image_del(P.box, xi.x, xi.y, xi.z, xj.x, xj.y, xj.z, dx, dy, dz);
    else min_image_del(P.box, xj.x, xj.y, xj.z, xi.x, xi.y, xi.z, dx, dy, dz);
    const double rsq = rsq_nofma(dx, dy, dz);
    if (rsq <= P.pc.cut_coulsq && (moli != molj || moli == 0)) {
      double sc = static_field_scalar(P.pc, rsq) * xj.w;
      if (!i_is_a) sc = -sc;  // ef_static[j] -= ef_temp*q_i*del  (pol.cpp:355-357)
      ex += sc * dx;
      ey += sc * dy;
      ez += sc * dz;
    }
  }
  ex = warp_sum(ex);
  ey = warp_sum(ey);
  ez = warp_sum(ez);
  if (lane == 0) ef[s] = make_double4(ex * P.pc.kq, ey * P.pc.kq, ez * P.pc.kq, 0.0);
}

// first guess mu = gamma*alpha*E_static unless use_previous (pol.cpp:376-385), owned atoms
__global__ void k_init_mu(int nloc, double gamma, const double4 *__restrict__ ef, double4 *__restrict__ mua)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nloc) return;
  double4 m = mua[s];
  const double4 e = ef[s];
  double a = m.w * e.x, b = m.w * e.y, c = m.w * e.z;
  a *= gamma;
  b *= gamma;
  c *= gamma;
  mua[s] = make_double4(a, b, c, m.w);
}

// mu = alpha*E_static without gamma: the divergence reset of pol.cpp:1227-1235
__global__ void k_reset_mu(int nloc, const double4 *__restrict__ ef, double4 *__restrict__ mua)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nloc) return;
  double4 m = mua[s];
  const double4 e = ef[s];
  mua[s] = make_double4(m.w * e.x, m.w * e.y, m.w * e.z, m.w);
}

// ---------------------------------------------------------------------------------------------------
// stage 3: one sweep of the induced-dipole iteration (pol.cpp:1158-1180 + 1198-1205)
// ---------------------------------------------------------------------------------------------------
// rows = positions [pos_beg,pos_end) of `order` (identity when order == nullptr).  Reads dipoles from
// mu_in (owned + ghost), writes mu_out[s] for the rows only and the per-block sum of squared changes.
// Jacobi: mu_in != mu_out over all rows.  Ranked colouring sweep: rows = one chunk, mu_out = staging.
template <bool LIST>
__global__ void __launch_bounds__(BLOCK)
k_sweep(int pos_beg, int pos_end, const int *__restrict__ order, DevParams P, ListRows L, AllPairRows A,
        const double4 *__restrict__ xq, const double4 *__restrict__ mu_in, const double4 *__restrict__ ef,
        double4 *__restrict__ mu_out, double *__restrict__ partial, const int *stop = nullptr)
{
  if (scf_stopped(stop)) return;
  const int lane = threadIdx.x & 31;
  const int pos = pos_beg + blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  double chg[1] = {0.0};
  if (pos < pos_end) {
    const int s = order ? order[pos] : pos;
    const double4 xi = xq[s];
    const double4 mi = mu_in[s];
    double ex = 0, ey = 0, ez = 0;
    if (mi.w != 0.0) {  // alpha_i == 0 => mu_new = 0 whatever the field
      if (LIST) {
        const unsigned long long beg = L.begin(s), end = L.end(s);
        for (unsigned long long k = beg + lane; k < end; k += 32) {
          const int j = L.neigh[k] & NEIGHMASK;
          const double4 xj = ld4(xq + j);
          const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
          const double rsq = dx * dx + dy * dy + dz * dz;
          if (rsq < P.pc.polar_cutsq) {
            const double4 mj = ld4(mu_in + j);
            induced_field_pair(P.pc, dx, dy, dz, rsq, mj.x, mj.y, mj.z, ex, ey, ez);
          }
        }
      } else {
        const int ci = A.perm[s];
        for (int j = lane; j < A.nloc; j += 32) {
          if (j == s) continue;
          const double4 xj = ld4(xq + j);
          double dx, dy, dz;
          // T is even in del, so the orientation only selects which atom's coordinates anchor the
          // minimum image, exactly as the matrix build does for i<j (pol.cpp:1279-1282)
          if (ci < A.perm[j]) min_image_del(P.box, xi.x, xi.y, xi.z, xj.x, xj.y, xj.z, dx, dy, dz);
          else min_image_del(P.box, xj.x, xj.y, xj.z, xi.x, xi.y, xi.z, dx, dy, dz);
          const double rsq = dx * dx + dy * dy + dz * dz;
          const double4 mj = ld4(mu_in + j);
          induced_field_pair(P.pc, dx, dy, dz, rsq, mj.x, mj.y, mj.z, ex, ey, ez);
        }
      }
      ex = warp_sum(ex);
      ey = warp_sum(ey);
      ez = warp_sum(ez);
    }
    if (lane == 0) {
      const double4 e = ef[s];
      const double nx = mi.w * (e.x + ex), ny = mi.w * (e.y + ey), nz = mi.w * (e.z + ez);
      mu_out[s] = make_double4(nx, ny, nz, mi.w);
      chg[0] = (nx - mi.x) * (nx - mi.x) + (ny - mi.y) * (ny - mi.y) + (nz - mi.z) * (nz - mi.z);
    }
  }
  block_reduce_store<1>(chg, partial);
}

// Optimised list-mode sweep (same mathematics as k_sweep<true>):
//   * reciprocal square root instead of sqrt + two divisions, damping polynomials in Horner form;
//   * neighbour indices prefetched two iterations ahead, so the dependent index -> gather chain overlaps
//     the FP64 work (PF is kept as a template parameter of the launch table; only PF == 1 is instantiated);
//   * DAMP is a template parameter: no per-pair branch on the damping type.
// radial part of T: s1 = d1/r^3, s2 = -3 d2/r^5  (T mu = s1 mu + s2 (del.mu) del)
template <bool DAMP>
__device__ __forceinline__ void radial_scalars(const PairConsts &pc, double r2, double &s1, double &s2)
{
  double rinv = rsqrt(r2);
  if (r2 == 0.0) rinv = 0.0;  // coincident sites: the damped tensor vanishes (d1 = d2 = 0 at r = 0)
  const double r = r2 * rinv;
  const double rinv2 = rinv * rinv;
  double r3 = rinv2 * rinv;
  const double r5 = r3 * rinv2;
  if (DAMP) {
    const double ar = pc.polar_damp * r;
    const double e = exp(-ar);
    const double p2 = fma(ar, fma(0.5, ar, 1.0), 1.0);          // 1 + ar + (ar)^2/2
    const double p3 = fma(ar * ar * ar, 1.0 / 6.0, p2);         // ... + (ar)^3/6
    s1 = fma(-e, p2, 1.0) * r3;
    s2 = -3.0 * fma(-e, p3, 1.0) * r5;
  } else {
    if (r2 == 0.0) r3 = DBL_MAX;  // undamped reference value (pol.cpp:1285-1286)
    s1 = r3;
    s2 = -3.0 * r5;
  }
}

template <bool DAMP>
__device__ __forceinline__ void induced_pair_fast(const PairConsts &pc, double dx, double dy, double dz, double r2,
                                                  const double4 &mj, double &ex, double &ey, double &ez)
{
  double s1, s2;
  radial_scalars<DAMP>(pc, r2, s1, s2);
  const double t = s2 * (dx * mj.x + dy * mj.y + dz * mj.z);
  ex -= fma(t, dx, s1 * mj.x);
  ey -= fma(t, dy, s1 * mj.y);
  ez -= fma(t, dz, s1 * mj.z);
}

// index stream: read once per sweep, keep it out of the way of the gathered records (ld.global.cs)
__device__ __forceinline__ int ld_index(const int *p) { return __ldcs(p) & NEIGHMASK; }

// WPB warps per block, MINB resident blocks per SM requested from ptxas.  CHANGE: also emit the squared
// dipole change of each row (per-ROW partials, no block barrier; summed in fixed order afterwards).
template <bool DAMP, int PF, int WPB, int MINB, bool CHANGE, bool PUSH = false>
__global__ void __launch_bounds__(WPB * 32, MINB)
k_sweep_list2(int pos_beg, int pos_end, const int *__restrict__ order, DevParams P, ListRows L,
              const double4 *__restrict__ xq, const double4 *__restrict__ mu_in, const double4 *__restrict__ ef,
              double4 *__restrict__ mu_out, double *__restrict__ row_change, PushArgs Q = PushArgs{},
              const int *stop = nullptr)
{
  const int lane = threadIdx.x & 31;
  const int pos = pos_beg + blockIdx.x * WPB + (threadIdx.x >> 5);
  if (pos >= pos_end || scf_stopped(stop)) return;
  const int s = order ? order[pos] : pos;
  const double4 xi = xq[s];
  const double4 mi = mu_in[s];
  double4 *push_dst = nullptr;
  if (PUSH) push_dst = push_prefetch(Q, s, lane);
  double ex = 0, ey = 0, ez = 0;
  if (mi.w != 0.0) {
    const int *__restrict__ row = L.neigh + L.begin(s);
    const int cnt = (int)(L.end(s) - L.begin(s));
    const double cutsq = P.pc.polar_cutsq;
    int jA = lane < cnt ? ld_index(row + lane) : -1;
    int jB = lane + 32 < cnt ? ld_index(row + lane + 32) : -1;
    for (int k = lane; k < cnt; k += 32) {
      const int jC = (k + 64 < cnt) ? ld_index(row + k + 64) : -1;
      const double4 xj = ld4(xq + jA);
      const double4 mj = ld4(mu_in + jA);
      const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
      const double r2 = dx * dx + dy * dy + dz * dz;
      if (r2 < cutsq) induced_pair_fast<DAMP>(P.pc, dx, dy, dz, r2, mj, ex, ey, ez);
      jA = jB;
      jB = jC;
    }
    ex = warp_sum(ex);
    ey = warp_sum(ey);
    ez = warp_sum(ez);
  }
  double nx = 0, ny = 0, nz = 0;
  if (lane == 0) {
    const double4 e = ef[s];
    nx = mi.w * (e.x + ex), ny = mi.w * (e.y + ey), nz = mi.w * (e.z + ez);
    mu_out[s] = make_double4(nx, ny, nz, mi.w);
    if (CHANGE)
      row_change[pos - pos_beg] = (nx - mi.x) * (nx - mi.x) + (ny - mi.y) * (ny - mi.y) + (nz - mi.z) * (nz - mi.z);
  }
  if (PUSH) push_row(push_dst, nx, ny, nz, mi.w);
}

// Per-step cache of the radial scalars of every tight-list entry (16 B per pair).  The geometry is
// frozen during the SCF, so rsqrt/exp/damping are evaluated once per step instead of once per sweep;
// the sweeps then stream 20 B per pair (index + scalars) and do 13 FP64 operations per pair.
template <bool DAMP>
__global__ void __launch_bounds__(BLOCK)
k_radial_cache(int nloc, DevParams P, ListRows L, const double4 *__restrict__ xq, double2 *__restrict__ s12)
{
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (s >= nloc) return;
  const double4 xi = xq[s];
  const unsigned long long beg = L.begin(s), end = L.end(s);
  for (unsigned long long k = beg + lane; k < end; k += 32) {
    const double4 xj = ld4(xq + (L.neigh[k] & NEIGHMASK));
    const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
    const double r2 = dx * dx + dy * dy + dz * dz;
    double s1 = 0.0, s2 = 0.0;
    if (r2 < P.pc.polar_cutsq) radial_scalars<DAMP>(P.pc, r2, s1, s2);
    s12[k] = make_double2(s1, s2);
  }
}

template <int WPB, int MINB, bool CHANGE, bool PUSH = false>
__global__ void __launch_bounds__(WPB * 32, MINB)
k_sweep_cached(int pos_beg, int pos_end, const int *__restrict__ order, ListRows L,
               const double2 *__restrict__ s12, const double4 *__restrict__ xq, const double4 *__restrict__ mu_in,
               const double4 *__restrict__ ef, double4 *__restrict__ mu_out, double *__restrict__ row_change,
               PushArgs Q = PushArgs{}, const int *stop = nullptr)
{
  const int lane = threadIdx.x & 31;
  const int pos = pos_beg + blockIdx.x * WPB + (threadIdx.x >> 5);
  if (pos >= pos_end || scf_stopped(stop)) return;
  const int s = order ? order[pos] : pos;
  const double4 xi = xq[s];
  const double4 mi = mu_in[s];
  double4 *push_dst = nullptr;
  if (PUSH) push_dst = push_prefetch(Q, s, lane);
  double ex = 0, ey = 0, ez = 0;
  if (mi.w != 0.0) {
    const unsigned long long beg = L.begin(s);
    const int *__restrict__ row = L.neigh + beg;
    const double2 *__restrict__ rs = s12 + beg;
    const int cnt = (int)(L.end(s) - beg);
    int jA = lane < cnt ? ld_index(row + lane) : -1;
    int jB = lane + 32 < cnt ? ld_index(row + lane + 32) : -1;
    double2 cA = lane < cnt ? __ldcs(rs + lane) : make_double2(0, 0);
    double2 cB = lane + 32 < cnt ? __ldcs(rs + lane + 32) : make_double2(0, 0);
    for (int k = lane; k < cnt; k += 32) {
      const bool more = k + 64 < cnt;
      const int jC = more ? ld_index(row + k + 64) : -1;
      const double2 cC = more ? __ldcs(rs + k + 64) : make_double2(0, 0);
      const double4 xj = ld4(xq + jA);
      const double4 mj = ld4(mu_in + jA);
      const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
      const double t = cA.y * (dx * mj.x + dy * mj.y + dz * mj.z);
      ex -= fma(t, dx, cA.x * mj.x);
      ey -= fma(t, dy, cA.x * mj.y);
      ez -= fma(t, dz, cA.x * mj.z);
      jA = jB;
      jB = jC;
      cA = cB;
      cB = cC;
    }
    ex = warp_sum(ex);
    ey = warp_sum(ey);
    ez = warp_sum(ez);
  }
  double nx = 0, ny = 0, nz = 0;
  if (lane == 0) {
    const double4 e = ef[s];
    nx = mi.w * (e.x + ex), ny = mi.w * (e.y + ey), nz = mi.w * (e.z + ez);
    mu_out[s] = make_double4(nx, ny, nz, mi.w);
    if (CHANGE)
      row_change[pos - pos_beg] = (nx - mi.x) * (nx - mi.x) + (ny - mi.y) * (ny - mi.y) + (nz - mi.z) * (nz - mi.z);
  }
  if (PUSH) push_row(push_dst, nx, ny, nz, mi.w);
}

// interleaved colouring: chunk c of the sweep holds the ranked positions c, c+C, c+2C, ... (atoms that are neighbours
// in the ranked order -- typically the sites of one molecule -- land in different chunks and see each other's new
// dipoles within the sweep).  Realised as a re-ordering of the visiting order: out = concat_c in[c::C].
__global__ void k_interleave_order(int n, int C, const int *__restrict__ in, int *__restrict__ out)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int c = p % C, m = p / C;
  out[c * (n / C) + min(c, n % C) + m] = in[p];
}

// commit a chunk of the ranked colouring sweep AND refresh the ghost copies of the committed atoms in one pass
// (own periodic images / neighbour bricks' ghost slots through the push tables)
__global__ void k_commit_push(int pos_beg, int pos_end, const int *__restrict__ order, const double4 *__restrict__ staged,
                              double4 *__restrict__ mua, PushArgs Q, const int *stop = nullptr)
{
  int pos = pos_beg + blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= pos_end || scf_stopped(stop)) return;
  const int s = order ? order[pos] : pos;
  const double4 v = staged[s];
  mua[s] = v;
  const unsigned long long b = Q.off[s], e = Q.off[s + 1];
  for (unsigned long long u = b; u < e; u++) *Q.ptr[u] = v;
}

// commit a chunk of the ranked colouring sweep: staged values become visible (owned records)
__global__ void k_commit_rows(int pos_beg, int pos_end, const int *__restrict__ order,
                              const double4 *__restrict__ staged, double4 *__restrict__ mua, const int *stop = nullptr)
{
  int pos = pos_beg + blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= pos_end || scf_stopped(stop)) return;
  int s = order ? order[pos] : pos;
  mua[s] = staged[s];
}

// Strictly sequential Gauss-Seidel sweep in ranked order with immediate write-back = the reference's
// polar_gs / polar_gs_ranked iteration (pol.cpp:1158-1180), all-pairs minimum image.  One CTA walks
// the ranked atoms; its 1024 threads share the partner loop of the current atom.  Dipoles are read
// through L2 (ld.global.cg) because they are rewritten inside the kernel.
constexpr int GS_THREADS = 1024;
__global__ void __launch_bounds__(GS_THREADS)
k_gs_sequential(int nloc, const int *__restrict__ order, DevParams P, const int *__restrict__ perm,
                const double4 *__restrict__ xq, double4 *__restrict__ mua, const double4 *__restrict__ ef,
                double *__restrict__ change_out, const int *stop = nullptr)
{
  if (scf_stopped(stop)) return;
  __shared__ double sm[GS_THREADS / 32][3];
  __shared__ double tot[3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double change = 0.0;  // thread 0 only
  for (int pos = 0; pos < nloc; pos++) {
    const int s = order ? order[pos] : pos;
    const double4 mi = ld4_cg(mua + s);
    if (mi.w == 0.0) {
      if (threadIdx.x == 0) {
        change += mi.x * mi.x + mi.y * mi.y + mi.z * mi.z;
        mua[s] = make_double4(0.0, 0.0, 0.0, 0.0);
      }
      __syncthreads();
      continue;
    }
    const double4 xi = xq[s];
    const int ci = perm[s];
    double ex = 0, ey = 0, ez = 0;
    for (int j = threadIdx.x; j < nloc; j += GS_THREADS) {
      if (j == s) continue;
      const double4 xj = ld4(xq + j);
      double dx, dy, dz;
      if (ci < perm[j]) min_image_del(P.box, xi.x, xi.y, xi.z, xj.x, xj.y, xj.z, dx, dy, dz);
      else min_image_del(P.box, xj.x, xj.y, xj.z, xi.x, xi.y, xi.z, dx, dy, dz);
      const double rsq = dx * dx + dy * dy + dz * dz;
      const double4 mj = ld4_cg(mua + j);
      induced_field_pair(P.pc, dx, dy, dz, rsq, mj.x, mj.y, mj.z, ex, ey, ez);
    }
    ex = warp_sum(ex);
    ey = warp_sum(ey);
    ez = warp_sum(ez);
    if (lane == 0) {
      sm[warp][0] = ex;
      sm[warp][1] = ey;
      sm[warp][2] = ez;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      double t = 0.0;
      for (int w = 0; w < GS_THREADS / 32; w++) t += sm[w][threadIdx.x];
      tot[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const double4 e = ef[s];
      const double nx = mi.w * (e.x + tot[0]), ny = mi.w * (e.y + tot[1]), nz = mi.w * (e.z + tot[2]);
      change += (nx - mi.x) * (nx - mi.x) + (ny - mi.y) * (ny - mi.y) + (nz - mi.z) * (nz - mi.z);
      mua[s] = make_double4(nx, ny, nz, mi.w);
      __threadfence();
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) change_out[0] = change;
}

// ---------------------------------------------------------------------------------------------------
// rank metric for polar_gs_ranked (pol.cpp:196-226) over the neighbour list (owned + ghost images)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK)
k_rmin(int nloc, ListRows L, const double4 *__restrict__ xq, const double4 *__restrict__ mua,
       const int2 *__restrict__ tm, unsigned long long *__restrict__ rmin_bits)
{
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (s >= nloc) return;
  const double4 xi = xq[s];
  const double ai = mua[s].w;
  const int moli = tm[s].y;
  double best = 1000.0;  // pol.cpp:196
  if (ai > 0) {
    const unsigned long long beg = L.begin(s), end = L.end(s);
    for (unsigned long long k = beg + lane; k < end; k += 32) {
      const int j = L.neigh[k] & NEIGHMASK;
      const double4 xj = ld4(xq + j);
      const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
      const double r = sqrt(rsq_nofma(dx, dy, dz));
      if (mua[j].w > 0 && best > r && (moli != tm[j].y || moli == 0)) best = r;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_down_sync(FULL, best, o));
  if (lane == 0) atomicMin(rmin_bits, (unsigned long long)__double_as_longlong(best));
}

// Index the reference would give an owned or ghost atom: owned atoms keep the caller's index, ghosts
// follow in CommBrick::borders creation order (src/comm_brick.cpp:726-790): dimension by dimension,
// the +prd swap before the -prd swap, each swap scanning the atoms created so far in index order.
// That order is the lexicographic order of (z stage, y stage, x stage, owner index), stage = 0 for
// no shift, 1 for +prd, 2 for -prd.
// Decomposed runs (tag != nullptr) have no global caller index: atoms are keyed by their tag instead.
__device__ __forceinline__ unsigned long long reference_index_key(int j, int nloc, const int *__restrict__ perm,
                                                                  const int *__restrict__ g_owner,
                                                                  const int *__restrict__ g_shift,
                                                                  const int *__restrict__ tag)
{
  if (j < nloc) return (unsigned long long)(tag ? tag[j] : perm[j]);
  const int code = g_shift[j - nloc];
  const int sx = (code & 3) - 1, sy = ((code >> 2) & 3) - 1, sz = ((code >> 4) & 3) - 1;
  const int cx = sx == 0 ? 0 : (sx > 0 ? 1 : 2), cy = sy == 0 ? 0 : (sy > 0 ? 1 : 2), cz = sz == 0 ? 0 : (sz > 0 ? 1 : 2);
  return ((unsigned long long)((cz * 3 + cy) * 3 + cx + 1) << 32) |
         (unsigned long long)(tag ? tag[j] : perm[g_owner[j - nloc]]);
}

// rank_metric[i] = sum of alpha_i*alpha_j over partners closer than 1.5*rmin (pol.cpp:214-226).
// Ties between metrics decide the Gauss-Seidel order, and metrics are sums of a few products whose
// rounding depends on the summation order, so the terms are added in the reference's j order.
constexpr int RANK_CAP = 64;
__global__ void __launch_bounds__(BLOCK)
k_rank_metric(int nloc, ListRows L, const double4 *__restrict__ xq, const double4 *__restrict__ mua,
              const int2 *__restrict__ tm, const unsigned long long *__restrict__ rmin_bits,
              const int *__restrict__ perm, const int *__restrict__ g_owner, const int *__restrict__ g_shift,
              const int *__restrict__ tag, double *__restrict__ metric_caller)
{
  __shared__ unsigned long long s_key[WARPS_PER_BLOCK][RANK_CAP];
  __shared__ double s_term[WARPS_PER_BLOCK][RANK_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.x * WARPS_PER_BLOCK + warp;
  if (s >= nloc) return;
  const double rmin = __longlong_as_double((long long)*rmin_bits);
  const double4 xi = xq[s];
  const double ai = mua[s].w;
  const int moli = tm[s].y;
  int cnt = 0;
  double overflow = 0.0;
  const unsigned long long beg = L.begin(s), end = L.end(s);
  for (unsigned long long k0 = beg; k0 < end; k0 += 32) {
    const unsigned long long k = k0 + lane;
    bool ok = false;
    int j = 0;
    if (k < end) {
      j = L.neigh[k] & NEIGHMASK;
      const double4 xj = ld4(xq + j);
      const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
      const double r = sqrt(rsq_nofma(dx, dy, dz));
      ok = rmin * 1.5 > r && (moli != tm[j].y || moli == 0);
    }
    const unsigned m = __ballot_sync(FULL, ok);
    if (ok) {
      const int slot = cnt + __popc(m & ((1u << lane) - 1));
      const double term = ai * mua[j].w;
      if (slot < RANK_CAP) {
        s_key[warp][slot] = reference_index_key(j, nloc, perm, g_owner, g_shift, tag);
        s_term[warp][slot] = term;
      } else overflow += term;
    }
    cnt += __popc(m);
  }
  overflow = warp_sum(overflow);
  __syncwarp();
  if (lane == 0) {
    const int n = cnt < RANK_CAP ? cnt : RANK_CAP;
    for (int a = 1; a < n; a++) {  // insertion sort by reference index
      const unsigned long long kk = s_key[warp][a];
      const double tt = s_term[warp][a];
      int b = a - 1;
      while (b >= 0 && s_key[warp][b] > kk) {
        s_key[warp][b + 1] = s_key[warp][b];
        s_term[warp][b + 1] = s_term[warp][b];
        b--;
      }
      s_key[warp][b + 1] = kk;
      s_term[warp][b + 1] = tt;
    }
    double msum = 0.0;
    for (int a = 0; a < n; a++) msum += s_term[warp][a];
    metric_caller[perm[s]] = msum + overflow;
  }
}

// ---------------------------------------------------------------------------------------------------
// stage 4: charge-dipole and dipole-dipole forces + energies (pol.cpp:425-631)
// ---------------------------------------------------------------------------------------------------
// per-block partials: u_self, u_ef, u_dd, then the polarization virial (6)
constexpr int NPOL_PART = 9;

template <bool LIST, bool EVFLAG, bool VPAIR, bool VATOM = false>
__global__ void __launch_bounds__(BLOCK, 3)
k_polforce(int nloc, DevParams P, ListRows L, AllPairRows A, const double4 *__restrict__ xq,
           const double4 *__restrict__ mua, const int2 *__restrict__ tm, double4 *__restrict__ f_pol,
           double *__restrict__ partial, double *__restrict__ vatom_row, int row0)
{
  // rows [row0, nloc): row0 = 0 unless the all-pairs work is shared by several GPUs (then `partial` is offset alike)
  const int lane = threadIdx.x & 31;
  const int s = row0 + blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  double acc[NPOL_PART] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (s < nloc) {
    const double4 xi = xq[s];
    const double4 mi = mua[s];
    const int moli = tm[s].y;
    double fx = 0, fy = 0, fz = 0;
    double va[6] = {0, 0, 0, 0, 0, 0};  // per-atom virial of the row atom (VATOM)
    PolPairIn in;
    const bool molecules = P.pc.has_molecules != 0;
    auto visit = [&](int j, const double4 &xj, bool i_is_a, double dx, double dy, double dz) {
      const double4 mj = ld4(mua + j);
      in.dx = dx; in.dy = dy; in.dz = dz;
      in.intermolecular = true;
      if (molecules) in.intermolecular = (moli != tm[j].y) || moli == 0;
      if (i_is_a) {
        in.qa = xi.w; in.qb = xj.w; in.alpha_a = mi.w; in.alpha_b = mj.w;
        in.max_ = mi.x; in.may = mi.y; in.maz = mi.z; in.mbx = mj.x; in.mby = mj.y; in.mbz = mj.z;
      } else {
        in.qa = xj.w; in.qb = xi.w; in.alpha_a = mj.w; in.alpha_b = mi.w;
        in.max_ = mj.x; in.may = mj.y; in.maz = mj.z; in.mbx = mi.x; in.mby = mi.y; in.mbz = mi.z;
      }
      double px, py, pz, uef, udd;
      pol_force_pair(P.pc, in, EVFLAG, px, py, pz, uef, udd);
      if (!i_is_a) { px = -px; py = -py; pz = -pz; }
      fx += px; fy += py; fz += pz;
      if (VATOM) {  // ev_tally_xyz, per-atom part (src/pair.cpp:1041-1087): half of del (x) F to each atom
        const double sx = i_is_a ? dx : -dx, sy = i_is_a ? dy : -dy, sz = i_is_a ? dz : -dz;
        va[0] += 0.5 * sx * px; va[1] += 0.5 * sy * py; va[2] += 0.5 * sz * pz;
        va[3] += 0.5 * sx * py; va[4] += 0.5 * sx * pz; va[5] += 0.5 * sy * pz;
      }
      if (EVFLAG) {
        acc[1] += 0.5 * uef;  // every pair is visited from both of its atoms
        acc[2] += 0.5 * udd;
        if (VPAIR) {  // ev_tally_xyz (src/pair.cpp:1001-1089): v = del (x) F with del = x_a - x_b
          const double sx = i_is_a ? dx : -dx, sy = i_is_a ? dy : -dy, sz = i_is_a ? dz : -dz;
          acc[3] += 0.5 * sx * px; acc[4] += 0.5 * sy * py; acc[5] += 0.5 * sz * pz;
          acc[6] += 0.5 * sx * py; acc[7] += 0.5 * sx * pz; acc[8] += 0.5 * sy * pz;
        }
      }
    };
    if (LIST) {
      const double reach = fmax(P.pc.cut_coulsq, P.pc.polar_cutsq);
      const int *__restrict__ row = L.neigh + L.begin(s);
      const int cnt = (int)(L.end(s) - L.begin(s));
      int jA = lane < cnt ? (__ldcs(row + lane) & NEIGHMASK) : 0;
      int jB = lane + 32 < cnt ? (__ldcs(row + lane + 32) & NEIGHMASK) : 0;
      for (int k = lane; k < cnt; k += 32) {
        const int jC = k + 64 < cnt ? (__ldcs(row + k + 64) & NEIGHMASK) : 0;
        const int j = jA;
        jA = jB;
        jB = jC;
        const double4 xj = ld4(xq + j);
        const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
        if (dx * dx + dy * dy + dz * dz < reach) visit(j, xj, true, dx, dy, dz);
      }
    } else {
      const int ci = A.perm[s];
      for (int j = lane; j < A.nloc; j += 32) {
        if (j == s) continue;
        const double4 xj = ld4(xq + j);
        const bool i_is_a = ci < A.perm[j];
        double dx, dy, dz;
        if (i_is_a) min_image_del(P.box, xi.x, xi.y, xi.z, xj.x, xj.y, xj.z, dx, dy, dz);
        else min_image_del(P.box, xj.x, xj.y, xj.z, xi.x, xi.y, xi.z, dx, dy, dz);
        visit(j, xj, i_is_a, dx, dy, dz);
      }
    }
    fx = warp_sum(fx);
    fy = warp_sum(fy);
    fz = warp_sum(fz);
    if (VATOM) {
#pragma unroll
      for (int k = 0; k < 6; k++) va[k] = warp_sum(va[k]);
      if (lane == 0)
        for (int k = 0; k < 6; k++) vatom_row[6 * (size_t)s + k] = va[k];
    }
    if (lane == 0) {
      f_pol[s] = make_double4(fx, fy, fz, 0.0);
      if (EVFLAG) {
        if (mi.w != 0.0) acc[0] = 0.5 * (mi.x * mi.x + mi.y * mi.y + mi.z * mi.z) / mi.w;  // pol.cpp:432-433
        if (!VPAIR) {
          // F.r virial of the reference (src/pair.cpp:1495-1543) restricted to this force field:
          // polarization forces act on owned atoms at their stored coordinates
          acc[3] += fx * xi.x; acc[4] += fy * xi.y; acc[5] += fz * xi.z;
          acc[6] += fy * xi.x; acc[7] += fz * xi.x; acc[8] += fz * xi.y;
        }
      }
    }
  }
  if (EVFLAG) {
    // acc[0] and the F.r terms live in lane 0 only; the rest are per-lane partial sums
#pragma unroll
    for (int k = 0; k < NPOL_PART; k++) acc[k] = warp_sum(acc[k]);
    block_reduce_store<NPOL_PART>(acc, partial);
  }
}

// ---------------------------------------------------------------------------------------------------
// stage 5: fused reduction into f / eng_vdwl / eng_coul / eng_pol / virial
// ---------------------------------------------------------------------------------------------------
// total pair force per owned atom back in caller order, plus dipoles and static field for the host
__global__ void k_scatter_out(int nloc, const int *__restrict__ perm, const double4 *__restrict__ f_pair,
                              const double4 *__restrict__ f_pol, const double4 *__restrict__ mua,
                              const double4 *__restrict__ ef, double *__restrict__ f_out,
                              double *__restrict__ mu_out, double *__restrict__ ef_out)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nloc) return;
  const int c = perm[s];
  const double4 a = f_pair[s], b = f_pol[s], m = mua[s], e = ef[s];
  f_out[3 * c] = a.x + b.x;
  f_out[3 * c + 1] = a.y + b.y;
  f_out[3 * c + 2] = a.z + b.z;
  mu_out[3 * c] = m.x;
  mu_out[3 * c + 1] = m.y;
  mu_out[3 * c + 2] = m.z;
  ef_out[3 * c] = e.x;
  ef_out[3 * c + 1] = e.y;
  ef_out[3 * c + 2] = e.z;
}

// per-atom tallies back in caller order: eatom (pair part only: the reference passes zero energies to
// ev_tally_xyz) and vatom = pair part + polarization part
__global__ void k_scatter_atomev(int nloc, const int *__restrict__ perm, const double *__restrict__ eatom_row,
                                 const double *__restrict__ vpair_row, const double *__restrict__ vpol_row,
                                 double *__restrict__ eatom_out, double *__restrict__ vatom_out)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nloc) return;
  const int c = perm[s];
  if (eatom_out) eatom_out[c] = eatom_row[s];
  if (vatom_out)
    for (int k = 0; k < 6; k++) vatom_out[6 * (size_t)c + k] = vpair_row[6 * (size_t)s + k] + vpol_row[6 * (size_t)s + k];
}

// deterministic final reduction of nblocks x NV per-block partials: one CTA, fixed tree
template <int NV>
__global__ void k_reduce_partials(int nblocks, const double *__restrict__ partial, double *__restrict__ out,
                                  int accumulate, const int *stop = nullptr)
{
  if (scf_stopped(stop)) return;
  __shared__ double sm[256];
  for (int v = 0; v < NV; v++) {
    double t = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) t += partial[(size_t)b * NV + v];
    sm[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[v] = accumulate ? out[v] + sm[0] : sm[0];
    __syncthreads();
  }
}

// displacement check of Neighbor::check_distance (src/neighbor.cpp:1989-1995): any atom moved more
// than half the skin since the last rebuild
__global__ void k_check_distance(int nloc, const double *__restrict__ x, const double *__restrict__ xhold,
                                 double triggersq, int *__restrict__ flag)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nloc) return;
  const double dx = x[3 * i] - xhold[3 * i], dy = x[3 * i + 1] - xhold[3 * i + 1], dz = x[3 * i + 2] - xhold[3 * i + 2];
  if (rsq_nofma(dx, dy, dz) > triggersq) atomicOr(flag, 1);
}

// number of list entries inside the dipole cutoff (the P of the roofline model)
__global__ void __launch_bounds__(BLOCK)
k_count_polar_pairs(int nloc, double cutsq, ListRows L, const double4 *__restrict__ xq,
                    unsigned long long *__restrict__ total)
{
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (s >= nloc) return;
  const double4 xi = xq[s];
  unsigned long long c = 0;
  for (unsigned long long k = L.begin(s) + lane; k < L.end(s); k += 32) {
    const double4 xj = ld4(xq + (L.neigh[k] & NEIGHMASK));
    const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
    if (dx * dx + dy * dy + dz * dz < cutsq) c++;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(FULL, c, o);
  if (lane == 0) atomicAdd(total, c);
}

// rows of a [n][w] caller-order int table into sorted order
__global__ void k_gather_rows(int n, int w, const int *__restrict__ perm, const int *__restrict__ in,
                              int *__restrict__ out)
{
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)n * w) return;
  int s = (int)(t / w), k = (int)(t % w);
  out[t] = in[(size_t)perm[s] * w + k];
}

// f += a  (device-resident callers)
__global__ void k_add_inplace(long n, const double *__restrict__ a, double *__restrict__ f)
{
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) f[t] += a[t];
}

// ---------------------------------------------------------------------------------------------------
// multi-GPU halo: send lists, pack / unpack of the boundary shell
// ---------------------------------------------------------------------------------------------------
// directions (see decomp.h) in which an owned atom is a ghost of a neighbour brick: same membership
// rule per dimension as CommBrick::borders (src/comm_brick.cpp:768-772): x <= lo+cut goes towards -1,
// x >= hi-cut goes towards +1, and the rules of the three dimensions combine to edges and corners.
__device__ __forceinline__ unsigned send_mask(const SendGeom &G, double x, double y, double z)
{
  const double p[3] = {x, y, z};
  unsigned a[3];
  for (int k = 0; k < 3; k++) a[k] = 2u | (p[k] <= G.lo[k] + G.cut ? 1u : 0u) | (p[k] >= G.hi[k] - G.cut ? 4u : 0u);
  unsigned m = 0;
  for (int d = 0; d < NDIR; d++) {
    const int vx = d % 3, vy = (d / 3) % 3, vz = d / 9;
    if (((a[0] >> vx) & 1u) && ((a[1] >> vy) & 1u) && ((a[2] >> vz) & 1u)) m |= 1u << d;
  }
  return m & G.valid;
}

__global__ void k_send_count(int n, const double4 *__restrict__ xq, SendGeom G, unsigned long long *__restrict__ cnt)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const double4 v = xq[s];
  cnt[s] = (unsigned long long)__popc(send_mask(G, v.x, v.y, v.z));
}

// (owner, direction) of every send slot, in (owner, direction) order; a stable sort by direction then
// gives direction-major segments with owners ascending inside each
__global__ void k_send_fill(int n, const double4 *__restrict__ xq, SendGeom G, const unsigned long long *__restrict__ off,
                            int *__restrict__ owner, int *__restrict__ dir, int *__restrict__ iota)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const double4 v = xq[s];
  unsigned m = send_mask(G, v.x, v.y, v.z);
  unsigned long long o = off[s];
  while (m) {
    const int d = __ffs(m) - 1;
    m &= m - 1;
    owner[o] = s;
    dir[o] = d;
    iota[o] = (int)o;
    o++;
  }
}

__global__ void k_gather_int(int n, const int *__restrict__ order, const int *__restrict__ in, int *__restrict__ out)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = in[order[t]];
}

// positions of the send slots with the periodic image shift applied: one rounding per shifted
// dimension, exactly the single-GPU ghost expression (k_ghost_gather) and AtomVecFull::pack_border
__global__ void k_pack_pos(int ns, const int *__restrict__ owner, const int *__restrict__ dir, SendGeom G,
                           const double4 *__restrict__ xq, double4 *__restrict__ sbuf)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ns) return;
  const double4 v = xq[owner[t]];
  const int d = dir[t];
  sbuf[t] = make_double4(G.shift[d][0] != 0.0 ? v.x + G.shift[d][0] : v.x, G.shift[d][1] != 0.0 ? v.y + G.shift[d][1] : v.y,
                         G.shift[d][2] != 0.0 ? v.z + G.shift[d][2] : v.z, v.w);
}

__global__ void k_pack_rec(int ns, const int *__restrict__ owner, const double4 *__restrict__ src,
                           double4 *__restrict__ sbuf)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < ns) sbuf[t] = src[owner[t]];
}

__global__ void k_pack_meta(int ns, const int *__restrict__ owner, const int *__restrict__ dir, SendGeom G,
                            const int2 *__restrict__ tm, const int *__restrict__ tag, int4 *__restrict__ sbuf)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ns) return;
  const int s = owner[t];
  const int2 a = tm[s];
  sbuf[t] = make_int4(a.x, a.y, tag[s], G.code[dir[t]]);
}

// received records into the cell-sorted ghost section (dst points at ext index nloc)
__global__ void k_unpack_rec(int ng, const int *__restrict__ gslot, const double4 *__restrict__ rbuf,
                             double4 *__restrict__ dst)
{
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < ng) dst[g] = rbuf[gslot[g]];
}

__global__ void k_unpack_meta(int ng, const int *__restrict__ gslot, const int4 *__restrict__ rbuf,
                              int2 *__restrict__ tm, int *__restrict__ tag, int *__restrict__ code)
{
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  const int4 a = rbuf[gslot[g]];
  tm[g] = make_int2(a.x, a.y);
  tag[g] = a.z;
  code[g] = a.w;
}

__global__ void k_ghost_keys(int ng, const double4 *__restrict__ rbuf, Grid g, int *__restrict__ key,
                             int *__restrict__ iota)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ng) return;
  const double4 v = rbuf[t];
  key[t] = sort_key_of(g, v.x, v.y, v.z);
  iota[t] = t;
}

__global__ void k_invert_perm(int n, const int *__restrict__ order, int *__restrict__ inv)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) inv[order[t]] = t;
}

// ext index (on THIS rank) of every receive slot: the answer the senders need for the peer push
__global__ void k_ghost_ext_index(int ng, int nloc, const int *__restrict__ gslot, int *__restrict__ per_slot)
{
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < ng) per_slot[gslot[g]] = nloc + g;
}

struct DirTable {
  int v[NDIR];
};

// push tables in (owner, direction) order: for every send slot the address of the ghost record it feeds on
// its destination rank -- in both dipole arrays and in the position array (peer mappings; the rank's own
// arrays for self images) -- and its direction (for the periodic shift of pushed positions)
__global__ void k_push_tables(int ns, const int *__restrict__ slot_of_u, const int *__restrict__ dir_of_slot, DirTable dest,
                              const int *__restrict__ remote_of_slot, PeerPush P, double4 **__restrict__ ptr0,
                              double4 **__restrict__ ptr1, double4 **__restrict__ ptrx, int *__restrict__ dir_of_u)
{
  int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= ns) return;
  const int t = slot_of_u[u];
  const int d = dir_of_slot[t];
  const int r = dest.v[d];
  const int idx = remote_of_slot[t];
  ptr0[u] = P.mu[0][r] + idx;
  ptr1[u] = P.mu[1][r] + idx;
  ptrx[u] = P.xq[r] + idx;
  dir_of_u[u] = d;
}

// once-per-step halo through peer memory: every send slot stores its (shifted) position / its dipole record
// straight into the ghost slot of the brick that needs it
__global__ void k_push_pos(int ns, const int *__restrict__ owner_u, const int *__restrict__ dir_of_u, SendGeom G,
                           const double4 *__restrict__ xq, double4 *const *__restrict__ ptrx)
{
  int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= ns) return;
  const double4 v = xq[owner_u[u]];
  const int d = dir_of_u[u];
  *ptrx[u] = make_double4(G.shift[d][0] != 0.0 ? v.x + G.shift[d][0] : v.x, G.shift[d][1] != 0.0 ? v.y + G.shift[d][1] : v.y,
                          G.shift[d][2] != 0.0 ? v.z + G.shift[d][2] : v.z, v.w);
}

__global__ void k_push_rec(int ns, const int *__restrict__ owner_u, const double4 *__restrict__ src,
                           double4 *const *__restrict__ ptr, const int *stop = nullptr)
{
  int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < ns && !scf_stopped(stop)) *ptr[u] = src[owner_u[u]];
}

// single GPU: the copies are the periodic images behind the owned atoms of the same arrays
__global__ void k_push_tables_local(int ng, int nloc, const int *__restrict__ sorted_pos_of_u, double4 *mua,
                                    double4 *mub, double4 **__restrict__ ptr0, double4 **__restrict__ ptr1)
{
  int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= ng) return;
  ptr0[u] = mua + nloc + sorted_pos_of_u[u];
  ptr1[u] = mub + nloc + sorted_pos_of_u[u];
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Inter-GPU barrier + all-reduce of one double over peer memory, one warp.  flags layout per rank:
// [0,MAX_PEERS) arrival epoch of each source rank, then two banks of MAX_PEERS partial sums (by epoch parity).
// The dipoles pushed by the sweep kernel were stored by an EARLIER kernel of the same stream, so they are
// performed before this kernel's release-store of the epoch.
__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// A brick that never arrives (crashed peer, ranks sharing one GPU so that the kernels cannot run at the same
// time) must not hang the device: the wait gives up after timeout_ns and raises *timeout_flag, which the host
// turns into an error at the end of the step.
__global__ void k_signal_wait(int rank, int nranks, unsigned long long epoch, PeerPush P, double *__restrict__ change,
                              unsigned long long timeout_ns, int *__restrict__ timeout_flag, const int *stop = nullptr)
{
  // (the stop decision is taken from the all-reduced change, identical on every brick: all bricks skip together)
  if (scf_stopped(stop)) return;
  const int r = threadIdx.x;
  const int bank = (1 + (int)(epoch & 1ull)) * MAX_PEERS;
  unsigned long long *mine = P.flag[rank];
  const unsigned long long t_in = global_ns();
  if (r < nranks) {
    unsigned long long *theirs = P.flag[r];
    if (change) theirs[bank + rank] = (unsigned long long)__double_as_longlong(*change);
    __threadfence_system();
    st_release_sys(theirs + rank, epoch);
    const unsigned long long t0 = global_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(mine + r) < epoch) {
      if ((++spins & 1023u) == 0 && global_ns() - t0 > timeout_ns) {
        atomicExch(timeout_flag, 1);
        break;
      }
    }
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    if (change) {
      double s = 0.0;
      for (int k = 0; k < nranks; k++) s += __longlong_as_double((long long)ld_acquire_sys(mine + bank + k));
      *change = s;
    }
    // statistics (slots 3*MAX_PEERS.. of this rank's own counters): time inside barriers, number of barriers.  The brick
    // that waits least is the one the others wait for.
    mine[3 * MAX_PEERS] += global_ns() - t_in;
    mine[3 * MAX_PEERS + 1] += 1ull;
  }
}

// ---------------------------------------------------------------------------------------------------
// stage 3, pair-group form of the Jacobi list sweep
// ---------------------------------------------------------------------------------------------------
// ncu shows the per-atom sweep bound by the L1 data pipe (l1tex__data_pipe_lsu_wavefronts 85 % of peak):
// every pair costs two divergent 32-byte gathers (x_j, mu_j), ~17 wavefronts each per warp trip.  Two
// atoms that are neighbours in the cell-sorted order (same (y,z) cell row, adjacent in x) share ~85 % of
// their partners, so ONE warp serves both: each gathered record feeds two pair evaluations, halving the
// wavefronts per pair.  A group row is the union of the two atoms' partner sets; the per-step cache holds
// the radial scalars of both members for every entry (zeros where an entry belongs to one member only).

// groups of one (y,z) cell row: atoms [rs,re) of the row pair up as (rs,rs+1),(rs+2,rs+3),...
__global__ void k_group_count(int nrows, int ncx, const int *__restrict__ cl_start, unsigned long long *__restrict__ cnt)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int c = cl_start[(size_t)(r + 1) * ncx] - cl_start[(size_t)r * ncx];
  cnt[r] = (unsigned long long)((c + 1) / 2);
}

// first atom of every group; the second member is first+1 when group_two is set
__global__ void k_group_fill(int nrows, int ncx, const int *__restrict__ cl_start, const unsigned long long *__restrict__ off,
                             int *__restrict__ group_first, int *__restrict__ group_two)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int rs = cl_start[(size_t)r * ncx], re = cl_start[(size_t)(r + 1) * ncx];
  unsigned long long g = off[r];
  for (int a = rs; a < re; a += 2, g++) {
    group_first[g] = a;
    group_two[g] = a + 1 < re ? 1 : 0;
  }
}

// union skin list of every group straight from the cells (same acceptance test as k_neigh_build, for either
// member).  The two members sit in the same cell row, so they share the stencil rows; only the x range widens.
template <bool FILL>
__global__ void __launch_bounds__(BLOCK)
k_group_build(int ngroups, int nloc, DevParams P, const double4 *__restrict__ xq, const int2 *__restrict__ tm,
              const int *__restrict__ group_first, const int *__restrict__ group_two,
              const int *__restrict__ cl_start, const int *__restrict__ cg_start, int nstencil,
              const int *__restrict__ stencil, unsigned long long *__restrict__ count,
              const unsigned long long *__restrict__ rowstart, int *__restrict__ neigh, int *__restrict__ rowcount)
{
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (g >= ngroups) return;
  const int a = group_first[g];
  const bool two = group_two[g] != 0;
  const int b = two ? a + 1 : a;
  const double4 xa = xq[a], xb = xq[b];
  const int ta = tm[a].x, tb = tm[b].x;
  const int n1 = P.pc.ntypes + 1;
  int cxa, cxb, cy, cz, t0, t1;
  cell_of(P.grid, xa.x, xa.y, xa.z, cxa, cy, cz);
  cell_of(P.grid, xb.x, xb.y, xb.z, cxb, t0, t1);
  unsigned long long n = 0;
  const unsigned long long base = FILL ? rowstart[g] : 0ull;
  for (int k = 0; k < nstencil; k++) {
    const int code = stencil[k];
    const int oy = cy + ((code & 63) - 16), oz = cz + (((code >> 6) & 63) - 16), xext = code >> 12;
    if (oy < 0 || oz < 0 || oy >= P.grid.nc[1] || oz >= P.grid.nc[2]) continue;
    const int xlo = max(min(cxa, cxb) - xext, 0), xhi = min(max(cxa, cxb) + xext, P.grid.nc[0] - 1);
    const int rowbase = (oz * P.grid.nc[1] + oy) * P.grid.nc[0];
    for (int part = 0; part < 2; part++) {
      const int beg = part ? nloc + cg_start[rowbase + xlo] : cl_start[rowbase + xlo];
      const int end = part ? nloc + cg_start[rowbase + xhi + 1] : cl_start[rowbase + xhi + 1];
      for (int j0 = beg; j0 < end; j0 += 32) {
        const int j = j0 + lane;
        bool ok = false;
        if (j < end) {
          const double4 xj = xq[j];
          const int tj = tm[j].x;
          const double ra = rsq_nofma(xa.x - xj.x, xa.y - xj.y, xa.z - xj.z);
          const double rb = rsq_nofma(xb.x - xj.x, xb.y - xj.y, xb.z - xj.z);
          ok = (j != a && ra <= P.cutneighsq[ta * n1 + tj]) || (two && j != b && rb <= P.cutneighsq[tb * n1 + tj]);
        }
        const unsigned m = __ballot_sync(FULL, ok);
        if (FILL && ok) neigh[base + n + __popc(m & ((1u << lane) - 1))] = j;
        n += __popc(m);
      }
    }
  }
  // rows start on 16-byte boundaries of the index array (bulk copies): the offsets are padded, the true
  // length of the row is kept in rowcount
  if (!FILL && lane == 0) count[g] = (n + 3ull) & ~3ull;
  if (FILL && lane == 0) rowcount[g] = (int)n;
}

// chunk-record layout of the tight group rows (TMA sweep): a row is a sequence of 64-entry records of
// GCHUNK*36 bytes = [64 x 16 B scalars of member a][64 x 16 B of member b][64 x 4 B indices], contiguous in memory,
// so that ONE bulk copy fetches a whole chunk (indices included) as a single 2304-byte request and the sweep's
// LDS.128 of either member have a lane stride of 16 B (conflict free; {s1a,s2a,s1b,s2b} per entry was 2-way)
constexpr int GCHUNK = 64;
constexpr int GCHUNK_BYTES = GCHUNK * 36;
__global__ void k_group_chunk_count(int ngroups, const int *__restrict__ rowcount, unsigned long long *__restrict__ cnt)
{
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < ngroups) cnt[g] = (unsigned long long)((rowcount[g] + GCHUNK - 1) / GCHUNK);
}

// per step: entries inside the dipole cutoff of either member at the current positions, compacted, with the
// radial scalars {s1a, s2a, s1b, s2b} of both members (zero for a member the entry does not belong to)
// RMIN: also the closest inter-molecular pair of polarizable atoms (the rmin of the rank metric, pol.cpp:196-212, over
// the partners inside the dipole cutoff), so that steps that only report rmin need no pass of their own over the list.
// The distance of a candidate is evaluated with the reference's un-fused expression, like k_rmin.
template <bool DAMP, bool CHUNKED, bool RMIN = false>
__global__ void __launch_bounds__(BLOCK)
k_group_cache(int ngroups, DevParams P, const int *__restrict__ group_first, const int *__restrict__ group_two,
              const unsigned long long *__restrict__ rowstart, const int *__restrict__ rowcount,
              const int *__restrict__ neigh, const double4 *__restrict__ xq, int *__restrict__ tneigh,
              int *__restrict__ tcount, double4 *__restrict__ s12ab, const unsigned long long *__restrict__ cstart,
              unsigned char *__restrict__ crec, const double4 *__restrict__ mua = nullptr, const int2 *__restrict__ tm = nullptr,
              unsigned long long *__restrict__ rmin_bits = nullptr)
{
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (g >= ngroups) return;
  const int a = group_first[g];
  const bool two = group_two[g] != 0;
  const int b = two ? a + 1 : a;
  const double4 xa = xq[a], xb = xq[b];
  const double cutsq = P.pc.polar_cutsq;
  const unsigned long long beg = rowstart[g], end = beg + (unsigned long long)rowcount[g];
  double best = 1000.0, guard = 1.0e6;  // pol.cpp:196; guard = best^2 with a margin for the fused distance
  double ala = 0.0, alb = 0.0;
  int mola = 0, molb = 0;
  if (RMIN) {
    ala = mua[a].w; alb = mua[b].w;
    mola = tm[a].y; molb = tm[b].y;
  }
  int n = 0;
  // CHUNKED: the radial scalars are not evaluated by the lane that finds a partner (63 % of the skin candidates are inside
  // the cutoff and a quarter of those for one member only, so that code ran with ~55 % of its lanes) but queued per warp
  // as (destination, r^2) items and evaluated 32 at a time with every lane busy
  __shared__ unsigned q_dst[CHUNKED ? WARPS_PER_BLOCK : 1][96];
  __shared__ double q_r2[CHUNKED ? WARPS_PER_BLOCK : 1][96];
  const int wq = CHUNKED ? (threadIdx.x >> 5) : 0;
  int qn = 0;
  auto drain = [&](int item) {   // item: queue slot of this lane
    const unsigned d = q_dst[wq][item];
    double s1, s2;
    radial_scalars<DAMP>(P.pc, q_r2[wq][item], s1, s2);
    const unsigned pos = d >> 1;
    unsigned char *rec = crec + (cstart[g] + (unsigned long long)(pos / GCHUNK)) * GCHUNK_BYTES + ((d & 1u) ? GCHUNK * 16 : 0);
    reinterpret_cast<double2 *>(rec)[pos % GCHUNK] = make_double2(s1, s2);
  };
  for (unsigned long long k0 = beg; k0 < end; k0 += 32) {
    const unsigned long long k = k0 + lane;
    bool ok = false, ina = false, inb = false;
    int j = 0;
    double ra = 0.0, rb = 0.0;
    double4 sc = make_double4(0, 0, 0, 0);
    if (k < end) {
      j = neigh[k];
      const double4 xj = ld4(xq + j);
      double dx = xa.x - xj.x, dy = xa.y - xj.y, dz = xa.z - xj.z;
      ra = dx * dx + dy * dy + dz * dz;
      dx = xb.x - xj.x, dy = xb.y - xj.y, dz = xb.z - xj.z;
      rb = dx * dx + dy * dy + dz * dz;
      ina = j != a && ra < cutsq;
      inb = two && j != b && rb < cutsq;
      ok = ina || inb;
      if (!CHUNKED) {
        if (ina) radial_scalars<DAMP>(P.pc, ra, sc.x, sc.y);
        if (inb) radial_scalars<DAMP>(P.pc, rb, sc.z, sc.w);
      }
      if (RMIN && ((ina && ra < guard) || (inb && rb < guard))) {
        const double aj = mua[j].w;
        const int molj = tm[j].y;
        if (aj > 0) {
          if (ina && ala > 0 && (mola != molj || mola == 0)) {
            const double r = sqrt(rsq_nofma(xa.x - xj.x, xa.y - xj.y, xa.z - xj.z));
            if (best > r) best = r;
          }
          if (inb && alb > 0 && (molb != molj || molb == 0)) {
            const double r = sqrt(rsq_nofma(xb.x - xj.x, xb.y - xj.y, xb.z - xj.z));
            if (best > r) best = r;
          }
          guard = best * best * 1.000001;
        }
      }
    }
    const unsigned m = __ballot_sync(FULL, ok);
    const unsigned below = (1u << lane) - 1;
    const int pos = n + __popc(m & below);
    if (CHUNKED) {
      if (ok) {
        unsigned char *rec = crec + (cstart[g] + (unsigned long long)(pos / GCHUNK)) * GCHUNK_BYTES;
        reinterpret_cast<int *>(rec + GCHUNK * 32)[pos % GCHUNK] = j;
        if (!ina) reinterpret_cast<double2 *>(rec)[pos % GCHUNK] = make_double2(0.0, 0.0);
        if (!inb) reinterpret_cast<double2 *>(rec + GCHUNK * 16)[pos % GCHUNK] = make_double2(0.0, 0.0);
      }
      const unsigned ma = __ballot_sync(FULL, ina), mb = __ballot_sync(FULL, inb);
      if (ina) {
        const int q = qn + __popc(ma & below);
        q_dst[wq][q] = (unsigned)pos << 1;
        q_r2[wq][q] = ra;
      }
      qn += __popc(ma);
      if (inb) {
        const int q = qn + __popc(mb & below);
        q_dst[wq][q] = ((unsigned)pos << 1) | 1u;
        q_r2[wq][q] = rb;
      }
      qn += __popc(mb);
      __syncwarp();
      while (qn >= 32) {
        drain(qn - 32 + lane);
        qn -= 32;
      }
      __syncwarp();
    } else if (ok) {
      tneigh[beg + pos] = j;
      s12ab[beg + pos] = sc;
    }
    n += __popc(m);
  }
  if (CHUNKED && lane < qn) drain(lane);
  if (lane == 0) tcount[g] = n;
  if (RMIN) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_down_sync(FULL, best, o));
    if (lane == 0 && best < 1000.0) atomicMin(rmin_bits, (unsigned long long)__double_as_longlong(best));
  }
}

// one Jacobi dipole iteration, one warp per group of two atoms (rows in cell-sorted order only)
template <int WPB, int MINB, bool CHANGE, bool PUSH, bool DEEP>
__global__ void __launch_bounds__(WPB * 32, MINB)
k_sweep_group(int ngroups, const int *__restrict__ group_first, const int *__restrict__ group_two,
              const unsigned long long *__restrict__ rowstart, const int *__restrict__ tneigh,
              const int *__restrict__ tcount, const double4 *__restrict__ s12ab, const double4 *__restrict__ xq,
              const double4 *__restrict__ mu_in, const double4 *__restrict__ ef, double4 *__restrict__ mu_out,
              double *__restrict__ row_change, PushArgs Q, const int *stop = nullptr)
{
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * WPB + (threadIdx.x >> 5);
  if (g >= ngroups || scf_stopped(stop)) return;
  const int a = group_first[g];
  const bool two = group_two[g] != 0;
  const int b = two ? a + 1 : a;
  const double4 xa = xq[a], xb = xq[b];
  double4 *push_a = nullptr, *push_b = nullptr;
  if (PUSH) {
    push_a = push_prefetch(Q, a, lane);
    if (two) push_b = push_prefetch(Q, b, lane);
  }
  double eax = 0, eay = 0, eaz = 0, ebx = 0, eby = 0, ebz = 0;
  {
    const unsigned long long beg = rowstart[g];
    const int *__restrict__ row = tneigh + beg;
    const double4 *__restrict__ rs = s12ab + beg;
    auto ldcs4 = [&](int k) {
      double4 v;
      asm volatile("ld.global.cs.v4.b64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(rs + k));
      return v;
    };
    const int cnt = tcount[g];
    int jA = lane < cnt ? __ldcs(row + lane) : -1;
    int jB = lane + 32 < cnt ? __ldcs(row + lane + 32) : -1;
    int jC = lane + 64 < cnt ? __ldcs(row + lane + 64) : -1;
    double4 cA = lane < cnt ? ldcs4(lane) : make_double4(0, 0, 0, 0);
    double4 cB = (DEEP && lane + 32 < cnt) ? ldcs4(lane + 32) : make_double4(0, 0, 0, 0);
    for (int k = lane; k < cnt; k += 32) {
      const int jD = k + 96 < cnt ? __ldcs(row + k + 96) : -1;
      // cache scalars DEEP ? two : one trip(s) ahead: more bytes of the HBM stream in flight per warp
      double4 cN;
      if (DEEP) cN = k + 64 < cnt ? ldcs4(k + 64) : make_double4(0, 0, 0, 0);
      else cN = k + 32 < cnt ? ldcs4(k + 32) : make_double4(0, 0, 0, 0);
      const double4 xj = ld4(xq + jA);
      const double4 mj = ld4(mu_in + jA);
      {
        const double dx = xa.x - xj.x, dy = xa.y - xj.y, dz = xa.z - xj.z;
        const double t = cA.y * (dx * mj.x + dy * mj.y + dz * mj.z);
        eax -= fma(t, dx, cA.x * mj.x);
        eay -= fma(t, dy, cA.x * mj.y);
        eaz -= fma(t, dz, cA.x * mj.z);
      }
      {
        const double dx = xb.x - xj.x, dy = xb.y - xj.y, dz = xb.z - xj.z;
        const double t = cA.w * (dx * mj.x + dy * mj.y + dz * mj.z);
        ebx -= fma(t, dx, cA.z * mj.x);
        eby -= fma(t, dy, cA.z * mj.y);
        ebz -= fma(t, dz, cA.z * mj.z);
      }
      jA = jB;
      jB = jC;
      jC = jD;
      if (DEEP) {
        cA = cB;
        cB = cN;
      } else cA = cN;
    }
    eax = warp_sum(eax); eay = warp_sum(eay); eaz = warp_sum(eaz);
    ebx = warp_sum(ebx); eby = warp_sum(eby); ebz = warp_sum(ebz);
  }
  double nax = 0, nay = 0, naz = 0, nbx = 0, nby = 0, nbz = 0, ala = 0, alb = 0;
  if (lane == 0) {
    const double4 ma = mu_in[a], e = ef[a];
    ala = ma.w;
    nax = ma.w * (e.x + eax), nay = ma.w * (e.y + eay), naz = ma.w * (e.z + eaz);
    mu_out[a] = make_double4(nax, nay, naz, ma.w);
    if (CHANGE) row_change[a] = (nax - ma.x) * (nax - ma.x) + (nay - ma.y) * (nay - ma.y) + (naz - ma.z) * (naz - ma.z);
    if (two) {
      const double4 mb = mu_in[b], eb = ef[b];
      alb = mb.w;
      nbx = mb.w * (eb.x + ebx), nby = mb.w * (eb.y + eby), nbz = mb.w * (eb.z + ebz);
      mu_out[b] = make_double4(nbx, nby, nbz, mb.w);
      if (CHANGE) row_change[b] = (nbx - mb.x) * (nbx - mb.x) + (nby - mb.y) * (nby - mb.y) + (nbz - mb.z) * (nbz - mb.z);
    }
  }
  if (PUSH) {
    ala = __shfl_sync(FULL, ala, 0);
    push_row(push_a, nax, nay, naz, ala);
    if (two) {
      alb = __shfl_sync(FULL, alb, 0);
      push_row(push_b, nbx, nby, nbz, alb);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// TMA-fed pair-group sweep
// ---------------------------------------------------------------------------------------------------
// With the gathers halved by the pair groups, ncu shows nothing saturated (HBM 52 %, L1 data pipe 53 %) and
// 86 % long-scoreboard stalls: the sweep is limited by the bytes of the HBM stream (neighbour indices +
// cached radial scalars, 36 B per entry) that register prefetching can keep in flight (~30 KB per SM against
// the ~60 KB the latency-bandwidth product asks for).  Here each warp streams its row through a ring of
// shared-memory stages filled by 1-D bulk async copies (cp.async.bulk, the TMA engine) that complete on
// mbarriers: up to NSTAGE*CHUNK*36 B per warp are in flight without costing a single register, and the
// indices of a whole chunk are available at once, so the gathers of all its trips are issued back to back.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// same copy, marked evict-first in L2: the row streams are read once per sweep and must not push the
// position / dipole records (gathered ~700 times each per sweep) out of the 126 MB L2
__device__ __forceinline__ unsigned long long l2_evict_first_policy()
{
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_stream(void *dst, const void *src, unsigned bytes, unsigned long long *bar,
                                                unsigned long long pol)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// GS = true: one chunk ("colour") of the group-coloured Gauss-Seidel sweep.  The warps walk the groups
// glist[gs.beg .. gs.end) instead of all groups, read the dipoles in place (mu_in = the live array), write the new
// dipoles of their rows to a staging array (mu_out; committed by k_commit_groups when the whole chunk is done, so the
// chunk is a Jacobi step and the result does not depend on the scheduling of the warps), and the two members of a
// group are updated one after the other: the second member sees the first member's NEW dipole (their mutual tensor
// is evaluated once per row from the positions -- the same radial_scalars() the per-step cache stores).
struct GsChunk {
  const int *glist;   // groups sorted by colour
  int beg, end;       // this chunk's slice of glist
  double polar_damp, polar_cutsq;
  int damping_exponential;
};

template <int WPB, int NSTAGE, int CHUNK, bool CHANGE, bool PUSH, bool EVICT, bool GS = false>
__global__ void __launch_bounds__(WPB * 32)
k_sweep_group_tma(int ngroups, const int *__restrict__ group_first, const int *__restrict__ group_two,
                  const unsigned long long *__restrict__ rowstart, const int *__restrict__ tneigh,
                  const int *__restrict__ tcount, const double4 *__restrict__ s12ab, const double4 *__restrict__ xq,
                  const double4 *__restrict__ mu_in, const double4 *__restrict__ ef, double4 *__restrict__ mu_out,
                  double *__restrict__ row_change, PushArgs Q, int *dbg, int reverse,
                  const unsigned long long *__restrict__ cstart, const unsigned char *__restrict__ crec,
                  const int *stop = nullptr, GsChunk gs = GsChunk{})
{
  if (scf_stopped(stop)) return;
  if (GS) ngroups = gs.end - gs.beg;
  static_assert(CHUNK == GCHUNK, "the chunk-record layout is built for 64-entry chunks");
  // stage layout: CHUNK x 32 B scalars {s1a,s2a,s1b,s2b}, then CHUNK x 4 B indices.  (Splitting the two members
  // into separate 16-byte streams makes the LDS conflict free but needs a third bulk copy per chunk: measured slower.)
  constexpr int STAGE_BYTES = CHUNK * 36;
  constexpr int TRIPS = CHUNK / 32;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *ring = smem + (size_t)warp * NSTAGE * STAGE_BYTES;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + (size_t)WPB * NSTAGE * STAGE_BYTES) + warp * NSTAGE;
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < NSTAGE; st++) mbar_init(bars + st, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  const unsigned long long l2pol = EVICT ? l2_evict_first_policy() : 0ull;
  unsigned cc = 0;  // chunks consumed by this warp so far: stage = cc % NSTAGE, parity = (cc / NSTAGE) & 1
  const int nwarps = gridDim.x * WPB;
  // `reverse` alternates from sweep to sweep: the 126 MB L2 still holds the tail of the stream the previous
  // sweep read last, so walking the groups in the opposite direction turns the first ~quarter of this sweep's
  // HBM stream into L2 hits (a Jacobi sweep does not care about the order of its rows)
  for (int gi = blockIdx.x * WPB + warp; gi < ngroups; gi += nwarps) {
    const int g = GS ? gs.glist[gs.beg + gi] : (reverse ? ngroups - 1 - gi : gi);
    const int a = group_first[g];
    const int gflag = group_two[g];  // bit 0: two members; bit 1 (GS): member b is updated first
    const bool two = (gflag & 1) != 0;
    const int b = two ? a + 1 : a;
    const int cnt = tcount[g];
    const int nchunks = (cnt + CHUNK - 1) / CHUNK;
    const unsigned char *__restrict__ recs = crec + cstart[g] * GCHUNK_BYTES;
    auto issue = [&](int c, unsigned slot) {  // lane 0: chunk record c of this row into ring slot `slot`
      unsigned char *dst = ring + (size_t)slot * STAGE_BYTES;
      mbar_expect_tx(bars + slot, (unsigned)GCHUNK_BYTES);
      if (EVICT) bulk_g2s_stream(dst, recs + (size_t)c * GCHUNK_BYTES, GCHUNK_BYTES, bars + slot, l2pol);
      else bulk_g2s(dst, recs + (size_t)c * GCHUNK_BYTES, GCHUNK_BYTES, bars + slot);
    };
    if (lane == 0) {
      const int pre = min(NSTAGE, nchunks);
      for (int c = 0; c < pre; c++) issue(c, (cc + c) % NSTAGE);
    }
    const double4 xa = xq[a], xb = xq[b];
    double4 *push_a = nullptr, *push_b = nullptr;
    if (PUSH) {
      push_a = push_prefetch(Q, a, lane);
      if (two) push_b = push_prefetch(Q, b, lane);
    }
    double eax = 0, eay = 0, eaz = 0, ebx = 0, eby = 0, ebz = 0;
    for (int c = 0; c < nchunks; c++, cc++) {
      const unsigned slot = cc % NSTAGE;
      mbar_wait(bars + slot, (cc / NSTAGE) & 1u);
      const double2 *sa_s = reinterpret_cast<const double2 *>(ring + (size_t)slot * STAGE_BYTES);
      const double2 *sb_s = sa_s + CHUNK;
      const int *ix_s = reinterpret_cast<const int *>(ring + (size_t)slot * STAGE_BYTES + CHUNK * 32);
      int j[TRIPS];
      double4 sc[TRIPS], xj[TRIPS], mj[TRIPS];
#pragma unroll
      for (int t = 0; t < TRIPS; t++) {
        const bool live = c * CHUNK + t * 32 + lane < cnt;
        j[t] = live ? ix_s[t * 32 + lane] : a;
        const double2 ua = live ? sa_s[t * 32 + lane] : make_double2(0, 0);
        const double2 ub = live ? sb_s[t * 32 + lane] : make_double2(0, 0);
        sc[t] = make_double4(ua.x, ua.y, ub.x, ub.y);
      }
#pragma unroll
      for (int t = 0; t < TRIPS; t++) {
        xj[t] = ld4(xq + j[t]);
        mj[t] = ld4(mu_in + j[t]);
      }
#pragma unroll
      for (int t = 0; t < TRIPS; t++) {
        {
          const double dx = xa.x - xj[t].x, dy = xa.y - xj[t].y, dz = xa.z - xj[t].z;
          const double q = sc[t].y * (dx * mj[t].x + dy * mj[t].y + dz * mj[t].z);
          eax -= fma(q, dx, sc[t].x * mj[t].x);
          eay -= fma(q, dy, sc[t].x * mj[t].y);
          eaz -= fma(q, dz, sc[t].x * mj[t].z);
        }
        {
          const double dx = xb.x - xj[t].x, dy = xb.y - xj[t].y, dz = xb.z - xj[t].z;
          const double q = sc[t].w * (dx * mj[t].x + dy * mj[t].y + dz * mj[t].z);
          ebx -= fma(q, dx, sc[t].z * mj[t].x);
          eby -= fma(q, dy, sc[t].z * mj[t].y);
          ebz -= fma(q, dz, sc[t].z * mj[t].z);
        }
      }
      // Refill only after the arithmetic that consumes this slot's scalars: the compiler is free to sink the
      // shared-memory loads of sc[] down to their first use, behind the long wait for the gathers, and a bulk
      // copy issued earlier could overwrite the slot before they execute (seen on hardware: wrong dipoles).
      __syncwarp();
      if (lane == 0 && c + NSTAGE < nchunks) issue(c + NSTAGE, slot);
    }
    eax = warp_sum(eax); eay = warp_sum(eay); eaz = warp_sum(eaz);
    ebx = warp_sum(ebx); eby = warp_sum(eby); ebz = warp_sum(ebz);
    double nax = 0, nay = 0, naz = 0, nbx = 0, nby = 0, nbz = 0, ala = 0, alb = 0;
    if (lane == 0) {
      const double4 ma = mu_in[a], e = ef[a];
      ala = ma.w;
      if (GS && two) {
        // in-group Gauss-Seidel: first member from the old dipoles, then the second member with the first one's
        // change folded into its field:  E_2 -= T_21 (mu_1_new - mu_1_old)
        const double4 mb = mu_in[b], eb = ef[b];
        alb = mb.w;
        const double dx = xa.x - xb.x, dy = xa.y - xb.y, dz = xa.z - xb.z;
        const double r2 = dx * dx + dy * dy + dz * dz;
        double s1 = 0.0, s2 = 0.0;
        if (r2 < gs.polar_cutsq) {
          PairConsts pc;
          pc.polar_damp = gs.polar_damp;
          if (gs.damping_exponential) radial_scalars<true>(pc, r2, s1, s2);
          else radial_scalars<false>(pc, r2, s1, s2);
        }
        if (gflag & 2) {
          nbx = mb.w * (eb.x + ebx), nby = mb.w * (eb.y + eby), nbz = mb.w * (eb.z + ebz);
          const double ux = nbx - mb.x, uy = nby - mb.y, uz = nbz - mb.z;
          const double t = s2 * (dx * ux + dy * uy + dz * uz);
          eax -= fma(t, dx, s1 * ux), eay -= fma(t, dy, s1 * uy), eaz -= fma(t, dz, s1 * uz);
          nax = ma.w * (e.x + eax), nay = ma.w * (e.y + eay), naz = ma.w * (e.z + eaz);
        } else {
          nax = ma.w * (e.x + eax), nay = ma.w * (e.y + eay), naz = ma.w * (e.z + eaz);
          const double ux = nax - ma.x, uy = nay - ma.y, uz = naz - ma.z;
          const double t = s2 * (dx * ux + dy * uy + dz * uz);
          ebx -= fma(t, dx, s1 * ux), eby -= fma(t, dy, s1 * uy), ebz -= fma(t, dz, s1 * uz);
          nbx = mb.w * (eb.x + ebx), nby = mb.w * (eb.y + eby), nbz = mb.w * (eb.z + ebz);
        }
        mu_out[a] = make_double4(nax, nay, naz, ma.w);
        mu_out[b] = make_double4(nbx, nby, nbz, mb.w);
        if (CHANGE) {
          row_change[a] = (nax - ma.x) * (nax - ma.x) + (nay - ma.y) * (nay - ma.y) + (naz - ma.z) * (naz - ma.z);
          row_change[b] = (nbx - mb.x) * (nbx - mb.x) + (nby - mb.y) * (nby - mb.y) + (nbz - mb.z) * (nbz - mb.z);
        }
      } else {
        nax = ma.w * (e.x + eax), nay = ma.w * (e.y + eay), naz = ma.w * (e.z + eaz);
        mu_out[a] = make_double4(nax, nay, naz, ma.w);
        if (CHANGE) row_change[a] = (nax - ma.x) * (nax - ma.x) + (nay - ma.y) * (nay - ma.y) + (naz - ma.z) * (naz - ma.z);
        if (two) {
          const double4 mb = mu_in[b], eb = ef[b];
          alb = mb.w;
          nbx = mb.w * (eb.x + ebx), nby = mb.w * (eb.y + eby), nbz = mb.w * (eb.z + ebz);
          mu_out[b] = make_double4(nbx, nby, nbz, mb.w);
          if (CHANGE) row_change[b] = (nbx - mb.x) * (nbx - mb.x) + (nby - mb.y) * (nby - mb.y) + (nbz - mb.z) * (nbz - mb.z);
        }
      }
    }
    if (PUSH) {
      ala = __shfl_sync(FULL, ala, 0);
      push_row(push_a, nax, nay, naz, ala);
      if (two) {
        alb = __shfl_sync(FULL, alb, 0);
        push_row(push_b, nbx, nby, nbz, alb);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// stages 2 and 4 on the pair-group rows
// ---------------------------------------------------------------------------------------------------
// When every interaction of the step reaches exactly as far as the dipole cutoff (max pair cutoff == cut_coul ==
// polar_cutoff: the shape of all BASELINE configs), the per-step tight group rows that the sweep streams hold every
// partner the LJ + Coulomb + static-field kernel and the polarization-force kernel need.  One warp serves the two
// members of a group: each gathered position / dipole / type record feeds two pair evaluations (half the gathers of the
// per-atom kernels), the per-atom tight list (k_tighten) is not built at all, and the index stream is the 256-byte index
// block of every 2304-byte chunk record.
__device__ __forceinline__ const int *group_index_ptr(const unsigned char *__restrict__ recs, int k)
{
  return reinterpret_cast<const int *>(recs + (size_t)(k >> 6) * GCHUNK_BYTES + GCHUNK * 32) + (k & 63);
}

constexpr int GPF_WARPS = 4;  // warps per CTA of the two kernels below: 5 CTAs of 128 threads per SM leave 102 registers
template <bool EVFLAG, int MINB>
__global__ void __launch_bounds__(GPF_WARPS * 32, MINB)
k_pair_group(int ngroups, DevParams P, const int *__restrict__ group_first, const int *__restrict__ group_two,
             const int *__restrict__ tcount, const unsigned long long *__restrict__ cstart,
             const unsigned char *__restrict__ crec, const double4 *__restrict__ xq, const int2 *__restrict__ tm,
             double4 *__restrict__ f_pair, double4 *__restrict__ ef, double *__restrict__ partial)
{
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * GPF_WARPS + (threadIdx.x >> 5);
  double acc[NPAIR_PART] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (g < ngroups) {
    const int a = group_first[g];
    const bool two = (group_two[g] & 1) != 0;
    const int b = two ? a + 1 : a;
    const double4 xa = xq[a], xb = xq[b];
    const int2 tma = tm[a], tmb = tm[b];
    const int n1 = P.pc.ntypes + 1;
    double fax = 0, fay = 0, faz = 0, fbx = 0, fby = 0, fbz = 0;
    double eax = 0, eay = 0, eaz = 0, ebx = 0, eby = 0, ebz = 0;
    const int cnt = tcount[g];
    const unsigned char *__restrict__ recs = crec + cstart[g] * GCHUNK_BYTES;
    int jA = lane < cnt ? __ldcs(group_index_ptr(recs, lane)) : 0;
    int jB = lane + 32 < cnt ? __ldcs(group_index_ptr(recs, lane + 32)) : 0;
    auto one = [&](const double4 &xi, const int2 &tmi, const double4 &xj, const int2 &tmj, double &fx, double &fy, double &fz,
                   double &ex, double &ey, double &ez) {
      const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
      const double rsq = rsq_nofma(dx, dy, dz);
      const int ij = tmi.x * n1 + tmj.x;
      if (rsq < P.lj.cutsq[ij]) {
        double evdwl, ecoul;
        const double fpair = lj_coul_pair(P.pc, P.lj, P.tb, ij, rsq, xi.w, xj.w, 0, EVFLAG, evdwl, ecoul);
        fx += dx * fpair;
        fy += dy * fpair;
        fz += dz * fpair;
        if (EVFLAG) {
          acc[0] += evdwl;
          acc[1] += ecoul;
          acc[2] += dx * dx * fpair;
          acc[3] += dy * dy * fpair;
          acc[4] += dz * dz * fpair;
          acc[5] += dx * dy * fpair;
          acc[6] += dx * dz * fpair;
          acc[7] += dy * dz * fpair;
        }
      }
      if (rsq <= P.pc.cut_coulsq && (tmi.y != tmj.y || tmi.y == 0)) {
        const double sc = static_field_scalar(P.pc, rsq) * xj.w;
        ex += sc * dx;
        ey += sc * dy;
        ez += sc * dz;
      }
    };
    for (int k = lane; k < cnt; k += 32) {
      const int jC = k + 64 < cnt ? __ldcs(group_index_ptr(recs, k + 64)) : 0;
      const int j = jA;
      jA = jB;
      jB = jC;
      const double4 xj = ld4(xq + j);
      const int2 tmj = tm[j];
      if (j != a) one(xa, tma, xj, tmj, fax, fay, faz, eax, eay, eaz);
      if (two && j != b) one(xb, tmb, xj, tmj, fbx, fby, fbz, ebx, eby, ebz);
    }
    fax = warp_sum(fax); fay = warp_sum(fay); faz = warp_sum(faz);
    eax = warp_sum(eax); eay = warp_sum(eay); eaz = warp_sum(eaz);
    if (two) {
      fbx = warp_sum(fbx); fby = warp_sum(fby); fbz = warp_sum(fbz);
      ebx = warp_sum(ebx); eby = warp_sum(eby); ebz = warp_sum(ebz);
    }
    if (lane == 0) {
      f_pair[a] = make_double4(fax, fay, faz, 0.0);
      ef[a] = make_double4(eax * P.pc.kq, eay * P.pc.kq, eaz * P.pc.kq, 0.0);  // pol.cpp:372-374
      if (two) {
        f_pair[b] = make_double4(fbx, fby, fbz, 0.0);
        ef[b] = make_double4(ebx * P.pc.kq, eby * P.pc.kq, ebz * P.pc.kq, 0.0);
      }
    }
  }
  if (EVFLAG) {
#pragma unroll
    for (int k = 0; k < NPAIR_PART; k++) acc[k] = 0.5 * warp_sum(acc[k]);
    block_reduce_store<NPAIR_PART, GPF_WARPS>(acc, partial);
  }
}

template <bool EVFLAG, int MINB>
__global__ void __launch_bounds__(GPF_WARPS * 32, MINB)
k_polforce_group(int ngroups, DevParams P, const int *__restrict__ group_first, const int *__restrict__ group_two,
                 const int *__restrict__ tcount, const unsigned long long *__restrict__ cstart,
                 const unsigned char *__restrict__ crec, const double4 *__restrict__ xq, const double4 *__restrict__ mua,
                 const int2 *__restrict__ tm, double4 *__restrict__ f_pol, double *__restrict__ partial)
{
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * GPF_WARPS + (threadIdx.x >> 5);
  double acc[NPOL_PART] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (g < ngroups) {
    const int a = group_first[g];
    const bool two = (group_two[g] & 1) != 0;
    const int b = two ? a + 1 : a;
    const double4 xa = xq[a], xb = xq[b];
    const double4 ma = mua[a], mb = mua[b];
    const int mola = tm[a].y, molb = tm[b].y;
    const bool molecules = P.pc.has_molecules != 0;
    double fax = 0, fay = 0, faz = 0, fbx = 0, fby = 0, fbz = 0;
    const int cnt = tcount[g];
    const unsigned char *__restrict__ recs = crec + cstart[g] * GCHUNK_BYTES;
    const double reach = fmax(P.pc.cut_coulsq, P.pc.polar_cutsq);
    int jA = lane < cnt ? __ldcs(group_index_ptr(recs, lane)) : 0;
    int jB = lane + 32 < cnt ? __ldcs(group_index_ptr(recs, lane + 32)) : 0;
    const bool damp = P.pc.damping_exponential != 0;
    auto one = [&](const double4 &xi, const double4 &mi, int moli, const double4 &xj, const double4 &mj, int molj, double &fx,
                   double &fy, double &fz) {
      const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
      if (dx * dx + dy * dy + dz * dz < reach) {
        const bool inter = !molecules || (moli != molj) || moli == 0;
        double px, py, pz, uef, udd;
        pol_force_pair_fast(P.pc, damp, dx, dy, dz, xi.w, xj.w, mi.w, mj.w, mi.x, mi.y, mi.z, mj.x, mj.y, mj.z, inter, EVFLAG,
                            px, py, pz, uef, udd);
        fx += px; fy += py; fz += pz;
        if (EVFLAG) {
          acc[1] += 0.5 * uef;  // every pair is visited from both of its atoms
          acc[2] += 0.5 * udd;
        }
      }
    };
    for (int k = lane; k < cnt; k += 32) {
      const int jC = k + 64 < cnt ? __ldcs(group_index_ptr(recs, k + 64)) : 0;
      const int j = jA;
      jA = jB;
      jB = jC;
      const double4 xj = ld4(xq + j);
      const double4 mj = ld4(mua + j);
      const int molj = molecules ? tm[j].y : 0;
      if (j != a) one(xa, ma, mola, xj, mj, molj, fax, fay, faz);
      if (two && j != b) one(xb, mb, molb, xj, mj, molj, fbx, fby, fbz);
    }
    fax = warp_sum(fax); fay = warp_sum(fay); faz = warp_sum(faz);
    if (two) { fbx = warp_sum(fbx); fby = warp_sum(fby); fbz = warp_sum(fbz); }
    if (lane == 0) {
      f_pol[a] = make_double4(fax, fay, faz, 0.0);
      if (two) f_pol[b] = make_double4(fbx, fby, fbz, 0.0);
      if (EVFLAG) {
        if (ma.w != 0.0) acc[0] = 0.5 * (ma.x * ma.x + ma.y * ma.y + ma.z * ma.z) / ma.w;  // pol.cpp:432-433
        // F.r virial of the reference (src/pair.cpp:1495-1543): polarization forces act on owned atoms
        acc[3] += fax * xa.x; acc[4] += fay * xa.y; acc[5] += faz * xa.z;
        acc[6] += fay * xa.x; acc[7] += faz * xa.x; acc[8] += faz * xa.y;
        if (two) {
          if (mb.w != 0.0) acc[0] += 0.5 * (mb.x * mb.x + mb.y * mb.y + mb.z * mb.z) / mb.w;
          acc[3] += fbx * xb.x; acc[4] += fby * xb.y; acc[5] += fbz * xb.z;
          acc[6] += fby * xb.x; acc[7] += fbz * xb.x; acc[8] += fbz * xb.y;
        }
      }
    }
  }
  if (EVFLAG) {
#pragma unroll
    for (int k = 0; k < NPOL_PART; k++) acc[k] = warp_sum(acc[k]);
    block_reduce_store<NPOL_PART, GPF_WARPS>(acc, partial);
  }
}

// ---------------------------------------------------------------------------------------------------
// group-coloured Gauss-Seidel (list mode polar_gs / polar_gs_ranked): colouring of the pair groups
// ---------------------------------------------------------------------------------------------------
// The ranked sweep of the reference (pol.cpp:1158-1180) is sequential.  Its device realisation is a multi-colour
// sweep: the pair groups are coloured so that groups whose members sit close to each other (the strong couplings:
// the sites of one molecule, nearest neighbours) get DIFFERENT colours; one sweep visits the colours in turn --
// Jacobi among the groups of a colour, Gauss-Seidel between colours and between the two members of a group.  Same
// fixed point as the reference's sweep; on the rigid-water box it converges in 14 iterations against 16 for the
// strictly sequential ranked sweep and 18 for the former per-atom interleaved colouring.  The rank metric keeps its
// role as the priority of the greedy colouring: groups with the larger metric choose their colour first.
constexpr int GS_MAXADJ = 32;

// group of every owned atom + which member of a two-member group is updated first (higher rank metric; ties: lower
// caller index, the reference's stable order); metric_caller == nullptr: caller index order (polar_gs)
__global__ void k_group_of(int ngroups, const int *__restrict__ group_first, int *__restrict__ group_two,
                           const int *__restrict__ perm, const double *__restrict__ metric_caller, int *__restrict__ gof,
                           float *__restrict__ gmetric)
{
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  const int a = group_first[g];
  const bool two = (group_two[g] & 1) != 0;
  gof[a] = g;
  const int ca = perm[a];
  double ma = metric_caller ? metric_caller[ca] : 0.0, mg = ma;
  int flag = two ? 1 : 0;
  if (two) {
    gof[a + 1] = g;
    const int cb = perm[a + 1];
    const double mb = metric_caller ? metric_caller[cb] : 0.0;
    if (mb > ma || (mb == ma && cb < ca)) flag |= 2;
    mg = fmax(ma, mb);
  }
  group_two[g] = flag;
  gmetric[g] = (float)mg;
}

__device__ __forceinline__ unsigned hash32(unsigned x)
{
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// strong-coupling adjacency of the groups: every other group with a member closer than `rs` to a member of this
// one (owned partners and this box's own periodic images; ghosts of other bricks are left out -- block-Jacobi
// across bricks).  One warp per group over its skin row.
__global__ void __launch_bounds__(BLOCK)
k_group_adjacency(int ngroups, int nloc, double rs2, const int *__restrict__ group_first, const int *__restrict__ group_two,
                  const unsigned long long *__restrict__ rowstart, const int *__restrict__ rowcount,
                  const int *__restrict__ neigh, const double4 *__restrict__ xq, const int *__restrict__ gof,
                  const int *__restrict__ g_owner, int *__restrict__ adj, int *__restrict__ adjn)
{
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (g >= ngroups) return;
  const int a = group_first[g];
  const int b = (group_two[g] & 1) ? a + 1 : a;
  const double4 xa = xq[a], xb = xq[b];
  const unsigned long long beg = rowstart[g], end = beg + (unsigned long long)rowcount[g];
  int n = 0;
  for (unsigned long long k0 = beg; k0 < end; k0 += 32) {
    const unsigned long long k = k0 + lane;
    int h = -1;
    if (k < end) {
      const int j = neigh[k];
      const double4 xj = ld4(xq + j);
      double dx = xa.x - xj.x, dy = xa.y - xj.y, dz = xa.z - xj.z;
      const double ra = dx * dx + dy * dy + dz * dz;
      dx = xb.x - xj.x, dy = xb.y - xj.y, dz = xb.z - xj.z;
      const double rb = dx * dx + dy * dy + dz * dz;
      if (ra < rs2 || rb < rs2) {
        const int o = j < nloc ? j : (g_owner ? g_owner[j - nloc] : -1);
        if (o >= 0) h = gof[o];
        if (h == g) h = -1;
      }
    }
    const unsigned m = __ballot_sync(FULL, h >= 0);
    if (h >= 0) {
      const int pos = n + __popc(m & ((1u << lane) - 1));
      if (pos < GS_MAXADJ) adj[(size_t)g * GS_MAXADJ + pos] = h;
    }
    n += __popc(m);
  }
  if (lane == 0) adjn[g] = min(n, GS_MAXADJ);
}

// priority of a group in the greedy colouring: larger rank metric first, ties by a hash of the group index
__device__ __forceinline__ unsigned long long group_priority(float metric, int g)
{
  return ((unsigned long long)__float_as_uint(metric) << 32) | (unsigned long long)(~hash32((unsigned)g));
}

// one round of the parallel greedy colouring (Jones-Plassmann): a group whose stronger-priority neighbours are all
// coloured takes the first colour, counted cyclically from its preferred one, that none of them uses (its preferred
// colour when all are taken).  The result is the colouring of the sequential greedy algorithm in priority order,
// whatever the number of rounds and the interleaving of the threads.
__global__ void k_colour_round(int ngroups, int ncolours, const int *__restrict__ adj, const int *__restrict__ adjn,
                               const float *__restrict__ gmetric, int *colour, int *__restrict__ remaining)
{
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  if (*(volatile int *)(colour + g) >= 0) return;
  const unsigned long long pg = group_priority(gmetric[g], g);
  unsigned used = 0;
  const int n = adjn[g];
  for (int k = 0; k < n; k++) {
    const int h = adj[(size_t)g * GS_MAXADJ + k];
    const unsigned long long ph = group_priority(gmetric[h], h);
    if (ph > pg) {
      const int c = *(volatile int *)(colour + h);
      if (c < 0) {  // a stronger neighbour is still undecided: next round
        atomicAdd(remaining, 1);
        return;
      }
      used |= 1u << c;
    }
  }
  const int c0 = (int)(hash32((unsigned)g * 2654435761u + 12345u) % (unsigned)ncolours);
  int c = c0;
  for (int t = 0; t < ncolours; t++) {
    const int cc = c0 + t < ncolours ? c0 + t : c0 + t - ncolours;
    if (!((used >> cc) & 1u)) {
      c = cc;
      break;
    }
  }
  *(volatile int *)(colour + g) = c;
}

__global__ void k_iota(int n, int *__restrict__ out)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = t;
}

// colour of every owned atom in caller order + its in-group predecessor (caller index of the member that is updated
// before it, -1 for first members and singles): what the oracle needs to replay the same sweep (tests)
__global__ void k_colour_export(int ngroups, const int *__restrict__ group_first, const int *__restrict__ group_two,
                                const int *__restrict__ colour, const int *__restrict__ perm, int *__restrict__ out_colour,
                                int *__restrict__ out_after)
{
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  const int a = group_first[g], fl = group_two[g], c = colour[g];
  const int ca = perm[a];
  out_colour[ca] = c;
  out_after[ca] = -1;
  if (fl & 1) {
    const int cb = perm[a + 1];
    out_colour[cb] = c;
    out_after[cb] = -1;
    if (fl & 2) out_after[ca] = cb;
    else out_after[cb] = ca;
  }
}

// commit one chunk of the group-coloured sweep: staged dipoles become visible, and their ghost copies (own periodic
// images / neighbour bricks' ghost slots) are refreshed through the push tables.  Two threads per group.
__global__ void k_commit_groups(GsChunk gs, const int *__restrict__ group_first, const int *__restrict__ group_two,
                                const double4 *__restrict__ staged, double4 *__restrict__ mua, PushArgs Q, int push,
                                const int *stop)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int gi = gs.beg + (t >> 1);
  if (gi >= gs.end || scf_stopped(stop)) return;
  const int g = gs.glist[gi];
  if ((t & 1) && !(group_two[g] & 1)) return;
  const int s = group_first[g] + (t & 1);
  const double4 v = staged[s];
  mua[s] = v;
  if (push) {
    const unsigned long long b = Q.off[s], e = Q.off[s + 1];
    for (unsigned long long u = b; u < e; u++) *Q.ptr[u] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// convergence test on the device (pol.cpp:1193-1237)
// ---------------------------------------------------------------------------------------------------
// ctl[0] stop flag, ctl[1] completed iterations, ctl[2] diverged, ctl[3] ticket of the two-level sum.
// k_change_sum: fixed slices of the per-row squared changes are summed by SCF_SUM_BLOCKS CTAs (fixed order inside a
// slice), the CTA that finishes last adds the slice sums in slice order and -- single GPU -- runs the test at once.
// Decomposed runs all-reduce the sum first (k_signal_wait / NCCL) and then launch k_scf_check.
constexpr int SCF_SUM_BLOCKS = 64;

__device__ __forceinline__ void scf_test(double sum, int *ctl, double prec2, double norm3n, int itmax)
{
  const double change = sum / norm3n;
  const bool keep = change > prec2;             // pol.cpp:1205-1209
  const int it = ctl[1] + 1;                    // "iterations++" after the copy (pol.cpp:1224)
  ctl[1] = it;
  if (it > itmax) {                             // pol.cpp:1227-1235: reset to alpha*E, warning, return
    ctl[2] = 1;
    ctl[0] = 1;
  } else if (!keep) ctl[0] = 1;
}

__global__ void __launch_bounds__(256)
k_change_sum(int n, const double *__restrict__ row_change, double *__restrict__ slice, double *__restrict__ out, int *ctl,
             int test, double prec2, double norm3n, int itmax, const int *stop)
{
  if (scf_stopped(stop)) return;
  __shared__ double sm[256];
  __shared__ int last;
  const int per = (n + SCF_SUM_BLOCKS - 1) / SCF_SUM_BLOCKS;
  const int beg = blockIdx.x * per, end = min(n, beg + per);
  double t = 0.0;
  for (int k = beg + threadIdx.x; k < end; k += 256) t += row_change[k];
  sm[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    slice[blockIdx.x] = sm[0];
    __threadfence();
    last = atomicAdd(ctl + 3, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!last || threadIdx.x != 0) return;
  __threadfence();
  ctl[3] = 0;
  double s = 0.0;
  for (int b = 0; b < (int)gridDim.x; b++) s += *(volatile double *)(slice + b);
  out[0] = s;
  if (test) scf_test(s, ctl, prec2, norm3n, itmax);
}

// test alone: the sum is already in *sum (exact-mode sweeps, all-reduced sums of the decomposed runs)
__global__ void k_scf_check(const double *__restrict__ sum, int *ctl, double prec2, double norm3n, int itmax)
{
  if (threadIdx.x == 0 && blockIdx.x == 0 && !scf_stopped(ctl)) scf_test(sum[0], ctl, prec2, norm3n, itmax);
}

// ---------------------------------------------------------------------------------------------------
// exact mode: sequential-equivalent Gauss-Seidel as a blocked forward substitution
// ---------------------------------------------------------------------------------------------------
// The reference's polar_gs / polar_gs_ranked sweep (pol.cpp:1158-1180) visits the atoms one by one in ranked
// order, each using the NEW dipoles of the atoms before it and the OLD dipoles of the atoms after it: one sweep
// solves the lower-triangular system  (I + alpha L) mu_new = alpha (E - U mu_old)  in ranked order.  Done atom by
// atom that is N dependent steps (k_gs_sequential: one CTA, 8 ms per sweep set at 750 atoms).  Blocked:
//   k_gsb_upper : R[p] = E[p] - sum over partners in LATER blocks of T mu_old          (all SMs, once per sweep)
//   per block b : k_gsb_step   - R[p] -= sum over the atoms of block b-1 of T mu_new for every row from block b on, then
//                                the 32x32 diagonal block: T of the block's pairs into shared memory, one warp substitutes
//                                sequentially (new dipoles of earlier atoms, old of later ones)
// Same operands per atom as the sequential sweep; only the order of the additions inside one field sum differs.
constexpr int GSB = 32;

// del of the pair (row atom s, partner j) with the reference's anchoring rule (lower caller index anchors the image)
__device__ __forceinline__ void pair_del(const Box &box, const int *__restrict__ perm, int s, int j, const double4 &xs,
                                         const double4 &xj, double &dx, double &dy, double &dz)
{
  if (perm[s] < perm[j]) min_image_del(box, xs.x, xs.y, xs.z, xj.x, xj.y, xj.z, dx, dy, dz);
  else min_image_del(box, xj.x, xj.y, xj.z, xs.x, xs.y, xs.z, dx, dy, dz);
}

// Per-step cache of the pair tensors in RANKED order for small systems: {s1, s2, del} of every ordered pair of ranked
// positions (p, q), five planes of n x n doubles (22 MB at 750 atoms).  The geometry and the order are frozen during the
// SCF, and the blocked sweep is a chain of dependent few-microsecond kernels: a square root, two divisions and an
// exponential per pair on that chain (~2000 cycles, twice per block) were a third of a sweep.  plane == nullptr: no cache.
struct GsPairCache {
  const double *plane;  // [5][n][n]
  int n;
  __device__ __forceinline__ bool on() const { return plane != nullptr; }
  __device__ __forceinline__ void load(int p, int q, double &s1, double &s2, double &dx, double &dy, double &dz) const
  {
    const size_t nn = (size_t)n * n, k = (size_t)p * n + q;
    s1 = plane[k]; s2 = plane[nn + k]; dx = plane[2 * nn + k]; dy = plane[3 * nn + k]; dz = plane[4 * nn + k];
  }
};

__global__ void __launch_bounds__(BLOCK)
k_gs_pair_cache(int n, const int *__restrict__ order, DevParams P, const int *__restrict__ perm,
                const double4 *__restrict__ xq, double *__restrict__ plane)
{
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (p >= n) return;
  const int s = order ? order[p] : p;
  const double4 xs = xq[s];
  const size_t nn = (size_t)n * n;
  for (int q = lane; q < n; q += 32) {
    double dx = 0, dy = 0, dz = 0, s1 = 0, s2 = 0;
    if (q != p) {
      const int j = order ? order[q] : q;
      pair_del(P.box, perm, s, j, xs, ld4(xq + j), dx, dy, dz);
      induced_field_scalars(P.pc, dx * dx + dy * dy + dz * dz, s1, s2);
    }
    const size_t k = (size_t)p * n + q;
    plane[k] = s1; plane[nn + k] = s2; plane[2 * nn + k] = dx; plane[3 * nn + k] = dy; plane[4 * nn + k] = dz;
  }
}

// e -= T(p, q) mu  from the cache
__device__ __forceinline__ void cached_field_pair(const GsPairCache &C, int p, int q, double mx, double my, double mz,
                                                  double &ex, double &ey, double &ez)
{
  double s1, s2, dx, dy, dz;
  C.load(p, q, s1, s2, dx, dy, dz);
  const double dm = dx * mx + dy * my + dz * mz;
  ex -= s1 * mx + s2 * dm * dx;
  ey -= s1 * my + s2 * dm * dy;
  ez -= s1 * mz + s2 * dm * dz;
}

__global__ void __launch_bounds__(BLOCK)
k_gsb_upper(int n, const int *__restrict__ order, DevParams P, const int *__restrict__ perm,
            const double4 *__restrict__ xq, const double4 *__restrict__ mua, const double4 *__restrict__ ef,
            double4 *__restrict__ R, const int *stop = nullptr, GsPairCache C = GsPairCache{nullptr, 0})
{
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (p >= n || scf_stopped(stop)) return;
  const int s = order ? order[p] : p;
  const double4 xs = xq[s];
  double ex = 0, ey = 0, ez = 0;
  const int q0 = (p / GSB + 1) * GSB;  // first position of the next block
  if (mua[s].w != 0.0)
    for (int q = q0 + lane; q < n; q += 32) {
      const int j = order ? order[q] : q;
      const double4 mj = ld4_cg(mua + j);
      if (C.on()) cached_field_pair(C, p, q, mj.x, mj.y, mj.z, ex, ey, ez);
      else {
        const double4 xj = ld4(xq + j);
        double dx, dy, dz;
        pair_del(P.box, perm, s, j, xs, xj, dx, dy, dz);
        induced_field_pair(P.pc, dx, dy, dz, dx * dx + dy * dy + dz * dz, mj.x, mj.y, mj.z, ex, ey, ez);
      }
    }
  ex = warp_sum(ex);
  ey = warp_sum(ey);
  ez = warp_sum(ez);
  if (lane == 0) {
    const double4 e = ef[s];
    R[p] = make_double4(e.x + ex, e.y + ey, e.z + ez, 0.0);
  }
}

// One launch per block instead of two: step b = "subtract block b-1's contribution from every row from block b on" fused
// with "solve block b".  CTA 0 (32 warps) first brings the 32 rows of block b up to date (warp w owns row w), then forms the
// block's pair tensors (thread (w, v) owns pair (w, v)) and substitutes; the other CTAs update the rows after block b.  The
// rows after block b are not needed before the next launch, so nothing waits inside the kernel.  Same operands per atom as
// the sequential sweep.  The launches of a sweep are replayed from a CUDA graph (engine.cu): at 750 atoms a sweep is
// 25 dependent kernels of a few microseconds each.
constexpr int GSS_THREADS = GSB * 32;
__global__ void __launch_bounds__(GSS_THREADS)
k_gsb_step(int n, int b, const int *__restrict__ order, DevParams P, const int *__restrict__ perm,
           const double4 *__restrict__ xq, double4 *__restrict__ mua, double4 *__restrict__ R,
           double *__restrict__ change_out, const int *stop = nullptr, GsPairCache C = GsPairCache{nullptr, 0})
{
  if (scf_stopped(stop)) return;
  __shared__ double sT[GSB][GSB][5];  // s1, s2, dx, dy, dz of every pair of the block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p0 = b * GSB, cnt = min(GSB, n - p0);
  // 1. rows from block b on: R[p] -= sum over the atoms of block b-1 of T mu_new (block b-1 is complete)
  if (b > 0) {
    const int p = blockIdx.x == 0 ? p0 + warp : p0 + GSB + (blockIdx.x - 1) * GSB + warp;
    if (p < n && (blockIdx.x > 0 || warp < cnt)) {
      const int s = order ? order[p] : p;
      const int q = (b - 1) * GSB + lane;
      const int j = order ? order[q] : q;
      double ex = 0, ey = 0, ez = 0;
      if (ld4_cg(mua + s).w != 0.0) {
        const double4 mj = ld4_cg(mua + j);
        if (C.on()) cached_field_pair(C, p, q, mj.x, mj.y, mj.z, ex, ey, ez);
        else {
          const double4 xs = xq[s], xj = ld4(xq + j);
          double dx, dy, dz;
          pair_del(P.box, perm, s, j, xs, xj, dx, dy, dz);
          induced_field_pair(P.pc, dx, dy, dz, dx * dx + dy * dy + dz * dz, mj.x, mj.y, mj.z, ex, ey, ez);
        }
      }
      ex = warp_sum(ex);
      ey = warp_sum(ey);
      ez = warp_sum(ez);
      if (lane == 0) {
        const double4 r = ld4_cg(R + p);
        R[p] = make_double4(r.x + ex, r.y + ey, r.z + ez, 0.0);
      }
    }
  }
  if (blockIdx.x != 0) return;
  // 2. CTA 0: the diagonal block
  if (warp < cnt && lane < cnt && warp != lane) {
    double dx, dy, dz, s1, s2;
    if (C.on()) C.load(p0 + warp, p0 + lane, s1, s2, dx, dy, dz);
    else {
      const int s = order ? order[p0 + warp] : p0 + warp, j = order ? order[p0 + lane] : p0 + lane;
      const double4 xs = xq[s], xj = xq[j];
      pair_del(P.box, perm, s, j, xs, xj, dx, dy, dz);
      induced_field_scalars(P.pc, dx * dx + dy * dy + dz * dz, s1, s2);
    }
    sT[warp][lane][0] = s1; sT[warp][lane][1] = s2;
    sT[warp][lane][2] = dx; sT[warp][lane][3] = dy; sT[warp][lane][4] = dz;
  }
  __syncthreads();  // (also orders the R updates of step 1 before the reads below)
  if (warp != 0) return;
  // Forward substitution in COLUMN form: lane v keeps the running field of block atom v; when atom w has its new dipole
  // it is broadcast and every lane subtracts T(v,w) mu_w from its own field -- no warp reduction on the critical path (a
  // reduction per atom made this loop ten times longer than everything else in the kernel).  T(v,w) is read as sT[w][v]
  // (T is even in del; consecutive lanes, 2-way bank conflicts instead of 32-way).
  const int sv = lane < cnt ? (order ? order[p0 + lane] : p0 + lane) : 0;
  double4 mv = lane < cnt ? ld4_cg(mua + sv) : make_double4(0, 0, 0, 0);
  const double4 rv = lane < cnt ? ld4_cg(R + p0 + lane) : make_double4(0, 0, 0, 0);
  double ax = rv.x, ay = rv.y, az = rv.z;
  auto subtract = [&](int w, double mx, double my, double mz) {  // field of lane's atom -= T(lane, w) (mx, my, mz)
    const double s1 = sT[w][lane][0], s2 = sT[w][lane][1];
    const double dx = sT[w][lane][2], dy = sT[w][lane][3], dz = sT[w][lane][4];
    const double t = s2 * (dx * mx + dy * my + dz * mz);
    ax -= fma(t, dx, s1 * mx);
    ay -= fma(t, dy, s1 * my);
    az -= fma(t, dz, s1 * mz);
  };
  // old dipoles of the atoms AFTER each lane's atom (independent steps: they pipeline)
  for (int u = 1; u < cnt; u++) {
    const double mx = __shfl_sync(FULL, mv.x, u), my = __shfl_sync(FULL, mv.y, u), mz = __shfl_sync(FULL, mv.z, u);
    if (lane < u) subtract(u, mx, my, mz);
  }
  double change = 0.0;
  for (int w = 0; w < cnt; w++) {
    double nx = 0, ny = 0, nz = 0;
    if (lane == w) {
      // alpha == 0: the field sum is skipped by the sequential kernel too; the product is zero either way
      nx = mv.w * ax, ny = mv.w * ay, nz = mv.w * az;
      change += (nx - mv.x) * (nx - mv.x) + (ny - mv.y) * (ny - mv.y) + (nz - mv.z) * (nz - mv.z);
      mv = make_double4(nx, ny, nz, mv.w);
    }
    nx = __shfl_sync(FULL, nx, w);
    ny = __shfl_sync(FULL, ny, w);
    nz = __shfl_sync(FULL, nz, w);
    if (lane > w && lane < cnt) subtract(w, nx, ny, nz);
  }
  if (lane < cnt) mua[sv] = mv;
  change = warp_sum(change);
  if (lane == 0) change_out[0] = b == 0 ? change : change_out[0] + change;
}

// The whole lower-triangular part of a sweep in ONE launch: a thread-block cluster walks the blocks.  The chain of
// k_gsb_step launches costs ~10 us per block (25 dependent launches at 750 atoms = 0.24 ms per sweep, the largest item of
// the reference's shipped examples); what a step really needs is ~2 us of dependent work.  Per block b:
//   CTA 0   : rows of block b  -=  T mu_new of block b-1 (warp w owns row w), tensors of the diagonal block into shared
//             memory, one warp substitutes (as k_gsb_step), new dipoles to global memory;
//   helpers : (CTAs 1..) rows of the blocks AFTER b  -=  T mu_new of block b-1 -- at the same time as CTA 0 solves block b;
//   barrier.cluster (release / acquire): block b's dipoles and the helpers' row updates are visible to the next step.
// The rows of block b+1 have then received every block before b from the helpers and get block b from CTA 0 itself.
// Same operands per atom and the same order of additions inside a row as the chain of k_gsb_step launches.
__device__ __forceinline__ void cluster_sync_all()
{
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_cta_rank()
{
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_num_ctas()
{
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

// dynamic shared memory of k_gsb_cluster: two diagonal tiles, two off-diagonal tiles, the block's right-hand sides and its
// new dipoles
constexpr int GSC_TILE = GSB * GSB * 5;                                   // doubles per off-diagonal tile: {s1, s2, del}
constexpr int GSC_DTILE = GSB * GSB * 6;                                  // per diagonal tile: the symmetric 3x3 tensor of a pair
constexpr int GSC_SMEM = (2 * GSC_DTILE + 2 * GSC_TILE + 2 * GSB * 3) * (int)sizeof(double);

__global__ void __launch_bounds__(GSS_THREADS)
k_gsb_cluster(int n, const int *__restrict__ order, DevParams P, const int *__restrict__ perm,
              const double4 *__restrict__ xq, double4 *__restrict__ mua, double4 *__restrict__ R,
              double *__restrict__ change_out, const int *stop, GsPairCache C)
{
  if (scf_stopped(stop)) return;   // the flag is stable while this kernel runs: every CTA takes the same branch
  extern __shared__ double gsc[];
  // CTA 0 keeps everything the NEXT block needs that does not depend on dipoles in shared memory, fetched by its 31 idle
  // warps while warp 0 substitutes: the tensors of the next diagonal block (sD) and of the pairs (row of the next block,
  // atom of this block) (sO).  After the barrier a step is then: read 32 right-hand sides, one shared-memory pass, substitute.
  double *sD = gsc, *sO = gsc + 2 * GSC_DTILE, *sR = sO + 2 * GSC_TILE, *sMu = sR + GSB * 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rank = (int)cluster_cta_rank(), helpers = (int)cluster_num_ctas() - 1;
  const int nblk = (n + GSB - 1) / GSB;
  double total = 0.0;   // lane 0 of CTA 0, warp 0: squared change of the sweep, block sums added in block order (as k_gsb_step)
  // {s1, s2, del} of the pair (position p, position q) of the sweep order
  auto tensor = [&](int p, int q, double *t) {
    double dx = 0, dy = 0, dz = 0, s1 = 0, s2 = 0;
    if (p < n && q < n && p != q) {
      if (C.on()) C.load(p, q, s1, s2, dx, dy, dz);
      else {
        const int s = order ? order[p] : p, j = order ? order[q] : q;
        pair_del(P.box, perm, s, j, xq[s], xq[j], dx, dy, dz);
        induced_field_scalars(P.pc, dx * dx + dy * dy + dz * dz, s1, s2);
      }
    }
    t[0] = s1; t[1] = s2; t[2] = dx; t[3] = dy; t[4] = dz;
  };
  // tiles of block b into buffer (b & 1): diagonal (b, b) and off-diagonal (rows of b, atoms of b - 1); row r by warp r0 + r
  auto prefetch = [&](int b, int w0, int nw) {
    const int p0 = b * GSB;
    double *d = sD + (b & 1) * GSC_DTILE, *o = sO + (b & 1) * GSC_TILE;
    for (int r = warp - w0; r < GSB; r += nw) {
      if (r < 0) continue;
      // diagonal tile: T = s1 I + s2 del del^T as six numbers, so that a substitution step is three short FMA chains
      // (the substitution is one dependent chain of 32 steps: two FP64 latencies less per step)
      double t[5];
      tensor(p0 + r, p0 + lane, t);
      double *m = d + (r * GSB + lane) * 6;
      m[0] = fma(t[1] * t[2], t[2], t[0]); m[1] = fma(t[1] * t[3], t[3], t[0]); m[2] = fma(t[1] * t[4], t[4], t[0]);
      m[3] = t[1] * t[2] * t[3]; m[4] = t[1] * t[2] * t[4]; m[5] = t[1] * t[3] * t[4];
      if (b > 0) tensor(p0 + r, p0 - GSB + lane, o + (r * GSB + lane) * 5);
    }
  };
  // helpers: R[p] -= sum over the atoms of block c of T(p, q) mu_new[q]  (global memory, coherent loads)
  auto apply = [&](int p, int c) {
    const int s = order ? order[p] : p;
    const int q = c * GSB + lane;   // block c is complete, hence full
    const int j = order ? order[q] : q;
    double ex = 0, ey = 0, ez = 0;
    if (ld4_cg(mua + s).w != 0.0) {
      const double4 mj = ld4_cg(mua + j);
      if (C.on()) cached_field_pair(C, p, q, mj.x, mj.y, mj.z, ex, ey, ez);
      else {
        const double4 xs = xq[s], xj = ld4(xq + j);
        double dx, dy, dz;
        pair_del(P.box, perm, s, j, xs, xj, dx, dy, dz);
        induced_field_pair(P.pc, dx, dy, dz, dx * dx + dy * dy + dz * dz, mj.x, mj.y, mj.z, ex, ey, ez);
      }
    }
    ex = warp_sum(ex);
    ey = warp_sum(ey);
    ez = warp_sum(ez);
    if (lane == 0) {
      const double4 r = ld4_cg(R + p);
      R[p] = make_double4(r.x + ex, r.y + ey, r.z + ez, 0.0);
    }
  };
  if (rank == 0) {
    prefetch(0, 0, GSB);
    __syncthreads();
  }
  for (int b = 0; b < nblk; b++) {
    const int p0 = b * GSB, cnt = min(GSB, n - p0);
    if (rank != 0) {
      // the rows after block b get block b-1's new dipoles while CTA 0 works on block b
      if (b > 0)
        for (int p = p0 + GSB + (rank - 1) * GSB + warp; p < n; p += helpers * GSB) apply(p, b - 1);
    } else {
      // (warp 0: the block's old dipoles travel while the right-hand sides are formed)
      const int sv = (warp == 0 && lane < cnt) ? (order ? order[p0 + lane] : p0 + lane) : 0;
      double4 mv = (warp == 0 && lane < cnt) ? ld4_cg(mua + sv) : make_double4(0, 0, 0, 0);
      // 1. right-hand sides of the block's rows: what the helpers and k_gsb_upper left in R, minus block b-1 (shared memory)
      if (warp < cnt) {
        const int s = order ? order[p0 + warp] : p0 + warp;
        const double4 r = ld4_cg(R + p0 + warp);
        double ex = 0, ey = 0, ez = 0;
        if (b > 0 && ld4_cg(mua + s).w != 0.0) {
          const double *t = sO + (b & 1) * GSC_TILE + (warp * GSB + lane) * 5;
          const double mx = sMu[lane * 3], my = sMu[lane * 3 + 1], mz = sMu[lane * 3 + 2];
          const double dm = t[2] * mx + t[3] * my + t[4] * mz;
          ex -= t[0] * mx + t[1] * dm * t[2];
          ey -= t[0] * my + t[1] * dm * t[3];
          ez -= t[0] * mz + t[1] * dm * t[4];
        }
        ex = warp_sum(ex);
        ey = warp_sum(ey);
        ez = warp_sum(ez);
        if (lane == 0) {
          sR[warp * 3] = r.x + ex;
          sR[warp * 3 + 1] = r.y + ey;
          sR[warp * 3 + 2] = r.z + ez;
        }
      }
      __syncthreads();
      if (warp == 0) {
        // 2. forward substitution in column form (see k_gsb_step)
        const double *sT = sD + (b & 1) * GSC_DTILE;
        double ax = lane < cnt ? sR[lane * 3] : 0.0, ay = lane < cnt ? sR[lane * 3 + 1] : 0.0, az = lane < cnt ? sR[lane * 3 + 2] : 0.0;
        auto subtract = [&](int w, double mx, double my, double mz) {   // T(lane, w) read as sT[w][lane] (T is symmetric in the pair)
          const double *t = sT + (w * GSB + lane) * 6;   // xx yy zz xy xz yz
          ax -= fma(t[4], mz, fma(t[3], my, t[0] * mx));
          ay -= fma(t[5], mz, fma(t[1], my, t[3] * mx));
          az -= fma(t[2], mz, fma(t[5], my, t[4] * mx));
        };
        for (int u = 1; u < cnt; u++) {
          const double mx = __shfl_sync(FULL, mv.x, u), my = __shfl_sync(FULL, mv.y, u), mz = __shfl_sync(FULL, mv.z, u);
          if (lane < u) subtract(u, mx, my, mz);
        }
        double change = 0.0;
        for (int w = 0; w < cnt; w++) {
          double nx = 0, ny = 0, nz = 0;
          if (lane == w) {
            nx = mv.w * ax, ny = mv.w * ay, nz = mv.w * az;
            change += (nx - mv.x) * (nx - mv.x) + (ny - mv.y) * (ny - mv.y) + (nz - mv.z) * (nz - mv.z);
            mv = make_double4(nx, ny, nz, mv.w);
          }
          nx = __shfl_sync(FULL, nx, w);
          ny = __shfl_sync(FULL, ny, w);
          nz = __shfl_sync(FULL, nz, w);
          if (lane > w && lane < cnt) subtract(w, nx, ny, nz);
        }
        if (lane < cnt) {
          mua[sv] = mv;
          sMu[lane * 3] = mv.x; sMu[lane * 3 + 1] = mv.y; sMu[lane * 3 + 2] = mv.z;
        }
        change = warp_sum(change);
        total = b == 0 ? change : total + change;
      } else if (b + 1 < nblk) {
        prefetch(b + 1, 1, GSB - 1);   // 3. meanwhile: the tiles of the next block into the other buffer
      }
      __syncthreads();
    }
    cluster_sync_all();
  }
  if (rank == 0 && warp == 0 && lane == 0) change_out[0] = total;
}

}  // namespace polb200
