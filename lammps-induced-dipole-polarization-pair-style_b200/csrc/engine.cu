// engine.cu -- device-side state of one pair-style instance and the orchestration of the five
// stages for one compute() call, plus the C ABI of include/polb200.h.
//
// Streams/graphs: everything of one call is enqueued on one stream; the host synchronises only
// where the reference's control flow needs a device value (ghost count at a rebuild, the convergence
// test of `precision` mode).  fixed_iteration mode enqueues all sweeps back to back with no sync.
// Also in this translation unit: the multi-GPU layer (comm.cuh, included after the handle definition) and the
// reciprocal-space Ewald (ewald.cuh + polb200_ewald_* at the end of the C ABI block).
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "comm_types.h"
#include "devbuf.h"
#include "host_style.h"
#include "kernels.cuh"
#include "polb200.h"

namespace polb200 {

// multi-GPU state (see comm.cuh)
struct CommState {
  bool active = false, geom_valid = false, want_push = true, flags_zeroed = false;
  bool shared_device = false;                          // two ranks on one GPU: spin barriers are not safe there
  unsigned long long barrier_timeout_ns = 20000000000ull;  // 20 s
  int rank = 0, nranks = 1;
  int pg[3] = {1, 1, 1};
  ncclComm_t nccl = nullptr;
  DecompPlan plan{};
  SendGeom geom{};
  int send_cnt[NDIR] = {}, recv_cnt[NDIR] = {}, send_off[NDIR] = {}, recv_off[NDIR] = {};
  int nsend = 0, nrecv = 0;
  long nglobal = 0;
  int nloc_of[64] = {};
  DBuf<int> send_owner_u, send_owner, send_dir, slot_of_u, dir_start, gslot, counts_dev, dir_of_u;
  DBuf<double4 *> push_ptrx;
  DBuf<unsigned long long> flags;
  DBuf<double4> sbuf, rbuf;
  DBuf<int4> sbufi, rbufi;
  DBuf<char> ipc_dev;
  void *peer_ptr[NPEERBUF][MAX_PEERS] = {};
  void *mapped_ptr[NPEERBUF] = {};
  PeerPush push{};
  unsigned long long epoch = 0;
  std::vector<void *> graveyard;  // dipole arrays replaced while peers may still map them
};

}  // namespace polb200

using namespace polb200;

struct polb200_handle {
  CommState comm;
  HostStyle style;
  std::string err;
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6] = {};
  long launches = 0;
  size_t partial_off = 0;        // where the list sweeps write their per-row squared changes

  Box box{};
  bool box_set = false;

  // options
  int sweep_block = BLOCK;
  double bin_div = 4.0;          // neighbor cutoff / cell width
  int xsort_bits = 10;           // resolution of the x position inside a cell in the sort key (0: cell order only)
  // 41 (default) / 40 / 44: TMA-fed pair-group sweep (Jacobi list mode); 30 / 31: pair groups with register prefetch;
  // 20 / 21: per-atom rows + radial cache (also the ranked colouring sweep); 6: matrix-free; 0: first version
  int sweep_variant = 41;
  bool use_tight = true;         // per-step tight list
  int gpf_minb = 4;              // resident CTAs per SM asked of the grouped force kernel (4: 128 registers, no spills; measured faster than 5)
  bool use_group_pairs = true;   // LJ + Coulomb + field and polarization forces on the pair-group rows when they qualify
  // all-pairs (exact) mode on several GPUs (polb200_comm_init_replicated): every process holds the whole system, the O(N^2)
  // stages are shared by rows and all-gathered (SURVEY 8e caveat)
  struct ExactShare {
    ncclComm_t nccl = nullptr;
    int rank = 0, nranks = 1;
    bool active = false;
  } xr;
  DBuf<double> xg;               // all-gather scratch
  int gs_cluster = 16;           // exact-mode blocked Gauss-Seidel: CTAs of the cluster that walks the blocks in one launch (0: one launch per block)
  bool gs_blocked = true;        // exact-mode Gauss-Seidel as blocked forward substitution (false: one atom at a time)
  // Owned atoms may lie up to this far OUTSIDE the box / the brick they are handed to (a caller that keeps rigid bodies
  // whole assigns a molecule to the brick of one of its atoms): the ghost shells are made that much deeper.
  double atom_slack = 0.0;
  bool use_graphs = true;        // ... its per-block launches replayed from a CUDA graph
  long params_version = 0;       // bumped whenever the kernel parameters (DevParams) change
  struct SweepGraph {
    cudaGraphExec_t exec = nullptr;
    int n = 0, nodes = 0;
    long version = -1;
    const void *key[8] = {};
  } gs_graph;
  DBuf<double> gs_planes;        // per-step pair-tensor cache of the blocked sweep (small systems)
  int gs_cache_max = 3000;       // ... up to this many atoms (5 n^2 doubles)
  DBuf<double4> gsR;
  bool l2_evict_first = true;    // TMA row streams are marked evict-first in L2
  bool alternate = true;         // sweeps walk the groups alternately forwards / backwards (L2 reuse of the stream tail)
  unsigned sweep_parity = 0;
  bool use_push = true;          // sweep kernel stores new dipoles into their ghost copies itself
  bool time_sweeps = false;     // record CUDA events around every k_sweep launch (bench roofline)
  std::vector<cudaEvent_t> sweep_ev;
  size_t sweep_ev_used = 0;
  double sweep_ms_accum = 0.0;
  long sweep_launches = 0;

  // device copies of host-style tables
  DBuf<double> d_coeff;   // 7 tables x (ntypes+1)^2 + cutneighsq
  DBuf<double> d_tables;  // 8 x ntable
  DevParams P{};
  bool params_uploaded = false;

  // caller-order staging (device)
  DBuf<double> c_x, c_q, c_alpha, c_mu, c_f, c_ef, c_xhold;
  DBuf<int> c_type, c_mol, c_tag, c_nspecial, c_special, c_mask;
  DBuf<int2> exb;                // exclusion-rule membership bits of owned atoms and ghosts
  ExclRules excl{};  // neigh_modify exclude rules (polb200_set_exclusions)
  // pinned host staging
  HPinned<double> h_stage;
  HPinned<double> h_scal;
  HPinned<int> h_int;

  // cell-sorted ext arrays
  DBuf<double4> xq, mua, mub, ef, f_pair, f_pol;
  DBuf<int2> tm;
  DBuf<int> tag, perm, invperm, keys, keys2, vals, vals2;
  DBuf<int> g_owner_u, g_shift_u, g_owner, g_shift;
  DBuf<int> cl_start, cg_start, stencil;
  DBuf<int> s_nspecial, s_special;
  DBuf<unsigned long long> cnt, rowstart;
  DBuf<int> neigh, tneigh, tcount;
  // fused sweep + ghost update: per owned atom the addresses of its ghost copies in mua / mub
  DBuf<unsigned long long> push_off;
  DBuf<double4 *> push_ptr0, push_ptr1;
  bool push_ready = false;
  // pair groups (two cell-row neighbours per warp) for the Jacobi list sweep
  DBuf<int> group_first, group_two, gneigh, gcount, tgneigh, tgcount;
  DBuf<unsigned long long> growstart;
  DBuf<double4> s12ab;           // per-step radial cache of the group rows: {s1a, s2a, s1b, s2b} per entry
  DBuf<unsigned long long> gcstart;  // chunk-record layout of the same (TMA sweep): first record of every group row
  DBuf<unsigned char> gcrec;
  unsigned long long gchunks = 0;
  int ngroups = 0;
  unsigned long long gpairs = 0;
  bool groups_built = false, group_cache_valid = false;
  // group-coloured Gauss-Seidel sweep (list mode polar_gs / polar_gs_ranked): colouring of the pair groups
  DBuf<int> gof, gadj, gadjn, gcolour, glist, cstart_dev, col_export;
  DBuf<float> gmetric;
  int ncolours = 8;              // colours = chunks of one sweep
  double gs_strong_m = 6.0;      // groups closer than the radius that holds this many atoms on average get different colours
  int chunk_beg[33] = {};
  int colour_rounds = 0;
  bool colours_valid = false;
  // device-side convergence test: ctl = {stop, iterations, diverged, ticket}
  DBuf<int> ctl;
  HPinned<int> h_ctl;
  DBuf<double> slice;
  cudaEvent_t ev_it[2] = {};
  bool rmin_fused = false;       // this step's group cache also produced rmin
  const int *scf_stop = nullptr; // non-null while the iterations of a precision-mode solve are being enqueued
  int scf_lag = 1;               // iterations enqueued ahead of the host's look at the stop flag (0: test every iteration)
  DBuf<double2> s12;             // per-step radial cache aligned with the tight list
  bool s12_valid = false;
  DBuf<char> cub_tmp;
  DBuf<double> partial, scal;
  DBuf<double> ea_row, va_pair_row, va_pol_row, c_eatom, c_vatom;  // per-atom tallies (sorted order / caller order)
  DBuf<int> flags;
  DBuf<double> metric, metric2;
  DBuf<int> ranked, ranked_in;
  DBuf<unsigned long long> rmin_bits;

  int nloc = 0, nghost = 0, maxspecial = 0, nstencil = 0;
  bool molecular = false;
  unsigned long long npairs = 0;
  bool have_lists = false;
  int ago_internal = 0;  // library-side Neighbor::decide state (ago < 0 calls)
  bool mu_is_b = false;
};

namespace polb200 {

template <class... KA, class... A>
static void launch_kernel(polb200_handle *h, void (*kernel)(KA...), int grid, int block, A &&...args)
{
  if (grid <= 0) return;  // empty range (e.g. a brick without ghosts)
  kernel<<<grid, block, 0, h->stream>>>(std::forward<A>(args)...);
  h->launches++;
  CUDA_CHECK(cudaGetLastError());
}
#define LAUNCH(h, kernel, grid, block, ...) launch_kernel(h, kernel, grid, block, __VA_ARGS__)

static void upload_params(polb200_handle *h)
{
  const HostStyle &st = h->style;
  const int n1 = st.ntypes + 1, nn = n1 * n1;
  std::vector<double> buf((size_t)8 * nn);
  const std::vector<double> *src[8] = {&st.cutsq, &st.cut_ljsq, &st.lj1,    &st.lj2,
                                       &st.lj3,   &st.lj4,      &st.offset, &st.cutneighsq};
  for (int k = 0; k < 8; k++) std::copy(src[k]->begin(), src[k]->end(), buf.begin() + (size_t)k * nn);
  h->d_coeff.ensure(buf.size());
  CUDA_CHECK(cudaMemcpy(h->d_coeff.p, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice));
  DevParams &P = h->P;
  P.lj.cutsq = h->d_coeff.p;
  P.lj.cut_ljsq = h->d_coeff.p + nn;
  P.lj.lj1 = h->d_coeff.p + 2 * nn;
  P.lj.lj2 = h->d_coeff.p + 3 * nn;
  P.lj.lj3 = h->d_coeff.p + 4 * nn;
  P.lj.lj4 = h->d_coeff.p + 5 * nn;
  P.lj.offset = h->d_coeff.p + 6 * nn;
  P.cutneighsq = h->d_coeff.p + 7 * nn;

  const int nt = st.ncoultablebits ? (1 << st.ncoultablebits) : 1;
  std::vector<double> tb((size_t)8 * nt, 0.0);
  if (st.ncoultablebits) {  // interleave: {r, dr, f, df, e, de, c, dc} per entry
    const std::vector<double> *ts[8] = {&st.tab.r, &st.tab.dr, &st.tab.f, &st.tab.df,
                                        &st.tab.e, &st.tab.de, &st.tab.c, &st.tab.dc};
    for (int i = 0; i < nt; i++)
      for (int k = 0; k < 8; k++) tb[(size_t)8 * i + k] = (*ts[k])[i];
  }
  h->d_tables.ensure(tb.size() + 8);
  CUDA_CHECK(cudaMemcpy(h->d_tables.p, tb.data(), tb.size() * sizeof(double), cudaMemcpyHostToDevice));
  P.tb.rec = reinterpret_cast<const CoulTabHalf *>(h->d_tables.p);

  PairConsts &pc = P.pc;
  pc.cut_coulsq = st.cut_coulsq;
  pc.f_shift = -1.0 / (st.cut_coul * st.cut_coul);
  pc.kq = sqrt(st.env.qqrd2e);
  pc.qqrd2e = st.env.qqrd2e;
  pc.g_ewald = st.env.g_ewald;
  pc.polar_damp = st.polar_damp;
  pc.polar_cutsq = st.polar_cutoff > 0.0 ? st.polar_cutoff * st.polar_cutoff : -1.0;
  pc.tabinnersq = st.tab.tabinnersq;
  pc.damping_exponential = st.damping_type == DAMP_EXPONENTIAL;
  pc.ncoultablebits = st.ncoultablebits;
  pc.ncoulmask = st.tab.mask;
  pc.ncoulshiftbits = st.tab.shift;
  pc.ntypes = st.ntypes;
  pc.has_molecules = 1;  // refined at every rebuild
  for (int k = 0; k < 4; k++) {
    pc.special_lj[k] = st.env.special_lj[k];
    pc.special_coul[k] = st.env.special_coul[k];
  }
  h->params_uploaded = true;
  h->have_lists = false;
  h->params_version++;
}

static int bits_for(int n)
{
  int b = 1;
  while ((1l << b) < n) b++;
  return b;
}

// Fine cell grid over the box extended by the ghost cutoff.  Cells are cut/bin_div wide (LAMMPS uses
// cut/2, src/nbin_standard.cpp:93-99; a finer grid gives a tighter candidate set and, because atoms are
// stored in cell order with x fastest, longer contiguous neighbour runs).  The stencil is a list of
// (dy,dz) cell rows with the x half-extent that can still be within the cutoff (closest-approach
// distance between cells, cf. NStencil::bin_distance, src/nstencil.cpp:205-223).
static void setup_grid(polb200_handle *h)
{
  const HostStyle &st = h->style;
  Grid &g = h->P.grid;
  const double cut = st.cutneighmax;
  const double binsize = cut / h->bin_div;
  std::vector<int> stencil;
  int sx[3];
  double cs[3];
  long ncell = 1;
  for (int d = 0; d < 3; d++) {
    const double ext = h->box.periodic[d] ? cut + h->atom_slack : 0.0;
    const double lo = h->box.lo[d] - ext, hi = h->box.hi[d] + ext;
    // margin: owned atoms may sit up to skin/2 outside the box between rebuilds
    const double margin = 0.5 * st.env.skin + h->atom_slack + 1e-9 * (hi - lo);
    g.lo[d] = lo - margin;
    const double len = (hi + margin) - g.lo[d];
    int nc = (int)(len / binsize);
    if (nc < 1) nc = 1;
    if (nc > 1024) nc = 1024;
    g.nc[d] = nc;
    cs[d] = len / nc;
    g.inv[d] = 1.0 / cs[d];
    sx[d] = (int)ceil(cut / cs[d]);
    if (sx[d] > 15) throw StyleError{POLB200_ERR_UNSUPPORTED, "neighbor stencil too wide"};
    ncell *= nc;
  }
  if (ncell > (1l << 30)) throw StyleError{POLB200_ERR_UNSUPPORTED, "Too many neighbor bins"};
  g.ncell = (int)ncell;
  g.xbits = std::max(0, std::min(h->xsort_bits, 31 - bits_for(g.ncell + 1)));
  auto gap = [](int o, double c) { return o > 0 ? (o - 1) * c : (o < 0 ? (o + 1) * c : 0.0); };
  for (int k = -sx[2]; k <= sx[2]; k++)
    for (int j = -sx[1]; j <= sx[1]; j++) {
      const double dy = gap(j, cs[1]), dz = gap(k, cs[2]);
      int xext = -1;
      for (int i = 0; i <= sx[0]; i++) {
        const double dx = gap(i, cs[0]);
        if (dx * dx + dy * dy + dz * dz <= cut * cut) xext = i;
      }
      if (xext >= 0) stencil.push_back((j + 16) | ((k + 16) << 6) | (xext << 12));
    }
  h->nstencil = (int)stencil.size();
  h->stencil.ensure(stencil.size());
  CUDA_CHECK(cudaMemcpy(h->stencil.p, stencil.data(), stencil.size() * sizeof(int), cudaMemcpyHostToDevice));
}

static void sort_pairs(polb200_handle *h, int n, const int *kin, int *kout, const int *vin, int *vout, int endbit)
{
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, endbit, h->stream);
  h->cub_tmp.ensure(bytes);
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(h->cub_tmp.p, bytes, kin, kout, vin, vout, n, 0, endbit, h->stream));
  h->launches += 3;
}

static void exclusive_sum(polb200_handle *h, int n, const unsigned long long *in, unsigned long long *out)
{
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, h->stream);
  h->cub_tmp.ensure(bytes);
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(h->cub_tmp.p, bytes, in, out, n, h->stream));
  h->launches += 2;
}


// ---- host <-> device staging of the caller's arrays ----------------------------------------------------
template <class T>
static void stage_in(polb200_handle *h, DBuf<T> &dst, const T *src, size_t n, bool on_device)
{
  dst.ensure(n);
  if (!src || n == 0) return;
  CUDA_CHECK(cudaMemcpyAsync(dst.p, src, n * sizeof(T), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                             h->stream));
}

// grow the ext (owned + ghost) arrays to `next` records, preserving the owned part [0,n)
static void grow_ext(polb200_handle *h, int n, size_t next)
{
  auto grow = [&](auto &buf) {
    using T = typename std::remove_reference<decltype(*buf.p)>::type;
    if (buf.cap >= next) return;
    DBuf<T> nb;
    nb.ensure(next, h->comm.active ? 1.3 : 1.1);
    CUDA_CHECK(cudaMemcpyAsync(nb.p, buf.p, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    auto *gy = buf.graveyard;
    buf.release();
    buf = nb;
    buf.graveyard = gy;
  };
  grow(h->xq); grow(h->mua); grow(h->tm); grow(h->tag);
  h->mub.ensure(next, h->comm.active ? 1.3 : 1.1);
}

}  // namespace polb200
#include "comm.cuh"
namespace polb200 {

// ghost copies follow their owners: positions (once per step) and/or one dipole array (once per sweep)
static void ghost_update(polb200_handle *h, bool pos, double4 *mu, bool fence_before = false)
{
  if (h->comm.active) {
    comm_refresh(h, pos, mu, fence_before);
    return;
  }
  const int ng = h->nghost, n = h->nloc;
  if (!ng) return;
  const int *stop = h->scf_stop;
  if (pos && mu) LAUNCH(h, (k_ghost_refresh<true, true>), cdiv(ng, 256), 256, stop, ng, n, h->g_owner.p, h->g_shift.p, h->box, h->xq.p, mu);
  else if (pos) LAUNCH(h, (k_ghost_refresh<true, false>), cdiv(ng, 256), 256, stop, ng, n, h->g_owner.p, h->g_shift.p, h->box, h->xq.p, mu);
  else if (mu) LAUNCH(h, (k_ghost_refresh<false, true>), cdiv(ng, 256), 256, stop, ng, n, h->g_owner.p, h->g_shift.p, h->box, h->xq.p, mu);
}

static void rebuild(polb200_handle *h, const polb200_atoms *at)
{
  const HostStyle &st = h->style;
  const int n = at->nlocal;
  const bool dev = at->on_device != 0;
  setup_grid(h);
  const Grid &g = h->P.grid;

  // static per-atom attributes (positions and dipoles were staged by the caller of rebuild)
  stage_in(h, h->c_q, at->q, n, dev);
  stage_in(h, h->c_type, at->type, n, dev);
  stage_in(h, h->c_alpha, at->alpha, n, dev);
  if (at->molecule) stage_in(h, h->c_mol, at->molecule, n, dev);
  if (at->tag) stage_in(h, h->c_tag, at->tag, n, dev);
  h->molecular = at->nspecial && at->special && at->maxspecial > 0;
  h->maxspecial = h->molecular ? at->maxspecial : 0;

  // 1. cell-sort the owned atoms
  h->keys.ensure(n); h->keys2.ensure(n); h->vals.ensure(n); h->vals2.ensure(n);
  h->flags.ensure(8);
  CUDA_CHECK(cudaMemsetAsync(h->flags.p, 0, 8 * sizeof(int), h->stream));
  LAUNCH(h, k_local_keys, cdiv(n, 256), 256, n, h->c_x.p, g, h->keys.p, h->vals.p, h->flags.p);
  sort_pairs(h, n, h->keys.p, h->keys2.p, h->vals.p, h->vals2.p, bits_for(g.ncell + 1) + g.xbits);
  h->perm.ensure(n); h->invperm.ensure(n);
  CUDA_CHECK(cudaMemcpyAsync(h->perm.p, h->vals2.p, n * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
  h->cl_start.ensure(g.ncell + 2);
  LAUNCH(h, k_cell_starts, cdiv(g.ncell + 1, 256), 256, g.ncell, n, h->keys2.p, h->cl_start.p, g.xbits);

  // capacity guess for ext arrays: grown after the ghost count is known
  h->xq.ensure(n); h->mua.ensure(n); h->mub.ensure(n); h->tm.ensure(n); h->tag.ensure(n);
  LAUNCH(h, k_gather_local, cdiv(n, 256), 256, n, h->perm.p, h->c_x.p, h->c_q.p, h->c_type.p,
         at->molecule ? h->c_mol.p : nullptr, at->tag ? h->c_tag.p : nullptr, h->c_alpha.p, h->c_mu.p,
         h->xq.p, h->mua.p, h->tm.p, h->tag.p, h->invperm.p, h->flags.p + 3);

  // 2. ghosts: periodic images of this box (single GPU) or the boundary shells of the neighbour bricks
  int ng = 0;
  if (h->comm.active) {
    comm_build_ghosts(h, n);
    ng = h->nghost;
  } else {
    h->cnt.ensure(n + 1); h->rowstart.ensure(n + 2);
    LAUNCH(h, k_ghost_count, cdiv(n, 256), 256, n, h->xq.p, h->box, st.cutneighmax + h->atom_slack, h->cnt.p);
    CUDA_CHECK(cudaMemsetAsync(h->cnt.p + n, 0, sizeof(unsigned long long), h->stream));
    exclusive_sum(h, n + 1, h->cnt.p, h->rowstart.p);
    unsigned long long ng64 = 0;
    CUDA_CHECK(cudaMemcpyAsync(&ng64, h->rowstart.p + n, sizeof(ng64), cudaMemcpyDeviceToHost, h->stream));
    int hflags[4];
    CUDA_CHECK(cudaMemcpyAsync(hflags, h->flags.p, sizeof(hflags), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (hflags[0] & 1) throw StyleError{POLB200_ERR_NAN, "Non-numeric positions - simulation unstable"};
    h->P.pc.has_molecules = hflags[3] ? 1 : 0;
    if (ng64 + (unsigned long long)n >= (1ull << 30))
      throw StyleError{POLB200_ERR_OVERFLOW, "owned+ghost atoms exceed the 30-bit neighbor index"};
    ng = (int)ng64;
    h->nghost = ng;
    const size_t next = (size_t)n + ng;
    if (ng > 0) {
      h->g_owner_u.ensure(ng); h->g_shift_u.ensure(ng); h->g_owner.ensure(ng); h->g_shift.ensure(ng);
      h->keys.ensure(std::max(n, ng)); h->keys2.ensure(std::max(n, ng));
      h->vals.ensure(std::max(n, ng)); h->vals2.ensure(std::max(n, ng));
      LAUNCH(h, k_ghost_fill, cdiv(n, 256), 256, n, h->xq.p, h->box, st.cutneighmax + h->atom_slack, g, h->rowstart.p,
             h->g_owner_u.p, h->g_shift_u.p, h->keys.p, h->vals.p);
      sort_pairs(h, ng, h->keys.p, h->keys2.p, h->vals.p, h->vals2.p, bits_for(g.ncell + 1) + g.xbits);
    }
    h->push_off.ensure(n + 2);  // the ghost count scan is also the CSR of every owned atom's images
    CUDA_CHECK(cudaMemcpyAsync(h->push_off.p, h->rowstart.p, (size_t)(n + 1) * sizeof(unsigned long long),
                               cudaMemcpyDeviceToDevice, h->stream));
    grow_ext(h, n, next);
    h->push_ptr0.ensure(ng + 1); h->push_ptr1.ensure(ng + 1);
    if (ng > 0) {
      LAUNCH(h, k_invert_perm, cdiv(ng, 256), 256, ng, h->vals2.p, h->keys.p);  // keys is free again: sorted position of u
      LAUNCH(h, k_push_tables_local, cdiv(ng, 256), 256, ng, n, h->keys.p, h->mua.p, h->mub.p, h->push_ptr0.p, h->push_ptr1.p);
    }
    h->push_ready = true;
    h->cg_start.ensure(g.ncell + 2);
    if (ng > 0) {
      LAUNCH(h, k_ghost_gather, cdiv(ng, 256), 256, ng, n, h->vals2.p, h->g_owner_u.p, h->g_shift_u.p, h->box,
             h->xq.p, h->mua.p, h->tm.p, h->tag.p, h->g_owner.p, h->g_shift.p);
      LAUNCH(h, k_cell_starts, cdiv(g.ncell + 1, 256), 256, g.ncell, ng, h->keys2.p, h->cg_start.p, g.xbits);
    } else {
      CUDA_CHECK(cudaMemsetAsync(h->cg_start.p, 0, (g.ncell + 2) * sizeof(int), h->stream));
    }
  }

  if (h->excl.n > 0) {  // group-membership bits of the exclusion rules: owned atoms (cell-sorted order), then their ghosts
    bool need_mask = false;
    for (int r = 0; r < h->excl.n; r++) need_mask |= h->excl.kind[r] != EXCL_TYPE;
    if (need_mask && !at->mask) throw StyleError{POLB200_ERR_ARG, "neigh_modify exclude group/molecule rules need polb200_atoms.mask"};
    if (need_mask) stage_in(h, h->c_mask, at->mask, n, dev);
    h->exb.ensure((size_t)n + ng + 1);
    LAUNCH(h, k_exbits, cdiv(n, 256), 256, n, h->perm.p, need_mask ? h->c_mask.p : (const int *)nullptr, h->excl, h->exb.p);
    if (ng > 0) {
      if (h->comm.active) comm_ghost_exbits(h, n, ng);
      else LAUNCH(h, k_exbits_ghost, cdiv(ng, 256), 256, ng, n, h->g_owner.p, h->exb.p);
    }
  }

  // special-bond lists in sorted order (tags)
  const int *d_nspecial = nullptr, *d_special = nullptr;
  if (h->molecular) {
    stage_in(h, h->c_nspecial, at->nspecial, (size_t)3 * n, dev);
    stage_in(h, h->c_special, at->special, (size_t)n * at->maxspecial, dev);
    h->s_nspecial.ensure((size_t)3 * n);
    h->s_special.ensure((size_t)n * at->maxspecial);
    LAUNCH(h, k_gather_rows, cdiv((long)n * 3, 256), 256, n, 3, h->perm.p, h->c_nspecial.p, h->s_nspecial.p);
    LAUNCH(h, k_gather_rows, cdiv((long)n * at->maxspecial, 256), 256, n, at->maxspecial, h->perm.p,
           h->c_special.p, h->s_special.p);
    d_nspecial = h->s_nspecial.p;
    d_special = h->s_special.p;
  }

  // 3. neighbor list: count, scan, fill
  LAUNCH(h, k_neigh_build<false>, cdiv(n, WARPS_PER_BLOCK), BLOCK, n, h->P, h->xq.p, h->tm.p, h->tag.p,
         h->cl_start.p, h->cg_start.p, h->nstencil, h->stencil.p, d_nspecial, d_special, h->maxspecial,
         h->cnt.p, (const unsigned long long *)nullptr, (int *)nullptr);
  CUDA_CHECK(cudaMemsetAsync(h->cnt.p + n, 0, sizeof(unsigned long long), h->stream));
  exclusive_sum(h, n + 1, h->cnt.p, h->rowstart.p);
  unsigned long long np = 0;
  CUDA_CHECK(cudaMemcpyAsync(&np, h->rowstart.p + n, sizeof(np), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->npairs = np;
  h->neigh.ensure(np + 32);
  LAUNCH(h, k_neigh_build<true>, cdiv(n, WARPS_PER_BLOCK), BLOCK, n, h->P, h->xq.p, h->tm.p, h->tag.p,
         h->cl_start.p, h->cg_start.p, h->nstencil, h->stencil.p, d_nspecial, d_special, h->maxspecial,
         h->cnt.p, h->rowstart.p, h->neigh.p);

  // 4. pair groups for the Jacobi list sweep (union skin list of two cell-row neighbours per warp)
  h->groups_built = false;
  h->colours_valid = false;
  const bool gs_mode = st.polar_gs || st.polar_gs_ranked;
  // Jacobi sweeps: any pair-group variant; Gauss-Seidel: the group-coloured sweep runs on the TMA kernel only, and an
  // explicit gs_chunks setting asks for the per-atom chunks the oracle emulates
  if (h->sweep_variant >= 30 && st.polar_cutoff > 0.0 && !st.zodid &&
      (!gs_mode || (h->sweep_variant >= 40 && st.gs_chunks == 0))) {
    const int nrows = g.nc[1] * g.nc[2];
    h->cnt.ensure(std::max(n, nrows) + 1); h->growstart.ensure((size_t)n + nrows + 2);
    LAUNCH(h, k_group_count, cdiv(nrows, 256), 256, nrows, g.nc[0], h->cl_start.p, h->cnt.p);
    CUDA_CHECK(cudaMemsetAsync(h->cnt.p + nrows, 0, sizeof(unsigned long long), h->stream));
    exclusive_sum(h, nrows + 1, h->cnt.p, h->growstart.p);
    unsigned long long ngr = 0;
    CUDA_CHECK(cudaMemcpyAsync(&ngr, h->growstart.p + nrows, sizeof(ngr), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    h->ngroups = (int)ngr;
    h->group_first.ensure(ngr + 1); h->group_two.ensure(ngr + 1);
    LAUNCH(h, k_group_fill, cdiv(nrows, 256), 256, nrows, g.nc[0], h->cl_start.p, h->growstart.p, h->group_first.p, h->group_two.p);
    const int ngb = cdiv(h->ngroups, WARPS_PER_BLOCK);
    h->cnt.ensure(ngr + 1);
    LAUNCH(h, k_group_build<false>, ngb, BLOCK, h->ngroups, n, h->P, h->xq.p, h->tm.p, h->group_first.p, h->group_two.p,
           h->cl_start.p, h->cg_start.p, h->nstencil, h->stencil.p, h->cnt.p, (const unsigned long long *)nullptr, (int *)nullptr, (int *)nullptr);
    CUDA_CHECK(cudaMemsetAsync(h->cnt.p + ngr, 0, sizeof(unsigned long long), h->stream));
    exclusive_sum(h, (int)ngr + 1, h->cnt.p, h->growstart.p);
    unsigned long long gp = 0;
    CUDA_CHECK(cudaMemcpyAsync(&gp, h->growstart.p + ngr, sizeof(gp), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    h->gpairs = gp;
    h->gneigh.ensure(gp + 32);
    h->gcount.ensure(ngr + 1);
    LAUNCH(h, k_group_build<true>, ngb, BLOCK, h->ngroups, n, h->P, h->xq.p, h->tm.p, h->group_first.p, h->group_two.p,
           h->cl_start.p, h->cg_start.p, h->nstencil, h->stencil.p, h->cnt.p, h->growstart.p, h->gneigh.p, h->gcount.p);
    // chunk-record layout for the TMA sweep
    h->cnt.ensure(ngr + 1); h->gcstart.ensure(ngr + 2);
    LAUNCH(h, k_group_chunk_count, cdiv(h->ngroups, 256), 256, h->ngroups, h->gcount.p, h->cnt.p);
    CUDA_CHECK(cudaMemsetAsync(h->cnt.p + ngr, 0, sizeof(unsigned long long), h->stream));
    exclusive_sum(h, (int)ngr + 1, h->cnt.p, h->gcstart.p);
    unsigned long long nch = 0;
    CUDA_CHECK(cudaMemcpyAsync(&nch, h->gcstart.p + ngr, sizeof(nch), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    h->gchunks = nch;
    h->groups_built = true;
  }

  // remember positions for the displacement trigger (Neighbor::build: xhold, src/neighbor.cpp:2032-2044)
  h->c_xhold.ensure((size_t)3 * n);
  CUDA_CHECK(cudaMemcpyAsync(h->c_xhold.p, h->c_x.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToDevice,
                             h->stream));
  h->nloc = n;
  h->have_lists = true;
  h->ago_internal = 0;
}

// Neighbor::decide (src/neighbor.cpp:1923-1937) for callers that pass ago < 0
static bool decide_rebuild(polb200_handle *h, int n)
{
  if (!h->have_lists) return true;
  if (n != h->nloc) {
    // atoms only change bricks at a rebuild, and the bricks must take this decision together
    if (h->comm.active) throw StyleError{POLB200_ERR_STATE, "number of owned atoms changed between neighbor rebuilds"};
    return true;
  }
  const polb200_env &e = h->style.env;
  h->ago_internal++;
  const int every = e.neigh_every > 0 ? e.neigh_every : 1;
  if (h->ago_internal >= e.neigh_delay && h->ago_internal % every == 0) {
    if (!e.neigh_check) return true;
    const double trig = 0.25 * e.skin * e.skin;  // triggersq = (skin/2)^2
    CUDA_CHECK(cudaMemsetAsync(h->flags.p + 1, 0, sizeof(int), h->stream));
    LAUNCH(h, k_check_distance, cdiv(n, 256), 256, n, h->c_x.p, h->c_xhold.p, trig, h->flags.p + 1);
    if (h->comm.active) comm_allreduce(h, h->flags.p + 1, 1, ncclInt, ncclMax);  // any atom on any brick
    int flag = 0;
    CUDA_CHECK(cudaMemcpyAsync(&flag, h->flags.p + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return flag != 0;
  }
  return false;
}

static void sweep_event(polb200_handle *h)
{
  if (!h->time_sweeps) return;
  if (h->sweep_ev_used == h->sweep_ev.size()) {
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreate(&e));
    h->sweep_ev.push_back(e);
  }
  CUDA_CHECK(cudaEventRecord(h->sweep_ev[h->sweep_ev_used++], h->stream));
}

static void sweep_events_collect(polb200_handle *h)
{
  for (size_t k = 0; k + 1 < h->sweep_ev_used; k += 2) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->sweep_ev[k], h->sweep_ev[k + 1]);
    h->sweep_ms_accum += ms;
    h->sweep_launches++;
  }
  h->sweep_ev_used = 0;
}

// list-mode sweep over ranked positions [beg,end): picks the kernel variant.  Returns the number of
// partial sums written to h->partial (rows for the v2 kernels, blocks for the first version).
// push: the kernel also stores each new dipole into the ghost slots other bricks hold for it.
static PushArgs push_args(polb200_handle *h, const double4 *nxt)
{
  PushArgs Q{};
  Q.off = h->push_off.p;
  Q.ptr = nxt == h->mub.p ? h->push_ptr1.p : h->push_ptr0.p;
  return Q;
}

template <bool DAMP, int PF, int WPB, int MINB>
static int launch_v2(polb200_handle *h, int beg, int end, const int *order, const DevParams &P, ListRows L,
                     const double4 *cur, double4 *nxt, bool change, bool push)
{
  const int nb = cdiv(end - beg, WPB);
  const PushArgs Q = push ? push_args(h, nxt) : PushArgs{};
#define GO(CH, PU) \
  LAUNCH(h, (k_sweep_list2<DAMP, PF, WPB, MINB, CH, PU>), nb, WPB * 32, beg, end, order, P, L, h->xq.p, cur, h->ef.p, nxt, h->partial.p + h->partial_off, Q, h->scf_stop)
  if (push) { if (change) GO(true, true); else GO(false, true); }
  else { if (change) GO(true, false); else GO(false, false); }
#undef GO
  return end - beg;
}

template <int WPB, int MINB>
static int launch_cached(polb200_handle *h, int beg, int end, const int *order, ListRows L, const double4 *cur,
                         double4 *nxt, bool change, bool push)
{
  const int nb = cdiv(end - beg, WPB);
  const PushArgs Q = push ? push_args(h, nxt) : PushArgs{};
#define GO(CH, PU) \
  LAUNCH(h, (k_sweep_cached<WPB, MINB, CH, PU>), nb, WPB * 32, beg, end, order, L, h->s12.p, h->xq.p, cur, h->ef.p, nxt, h->partial.p + h->partial_off, Q, h->scf_stop)
  if (push) { if (change) GO(true, true); else GO(false, true); }
  else { if (change) GO(true, false); else GO(false, false); }
#undef GO
  return end - beg;
}

// per-step radial cache of the pair-group rows (k_group_cache), built by the first sweep of a step.  Returns false
// (and drops the groups) when the cache does not fit comfortably in free HBM.
static bool ensure_group_cache(polb200_handle *h, const DevParams &P, bool with_rmin = false)
{
  if (!h->groups_built) return false;
  if (h->group_cache_valid) return true;
  const bool damp = P.pc.damping_exponential != 0;
  bool fits = true;
  const bool chunked = h->sweep_variant >= 40;
  const size_t need_b = chunked ? (size_t)h->gchunks * GCHUNK_BYTES : (size_t)h->gneigh.cap * 36;
  if (chunked ? h->gcrec.cap < need_b : h->s12ab.cap < h->gneigh.cap) {  // only while it fits comfortably in free HBM
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    fits = (double)need_b < 0.6 * (double)free_b;
  }
  if (!fits) {
    h->groups_built = false;
    return false;
  }
  h->tgcount.ensure(h->ngroups + 1);
  if (chunked) h->gcrec.ensure(need_b + 64, 1.0);
  else { h->tgneigh.ensure(h->gneigh.cap); h->s12ab.ensure(h->gneigh.cap); }
  const int ngb = cdiv(h->ngroups, WARPS_PER_BLOCK);
#define GC(DA, CK, RM) \
  LAUNCH(h, (k_group_cache<DA, CK, RM>), ngb, BLOCK, h->ngroups, P, h->group_first.p, h->group_two.p, h->growstart.p, h->gcount.p, \
         h->gneigh.p, h->xq.p, h->tgneigh.p, h->tgcount.p, h->s12ab.p, h->gcstart.p, h->gcrec.p, (const double4 *)h->mua.p, \
         (const int2 *)h->tm.p, h->rmin_bits.p)
  if (with_rmin && chunked) { if (damp) GC(true, true, true); else GC(false, true, true); }
  else if (damp) { if (chunked) GC(true, true, false); else GC(true, false, false); }
  else { if (chunked) GC(false, true, false); else GC(false, false, false); }
#undef GC
  h->rmin_fused = with_rmin && chunked;
  h->group_cache_valid = true;
  return true;
}

// launch of the TMA-fed pair-group sweep: whole-system Jacobi sweep (gsc == nullptr) or one colour of the
// group-coloured Gauss-Seidel sweep
template <int GW, int NS, int CK, bool CH, bool PU, bool EV, bool GS>
static void launch_tma(polb200_handle *h, const double4 *cur, double4 *nxt, const PushArgs &Q, const GsChunk &gsc)
{
  auto kern = k_sweep_group_tma<GW, NS, CK, CH, PU, EV, GS>;
  const int smem = GW * NS * CK * 36 + GW * NS * 8;
  static bool attr_set = false;
  static int per_sm = 0;
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GW * 32, smem));
    attr_set = true;
  }
  const int rows = GS ? gsc.end - gsc.beg : h->ngroups;
  if (rows <= 0) return;
  const int grid = std::min(cdiv(rows, GW), std::max(per_sm, 1) * h->num_sms);
  const int reverse = (!GS && h->alternate && (h->sweep_parity++ & 1)) ? 1 : 0;
  kern<<<grid, GW * 32, smem, h->stream>>>(h->ngroups, h->group_first.p, h->group_two.p, h->growstart.p, h->tgneigh.p,
                                           h->tgcount.p, h->s12ab.p, h->xq.p, cur, h->ef.p, nxt, h->partial.p, Q, h->flags.p + 4,
                                           reverse, h->gcstart.p, h->gcrec.p, h->scf_stop, gsc);
  h->launches++;
  CUDA_CHECK(cudaGetLastError());
}

template <int GW, int NS, int CK, bool GS>
static void launch_tma_flags(polb200_handle *h, const double4 *cur, double4 *nxt, bool change, bool push, const GsChunk &gsc)
{
  const PushArgs Q = push ? push_args(h, nxt) : PushArgs{};
#define T4(CH, PU) \
  do { if (h->l2_evict_first) launch_tma<GW, NS, CK, CH, PU, true, GS>(h, cur, nxt, Q, gsc); \
       else launch_tma<GW, NS, CK, CH, PU, false, GS>(h, cur, nxt, Q, gsc); } while (0)
  if (push) { if (change) T4(true, true); else T4(false, true); }
  else { if (change) T4(true, false); else T4(false, false); }
#undef T4
}

// one colour of the group-coloured Gauss-Seidel sweep: rows = the groups glist[chunk_beg[c] .. chunk_beg[c+1])
static void launch_group_gs_chunk(polb200_handle *h, int c, const DevParams &P, const double4 *cur, double4 *staging, bool change)
{
  GsChunk gsc{};
  gsc.glist = h->glist.p;
  gsc.beg = h->chunk_beg[c];
  gsc.end = h->chunk_beg[c + 1];
  gsc.polar_damp = P.pc.polar_damp;
  gsc.polar_cutsq = P.pc.polar_cutsq;
  gsc.damping_exponential = P.pc.damping_exponential;
  switch (h->sweep_variant) {
    case 40: launch_tma_flags<4, 4, 64, true>(h, cur, staging, change, false, gsc); break;
    case 44: launch_tma_flags<8, 4, 64, true>(h, cur, staging, change, false, gsc); break;
    default: launch_tma_flags<4, 3, 64, true>(h, cur, staging, change, false, gsc); break;
  }
}

static GsChunk gs_chunk_args(polb200_handle *h, int c)
{
  GsChunk gsc{};
  gsc.glist = h->glist.p;
  gsc.beg = h->chunk_beg[c];
  gsc.end = h->chunk_beg[c + 1];
  return gsc;
}

static int launch_list_sweep(polb200_handle *h, int beg, int end, const int *order, const DevParams &P, ListRows L,
                             AllPairRows A, const double4 *cur, double4 *nxt, bool change, bool push)
{
  const bool damp = P.pc.damping_exponential != 0;
  h->partial.ensure((size_t)(end - beg) + 64);
  if (h->groups_built && order == nullptr && beg == 0 && end == h->nloc) {
    // whole-system Jacobi sweep: two cell-row neighbours per warp
    if (ensure_group_cache(h, P)) {
      const PushArgs Q = push ? push_args(h, nxt) : PushArgs{};
#define GOG4(GW, MB, DP, CH, PU) \
  LAUNCH(h, (k_sweep_group<GW, MB, CH, PU, DP>), cdiv(h->ngroups, GW), GW * 32, h->ngroups, h->group_first.p, h->group_two.p, h->growstart.p, \
         h->tgneigh.p, h->tgcount.p, h->s12ab.p, h->xq.p, cur, h->ef.p, nxt, h->partial.p + h->partial_off, Q, h->scf_stop)
#define GOG(GW, MB, DP) \
  do { if (push) { if (change) GOG4(GW, MB, DP, true, true); else GOG4(GW, MB, DP, false, true); } \
       else { if (change) GOG4(GW, MB, DP, true, false); else GOG4(GW, MB, DP, false, false); } } while (0)
      const GsChunk none{};
      switch (h->sweep_variant) {
        case 40: launch_tma_flags<4, 4, 64, false>(h, cur, nxt, change, push, none); break;
        case 41: launch_tma_flags<4, 3, 64, false>(h, cur, nxt, change, push, none); break;
        case 44: launch_tma_flags<8, 4, 64, false>(h, cur, nxt, change, push, none); break;
        case 31: GOG(4, 5, true); break;
        default: GOG(4, 6, false); break;
      }
#undef GOG4
#undef GOG
      return end - beg;
    }
  }
  if (h->sweep_variant >= 20 && !h->s12_valid && h->s12.cap < h->neigh.cap) {
    // the radial cache costs 16 B per list entry: keep it only while it fits comfortably in free HBM
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if ((double)h->neigh.cap * sizeof(double2) > 0.5 * (double)free_b) h->sweep_variant = 6;  // matrix-free
  }
  if (h->sweep_variant >= 20) {
    if (!h->s12_valid) {  // first sweep of this step: build the radial cache
      h->s12.ensure(h->neigh.cap);
      const int nrb = cdiv(h->nloc, WARPS_PER_BLOCK);
      if (damp) LAUNCH(h, (k_radial_cache<true>), nrb, BLOCK, h->nloc, P, L, h->xq.p, h->s12.p);
      else LAUNCH(h, (k_radial_cache<false>), nrb, BLOCK, h->nloc, P, L, h->xq.p, h->s12.p);
      h->s12_valid = true;
    }
    if (h->sweep_variant == 21) return launch_cached<8, 4>(h, beg, end, order, L, cur, nxt, change, push);
    return launch_cached<4, 8>(h, beg, end, order, L, cur, nxt, change, push);
  }
  if (!order && h->sweep_variant == 0) {  // first version (kept as the simplest statement of the sweep; per-block partials)
    const int nb = cdiv(end - beg, WARPS_PER_BLOCK);
    LAUNCH(h, (k_sweep<true>), nb, BLOCK, beg, end, order, P, L, A, h->xq.p, cur, h->ef.p, nxt, h->partial.p, h->scf_stop);
    return nb;
  }
  // matrix-free: no per-pair cache (used when 16 B per pair do not fit in HBM, or on request: variant 6)
  return damp ? launch_v2<true, 1, 4, 10>(h, beg, end, order, P, L, cur, nxt, change, push)
              : launch_v2<false, 1, 4, 10>(h, beg, end, order, P, L, cur, nxt, change, push);
}

// Colouring of the pair groups for the group-coloured Gauss-Seidel sweep (kernels.cuh), once per rebuild:
// strong-coupling adjacency -> parallel greedy colouring in priority order (rank metric, then a hash) -> the groups
// sorted by colour (glist) and the start of every colour's slice (chunk_beg).
static void build_colouring(polb200_handle *h, const double *metric_caller)
{
  const int ngr = h->ngroups, n = h->nloc, C = h->ncolours;
  h->gof.ensure(n); h->gmetric.ensure(ngr); h->gadj.ensure((size_t)ngr * GS_MAXADJ); h->gadjn.ensure(ngr);
  h->gcolour.ensure(ngr); h->glist.ensure(ngr); h->cstart_dev.ensure(C + 2);
  LAUNCH(h, k_group_of, cdiv(ngr, 256), 256, ngr, h->group_first.p, h->group_two.p, h->perm.p, metric_caller, h->gof.p, h->gmetric.p);
  // radius of a "strong" coupling: the sphere that holds gs_strong_m atoms at the mean density of the system
  const double pi = 3.14159265358979323846;
  const double natoms = h->comm.active ? (double)h->comm.nglobal : (double)n;
  const double rho = natoms / (h->box.prd[0] * h->box.prd[1] * h->box.prd[2]);
  double rs = cbrt(3.0 * h->gs_strong_m / (4.0 * pi * rho));
  rs = std::min(rs, h->style.cutneighmax);
  LAUNCH(h, k_group_adjacency, cdiv(ngr, WARPS_PER_BLOCK), BLOCK, ngr, n, rs * rs, h->group_first.p, h->group_two.p, h->growstart.p,
         h->gcount.p, h->gneigh.p, h->xq.p, h->gof.p, (h->comm.active || h->nghost == 0) ? (const int *)nullptr : h->g_owner.p,
         h->gadj.p, h->gadjn.p);
  CUDA_CHECK(cudaMemsetAsync(h->gcolour.p, 0xFF, (size_t)ngr * sizeof(int), h->stream));
  int *remaining = h->flags.p + 7;
  h->colour_rounds = 0;
  for (int batch = 0; batch < 4096; batch++) {
    for (int r = 0; r < 8; r++) {
      if (r == 7) CUDA_CHECK(cudaMemsetAsync(remaining, 0, sizeof(int), h->stream));
      LAUNCH(h, k_colour_round, cdiv(ngr, 256), 256, ngr, C, h->gadj.p, h->gadjn.p, h->gmetric.p, h->gcolour.p, remaining);
    }
    h->colour_rounds += 8;
    int left = 0;
    CUDA_CHECK(cudaMemcpyAsync(&left, remaining, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (left == 0) break;
    if (batch == 4095) throw CudaError{"group colouring did not terminate"};
  }
  // groups by colour (stable: ascending group index inside a colour, i.e. the cell-sorted streaming order)
  const int m = std::max(ngr, 64);
  h->keys.ensure(m); h->keys2.ensure(m); h->vals.ensure(m);
  CUDA_CHECK(cudaMemcpyAsync(h->keys.p, h->gcolour.p, (size_t)ngr * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
  LAUNCH(h, k_iota, cdiv(ngr, 256), 256, ngr, h->vals.p);
  sort_pairs(h, ngr, h->keys.p, h->keys2.p, h->vals.p, h->glist.p, 6);
  LAUNCH(h, k_cell_starts, 1, 64, C, ngr, h->keys2.p, h->cstart_dev.p, 0);
  CUDA_CHECK(cudaMemcpyAsync(h->chunk_beg, h->cstart_dev.p, (size_t)(C + 1) * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->chunk_beg[C] = ngr;
  h->colours_valid = true;
}

template <int NV>
static void reduce_partials(polb200_handle *h, int nblocks, double *out, int accumulate)
{
  LAUNCH(h, k_reduce_partials<NV>, 1, 256, nblocks, h->partial.p, out, accumulate, (const int *)nullptr);
}

// scal layout (device doubles): 0..7 pair partial sums, 8..16 polarization sums, 17 change, 18 rmin
enum { S_PAIR = 0, S_POL = 8, S_CHANGE = 17, S_RMIN = 18, S_N = 24 };

static void compute_impl(polb200_handle *h, const polb200_atoms *at, int eflag, int vflag, int ago,
                         polb200_result *out)
{
  HostStyle &st = h->style;
  if (!st.initialized) throw StyleError{POLB200_ERR_STATE, "polb200_init has not been called"};
  if (!h->params_uploaded) upload_params(h);
  if (!h->box_set) throw StyleError{POLB200_ERR_STATE, "polb200_set_box has not been called"};
  const bool eflag_atom = (eflag / 2) != 0, vflag_atom = (vflag / 4) != 0;  // src/pair.cpp:763-771
  if ((eflag_atom && !at->eatom) || (vflag_atom && !at->vatom))
    throw StyleError{POLB200_ERR_ARG, "per-atom tallies requested (eflag & 2 / vflag & 4) without eatom / vatom arrays"};
  const int n = at->nlocal;
  memset(out, 0, sizeof(*out));
  const bool comm = h->comm.active;
  if (n <= 0) {
    // decomposed: bricks are only re-populated at rebuild steps, whose first collective is the error agreement of
    // comm_build_ghosts -- take part in it so that every brick stops with the same message
    if (comm) comm_throw(std::max(comm_agree(h, COMM_ERR_EMPTY), (int)COMM_ERR_EMPTY));
    return;
  }
  if (!at->x || !at->q || !at->type || !at->alpha || !at->mu || !at->f)
    throw StyleError{POLB200_ERR_ARG, "polb200_compute: the arrays x, q, type, alpha, mu and f are required"};
  const bool dev = at->on_device != 0;
  const bool list_mode = st.polar_cutoff > 0.0;
  if (comm && !list_mode)
    throw StyleError{POLB200_ERR_UNSUPPORTED,
                     "spatial decomposition needs a dipole cutoff (polar_cutoff <r>): the reference's all-pairs "
                     "minimum-image interaction set does not decompose into bricks (polb200_comm_init_replicated shares it by rows)"};
  // all-pairs mode shared by rows: every process holds the whole system (polb200_comm_init_replicated)
  const bool xsplit = h->xr.active;
  if (xsplit && list_mode)
    throw StyleError{POLB200_ERR_UNSUPPORTED, "polb200_comm_init_replicated shares the all-pairs mode; with polar_cutoff use polb200_comm_init (bricks)"};
  if (xsplit && (eflag_atom || vflag_atom))
    throw StyleError{POLB200_ERR_UNSUPPORTED, "per-atom tallies are not available when the all-pairs mode is shared by several GPUs"};
  // rows of this process: equal chunks, multiples of the row-block size, so that every per-block partial sum is the
  // single-GPU one and the fixed-order reductions give the same bits on every process
  const int xchunk = xsplit ? cdiv(cdiv(n, h->xr.nranks), WARPS_PER_BLOCK) * WARPS_PER_BLOCK : n;
  const int xr0 = xsplit ? std::min(n, h->xr.rank * xchunk) : 0, xr1 = xsplit ? std::min(n, xr0 + xchunk) : n;
  // all-gather of `per_row` doubles per row (or per row block: rows = blocks) of an array whose rows [lo, hi) this process wrote
  auto xgather = [&](double *arr, int per_row, int chunk_rows, int lo, int hi, int total_rows) {
    const size_t cnt = (size_t)chunk_rows * per_row;
    h->xg.ensure(cnt * h->xr.nranks + 8);
    double *mine = h->xg.p + cnt * h->xr.rank;
    CUDA_CHECK(cudaMemsetAsync(mine, 0, cnt * sizeof(double), h->stream));
    if (hi > lo)
      CUDA_CHECK(cudaMemcpyAsync(mine, arr + (size_t)lo * per_row, (size_t)(hi - lo) * per_row * sizeof(double), cudaMemcpyDeviceToDevice,
                                 h->stream));
    NCCL_CHECK(g_nccl.AllGather(mine, h->xg.p, cnt, ncclDouble, h->xr.nccl, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(arr, h->xg.p, (size_t)total_rows * per_row * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  };
  for (int d = 0; d < 3; d++) {
    if (!h->box.periodic[d]) continue;
    if (st.cutneighmax + h->atom_slack > h->box.prd[d])
      throw StyleError{POLB200_ERR_UNSUPPORTED, "neighbor cutoff exceeds the periodic box length"};
    if (list_mode && (st.polar_cutoff > h->box.half[d] || st.cut_coul > h->box.half[d]))
      throw StyleError{POLB200_ERR_UNSUPPORTED,
                       "polar_cutoff (list) mode needs polar_cutoff and cut_coul <= half the box length"};
  }
  if (list_mode && st.polar_cutoff > st.cutforce)
    throw StyleError{POLB200_ERR_UNSUPPORTED, "polar_cutoff must not exceed the largest pair cutoff"};

  const int eflag_global = eflag & 1;
  int vflag_global = vflag % 4;
  const bool evflag = eflag_global || vflag_global || eflag_atom || vflag_atom;
  const bool vpair = vflag_global == 1;  // pairwise tallies; 2 = F.r (src/pair.cpp:809-815)

  CUDA_CHECK(cudaEventRecord(h->ev[0], h->stream));
  // ---- inputs of this step ----
  stage_in(h, h->c_x, at->x, (size_t)3 * n, dev);
  stage_in(h, h->c_mu, at->mu, (size_t)3 * n, dev);
  bool need = (ago == 0) || !h->have_lists || n != h->nloc;
  if (comm && ago > 0 && h->have_lists && n != h->nloc)
    throw StyleError{POLB200_ERR_STATE, "number of owned atoms changed between neighbor rebuilds"};
  if (ago < 0) need = decide_rebuild(h, n);
  if (need) rebuild(h, at);
  else {
    stage_in(h, h->c_q, at->q, n, dev);
    stage_in(h, h->c_alpha, at->alpha, n, dev);
    LAUNCH(h, k_refresh_local, cdiv(n, 256), 256, n, h->perm.p, h->c_x.p, h->c_mu.p, h->c_q.p, h->c_alpha.p, h->xq.p, h->mua.p);
    // (peer push: the neighbours may still be reading these ghosts in their previous step => barrier first)
    ghost_update(h, true, h->mua.p, true);
  }
  const int ng = h->nghost;
  const int nrowblocks = cdiv(n, WARPS_PER_BLOCK);
  h->ef.ensure(n); h->f_pair.ensure(n); h->f_pol.ensure(n);
  h->partial.ensure((size_t)nrowblocks * 16 + 64);
  h->scal.ensure(S_N);
  h->h_scal.ensure(S_N);
  CUDA_CHECK(cudaMemsetAsync(h->scal.p, 0, S_N * sizeof(double), h->stream));
  CUDA_CHECK(cudaEventRecord(h->ev[1], h->stream));

  const DevParams &P = h->P;
  ListRows L{h->rowstart.p, h->neigh.p, nullptr};
  AllPairRows A{n, h->perm.p};
  h->s12_valid = false;
  h->group_cache_valid = false;
  h->rmin_fused = false;
  // how this step's sweeps and pair kernels run
  const bool gs_mode = !st.zodid && (st.polar_gs || st.polar_gs_ranked);
  // list-mode Gauss-Seidel: group-coloured sweep on the TMA pair-group kernel (default), or -- with an explicit
  // gs_chunks setting, without pair groups, or when their cache does not fit -- per-atom chunks of the ranked order
  bool coloured = gs_mode && list_mode && st.gs_chunks == 0 && h->groups_built && h->sweep_variant >= 40;
  // LJ + Coulomb + field and the polarization forces on the pair-group rows: every interaction of the step must
  // reach exactly as far as the dipole cutoff (the tight group rows hold the partners inside it), no exclusion rules,
  // no per-atom / pairwise tallies.  The pair kernel also needs "no special bonds" (the group rows carry no special-bond
  // classes); the polarization forces only look at molecule ids, so molecular systems (BASELINE config 4) take the
  // grouped force kernel with the per-atom pair kernel.
  const bool grouped_force_ok = list_mode && h->groups_built && h->sweep_variant >= 40 && h->use_group_pairs &&
                                h->excl.n == 0 && !eflag_atom && !vflag_atom && !vpair &&
                                std::max(st.cutforce, st.cut_coul) <= st.polar_cutoff;
  bool grouped_pf = grouped_force_ok && !h->molecular;   // stage 2 on the group rows
  bool grouped_force = grouped_force_ok;                 // stage 4 on the group rows
  if (coloured || grouped_force) {
    // steps that keep their colouring only report rmin: the group cache produces it on the way
    const bool fuse_rmin = coloured && st.polar_gs_ranked && h->colours_valid;
    if (fuse_rmin) {
      h->rmin_bits.ensure(1);
      const unsigned long long init = (unsigned long long)0x408F400000000000ull;  // bits of 1000.0 (pol.cpp:196)
      CUDA_CHECK(cudaMemcpyAsync(h->rmin_bits.p, &init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    }
    if (!ensure_group_cache(h, P, fuse_rmin)) coloured = grouped_pf = grouped_force = false;
  }
  const bool jacobi_groups = !gs_mode && h->groups_built;
  const bool need_tight = h->use_tight && !(grouped_pf && (st.zodid || coloured || jacobi_groups));
  if (need_tight) {
    // every pair kernel of this step only needs partners within the largest interaction cutoff
    double reach = st.cutforce > st.cut_coul ? st.cutforce : st.cut_coul;
    h->tneigh.ensure(h->neigh.cap);
    h->tcount.ensure(n);
    LAUNCH(h, k_tighten, nrowblocks, BLOCK, n, reach * reach, h->rowstart.p, h->neigh.p, h->xq.p, h->tneigh.p, h->tcount.p);
    L = ListRows{h->rowstart.p, h->tneigh.p, h->tcount.p};
  }

  double *ea_ptr = nullptr, *vp_ptr = nullptr, *vq_ptr = nullptr;
  if (eflag_atom) { h->ea_row.ensure(n); ea_ptr = h->ea_row.p; }
  if (vflag_atom) {
    h->va_pair_row.ensure((size_t)6 * n); h->va_pol_row.ensure((size_t)6 * n);
    vp_ptr = h->va_pair_row.p; vq_ptr = h->va_pol_row.p;
  }
  // ---- stage 2: LJ + Coulomb (+ static field) ----
#define PAIR_ARGS n, P, h->xq.p, h->tm.p, L, h->f_pair.p, h->ef.p, h->partial.p, ea_ptr, vp_ptr, h->excl, h->exb.p, h->g_owner.p
  const int ngroupblocks = cdiv(h->ngroups, GPF_WARPS);
  if (grouped_pf) {
    h->partial.ensure((size_t)ngroupblocks * 16 + 64);
#define PG(EV) LAUNCH(h, (k_pair_group<EV, 4>), ngroupblocks, GPF_WARPS * 32, h->ngroups, P, h->group_first.p, h->group_two.p, h->tgcount.p, \
                      h->gcstart.p, h->gcrec.p, h->xq.p, h->tm.p, h->f_pair.p, h->ef.p, h->partial.p)
    if (evflag) PG(true); else PG(false);
#undef PG
  } else if (h->excl.n > 0) {  // neigh_modify exclude: separate instantiations, the default kernels are untouched
    if (list_mode) {
      if (evflag) LAUNCH(h, (k_pair<true, true, true>), nrowblocks, BLOCK, PAIR_ARGS);
      else LAUNCH(h, (k_pair<false, true, true>), nrowblocks, BLOCK, PAIR_ARGS);
    } else {
      if (evflag) LAUNCH(h, (k_pair<true, false, true>), nrowblocks, BLOCK, PAIR_ARGS);
      else LAUNCH(h, (k_pair<false, false, true>), nrowblocks, BLOCK, PAIR_ARGS);
    }
  } else if (list_mode) {
    if (evflag) LAUNCH(h, (k_pair<true, true, false>), nrowblocks, BLOCK, PAIR_ARGS);
    else LAUNCH(h, (k_pair<false, true, false>), nrowblocks, BLOCK, PAIR_ARGS);
  } else {
    if (evflag) LAUNCH(h, (k_pair<true, false, false>), nrowblocks, BLOCK, PAIR_ARGS);
    else LAUNCH(h, (k_pair<false, false, false>), nrowblocks, BLOCK, PAIR_ARGS);
  }
#undef PAIR_ARGS
  if (evflag) reduce_partials<NPAIR_PART>(h, grouped_pf ? ngroupblocks : nrowblocks, h->scal.p + S_PAIR, 0);
  if (!list_mode) {
    if (xr1 > xr0) LAUNCH(h, k_static_allpairs, cdiv(xr1 - xr0, WARPS_PER_BLOCK), BLOCK, n, P, h->xq.p, h->tm.p, h->perm.p, h->ef.p, xr0, xr1);
    if (xsplit) xgather(reinterpret_cast<double *>(h->ef.p), 4, xchunk, xr0, xr1, n);
  }
  if (!st.use_previous) LAUNCH(h, k_init_mu, cdiv(n, 256), 256, n, st.polar_gamma, h->ef.p, h->mua.p);
  ghost_update(h, false, h->mua.p);
  CUDA_CHECK(cudaEventRecord(h->ev[2], h->stream));

  // ---- stage 3: self-consistent dipoles (DipoleSolverIterative, pol.cpp:1113-1238) ----
  int iterations = 0;
  bool diverged = false;
  if (!st.zodid) {
    const bool gs = gs_mode;
    // Gauss-Seidel visits atoms in the CALLER's index order (ranked_array = identity, pol.cpp:1127),
    // i.e. position c -> cell-sorted index invperm[c]; Jacobi does not care about the order.
    const int *order = gs ? h->invperm.p : nullptr;
    if (st.polar_gs_ranked) {
      // rank metric + stable descending sort, ties by caller index (pol.cpp:192-227,1127-1143)
      h->rmin_bits.ensure(1);
      h->metric.ensure(n); h->metric2.ensure(n); h->ranked.ensure(n); h->ranked_in.ensure(n);
      ListRows Lfull{h->rowstart.p, h->neigh.p, nullptr};
      if (!h->rmin_fused) {
        const unsigned long long init = (unsigned long long)0x408F400000000000ull;  // bits of 1000.0
        CUDA_CHECK(cudaMemcpyAsync(h->rmin_bits.p, &init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
        LAUNCH(h, k_rmin, nrowblocks, BLOCK, n, Lfull, h->xq.p, h->mua.p, h->tm.p, h->rmin_bits.p);
      }
      if (comm) comm_allreduce(h, h->rmin_bits.p, 1, ncclUint64, ncclMin);  // positive doubles order like their bits
      // (rmin travels to the host with the step's scalars: no synchronisation here)
      CUDA_CHECK(cudaMemcpyAsync(h->scal.p + S_RMIN, h->rmin_bits.p, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      // the coloured sweep uses the metric as the priority of its colouring, fixed between two rebuilds
      if (!coloured || !h->colours_valid)
        LAUNCH(h, k_rank_metric, nrowblocks, BLOCK, n, Lfull, h->xq.p, h->mua.p, h->tm.p, h->rmin_bits.p, h->perm.p, h->g_owner.p, h->g_shift.p,
               comm ? h->tag.p : (const int *)nullptr, h->metric.p);
      if (!coloured) {
        // values in caller order = sorted index of caller atom c
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, h->metric.p, h->metric2.p, h->invperm.p, h->ranked.p, n, 0, 64, h->stream);
        h->cub_tmp.ensure(bytes);
        CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(h->cub_tmp.p, bytes, h->metric.p, h->metric2.p, h->invperm.p, h->ranked.p, n, 0, 64, h->stream));
        h->launches += 3;
        order = h->ranked.p;
      }
    }
    if (coloured && !h->colours_valid) build_colouring(h, st.polar_gs_ranked ? h->metric.p : nullptr);
    // which sweep realisation
    int nchunks = 0;          // 0: Jacobi (one block of rows, separate output array)
    bool sequential = false;  // reference-exact Gauss-Seidel (all-pairs mode only)
    if (gs) {
      if (coloured) nchunks = h->ncolours;
      else if (st.gs_chunks != 0) nchunks = abs(st.gs_chunks);
      else if (list_mode) nchunks = std::min(8, n);  // per-atom fallback: 8 INTERLEAVED chunks of the ranked order
      else sequential = true;
    }
    if (gs && !coloured && nchunks > 1 && st.gs_chunks <= 0 && !sequential) {
      h->ranked_in.ensure(n);
      LAUNCH(h, k_interleave_order, cdiv(n, 256), 256, n, nchunks, order, h->ranked_in.p);
      order = h->ranked_in.p;
    }
    // fused sweep + halo: new dipoles go straight into the neighbour bricks' ghost slots (Jacobi sweeps)
    const bool push = h->push_ready && h->use_push && !gs && list_mode && h->sweep_variant != 0 &&
                      !(h->sweep_variant >= 11 && h->sweep_variant <= 14);
    const bool fused_commit = h->push_ready && h->use_push && (!comm || h->comm.push.enabled);
    const double natoms_norm = comm ? (double)h->comm.nglobal : (double)n;
    const double prec2 = st.polar_precision * st.polar_precision;
    const bool want_change = !st.fixed_iteration;
    const int itmax = st.iterations_max;
    if (list_mode) h->partial.ensure((size_t)n + 64);
    h->slice.ensure(SCF_SUM_BLOCKS);

    // precision mode: the convergence test runs on the device and raises a stop flag that every kernel of the loop
    // honours, so the host may enqueue iterations ahead of its (lagged) look at the flag
    int *ctl = nullptr;
    if (want_change) {
      h->ctl.ensure(4);
      h->h_ctl.ensure(8);
      if (!h->ev_it[0])
        for (auto &e : h->ev_it) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctl = h->ctl.p;
      CUDA_CHECK(cudaMemsetAsync(ctl, 0, 4 * sizeof(int), h->stream));
    }
    h->scf_stop = ctl;
    const int *stop = ctl;
    GsPairCache gs_cache{nullptr, 0};
    if (sequential && h->gs_blocked && n <= h->gs_cache_max) {
      h->gs_planes.ensure((size_t)5 * n * n);
      LAUNCH(h, k_gs_pair_cache, nrowblocks, BLOCK, n, order, P, h->perm.p, h->xq.p, h->gs_planes.p);
      gs_cache = GsPairCache{h->gs_planes.p, n};
    }

    // one iteration `it` of the solver (Jacobi: reads the buffer of parity it, writes the other one)
    auto enqueue_iteration = [&](int it) {
      double4 *cur = h->mua.p, *nxt = h->mub.p;
      if (!gs && (it & 1)) std::swap(cur, nxt);
      double *chg = h->scal.p + S_CHANGE;
      bool summed = false;  // the squared change is already in *chg
      int nparts = 0;
      if (sequential && h->gs_blocked) {
        // blocked forward substitution = the same sweep, N/32 dependent steps instead of N (kernels.cuh)
        h->gsR.ensure(n);
        LAUNCH(h, k_gsb_upper, nrowblocks, BLOCK, n, order, P, h->perm.p, h->xq.p, cur, h->ef.p, h->gsR.p, stop, gs_cache);
        // one launch per block (k_gsb_step); the launches of a sweep are replayed from a CUDA graph, captured once and
        // kept for as long as its arguments stay the same (at 750 atoms a sweep is 25 dependent kernels of a few us)
        const int nblk = cdiv(n, GSB);
        // the lower-triangular part in one launch by a thread-block cluster (k_gsb_cluster) when the device can schedule it
        bool by_cluster = false;
        if (h->gs_cluster > 0 && nblk > 1) {
          int ncta = std::min(h->gs_cluster, 16);
          while (ncta > 1 && (ncta - 1) * GSB >= 2 * n) ncta /= 2;   // no more helpers than rows to update
          static int nonportable = -1;
          if (nonportable < 0) {
            nonportable = cudaFuncSetAttribute(k_gsb_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess ? 1 : 0;
            CUDA_CHECK(cudaFuncSetAttribute(k_gsb_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, GSC_SMEM));
          }
          if (!nonportable) ncta = std::min(ncta, 8);
          cudaLaunchConfig_t cfg = {};
          cudaLaunchAttribute attr[1];
          bool launched = false;
          for (; ncta >= 2 && !launched; ncta /= 2) {
            cfg.gridDim = dim3(ncta);
            cfg.blockDim = dim3(GSS_THREADS);
            cfg.dynamicSmemBytes = GSC_SMEM;
            cfg.stream = h->stream;
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = ncta;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, k_gsb_cluster, &cfg) != cudaSuccess || nclusters < 1) {
              cudaGetLastError();
              continue;
            }
            CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_gsb_cluster, n, order, P, (const int *)h->perm.p, (const double4 *)h->xq.p, cur, h->gsR.p, chg,
                                          stop, gs_cache));
            h->launches++;
            launched = true;
          }
          by_cluster = launched;
        }
        if (!by_cluster) {
        const void *key[8] = {order, h->perm.p, h->xq.p, cur, h->gsR.p, chg, stop, gs_cache.plane};
        auto launch_steps = [&] {
          for (int b = 0; b < nblk; b++) {
            const int rows_after = n - (b + 1) * GSB;
            const int grid = 1 + (b > 0 && rows_after > 0 ? cdiv(rows_after, GSB) : 0);
            LAUNCH(h, k_gsb_step, grid, GSS_THREADS, n, b, order, P, h->perm.p, h->xq.p, cur, h->gsR.p, chg, stop, gs_cache);
          }
        };
        polb200_handle::SweepGraph &G = h->gs_graph;
        if (!h->use_graphs) launch_steps();
        else {
          if (!G.exec || G.n != n || G.version != h->params_version || memcmp(G.key, key, sizeof(key)) != 0) {
            if (G.exec) cudaGraphExecDestroy(G.exec);
            G.exec = nullptr;
            cudaGraph_t graph = nullptr;
            const long before = h->launches;
            CUDA_CHECK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            try {
              launch_steps();
            } catch (...) {
              cudaStreamEndCapture(h->stream, &graph);
              if (graph) cudaGraphDestroy(graph);
              throw;
            }
            CUDA_CHECK(cudaStreamEndCapture(h->stream, &graph));
            G.nodes = (int)(h->launches - before);
            h->launches = before;
            const cudaError_t ie = cudaGraphInstantiate(&G.exec, graph, 0);
            cudaGraphDestroy(graph);
            CUDA_CHECK(ie);
            G.n = n;
            G.version = h->params_version;
            memcpy(G.key, key, sizeof(key));
          }
          CUDA_CHECK(cudaGraphLaunch(G.exec, h->stream));
          h->launches += G.nodes;
        }
        }
        summed = true;
      } else if (sequential) {
        LAUNCH(h, k_gs_sequential, 1, GS_THREADS, n, order, P, h->perm.p, h->xq.p, cur, h->ef.p, chg, stop);
        summed = true;
      } else if (!gs) {
        sweep_event(h);
        nparts = nrowblocks;
        if (list_mode) nparts = launch_list_sweep(h, 0, n, order, P, L, A, cur, nxt, want_change, push);
        else if (!xsplit) LAUNCH(h, (k_sweep<false>), nrowblocks, BLOCK, 0, n, order, P, L, A, h->xq.p, cur, h->ef.p, nxt, h->partial.p, stop);
        else {
          // this process's rows, then every process gets all new dipoles (24 N bytes per sweep; the arrays carry alpha
          // along) and all per-block squared changes: the reduction below sees the blocks of the single-GPU launch
          const int bchunk = xchunk / WARPS_PER_BLOCK, b0 = xr0 / WARPS_PER_BLOCK, b1 = cdiv(xr1, WARPS_PER_BLOCK);
          if (xr1 > xr0)
            LAUNCH(h, (k_sweep<false>), b1 - b0, BLOCK, xr0, xr1, order, P, L, A, h->xq.p, cur, h->ef.p, nxt, h->partial.p + b0, stop);
          xgather(reinterpret_cast<double *>(nxt), 4, xchunk, xr0, xr1, n);
          if (want_change) xgather(h->partial.p, 1, bchunk, b0, b1, nrowblocks);
        }
        sweep_event(h);
      } else if (coloured) {
        // group-coloured sweep: the colours in turn; Jacobi inside a colour (staging array + commit)
        sweep_event(h);
        for (int c = 0; c < nchunks; c++) {
          if (h->chunk_beg[c + 1] <= h->chunk_beg[c]) continue;
          launch_group_gs_chunk(h, c, P, cur, nxt, want_change);
          const int nthreads = 2 * (h->chunk_beg[c + 1] - h->chunk_beg[c]);
          if (fused_commit) {
            // in place: the neighbour bricks must have finished reading this chunk's input before their ghost
            // slots change, and must see the new values before the next chunk (two barriers around one kernel)
            if (comm) comm_signal_wait(h, nullptr);
            LAUNCH(h, k_commit_groups, cdiv(nthreads, 256), 256, gs_chunk_args(h, c), h->group_first.p, h->group_two.p, nxt, cur,
                   push_args(h, cur), 1, stop);
            if (comm) comm_signal_wait(h, nullptr);
          } else {
            LAUNCH(h, k_commit_groups, cdiv(nthreads, 256), 256, gs_chunk_args(h, c), h->group_first.p, h->group_two.p, nxt, cur,
                   PushArgs{}, 0, stop);
            ghost_update(h, false, cur, true);
          }
        }
        sweep_event(h);
        nparts = n;
      } else {
        // per-atom chunks of the ranked order, Jacobi inside, Gauss-Seidel between
        for (int c = 0; c < nchunks; c++) {
          const int beg = (int)(((long)c * n) / nchunks), end = (int)(((long)(c + 1) * n) / nchunks);
          if (end <= beg) continue;
          const int nb = cdiv(end - beg, WARPS_PER_BLOCK);
          if (list_mode) {
            // per-row squared changes of every chunk land in one array: a single reduction per iteration
            h->partial_off = (size_t)beg;
            launch_list_sweep(h, beg, end, order, P, L, A, cur, nxt, want_change, false);
            h->partial_off = 0;
          } else {
            LAUNCH(h, (k_sweep<false>), nb, BLOCK, beg, end, order, P, L, A, h->xq.p, cur, h->ef.p, nxt, h->partial.p, stop);
            if (want_change) LAUNCH(h, k_reduce_partials<1>, 1, 256, nb, h->partial.p, chg, c > 0 ? 1 : 0, stop);
            summed = true;
          }
          if (fused_commit) {
            if (comm) comm_signal_wait(h, nullptr);
            LAUNCH(h, k_commit_push, cdiv(end - beg, 256), 256, beg, end, order, nxt, cur, push_args(h, cur), stop);
            if (comm) comm_signal_wait(h, nullptr);
          } else {
            LAUNCH(h, k_commit_rows, cdiv(end - beg, 256), 256, beg, end, order, nxt, cur, stop);
            ghost_update(h, false, cur, true);
          }
        }
        nparts = n;
      }
      // squared change of this iteration -> (all-reduce) -> test
      if (want_change) {
        const double norm3n = natoms_norm * 3.0;
        if (!summed)
          LAUNCH(h, k_change_sum, SCF_SUM_BLOCKS, 256, nparts, h->partial.p, h->slice.p, chg, ctl, comm ? 0 : 1, prec2, norm3n,
                 itmax, stop);
        if (!gs) {
          if (push) {
            if (comm) comm_signal_wait(h, chg);  // barrier (+ all-reduce)
          } else {
            ghost_update(h, false, nxt);
            if (comm) comm_allreduce(h, chg, 1, ncclDouble, ncclSum);
          }
        } else if (comm) {
          comm_allreduce(h, chg, 1, ncclDouble, ncclSum);
        }
        if (summed || comm) LAUNCH(h, k_scf_check, 1, 32, chg, ctl, prec2, norm3n, itmax);
      } else if (!gs) {
        if (push) {
          if (comm) comm_signal_wait(h, nullptr);
        } else ghost_update(h, false, nxt);
      }
    };

    const size_t ev_base = h->sweep_ev_used;
    if (!want_change) {
      // fixed_iteration: Jacobi runs max_iterations+1 sweeps in the reference and discards the last (pol.cpp:1214
      // returns before the copy), so only max_iterations sweeps shape the result; the in-place writes of the
      // Gauss-Seidel modes' extra sweep stay
      const int nsweeps = gs ? itmax + 1 : itmax;
      for (int it = 0; it < nsweeps; it++) enqueue_iteration(it);
      iterations = itmax;
    } else {
      const int lag = ((comm && !h->comm.push.enabled) || xsplit) ? 0 : h->scf_lag;
      int it = 0;
      const int *seen = nullptr;
      while (!seen) {
        enqueue_iteration(it);
        int *slot = h->h_ctl.p + 4 * (it & 1);
        CUDA_CHECK(cudaMemcpyAsync(slot, ctl, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaEventRecord(h->ev_it[it & 1], h->stream));
        it++;
        const int w = it - 1 - lag;  // the iteration whose test the host looks at now
        if (w >= 0) {
          CUDA_CHECK(cudaEventSynchronize(h->ev_it[w & 1]));
          if (h->h_ctl.p[4 * (w & 1)]) seen = h->h_ctl.p + 4 * (w & 1);
        }
        if (!seen && it > itmax + 1) {  // the device stops at max_iterations+1 at the latest (divergence path)
          CUDA_CHECK(cudaStreamSynchronize(h->stream));
          seen = h->h_ctl.p + 4 * ((it - 1) & 1);
          if (!seen[0]) throw CudaError{"SCF stop flag was not raised after max_iterations + 1 iterations"};
        }
      }
      iterations = seen[1];
      diverged = seen[2] != 0;
      // events of the iterations that were enqueued past the stop are not sweeps
      h->sweep_ev_used = std::min(h->sweep_ev_used, ev_base + (size_t)2 * iterations);
    }
    h->scf_stop = nullptr;
    double4 *cur = (!gs && (iterations & 1)) ? h->mub.p : h->mua.p;
    if (diverged) LAUNCH(h, k_reset_mu, cdiv(n, 256), 256, n, h->ef.p, cur);  // pol.cpp:1227-1235
    // after fused sweeps the ghost copies of the final array are already current on every brick
    const bool ghosts_current = push && !diverged;
    if (cur != h->mua.p) {  // keep the canonical buffer
      CUDA_CHECK(cudaMemcpyAsync(h->mua.p, cur, (size_t)(ghosts_current ? n + ng : n) * sizeof(double4),
                                 cudaMemcpyDeviceToDevice, h->stream));
    }
    if (!ghosts_current) ghost_update(h, false, h->mua.p);
  }
  CUDA_CHECK(cudaEventRecord(h->ev[3], h->stream));

  // ---- stage 4: polarization forces ----
#define POLFORCE(LISTM, EV, VP)                                                                                          \
  do {                                                                                                                   \
    if (vflag_atom)                                                                                                      \
      LAUNCH(h, (k_polforce<LISTM, EV, VP, true>), nrowblocks, BLOCK, n, P, L, A, h->xq.p, h->mua.p, h->tm.p, h->f_pol.p, \
             h->partial.p, vq_ptr, 0);                                                                                   \
    else if (!xsplit)                                                                                                    \
      LAUNCH(h, (k_polforce<LISTM, EV, VP, false>), nrowblocks, BLOCK, n, P, L, A, h->xq.p, h->mua.p, h->tm.p,            \
             h->f_pol.p, h->partial.p, (double *)nullptr, 0);                                                            \
    else {                                                                                                               \
      /* all-pairs mode shared by rows: own rows, then all forces and all per-block energy / virial partials */         \
      const int bchunk = xchunk / WARPS_PER_BLOCK, b0 = xr0 / WARPS_PER_BLOCK, b1 = cdiv(xr1, WARPS_PER_BLOCK);           \
      if (xr1 > xr0)                                                                                                     \
        LAUNCH(h, (k_polforce<LISTM, EV, VP, false>), b1 - b0, BLOCK, xr1, P, L, A, h->xq.p, h->mua.p, h->tm.p,          \
               h->f_pol.p, h->partial.p + (size_t)b0 * NPOL_PART, (double *)nullptr, xr0);                               \
      xgather(reinterpret_cast<double *>(h->f_pol.p), 4, xchunk, xr0, xr1, n);                                           \
      if (EV) xgather(h->partial.p, NPOL_PART, bchunk, b0, b1, nrowblocks);                                              \
    }                                                                                                                    \
  } while (0)
  // the reference tallies polarization energies whenever eflag is set, virial via F.r or pairwise
  const bool ev4 = evflag;
  if (grouped_force) {
#define FG4(EV) LAUNCH(h, (k_polforce_group<EV, 4>), ngroupblocks, GPF_WARPS * 32, h->ngroups, P, h->group_first.p, h->group_two.p, h->tgcount.p, \
                       h->gcstart.p, h->gcrec.p, h->xq.p, h->mua.p, h->tm.p, h->f_pol.p, h->partial.p)
#define FG(EV) if (h->gpf_minb == 4) FG4(EV); else LAUNCH(h, (k_polforce_group<EV, 5>), ngroupblocks, GPF_WARPS * 32, h->ngroups, P, h->group_first.p, h->group_two.p, h->tgcount.p, \
                      h->gcstart.p, h->gcrec.p, h->xq.p, h->mua.p, h->tm.p, h->f_pol.p, h->partial.p)
    if (ev4) { FG(true); } else { FG(false); }
#undef FG
#undef FG4
  } else if (list_mode) {
    if (!ev4) POLFORCE(true, false, false);
    else if (vpair) POLFORCE(true, true, true);
    else POLFORCE(true, true, false);
  } else {
    if (!ev4) POLFORCE(false, false, false);
    else if (vpair) POLFORCE(false, true, true);
    else POLFORCE(false, true, false);
  }
#undef POLFORCE
  if (ev4) reduce_partials<NPOL_PART>(h, grouped_force ? ngroupblocks : nrowblocks, h->scal.p + S_POL, 0);

  // ---- stage 5: outputs ----
  h->c_f.ensure((size_t)3 * n); h->c_ef.ensure((size_t)3 * n);
  LAUNCH(h, k_scatter_out, cdiv(n, 256), 256, n, h->perm.p, h->f_pair.p, h->f_pol.p, h->mua.p, h->ef.p, h->c_f.p,
         h->c_mu.p, h->c_ef.p);
  if (eflag_atom || vflag_atom) {
    if (eflag_atom) h->c_eatom.ensure(n);
    if (vflag_atom) h->c_vatom.ensure((size_t)6 * n);
    LAUNCH(h, k_scatter_atomev, cdiv(n, 256), 256, n, h->perm.p, ea_ptr, vp_ptr, vq_ptr, eflag_atom ? h->c_eatom.p : nullptr,
           vflag_atom ? h->c_vatom.p : nullptr);
  }
  CUDA_CHECK(cudaEventRecord(h->ev[4], h->stream));
  CUDA_CHECK(cudaMemcpyAsync(h->h_scal.p, h->scal.p, S_N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  h->h_int.ensure(8);
  CUDA_CHECK(cudaMemcpyAsync(h->h_int.p, h->flags.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (dev) {
    // device-resident caller: f += pair force, mu/ef overwritten, all on the stream
    LAUNCH(h, k_add_inplace, cdiv((long)3 * n, 256), 256, (long)3 * n, h->c_f.p, at->f);
    if (eflag_atom) LAUNCH(h, k_add_inplace, cdiv((long)n, 256), 256, (long)n, h->c_eatom.p, at->eatom);
    if (vflag_atom) LAUNCH(h, k_add_inplace, cdiv((long)6 * n, 256), 256, (long)6 * n, h->c_vatom.p, at->vatom);
    CUDA_CHECK(cudaMemcpyAsync(at->mu, h->c_mu.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (at->ef_static)
      CUDA_CHECK(cudaMemcpyAsync(at->ef_static, h->c_ef.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CUDA_CHECK(cudaEventRecord(h->ev[5], h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  } else {
    h->h_stage.ensure((size_t)3 * n + (eflag_atom ? n : 0) + (vflag_atom ? (size_t)6 * n : 0));
    CUDA_CHECK(cudaMemcpyAsync(h->h_stage.p, h->c_f.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    double *h_ea = h->h_stage.p + (size_t)3 * n, *h_va = h_ea + (eflag_atom ? n : 0);
    if (eflag_atom) CUDA_CHECK(cudaMemcpyAsync(h_ea, h->c_eatom.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (vflag_atom) CUDA_CHECK(cudaMemcpyAsync(h_va, h->c_vatom.p, (size_t)6 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(at->mu, h->c_mu.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (at->ef_static)
      CUDA_CHECK(cudaMemcpyAsync(at->ef_static, h->c_ef.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaEventRecord(h->ev[5], h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    double *f = at->f;
    const double *a = h->h_stage.p;
    for (size_t k = 0; k < (size_t)3 * n; k++) f[k] += a[k];
    if (eflag_atom)
      for (size_t k = 0; k < (size_t)n; k++) at->eatom[k] += h_ea[k];
    if (vflag_atom)
      for (size_t k = 0; k < (size_t)6 * n; k++) at->vatom[k] += h_va[k];
  }

  sweep_events_collect(h);
  if (comm) {
    const int tflag = h->h_int.p[6];
    if (tflag) throw CudaError{"a neighbour brick did not reach the inter-GPU barrier within the timeout "
                               "(peer process gone, or several ranks on one GPU)"};
  }
  const double *sc = h->h_scal.p;
  if (eflag_global) {
    out->eng_vdwl = sc[S_PAIR + 0];
    out->eng_coul = sc[S_PAIR + 1];
  }
  if (evflag) {
    // u_polar is assigned to eng_pol unconditionally (pol.cpp:632,641); its pieces are only
    // accumulated under eflag (pol.cpp:432,476,499,538)
    if (eflag) {
      out->u_self = sc[S_POL + 0];
      out->u_ef = sc[S_POL + 1];
      out->u_dd = sc[S_POL + 2];
    }
    out->eng_pol = out->u_self + out->u_ef + out->u_dd;
    if (vflag_global)
      for (int k = 0; k < 6; k++) out->virial[k] = sc[S_PAIR + 2 + k] + sc[S_POL + 3 + k];
  }
  out->iterations = iterations;
  out->rmin = sc[S_RMIN];
  out->status = (diverged ? POLB200_STATUS_DIVERGED : 0) | (need ? POLB200_STATUS_REBUILT : 0) |
                (list_mode ? 0 : POLB200_STATUS_EXACT);
  out->npairs_full = (long)h->npairs;
  out->nghost = ng;
  cudaEventElapsedTime(&out->ms_neigh, h->ev[0], h->ev[1]);
  cudaEventElapsedTime(&out->ms_pair, h->ev[1], h->ev[2]);
  cudaEventElapsedTime(&out->ms_scf, h->ev[2], h->ev[3]);
  cudaEventElapsedTime(&out->ms_force, h->ev[3], h->ev[4]);
  cudaEventElapsedTime(&out->ms_total, h->ev[0], h->ev[5]);
}

}  // namespace polb200

// =====================================================================================================
// reciprocal-space Ewald
// =====================================================================================================
#include "ewald.cuh"

struct polb200_ewald {
  std::string err;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {};
  long launches = 0;
  bool ready = false;
  // plan (Ewald::init + setup, ewald.cpp:87-340)
  double g_ewald = 0.0, gsqmx = 0.0, accuracy = 0.0, qqrd2e = 0.0, qsum = 0.0, qsqsum = 0.0, volume = 0.0;
  double unitk[3] = {0, 0, 0};
  int kxmax = 0, kymax = 0, kzmax = 0, kmax = 0, nk = 0, nquads = 0, slots = 0;
  DBuf<EwaldK> kv;
  // column form of the same k set (ewald.cuh, round 2): the default; the quad kernels stay as the cross-check (POLB200_EWALD_QUADS=1)
  DBuf<EwaldCol> cols, vcols;
  DBuf<double> ugv;
  DBuf<int> kxyz;
  DBuf<double2> W;
  int nvalid = 0, kxhi_max = 0;
  bool use_cols = true;
  DBuf<double2> S, Spart, phase;
  DBuf<double> c_x, c_q, c_f, out, fpart;
  HPinned<double> h_out, h_f;
  int num_sms = 148;
  int sfac_smem_set = 0, force_smem_set = 0;
  float ms_last = 0.f;
  // multi-GPU (polb200_ewald_comm_init): every rank holds its own atoms, the structure factors are all-reduced
  ncclComm_t nccl = nullptr;
  int rank = 0, nranks = 1;
};

namespace polb200 {

static double ewald_rms(int km, double prd, long natoms, double q2, double g)
{
  if (natoms == 0) natoms = 1;  // ewald.cpp:345
  const double pi = 3.14159265358979323846;
  return 2.0 * q2 * g / prd * sqrt(1.0 / (pi * km * natoms)) * exp(-pi * pi * km * km / (g * g * prd * prd));
}

// host: the half-space k set, kx fastest inside (kz, ky) so that warps of the structure-factor kernel read
// consecutive shared-memory slots.  Same membership as Ewald::coeffs (ewald.cpp:760-1026): first non-zero
// component positive; axis vectors up to kmax, others up to the per-dimension maxima; |k|^2 <= gsqmx.
static bool ewald_in_set(const polb200_ewald *e, int kx, int ky, int kz, double &sqk)
{
  if (kx < 0 || (kx == 0 && (ky < 0 || (ky == 0 && kz <= 0)))) return false;
  const int nz = (kx != 0) + (ky != 0) + (kz != 0);
  if (nz == 1) {
    if (abs(kx) > e->kmax || abs(ky) > e->kmax || abs(kz) > e->kmax) return false;
  } else if (abs(kx) > e->kxmax || abs(ky) > e->kymax || abs(kz) > e->kzmax) return false;
  const double a = kx * e->unitk[0], b = ky * e->unitk[1], c = kz * e->unitk[2];
  sqk = a * a + b * b + c * c;
  return sqk <= e->gsqmx;
}

// quads of four consecutive kx per (ky,kz) row; slots of a row's last quad that fall outside the set get ug = 0.
// Returns the number of real k-vectors (Ewald::kcount).
static int ewald_build_kset(polb200_ewald *e, std::vector<EwaldK> &out)
{
  const double pi = 3.14159265358979323846;
  const double ginv2 = 1.0 / (e->g_ewald * e->g_ewald), preu = 4.0 * pi / e->volume;
  out.clear();
  int real = 0;
  for (int kz = -e->kmax; kz <= e->kmax; kz++)
    for (int ky = -e->kmax; ky <= e->kmax; ky++) {
      int lo = -1, hi = -1;
      double sqk;
      for (int kx = 0; kx <= e->kmax; kx++)
        if (ewald_in_set(e, kx, ky, kz, sqk)) {
          if (lo < 0) lo = kx;
          hi = kx;
        }
      if (lo < 0) continue;
      for (int q0 = lo; q0 <= hi; q0 += 4) {
        EwaldK k;
        k.kx0 = q0; k.ky = ky; k.kz = kz; k.pad = 0;
        for (int j = 0; j < 4; j++) {
          k.ug[j] = 0.0;
          if (q0 + j <= hi && ewald_in_set(e, q0 + j, ky, kz, sqk)) {
            k.ug[j] = preu * exp(-0.25 * sqk * ginv2) / sqk;
            real++;
          }
        }
        out.push_back(k);
      }
    }
  return real;
}

// column form: per (ky,kz) the contiguous run kx = lo..hi of the same k set; flat arrays ordered ky, kz, kx
static int ewald_build_columns(polb200_ewald *e, std::vector<EwaldCol> &cols, std::vector<EwaldCol> &valid, std::vector<double> &ug,
                               std::vector<int> &kxyz)
{
  const double pi = 3.14159265358979323846;
  const double ginv2 = 1.0 / (e->g_ewald * e->g_ewald), preu = 4.0 * pi / e->volume;
  cols.clear(); valid.clear(); ug.clear(); kxyz.clear();
  e->kxhi_max = 0;
  for (int ky = -e->kmax; ky <= e->kmax; ky++)
    for (int kz = -e->kmax; kz <= e->kmax; kz++) {
      EwaldCol c;
      c.ky = (short)ky; c.kz = (short)kz; c.lo = 0; c.hi = -1; c.off = (int)ug.size(); c.pad = 0;
      int lo = -1, hi = -1, cnt = 0;
      double sqk;
      for (int kx = 0; kx <= e->kmax; kx++)
        if (ewald_in_set(e, kx, ky, kz, sqk)) {
          if (lo < 0) lo = kx;
          hi = kx;
          cnt++;
        }
      if (lo >= 0) {
        if (cnt != hi - lo + 1) throw StyleError{POLB200_ERR_STATE, "polb200_ewald: the k set of a (ky,kz) column is not one run of kx"};
        c.lo = (short)lo; c.hi = (short)hi;
        for (int kx = lo; kx <= hi; kx++) {
          ewald_in_set(e, kx, ky, kz, sqk);
          ug.push_back(preu * exp(-0.25 * sqk * ginv2) / sqk);
          kxyz.push_back(kx | ((ky + 512) << 10) | ((kz + 512) << 20));
        }
        e->kxhi_max = std::max(e->kxhi_max, hi);
        valid.push_back(c);
      }
      cols.push_back(c);
    }
  // the structure-factor kernel skips blocks of kx above the largest `hi` of a warp's columns: hand it the columns
  // longest first, so that the columns of a warp are (nearly) equally long and the sphere, not the box, is evaluated
  std::stable_sort(valid.begin(), valid.end(), [](const EwaldCol &a, const EwaldCol &b) { return a.hi > b.hi; });
  return (int)ug.size();
}

}  // namespace polb200

// =====================================================================================================
// C ABI
// =====================================================================================================

template <class F>
static int guarded(polb200_t *h, F &&fn)
{
  try {
    fn();
    return POLB200_OK;
  } catch (const StyleError &e) {
    if (h) h->err = e.msg;
    return e.code;
  } catch (const CudaError &e) {
    if (h) h->err = e.msg;
    return POLB200_ERR_CUDA;
  } catch (const std::exception &e) {
    if (h) h->err = e.what();
    return POLB200_ERR_ARG;
  }
}

template <class F>
static int ewald_guarded(polb200_ewald *e, F &&fn)
{
  try {
    fn();
    return POLB200_OK;
  } catch (const StyleError &x) {
    e->err = x.msg;
    return x.code;
  } catch (const CudaError &x) {
    e->err = x.msg;
    return POLB200_ERR_CUDA;
  } catch (const std::exception &x) {
    e->err = x.what();
    return POLB200_ERR_ARG;
  }
}

extern "C" {

int polb200_abi_version(void) { return POLB200_ABI_VERSION; }

int polb200_create(polb200_t **out, int device)
{
  if (!out) return POLB200_ERR_ARG;
  *out = nullptr;
  if (device == POLB200_DEVICE_NONE) {  // configuration-only handle: every device entry point fails loudly
    polb200_t *hh = new polb200_handle();
    hh->device = POLB200_DEVICE_NONE;
    *out = hh;
    return POLB200_OK;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    fprintf(stderr, "polb200_create: no usable CUDA device %d (found %d); there is no CPU fallback\n", device, count);
    return POLB200_ERR_CUDA;
  }
  polb200_t *h = new polb200_handle();
  h->device = device;
  if (const char *v = getenv("POLB200_SWEEP_VARIANT")) h->sweep_variant = atoi(v);  // experiments only
  if (const char *v = getenv("POLB200_USE_PUSH")) h->use_push = atoi(v) != 0;
  if (const char *v = getenv("POLB200_GPF_MINB")) h->gpf_minb = atoi(v);
  if (const char *v = getenv("POLB200_GROUP_PAIRS")) h->use_group_pairs = atoi(v) != 0;
  if (const char *v = getenv("POLB200_GS_CLUSTER")) h->gs_cluster = std::max(0, std::min(16, atoi(v)));
  if (const char *v = getenv("POLB200_XSORT_BITS")) h->xsort_bits = atoi(v);
  if (const char *v = getenv("POLB200_BIN_DIV")) h->bin_div = atof(v);
  int rc = guarded(h, [&] {
    CUDA_CHECK(cudaSetDevice(device));
    CUDA_CHECK(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
    CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto &e : h->ev) CUDA_CHECK(cudaEventCreate(&e));
  });
  if (rc) {
    fprintf(stderr, "polb200_create: %s\n", h->err.c_str());
    delete h;
    return rc;
  }
  *out = h;
  return POLB200_OK;
}

void polb200_destroy(polb200_t *h)
{
  if (!h) return;
  if (h->device == POLB200_DEVICE_NONE) {
    delete h;
    return;
  }
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->comm.active) {
    CommState &c = h->comm;
    comm_close_peers(h);
    h->mua.graveyard = h->mub.graveyard = h->xq.graveyard = nullptr;
    for (void *q : c.graveyard) cudaFree(q);
    c.graveyard.clear();
    for (auto *b : {&c.send_owner_u, &c.send_owner, &c.send_dir, &c.slot_of_u, &c.dir_start, &c.gslot, &c.counts_dev,
                    &c.dir_of_u})
      b->release();
    c.push_ptrx.release();
    c.flags.release(); c.sbuf.release(); c.rbuf.release(); c.sbufi.release(); c.rbufi.release();
    c.ipc_dev.release();
    if (c.nccl) g_nccl.CommDestroy(c.nccl);
    c.active = false;
  }
  if (h->xr.nccl) g_nccl.CommDestroy(h->xr.nccl);
  h->xr = polb200_handle::ExactShare{};
  h->xg.release();
  for (auto *b : {&h->c_x, &h->c_q, &h->c_alpha, &h->c_mu, &h->c_f, &h->c_ef, &h->c_xhold, &h->d_coeff,
                  &h->d_tables, &h->partial, &h->scal, &h->metric, &h->metric2, &h->ea_row, &h->va_pair_row,
                  &h->va_pol_row, &h->c_eatom, &h->c_vatom})
    b->release();
  for (auto *b : {&h->c_type, &h->c_mol, &h->c_tag, &h->c_nspecial, &h->c_special, &h->tag, &h->perm,
                  &h->invperm, &h->keys, &h->keys2, &h->vals, &h->vals2, &h->g_owner_u, &h->g_shift_u,
                  &h->g_owner, &h->g_shift, &h->cl_start, &h->cg_start, &h->stencil, &h->s_nspecial,
                  &h->s_special, &h->neigh, &h->tneigh, &h->tcount, &h->flags, &h->ranked, &h->ranked_in})
    b->release();
  for (auto *b : {&h->xq, &h->mua, &h->mub, &h->ef, &h->f_pair, &h->f_pol}) b->release();
  h->group_first.release(); h->group_two.release(); h->gneigh.release(); h->gcount.release(); h->tgneigh.release(); h->tgcount.release();
  h->gsR.release(); h->growstart.release(); h->s12ab.release(); h->gcstart.release(); h->gcrec.release();
  h->s12.release(); h->push_off.release(); h->push_ptr0.release(); h->push_ptr1.release();
  h->gof.release(); h->gadj.release(); h->gadjn.release(); h->gcolour.release(); h->glist.release(); h->cstart_dev.release();
  h->col_export.release(); h->gmetric.release(); h->ctl.release(); h->h_ctl.release(); h->slice.release();
  for (auto &e : h->ev_it) if (e) cudaEventDestroy(e);
  if (h->gs_graph.exec) cudaGraphExecDestroy(h->gs_graph.exec);
  h->gs_planes.release();
  h->tm.release(); h->cnt.release(); h->rowstart.release(); h->cub_tmp.release(); h->rmin_bits.release();
  h->h_stage.release(); h->h_scal.release(); h->h_int.release();
  for (auto &e : h->ev) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char *polb200_last_error(const polb200_t *h) { return h ? h->err.c_str() : "null handle"; }

int polb200_settings(polb200_t *h, int narg, const char *const *arg)
{
  if (!h) return POLB200_ERR_ARG;
  return guarded(h, [&] { h->style.settings(narg, arg); h->params_uploaded = false; });
}

int polb200_set_ntypes(polb200_t *h, int ntypes)
{
  if (!h) return POLB200_ERR_ARG;
  return guarded(h, [&] { h->style.set_ntypes(ntypes); h->params_uploaded = false; });
}

int polb200_coeff(polb200_t *h, int narg, const char *const *arg)
{
  if (!h) return POLB200_ERR_ARG;
  return guarded(h, [&] { h->style.coeff(narg, arg); h->params_uploaded = false; });
}

int polb200_pair_modify(polb200_t *h, int narg, const char *const *arg)
{
  if (!h) return POLB200_ERR_ARG;
  return guarded(h, [&] { h->style.pair_modify(narg, arg); h->params_uploaded = false; });
}

int polb200_init(polb200_t *h, const polb200_env *env)
{
  if (!h || !env) return POLB200_ERR_ARG;
  return guarded(h, [&] {
    h->style.init(*env);
    h->params_uploaded = false;  // device copies are refreshed by the next compute()
  });
}

int polb200_init_one(const polb200_t *h, int i, int j, double *cut)
{
  if (!h || !cut || !h->style.initialized || i < 1 || j < 1 || i > h->style.ntypes || j > h->style.ntypes)
    return POLB200_ERR_ARG;
  *cut = h->style.cut_pair[h->style.idx(i, j)];
  return POLB200_OK;
}

int polb200_tail(const polb200_t *h, int i, int j, double count_i, double count_j, double *etail_ij, double *ptail_ij)
{
  if (!h || !etail_ij || !ptail_ij || !h->style.initialized || i < 1 || j < 1 || i > h->style.ntypes || j > h->style.ntypes)
    return POLB200_ERR_ARG;
  h->style.tail_correction(i, j, count_i, count_j, *etail_ij, *ptail_ij);
  return POLB200_OK;
}

const void *polb200_extract(const polb200_t *h, const char *name, int *dim)
{
  if (!h || !name || !dim) return nullptr;
  *dim = 0;
  if (!strcmp(name, "cut_coul")) return &h->style.cut_coul;
  *dim = 2;
  if (!strcmp(name, "epsilon")) return h->style.epsilon.data();
  if (!strcmp(name, "sigma")) return h->style.sigma.data();
  // integration-layer extras (not part of the reference's extract()): int setflag[(ntypes+1)^2], double cut_lj[...]
  if (!strcmp(name, "setflag")) return h->style.setflag.data();
  if (!strcmp(name, "cut_lj")) return h->style.cut_lj.data();
  return nullptr;
}

int polb200_single(const polb200_t *h, int itype, int jtype, double qi, double qj, double rsq,
                   double factor_coul, double factor_lj, double *fforce, double *eng)
{
  if (!h || !fforce || !eng || !h->style.initialized) return POLB200_ERR_STATE;
  if (itype < 1 || jtype < 1 || itype > h->style.ntypes || jtype > h->style.ntypes) return POLB200_ERR_ARG;
  *eng = h->style.single(itype, jtype, qi, qj, rsq, factor_coul, factor_lj, *fforce);
  return POLB200_OK;
}

int polb200_restart_size(const polb200_t *h, long *nbytes)
{
  if (!h || !nbytes) return POLB200_ERR_ARG;
  *nbytes = (long)h->style.restart_image().size();
  return POLB200_OK;
}

int polb200_write_restart(const polb200_t *h, void *buf, long nbytes)
{
  if (!h || !buf) return POLB200_ERR_ARG;
  std::vector<char> img = h->style.restart_image();
  if ((long)img.size() > nbytes) return POLB200_ERR_ARG;
  memcpy(buf, img.data(), img.size());
  return POLB200_OK;
}

int polb200_restart_settings_size(const polb200_t *h, long *nbytes)
{
  if (!h || !nbytes) return POLB200_ERR_ARG;
  *nbytes = (long)h->style.restart_settings_image().size();
  return POLB200_OK;
}

int polb200_read_restart_settings(polb200_t *h, const void *buf, long nbytes, long *consumed)
{
  if (!h || !buf) return POLB200_ERR_ARG;
  return guarded(h, [&] {
    const long used = h->style.read_restart_settings_image(buf, nbytes);
    if (consumed) *consumed = used;
    h->params_uploaded = false;
  });
}

int polb200_read_restart(polb200_t *h, const void *buf, long nbytes)
{
  if (!h || !buf) return POLB200_ERR_ARG;
  return guarded(h, [&] { h->style.read_restart_image(buf, nbytes); h->params_uploaded = false; });
}

int polb200_set_box(polb200_t *h, const double boxlo[3], const double boxhi[3], const int periodic[3])
{
  if (!h || !boxlo || !boxhi || !periodic) return POLB200_ERR_ARG;
  return guarded(h, [&] {
    bool changed = !h->box_set;
    for (int d = 0; d < 3; d++) {
      if (!(boxhi[d] > boxlo[d])) throw StyleError{POLB200_ERR_ARG, "Box bounds are invalid"};
      changed = changed || h->box.lo[d] != boxlo[d] || h->box.hi[d] != boxhi[d] || h->box.periodic[d] != periodic[d];
      h->box.lo[d] = boxlo[d];
      h->box.hi[d] = boxhi[d];
      h->box.prd[d] = boxhi[d] - boxlo[d];
      h->box.half[d] = 0.5 * h->box.prd[d];
      h->box.periodic[d] = periodic[d];
    }
    h->P.box = h->box;
    h->box_set = true;
    if (changed) h->params_version++;
    if (changed) h->have_lists = false;
    if (h->comm.active && (changed || !h->comm.geom_valid)) comm_setup_geom(h);
  });
}

int polb200_compute(polb200_t *h, const polb200_atoms *atoms, int eflag, int vflag, int ago, polb200_result *out)
{
  if (!h || !atoms || !out) return POLB200_ERR_ARG;
  return guarded(h, [&] {
    if (h->device == POLB200_DEVICE_NONE)
      throw CudaError{"polb200_compute on a configuration-only handle: a CUDA device is required (no CPU fallback)"};
    CUDA_CHECK(cudaSetDevice(h->device));
    compute_impl(h, atoms, eflag, vflag, ago, out);
  });
}

int polb200_set_exclusions(polb200_t *h, int nrules, const polb200_exclusion *rules)
{
  if (!h || nrules < 0 || (nrules > 0 && !rules)) return POLB200_ERR_ARG;
  return guarded(h, [&] {
    if (nrules > MAX_EXCL) throw StyleError{POLB200_ERR_UNSUPPORTED, "more than 32 neigh_modify exclude / include rules"};
    ExclRules X{};
    X.n = nrules;
    for (int r = 0; r < nrules; r++) {
      if (rules[r].kind < EXCL_TYPE || rules[r].kind > EXCL_INCLUDE) throw StyleError{POLB200_ERR_ARG, "Illegal neigh_modify command"};
      X.kind[r] = rules[r].kind;
      X.a[r] = rules[r].a;
      X.b[r] = rules[r].b;
    }
    h->excl = X;
    h->have_lists = false;  // the membership bits are rebuilt with the neighbor structures
  });
}

long polb200_launch_count(polb200_t *h, int reset)
{
  if (!h) return -1;
  long v = h->launches;
  if (reset) h->launches = 0;
  return v;
}

int polb200_set_option(polb200_t *h, const char *name, double value)
{
  if (!h || !name) return POLB200_ERR_ARG;
  if (!strcmp(name, "sweep_variant")) {
    h->sweep_variant = (int)value;
    h->have_lists = false;  // the pair groups are built with the neighbor structures
    return POLB200_OK;
  }
  if (!strcmp(name, "bin_div")) {
    if (!(value >= 1.0 && value <= 12.0)) return POLB200_ERR_ARG;
    h->bin_div = value;
    h->have_lists = false;
    return POLB200_OK;
  }
  if (!strcmp(name, "p2p_push")) {  // 0: NCCL halo per sweep; 1: fused sweep + peer-memory push (takes effect at the next rebuild)
    h->comm.want_push = value != 0.0;
    h->comm.mapped_ptr[0] = nullptr;
    h->have_lists = false;
    return POLB200_OK;
  }
  if (!strcmp(name, "xsort_bits")) {
    h->xsort_bits = (int)value;
    h->have_lists = false;
    return POLB200_OK;
  }
  if (!strcmp(name, "gs_cache_max")) {
    h->gs_cache_max = (int)value;
    return POLB200_OK;
  }
  if (!strcmp(name, "atom_slack")) {
    if (!(value >= 0.0 && value <= 16.0)) return POLB200_ERR_ARG;
    h->atom_slack = value;
    h->have_lists = false;
    return POLB200_OK;
  }
  if (!strcmp(name, "use_graphs")) {
    h->use_graphs = value != 0.0;
    return POLB200_OK;
  }
  if (!strcmp(name, "gs_blocked")) {
    h->gs_blocked = value != 0.0;
    return POLB200_OK;
  }
  if (!strcmp(name, "l2_evict_first")) {
    h->l2_evict_first = value != 0.0;
    return POLB200_OK;
  }
  if (!strcmp(name, "alternate")) {
    h->alternate = value != 0.0;
    return POLB200_OK;
  }
  if (!strcmp(name, "use_push")) {
    h->use_push = value != 0.0;
    return POLB200_OK;
  }
  if (!strcmp(name, "gpf_minb")) {
    h->gpf_minb = (int)value;
    return POLB200_OK;
  }
  if (!strcmp(name, "use_group_pairs")) {
    h->use_group_pairs = value != 0.0;
    return POLB200_OK;
  }
  if (!strcmp(name, "gs_cluster")) {
    if (value < 0.0 || value > 16.0) return POLB200_ERR_ARG;
    h->gs_cluster = (int)value;
    return POLB200_OK;
  }
  if (!strcmp(name, "use_tight")) {
    h->use_tight = value != 0.0;
    return POLB200_OK;
  }
  if (!strcmp(name, "gs_colours")) {  // colours (= chunks per sweep) of the group-coloured Gauss-Seidel sweep
    if (!(value >= 1.0 && value <= 32.0)) return POLB200_ERR_ARG;
    h->ncolours = (int)value;
    h->colours_valid = false;
    return POLB200_OK;
  }
  if (!strcmp(name, "gs_strong_m")) {  // atoms inside the "strong coupling" radius of the colouring (mean density)
    if (!(value >= 0.0 && value <= 64.0)) return POLB200_ERR_ARG;
    h->gs_strong_m = value;
    h->colours_valid = false;
    return POLB200_OK;
  }
  if (!strcmp(name, "scf_lag")) {  // 1: one iteration enqueued ahead of the host's look at the stop flag; 0: none
    h->scf_lag = value != 0.0 ? 1 : 0;
    return POLB200_OK;
  }
  if (!strcmp(name, "time_sweeps")) {
    h->time_sweeps = value != 0.0;
    h->sweep_ms_accum = 0.0;
    h->sweep_launches = 0;
    return POLB200_OK;
  }
  h->err = std::string("unknown option ") + name;
  return POLB200_ERR_ARG;
}

long polb200_debug_fetch(polb200_t *h, const char *name, void *dst, long capacity_bytes)
{
  if (!h || !name || !dst) return -1;
  long result = -1;
  {  // host-side tables (no device needed): "h_<name>"
    const HostStyle &st = h->style;
    const std::vector<double> *v = nullptr;
    const struct { const char *n; const std::vector<double> *v; } tabs[] = {
        {"h_rtable", &st.tab.r}, {"h_drtable", &st.tab.dr}, {"h_ftable", &st.tab.f}, {"h_dftable", &st.tab.df},
        {"h_ctable", &st.tab.c}, {"h_dctable", &st.tab.dc}, {"h_etable", &st.tab.e}, {"h_detable", &st.tab.de},
        {"h_cutsq", &st.cutsq}, {"h_cut_ljsq", &st.cut_ljsq}, {"h_lj1", &st.lj1}, {"h_lj2", &st.lj2},
        {"h_lj3", &st.lj3}, {"h_lj4", &st.lj4}, {"h_offset", &st.offset}, {"h_cutneighsq", &st.cutneighsq}};
    for (auto &e : tabs)
      if (!strcmp(name, e.n)) v = e.v;
    if (v) {
      if ((long)(v->size() * 8) > capacity_bytes) return -1;
      memcpy(dst, v->data(), v->size() * 8);
      return (long)v->size();
    }
    if (!strcmp(name, "h_tabmeta")) {  // mask, shift, tabinnersq, cutneighmax
      if (capacity_bytes < 32) return -1;
      double m[4] = {(double)st.tab.mask, (double)st.tab.shift, st.tab.tabinnersq, st.cutneighmax};
      memcpy(dst, m, 32);
      return 4;
    }
  }
  if (h->device == POLB200_DEVICE_NONE) return -1;
  guarded(h, [&] {
    CUDA_CHECK(cudaSetDevice(h->device));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    const int n = h->nloc, ng = h->nghost;
    auto fetch = [&](const void *src, size_t bytes, long count) {
      if ((long)bytes > capacity_bytes) throw StyleError{POLB200_ERR_ARG, "debug_fetch: buffer too small"};
      CUDA_CHECK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
      result = count;
    };
    if (!strcmp(name, "sweep_timing")) {  // {accumulated ms, launches} since time_sweeps was set
      if (capacity_bytes < 16) throw StyleError{POLB200_ERR_ARG, "debug_fetch: buffer too small"};
      double v[2] = {h->sweep_ms_accum, (double)h->sweep_launches};
      memcpy(dst, v, 16);
      result = 2;
    } else if (!strcmp(name, "polar_pairs")) {  // ordered pairs inside the dipole cutoff (list mode)
      if (capacity_bytes < 8) throw StyleError{POLB200_ERR_ARG, "debug_fetch: buffer too small"};
      h->cnt.ensure(1);
      CUDA_CHECK(cudaMemsetAsync(h->cnt.p, 0, 8, h->stream));
      ListRows L{h->rowstart.p, h->neigh.p, nullptr};
      LAUNCH(h, k_count_polar_pairs, cdiv(n, WARPS_PER_BLOCK), BLOCK, n, h->P.pc.polar_cutsq, L, h->xq.p, h->cnt.p);
      CUDA_CHECK(cudaMemcpyAsync(dst, h->cnt.p, 8, cudaMemcpyDeviceToHost, h->stream));
      CUDA_CHECK(cudaStreamSynchronize(h->stream));
      result = 1;
    } else if (!strcmp(name, "comm_stats")) {  // {nsend, nrecv, nglobal, push enabled, nranks}
      if (capacity_bytes < 40) throw StyleError{POLB200_ERR_ARG, "debug_fetch: buffer too small"};
      double v[5] = {(double)h->comm.nsend, (double)h->comm.nrecv, (double)h->comm.nglobal,
                     (double)h->comm.push.enabled, (double)h->comm.nranks};
      memcpy(dst, v, 40);
      result = 5;
    } else if (!strcmp(name, "barrier_stats")) {  // {ns spent inside inter-GPU barriers, barriers} since comm_init (this rank)
      if (capacity_bytes < 16) throw StyleError{POLB200_ERR_ARG, "debug_fetch: buffer too small"};
      double v[2] = {0.0, 0.0};
      if (h->comm.active && h->comm.flags.p) {
        unsigned long long u[2];
        CUDA_CHECK(cudaMemcpy(u, h->comm.flags.p + 3 * MAX_PEERS, sizeof(u), cudaMemcpyDeviceToHost));
        v[0] = (double)u[0];
        v[1] = (double)u[1];
      }
      memcpy(dst, v, 16);
      result = 2;
    } else if (!strcmp(name, "group_stats")) {  // {groups, union entries inside the cutoff (last step), skin entries, built}
      if (capacity_bytes < 32) throw StyleError{POLB200_ERR_ARG, "debug_fetch: buffer too small"};
      double v[4] = {(double)h->ngroups, 0.0, (double)h->gpairs, h->groups_built ? 1.0 : 0.0};
      if (h->groups_built && h->group_cache_valid && h->ngroups > 0) {
        std::vector<int> tc(h->ngroups);
        CUDA_CHECK(cudaMemcpy(tc.data(), h->tgcount.p, (size_t)h->ngroups * sizeof(int), cudaMemcpyDeviceToHost));
        for (int c : tc) v[1] += c;
      }
      memcpy(dst, v, 32);
      result = 4;
    } else if (!strcmp(name, "flags")) fetch(h->flags.p, 32, 8);
    else if (!strcmp(name, "tag")) fetch(h->tag.p, (size_t)(n + ng) * 4, n + ng);
    else if (!strcmp(name, "perm")) fetch(h->perm.p, (size_t)n * 4, n);
    else if (!strcmp(name, "ghost_owner")) fetch(h->g_owner.p, (size_t)ng * 4, ng);   // sorted owned index
    else if (!strcmp(name, "ghost_shift")) fetch(h->g_shift.p, (size_t)ng * 4, ng);   // packed code
    else if (!strcmp(name, "rowstart")) fetch(h->rowstart.p, (size_t)(n + 1) * 8, n + 1);
    else if (!strcmp(name, "neigh")) fetch(h->neigh.p, (size_t)h->npairs * 4, (long)h->npairs);
    else if (!strcmp(name, "xq")) fetch(h->xq.p, (size_t)(n + ng) * 32, (long)(n + ng) * 4);
    else if (!strcmp(name, "mua")) fetch(h->mua.p, (size_t)(n + ng) * 32, (long)(n + ng) * 4);
    else if (!strcmp(name, "ef")) fetch(h->ef.p, (size_t)n * 32, (long)n * 4);
    else if (!strcmp(name, "ranked")) fetch(h->ranked.p, (size_t)n * 4, n);
    else if (!strcmp(name, "gs_colouring")) {
      // group-coloured sweep: {colour, in-group predecessor (caller index or -1)} of every owned atom, caller order;
      // then {number of colours, rounds of the greedy colouring, groups}
      if (!h->colours_valid) throw StyleError{POLB200_ERR_STATE, "debug_fetch: no group colouring (not a list-mode Gauss-Seidel run)"};
      h->col_export.ensure((size_t)2 * n + 4);
      LAUNCH(h, k_colour_export, cdiv(h->ngroups, 256), 256, h->ngroups, h->group_first.p, h->group_two.p, h->gcolour.p, h->perm.p,
             h->col_export.p, h->col_export.p + n);
      const int tail[3] = {h->ncolours, h->colour_rounds, h->ngroups};
      CUDA_CHECK(cudaMemcpyAsync(h->col_export.p + 2 * (size_t)n, tail, sizeof(tail), cudaMemcpyHostToDevice, h->stream));
      CUDA_CHECK(cudaStreamSynchronize(h->stream));
      fetch(h->col_export.p, ((size_t)2 * n + 3) * 4, (long)2 * n + 3);
    }
    else if (!strcmp(name, "metric")) fetch(h->metric.p, (size_t)n * 8, n);
    else throw StyleError{POLB200_ERR_ARG, std::string("debug_fetch: unknown array ") + name};
  });
  return result;
}

// ---- device buffers for resident callers ------------------------------------------------------------------
void *polb200_dev_alloc(int device, size_t bytes)
{
  void *p = nullptr;
  if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void polb200_dev_free(int device, void *p)
{
  if (!p) return;
  cudaSetDevice(device);
  cudaFree(p);
}

int polb200_dev_copy(int device, void *dst, const void *src, size_t bytes, int kind)
{
  if (bytes == 0) return POLB200_OK;
  if (!dst || !src || kind < 0 || kind > 2) return POLB200_ERR_ARG;
  const cudaMemcpyKind k = kind == POLB200_COPY_H2D ? cudaMemcpyHostToDevice
                                                    : (kind == POLB200_COPY_D2H ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
  if (cudaSetDevice(device) != cudaSuccess || cudaMemcpy(dst, src, bytes, k) != cudaSuccess) {
    cudaGetLastError();
    return POLB200_ERR_CUDA;
  }
  return POLB200_OK;
}

int polb200_dev_zero(int device, void *p, size_t bytes)
{
  if (bytes == 0) return POLB200_OK;
  if (!p) return POLB200_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess || cudaMemset(p, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    cudaGetLastError();
    return POLB200_ERR_CUDA;
  }
  return POLB200_OK;
}

int polb200_host_register(void *p, size_t bytes)
{
  if (!p || !bytes) return POLB200_ERR_ARG;
  if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) {
    cudaGetLastError();
    return POLB200_ERR_CUDA;
  }
  return POLB200_OK;
}

int polb200_host_unregister(void *p)
{
  if (!p) return POLB200_ERR_ARG;
  if (cudaHostUnregister(p) != cudaSuccess) {
    cudaGetLastError();
    return POLB200_ERR_CUDA;
  }
  return POLB200_OK;
}

// ---- multi-GPU entry points (comm.cuh) ---------------------------------------------------------------
int polb200_comm_id_size(void) { return (int)sizeof(ncclUniqueId); }

int polb200_comm_create_id(void *id_bytes)
{
  if (!id_bytes) return POLB200_ERR_ARG;
  std::string err;
  if (!g_nccl.load(err)) {
    fprintf(stderr, "polb200_comm_create_id: %s\n", err.c_str());
    return POLB200_ERR_UNSUPPORTED;
  }
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return POLB200_ERR_CUDA;
  memcpy(id_bytes, &id, sizeof(id));
  return POLB200_OK;
}

int polb200_comm_init_replicated(polb200_t *h, int rank, int nranks, const void *id_bytes)
{
  if (!h || !id_bytes) return POLB200_ERR_ARG;
  return guarded(h, [&] {
    if (h->device == POLB200_DEVICE_NONE) throw CudaError{"polb200_comm_init_replicated needs a CUDA device"};
    if (h->comm.active || h->xr.active) throw StyleError{POLB200_ERR_STATE, "a communicator was already set up for this handle"};
    if (nranks < 1 || rank < 0 || rank >= nranks || nranks > 64) throw StyleError{POLB200_ERR_ARG, "Bad grid of processors"};
    std::string err;
    if (!g_nccl.load(err)) throw StyleError{POLB200_ERR_UNSUPPORTED, err};
    CUDA_CHECK(cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    NCCL_CHECK(g_nccl.CommInitRank(&h->xr.nccl, nranks, id, rank));
    h->xr.rank = rank;
    h->xr.nranks = nranks;
    h->xr.active = true;
  });
}

int polb200_comm_init(polb200_t *h, int rank, int nranks, const void *id_bytes, const int procgrid[3])
{
  if (!h || !id_bytes || !procgrid) return POLB200_ERR_ARG;
  return guarded(h, [&] {
    if (h->device == POLB200_DEVICE_NONE) throw CudaError{"polb200_comm_init needs a CUDA device"};
    if (h->comm.active || h->xr.active) throw StyleError{POLB200_ERR_STATE, "polb200_comm_init was already called"};
    if (nranks < 1 || rank < 0 || rank >= nranks || procgrid[0] * procgrid[1] * procgrid[2] != nranks || nranks > 64)
      throw StyleError{POLB200_ERR_ARG, "Bad grid of processors"};
    std::string err;
    if (!g_nccl.load(err)) throw StyleError{POLB200_ERR_UNSUPPORTED, err};
    CUDA_CHECK(cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    CommState &c = h->comm;
    NCCL_CHECK(g_nccl.CommInitRank(&c.nccl, nranks, id, rank));
    c.rank = rank;
    c.nranks = nranks;
    for (int k = 0; k < 3; k++) c.pg[k] = procgrid[k];
    c.active = true;
    {  // ranks that share a GPU cannot wait on each other inside kernels (nothing guarantees co-scheduling):
       // detect it once (all-gather of the PCI identity) and keep such runs on the NCCL halo
      cudaDeviceProp prop;
      CUDA_CHECK(cudaGetDeviceProperties(&prop, h->device));
      int id[2] = {(prop.pciDomainID << 16) | (prop.pciBusID << 8) | prop.pciDeviceID, 0};
      DBuf<int> d;
      d.ensure((size_t)2 * (nranks + 1));
      std::vector<int> all((size_t)2 * nranks);
      CUDA_CHECK(cudaMemcpyAsync(d.p + 2 * nranks, id, sizeof(id), cudaMemcpyHostToDevice, h->stream));
      NCCL_CHECK(g_nccl.AllGather(d.p + 2 * nranks, d.p, 2, ncclInt, c.nccl, h->stream));
      CUDA_CHECK(cudaMemcpyAsync(all.data(), d.p, all.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CUDA_CHECK(cudaStreamSynchronize(h->stream));
      for (int a = 0; a < nranks; a++)
        for (int b = a + 1; b < nranks; b++)
          if (all[2 * a] == all[2 * b]) c.shared_device = true;
      d.release();
    }
    h->mua.graveyard = h->mub.graveyard = h->xq.graveyard = &c.graveyard;
    if (const char *v = getenv("POLB200_P2P_PUSH")) c.want_push = atoi(v) != 0;
    h->have_lists = false;
    if (h->box_set) comm_setup_geom(h);
  });
}

int polb200_subdomain(const polb200_t *h, double sublo[3], double subhi[3])
{
  if (!h || !h->box_set || !sublo || !subhi) return POLB200_ERR_STATE;
  DecompPlan p;
  const int one[3] = {1, 1, 1};
  if (make_plan(h->comm.active ? h->comm.nranks : 1, h->comm.active ? h->comm.rank : 0, h->comm.active ? h->comm.pg : one,
                h->box.periodic, h->box.lo, h->box.hi, p))
    return POLB200_ERR_ARG;
  for (int d = 0; d < 3; d++) {
    sublo[d] = p.sublo[d];
    subhi[d] = p.subhi[d];
  }
  return POLB200_OK;
}

int polb200_decomp_plan(int nranks, int rank, const int procgrid[3], const int periodic[3], const double boxlo[3],
                        const double boxhi[3], int dest[27], int src[27], int wrap[81], double sublo[3], double subhi[3])
{
  if (!procgrid || !periodic || !boxlo || !boxhi || !dest || !src || !wrap || !sublo || !subhi) return POLB200_ERR_ARG;
  DecompPlan p;
  if (make_plan(nranks, rank, procgrid, periodic, boxlo, boxhi, p)) return POLB200_ERR_ARG;
  for (int d = 0; d < NDIR; d++) {
    dest[d] = p.dest[d];
    src[d] = p.src[d];
    for (int k = 0; k < 3; k++) wrap[3 * d + k] = p.wrap[d][k];
  }
  for (int k = 0; k < 3; k++) {
    sublo[k] = p.sublo[k];
    subhi[k] = p.subhi[k];
  }
  return POLB200_OK;
}


// ---- reciprocal-space Ewald ------------------------------------------------------------------------------
int polb200_ewald_create(polb200_ewald_t **out, int device)
{
  if (!out) return POLB200_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    fprintf(stderr, "polb200_ewald_create: no usable CUDA device %d (found %d); there is no CPU fallback\n", device, count);
    return POLB200_ERR_CUDA;
  }
  polb200_ewald *e = new polb200_ewald();
  e->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&e->ev[0]) != cudaSuccess || cudaEventCreate(&e->ev[1]) != cudaSuccess) {
    delete e;
    return POLB200_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, device);
  *out = e;
  return POLB200_OK;
}

void polb200_ewald_destroy(polb200_ewald_t *e)
{
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  e->kv.release(); e->S.release(); e->Spart.release(); e->phase.release();
  e->cols.release(); e->vcols.release(); e->ugv.release(); e->kxyz.release(); e->W.release(); e->fpart.release();
  e->c_x.release(); e->c_q.release(); e->c_f.release(); e->out.release();
  e->h_out.release(); e->h_f.release();
  if (e->nccl) g_nccl.CommDestroy(e->nccl);
  cudaEventDestroy(e->ev[0]); cudaEventDestroy(e->ev[1]);
  cudaStreamDestroy(e->stream);
  delete e;
}

const char *polb200_ewald_last_error(const polb200_ewald_t *e) { return e ? e->err.c_str() : "null handle"; }

int polb200_ewald_init(polb200_ewald_t *e, const polb200_ewald_setup *in, polb200_ewald_info *info)
{
  if (!e || !in) return POLB200_ERR_ARG;
  return ewald_guarded(e, [&] {
    CUDA_CHECK(cudaSetDevice(e->device));
    const double pi = 3.14159265358979323846;
    double prd[3];
    for (int d = 0; d < 3; d++) {
      prd[d] = in->boxhi[d] - in->boxlo[d];
      if (!(prd[d] > 0.0)) throw StyleError{POLB200_ERR_ARG, "Box bounds are invalid"};
      if (!in->periodic[d]) throw StyleError{POLB200_ERR_UNSUPPORTED, "Cannot use nonperiodic boundaries with Ewald"};
    }
    e->qqrd2e = in->qqrd2e;
    e->qsum = in->qsum;
    e->qsqsum = in->qsqsum;
    e->accuracy = in->accuracy_relative * in->two_charge_force;  // ewald.cpp:131-132
    const double q2 = in->qsqsum * in->qqrd2e;                   // kspace.cpp:293
    double g = in->g_ewald;
    if (!(g > 0.0)) {  // ewald.cpp:153-160
      if (e->accuracy <= 0.0) throw StyleError{POLB200_ERR_ARG, "KSpace accuracy must be > 0"};
      if (q2 == 0.0) throw StyleError{POLB200_ERR_ARG, "Must use 'kspace_modify gewald' for uncharged system"};
      g = e->accuracy * sqrt((double)in->natoms * in->cutoff * prd[0] * prd[1] * prd[2]) / (2.0 * q2);
      if (g >= 1.0) g = (1.35 - 0.15 * log(e->accuracy)) / in->cutoff;
      else g = sqrt(-log(g)) / in->cutoff;
    }
    e->g_ewald = g;
    e->volume = prd[0] * prd[1] * prd[2];
    int km[3];
    double gs = 0.0;
    for (int d = 0; d < 3; d++) {  // ewald.cpp:241-275
      e->unitk[d] = 2.0 * pi / prd[d];
      km[d] = 1;
      while (ewald_rms(km[d], prd[d], in->natoms, q2, g) > e->accuracy) km[d]++;
      gs = std::max(gs, e->unitk[d] * e->unitk[d] * km[d] * km[d]);
    }
    e->kxmax = km[0]; e->kymax = km[1]; e->kzmax = km[2];
    e->kmax = std::max(km[0], std::max(km[1], km[2]));
    e->gsqmx = gs * 1.00001;  // ewald.cpp:311
    std::vector<EwaldK> ks;
    e->nk = ewald_build_kset(e, ks);
    e->nquads = (int)ks.size();
    e->slots = ew_row_slots(e->kmax);
    e->kv.ensure(ks.size() + 1);
    e->S.ensure(4 * ks.size() + 4);
    e->out.ensure(8);
    e->h_out.ensure(8);
    if (!ks.empty()) CUDA_CHECK(cudaMemcpy(e->kv.p, ks.data(), ks.size() * sizeof(EwaldK), cudaMemcpyHostToDevice));
    {
      if (e->kmax >= 512) throw StyleError{POLB200_ERR_UNSUPPORTED, "polb200_ewald: kmax >= 512"};
      std::vector<EwaldCol> cols, valid;
      std::vector<double> ug;
      std::vector<int> kxyz;
      const int nkc = ewald_build_columns(e, cols, valid, ug, kxyz);
      if (nkc != e->nk) throw StyleError{POLB200_ERR_STATE, "polb200_ewald: column form and quad form of the k set differ"};
      e->nvalid = (int)valid.size();
      e->cols.ensure(cols.size() + 1); e->vcols.ensure(valid.size() + 1); e->ugv.ensure(ug.size() + 1); e->kxyz.ensure(kxyz.size() + 1);
      e->W.ensure(ug.size() + 1);
      CUDA_CHECK(cudaMemcpy(e->cols.p, cols.data(), cols.size() * sizeof(EwaldCol), cudaMemcpyHostToDevice));
      if (!valid.empty()) {
        CUDA_CHECK(cudaMemcpy(e->vcols.p, valid.data(), valid.size() * sizeof(EwaldCol), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(e->ugv.p, ug.data(), ug.size() * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(e->kxyz.p, kxyz.data(), kxyz.size() * sizeof(int), cudaMemcpyHostToDevice));
      }
      if (const char *v = getenv("POLB200_EWALD_QUADS")) e->use_cols = atoi(v) == 0;
    }
    e->ready = true;
    if (info) {
      info->g_ewald = g;
      info->kxmax = km[0]; info->kymax = km[1]; info->kzmax = km[2];
      info->kmax = e->kmax;
      info->kcount = e->nk;
      info->gsqmx = e->gsqmx;
    }
  });
}

int polb200_ewald_compute(polb200_ewald_t *e, int nlocal, const double *x, const double *q, double *f, int eflag, int vflag,
                          int on_device, double *energy, double virial[6])
{
  if (!e || nlocal < 0 || (nlocal > 0 && (!x || !q || !f))) return POLB200_ERR_ARG;
  return ewald_guarded(e, [&] {
    if (!e->ready) throw StyleError{POLB200_ERR_STATE, "polb200_ewald_init has not been called"};
    if ((eflag / 2) || (vflag / 4)) throw StyleError{POLB200_ERR_UNSUPPORTED, "per-atom KSpace tallies are not implemented on the device"};
    CUDA_CHECK(cudaSetDevice(e->device));
    if (energy) *energy = 0.0;
    if (virial) for (int k = 0; k < 6; k++) virial[k] = 0.0;
    const bool cols = e->use_cols;
    const int n = nlocal, nk = cols ? e->nk : e->nquads;  // quad kernels walk quads of four k-vectors, column kernels the nk real ones
    if (e->qsqsum == 0.0 || nk == 0 || (n == 0 && !e->nccl)) return;  // ewald.cpp:376 (a rank without atoms still joins the sum)
    const double pi = 3.14159265358979323846;
    CUDA_CHECK(cudaEventRecord(e->ev[0], e->stream));
    const double *dx = x, *dq = q;
    double *df = f;
    if (!on_device) {
      e->c_x.ensure((size_t)3 * n); e->c_q.ensure(n); e->c_f.ensure((size_t)3 * n);
      CUDA_CHECK(cudaMemcpyAsync(e->c_x.p, x, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
      CUDA_CHECK(cudaMemcpyAsync(e->c_q.p, q, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
      CUDA_CHECK(cudaMemsetAsync(e->c_f.p, 0, (size_t)3 * n * sizeof(double), e->stream));
      dx = e->c_x.p; dq = e->c_q.p; df = e->c_f.p;
    }
    const int slots = e->slots;
    const size_t s_slots = cols ? (size_t)nk : (size_t)4 * nk;   // complex entries of S
    if (n == 0) CUDA_CHECK(cudaMemsetAsync(e->S.p, 0, s_slots * sizeof(double2), e->stream));
    else {
    e->phase.ensure((size_t)3 * n * slots);
    k_ewald_phase<<<cdiv((long)3 * n, 256), 256, 0, e->stream>>>(n, dx, e->unitk[0], e->unitk[1], e->unitk[2], slots, e->phase.p);
    CUDA_CHECK(cudaGetLastError());
    if (cols) {
      // structure factors, column form: threads = columns, slices of atoms over grid.y until the machine is full
      const int nkx = e->kxhi_max + 1;
      const int NK = std::min(32, 4 * cdiv(nkx, 4));   // kx held in registers per pass (multiples of 4, at most 32)
      const int ncol = (NK <= 16 || NK == 24) ? 2 : 1;  // columns per thread (measured: 20 kx are better served by one + look-ahead)
      const int kblocks = cdiv(e->nvalid, EWC_THREADS * ncol);
      const int smem = 2 * (EWC_TILE * NK + 2 * EWC_TILE * slots + EWC_TILE) * (int)sizeof(double2);   // two tile buffers
      int slices = 1;
      for (int kxbase = 0; kxbase < nkx; kxbase += 32) {
#define SF(NKX, CO)                                                                                                          \
  do {                                                                                                                       \
    auto kern = k_ewald_sfac_col<NKX, CO>;                                                                                   \
    static int set = 0;                                                                                                      \
    if (smem > set) {                                                                                                        \
      CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                             \
      set = smem;                                                                                                            \
    }                                                                                                                        \
    /* slices of atoms over grid.y.  The columns are sorted by length, so the CTAs of a slice differ in work (a CTA's    */ \
    /* four warps do not: no waiting at the tile barrier): several waves of CTAs let the hardware scheduler level that   */ \
    /* out; for small systems exactly one wave (a partial second wave cost 40 %)                                         */ \
    int occ = 1;                                                                                                             \
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, EWC_THREADS, smem));                                \
    const int wave = std::max(1, occ) * e->num_sms;                                                                          \
    slices = std::max(1, wave / kblocks);                                                                                    \
    if (n >= 8 * EWC_TILE * 4 * slices) slices *= 4;                                                                         \
    e->Spart.ensure((size_t)(slices + cdiv(slices, 16)) * nk + 4);                                                           \
    kern<<<dim3(kblocks, slices), EWC_THREADS, smem, e->stream>>>(n, e->nvalid, nk, e->vcols.p, dq, e->phase.p, slots, kxbase, \
                                                                  e->Spart.p);                                                \
  } while (0)
        switch (NK) {
          case 4: SF(4, 2); break;
          case 8: SF(8, 2); break;
          case 12: SF(12, 2); break;
          case 16: SF(16, 2); break;
          case 20: SF(20, 1); break;
          case 24: SF(24, 2); break;
          case 28: SF(28, 1); break;
          default: SF(32, 1); break;
        }
#undef SF
        CUDA_CHECK(cudaGetLastError());
        e->launches++;
      }
      // slices -> groups of 16 -> S, fixed order
      const int ngroups = cdiv(slices, 16);
      double2 *gsum = e->Spart.p + (size_t)slices * nk;
      k_ewald_sum_groups<<<dim3(cdiv(nk, 128), ngroups), 128, 0, e->stream>>>(nk, slices, 16, e->Spart.p, gsum);
      CUDA_CHECK(cudaGetLastError());
      k_ewald_sum_groups<<<dim3(cdiv(nk, 128), 1), 128, 0, e->stream>>>(nk, ngroups, ngroups, gsum, e->S.p);
      CUDA_CHECK(cudaGetLastError());
      e->launches += 2;
    } else {

    // structure factors: k-vectors x atom slices (enough CTAs to fill the machine, few enough atomics)
    const int kblocks = cdiv(nk, EW_KTHREADS);
    int slices = std::max(1, std::min(cdiv(n, 4 * EW_TILE), cdiv(4 * 148, kblocks)));
    const int sfac_smem = (3 * EW_TILE * slots + EW_TILE) * (int)sizeof(double2);
    if (sfac_smem > e->sfac_smem_set) {
      CUDA_CHECK(cudaFuncSetAttribute(k_ewald_sfac, cudaFuncAttributeMaxDynamicSharedMemorySize, sfac_smem));
      e->sfac_smem_set = sfac_smem;
    }
    e->Spart.ensure((size_t)slices * 4 * nk + 4);
    k_ewald_sfac<<<dim3(kblocks, slices), EW_KTHREADS, sfac_smem, e->stream>>>(n, nk, e->kv.p, dq, e->phase.p, slots, e->Spart.p);
    CUDA_CHECK(cudaGetLastError());
    k_ewald_sum_slices<<<cdiv((long)4 * nk, 256), 256, 0, e->stream>>>(4 * nk, slices, e->Spart.p, e->S.p);
    CUDA_CHECK(cudaGetLastError());
    e->launches++;
    }
    }
    // decomposed run: S(k) = sum over the ranks of their atoms' contributions (the reference's MPI_Allreduce of
    // sfacrl / sfacim, ewald.cpp:395-400)
    if (e->nccl) NCCL_CHECK(g_nccl.AllReduce(e->S.p, e->S.p, 2 * s_slots, ncclDouble, ncclSum, e->nccl, e->stream));
    // forces: one thread per atom with its phase rows in shared memory
    if (n > 0 && cols) {
      const int side = 2 * e->kmax + 1;
      const int smem = side * (e->kmax + 1) * (int)sizeof(double2) + side * (int)sizeof(EwaldCol);
      if (smem > 200 * 1024) throw StyleError{POLB200_ERR_UNSUPPORTED, "Ewald kmax too large for the device force kernel"};
      if (smem > e->force_smem_set) {
        CUDA_CHECK(cudaFuncSetAttribute(k_ewald_force_col, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        e->force_smem_set = smem;
      }
      k_ewald_w<<<cdiv(nk, 256), 256, 0, e->stream>>>(nk, e->ugv.p, e->S.p, e->W.p);
      CUDA_CHECK(cudaGetLastError());
      // small systems: the ky rows are split over grid.y until ~16 warps per SM are in flight
      const int athreads = EWC_THREADS;
      const int ablocks = cdiv(n, athreads);
      const int fslices = std::max(1, std::min(side, (4 * e->num_sms) / ablocks));
      if (fslices > 1) e->fpart.ensure((size_t)fslices * 3 * n);
      k_ewald_force_col<<<dim3(ablocks, fslices), athreads, smem, e->stream>>>(n, e->kmax, e->cols.p, e->W.p, dq, e->phase.p, slots,
                                                                               e->unitk[0], e->unitk[1], e->unitk[2], e->qqrd2e, df, e->fpart.p);
      CUDA_CHECK(cudaGetLastError());
      if (fslices > 1) {
        k_ewald_force_sum<<<cdiv((long)3 * n, 256), 256, 0, e->stream>>>((long)3 * n, fslices, e->fpart.p, df);
        CUDA_CHECK(cudaGetLastError());
        e->launches++;
      }
      e->launches += 3;
    } else if (n > 0) {
    int athreads = EW_ATHREADS;
    const int tile_b = EW_KTILE * (int)(sizeof(EwaldK) + 4 * sizeof(double2));
    while (athreads > 32 && athreads * 3 * slots * (int)sizeof(double2) + tile_b > 200 * 1024) athreads -= 32;
    const int force_smem = athreads * 3 * slots * (int)sizeof(double2) + tile_b;
    if (force_smem > 220 * 1024) throw StyleError{POLB200_ERR_UNSUPPORTED, "Ewald kmax too large for the device force kernel"};
    if (force_smem > e->force_smem_set) {
      CUDA_CHECK(cudaFuncSetAttribute(k_ewald_force, cudaFuncAttributeMaxDynamicSharedMemorySize, force_smem));
      e->force_smem_set = force_smem;
    }
    k_ewald_force<<<cdiv(n, athreads), athreads, force_smem, e->stream>>>(n, nk, e->kv.p, e->S.p, dq, e->phase.p, slots, e->unitk[0],
                                                                          e->unitk[1], e->unitk[2], e->qqrd2e, df);
    CUDA_CHECK(cudaGetLastError());
    e->launches += 3;
    }
    const bool ev = (eflag & 1) || (vflag % 4);
    if (ev) {
      if (cols)
        k_ewald_energy_col<<<1, 256, 0, e->stream>>>(nk, e->ugv.p, e->kxyz.p, e->S.p, e->unitk[0], e->unitk[1], e->unitk[2],
                                                     1.0 / (e->g_ewald * e->g_ewald), e->out.p);
      else
      k_ewald_energy<<<1, 256, 0, e->stream>>>(nk, e->kv.p, e->S.p, e->unitk[0], e->unitk[1], e->unitk[2],
                                               1.0 / (e->g_ewald * e->g_ewald), e->out.p);
      CUDA_CHECK(cudaGetLastError());
      e->launches++;
      CUDA_CHECK(cudaMemcpyAsync(e->h_out.p, e->out.p, 7 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    }
    if (!on_device && n > 0) {
      e->h_f.ensure((size_t)3 * n);
      CUDA_CHECK(cudaMemcpyAsync(e->h_f.p, e->c_f.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    }
    CUDA_CHECK(cudaEventRecord(e->ev[1], e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    cudaEventElapsedTime(&e->ms_last, e->ev[0], e->ev[1]);
    if (!on_device)
      for (size_t k = 0; k < (size_t)3 * n; k++) f[k] += e->h_f.p[k];
    if (ev) {
      if ((eflag & 1) && energy) {  // ewald.cpp:455-462
        double en = e->h_out.p[0];
        en -= e->g_ewald * e->qsqsum / sqrt(pi) + 0.5 * pi * e->qsum * e->qsum / (e->g_ewald * e->g_ewald * e->volume);
        *energy = en * e->qqrd2e / e->nranks;   // every rank holds the global sums: per-rank partials that add up (thermo all-reduces)
      }
      if ((vflag % 4) && virial)
        for (int k = 0; k < 6; k++) virial[k] = e->h_out.p[1 + k] * e->qqrd2e / e->nranks;  // ewald.cpp:466-474
    }
  });
}

double polb200_ewald_last_ms(const polb200_ewald_t *e) { return e ? (double)e->ms_last : 0.0; }

int polb200_ewald_comm_init(polb200_ewald_t *e, int rank, int nranks, const void *id_bytes)
{
  if (!e || !id_bytes || nranks < 1 || rank < 0 || rank >= nranks) return POLB200_ERR_ARG;
  return ewald_guarded(e, [&] {
    if (e->nccl) throw StyleError{POLB200_ERR_STATE, "polb200_ewald_comm_init was already called"};
    std::string err;
    if (!g_nccl.load(err)) throw StyleError{POLB200_ERR_UNSUPPORTED, err};
    CUDA_CHECK(cudaSetDevice(e->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    NCCL_CHECK(g_nccl.CommInitRank(&e->nccl, nranks, id, rank));
    e->rank = rank;
    e->nranks = nranks;
  });
}

}  // extern "C"
