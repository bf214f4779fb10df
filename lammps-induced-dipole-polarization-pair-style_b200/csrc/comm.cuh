// comm.cuh -- multi-GPU layer: spatial decomposition of the box into bricks, one process
// per GPU.  Included by engine.cu after the handle definition.
//
//   * at a rebuild: send lists (which owned atoms are ghosts of which neighbour brick), counts, then
//     positions / dipoles / type+molecule+tag of the shell, all neighbours in ONE grouped NCCL step;
//   * every step: ghost positions once;
//   * every SCF sweep: ghost dipoles.  Two realisations:
//       - NCCL halo: pack kernel -> grouped ncclSend/ncclRecv -> unpack kernel;
//       - peer push (default when the peers' buffers can be mapped): the sweep kernel itself stores
//         each new dipole into the ghost slots of the neighbour bricks through NVLink peer memory
//         (cudaIpc mappings), followed by one tiny signal/wait kernel that is the inter-GPU barrier
//         and the convergence all-reduce in one (no NCCL call, no pack/unpack, no host sync);
//   * convergence / rmin / rebuild trigger: all-reduces of one or two scalars.
//
// Energies and the virial are returned per rank (LAMMPS convention: thermo all-reduces them).
#pragma once

namespace polb200 {

static NcclApi g_nccl;

#define NCCL_CHECK(expr)                                                                                   \
  do {                                                                                                     \
    ncclResult_t _r = (expr);                                                                              \
    if (_r != ncclSuccess)                                                                                 \
      throw CudaError{std::string(#expr) + ": " + g_nccl.GetErrorString(_r) + " (" + __FILE__ + ":" +     \
                      std::to_string(__LINE__) + ")"};                                                     \
  } while (0)

static void comm_setup_geom(polb200_handle *h)
{
  CommState &c = h->comm;
  if (!c.active || !h->box_set) return;
  if (make_plan(c.nranks, c.rank, c.pg, h->box.periodic, h->box.lo, h->box.hi, c.plan))
    throw StyleError{POLB200_ERR_ARG, "Bad grid of processors"};
  SendGeom &G = c.geom;
  G.valid = 0;
  for (int k = 0; k < 3; k++) {
    G.lo[k] = c.plan.sublo[k];
    G.hi[k] = c.plan.subhi[k];
  }
  for (int d = 0; d < NDIR; d++) {
    if (c.plan.dest[d] >= 0) G.valid |= 1u << d;
    int code = 0;
    for (int k = 0; k < 3; k++) {
      G.shift[d][k] = c.plan.wrap[d][k] * h->box.prd[k];
      code |= (c.plan.wrap[d][k] + 1) << (2 * k);
    }
    G.code[d] = code;
  }
  c.geom_valid = true;
  h->have_lists = false;
}

// one grouped exchange of per-slot records of `eb` bytes: send segment d goes to dest[d], what arrives
// from src[d] lands in receive segment d.  Messages between the same two ranks are matched by issue
// order, which is the direction order on both sides (decomp.h).
static void comm_exchange(polb200_handle *h, const void *sbuf, void *rbuf, size_t eb)
{
  CommState &c = h->comm;
  const char *s = static_cast<const char *>(sbuf);
  char *r = static_cast<char *>(rbuf);
  bool any_remote = false;
  for (int d = 0; d < NDIR; d++)
    if (d != DIR_SELF && c.plan.dest[d] != c.rank && (c.send_cnt[d] || c.recv_cnt[d])) any_remote = true;
  if (any_remote) NCCL_CHECK(g_nccl.GroupStart());
  for (int d = 0; d < NDIR; d++) {
    if (d == DIR_SELF) continue;
    if (c.plan.dest[d] == c.rank) {  // periodic image of this brick itself (one brick along that dimension)
      if (c.send_cnt[d])
        CUDA_CHECK(cudaMemcpyAsync(r + (size_t)c.recv_off[d] * eb, s + (size_t)c.send_off[d] * eb,
                                   (size_t)c.send_cnt[d] * eb, cudaMemcpyDeviceToDevice, h->stream));
      continue;
    }
    if (c.send_cnt[d])
      NCCL_CHECK(g_nccl.Send(s + (size_t)c.send_off[d] * eb, (size_t)c.send_cnt[d] * eb, ncclChar, c.plan.dest[d],
                             c.nccl, h->stream));
    if (c.recv_cnt[d])
      NCCL_CHECK(g_nccl.Recv(r + (size_t)c.recv_off[d] * eb, (size_t)c.recv_cnt[d] * eb, ncclChar, c.plan.src[d],
                             c.nccl, h->stream));
  }
  if (any_remote) NCCL_CHECK(g_nccl.GroupEnd());
}

// reverse direction: per RECEIVE slot records travel back to the rank that sent the slot
static void comm_exchange_reverse(polb200_handle *h, const void *rbuf, void *sbuf, size_t eb)
{
  CommState &c = h->comm;
  const char *r = static_cast<const char *>(rbuf);
  char *s = static_cast<char *>(sbuf);
  bool any_remote = false;
  for (int d = 0; d < NDIR; d++)
    if (d != DIR_SELF && c.plan.dest[d] != c.rank && (c.send_cnt[d] || c.recv_cnt[d])) any_remote = true;
  if (any_remote) NCCL_CHECK(g_nccl.GroupStart());
  for (int d = 0; d < NDIR; d++) {
    if (d == DIR_SELF) continue;
    if (c.plan.dest[d] == c.rank) {
      if (c.send_cnt[d])
        CUDA_CHECK(cudaMemcpyAsync(s + (size_t)c.send_off[d] * eb, r + (size_t)c.recv_off[d] * eb,
                                   (size_t)c.send_cnt[d] * eb, cudaMemcpyDeviceToDevice, h->stream));
      continue;
    }
    if (c.recv_cnt[d])
      NCCL_CHECK(g_nccl.Send(r + (size_t)c.recv_off[d] * eb, (size_t)c.recv_cnt[d] * eb, ncclChar, c.plan.src[d],
                             c.nccl, h->stream));
    if (c.send_cnt[d])
      NCCL_CHECK(g_nccl.Recv(s + (size_t)c.send_off[d] * eb, (size_t)c.send_cnt[d] * eb, ncclChar, c.plan.dest[d],
                             c.nccl, h->stream));
  }
  if (any_remote) NCCL_CHECK(g_nccl.GroupEnd());
}

static void comm_allreduce(polb200_handle *h, void *dptr, size_t count, ncclDataType_t t, ncclRedOp_t op)
{
  NCCL_CHECK(g_nccl.AllReduce(dptr, dptr, count, t, op, h->comm.nccl, h->stream));
}

// Collective agreement on an error: every brick contributes its local error code (0 = fine), all get the largest.
// A brick that fails alone -- before or between collectives -- would leave the others blocked in NCCL or spinning in the
// peer barrier until the timeout; with this every brick throws the same error at the same point instead.
enum { COMM_OK = 0, COMM_ERR_EMPTY = 1, COMM_ERR_OVERFLOW = 2, COMM_ERR_ALLOC = 3, COMM_ERR_SUBDOMAIN = 4 };
static int comm_agree(polb200_handle *h, int local_err)
{
  h->flags.ensure(8);
  int *flag = h->flags.p + 5;
  int v = local_err;
  CUDA_CHECK(cudaMemcpyAsync(flag, &v, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  comm_allreduce(h, flag, 1, ncclInt, ncclMax);
  CUDA_CHECK(cudaMemcpyAsync(&v, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return v;
}

static void comm_throw(int err)
{
  switch (err) {
    case COMM_OK: return;
    case COMM_ERR_EMPTY:
      throw StyleError{POLB200_ERR_UNSUPPORTED, "a brick of the decomposition owns no atoms"};
    case COMM_ERR_OVERFLOW:
      throw StyleError{POLB200_ERR_OVERFLOW, "owned+ghost atoms of a brick exceed the 30-bit neighbor index"};
    case COMM_ERR_SUBDOMAIN:
      throw StyleError{POLB200_ERR_UNSUPPORTED, "neighbor cutoff exceeds the sub-domain length of a brick"};
    default:
      throw CudaError{"a brick of the decomposition could not allocate its device arrays (cudaMalloc failed)"};
  }
}

static void comm_close_peers(polb200_handle *h)
{
  CommState &c = h->comm;
  for (int r = 0; r < MAX_PEERS; r++)
    for (int k = 0; k < NPEERBUF; k++) {
      if (c.peer_ptr[k][r] && r != c.rank) cudaIpcCloseMemHandle(c.peer_ptr[k][r]);
      c.peer_ptr[k][r] = nullptr;
    }
  c.push.enabled = 0;
}

// Map the peers' dipole buffers (both parities) and arrival counters into this process.
// Collective: every rank calls it in the same rebuild.  On any failure the push is disabled on ALL ranks
// (the decision is all-reduced) and the NCCL halo is used.
static void comm_map_peers(polb200_handle *h)
{
  CommState &c = h->comm;
  comm_close_peers(h);
  c.mapped_ptr[0] = h->mua.p;
  c.mapped_ptr[1] = h->mub.p;
  c.mapped_ptr[2] = c.flags.p;
  c.mapped_ptr[3] = h->xq.p;
  int ok = c.want_push && c.nranks <= MAX_PEERS && !c.shared_device ? 1 : 0;
  cudaIpcMemHandle_t mine[NPEERBUF];
  memset(mine, 0, sizeof(mine));
  if (ok) {
    for (int k = 0; k < NPEERBUF; k++)
      if (cudaIpcGetMemHandle(&mine[k], c.mapped_ptr[k]) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
      }
  }
  const size_t hb = sizeof(mine);
  c.ipc_dev.ensure(hb * (size_t)(c.nranks + 1));
  std::vector<char> all(hb * c.nranks);
  CUDA_CHECK(cudaMemcpyAsync(c.ipc_dev.p + hb * c.nranks, mine, hb, cudaMemcpyHostToDevice, h->stream));
  NCCL_CHECK(g_nccl.AllGather(c.ipc_dev.p + hb * c.nranks, c.ipc_dev.p, hb, ncclChar, c.nccl, h->stream));
  CUDA_CHECK(cudaMemcpyAsync(all.data(), c.ipc_dev.p, hb * c.nranks, cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (ok) {
    for (int r = 0; r < c.nranks && ok; r++) {
      for (int k = 0; k < NPEERBUF && ok; k++) {
        if (r == c.rank) {
          c.peer_ptr[k][r] = c.mapped_ptr[k];
          continue;
        }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, all.data() + hb * r + sizeof(hd) * k, sizeof(hd));
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          ok = 0;
        } else
          c.peer_ptr[k][r] = p;
      }
    }
  }
  // unanimous decision
  int *flag = h->flags.p + 2;
  CUDA_CHECK(cudaMemcpyAsync(flag, &ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  comm_allreduce(h, flag, 1, ncclInt, ncclMin);
  CUDA_CHECK(cudaMemcpyAsync(&ok, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (!ok) {
    comm_close_peers(h);
    return;
  }
  for (int r = 0; r < c.nranks; r++) {
    c.push.mu[0][r] = static_cast<double4 *>(c.peer_ptr[0][r]);
    c.push.mu[1][r] = static_cast<double4 *>(c.peer_ptr[1][r]);
    c.push.flag[r] = static_cast<unsigned long long *>(c.peer_ptr[2][r]);
    c.push.xq[r] = static_cast<double4 *>(c.peer_ptr[3][r]);
  }
  c.push.enabled = 1;
}

// Ghost shell of this brick at a rebuild.  On entry the owned atoms are cell-sorted in xq/mua/tm/tag
// [0,n).  On return nghost is known, the ext arrays hold the ghosts cell-sorted behind the owned atoms
// and cg_start indexes them by cell.
static void comm_build_ghosts(polb200_handle *h, int n)
{
  CommState &c = h->comm;
  const HostStyle &st = h->style;
  const Grid &g = h->P.grid;
  if (!c.geom_valid) comm_setup_geom(h);
  {
    // first collective of every rebuild: local preconditions, agreed by all bricks (a brick that was handed no atoms
    // takes part in this agreement from compute_impl and nothing else)
    int err = COMM_OK;
    for (int k = 0; k < 3; k++) {
      const double len = c.plan.subhi[k] - c.plan.sublo[k];
      if ((h->box.periodic[k] || c.pg[k] > 1) && st.cutneighmax + h->atom_slack > len) err = COMM_ERR_SUBDOMAIN;
    }
    comm_throw(comm_agree(h, err));
  }
  c.geom.cut = st.cutneighmax + h->atom_slack;

  // 1. send lists
  h->cnt.ensure(n + 1);
  h->rowstart.ensure(n + 2);
  LAUNCH(h, k_send_count, cdiv(n, 256), 256, n, h->xq.p, c.geom, h->cnt.p);
  CUDA_CHECK(cudaMemsetAsync(h->cnt.p + n, 0, sizeof(unsigned long long), h->stream));
  exclusive_sum(h, n + 1, h->cnt.p, h->rowstart.p);
  unsigned long long ns64 = 0;
  int hflags[4];
  CUDA_CHECK(cudaMemcpyAsync(&ns64, h->rowstart.p + n, sizeof(ns64), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaMemcpyAsync(hflags, h->flags.p, sizeof(hflags), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  const int nan_local = hflags[0] & 1;
  h->P.pc.has_molecules = hflags[3] ? 1 : 0;
  const int ns = (int)ns64;
  c.nsend = ns;
  h->push_off.ensure(n + 2);
  CUDA_CHECK(cudaMemcpyAsync(h->push_off.p, h->rowstart.p, (size_t)(n + 1) * sizeof(unsigned long long),
                             cudaMemcpyDeviceToDevice, h->stream));
  const int nmax = std::max(std::max(n, ns), 64);
  h->keys.ensure(nmax); h->keys2.ensure(nmax); h->vals.ensure(nmax); h->vals2.ensure(nmax);
  c.send_owner_u.ensure(ns + 1); c.send_owner.ensure(ns + 1); c.send_dir.ensure(ns + 1); c.slot_of_u.ensure(ns + 1);
  c.dir_start.ensure(NDIR + 2);
  int dstart[NDIR + 1];
  memset(dstart, 0, sizeof(dstart));
  if (ns > 0) {
    LAUNCH(h, k_send_fill, cdiv(n, 256), 256, n, h->xq.p, c.geom, h->rowstart.p, c.send_owner_u.p, h->keys.p, h->vals.p);
    sort_pairs(h, ns, h->keys.p, c.send_dir.p, h->vals.p, h->vals2.p, 5);  // stable: owners ascending per direction
    LAUNCH(h, k_gather_int, cdiv(ns, 256), 256, ns, h->vals2.p, c.send_owner_u.p, c.send_owner.p);
    LAUNCH(h, k_invert_perm, cdiv(ns, 256), 256, ns, h->vals2.p, c.slot_of_u.p);
    LAUNCH(h, k_cell_starts, 1, 64, NDIR, ns, c.send_dir.p, c.dir_start.p, 0);
    CUDA_CHECK(cudaMemcpyAsync(dstart, c.dir_start.p, sizeof(dstart), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  }
  // 2. counts of every rank (+ owned atoms, NaN flag, "my buffers moved" flag in the spare slots)
  constexpr int NV = 32;
  int mine[NV];
  memset(mine, 0, sizeof(mine));
  for (int d = 0; d < NDIR; d++) {
    c.send_off[d] = dstart[d];
    c.send_cnt[d] = dstart[d + 1] - dstart[d];
    mine[d] = c.send_cnt[d];
  }
  mine[DIR_SELF] = 0;
  mine[27] = n;
  mine[28] = nan_local;
  c.counts_dev.ensure((size_t)NV * (c.nranks + 1));
  std::vector<int> all((size_t)NV * c.nranks);
  CUDA_CHECK(cudaMemcpyAsync(c.counts_dev.p + (size_t)NV * c.nranks, mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
  NCCL_CHECK(g_nccl.AllGather(c.counts_dev.p + (size_t)NV * c.nranks, c.counts_dev.p, NV, ncclInt, c.nccl, h->stream));
  CUDA_CHECK(cudaMemcpyAsync(all.data(), c.counts_dev.p, all.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  long nglobal = 0;
  int nan_any = 0;
  for (int r = 0; r < c.nranks; r++) {
    nglobal += all[(size_t)NV * r + 27];
    nan_any |= all[(size_t)NV * r + 28];
    c.nloc_of[r] = all[(size_t)NV * r + 27];
  }
  if (nan_any) throw StyleError{POLB200_ERR_NAN, "Non-numeric positions - simulation unstable"};
  c.nglobal = nglobal;
  int nrecv = 0;
  for (int d = 0; d < NDIR; d++) {
    c.recv_off[d] = nrecv;
    c.recv_cnt[d] = (d != DIR_SELF && c.plan.src[d] >= 0) ? all[(size_t)NV * c.plan.src[d] + d] : 0;
    nrecv += c.recv_cnt[d];
  }
  c.nrecv = nrecv;
  const int overflow = ((long)n + nrecv >= (1l << 30)) ? 1 : 0;  // agreed below, together with `moved`
  const int ng = nrecv;
  h->nghost = ng;
  const size_t next = (size_t)n + ng;

  // 3. payloads.  Peers map each other's dipole arrays (cudaIpc) for the fused sweep + push, and nobody may
  //    free an array that a peer still has mapped: the decision to (re)allocate is taken collectively, the
  //    mappings are closed first, and arrays replaced meanwhile wait in the graveyard until then.
  c.flags.ensure(4 * MAX_PEERS);
  if (!c.flags_zeroed) {
    CUDA_CHECK(cudaMemsetAsync(c.flags.p, 0, c.flags.cap * sizeof(unsigned long long), h->stream));
    c.flags_zeroed = true;
  }
  int moved = (next > h->mua.cap || next > h->mub.cap || next > h->xq.cap || c.mapped_ptr[0] != h->mua.p ||
               c.mapped_ptr[1] != h->mub.p || c.mapped_ptr[2] != c.flags.p || c.mapped_ptr[3] != h->xq.p) ? 1 : 0;
  int *flag = h->flags.p + 2;
  moved += 2 * overflow;  // one all-reduce carries both
  CUDA_CHECK(cudaMemcpyAsync(flag, &moved, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  comm_allreduce(h, flag, 1, ncclInt, ncclMax);
  CUDA_CHECK(cudaMemcpyAsync(&moved, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (moved >= 2) comm_throw(COMM_ERR_OVERFLOW);
  if (moved) {
    comm_close_peers(h);
    comm_allreduce(h, flag, 1, ncclInt, ncclMax);  // barrier: every rank has closed its mappings
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    for (void *q : c.graveyard) cudaFree(q);
    c.graveyard.clear();
  }
  const int nmax2 = std::max(nmax, ng);
  int alloc_err = COMM_OK;
  try {
    grow_ext(h, n, next);
    if (moved) {
      for (void *q : c.graveyard) cudaFree(q);  // replaced just now, and no peer maps anything at this point
      c.graveyard.clear();
    }
    c.sbuf.ensure(ns + 1); c.rbuf.ensure(ng + 1);
    c.sbufi.ensure(ns + 1); c.rbufi.ensure(ng + 1);
    c.gslot.ensure(ng + 1);
    h->g_shift.ensure(ng + 1);
    h->keys.ensure(nmax2); h->keys2.ensure(nmax2); h->vals.ensure(nmax2); h->vals2.ensure(nmax2);
    h->cg_start.ensure(g.ncell + 2);
  } catch (const CudaError &) {
    cudaGetLastError();
    alloc_err = COMM_ERR_ALLOC;
  }
  comm_throw(comm_agree(h, alloc_err));  // a brick out of memory must not leave the others in the exchange below
  if (ns) LAUNCH(h, k_pack_pos, cdiv(ns, 256), 256, ns, c.send_owner.p, c.send_dir.p, c.geom, h->xq.p, c.sbuf.p);
  comm_exchange(h, c.sbuf.p, c.rbuf.p, sizeof(double4));
  if (ng) {
    LAUNCH(h, k_ghost_keys, cdiv(ng, 256), 256, ng, c.rbuf.p, g, h->keys.p, h->vals.p);
    sort_pairs(h, ng, h->keys.p, h->keys2.p, h->vals.p, h->vals2.p, bits_for(g.ncell + 1) + g.xbits);
    CUDA_CHECK(cudaMemcpyAsync(c.gslot.p, h->vals2.p, (size_t)ng * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    LAUNCH(h, k_cell_starts, cdiv(g.ncell + 1, 256), 256, g.ncell, ng, h->keys2.p, h->cg_start.p, g.xbits);
    LAUNCH(h, k_unpack_rec, cdiv(ng, 256), 256, ng, c.gslot.p, c.rbuf.p, h->xq.p + n);
  } else {
    CUDA_CHECK(cudaMemsetAsync(h->cg_start.p, 0, (g.ncell + 2) * sizeof(int), h->stream));
  }
  if (ns) LAUNCH(h, k_pack_rec, cdiv(ns, 256), 256, ns, c.send_owner.p, h->mua.p, c.sbuf.p);
  comm_exchange(h, c.sbuf.p, c.rbuf.p, sizeof(double4));
  if (ng) LAUNCH(h, k_unpack_rec, cdiv(ng, 256), 256, ng, c.gslot.p, c.rbuf.p, h->mua.p + n);
  if (ns) LAUNCH(h, k_pack_meta, cdiv(ns, 256), 256, ns, c.send_owner.p, c.send_dir.p, c.geom, h->tm.p, h->tag.p, c.sbufi.p);
  comm_exchange(h, c.sbufi.p, c.rbufi.p, sizeof(int4));
  if (ng) LAUNCH(h, k_unpack_meta, cdiv(ng, 256), 256, ng, c.gslot.p, c.rbufi.p, h->tm.p + n, h->tag.p + n, h->g_shift.p);

  // 4. peer push tables: for every send slot the ext index it occupies on its destination rank,
  //    returned by the receivers, stored per owned atom (CSR push_off in (owner, direction) order)
  if (moved) comm_map_peers(h);
  h->push_ptr0.ensure(ns + 1); h->push_ptr1.ensure(ns + 1); c.push_ptrx.ensure(ns + 1); c.dir_of_u.ensure(ns + 1);
  h->push_ready = false;
  if (ng) LAUNCH(h, k_ghost_ext_index, cdiv(ng, 256), 256, ng, n, c.gslot.p, (int *)c.rbufi.p);  // per recv slot
  comm_exchange_reverse(h, c.rbufi.p, c.sbufi.p, sizeof(int));
  if (c.push.enabled) {
    DirTable T;
    for (int d = 0; d < NDIR; d++) T.v[d] = c.plan.dest[d];
    if (ns)
      LAUNCH(h, k_push_tables, cdiv(ns, 256), 256, ns, c.slot_of_u.p, c.send_dir.p, T, (const int *)c.sbufi.p, c.push,
             h->push_ptr0.p, h->push_ptr1.p, c.push_ptrx.p, c.dir_of_u.p);
    h->push_ready = true;
  }
}

// exclusion-rule bits of the ghosts: the owners' bits travel like any other per-slot record (rebuild only)
static void comm_ghost_exbits(polb200_handle *h, int n, int ng)
{
  CommState &c = h->comm;
  const int ns = c.nsend;
  static_assert(sizeof(int2) <= sizeof(int4), "exchange buffers are sized for int4 records");
  if (ns) LAUNCH(h, k_pack_int2, cdiv(ns, 256), 256, ns, c.send_owner.p, h->exb.p, reinterpret_cast<int2 *>(c.sbufi.p));
  comm_exchange(h, c.sbufi.p, c.rbufi.p, sizeof(int2));
  if (ng) LAUNCH(h, k_unpack_int2, cdiv(ng, 256), 256, ng, c.gslot.p, reinterpret_cast<const int2 *>(c.rbufi.p), h->exb.p + n);
}

// inter-GPU barrier of the peer-push path, optionally carrying this rank's partial squared change to
// everybody: afterwards *change_inout holds the global sum (ranks added in rank order: identical everywhere)
static void comm_signal_wait(polb200_handle *h, double *change_inout)
{
  CommState &c = h->comm;
  c.epoch++;
  LAUNCH(h, k_signal_wait, 1, 32, c.rank, c.nranks, c.epoch, c.push, change_inout, c.barrier_timeout_ns, h->flags.p + 6, h->scf_stop);
}

// ghost refresh: positions (shifted) and/or one dipole array (mua or mub).
//   peer push: one store kernel per array straight into the neighbours' ghost slots + one barrier kernel.
//     `fence_before`: the targets may still be read by the peers' previous step (positions at the start of a
//     step), so a barrier also precedes the stores;
//   otherwise: pack -> grouped NCCL send/recv -> unpack.
static void comm_refresh(polb200_handle *h, bool pos, double4 *mu, bool fence_before)
{
  CommState &c = h->comm;
  const int n = h->nloc, ng = h->nghost, ns = c.nsend;
  if (h->push_ready && c.push.enabled && (mu == nullptr || mu == h->mua.p || mu == h->mub.p)) {
    if (fence_before) comm_signal_wait(h, nullptr);
    if (pos && ns) LAUNCH(h, k_push_pos, cdiv(ns, 256), 256, ns, c.send_owner_u.p, c.dir_of_u.p, c.geom, h->xq.p, c.push_ptrx.p);
    if (mu && ns) LAUNCH(h, k_push_rec, cdiv(ns, 256), 256, ns, c.send_owner_u.p, mu, mu == h->mub.p ? h->push_ptr1.p : h->push_ptr0.p, h->scf_stop);
    comm_signal_wait(h, nullptr);
    return;
  }
  if (pos) {
    if (ns) LAUNCH(h, k_pack_pos, cdiv(ns, 256), 256, ns, c.send_owner.p, c.send_dir.p, c.geom, h->xq.p, c.sbuf.p);
    comm_exchange(h, c.sbuf.p, c.rbuf.p, sizeof(double4));
    if (ng) LAUNCH(h, k_unpack_rec, cdiv(ng, 256), 256, ng, c.gslot.p, c.rbuf.p, h->xq.p + n);
  }
  if (mu) {
    if (ns) LAUNCH(h, k_pack_rec, cdiv(ns, 256), 256, ns, c.send_owner.p, mu, c.sbuf.p);
    comm_exchange(h, c.sbuf.p, c.rbuf.p, sizeof(double4));
    if (ng) LAUNCH(h, k_unpack_rec, cdiv(ng, 256), 256, ng, c.gslot.p, c.rbuf.p, mu + n);
  }
}

}  // namespace polb200
