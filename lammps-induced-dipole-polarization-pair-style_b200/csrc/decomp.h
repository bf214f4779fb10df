// decomp.h -- host-side plan of the spatial decomposition: a 3-D brick grid of GPUs over
// the orthogonal periodic box, each brick exchanging its boundary shell with its <= 26 neighbour
// bricks in ONE grouped step (NVSwitch is uniform, so there is no reason for the three staged
// sweeps of CommBrick::borders, src/comm_brick.cpp:712-880).
//
// Plain host C++: no CUDA, no NCCL.  The same plan drives the NCCL halo (comm.cuh), the peer-memory
// push of the sweep kernel and the gloo-based CPU test of the decomposition (tests/test_decomp_gloo.py).
#pragma once

namespace polb200 {

constexpr int NDIR = 27;      // (dz+1)*9 + (dy+1)*3 + (dx+1); slot 13 (0,0,0) is unused
constexpr int DIR_SELF = 13;

struct DecompPlan {
  int nranks, rank;
  int pg[3];        // bricks per dimension
  int coords[3];    // this rank's brick, rank = (cz*pg[1] + cy)*pg[0] + cx
  int dest[NDIR];   // rank that receives what this rank sends in direction d (-1: no such neighbour)
  int src[NDIR];    // rank whose direction-d message this rank receives (-1: none)
  int wrap[NDIR][3];  // periodic image shift (units of the box length) the SENDER adds to coordinates
  int rwrap[NDIR][3]; // the shift that was applied to what arrives as direction d
  double sublo[3], subhi[3];
};

inline void dir_vec(int d, int v[3])
{
  v[0] = d % 3 - 1;
  v[1] = (d / 3) % 3 - 1;
  v[2] = d / 9 - 1;
}

// returns 0, or 1 when pg does not multiply to nranks / rank is out of range
inline int make_plan(int nranks, int rank, const int pg[3], const int periodic[3], const double boxlo[3],
                     const double boxhi[3], DecompPlan &p)
{
  if (pg[0] < 1 || pg[1] < 1 || pg[2] < 1 || pg[0] * pg[1] * pg[2] != nranks || rank < 0 || rank >= nranks) return 1;
  p.nranks = nranks;
  p.rank = rank;
  for (int k = 0; k < 3; k++) p.pg[k] = pg[k];
  p.coords[0] = rank % pg[0];
  p.coords[1] = (rank / pg[0]) % pg[1];
  p.coords[2] = rank / (pg[0] * pg[1]);
  for (int k = 0; k < 3; k++) {
    // same split points as Domain::set_local_box / Comm::xsplit (src/domain.cpp:335-360): lo + prd*i/n
    const double prd = boxhi[k] - boxlo[k];
    p.sublo[k] = boxlo[k] + prd * p.coords[k] / pg[k];
    p.subhi[k] = p.coords[k] == pg[k] - 1 ? boxhi[k] : boxlo[k] + prd * (p.coords[k] + 1) / pg[k];
  }
  for (int d = 0; d < NDIR; d++) {
    int v[3];
    dir_vec(d, v);
    int cd[3], cs[3];
    bool okd = d != DIR_SELF, oks = d != DIR_SELF;
    for (int k = 0; k < 3; k++) {
      p.wrap[d][k] = p.rwrap[d][k] = 0;
      // destination brick of a message sent towards +v
      int c = p.coords[k] + v[k];
      if (c < 0) { c += pg[k]; p.wrap[d][k] = +1; }          // leaves through the low face: image at x + prd
      else if (c >= pg[k]) { c -= pg[k]; p.wrap[d][k] = -1; } // leaves through the high face: image at x - prd
      if (p.wrap[d][k] && !periodic[k]) okd = false;
      cd[k] = c;
      // source brick: the one for which (its coords + v) lands here
      int s = p.coords[k] - v[k];
      if (s < 0) { s += pg[k]; p.rwrap[d][k] = -1; }          // the sender wrapped through its high face
      else if (s >= pg[k]) { s -= pg[k]; p.rwrap[d][k] = +1; }
      if (p.rwrap[d][k] && !periodic[k]) oks = false;
      cs[k] = s;
    }
    p.dest[d] = okd ? (cd[2] * pg[1] + cd[1]) * pg[0] + cd[0] : -1;
    p.src[d] = oks ? (cs[2] * pg[1] + cs[1]) * pg[0] + cs[0] : -1;
  }
  return 0;
}

}  // namespace polb200
