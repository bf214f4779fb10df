/* polb200 oracle shim for the git-ignored upstream header accelerator_kokkos.h.
   Written from scratch: dummy classes so that a KOKKOS-less serial build links. */
#ifndef LMP_ACCELERATOR_KOKKOS_H
#define LMP_ACCELERATOR_KOKKOS_H

#include "atom.h"
#include "comm_brick.h"
#include "comm_tiled.h"
#include "domain.h"
#include "neighbor.h"
#include "memory.h"
#include "modify.h"

#define SPECIAL_MASK 0
enum ExecutionSpace { Host, Device };

namespace LAMMPS_NS {

class KokkosLMP {
 public:
  int kokkos_exists, num_threads, ngpu, numa;
  KokkosLMP(class LAMMPS *, int, char **) : kokkos_exists(0), num_threads(1), ngpu(0), numa(1) {}
  ~KokkosLMP() {}
  void accelerator(int, char **) {}
  int neigh_list_kokkos(int) { return 0; }
  int neigh_count(int) { return 0; }
};

class AtomKokkos : public Atom {
 public:
  tagint **k_special;
  AtomKokkos(class LAMMPS *lmp) : Atom(lmp), k_special(NULL) {}
  void sync(const ExecutionSpace, unsigned int) {}
  void modified(const ExecutionSpace, unsigned int) {}
};

class CommKokkos : public CommBrick {
 public:
  CommKokkos(class LAMMPS *lmp) : CommBrick(lmp) {}
};

class CommTiledKokkos : public CommTiled {
 public:
  CommTiledKokkos(class LAMMPS *lmp) : CommTiled(lmp) {}
  CommTiledKokkos(class LAMMPS *lmp, Comm *oldcomm) : CommTiled(lmp, oldcomm) {}
};

class DomainKokkos : public Domain {
 public:
  DomainKokkos(class LAMMPS *lmp) : Domain(lmp) {}
};

class NeighborKokkos : public Neighbor {
 public:
  NeighborKokkos(class LAMMPS *lmp) : Neighbor(lmp) {}
};

class MemoryKokkos : public Memory {
 public:
  MemoryKokkos(class LAMMPS *lmp) : Memory(lmp) {}
  void grow_kokkos(tagint **, tagint **, int, int, const char *) {}
};

class ModifyKokkos : public Modify {
 public:
  ModifyKokkos(class LAMMPS *lmp) : Modify(lmp) {}
};

class DAT {
 public:
  typedef double tdual_xfloat_1d;
  typedef double tdual_FFT_SCALAR_1d;
  typedef int t_int_1d;
  typedef int tdual_int_2d;
};

}
#endif
