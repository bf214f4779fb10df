/* -*- c++ -*- ----------------------------------------------------------
   atom_style full with the three per-atom arrays of the polarization pair style.

   The reference declares static_polarizability / ef_static / mu_induced in Atom (src/atom.h:160-163) and a `set`
   keyword for the first (src/set.cpp:175-181), but ships no atom style that allocates them (the author's
   atom_vec_full.cpp is git-ignored, src/.gitignore:183): as shipped, init_style() of the pair style errors out.
   This class supplies that missing piece as a subclass of the stock AtomVecFull, registered under the unchanged
   style name `full` (the lmp_b200 build renames the stock registration to `full/stock`), and carries the arrays
   through everything an atom style is responsible for:
     grow / copy (atom sorting, deletion)       all three arrays
     border communication                       static_polarizability (ghost polarizabilities: pol.cpp:196-227)
     exchange between processors                static_polarizability, mu_induced, ef_static
     restart files                              static_polarizability, mu_induced (use_previous continues the SCF
                                                from the stored dipoles; no `set` needed after read_restart)
     create_atom / data_atom                    zero-initialised
   Dipoles and fields are exposed to dump / fix ave / Python through compute polarization/atom
   (compute_polarization_atom_b200.{h,cpp}), since Atom::extract and dump custom have closed keyword lists.
------------------------------------------------------------------------- */

#ifdef ATOM_CLASS

AtomStyle(full,AtomVecFullPolar)

#else

#ifndef LMP_ATOM_VEC_FULL_POLAR_B200_H
#define LMP_ATOM_VEC_FULL_POLAR_B200_H

#include "atom_vec_full.h"

namespace LAMMPS_NS {

class AtomVecFullPolar : public AtomVecFull {
 public:
  AtomVecFullPolar(class LAMMPS *);
  void grow(int);
  void grow_reset();
  void copy(int, int, int);
  int pack_border(int, int *, double *, int, int *);
  int pack_border_vel(int, int *, double *, int, int *);
  int pack_border_hybrid(int, int *, double *);
  void unpack_border(int, int, double *);
  void unpack_border_vel(int, int, double *);
  int unpack_border_hybrid(int, int, double *);
  int pack_exchange(int, double *);
  int unpack_exchange(double *);
  int size_restart();
  int pack_restart(int, double *);
  int unpack_restart(double *);
  void create_atom(int, double *);
  void data_atom(double *, imageint, char **);
  int data_atom_hybrid(int, char **);
  bigint memory_usage();

 private:
  double *alpha,**efield,**dipole;        // atom->static_polarizability / ef_static / mu_induced
  enum { NEXCHANGE = 7, NRESTART = 4 };   // trailer lengths behind the stock record
  template <int VEL> int pack_border_any(int, int *, double *, int, int *);
  template <int VEL> void unpack_border_any(int, int, double *);
  void clear_polar(int);
};

}

#endif
#endif
