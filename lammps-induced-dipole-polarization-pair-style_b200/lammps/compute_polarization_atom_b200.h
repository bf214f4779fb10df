/* -*- c++ -*- ----------------------------------------------------------
   compute ID group polarization/atom  

   Per-atom array with 7 columns: static_polarizability, mu_induced x y z, ef_static x y z -- the three arrays of
   the polarization pair style (src/atom.h:160-163), which the reference never exposes (Atom::extract and dump
   custom have closed keyword lists).  `dump d all custom 100 out.dump id type x y z c_pol[2] c_pol[3] c_pol[4]`
   writes the induced dipoles; fix ave/atom, variables and the Python library interface (extract_compute) work
   the same way.  Values are those of the last pair compute(); atoms outside the group read 0.
------------------------------------------------------------------------- */

#ifdef COMPUTE_CLASS

ComputeStyle(polarization/atom,ComputePolarizationAtom)

#else

#ifndef LMP_COMPUTE_POLARIZATION_ATOM_B200_H
#define LMP_COMPUTE_POLARIZATION_ATOM_B200_H

#include "compute.h"

namespace LAMMPS_NS {

class ComputePolarizationAtom : public Compute {
 public:
  ComputePolarizationAtom(class LAMMPS *, int, char **);
  ~ComputePolarizationAtom();
  void init() {}
  void compute_peratom();
  double memory_usage();

 private:
  int nmax;
  double **pol;
};

}

#endif
#endif
