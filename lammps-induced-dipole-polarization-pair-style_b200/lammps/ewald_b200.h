/* -*- c++ -*- ----------------------------------------------------------
   B200 drop-in for `kspace_style ewald`.

   Replaces src/KSPACE/ewald.{h,cpp} of the reference tree: same style name and KSpace virtuals, so the
   `kspace_style ewald <accuracy>` line of an unchanged input selects it.  The reciprocal-space sums run
   behind the C ABI of include/polb200.h (polb200_ewald_*: hand-written sm_100a CUDA in libpolb200.so).
------------------------------------------------------------------------- */

#ifdef KSPACE_CLASS

KSpaceStyle(ewald,Ewald)

#else

#ifndef LMP_EWALD_H
#define LMP_EWALD_H

#include "kspace.h"

struct polb200_ewald;

namespace LAMMPS_NS {

class Ewald : public KSpace {
 public:
  Ewald(class LAMMPS *lmp, int narg, char **arg);
  ~Ewald();
  void init();                        // reference ewald.cpp:87-209
  void setup();                       // :211-337
  void compute(int eflag, int vflag); // :357-497
  double memory_usage();

 private:
  struct polb200_ewald *handle;
  double cutoff;                      // the pair style's cut_coul
  int kcount, kxmax, kymax, kzmax, kmax;
  void plan(double g_in, int print);
};

}

#endif
#endif
