/* -*- c++ -*- ----------------------------------------------------------
   B200 drop-in for `kspace_style pppm` (second half).

   Replaces src/KSPACE/pppm.{h,cpp} of the reference tree: same style name and KSpace virtuals, so the
   `kspace_style pppm <accuracy>` line of an unchanged input selects it; `kspace_modify order / mesh / gewald` reach it
   through the KSpace base class as before.  Charge assignment, FFTs (cuFFT), the Poisson solve with the optimal
   influence function and the field interpolation run behind the C ABI of include/polb200.h (polb200_pppm_*).
   In the lmp_b200 build the classes derived from the reference's PPPM (pppm/cg, pppm/stagger, pppm/tip4p) leave with it.
------------------------------------------------------------------------- */

#ifdef KSPACE_CLASS

KSpaceStyle(pppm,PPPM)

#else

#ifndef LMP_PPPM_H
#define LMP_PPPM_H

#include "kspace.h"

struct polb200_pppm;

namespace LAMMPS_NS {

class PPPM : public KSpace {
 public:
  PPPM(class LAMMPS *lmp, int narg, char **arg);
  ~PPPM();
  void init();                        // reference pppm.cpp:184-395
  void setup();                       // :400-495
  void compute(int eflag, int vflag); // :622-765
  double memory_usage();

 private:
  struct polb200_pppm *handle;
  double cutoff;                      // the pair style's cut_coul
  void plan(int print);
};

}

#endif
#endif
