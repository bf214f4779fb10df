/* -*- c++ -*- ----------------------------------------------------------
   B200 drop-in for `fix rigid/nve molecule` and `fix rigid/nvt molecule`.

   Registered under the reference's own style names (src/RIGID/fix_rigid_nve.h:16, fix_rigid_nvt.h:16), so the
   shipped polarization inputs (`fix rigid_nve all rigid/nve molecule`, `fix rigid_nvt moving rigid/nvt molecule
   temp T T 100.0 tparam 50 1 3`) run unchanged.  The class keeps no body state: centre of mass, quaternion,
   conjugate momenta and the Nose-Hoover chains live on the GPU behind polb200_rigid_* (include/polb200.h); the
   virtuals below marshal atom->x / v / f / tag / image and forward.  In the lmp_b200 build this header and its
   .cpp take the file names fix_rigid_nve.{h,cpp}; fix_rigid_nvt.{h,cpp} leave the build.
------------------------------------------------------------------------- */

#ifdef FIX_CLASS

FixStyle(rigid/nve,FixRigidNHB200)
FixStyle(rigid/nvt,FixRigidNHB200)

#else

#ifndef LMP_FIX_RIGID_NH_B200_H
#define LMP_FIX_RIGID_NH_B200_H

#include <cstdio>
#include "fix.h"

struct polb200_rigid;

namespace LAMMPS_NS {

class FixRigidNHB200 : public Fix {
 public:
  FixRigidNHB200(class LAMMPS *, int, char **);
  ~FixRigidNHB200();
  int setmask();
  void init();
  void setup(int);
  void setup_pre_neighbor();
  void initial_integrate(int);
  void final_integrate();
  void pre_exchange();
  void pre_neighbor();
  int dof(int);
  void deform(int);
  void reset_dt();
  double compute_scalar();
  double memory_usage();
  void *extract(const char *, int &);
  void write_restart(FILE *);
  void restart(char *);

 private:
  struct polb200_rigid *handle;
  int tstat_flag, t_chain, t_iter, t_order, nbody, setupflag;
  double t_start, t_stop, t_period, t_target;
  int nmax_work;
  int nchain_restart;    // thermostat state read from a restart file, handed to the library at the next init()
  double *chain_restart;
  int *ingroup;          // work: 1 where mask & groupbit
  double *massone;       // work: per-atom mass

  void fail();           // error->all with the library's message
  void fill_work();
};

}

#endif
#endif
