/* -*- c++ -*- ----------------------------------------------------------
   B200 drop-in for pair style lj/cut/coul/long/polarization.

   Replaces src/pair_lj_cut_coul_long_polarization.{h,cpp} of the reference in a LAMMPS 16Mar2018
   source tree: same style name, same Pair virtuals (reference header :30-48), so an unchanged input
   script selects it.  Everything numerical happens behind the C ABI of include/polb200.h
   (libpolb200.so, hand-written sm_100a CUDA); this class only marshals LAMMPS' arrays.
------------------------------------------------------------------------- */

#ifdef PAIR_CLASS

PairStyle(lj/cut/coul/long/polarization,PairLJCutCoulLongPolarization)

#else

#ifndef LMP_PAIR_LJ_CUT_COUL_LONG_POLARIZATION_H
#define LMP_PAIR_LJ_CUT_COUL_LONG_POLARIZATION_H

#include "pair.h"
#include <vector>

struct polb200_handle;

namespace LAMMPS_NS {

class PairLJCutCoulLongPolarization : public Pair {
 public:
  explicit PairLJCutCoulLongPolarization(class LAMMPS *);
  ~PairLJCutCoulLongPolarization();

  // configuration (forwarded to the host-side mirror behind the C ABI)
  void settings(int narg, char **arg);
  void coeff(int narg, char **arg);
  void init_style();
  double init_one(int itype, int jtype);

  // the hot path
  void compute(int eflag, int vflag);

  // queries, restart and data files
  double single(int i, int j, int itype, int jtype, double rsq, double factor_coul, double factor_lj, double &fforce);
  void *extract(const char *name, int &dim);
  void write_restart(FILE *fp);
  void write_restart_settings(FILE *fp);
  void read_restart(FILE *fp);
  void read_restart_settings(FILE *fp);
  void write_data(FILE *fp);
  void write_data_all(FILE *fp);

 private:
  struct polb200_handle *handle;        // opaque library state (device arrays, kernels' parameters)
  int device;                           // CUDA ordinal, environment POLB200_DEVICE (default 0)
  int debug;                            // `debug yes`: per-step prints like the reference (:391,:635-639)
  int ntypes_set;
  int nexclude_sent;                    // neigh_modify exclude rules last handed to the library
  double **epsilon_rows, **sigma_rows;  // row tables over the library's flat arrays, for extract()

  void ensure_types();
  void apply_settings_block(const std::vector<char> &img);
  void check(int rc, const char *file, int line);
  void sync_modify_params();
  void free_rows();
};

}

#endif
#endif
