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

struct polb200_handle;

namespace LAMMPS_NS {

class PairLJCutCoulLongPolarization : public Pair {
 public:
  PairLJCutCoulLongPolarization(class LAMMPS *);
  virtual ~PairLJCutCoulLongPolarization();
  virtual void compute(int, int);
  virtual void settings(int, char **);
  void coeff(int, char **);
  virtual void init_style();
  virtual double init_one(int, int);
  void write_restart(FILE *);
  void read_restart(FILE *);
  virtual void write_restart_settings(FILE *);
  virtual void read_restart_settings(FILE *);
  void write_data(FILE *);
  void write_data_all(FILE *);
  virtual double single(int, int, int, int, double, double, double, double &);
  virtual void *extract(const char *, int &);

 protected:
  struct polb200_handle *handle;   // opaque library state (device arrays, kernels' parameters)
  int device;                      // CUDA ordinal, environment POLB200_DEVICE (default 0)
  int debug;                       // `debug yes`: per-step prints like the reference (:391,:635-639)
  int ntypes_set;
  double **epsilon_rows, **sigma_rows;  // row tables over the library's flat arrays, for extract()

  void ensure_types();
  void check(int rc, const char *file, int line);
  void sync_modify_params();
  void free_rows();
};

}

#endif
#endif
