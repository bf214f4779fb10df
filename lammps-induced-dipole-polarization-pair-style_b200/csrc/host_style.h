// host_style.h -- host-side (C++) state of the pair style: the mirror of the reference's
// settings / coeff / allocate / init_style / init_one / init_tables / single / extract / restart
// (src/pair_lj_cut_coul_long_polarization.cpp:651-1109, src/pair.cpp:132-185,313-520,660-685).
// Pure host code: no CUDA, no LAMMPS headers.  Errors are reported as the reference's own
// error->all() strings through StyleError.
#pragma once
#include <string>
#include <vector>

#include "polb200.h"

namespace polb200 {

struct StyleError {
  int code;
  std::string msg;
};

enum { DAMP_EXPONENTIAL = 0, DAMP_NONE = 1 };             // pol.cpp:51
enum { MIX_GEOMETRIC = 0, MIX_ARITHMETIC = 1, MIX_SIXTHPOWER = 2 };

struct CoulTables {
  int nbits = 0, mask = 0, shift = 0;
  double tabinnersq = 0.0;
  std::vector<double> r, dr, f, df, c, dc, e, de;
};

class HostStyle {
 public:
  HostStyle();

  void settings(int narg, const char *const *arg);        // pol.cpp:678-766
  void set_ntypes(int n);                                  // allocate(), pol.cpp:651-672
  void coeff(int narg, const char *const *arg);            // pol.cpp:772-800
  void pair_modify(int narg, const char *const *arg);      // src/pair.cpp:132-185 (subset)
  void init(const polb200_env &env);                       // Pair::init + init_style + init_one + tables
  double single(int itype, int jtype, double qi, double qj, double rsq, double factor_coul,
                double factor_lj, double &fforce) const;   // pol.cpp:1035-1097
  void tail_correction(int i, int j, double count_i, double count_j, double &etail_ij,
                       double &ptail_ij) const;            // pol.cpp:897-918 (pair_modify tail yes)
  std::vector<char> restart_image() const;                 // pol.cpp:927-941,976-985
  void read_restart_image(const void *buf, long nbytes);   // pol.cpp:947-970,991-1009
  // the settings block alone (write_restart_settings / read_restart_settings, pol.cpp:976-1009): 40 bytes in the
  // reference's layout, followed -- only when `restart_keywords yes` was given -- by the keyword extension record
  std::vector<char> restart_settings_image() const;
  long read_restart_settings_image(const void *buf, long nbytes);  // returns the bytes consumed

  int idx(int i, int j) const { return i * (ntypes + 1) + j; }

  // settings (defaults pol.cpp:65-78)
  double cut_lj_global = 0.0, cut_coul = 0.0;
  int iterations_max = 50, damping_type = DAMP_NONE, zodid = 0, fixed_iteration = 0;
  int polar_gs = 0, polar_gs_ranked = 1, use_previous = 0, debug = 0;
  double polar_damp = 2.1304, polar_precision = 0.00000000001, polar_gamma = 1.03;
  // extensions
  double polar_cutoff = 0.0;  // <= 0: none (reference all-pairs minimum image)
  int gs_chunks = 0;
  int restart_keywords = 0;   // 1: restart files also carry the polarization keywords (extension record; such a file is
                              // no longer readable by the reference binary, which stores none of them)

  // Pair base-class state that shapes this style (src/pair.cpp:82-88)
  int offset_flag = 0, mix_flag = MIX_GEOMETRIC, tail_flag = 0, ncoultablebits = 12;
  double tabinner;

  int ntypes = 0;
  bool allocated = false, initialized = false;
  std::vector<int> setflag;
  std::vector<double> epsilon, sigma, cut_lj, cut_ljsq, cutsq, lj1, lj2, lj3, lj4, offset, cut_pair;
  double cut_coulsq = 0.0, cutforce = 0.0;
  polb200_env env{};
  CoulTables tab;
  std::vector<double> cutneighsq;  // (cut+skin)^2, src/neighbor.cpp:293-320
  double cutneighmax = 0.0;
  int special_flag[4] = {0, 2, 2, 2};  // src/neighbor.cpp:361-382

 private:
  void init_tables();
};

}  // namespace polb200
