/* polb200 oracle shim: declaration-only stand-in for the git-ignored upstream
   atom_vec_ellipsoid.h. No ATOM_CLASS section, so the style is never registered;
   core files only need the type to compile their (unused here) ellipsoid branches. */
#ifdef ATOM_CLASS
#else
#ifndef LMP_ATOM_VEC_ELLIPSOID_H
#define LMP_ATOM_VEC_ELLIPSOID_H
#include "atom_vec.h"
namespace LAMMPS_NS {
class AtomVecEllipsoid : public AtomVec {
 public:
  struct Bonus {
    double shape[3];
    double quat[4];
    int ilocal;
  };
  struct Bonus *bonus;
  void set_shape(int, double, double, double) {}
};
}
#endif
#endif
