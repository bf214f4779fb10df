/* polb200 oracle shim for the git-ignored upstream dihedral_hybrid.h */
#ifdef DIHEDRAL_CLASS
#else
#ifndef LMP_DIHEDRAL_HYBRID_H
#define LMP_DIHEDRAL_HYBRID_H
#include "dihedral.h"
namespace LAMMPS_NS {
class DihedralHybrid : public Dihedral {
 public:
  int nstyles;
  Dihedral **styles;
  char **keywords;
};
}
#endif
#endif
