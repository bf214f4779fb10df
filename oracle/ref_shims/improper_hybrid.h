/* polb200 oracle shim for the git-ignored upstream improper_hybrid.h */
#ifdef IMPROPER_CLASS
#else
#ifndef LMP_IMPROPER_HYBRID_H
#define LMP_IMPROPER_HYBRID_H
#include "improper.h"
namespace LAMMPS_NS {
class ImproperHybrid : public Improper {
 public:
  int nstyles;
  Improper **styles;
  char **keywords;
};
}
#endif
#endif
