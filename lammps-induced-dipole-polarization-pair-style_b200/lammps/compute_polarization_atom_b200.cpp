/* ----------------------------------------------------------------------
   compute polarization/atom (see the header)
------------------------------------------------------------------------- */

#include "compute_polarization_atom_b200.h"
#include "atom.h"
#include "error.h"
#include "memory.h"
#include "update.h"

using namespace LAMMPS_NS;

ComputePolarizationAtom::ComputePolarizationAtom(LAMMPS *lmp, int narg, char **arg) :
  Compute(lmp, narg, arg), nmax(0), pol(NULL)
{
  if (narg != 3) error->all(FLERR,"Illegal compute polarization/atom command");
  if (!atom->static_polarizability_flag)
    error->all(FLERR,"Compute polarization/atom requires atom attribute polarizability");
  peratom_flag = 1;
  size_peratom_cols = 7;
}

ComputePolarizationAtom::~ComputePolarizationAtom()
{
  memory->destroy(pol);
}

void ComputePolarizationAtom::compute_peratom()
{
  invoked_peratom = update->ntimestep;
  if (atom->nmax > nmax) {
    memory->destroy(pol);
    nmax = atom->nmax;
    memory->create(pol,nmax,7,"polarization/atom:pol");
    array_atom = pol;
  }
  const int nlocal = atom->nlocal;
  const int *mask = atom->mask;
  for (int i = 0; i < nlocal; i++) {
    if (mask[i] & groupbit) {
      pol[i][0] = atom->static_polarizability[i];
      for (int k = 0; k < 3; k++) {
        pol[i][1+k] = atom->mu_induced[i][k];
        pol[i][4+k] = atom->ef_static[i][k];
      }
    } else {
      for (int k = 0; k < 7; k++) pol[i][k] = 0.0;
    }
  }
}

double ComputePolarizationAtom::memory_usage()
{
  return (double) nmax * 7 * sizeof(double);
}
