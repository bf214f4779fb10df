/* ----------------------------------------------------------------------
   B200 drop-in for `kspace_style pppm` (see the header).  Host side only: checks, the quantities the KSpace base
   class owns (qsum/qsqsum, accuracy, order / mesh / gewald of kspace_modify, g_ewald handed to the pair style), and
   the marshalling of atom->x / q / f into polb200_pppm_compute.
------------------------------------------------------------------------- */

#include <mpi.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "pppm.h"
#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "pair.h"
#include "polb200.h"
#include "device_atoms_b200.h"

using namespace LAMMPS_NS;

PPPM::PPPM(LAMMPS *lmp, int narg, char **arg) : KSpace(lmp, narg, arg), handle(NULL)
{
  if (narg < 1) error->all(FLERR,"Illegal kspace_style pppm command");
  pppmflag = 1;
  group_group_enable = 0;             // compute group/group kspace is not offered by the device path
  accuracy_relative = fabs(force->numeric(FLERR,arg[0]));
  nx_pppm = ny_pppm = nz_pppm = 0;
  cutoff = 0.0;
  if (comm->nprocs != 1) error->all(FLERR,"kspace_style pppm (B200) runs on one MPI rank");
  const char *dev = getenv("POLB200_DEVICE");
  polb200_pppm_t *h = NULL;
  if (polb200_pppm_create(&h, dev ? atoi(dev) : 0) != POLB200_OK)
    error->all(FLERR,"kspace_style pppm: no usable CUDA device (there is no CPU path)");
  handle = h;
}

PPPM::~PPPM()
{
  if (handle) polb200_pppm_destroy(handle);
}

/* hand the current box / charges / accuracy / kspace_modify settings to the library: PPPM::init + setup */

void PPPM::plan(int print)
{
  polb200_pppm_setup in;
  memset(&in, 0, sizeof(in));
  in.accuracy_relative = accuracy_absolute >= 0.0 ? accuracy_absolute / two_charge_force : accuracy_relative;
  in.g_ewald = gewaldflag ? g_ewald : 0.0;
  in.order = order;
  if (gridflag) { in.mesh[0] = nx_pppm; in.mesh[1] = ny_pppm; in.mesh[2] = nz_pppm; }
  in.qqrd2e = qqrd2e;
  in.two_charge_force = two_charge_force;
  in.qsum = qsum;
  in.qsqsum = qsqsum;
  in.natoms = (long) atom->natoms;
  in.cutoff = cutoff;
  for (int d = 0; d < 3; d++) {
    in.boxlo[d] = domain->boxlo[d];
    in.boxhi[d] = domain->boxhi[d];
    in.periodic[d] = domain->periodicity[d];
  }
  polb200_pppm_info info;
  if (polb200_pppm_init(handle, &in, &info) != POLB200_OK)
    error->all(FLERR, polb200_pppm_last_error(handle));
  if (!gewaldflag) g_ewald = info.g_ewald;   // stays "not set by the user": the next init estimates it again, as the reference does
  if (!gridflag) { nx_pppm = info.nx; ny_pppm = info.ny; nz_pppm = info.nz; }
  if (print && comm->me == 0) {
    FILE *out[2] = {screen, logfile};
    for (int k = 0; k < 2; k++)
      if (out[k]) {
        fprintf(out[k],"  G vector (1/distance) = %g\n",info.g_ewald);
        fprintf(out[k],"  grid = %d %d %d\n",info.nx,info.ny,info.nz);
        fprintf(out[k],"  stencil order = %d\n",info.order);
        fprintf(out[k],"  using double precision FFTs (cuFFT)\n");
      }
  }
}

void PPPM::init()
{
  if (comm->me == 0) {
    if (screen) fprintf(screen,"PPPM initialization (B200) ...\n");
    if (logfile) fprintf(logfile,"PPPM initialization (B200) ...\n");
  }
  triclinic_check();
  if (domain->triclinic) error->all(FLERR,"kspace_style pppm (B200) requires an orthogonal box");
  if (domain->dimension == 2) error->all(FLERR,"Cannot use PPPM with 2d simulation");
  if (!atom->q_flag) error->all(FLERR,"Kspace style requires atom attribute q");
  if (slabflag) error->all(FLERR,"kspace_modify slab is not offered by the B200 PPPM");
  if (domain->nonperiodic > 0) error->all(FLERR,"Cannot use nonperiodic boundaries with PPPM");
  if (differentiation_flag == 1) error->all(FLERR,"kspace_modify diff ad is not offered by the B200 PPPM");

  pair_check();
  int itmp;
  double *p_cutoff = (double *) force->pair->extract("cut_coul",itmp);
  if (p_cutoff == NULL) error->all(FLERR,"KSpace style is incompatible with Pair style");
  cutoff = *p_cutoff;

  scale = 1.0;
  qqrd2e = force->qqrd2e;
  qsum_qsq();
  natoms_original = atom->natoms;
  const int user_grid = gridflag;
  if (!user_grid) nx_pppm = ny_pppm = nz_pppm = 0;
  plan(1);
}

/* box changed: the reference's setup() keeps grid and g_ewald and recomputes the influence function */

void PPPM::setup()
{
  const int gf = gridflag, ef = gewaldflag;
  gridflag = gewaldflag = 1;          // freeze what init chose
  plan(0);
  gridflag = gf;
  gewaldflag = ef;
}

void PPPM::compute(int eflag, int vflag)
{
  if (eflag || vflag) ev_setup(eflag,vflag);
  else evflag = evflag_atom = eflag_global = vflag_global = eflag_atom = vflag_atom = 0;
  if (evflag_atom) error->all(FLERR,"per-atom KSpace energy/virial is not offered by the B200 PPPM");

  if (atom->natoms != natoms_original) {
    qsum_qsq();
    natoms_original = atom->natoms;
    setup();
  }
  if (qsqsum == 0.0 || atom->nlocal == 0) return;

  double e = 0.0, v[6];
  // device-resident atoms (device_atoms_b200.h): positions and charges are read from, forces added to the shared mirror
  DeviceAtomsB200 &da = DeviceAtomsB200::instance();
  const bool resident = da.resident && !update->setupflag && da.xv_on_device && da.static_on_device;
  if (polb200_pppm_compute(handle, atom->nlocal, resident ? da.x : atom->x[0], resident ? da.q : atom->q, resident ? da.f : atom->f[0],
      eflag_global, vflag_global ? 1 : 0, resident ? 1 : 0,
                           &e, v) != POLB200_OK)
    error->all(FLERR, polb200_pppm_last_error(handle));
  if (resident) da.f_host_current = false;
  if (eflag_global) energy += e;
  if (vflag_global)
    for (int k = 0; k < 6; k++) virial[k] += v[k];
}

double PPPM::memory_usage()
{
  return 0.0;  // everything lives in device memory
}
