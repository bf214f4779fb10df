/* ----------------------------------------------------------------------
   B200 drop-in for `kspace_style ewald` (see the header).  Host side only: checks, the quantities the
   KSpace base class owns (qsum/qsqsum, accuracy, g_ewald handed to the pair style), and the marshalling of
   atom->x / q / f into polb200_ewald_compute.
------------------------------------------------------------------------- */

#include <mpi.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "ewald.h"
#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "pair.h"
#include "polb200.h"
#include "device_atoms_b200.h"

using namespace LAMMPS_NS;

Ewald::Ewald(LAMMPS *lmp, int narg, char **arg) : KSpace(lmp, narg, arg), handle(NULL)
{
  if (narg != 1) error->all(FLERR,"Illegal kspace_style ewald command");
  ewaldflag = 1;
  group_group_enable = 0;             // compute group/group kspace is not offered by the device path
  accuracy_relative = fabs(force->numeric(FLERR,arg[0]));
  kcount = kxmax = kymax = kzmax = kmax = 0;
  cutoff = 0.0;
  if (comm->nprocs != 1)
    error->all(FLERR,"kspace_style ewald (B200) runs on one MPI rank");
  const char *dev = getenv("POLB200_DEVICE");
  polb200_ewald_t *h = NULL;
  if (polb200_ewald_create(&h, dev ? atoi(dev) : 0) != POLB200_OK)
    error->all(FLERR,"kspace_style ewald: no usable CUDA device (there is no CPU path)");
  handle = h;
}

Ewald::~Ewald()
{
  if (handle) polb200_ewald_destroy(handle);
}

/* hand the current box / charges / accuracy to the library: Ewald::init + setup of the reference */

void Ewald::plan(double g_in, int print)
{
  polb200_ewald_setup in;
  memset(&in, 0, sizeof(in));
  in.accuracy_relative = accuracy_absolute >= 0.0 ? accuracy_absolute / two_charge_force : accuracy_relative;
  in.g_ewald = g_in;
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
  polb200_ewald_info info;
  if (polb200_ewald_init(handle, &in, &info) != POLB200_OK)
    error->all(FLERR, polb200_ewald_last_error(handle));
  g_ewald = info.g_ewald;
  kcount = info.kcount;
  kxmax = info.kxmax; kymax = info.kymax; kzmax = info.kzmax; kmax = info.kmax;
  if (print && comm->me == 0) {
    const int kmax3d = 4*kmax*kmax*kmax + 6*kmax*kmax + 3*kmax;
    FILE *out[2] = {screen, logfile};
    for (int k = 0; k < 2; k++)
      if (out[k]) {
        fprintf(out[k],"  G vector (1/distance) = %g\n",g_ewald);
        fprintf(out[k],"  KSpace vectors: actual max1d max3d = %d %d %d\n",kcount,kmax,kmax3d);
        fprintf(out[k],"                  kxmax kymax kzmax  = %d %d %d\n",kxmax,kymax,kzmax);
      }
  }
}

void Ewald::init()
{
  if (comm->me == 0) {
    if (screen) fprintf(screen,"Ewald initialization (B200) ...\n");
    if (logfile) fprintf(logfile,"Ewald initialization (B200) ...\n");
  }
  triclinic_check();
  if (domain->triclinic) error->all(FLERR,"kspace_style ewald (B200) requires an orthogonal box");
  if (domain->dimension == 2) error->all(FLERR,"Cannot use Ewald with 2d simulation");
  if (!atom->q_flag) error->all(FLERR,"Kspace style requires atom attribute q");
  if (slabflag) error->all(FLERR,"kspace_modify slab is not offered by the B200 Ewald");
  if (domain->nonperiodic > 0) error->all(FLERR,"Cannot use nonperiodic boundaries with Ewald");

  pair_check();
  int itmp;
  double *p_cutoff = (double *) force->pair->extract("cut_coul",itmp);
  if (p_cutoff == NULL) error->all(FLERR,"KSpace style is incompatible with Pair style");
  cutoff = *p_cutoff;

  scale = 1.0;
  qqrd2e = force->qqrd2e;
  qsum_qsq();
  natoms_original = atom->natoms;
  plan(gewaldflag ? g_ewald : 0.0, 1);
}

/* box or charges changed: same g_ewald, new k set */

void Ewald::setup()
{
  plan(g_ewald, 0);
}

void Ewald::compute(int eflag, int vflag)
{
  if (eflag || vflag) ev_setup(eflag,vflag);
  else evflag = evflag_atom = eflag_global = vflag_global = eflag_atom = vflag_atom = 0;
  if (evflag_atom) error->all(FLERR,"per-atom KSpace energy/virial is not offered by the B200 Ewald");

  if (atom->natoms != natoms_original) {
    qsum_qsq();
    natoms_original = atom->natoms;
    plan(g_ewald, 0);
  }
  if (qsqsum == 0.0 || atom->nlocal == 0) return;

  double e = 0.0, v[6];
  // device-resident atoms (device_atoms_b200.h): positions and charges are read from, forces added to the shared mirror
  DeviceAtomsB200 &da = DeviceAtomsB200::instance();
  const bool resident = da.resident && !update->setupflag && da.xv_on_device && da.static_on_device;
  if (polb200_ewald_compute(handle, atom->nlocal, resident ? da.x : atom->x[0], resident ? da.q : atom->q, resident ? da.f : atom->f[0],
      eflag_global, vflag_global ? 1 : 0, resident ? 1 : 0,
                            &e, v) != POLB200_OK)
    error->all(FLERR, polb200_ewald_last_error(handle));
  if (resident) da.f_host_current = false;
  if (eflag_global) energy += e;
  if (vflag_global)
    for (int k = 0; k < 6; k++) virial[k] += v[k];
}

double Ewald::memory_usage()
{
  return 0.0;  // everything lives in device memory
}
