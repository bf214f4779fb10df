/* ----------------------------------------------------------------------
   B200 drop-in for `fix rigid/nve|nvt molecule` (see the header).  Host side only: argument parsing with the
   reference's keywords and error texts for what is supported, the Fix flags other parts of LAMMPS read
   (rigid_flag, dof_flag, virial_flag, time_integrate), and the marshalling into polb200_rigid_*.
------------------------------------------------------------------------- */

#include <mpi.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "fix_rigid_nve.h"
#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "group.h"
#include "memory.h"
#include "update.h"
#include "polb200.h"
#include "device_atoms_b200.h"

using namespace LAMMPS_NS;
using namespace FixConst;

FixRigidNHB200::FixRigidNHB200(LAMMPS *lmp, int narg, char **arg) :
  Fix(lmp, narg, arg), handle(NULL), ingroup(NULL), massone(NULL), nchain_restart(0), chain_restart(NULL)
{
  // the flags FixRigid's constructor sets (src/RIGID/fix_rigid.cpp:75-83)
  scalar_flag = 1;
  extscalar = 0;
  time_integrate = 1;
  rigid_flag = 1;
  virial_flag = 1;
  thermo_virial = 1;
  dof_flag = 1;
  nbody = 0;
  setupflag = 0;
  nmax_work = 0;
  t_target = 0.0;

  if (narg < 4) error->all(FLERR,"Illegal fix rigid command");
  if (strcmp(arg[3],"molecule") != 0)
    error->all(FLERR,"fix rigid/nve|nvt (B200): only bodystyle 'molecule' is on the device");
  if (atom->molecule_flag == 0)
    error->all(FLERR,"Fix rigid molecule requires atom attribute molecule");
  if (comm->nprocs != 1) error->all(FLERR,"fix rigid/nve|nvt (B200) runs on one MPI rank");
  if (domain->dimension != 3) error->all(FLERR,"fix rigid/nve|nvt (B200) needs a 3d simulation");

  tstat_flag = 0;
  t_chain = 10; t_iter = 1; t_order = 3;         // fix_rigid.cpp:323-325
  t_start = t_stop = t_period = 0.0;
  int iarg = 4;
  while (iarg < narg) {
    if (strcmp(arg[iarg],"temp") == 0) {          // fix_rigid.cpp:418-427
      if (iarg+4 > narg) error->all(FLERR,"Illegal fix rigid command");
      tstat_flag = 1;
      t_start = force->numeric(FLERR,arg[iarg+1]);
      t_stop = force->numeric(FLERR,arg[iarg+2]);
      t_period = force->numeric(FLERR,arg[iarg+3]);
      iarg += 4;
    } else if (strcmp(arg[iarg],"tparam") == 0) { // fix_rigid.cpp:528-537
      if (iarg+4 > narg) error->all(FLERR,"Illegal fix rigid command");
      t_chain = force->inumeric(FLERR,arg[iarg+1]);
      t_iter = force->inumeric(FLERR,arg[iarg+2]);
      t_order = force->inumeric(FLERR,arg[iarg+3]);
      iarg += 4;
    } else {
      char msg[160];
      snprintf(msg,sizeof(msg),"fix rigid/nve|nvt (B200): keyword '%s' is not on the device",arg[iarg]);
      error->all(FLERR,msg);
    }
  }
  if (strcmp(style,"rigid/nvt") == 0) {            // fix_rigid_nvt.cpp:29-49
    scalar_flag = 1;
    restart_global = 1;
    extscalar = 1;
    if (tstat_flag == 0) error->all(FLERR,"Did not set temperature for fix rigid/nvt");
    if (t_start < 0.0 || t_stop <= 0.0)
      error->all(FLERR,"Target temperature for fix rigid/nvt cannot be 0.0");
    if (t_period <= 0.0) error->all(FLERR,"Fix rigid/nvt period must be > 0.0");
    if (t_chain < 1) error->all(FLERR,"Illegal fix rigid/nvt command");
    if (t_iter < 1) error->all(FLERR,"Illegal fix rigid/nvt  command");
    if (t_order != 3 && t_order != 5)
      error->all(FLERR,"Fix rigid/nvt temperature order must be 3 or 5");
  }

  const char *dev = getenv("POLB200_DEVICE");
  polb200_rigid_t *h = NULL;
  if (polb200_rigid_create(&h, dev ? atoi(dev) : 0) != POLB200_OK)
    error->all(FLERR,"fix rigid/nve|nvt: no usable CUDA device (there is no CPU path)");
  handle = h;

  // the statistics line of the reference's constructor (fix_rigid.cpp:627-636): count the bodies here
  {
    const int nlocal = atom->nlocal;
    tagint maxmol = -1;
    for (int i = 0; i < nlocal; i++)
      if (atom->mask[i] & groupbit) maxmol = MAX(maxmol,atom->molecule[i]);
    int nsum = 0;
    if (maxmol >= 0) {
      int *ncount = new int[maxmol+1];
      for (tagint m = 0; m <= maxmol; m++) ncount[m] = 0;
      for (int i = 0; i < nlocal; i++)
        if (atom->mask[i] & groupbit) { ncount[atom->molecule[i]]++; nsum++; }
      for (tagint m = 0; m <= maxmol; m++) {
        if (ncount[m]) nbody++;
        if (ncount[m] == 1) error->all(FLERR,"One or zero atoms in rigid body");
      }
      delete [] ncount;
    }
    if (comm->me == 0) {
      if (screen) fprintf(screen,"%d rigid bodies with %d atoms\n",nbody,nsum);
      if (logfile) fprintf(logfile,"%d rigid bodies with %d atoms\n",nbody,nsum);
    }
  }
}

FixRigidNHB200::~FixRigidNHB200()
{
  if (handle) polb200_rigid_destroy(handle);
  memory->destroy(ingroup);
  memory->destroy(massone);
  delete [] chain_restart;
}

void FixRigidNHB200::fail()
{
  error->all(FLERR, polb200_rigid_last_error(handle));
}

int FixRigidNHB200::setmask()
{
  int mask = 0;
  mask |= INITIAL_INTEGRATE;
  mask |= FINAL_INTEGRATE;
  mask |= PRE_NEIGHBOR;
  mask |= PRE_EXCHANGE;                     // device-resident atoms: host copies made current before the host re-orders them
  if (tstat_flag) mask |= THERMO_ENERGY;   // fix_rigid_nh.cpp:196-203
  return mask;
}

void FixRigidNHB200::fill_work()
{
  const int nlocal = atom->nlocal;
  if (nlocal > nmax_work) {
    nmax_work = atom->nmax;
    memory->destroy(ingroup);
    memory->destroy(massone);
    memory->create(ingroup,nmax_work,"rigid/b200:ingroup");
    memory->create(massone,nmax_work,"rigid/b200:massone");
  }
  for (int i = 0; i < nlocal; i++) {
    ingroup[i] = (atom->mask[i] & groupbit) ? 1 : 0;
    massone[i] = atom->rmass ? atom->rmass[i] : atom->mass[atom->type[i]];
  }
}

/* FixRigid::init + FixRigidNH::init: the bodies are (re)built from the current atoms on every init
   (reinitflag = 1, the reference's default) */

void FixRigidNHB200::init()
{
  if (domain->triclinic) error->all(FLERR,"fix rigid/nve|nvt (B200): triclinic boxes are not supported");
  if (strstr(update->integrate_style,"respa"))
    error->all(FLERR,"fix rigid/nve|nvt (B200) does not support run_style respa");
  fill_work();
  DeviceAtomsB200::instance().evaluate(lmp);   // may this run keep its atoms on the device? (device_atoms_b200.h)
  polb200_rigid_params p;
  memset(&p, 0, sizeof(p));
  p.thermostat = tstat_flag;
  p.t_start = t_start; p.t_stop = t_stop; p.t_period = t_period;
  p.t_chain = t_chain; p.t_iter = t_iter; p.t_order = t_order;
  p.dt = update->dt;
  p.ftm2v = force->ftm2v; p.mvv2e = force->mvv2e; p.boltz = force->boltz;
  for (int d = 0; d < 3; d++) {
    p.boxlo[d] = domain->boxlo[d];
    p.boxhi[d] = domain->boxhi[d];
    p.periodic[d] = domain->periodicity[d];
  }
  polb200_rigid_info info;
  if (polb200_rigid_init(handle, &p, atom->nlocal, atom->tag, atom->molecule, ingroup, massone, atom->image,
                         atom->x[0], atom->v[0], &info) != POLB200_OK) fail();
  nbody = info.nbody;
  setupflag = 1;
  t_target = t_start;
  if (chain_restart) {   // FixRigidNH::restart ran before the first init: the chains start from the stored state
    if (polb200_rigid_set_chain(handle, chain_restart, nchain_restart) != POLB200_OK) fail();
    delete [] chain_restart;
    chain_restart = NULL;
  }
}

void FixRigidNHB200::setup_pre_neighbor()
{
  pre_neighbor();
}

void FixRigidNHB200::setup(int vflag)
{
  polb200_rigid_atoms a;
  a.nlocal = atom->nlocal; a.tag = atom->tag; a.x = atom->x[0]; a.v = atom->v[0]; a.f = atom->f[0]; a.on_device = 0;
  if (vflag) v_setup(vflag);
  else evflag = 0;
  if (polb200_rigid_setup(handle, &a, evflag && vflag_global) != POLB200_OK) fail();
  if (evflag && vflag_global && polb200_rigid_virial(handle, virial) != POLB200_OK) fail();
}

/* device-resident atoms (device_atoms_b200.h): the integrator works on the shared device mirror; the host gets the new
   positions every step (Neighbor::decide reads them) and the new velocities / forces on output steps only */

void FixRigidNHB200::initial_integrate(int vflag)
{
  DeviceAtomsB200 &da = DeviceAtomsB200::instance();
  const bool resident = da.resident && atom->nlocal > 0;
  polb200_rigid_atoms a;
  a.nlocal = atom->nlocal; a.tag = atom->tag; a.x = atom->x[0]; a.v = atom->v[0]; a.f = atom->f[0]; a.on_device = 0;
  if (resident) {
    da.ensure_xv(lmp);
    a.tag = da.tag; a.x = da.x; a.v = da.v; a.f = da.f; a.on_device = 1;
  }
  if (vflag) v_setup(vflag);
  else evflag = 0;
  double delta = update->ntimestep - update->beginstep;       // fix_rigid_nh.cpp:1109-1115
  if (delta != 0.0) delta /= update->endstep - update->beginstep;
  t_target = t_start + delta * (t_stop-t_start);
  if (polb200_rigid_initial_integrate(handle, &a, evflag && vflag_global, delta) != POLB200_OK) fail();
  if (resident) {
    da.v_host_current = false;
    da.download_x(lmp);
  }
}

void FixRigidNHB200::final_integrate()
{
  DeviceAtomsB200 &da = DeviceAtomsB200::instance();
  const bool resident = da.resident && atom->nlocal > 0 && da.xv_on_device;
  polb200_rigid_atoms a;
  a.nlocal = atom->nlocal; a.tag = atom->tag; a.x = atom->x[0]; a.v = atom->v[0]; a.f = atom->f[0]; a.on_device = 0;
  if (resident) {
    a.tag = da.tag; a.x = da.x; a.v = da.v; a.f = da.f; a.on_device = 1;
  }
  if (polb200_rigid_final_integrate(handle, &a) != POLB200_OK) fail();
  if (evflag && vflag_global && polb200_rigid_virial(handle, virial) != POLB200_OK) fail();
  if (resident) {
    da.v_host_current = false;
    if (da.output_step(lmp)) {
      da.download_v(lmp);
      da.download_f(lmp);
    }
  }
}

/* re-neighboring step, before Domain::pbc / Comm::exchange / Atom::sort touch the host arrays: the host copies of the
   velocities and dipoles must be the current ones */

void FixRigidNHB200::pre_exchange()
{
  DeviceAtomsB200 &da = DeviceAtomsB200::instance();
  if (da.resident) da.before_reneighbor(lmp);
}

void FixRigidNHB200::pre_neighbor()
{
  DeviceAtomsB200 &da = DeviceAtomsB200::instance();
  if (da.resident) da.after_reneighbor();   // the atoms may have been wrapped and re-ordered: the host is authoritative
  if (!setupflag) return;
  if (polb200_rigid_pre_neighbor(handle, atom->nlocal, atom->tag, atom->image, 0) != POLB200_OK) fail();
}

int FixRigidNHB200::dof(int tgroup)
{
  if (!setupflag) {   // fix_rigid.cpp:1185-1190
    if (comm->me == 0)
      error->warning(FLERR,"Cannot count rigid body degrees-of-freedom before bodies are initialized");
    return 0;
  }
  const int tgroupbit = group->bitmask[tgroup];
  const int nlocal = atom->nlocal;
  int *tg = new int[nlocal > 0 ? nlocal : 1];
  for (int i = 0; i < nlocal; i++) tg[i] = (atom->mask[i] & tgroupbit) ? 1 : 0;
  int n = 0;
  const int rc = polb200_rigid_dof(handle, nlocal, atom->tag, tg, &n);
  delete [] tg;
  if (rc != POLB200_OK) fail();
  return n;
}

void FixRigidNHB200::deform(int)
{
  error->all(FLERR,"fix rigid/nve|nvt (B200): box-changing fixes are not supported");
}

void FixRigidNHB200::reset_dt()
{
  if (polb200_rigid_reset_dt(handle, update->dt) != POLB200_OK) fail();
}

double FixRigidNHB200::compute_scalar()
{
  double s = 0.0;
  if (polb200_rigid_scalar(handle, &s, NULL, NULL) != POLB200_OK) fail();
  return s;
}

void *FixRigidNHB200::extract(const char *str, int &dim)
{
  if (strcmp(str,"t_target") == 0) {
    dim = 0;
    return &t_target;
  }
  return NULL;
}

/* the record FixRigidNH::write_restart writes (fix_rigid_nh.cpp:1171-1224): tstat_flag, t_chain, 4 doubles per chain
   link, pstat_flag = 0 -- restart files move freely between the reference and the drop-in */

void FixRigidNHB200::write_restart(FILE *fp)
{
  if (tstat_flag == 0) return;
  const int nsize = 2 + 1 + 4*t_chain;
  double *list = new double[nsize];
  int n = 0, nc = 0;
  list[n++] = tstat_flag;
  list[n++] = t_chain;
  if (!setupflag) nc = -1;                                                  // before the first init: chains at rest
  else if (polb200_rigid_get_chain(handle, list+n, 4*t_chain, &nc) != POLB200_OK) fail();
  if (nc != t_chain) for (int i = 0; i < 4*t_chain; i++) list[n+i] = 0.0;
  n += 4*t_chain;
  list[n++] = 0;
  if (comm->me == 0) {
    int size = nsize*sizeof(double);
    fwrite(&size,sizeof(int),1,fp);
    fwrite(list,sizeof(double),nsize,fp);
  }
  delete [] list;
}

void FixRigidNHB200::restart(char *buf)
{
  double *list = (double *) buf;
  int n = 0;
  const int flag = static_cast<int> (list[n++]);
  if (!flag) return;
  const int m = static_cast<int> (list[n++]);
  if (tstat_flag && m == t_chain) {
    delete [] chain_restart;
    chain_restart = new double[4*m];
    for (int i = 0; i < 4*m; i++) chain_restart[i] = list[n++];
    nchain_restart = m;
  }
}

double FixRigidNHB200::memory_usage()
{
  return (double) nmax_work * (sizeof(int) + sizeof(double));
}
