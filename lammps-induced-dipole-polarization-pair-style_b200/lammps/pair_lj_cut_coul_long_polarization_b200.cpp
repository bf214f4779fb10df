/* ----------------------------------------------------------------------
   B200 drop-in for pair style lj/cut/coul/long/polarization (see the header).

   Call map (reference = src/pair_lj_cut_coul_long_polarization.cpp):
     ctor/dtor        :55-121    -> polb200_create / polb200_destroy
     settings         :678-766   -> polb200_settings        (same keywords, defaults, error texts)
     coeff            :772-800   -> polb200_coeff           (+ Pair::setflag kept in step)
     init_style       :806-852   -> polb200_init            (g_ewald, qqrd2e, special_*, skin, neigh_modify)
     init_one         :858-921   -> polb200_init_one
     compute          :125-645   -> polb200_set_box + polb200_compute
     single           :1035-1097 -> polb200_single
     extract          :1101-1109 -> polb200_extract
     restart / data   :927-1031  -> polb200_write_restart / polb200_read_restart (byte-identical records)
------------------------------------------------------------------------- */

#include <mpi.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "pair_lj_cut_coul_long_polarization.h"
#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "group.h"
#include "kspace.h"
#include "memory.h"
#include "neighbor.h"
#include "update.h"
#include "polb200.h"
#include "device_atoms_b200.h"

using namespace LAMMPS_NS;

#define CHECK(call) check((call), FLERR)

/* ---------------------------------------------------------------------- */

PairLJCutCoulLongPolarization::PairLJCutCoulLongPolarization(LAMMPS *lmp) : Pair(lmp)
{
  ewaldflag = pppmflag = 1;
  respa_enable = 0;
  writedata = 1;
  ftable = NULL;
  // the device returns the finished virial (pair part pairwise, polarization part as the reference's F.r sum
  // over owned atoms), so ev_setup must keep vflag_global instead of deferring to
  // virial_fdotr_compute() (src/pair.cpp:809-815)
  no_virial_fdotr_compute = 1;
  handle = NULL;
  debug = 0;
  ntypes_set = 0;
  nexclude_sent = 0;
  epsilon_rows = sigma_rows = NULL;
  const char *dev = getenv("POLB200_DEVICE");
  device = dev ? atoi(dev) : 0;
  if (comm->nprocs != 1)
    error->all(FLERR,"Pair style lj/cut/coul/long/polarization (B200) runs one MPI rank per pair instance; "
                     "multi-GPU runs decompose inside the library");
  polb200_t *h = NULL;
  if (polb200_create(&h, device) != POLB200_OK)
    error->all(FLERR,"Pair style lj/cut/coul/long/polarization: no usable CUDA device (there is no CPU path)");
  handle = h;
}

PairLJCutCoulLongPolarization::~PairLJCutCoulLongPolarization()
{
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
  }
  free_rows();
  if (handle) polb200_destroy(handle);
}

void PairLJCutCoulLongPolarization::free_rows()
{
  delete [] epsilon_rows;
  delete [] sigma_rows;
  epsilon_rows = sigma_rows = NULL;
}

void PairLJCutCoulLongPolarization::check(int rc, const char *file, int line)
{
  if (rc == POLB200_OK) return;
  // the library hands back the reference's own message text for configuration errors
  error->all(file, line, polb200_last_error(handle));
}

/* Pair::setflag / cutsq are what Pair::init() and Neighbor read; everything else lives in the library */

void PairLJCutCoulLongPolarization::ensure_types()
{
  if (ntypes_set == atom->ntypes && allocated) return;
  const int n = atom->ntypes;
  CHECK(polb200_set_ntypes(handle, n));
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
  }
  memory->create(setflag, n+1, n+1, "pair:setflag");
  memory->create(cutsq, n+1, n+1, "pair:cutsq");
  for (int i = 0; i <= n; i++)
    for (int j = 0; j <= n; j++) {
      setflag[i][j] = 0;
      cutsq[i][j] = 0.0;
    }
  allocated = 1;
  ntypes_set = n;
  free_rows();
}

/* ---------------------------------------------------------------------- */

void PairLJCutCoulLongPolarization::settings(int narg, char **arg)
{
  CHECK(polb200_settings(handle, narg, arg));
  for (int k = 2; k + 1 < narg; k += 2)
    if (strcmp(arg[k], "debug") == 0) debug = strcmp(arg[k+1], "yes") == 0;
  // a re-issued pair_style resets explicitly set per-pair LJ cutoffs, as the reference does (:757-765);
  // the library applies the same rule to its own tables
}

void PairLJCutCoulLongPolarization::coeff(int narg, char **arg)
{
  ensure_types();
  CHECK(polb200_coeff(handle, narg, arg));
  int ilo, ihi, jlo, jhi;
  force->bounds(FLERR, arg[0], atom->ntypes, ilo, ihi);
  force->bounds(FLERR, arg[1], atom->ntypes, jlo, jhi);
  for (int i = ilo; i <= ihi; i++)
    for (int j = (jlo > i ? jlo : i); j <= jhi; j++) setflag[i][j] = 1;
}

/* pair_modify is handled by the non-virtual Pair::modify_params: forward its results */

void PairLJCutCoulLongPolarization::sync_modify_params()
{
  char table[32], tabin[64];
  snprintf(table, sizeof(table), "%d", ncoultablebits);
  snprintf(tabin, sizeof(tabin), "%.17g", tabinner);
  const char *mixname = mix_flag == GEOMETRIC ? "geometric" : (mix_flag == ARITHMETIC ? "arithmetic" : "sixthpower");
  const char *words[] = {"shift", offset_flag ? "yes" : "no", "mix", mixname, "table", table, "tabinner", tabin,
                         "tail", tail_flag ? "yes" : "no"};
  CHECK(polb200_pair_modify(handle, 10, words));
}

void PairLJCutCoulLongPolarization::init_style()
{
  ensure_types();
  sync_modify_params();
  polb200_env env;
  memset(&env, 0, sizeof(env));
  env.kspace_present = force->kspace != NULL;
  env.g_ewald = force->kspace ? force->kspace->g_ewald : 0.0;
  env.qqrd2e = force->qqrd2e;
  for (int k = 0; k < 4; k++) {
    env.special_lj[k] = force->special_lj[k];
    env.special_coul[k] = force->special_coul[k];
  }
  env.newton_pair = force->newton_pair;
  env.skin = neighbor->skin;
  env.neigh_every = neighbor->every;
  env.neigh_delay = neighbor->delay;
  env.neigh_check = neighbor->dist_check;
  env.q_flag = atom->q_flag;
  env.polarizability_flag = atom->static_polarizability_flag;
  env.molecular = atom->molecular;
  CHECK(polb200_init(handle, &env));
  DeviceAtomsB200::instance().evaluate(lmp);   // may this run keep its atoms on the device? (device_atoms_b200.h)
  if (domain->triclinic)
    error->all(FLERR,"Pair style lj/cut/coul/long/polarization (B200) requires an orthogonal box");
  // no NeighRequest: the library builds its own cell-sorted device list from atom->x on the steps where
  // LAMMPS re-neighbors (neighbor->ago == 0) and refreshes positions otherwise
}

double PairLJCutCoulLongPolarization::init_one(int i, int j)
{
  double cut = 0.0;
  CHECK(polb200_init_one(handle, i, j, &cut));
  setflag[i][j] = 1;  // mixed pairs are now defined, like the reference's init_one (:860-866)
  if (tail_flag) {    // long-range LJ correction (:897-918): type populations summed over all ranks
    double count[2] = {0.0, 0.0}, all[2];
    for (int k = 0; k < atom->nlocal; k++) {
      if (atom->type[k] == i) count[0] += 1.0;
      if (atom->type[k] == j) count[1] += 1.0;
    }
    MPI_Allreduce(count, all, 2, MPI_DOUBLE, MPI_SUM, world);
    CHECK(polb200_tail(handle, i, j, all[0], all[1], &etail_ij, &ptail_ij));
  }
  return cut;
}

/* ---------------------------------------------------------------------- */

void PairLJCutCoulLongPolarization::compute(int eflag, int vflag)
{
  if (eflag || vflag) ev_setup(eflag, vflag);
  else evflag = vflag_fdotr = eflag_global = eflag_atom = vflag_global = vflag_atom = 0;

  int periodic[3] = {domain->xperiodic, domain->yperiodic, domain->zperiodic};
  CHECK(polb200_set_box(handle, domain->boxlo, domain->boxhi, periodic));

  // device-resident atoms (device_atoms_b200.h): the pair style reads x / q / mu from and adds its forces to the shared
  // device mirror; per-atom tallies (eatom / vatom are host arrays of Pair) take the host-buffer path for that step
  DeviceAtomsB200 &da = DeviceAtomsB200::instance();
  const bool resident = da.resident && !update->setupflag && atom->nlocal > 0;
  const bool resident_call = resident && !eflag_atom && !vflag_atom;
  if (resident) {
    da.ensure_xv(lmp);
    da.ensure_static(lmp);
    if (!resident_call) da.download_mu(lmp);
  }

  polb200_atoms a;
  memset(&a, 0, sizeof(a));
  a.nlocal = atom->nlocal;
  if (resident_call) {
    da.zero_f(lmp);                      // Verlet::force_clear cleared the host copy; the pair style is the first contributor
    a.x = da.x; a.q = da.q; a.type = da.type; a.molecule = da.molecule; a.tag = da.tag; a.alpha = da.alpha;
    a.mu = da.mu; a.ef_static = da.ef; a.f = da.f; a.mask = da.mask;
    if (da.maxspecial > 0) {
      a.nspecial = da.nspecial;
      a.special = da.special;
      a.maxspecial = da.maxspecial;
    }
  } else if (a.nlocal > 0) {
    a.x = atom->x[0];
    a.q = atom->q;
    a.type = atom->type;
    a.molecule = atom->molecule;        // 32-bit tagint (LAMMPS_SMALLBIG, src/lmptype.h:83-85)
    a.tag = atom->tag;
    a.alpha = atom->static_polarizability;
    a.mu = atom->mu_induced[0];
    a.ef_static = atom->ef_static[0];
    a.f = atom->f[0];
    if (atom->molecular && atom->maxspecial > 0 && atom->special) {
      a.nspecial = atom->nspecial[0];
      a.special = atom->special[0];
      a.maxspecial = atom->maxspecial;
    }
  }
  a.on_device = resident_call ? 1 : 0;
  // `neigh_modify exclude` (src/neighbor.cpp:2276-2333): LAMMPS' own pair list is not used, so the rules travel to the
  // device list on every re-neighboring step (NPair::exclusion, src/npair.cpp:173-203)
  if (neighbor->ago == 0) {
    std::vector<polb200_exclusion> rules;
    for (int m = 0; m < neighbor->nex_type; m++)
      rules.push_back({POLB200_EXCL_TYPE, neighbor->ex1_type[m], neighbor->ex2_type[m]});
    for (int m = 0; m < neighbor->nex_group; m++)
      rules.push_back({POLB200_EXCL_GROUP, group->bitmask[neighbor->ex1_group[m]], group->bitmask[neighbor->ex2_group[m]]});
    for (int m = 0; m < neighbor->nex_mol; m++)
      rules.push_back({neighbor->ex_mol_intra[m] ? POLB200_EXCL_MOL_INTRA : POLB200_EXCL_MOL_INTER,
                       group->bitmask[neighbor->ex_mol_group[m]], 0});
    // `neigh_modify include g` (with `atom_modify first g`): only pairs of two atoms of g reach the pair loop
    if (neighbor->includegroup)
      rules.push_back({POLB200_EXCL_INCLUDE, group->bitmask[neighbor->includegroup], 0});
    if (!rules.empty() || nexclude_sent > 0) {
      CHECK(polb200_set_exclusions(handle, (int) rules.size(), rules.empty() ? NULL : rules.data()));
      nexclude_sent = (int) rules.size();
    }
  }
  if (!resident_call) a.mask = atom->mask;
  a.eatom = eflag_atom ? eatom : NULL;           // Pair::eatom / vatom, zeroed by ev_setup (src/pair.cpp:789-804)
  a.vatom = (vflag_atom && a.nlocal > 0) ? vatom[0] : NULL;

  polb200_result res;
  CHECK(polb200_compute(handle, &a, eflag, vflag, neighbor->ago, &res));
  if (resident_call) {
    da.mu_host_current = false;
    if (da.output_step(lmp)) da.download_mu(lmp);
  } else if (resident) {
    // host-buffer call inside a resident run: the mirror takes over its results (atom->f holds the pair forces only,
    // Verlet::force_clear zeroed it before)
    da.static_on_device = false;
    da.ensure_static(lmp);       // (mu, ef: the host copies are the new ones)
    da.upload_f(lmp);
  }

  if (res.status & POLB200_STATUS_DIVERGED)
    error->warning(FLERR,"Number of iterations exceeding max_iterations, setting dipoles to alpha*E");
  if (eflag_global) {
    eng_vdwl += res.eng_vdwl;
    eng_coul += res.eng_coul;
  }
  eng_pol = res.eng_pol;
  if (vflag_global)
    for (int k = 0; k < 6; k++) virial[k] += res.virial[k];

  if (debug) {
    fprintf(screen, "iterations: %d\n", res.iterations);
    fprintf(screen, "self %.18f\nef %.18f\ndd %.18f\nu_polar calc %.18f\n", res.u_self, res.u_ef, res.u_dd, res.eng_pol);
  }
}

/* ---------------------------------------------------------------------- */

double PairLJCutCoulLongPolarization::single(int i, int j, int itype, int jtype, double rsq, double factor_coul,
                                             double factor_lj, double &fforce)
{
  double eng = 0.0;
  CHECK(polb200_single(handle, itype, jtype, atom->q[i], atom->q[j], rsq, factor_coul, factor_lj, &fforce, &eng));
  return eng;
}

void *PairLJCutCoulLongPolarization::extract(const char *str, int &dim)
{
  const void *p = polb200_extract(handle, str, &dim);
  if (!p) return NULL;
  if (dim == 0) return (void *) p;
  // dim 2: LAMMPS callers expect double** rows over an (ntypes+1)^2 table
  const int n = atom->ntypes + 1;
  double ***rows = strcmp(str, "epsilon") == 0 ? &epsilon_rows : (strcmp(str, "sigma") == 0 ? &sigma_rows : NULL);
  if (!rows) return NULL;
  if (!*rows) *rows = new double*[n];
  for (int i = 0; i < n; i++) (*rows)[i] = (double *) p + (size_t) i * n;
  return (void *) *rows;
}

/* ---------------------------------------------------------------------- restart: proc 0 writes / reads, bcast */

void PairLJCutCoulLongPolarization::write_restart(FILE *fp)
{
  long nbytes = 0;
  CHECK(polb200_restart_size(handle, &nbytes));
  std::vector<char> img(nbytes);
  CHECK(polb200_write_restart(handle, img.data(), nbytes));
  fwrite(img.data(), 1, nbytes, fp);   // = write_restart_settings + per-pair records of the reference (:927-985)
}

void PairLJCutCoulLongPolarization::write_restart_settings(FILE *fp)
{
  long nbytes = 0, nset = 0;
  CHECK(polb200_restart_size(handle, &nbytes));
  CHECK(polb200_restart_settings_size(handle, &nset));
  std::vector<char> img(nbytes);
  CHECK(polb200_write_restart(handle, img.data(), nbytes));
  // the 7 settings fields (2 doubles, 4 ints, 1 double) [+ the keyword record of `restart_keywords yes`]
  fwrite(img.data(), 1, nset, fp);
}

// the settings block as bytes: rank 0 reads, everybody gets it (pol.cpp:991-1009).  A keyword record follows the 7
// reference fields only in files written with `restart_keywords yes`; it announces itself with its magic.
static void read_settings_block(FILE *fp, int me, MPI_Comm world, Error *error, std::vector<char> &img)
{
  int len = 0;
  if (me == 0) {
    img.resize(40);
    if (fread(img.data(), 1, 40, fp) != 40) error->one(FLERR,"Unexpected end of restart file");
    char peek[8];
    const long pos = ftell(fp);
    const size_t got = fread(peek, 1, 8, fp);
    if (got == 8 && memcmp(peek, POLB200_RESTART_MAGIC, 8) == 0) {
      img.insert(img.end(), peek, peek + 8);
      char rest[80];
      if (fread(rest, 1, 80, fp) != 80) error->one(FLERR,"Unexpected end of restart file");
      img.insert(img.end(), rest, rest + 80);
    } else fseek(fp, pos, SEEK_SET);
    len = (int) img.size();
  }
  MPI_Bcast(&len, 1, MPI_INT, 0, world);
  img.resize(len);
  MPI_Bcast(img.data(), len, MPI_CHAR, 0, world);
}

void PairLJCutCoulLongPolarization::apply_settings_block(const std::vector<char> &img)
{
  // base-class copies of the settings that Pair::init()/modify_params consult
  memcpy(&offset_flag, img.data() + 16, 4);
  memcpy(&mix_flag, img.data() + 20, 4);
  memcpy(&tail_flag, img.data() + 24, 4);
  memcpy(&ncoultablebits, img.data() + 28, 4);
  memcpy(&tabinner, img.data() + 32, 8);
}

void PairLJCutCoulLongPolarization::read_restart_settings(FILE *fp)
{
  std::vector<char> img;
  read_settings_block(fp, comm->me, world, error, img);
  CHECK(polb200_read_restart_settings(handle, img.data(), (long) img.size(), NULL));
  apply_settings_block(img);
}

void PairLJCutCoulLongPolarization::read_restart(FILE *fp)
{
  ensure_types();
  const int n = atom->ntypes;
  std::vector<char> img;
  read_settings_block(fp, comm->me, world, error, img);   // = read_restart_settings(fp) of the reference (:949)
  std::vector<char> pairs;
  int len = 0;
  if (comm->me == 0) {
    for (int i = 1; i <= n; i++)
      for (int j = i; j <= n; j++) {
        int flag = 0;
        if (fread(&flag, sizeof(int), 1, fp) != 1) error->one(FLERR,"Unexpected end of restart file");
        pairs.insert(pairs.end(), (char *) &flag, (char *) &flag + sizeof(int));
        if (flag) {
          char rec[24];
          if (fread(rec, 1, 24, fp) != 24) error->one(FLERR,"Unexpected end of restart file");
          pairs.insert(pairs.end(), rec, rec + 24);
        }
      }
    len = (int) pairs.size();
  }
  MPI_Bcast(&len, 1, MPI_INT, 0, world);
  pairs.resize(len);
  MPI_Bcast(pairs.data(), len, MPI_CHAR, 0, world);
  img.insert(img.end(), pairs.begin(), pairs.end());
  CHECK(polb200_read_restart(handle, img.data(), (long) img.size()));
  int dim;
  const int *flags = (const int *) polb200_extract(handle, "setflag", &dim);
  for (int i = 1; i <= n; i++)
    for (int j = i; j <= n; j++) setflag[i][j] = flags[i * (n + 1) + j];
  apply_settings_block(img);
}

/* ---------------------------------------------------------------------- data file: eps/sigma like the reference */

void PairLJCutCoulLongPolarization::write_data(FILE *fp)
{
  int dim;
  const int n = atom->ntypes + 1;
  const double *eps = (const double *) polb200_extract(handle, "epsilon", &dim);
  const double *sig = (const double *) polb200_extract(handle, "sigma", &dim);
  for (int i = 1; i <= atom->ntypes; i++) fprintf(fp, "%d %g %g\n", i, eps[i * n + i], sig[i * n + i]);
}

void PairLJCutCoulLongPolarization::write_data_all(FILE *fp)
{
  int dim;
  const int n = atom->ntypes + 1;
  const double *eps = (const double *) polb200_extract(handle, "epsilon", &dim);
  const double *sig = (const double *) polb200_extract(handle, "sigma", &dim);
  const double *cl = (const double *) polb200_extract(handle, "cut_lj", &dim);
  for (int i = 1; i <= atom->ntypes; i++)
    for (int j = i; j <= atom->ntypes; j++)
      fprintf(fp, "%d %d %g %g %g\n", i, j, eps[i * n + j], sig[i * n + j], cl[i * n + j]);
}
