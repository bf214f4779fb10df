/* polb200.h -- C ABI of the B200-native hot path of pair style lj/cut/coul/long/polarization.
 *
 * This is the drop-in boundary: a LAMMPS Pair subclass registered under the reference's own style
 * name (see lammps-induced-dipole-polarization-pair-style_b200/lammps/) forwards its virtuals to
 * these entry points and nothing else crosses the boundary -- plain pointers, sizes and C structs,
 * no C++ or torch types, no exceptions.  Every function returns POLB200_OK (0) or an error code;
 * polb200_last_error() gives the message the caller passes to error->all()/error->one().
 * The shape follows the reference's own precedent for "Pair subclass -> extern C -> CUDA library",
 * the GPU package (src/GPU/pair_lj_cut_coul_long_gpu.cpp:50-80: ljcl_gpu_init / _compute_n / _clear).
 *
 * Citations are file:line under the reference tree; "pol.cpp" abbreviates
 * src/pair_lj_cut_coul_long_polarization.cpp.
 *
 * There is NO CPU fallback: without a CUDA device polb200_create() fails with POLB200_ERR_CUDA.
 */
#ifndef POLB200_H
#define POLB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POLB200_ABI_VERSION 5

typedef struct polb200_handle polb200_t;

enum {
  POLB200_OK = 0,
  POLB200_ERR_ARG = 1,      /* message = the reference's error->all text, e.g. "Illegal pair_style command" */
  POLB200_ERR_CUDA = 2,     /* CUDA runtime failure or no device */
  POLB200_ERR_STATE = 3,    /* call order (e.g. compute before init) */
  POLB200_ERR_UNSUPPORTED = 4, /* triclinic box, cutoffs beyond half the box in list mode, ... */
  POLB200_ERR_OVERFLOW = 5, /* neighbor capacity ("Neighbor list overflow, boost neigh_modify one") */
  POLB200_ERR_NAN = 6       /* "Non-numeric positions - simulation unstable" (src/nbin.cpp:120-121) */
};

/* status bits in polb200_result.status */
#define POLB200_STATUS_DIVERGED 1 /* pol.cpp:1227-1235: caller emits the reference's warning text */
#define POLB200_STATUS_REBUILT 2  /* neighbor structures were rebuilt in this call */
#define POLB200_STATUS_EXACT 4    /* all-pairs minimum-image (reference) polarization path was used */

/* ---- life cycle -------------------------------------------------------------------------------- */

/* replaces the constructor pol.cpp:55-91 (defaults :65-78).  device = CUDA ordinal.
 * POLB200_DEVICE_NONE makes a configuration-only handle (settings/coeff/init/single/extract/restart
 * work on the host; polb200_compute fails with POLB200_ERR_CUDA) for input validation and restart
 * tools on machines without a GPU. */
#define POLB200_DEVICE_NONE (-1)
int polb200_create(polb200_t **h, int device);
/* replaces the destructor pol.cpp:95-121 */
void polb200_destroy(polb200_t *h);
const char *polb200_last_error(const polb200_t *h);
int polb200_abi_version(void);

/* ---- configuration: same arguments, defaults and error texts as the reference ------------------ */

/* PairLJCutCoulLongPolarization::settings(narg,arg), pol.cpp:678-766.  arg[0]=cut_lj_global,
 * arg[1]=cut_coul, then keyword/value pairs: precision zodid fixed_iteration damp max_iterations
 * damp_type polar_gs polar_gs_ranked polar_gamma debug use_previous -- plus the documented extensions
 *   polar_cutoff <r|none>  dipole-dipole cutoff (none = reference all-pairs semantics, the default)
 *   restart_keywords <yes|no>  write the polarization keywords into restart files (extension record, default no)
 *   gs_chunks <n>          list-mode Gauss-Seidel (polar_gs / polar_gs_ranked): 0 (default) = group-coloured sweep
 *                          (pair groups coloured so that close groups differ, colours visited in turn, rank metric =
 *                          priority of the colouring); n > 0 = n contiguous per-atom chunks of the ranked order,
 *                          n < 0 = |n| interleaved per-atom chunks (the forms the oracle's gs_chunks emulates)
 * Order-dependent validation is the reference's (e.g. "zodid" errors while polar_gs_ranked is on). */
int polb200_settings(polb200_t *h, int narg, const char *const *arg);
/* Atom::ntypes; allocates the (ntypes+1)^2 coefficient arrays = allocate(), pol.cpp:651-672 */
int polb200_set_ntypes(polb200_t *h, int ntypes);
/* coeff(narg,arg), pol.cpp:772-800: "I J epsilon sigma [cut_lj]" with Force::bounds wildcards */
int polb200_coeff(polb200_t *h, int narg, const char *const *arg);
/* Pair::modify_params (src/pair.cpp:132-185) subset that shapes this style: mix, shift, table, tabinner */
int polb200_pair_modify(polb200_t *h, int narg, const char *const *arg);

typedef struct {
  double g_ewald;          /* force->kspace->g_ewald, pol.cpp:845-847 */
  double qqrd2e;           /* force->qqrd2e */
  double special_lj[4];    /* force->special_lj */
  double special_coul[4];  /* force->special_coul */
  int newton_pair;         /* force->newton_pair (0 and 1: the owner-computes device path serves both; `newton off` selects the
                              pairwise virial tally, as in the reference) */
  double skin;             /* neighbor->skin */
  int neigh_every, neigh_delay, neigh_check; /* neigh_modify every/delay/check (1,10,1) */
  int kspace_present;      /* 0 => "Pair style requires a KSpace style" */
  int q_flag, polarizability_flag; /* atom->q_flag, atom->static_polarizability_flag */
  int molecular;           /* atom->molecular: special-bond lists are meaningful */
} polb200_env;

/* init_style(), pol.cpp:806-852, followed by Pair::init()'s loop over init_one(i,j)
 * (src/pair.cpp:227-255, pol.cpp:858-921) and init_tables (src/pair.cpp:313-520). */
int polb200_init(polb200_t *h, const polb200_env *env);
/* value init_one(i,j) returned (the pair cutoff); valid after polb200_init */
int polb200_init_one(const polb200_t *h, int i, int j, double *cut);
/* pair_modify tail yes: etail_ij / ptail_ij of init_one (pol.cpp:897-918); count_i / count_j = number of atoms
 * of types i and j over all ranks (the reference MPI_Allreduces them; the caller owns that sum).  Zero unless
 * `pair_modify tail yes` was given. */
int polb200_tail(const polb200_t *h, int i, int j, double count_i, double count_j, double *etail_ij, double *ptail_ij);
/* extract(), pol.cpp:1101-1109: "cut_coul" (dim 0), "epsilon"/"sigma" (dim 2, (ntypes+1)^2 row-major).
 * Returns a pointer into host memory owned by the handle, or NULL. */
const void *polb200_extract(const polb200_t *h, const char *name, int *dim);
/* single(), pol.cpp:1035-1097 (host arithmetic; qi,qj passed explicitly) */
int polb200_single(const polb200_t *h, int itype, int jtype, double qi, double qj, double rsq,
                   double factor_coul, double factor_lj, double *fforce, double *eng);
/* write_restart_settings / read_restart_settings payload (pol.cpp:976-1009): 7 fields, and
 * write_restart / read_restart per-pair records (pol.cpp:927-970), as a flat byte image. */
int polb200_restart_size(const polb200_t *h, long *nbytes);
int polb200_write_restart(const polb200_t *h, void *buf, long nbytes);
int polb200_read_restart(polb200_t *h, const void *buf, long nbytes);
/* The settings block alone = the first polb200_restart_settings_size() bytes of the image: what
 * write_restart_settings / read_restart_settings exchange (pol.cpp:976-1009; PairHybrid calls exactly these for its
 * sub-styles, src/pair_hybrid.cpp:650,691).  40 bytes in the reference's layout; with the extension keyword
 * `restart_keywords yes` an 88-byte record with every polarization keyword follows (magic "POLB2KW1"; such files are not
 * readable by the reference, which stores none of them).  polb200_read_restart_settings accepts both forms and reports
 * the bytes it consumed: a caller reading a stream reads 40 bytes, peeks 8 more, and reads 80 more when they are the
 * magic.  ABI version 5. */
#define POLB200_RESTART_MAGIC "POLB2KW1"
int polb200_restart_settings_size(const polb200_t *h, long *nbytes);
int polb200_read_restart_settings(polb200_t *h, const void *buf, long nbytes, long *consumed);

/* `neigh_modify exclude ...` (src/neighbor.cpp:2276-2333): the rules of NPair::exclusion (src/npair.cpp:173-203), at most 32,
 * and `neigh_modify include g` (src/neighbor.cpp:2264-2274; the list then holds pairs of two atoms of g only,
 * src/nbin_standard.cpp:209-223, src/npair_half_bin_newton.cpp:51).
 * The reference removes excluded pairs from the list its LJ / real-space Coulomb loop walks (pol.cpp:232-321) and from
 * nothing else -- static field, dipole solve and dipole forces loop over all pairs -- and so does the device path.
 *   POLB200_EXCL_TYPE       exclude type a b            (pairs of atom types a and b, either order)
 *   POLB200_EXCL_GROUP      exclude group g1 g2         a, b = the groups' bitmasks (Group::bitmask)
 *   POLB200_EXCL_MOL_INTRA  exclude molecule/intra g    a = bitmask; both atoms in g and in the same molecule
 *   POLB200_EXCL_MOL_INTER  exclude molecule/inter g    a = bitmask; both atoms in g and in different molecules
 *   POLB200_EXCL_INCLUDE    include g                   a = bitmask; every pair with an atom OUTSIDE g is left out
 * nrules = 0 clears them (`exclude none`).  Takes effect at the next rebuild. */
enum { POLB200_EXCL_TYPE = 0, POLB200_EXCL_GROUP = 1, POLB200_EXCL_MOL_INTRA = 2, POLB200_EXCL_MOL_INTER = 3, POLB200_EXCL_INCLUDE = 4 };
typedef struct {
  int kind, a, b;
} polb200_exclusion;
int polb200_set_exclusions(polb200_t *h, int nrules, const polb200_exclusion *rules);

/* orthogonal periodic box (Domain::boxlo/boxhi/periodicity); triclinic => POLB200_ERR_UNSUPPORTED */
int polb200_set_box(polb200_t *h, const double boxlo[3], const double boxhi[3], const int periodic[3]);

/* ---- the hot path ------------------------------------------------------------------------------ */

/* x, q, type, alpha, mu and f are required (POLB200_ERR_ARG otherwise).  x, mu, q and alpha are read on every call;
 * type, molecule, tag, mask and the special lists are sampled at the rebuilds (calls with ago == 0, or whenever the
 * library's own schedule decides to rebuild). */
typedef struct {
  int nlocal;
  const double *x;          /* [nlocal][3] AoS, atom->x[0]  (host or device pointer, see `on_device`) */
  const double *q;          /* [nlocal] */
  const int *type;          /* [nlocal] 1-based */
  const int *molecule;      /* [nlocal] (32-bit tagint, src/lmptype.h:83-85) or NULL = all 0 */
  const int *tag;           /* [nlocal] or NULL = 1..nlocal */
  const double *alpha;      /* atom->static_polarizability [nlocal] */
  double *mu;               /* atom->mu_induced [nlocal][3]  in/out */
  double *ef_static;        /* atom->ef_static  [nlocal][3]  out (may be NULL) */
  double *f;                /* atom->f [nlocal][3]  accumulated (+=) */
  const int *nspecial;      /* atom->nspecial [nlocal][3] or NULL */
  const int *special;       /* atom->special [nlocal][maxspecial] (tags) or NULL */
  int maxspecial;
  int on_device;            /* 0: all pointers are host memory; 1: all are device memory on this GPU */
  /* per-atom tallies (Pair::eatom / Pair::vatom, src/pair.h:38), accumulated (+=); required when the
   * per-atom bits of eflag / vflag are set (eflag & 2, vflag & 4), ignored otherwise.  ABI version 2. */
  double *eatom;            /* [nlocal] */
  double *vatom;            /* [nlocal][6]  xx yy zz xy xz yz */
  /* ABI version 4 */
  const int *mask;          /* atom->mask [nlocal] (group bits); needed only by group / molecule exclusion rules, else NULL */
} polb200_atoms;

typedef struct {
  double eng_vdwl, eng_coul, eng_pol; /* Pair::eng_vdwl / eng_coul / eng_pol (src/pair.h:36) */
  double virial[6];                   /* Pair::virial, xx yy zz xy xz yz */
  double u_self, u_ef, u_dd;          /* the `debug yes` partial sums, pol.cpp:632-636 */
  double rmin;                        /* pol.cpp:196-209 (gs_ranked only) */
  int iterations;                     /* DipoleSolverIterative() return value */
  int status;                         /* POLB200_STATUS_* bits */
  long npairs_full;                   /* entries of the device full neighbor list */
  int nghost;
  float ms_neigh, ms_pair, ms_scf, ms_force, ms_total; /* CUDA-event stage times of this call */
} polb200_result;

/* compute(eflag,vflag), pol.cpp:125-645.  `ago` = neighbor->ago (0: LAMMPS rebuilt its lists this
 * step => rebuild device structures; >0: reuse, refresh ghost positions).  ago < 0: the library runs
 * the reference's own rebuild schedule itself (Neighbor::decide/check_distance,
 * src/neighbor.cpp:1923-2001: every/delay/check, half-skin displacement trigger).
 * eflag/vflag use LAMMPS' encoding (src/integrate.cpp:122-157): eflag&1 global energy,
 * vflag%4 == 1 pairwise virial, == 2 F.r virial; eflag & 2 / vflag & 4 request the per-atom tallies of
 * Pair::ev_tally / ev_tally_xyz (half of every pair's energy and virial to each of its atoms; the
 * polarization terms contribute to the per-atom virial only, as in the reference). */
int polb200_compute(polb200_t *h, const polb200_atoms *atoms, int eflag, int vflag, int ago,
                    polb200_result *out);

/* ---- multi-GPU (one process per GPU; spatial decomposition) ------------------------- */

/* Size of the opaque NCCL unique id the ranks must share (rank 0 creates it). */
int polb200_comm_id_size(void);
int polb200_comm_create_id(void *id_bytes);
/* The all-pairs (exact) mode on several GPUs -- SURVEY 8e's caveat: that interaction set (every minimum-image pair, no
 * dipole cutoff; what the reference computes, pol.cpp:1243-1316) does not decompose into bricks, so it is shared as a 1-D
 * row partition instead.  EVERY process passes the WHOLE system to polb200_compute (identical arrays); the O(N^2) stages --
 * static field, Jacobi dipole sweeps, polarization forces -- are computed for the process's rows and all-gathered
 * (dipoles: 24 N bytes per sweep, as double4 records), the O(N) stages and the sequential Gauss-Seidel sweeps run
 * replicated, and every process returns the full, identical result (do not sum it over the processes).  Rows are handed
 * out in blocks of the kernels' row-block size, so all fixed-order reductions see the partial sums of a single-GPU run:
 * results and iteration counts are bit-identical to one GPU.  Not combinable with polb200_comm_init / polar_cutoff,
 * no per-atom tallies.  nccl_unique_id as for polb200_comm_init. */
int polb200_comm_init_replicated(polb200_t *h, int rank, int nranks, const void *nccl_unique_id);
/* Join a communicator of nranks processes; procgrid = bricks per dimension (px*py*pz == nranks). */
int polb200_comm_init(polb200_t *h, int rank, int nranks, const void *id_bytes, const int procgrid[3]);
/* Sub-domain owned by this rank, valid after polb200_set_box + polb200_comm_init. */
int polb200_subdomain(const polb200_t *h, double sublo[3], double subhi[3]);
/* The decomposition plan as plain host arithmetic (no device, no communicator): for each of the 27
 * directions d = (dz+1)*9 + (dy+1)*3 + (dx+1) the rank that receives what `rank` sends towards d
 * (dest, -1 = none), the rank whose direction-d message `rank` receives (src), and the periodic image
 * shift (units of the box length) the sender adds to coordinates (wrap[3*d+k]).  Used by callers that
 * migrate atoms themselves and by the CPU (gloo) test of the decomposition. */
int polb200_decomp_plan(int nranks, int rank, const int procgrid[3], const int periodic[3], const double boxlo[3],
                        const double boxhi[3], int dest[27], int src[27], int wrap[81], double sublo[3],
                        double subhi[3]);

/* ---- KSpace: reciprocal-space Ewald ------------------------------------------------
 * Replaces class Ewald of the reference (src/KSPACE/ewald.{h,cpp}, `kspace_style ewald <accuracy>`), the
 * long-range partner every input of the pair style needs: same g_ewald / kmax selection, same half-space k set,
 * same energy (incl. self and neutralising terms), forces and virial.  Orthogonal, fully periodic boxes; no slab
 * correction, no per-atom tallies, no group/group.  A separate handle: LAMMPS owns Pair and KSpace separately. */
typedef struct polb200_ewald polb200_ewald_t;

typedef struct {
  double accuracy_relative;   /* the argument of kspace_style ewald (KSpace::accuracy_relative) */
  double g_ewald;             /* > 0: kspace_modify gewald; <= 0: estimate as Ewald::init (ewald.cpp:153-160) */
  double qqrd2e;              /* force->qqrd2e */
  double two_charge_force;    /* KSpace::two_charge_force (src/kspace.cpp:79-81) */
  double qsum, qsqsum;        /* KSpace::qsum_qsq over all atoms (src/kspace.cpp:271-306) */
  long natoms;                /* atom->natoms */
  double cutoff;              /* the pair style's cut_coul (extract("cut_coul")) */
  double boxlo[3], boxhi[3];
  int periodic[3];
} polb200_ewald_setup;

typedef struct {
  double g_ewald, gsqmx;
  int kxmax, kymax, kzmax, kmax, kcount;  /* what Ewald::init prints (ewald.cpp:186-205) */
} polb200_ewald_info;

int polb200_ewald_create(polb200_ewald_t **e, int device);
void polb200_ewald_destroy(polb200_ewald_t *e);
const char *polb200_ewald_last_error(const polb200_ewald_t *e);
/* Ewald::init + setup (ewald.cpp:87-340); call again when the box or the charges change (Ewald::setup) */
int polb200_ewald_init(polb200_ewald_t *e, const polb200_ewald_setup *in, polb200_ewald_info *info);
/* Ewald::compute (ewald.cpp:357-497): f[nlocal][3] += KSpace forces; *energy / virial[6] set when the global bits
 * of eflag / vflag ask for them.  on_device: x, q, f are device pointers on this GPU. */
int polb200_ewald_compute(polb200_ewald_t *e, int nlocal, const double *x, const double *q, double *f, int eflag,
                          int vflag, int on_device, double *energy, double virial[6]);
double polb200_ewald_last_ms(const polb200_ewald_t *e);  /* CUDA-event time of the last compute */
/* Multi-GPU (one process per GPU, any partition of the atoms, e.g. the bricks of polb200_comm_init): every rank calls this once
 * with the same id (polb200_comm_create_id) and then polb200_ewald_init with the GLOBAL qsum / qsqsum / natoms (what
 * KSpace::qsum_qsq all-reduces) and polb200_ewald_compute with its own atoms.  The structure factors are all-reduced over
 * NCCL (the reference's MPI_Allreduce, ewald.cpp:395-400); energy and virial come back as per-rank partials that add up. */
int polb200_ewald_comm_init(polb200_ewald_t *e, int rank, int nranks, const void *id_bytes);

/* ---- KSpace: PPPM (second half) ---------------------------------------------------------
 * Replaces class PPPM of the reference (src/KSPACE/pppm.{h,cpp}, `kspace_style pppm <accuracy>`): ik differentiation, no
 * stagger, orthogonal fully periodic box, one GPU.  Same g_ewald and grid selection (set_grid_global :985-1135,
 * adjust_gewald :1287-1340), same optimal influence function (compute_gf_ik :1549-1627), same order-n charge assignment
 * (:1951-1995, :2844-2952), energy incl. self and neutralising terms, forces, virial (:622-765).  FFTs by cuFFT (loaded at
 * run time).  Not offered: `kspace_modify diff ad`, stagger, slab, triclinic, per-atom tallies, group/group, TIP4P. */
typedef struct polb200_pppm polb200_pppm_t;

typedef struct {
  double accuracy_relative;   /* the argument of kspace_style pppm */
  double g_ewald;             /* > 0: kspace_modify gewald; <= 0: estimate and refine as the reference */
  int order;                  /* kspace_modify order (2..7); <= 0: 5 */
  int mesh[3];                /* kspace_modify mesh nx ny nz; any <= 0: choose from the accuracy */
  double qqrd2e, two_charge_force;
  double qsum, qsqsum;        /* KSpace::qsum_qsq */
  long natoms;
  double cutoff;              /* the pair style's cut_coul */
  double boxlo[3], boxhi[3];
  int periodic[3];
} polb200_pppm_setup;

typedef struct {
  double g_ewald;
  int nx, ny, nz, order;      /* what PPPM::init prints (pppm.cpp:340-365) */
} polb200_pppm_info;

int polb200_pppm_create(polb200_pppm_t **p, int device);
void polb200_pppm_destroy(polb200_pppm_t *p);
const char *polb200_pppm_last_error(const polb200_pppm_t *p);
/* PPPM::init + setup (pppm.cpp:184-395, 400-495) */
int polb200_pppm_init(polb200_pppm_t *p, const polb200_pppm_setup *in, polb200_pppm_info *info);
/* PPPM::compute (pppm.cpp:622-765): f[nlocal][3] += KSpace forces; *energy / virial[6] as the global bits of eflag / vflag ask */
int polb200_pppm_compute(polb200_pppm_t *p, int nlocal, const double *x, const double *q, double *f, int eflag, int vflag,
                         int on_device, double *energy, double virial[6]);
double polb200_pppm_last_ms(const polb200_pppm_t *p);
/* Multi-GPU, as polb200_ewald_comm_init: every rank assigns the charges of its own atoms to ONE shared grid, the grid is
 * all-reduced over NCCL (16 B per grid point and step), FFTs and the Poisson solve are replicated, every rank interpolates
 * the forces of its own atoms.  init takes the GLOBAL qsum / qsqsum / natoms. */
int polb200_pppm_comm_init(polb200_pppm_t *p, int rank, int nranks, const void *id_bytes);

/* ---- Rigid-body integrator -----------------------------------------------------------
 * Replaces `fix rigid/nve molecule` and `fix rigid/nvt molecule` of the reference's RIGID package, the integrator of
 * every shipped polarization example (src/RIGID/fix_rigid_nh.{h,cpp} on top of fix_rigid.{h,cpp}): bodies = molecules
 * of the fix group, point particles, orthogonal periodic box.  Body state (centre of mass, quaternion, conjugate
 * quaternion momentum, principal axes, thermostat chains) lives on the device; every per-step entry point accepts
 * host pointers (atom->x / v / f of a LAMMPS Fix) or device pointers (on_device = 1: positions, velocities and forces
 * stay in HBM across steps together with polb200_compute / polb200_ewald_compute).  Per-atom body data is keyed by
 * atom id, so the caller may reorder its atoms between calls (atom sorting) as long as it passes the matching `tag`.
 * Not offered (clean errors): single/group/custom bodies, extended particles, force/torque keywords, langevin,
 * infile, rigid/npt|nph, triclinic boxes, 2d. */
typedef struct polb200_rigid polb200_rigid_t;

typedef struct {
  int thermostat;                    /* 0 = rigid/nve, 1 = rigid/nvt (FixRigidNH::tstat_flag) */
  double t_start, t_stop, t_period;  /* `temp Tstart Tstop Tdamp` (fix_rigid.cpp:418-427) */
  int t_chain, t_iter, t_order;      /* `tparam` (defaults 10 1 3, fix_rigid.cpp:323-325); order 3 or 5 */
  double dt;                         /* update->dt */
  double ftm2v, mvv2e, boltz;        /* force->ftm2v / mvv2e / boltz */
  double boxlo[3], boxhi[3];
  int periodic[3];
} polb200_rigid_params;

typedef struct {
  int nbody;       /* FixRigid::nbody */
  int nlinear;     /* bodies with a zero principal moment */
  int nf_t, nf_r;  /* FixRigidNH::nf_t / nf_r (fix_rigid_nh.cpp:232-244) */
  int maxmembers;  /* atoms of the largest body */
} polb200_rigid_info;

typedef struct {
  int nlocal;
  const int *tag;   /* atom->tag [nlocal] */
  double *x;        /* atom->x [nlocal][3], in/out */
  double *v;        /* atom->v [nlocal][3], in/out */
  const double *f;  /* atom->f [nlocal][3] */
  int on_device;    /* 0: host pointers (copied in and out inside the call); 1: device pointers on this GPU */
} polb200_rigid_atoms;

int polb200_rigid_create(polb200_rigid_t **r, int device);
void polb200_rigid_destroy(polb200_rigid_t *r);
const char *polb200_rigid_last_error(const polb200_rigid_t *r);
/* constructor body numbering (fix_rigid.cpp:130-220: molecules of the group in ascending id) + FixRigid::init
 * (:701-765: setup_bodies_static :1605-2112 incl. the Jacobi diagonalisation and the "Bad principal moments" checks,
 * setup_bodies_dynamic :2120-2211) + FixRigidNH::init (fix_rigid_nh.cpp:208-262).  Host pointers: mass = per-atom
 * mass (rmass[i] or mass[type[i]]), image = atom->image (packed 10-bit fields, src/lmptype.h:96-103),
 * ingroup = 1 where mask[i] & groupbit (NULL = all atoms). */
int polb200_rigid_init(polb200_rigid_t *r, const polb200_rigid_params *p, int nlocal, const int *tag, const int *molecule,
                       const int *ingroup, const double *mass, const int *image, const double *x, const double *v,
                       polb200_rigid_info *info);
/* FixRigid::dof (fix_rigid.cpp:1181-1262); tgroup = 1 where the atom is in the temperature group (NULL = all) */
int polb200_rigid_dof(polb200_rigid_t *r, int nlocal, const int *tag, const int *tgroup, int *dof);
/* FixRigid::setup + FixRigidNH::setup (fix_rigid.cpp:782-889, fix_rigid_nh.cpp:323-421): body force and torque from
 * the current forces, velocities made rigid-consistent (v is written), doubled virial, conjugate momenta, thermostat
 * masses.  vflag != 0 tallies the fix's virial. */
int polb200_rigid_setup(polb200_rigid_t *r, const polb200_rigid_atoms *a, int vflag);
/* FixRigidNH::initial_integrate (fix_rigid_nh.cpp:428-603): x and v are written; f = the forces the last force call
 * left (set_xv's constraint virial reads them).  run_fraction = (ntimestep - beginstep) / (endstep - beginstep), the
 * argument of compute_temp_target (:1109-1115; ignored by rigid/nve). */
int polb200_rigid_initial_integrate(polb200_rigid_t *r, const polb200_rigid_atoms *a, int vflag, double run_fraction);
/* FixRigidNH::final_integrate (fix_rigid_nh.cpp:607-790): v is written; f = forces at the new positions */
int polb200_rigid_final_integrate(polb200_rigid_t *r, const polb200_rigid_atoms *a);
/* FixRigid::pre_neighbor (fix_rigid.cpp:1137-1175) after the caller wrapped its atoms (Domain::pbc): bodies are
 * remapped into the box, the atoms' body-relative image flags are recomputed from `image` (atom->image). */
int polb200_rigid_pre_neighbor(polb200_rigid_t *r, int nlocal, const int *tag, const int *image, int on_device);
/* Fix::virial of the last step (set_xv + set_v halves), xx yy zz xy xz yz */
int polb200_rigid_virial(polb200_rigid_t *r, double virial[6]);
/* FixRigidNH::compute_scalar (fix_rigid_nh.cpp:991-1016; = FixRigid::compute_scalar :2595-2622 for rigid/nve),
 * FixRigid::extract_ke (:2650-2659) and extract_erotational (:2665-2689); any pointer may be NULL */
int polb200_rigid_scalar(polb200_rigid_t *r, double *scalar, double *ke_translational, double *ke_rotational);
/* The thermostat state FixRigidNH::write_restart / restart carry (fix_rigid_nh.cpp:1171-1267): per chain link
 * eta_t, eta_r, eta_dot_t, eta_dot_r, interleaved in that order (4 * t_chain doubles).  set before polb200_rigid_setup. */
int polb200_rigid_get_chain(polb200_rigid_t *r, double *state, int capacity, int *t_chain);
int polb200_rigid_set_chain(polb200_rigid_t *r, const double *state, int t_chain);
/* More than one process (one per GPU), the reference's scheme for `fix rigid` (fix_rigid.cpp:782-855, :1181-1262,
 * :1605-2211): every process holds EVERY body, each atom is owned by exactly one process, and the per-body force and
 * torque sums are all-reduced (one ncclAllReduce of 6 * nbody doubles in setup / final_integrate, where the reference has
 * MPI_Allreduce(sum, all, 6*nbody)).  Call once BEFORE polb200_rigid_init with the bytes of polb200_comm_create_id();
 * afterwards `nlocal` of every call is this process's own atoms (0 is allowed), init gathers the atoms of all
 * processes and builds bit-identical bodies everywhere, dof all-reduces its member counts, the body state (scalar,
 * chain, fetch) is the same on every process, and polb200_rigid_virial returns this process's share (sum it over the
 * processes like any per-process virial, fix_rigid.cpp:1112-1129). */
int polb200_rigid_comm_init(polb200_rigid_t *r, int rank, int nranks, const void *nccl_unique_id);
/* FixRigid::reset_dt (fix_rigid.cpp:2553-2558) */
int polb200_rigid_reset_dt(polb200_rigid_t *r, double dt);
/* body arrays for tests / compute_array: "xcm" "vcm" "fcm" "torque" "angmom" "omega" "ex" "ey" "ez" "inertia" [nbody][3],
 * "quat" "conjqm" [nbody][4], "masstotal" [nbody]; returns the number of doubles copied or < 0 */
long polb200_rigid_fetch(polb200_rigid_t *r, const char *name, double *dst, long capacity);
long polb200_rigid_launch_count(polb200_rigid_t *r, int reset);
double polb200_rigid_last_ms(const polb200_rigid_t *r); /* CUDA-event time of the last per-step call */

/* ---- device buffers for callers that keep their atom arrays resident ---------------------------------------------- */
/* A caller that passes device pointers (`on_device = 1` of polb200_compute / _ewald_compute / _pppm_compute / the rigid
 * entry points) needs device memory of its own.  A host framework without a CUDA tool chain -- the LAMMPS binding's shared
 * device mirror of atom->x / v / f / q / mu, lammps/device_atoms_b200.h -- gets it here: plain allocation and synchronous
 * copies on `device` (every library entry point returns with its own stream drained, so these copies are ordered with the
 * library's work).  polb200_host_register pins a host range (faster copies); all return POLB200_OK or POLB200_ERR_CUDA. */
enum { POLB200_COPY_H2D = 0, POLB200_COPY_D2H = 1, POLB200_COPY_D2D = 2 };
void *polb200_dev_alloc(int device, size_t bytes);
void polb200_dev_free(int device, void *p);
int polb200_dev_copy(int device, void *dst, const void *src, size_t bytes, int kind);
int polb200_dev_zero(int device, void *p, size_t bytes);
int polb200_host_register(void *p, size_t bytes);
int polb200_host_unregister(void *p);

/* ---- introspection for tests and profiling ------------------------------------------------------ */

/* Copy internal state to host (tests, profiling).  Device arrays: "perm" (int nlocal: cell-sorted -> caller index), "tag",
 * "ghost_owner" (int nghost: cell-sorted owner index), "ghost_shift" (int nghost: packed image code), "rowstart" (u64
 * nlocal+1), "neigh" (int: full list, entry = ext index | special << 30), "xq" / "mua" (double4 per owned+ghost atom), "ef"
 * (double4 per owned atom), "ranked", "metric", "flags"; "gs_colouring" (int 2*nlocal+3: colour and in-group predecessor
 * of every atom in caller order, then {colours, colouring rounds, groups}).  Statistics: "sweep_timing", "polar_pairs",
 * "group_stats", "comm_stats", "barrier_stats".  Host tables: "h_rtable" ... "h_cutneighsq", "h_tabmeta".
 * Returns the number of elements copied or < 0. */
long polb200_debug_fetch(polb200_t *h, const char *name, void *dst, long capacity_bytes);
/* number of kernel launches issued by this handle since creation / since last reset */
long polb200_launch_count(polb200_t *h, int reset);
/* knobs for experiments and A/B measurements: "sweep_variant" (41 = TMA-fed pair-group sweep, default; 40 / 44; 30 / 31
 * register-prefetch pair groups; 20 / 21 per-atom rows + radial cache; 6 matrix-free; 0 first version), "bin_div",
 * "xsort_bits", "gs_blocked", "gs_cluster" (CTAs of the cluster kernel of the exact-mode Gauss-Seidel sweep, default 16;
 * 0 = one launch per block), "gs_cache_max", "use_graphs", "gs_colours", "gs_strong_m", "scf_lag", "use_group_pairs",
 * "gpf_minb", "use_tight", "use_push", "p2p_push", "l2_evict_first", "alternate", "time_sweeps".
 * One option is part of the multi-GPU interface rather than a knob: "atom_slack" (Angstrom, 0..16, default 0) lets a
 * decomposed caller hand a brick atoms up to that far OUTSIDE its sub-domain (whole molecules kept on one process): ghost
 * shells, send lists and the cell grid reach cutneighmax + slack.  POLB200_ERR_ARG for unknown names. */
int polb200_set_option(polb200_t *h, const char *name, double value);

#ifdef __cplusplus
}
#endif
#endif
