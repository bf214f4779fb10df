/* polref -- CPU restatement (ORACLE, test infrastructure only) of the per-timestep hot path of
 * pair style lj/cut/coul/long/polarization.
 *
 * NOT PART OF THE PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (libpolb200.so) never links it and
 * has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against
 *   (a) the thermo tables of the reference's own committed logs
 *       (polarization/examples/Bulk H2/log.lammps:92-100, MOF5+Methane/log.lammps:150-156), and
 *   (b) per-atom dipoles / fields / forces / neighbor lists dumped from the repaired reference
 *       binary oracle/_ref/lmp_serial (recipe: oracle/build_ref.sh, fixtures: tests/golden/).
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 */
#ifndef POLREF_H
#define POLREF_H

#ifdef __cplusplus
extern "C" {
#endif

#define POLREF_DAMP_EXPONENTIAL 0 /* src/pair_lj_cut_coul_long_polarization.cpp:51 */
#define POLREF_DAMP_NONE 1

#define POLREF_MIX_GEOMETRIC 0 /* src/pair.cpp:660-685 */
#define POLREF_MIX_ARITHMETIC 1
#define POLREF_MIX_SIXTHPOWER 2

typedef struct {
  int ntypes;
  /* (ntypes+1)*(ntypes+1) row-major tables, index [itype*(ntypes+1)+jtype], types are 1-based */
  const double *cutsq, *cut_ljsq, *lj1, *lj2, *lj3, *lj4, *offset;
  double cut_coul, g_ewald, qqrd2e;
  double special_lj[4], special_coul[4];
  /* Coulomb lookup tables (src/pair.cpp:313-520); ncoultablebits==0 => analytic erfc only */
  int ncoultablebits, ncoulmask, ncoulshiftbits;
  double tabinnersq;
  const double *rtable, *drtable, *ftable, *dftable, *ctable, *dctable, *etable, *detable;
  /* polarization keywords (src/pair_lj_cut_coul_long_polarization.cpp:65-78,678-766) */
  int iterations_max, damping_type, zodid, fixed_iteration, polar_gs, polar_gs_ranked, use_previous;
  double polar_damp, polar_precision, polar_gamma;
  /* EXTENSIONS (not in the reference; defaults reproduce it):
   *   polar_cut <= 0 : dipole-dipole over all minimum-image pairs (reference semantics)
   *   polar_cut  > 0 : dipole-dipole restricted to minimum-image pairs with rsq < polar_cut^2
   *   gs_chunks == 0 : Gauss-Seidel strictly sequential in ranked order (reference semantics)
   *   gs_chunks  > 0 : ranked order cut into gs_chunks contiguous chunks; Jacobi inside a chunk,
   *                    Gauss-Seidel between chunks ("ranked colouring sweep" of the CUDA list path) */
  double polar_cut;
  int gs_chunks;
  /* orthogonal box */
  double boxlo[3], boxhi[3];
  int periodic[3];
  /* optional per-atom tallies of Pair::ev_tally / ev_tally_xyz (src/pair.cpp:854-949,1001-1089), newton on:
   * half of every pair's energy and virial to each of its two atoms.  (nlocal+nghost) and (nlocal+nghost)*6
   * doubles, accumulated (+=); NULL = off (eflag_atom / vflag_atom not set). */
  double *eatom, *vatom;
  /* EXTENSION (polref_polar_rows only): explicit colouring of the Gauss-Seidel sweep, as the CUDA list path's
   * group-coloured sweep chooses it (exported by the device, tests only).  gs_colour[i] in [0, gs_ncolours): the
   * colours are visited in turn, Jacobi inside a colour; gs_after[i] = a >= 0: atom i (second member of a pair group)
   * is updated right after atom a of the same colour and sees its new dipole.  NULL = off. */
  const int *gs_colour, *gs_after;
  int gs_ncolours;
} polref_params;

typedef struct {
  double eng_vdwl, eng_coul, eng_pol;
  double virial[6];
  double u_self, u_ef, u_dd;
  double rmin;
  int iterations;
  int diverged; /* 1 if "Number of iterations exceeding max_iterations" path was taken */
} polref_result;

/* ---- host-side setup restatements ---- */

/* src/pair.cpp:1676-1725 + 313-520 (cut_respa==NULL, msmflag==0 branch). Arrays of 2^bits doubles. */
int polref_init_tables(double cut_coul, double g_ewald, double qqrd2e, int ncoultablebits,
                       double tabinner, int *ncoulmask, int *ncoulshiftbits, double *tabinnersq,
                       double *rtable, double *drtable, double *ftable, double *dftable,
                       double *ctable, double *dctable, double *etable, double *detable);

/* src/pair_lj_cut_coul_long_polarization.cpp:858-921 + src/pair.cpp:189-255,660-685.
 * epsilon/sigma/cut_lj/setflag are (ntypes+1)^2 in/out; derived tables are outputs. */
int polref_init_coeffs(int ntypes, double *epsilon, double *sigma, double *cut_lj, const int *setflag,
                       int mix_flag, int offset_flag, double cut_coul, double *cutsq, double *cut_ljsq,
                       double *lj1, double *lj2, double *lj3, double *lj4, double *offset);

/* src/KSPACE/ewald.cpp:153-162: initial (and, for kspace_style ewald, final) g_ewald estimate */
double polref_ewald_g(double accuracy_relative, double qqrd2e, double two_charge_force, double q2sum,
                      long natoms, double cutoff, double xprd, double yprd, double zprd);

/* ---- KSpace: reciprocal-space Ewald (src/KSPACE/ewald.cpp) ---- */

typedef struct {
  double g_ewald;
  int kxmax, kymax, kzmax, kmax; /* per-dimension and overall integer bounds, ewald.cpp:241-266 */
  double gsqmx;                  /* |k|^2 cutoff incl. the 1.00001 factor, ewald.cpp:267-275,311 */
  int kcount;                    /* number of half-space k-vectors, ewald.cpp:760-1026 */
} polref_ewald_plan;

/* Ewald::init + setup (ewald.cpp:87-340): g_ewald from the relative accuracy unless g_ewald_in > 0
 * (kspace_modify gewald), kmax per dimension from the rms() criterion (:343-351). */
int polref_ewald_plan_make(double accuracy_relative, double qqrd2e, double two_charge_force, double qsqsum,
                           long natoms, double cutoff, const double prd[3], double g_ewald_in,
                           polref_ewald_plan *plan);

/* Ewald::compute (ewald.cpp:357-497) with eik_dot_r (:501-680) and coeffs (:760-1026), orthogonal box, no slab
 * correction: structure factors over the half-space k set, energy incl. self and neutralising terms, forces (+=
 * into f) and virial.  Direct evaluation of exp(i k.r) (no recurrences): an independent statement of the sums. */
int polref_ewald_compute(const polref_ewald_plan *plan, int n, const double *x, const double *q,
                         const double prd[3], double qqrd2e, double *f, double *energy, double virial[6]);

/* ---- ghost atoms + half neighbor list, single process ---- */

/* src/comm_brick.cpp:164-411,712-880 (one proc, mode SINGLE, uniform layout).
 * Returns nghost (or -needed-1 if maxghost too small).  ghost_owner[g] = local index,
 * ghost_shift[3g..] = integer periodic image shift, xall gets nlocal+nghost coordinates. */
int polref_build_ghosts(int nlocal, const double *x, const double boxlo[3], const double boxhi[3],
                        const int periodic[3], double cutghost, int maxghost, double *xall,
                        int *ghost_owner, int *ghost_shift);

/* src/nbin_standard.cpp:55-234, src/nbin.cpp:116-147, src/nstencil.cpp:142-223,
 * src/nstencil_half_bin_3d_newton.cpp:27-40, src/npair_half_bin_newton.cpp:36-161,
 * src/npair.h:111-137.  special/nspecial may be NULL (atomic system).  Returns total pairs or <0. */
long polref_build_half_list(int nlocal, int nghost, const double *xall, const int *type_all,
                            const int *tag_all, const double boxlo[3], const double boxhi[3],
                            const int periodic[3], int ntypes, const double *cutneighsq,
                            double cutneighmax, double cutghost, const int *nspecial,
                            const int *special, int maxspecial, const int special_flag[4],
                            long maxpairs, int *numneigh, long *firstoffset, int *neigh);

/* ---- the hot path ---- */

/* Literal restatement of compute() (src/pair_lj_cut_coul_long_polarization.cpp:125-645) and
 * DipoleSolverIterative()/build_dipole_field_matrix() (:1113-1316) in their original loop order.
 * x,q,type,molecule,alpha: nlocal+nghost entries.  mu: nlocal*3 in/out.  f: (nlocal+nghost)*3 +=.
 * trace (optional, may be NULL): receives mu after every sweep, trace_stride = 3*nlocal doubles per
 * sweep, at most trace_max sweeps.  use_matrix!=0 materialises the 3N x 3N matrix like the reference. */
int polref_compute(const polref_params *p, int nlocal, int nghost, const double *x, const double *q,
                   const int *type, const int *molecule, const double *alpha, int inum,
                   const int *ilist, const int *numneigh, const long *firstoffset, const int *neigh,
                   double *mu, double *ef_static, double *f, int eflag, int vflag, int use_matrix,
                   polref_result *out, double *trace, int trace_max, int *ranked_out);

/* Row-gather form of the polarization part only (static field, solver, dipole forces/energies) for
 * sizes where the literal O(N^2) loops are too slow: per-atom minimum-image partner lists built
 * with cells, partners visited in ascending index order (the order in which the reference's i<j
 * scatter loops deliver contributions to each atom), OpenMP over rows.  Same semantics as
 * polref_compute when polar_cut<=0 is replaced by list_cut >= every minimum-image distance.
 * Adds its forces to f (nlocal*3) and returns energies in out (eng_vdwl/eng_coul untouched = 0). */
int polref_polar_rows(const polref_params *p, int nlocal, const double *x, const double *q,
                      const int *molecule, const double *alpha, double *mu, double *ef_static,
                      double *f, int eflag, polref_result *out, double *trace, int trace_max,
                      int nthreads);

/* One Jacobi sweep timing kernel for bench.py cpu baselines: rows [row0,row1) only. */
double polref_bench_rows(const polref_params *p, int nlocal, const double *x, const double *q,
                         const int *molecule, const double *alpha, int row0, int row1, int nsweeps,
                         int nthreads, double *checksum);

#ifdef __cplusplus
}
#endif
#endif
