/* polref.c -- CPU ORACLE (test infrastructure only; see polref.h header for status and rules).
 *
 * Plain-C restatement, on flat arrays, of the reference algorithm for the hot path of
 * pair style lj/cut/coul/long/polarization.  Citations are file:line under /root/reference.
 * Arithmetic order follows the reference expression by expression, because the 1e-10 parity
 * target of the CUDA path is measured against this file and this file is itself pinned against
 * the repaired reference binary to ~1e-13.
 */
#include "polref.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EWALD_F 1.12837917 /* src/pair_lj_cut_coul_long_polarization.cpp:43-49 */
#define EWALD_P 0.3275911
#define A1 0.254829592
#define A2 -0.284496736
#define A3 1.421413741
#define A4 -1.453152027
#define A5 1.061405429
#define MY_ISPI4 1.12837916709551257389 /* src/math_const.h:29 */

#define SBBITS 30 /* src/lmptype.h:58-59 */
#define NEIGHMASK 0x3FFFFFFF

typedef union {
  int i;
  float f;
} int_float_t;

typedef struct {
  double prd[3], half[3];
  int periodic[3];
} box_t;

static void box_init(box_t *b, const double lo[3], const double hi[3], const int per[3])
{
  for (int d = 0; d < 3; d++) {
    b->prd[d] = hi[d] - lo[d]; /* src/domain.cpp set_global_box: prd = boxhi - boxlo */
    b->half[d] = 0.5 * b->prd[d];
    b->periodic[d] = per[d];
  }
}

/* src/domain.cpp:1220-1318, orthogonal branch.  Returns the image of xj closest to xi. */
static void closest_image(const box_t *b, const double *xi, const double *xj, double *xjimage)
{
  for (int d = 0; d < 3; d++) {
    double dx = xj[d] - xi[d];
    if (b->periodic[d]) {
      if (dx < 0.0) {
        while (dx < 0.0) dx += b->prd[d];
        if (dx > b->half[d]) dx -= b->prd[d];
      } else {
        while (dx > 0.0) dx -= b->prd[d];
        if (dx < -b->half[d]) dx += b->prd[d];
      }
    }
    xjimage[d] = xi[d] + dx;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Coulomb tables: src/pair.cpp:1676-1725 (init_bitmap) and :313-520 (init_tables)             */
/* ------------------------------------------------------------------------------------------ */

static int init_bitmap(double inner, double outer, int ntablebits, int *masklo, int *maskhi,
                       int *nmask, int *nshiftbits)
{
  if (ntablebits > (int)(sizeof(float) * 8)) return -1;
  int nlowermin = 1;
  while (!((pow(2.0, (double)nlowermin) <= inner * inner) &&
           (pow(2.0, (double)nlowermin + 1.0) > inner * inner))) {
    if (pow(2.0, (double)nlowermin) <= inner * inner) nlowermin++;
    else nlowermin--;
  }
  int nexpbits = 0;
  double required_range = outer * outer / pow(2.0, (double)nlowermin);
  double available_range = 2.0;
  while (available_range < required_range) {
    nexpbits++;
    available_range = pow(2.0, pow(2.0, (double)nexpbits));
  }
  int nmantbits = ntablebits - nexpbits;
  if (nexpbits > (int)(sizeof(float) * 8) - FLT_MANT_DIG) return -2;
  if (nmantbits + 1 > FLT_MANT_DIG) return -3;
  if (nmantbits < 3) return -4;
  *nshiftbits = FLT_MANT_DIG - (nmantbits + 1);
  int m = 1;
  for (int j = 0; j < ntablebits + *nshiftbits; j++) m *= 2;
  m -= 1;
  *nmask = m;
  int_float_t u;
  u.f = outer * outer;
  *maskhi = u.i & ~m;
  u.f = inner * inner;
  *masklo = u.i & ~m;
  return 0;
}

int polref_init_tables(double cut_coul, double g_ewald, double qqrd2e, int ncoultablebits,
                       double tabinner, int *ncoulmask, int *ncoulshiftbits, double *tabinnersq_out,
                       double *rtable, double *drtable, double *ftable, double *dftable,
                       double *ctable, double *dctable, double *etable, double *detable)
{
  int masklo, maskhi;
  double cut_coulsq = cut_coul * cut_coul;
  double tabinnersq = tabinner * tabinner;
  int err = init_bitmap(tabinner, cut_coul, ncoultablebits, &masklo, &maskhi, ncoulmask,
                        ncoulshiftbits);
  if (err) return err;
  int shift = *ncoulshiftbits;
  int ntable = 1;
  for (int i = 0; i < ncoultablebits; i++) ntable *= 2;

  int_float_t rsq_lookup, minrsq_lookup;
  minrsq_lookup.i = 0 << shift;
  minrsq_lookup.i |= maskhi;
  for (int i = 0; i < ntable; i++) {
    rsq_lookup.i = i << shift;
    rsq_lookup.i |= masklo;
    if (rsq_lookup.f < tabinnersq) {
      rsq_lookup.i = i << shift;
      rsq_lookup.i |= maskhi;
    }
    double r = sqrtf(rsq_lookup.f);
    double grij = g_ewald * r;
    double expm2 = exp(-grij * grij);
    double derfc = erfc(grij);
    rtable[i] = rsq_lookup.f;
    ctable[i] = qqrd2e / r;
    ftable[i] = qqrd2e / r * (derfc + MY_ISPI4 * grij * expm2);
    etable[i] = qqrd2e / r * derfc;
    minrsq_lookup.f = (minrsq_lookup.f < rsq_lookup.f) ? minrsq_lookup.f : rsq_lookup.f;
  }
  *tabinnersq_out = minrsq_lookup.f;

  int ntablem1 = ntable - 1;
  for (int i = 0; i < ntablem1; i++) {
    drtable[i] = 1.0 / (rtable[i + 1] - rtable[i]);
    dftable[i] = ftable[i + 1] - ftable[i];
    dctable[i] = ctable[i + 1] - ctable[i];
    detable[i] = etable[i + 1] - etable[i];
  }
  drtable[ntablem1] = 1.0 / (rtable[0] - rtable[ntablem1]);
  dftable[ntablem1] = ftable[0] - ftable[ntablem1];
  dctable[ntablem1] = ctable[0] - ctable[ntablem1];
  detable[ntablem1] = etable[0] - etable[ntablem1];

  int itablemin = minrsq_lookup.i & *ncoulmask;
  itablemin >>= shift;
  int itablemax = itablemin - 1;
  if (itablemin == 0) itablemax = ntablem1;
  rsq_lookup.i = itablemax << shift;
  rsq_lookup.i |= maskhi;
  if (rsq_lookup.f < cut_coulsq) {
    rsq_lookup.f = cut_coulsq;
    double r = sqrtf(rsq_lookup.f);
    double grij = g_ewald * r;
    double expm2 = exp(-grij * grij);
    double derfc = erfc(grij);
    double c_tmp = qqrd2e / r;
    double f_tmp = qqrd2e / r * (derfc + MY_ISPI4 * grij * expm2);
    double e_tmp = qqrd2e / r * derfc;
    drtable[itablemax] = 1.0 / (rsq_lookup.f - rtable[itablemax]);
    dftable[itablemax] = f_tmp - ftable[itablemax];
    dctable[itablemax] = c_tmp - ctable[itablemax];
    detable[itablemax] = e_tmp - etable[itablemax];
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* coefficients: src/pair.cpp:189-255 (Pair::init loop), :660-685 (mixing),                   */
/* src/pair_lj_cut_coul_long_polarization.cpp:858-921 (init_one, qdist = 0, no tail)           */
/* ------------------------------------------------------------------------------------------ */

static double mix_energy(int mix, double e1, double e2, double s1, double s2)
{
  if (mix == POLREF_MIX_GEOMETRIC || mix == POLREF_MIX_ARITHMETIC) return sqrt(e1 * e2);
  return 2.0 * sqrt(e1 * e2) * pow(s1, 3.0) * pow(s2, 3.0) / (pow(s1, 6.0) + pow(s2, 6.0));
}

static double mix_distance(int mix, double s1, double s2)
{
  if (mix == POLREF_MIX_GEOMETRIC) return sqrt(s1 * s2);
  if (mix == POLREF_MIX_ARITHMETIC) return 0.5 * (s1 + s2);
  return pow(0.5 * (pow(s1, 6.0) + pow(s2, 6.0)), 1.0 / 6.0);
}

int polref_init_coeffs(int ntypes, double *epsilon, double *sigma, double *cut_lj, const int *setflag,
                       int mix_flag, int offset_flag, double cut_coul, double *cutsq, double *cut_ljsq,
                       double *lj1, double *lj2, double *lj3, double *lj4, double *offset)
{
  int n1 = ntypes + 1;
  for (int i = 1; i <= ntypes; i++)
    if (!setflag[i * n1 + i]) return -1; /* "All pair coeffs are not set" */
  for (int i = 1; i <= ntypes; i++)
    for (int j = i; j <= ntypes; j++) {
      int ij = i * n1 + j, ji = j * n1 + i, ii = i * n1 + i, jj = j * n1 + j;
      if (!setflag[ij]) {
        epsilon[ij] = mix_energy(mix_flag, epsilon[ii], epsilon[jj], sigma[ii], sigma[jj]);
        sigma[ij] = mix_distance(mix_flag, sigma[ii], sigma[jj]);
        cut_lj[ij] = mix_distance(mix_flag, cut_lj[ii], cut_lj[jj]);
      }
      double cut = (cut_lj[ij] > cut_coul) ? cut_lj[ij] : cut_coul;
      cut_ljsq[ij] = cut_lj[ij] * cut_lj[ij];
      lj1[ij] = 48.0 * epsilon[ij] * pow(sigma[ij], 12.0);
      lj2[ij] = 24.0 * epsilon[ij] * pow(sigma[ij], 6.0);
      lj3[ij] = 4.0 * epsilon[ij] * pow(sigma[ij], 12.0);
      lj4[ij] = 4.0 * epsilon[ij] * pow(sigma[ij], 6.0);
      if (offset_flag && (cut_lj[ij] > 0.0)) {
        double ratio = sigma[ij] / cut_lj[ij];
        offset[ij] = 4.0 * epsilon[ij] * (pow(ratio, 12.0) - pow(ratio, 6.0));
      } else offset[ij] = 0.0;
      cut_ljsq[ji] = cut_ljsq[ij];
      lj1[ji] = lj1[ij];
      lj2[ji] = lj2[ij];
      lj3[ji] = lj3[ij];
      lj4[ji] = lj4[ij];
      offset[ji] = offset[ij];
      cutsq[ij] = cutsq[ji] = cut * cut;
    }
  return 0;
}

double polref_ewald_g(double accuracy_relative, double qqrd2e, double two_charge_force, double q2sum,
                      long natoms, double cutoff, double xprd, double yprd, double zprd)
{
  /* src/KSPACE/ewald.cpp:133-134,153-162; src/kspace.cpp:294 (q2 = qsqsum*qqrd2e) */
  double accuracy = accuracy_relative * two_charge_force;
  double q2 = q2sum * qqrd2e;
  double g = accuracy * sqrt(natoms * cutoff * xprd * yprd * zprd) / (2.0 * q2);
  if (g >= 1.0) g = (1.35 - 0.15 * log(accuracy)) / cutoff;
  else g = sqrt(-log(g)) / cutoff;
  return g;
}

/* ------------------------------------------------------------------------------------------ */
/* ghost atoms, one process: src/comm_brick.cpp:164-411 (setup) and :712-880 (borders)          */
/* ------------------------------------------------------------------------------------------ */

int polref_build_ghosts(int nlocal, const double *x, const double boxlo[3], const double boxhi[3],
                        const int periodic[3], double cutghost, int maxghost, double *xall,
                        int *ghost_owner, int *ghost_shift)
{
  const double BIG = 1.0e20;
  int nall = nlocal;
  int overflow = 0;
  memcpy(xall, x, sizeof(double) * 3 * (size_t)nlocal);
  /* owner/shift for every atom in xall so that ghosts of ghosts resolve to a local owner */
  int cap = nlocal + maxghost;
  int *own = (int *)malloc(sizeof(int) * (size_t)cap);
  int *shf = (int *)malloc(sizeof(int) * 3 * (size_t)cap);
  for (int i = 0; i < nlocal; i++) {
    own[i] = i;
    shf[3 * i] = shf[3 * i + 1] = shf[3 * i + 2] = 0;
  }
  for (int dim = 0; dim < 3; dim++) {
    double prd = boxhi[dim] - boxlo[dim];
    /* comm_brick.cpp:239-244: maxneed = int(cutghost*procgrid/prd)+1, procgrid = 1 */
    int maxneed = (int)(cutghost * 1 / prd) + 1;
    if (!periodic[dim]) maxneed = (maxneed < 0) ? maxneed : 0; /* MIN(maxneed,procgrid-1) */
    int nfirst = 0, nlast = 0;
    for (int ineed = 0; ineed < 2 * maxneed; ineed++) {
      double lo, hi;
      int pbc;
      if (ineed % 2 == 0) { /* comm_brick.cpp:359-381 */
        lo = (ineed < 2) ? -BIG : 0.5 * (boxlo[dim] + boxhi[dim]);
        hi = boxlo[dim] + cutghost;
        pbc = 1;
        nfirst = nlast; /* comm_brick.cpp:748-751 */
        nlast = nall;
      } else { /* comm_brick.cpp:383-405 */
        lo = boxhi[dim] - cutghost;
        hi = (ineed < 2) ? BIG : 0.5 * (boxlo[dim] + boxhi[dim]);
        pbc = -1;
      }
      int nstart = nall;
      for (int i = nfirst; i < nlast; i++) {
        double xi = xall[3 * i + dim];
        if (xi >= lo && xi <= hi) { /* comm_brick.cpp:768-772 */
          if (nall - nlocal >= maxghost) {
            overflow++;
            continue;
          }
          int g = nall++;
          /* atom_vec_full.cpp:403-421 pack_border with pbc: x + pbc*prd */
          for (int d = 0; d < 3; d++) xall[3 * g + d] = xall[3 * i + d];
          xall[3 * g + dim] = xall[3 * i + dim] + pbc * prd;
          own[g] = own[i];
          shf[3 * g] = shf[3 * i];
          shf[3 * g + 1] = shf[3 * i + 1];
          shf[3 * g + 2] = shf[3 * i + 2];
          shf[3 * g + dim] += pbc;
        }
      }
      (void)nstart;
    }
  }
  int nghost = nall - nlocal;
  for (int g = 0; g < nghost; g++) {
    ghost_owner[g] = own[nlocal + g];
    for (int d = 0; d < 3; d++) ghost_shift[3 * g + d] = shf[3 * (nlocal + g) + d];
  }
  free(own);
  free(shf);
  if (overflow) return -(nghost + overflow) - 1;
  return nghost;
}

/* ------------------------------------------------------------------------------------------ */
/* binning + stencil + half/bin/newton pair build                                              */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  int nbinx, nbiny, nbinz, mbinx, mbiny, mbinz, mbinxlo, mbinylo, mbinzlo, mbins;
  double binsizex, binsizey, binsizez, bininvx, bininvy, bininvz;
  double bboxlo[3], bboxhi[3];
} bins_t;

/* src/nbin_standard.cpp:55-188 (orthogonal, style BIN, no user binsize) */
static int setup_bins(bins_t *b, const double boxlo[3], const double boxhi[3], double cutneighmax,
                      double cutghost)
{
  const double SMALL = 1.0e-6;
  double bbox[3], bsubboxlo[3], bsubboxhi[3];
  for (int d = 0; d < 3; d++) {
    b->bboxlo[d] = boxlo[d];
    b->bboxhi[d] = boxhi[d];
    bsubboxlo[d] = boxlo[d] - cutghost;
    bsubboxhi[d] = boxhi[d] + cutghost;
    bbox[d] = boxhi[d] - boxlo[d];
  }
  double binsize_optimal = 0.5 * cutneighmax;
  if (binsize_optimal == 0.0) binsize_optimal = bbox[0];
  double binsizeinv = 1.0 / binsize_optimal;
  b->nbinx = (int)(bbox[0] * binsizeinv);
  b->nbiny = (int)(bbox[1] * binsizeinv);
  b->nbinz = (int)(bbox[2] * binsizeinv);
  if (b->nbinx == 0) b->nbinx = 1;
  if (b->nbiny == 0) b->nbiny = 1;
  if (b->nbinz == 0) b->nbinz = 1;
  b->binsizex = bbox[0] / b->nbinx;
  b->binsizey = bbox[1] / b->nbiny;
  b->binsizez = bbox[2] / b->nbinz;
  b->bininvx = 1.0 / b->binsizex;
  b->bininvy = 1.0 / b->binsizey;
  b->bininvz = 1.0 / b->binsizez;

  int mbinxhi, mbinyhi, mbinzhi;
  double coord;
  coord = bsubboxlo[0] - SMALL * bbox[0];
  b->mbinxlo = (int)((coord - b->bboxlo[0]) * b->bininvx);
  if (coord < b->bboxlo[0]) b->mbinxlo = b->mbinxlo - 1;
  coord = bsubboxhi[0] + SMALL * bbox[0];
  mbinxhi = (int)((coord - b->bboxlo[0]) * b->bininvx);
  coord = bsubboxlo[1] - SMALL * bbox[1];
  b->mbinylo = (int)((coord - b->bboxlo[1]) * b->bininvy);
  if (coord < b->bboxlo[1]) b->mbinylo = b->mbinylo - 1;
  coord = bsubboxhi[1] + SMALL * bbox[1];
  mbinyhi = (int)((coord - b->bboxlo[1]) * b->bininvy);
  coord = bsubboxlo[2] - SMALL * bbox[2];
  b->mbinzlo = (int)((coord - b->bboxlo[2]) * b->bininvz);
  if (coord < b->bboxlo[2]) b->mbinzlo = b->mbinzlo - 1;
  coord = bsubboxhi[2] + SMALL * bbox[2];
  mbinzhi = (int)((coord - b->bboxlo[2]) * b->bininvz);

  b->mbinxlo -= 1;
  mbinxhi += 1;
  b->mbinx = mbinxhi - b->mbinxlo + 1;
  b->mbinylo -= 1;
  mbinyhi += 1;
  b->mbiny = mbinyhi - b->mbinylo + 1;
  b->mbinzlo -= 1;
  mbinzhi += 1;
  b->mbinz = mbinzhi - b->mbinzlo + 1;
  double bbin = (double)b->mbinx * b->mbiny * b->mbinz + 1;
  if (bbin > 2147483647.0) return -1;
  b->mbins = (int)bbin;
  return 0;
}

/* src/nbin.cpp:116-147 */
static int coord2bin(const bins_t *b, const double *x)
{
  int ix, iy, iz;
  if (x[0] >= b->bboxhi[0]) ix = (int)((x[0] - b->bboxhi[0]) * b->bininvx) + b->nbinx;
  else if (x[0] >= b->bboxlo[0]) {
    ix = (int)((x[0] - b->bboxlo[0]) * b->bininvx);
    if (ix > b->nbinx - 1) ix = b->nbinx - 1;
  } else ix = (int)((x[0] - b->bboxlo[0]) * b->bininvx) - 1;
  if (x[1] >= b->bboxhi[1]) iy = (int)((x[1] - b->bboxhi[1]) * b->bininvy) + b->nbiny;
  else if (x[1] >= b->bboxlo[1]) {
    iy = (int)((x[1] - b->bboxlo[1]) * b->bininvy);
    if (iy > b->nbiny - 1) iy = b->nbiny - 1;
  } else iy = (int)((x[1] - b->bboxlo[1]) * b->bininvy) - 1;
  if (x[2] >= b->bboxhi[2]) iz = (int)((x[2] - b->bboxhi[2]) * b->bininvz) + b->nbinz;
  else if (x[2] >= b->bboxlo[2]) {
    iz = (int)((x[2] - b->bboxlo[2]) * b->bininvz);
    if (iz > b->nbinz - 1) iz = b->nbinz - 1;
  } else iz = (int)((x[2] - b->bboxlo[2]) * b->bininvz) - 1;
  return (iz - b->mbinzlo) * b->mbiny * b->mbinx + (iy - b->mbinylo) * b->mbinx + (ix - b->mbinxlo);
}

/* src/nstencil.cpp:205-223 */
static double bin_distance(const bins_t *b, int i, int j, int k)
{
  double delx, dely, delz;
  if (i > 0) delx = (i - 1) * b->binsizex;
  else if (i == 0) delx = 0.0;
  else delx = (i + 1) * b->binsizex;
  if (j > 0) dely = (j - 1) * b->binsizey;
  else if (j == 0) dely = 0.0;
  else dely = (j + 1) * b->binsizey;
  if (k > 0) delz = (k - 1) * b->binsizez;
  else if (k == 0) delz = 0.0;
  else delz = (k + 1) * b->binsizez;
  return delx * delx + dely * dely + delz * delz;
}

/* src/npair.h:111-137 */
static int find_special(const int *list, const int *nspecial, int tag, const int special_flag[4])
{
  int n1 = nspecial[0], n2 = nspecial[1], n3 = nspecial[2];
  for (int i = 0; i < n3; i++) {
    if (list[i] == tag) {
      int lvl = (i < n1) ? 1 : (i < n2) ? 2 : 3;
      if (special_flag[lvl] == 0) return -1;
      if (special_flag[lvl] == 1) return 0;
      return lvl;
    }
  }
  return 0;
}

long polref_build_half_list(int nlocal, int nghost, const double *x, const int *type, const int *tag,
                            const double boxlo[3], const double boxhi[3], const int periodic[3],
                            int ntypes, const double *cutneighsq, double cutneighmax,
                            double cutghost, const int *nspecial, const int *special,
                            int maxspecial, const int special_flag[4], long maxpairs,
                            int *numneigh, long *firstoffset, int *neigh)
{
  bins_t b;
  if (setup_bins(&b, boxlo, boxhi, cutneighmax, cutghost)) return -1;
  int nall = nlocal + nghost;
  int n1 = ntypes + 1;
  double half[3];
  for (int d = 0; d < 3; d++) half[d] = 0.5 * (boxhi[d] - boxlo[d]);

  /* stencil: src/nstencil.cpp:142-160 + src/nstencil_half_bin_3d_newton.cpp:27-40 */
  double cutneighmaxsq = cutneighmax * cutneighmax;
  int sx = (int)(cutneighmax * b.bininvx);
  if (sx * b.binsizex < cutneighmax) sx++;
  int sy = (int)(cutneighmax * b.bininvy);
  if (sy * b.binsizey < cutneighmax) sy++;
  int sz = (int)(cutneighmax * b.bininvz);
  if (sz * b.binsizez < cutneighmax) sz++;
  int smax = (2 * sx + 1) * (2 * sy + 1) * (2 * sz + 1);
  int *stencil = (int *)malloc(sizeof(int) * (size_t)smax);
  int nstencil = 0;
  for (int k = 0; k <= sz; k++)
    for (int j = -sy; j <= sy; j++)
      for (int i = -sx; i <= sx; i++)
        if (k > 0 || j > 0 || (j == 0 && i > 0))
          if (bin_distance(&b, i, j, k) < cutneighmaxsq)
            stencil[nstencil++] = k * b.mbiny * b.mbinx + j * b.mbinx + i;

  /* bin atoms: src/nbin_standard.cpp:194-234 (reverse order => forward linked lists, ghosts last) */
  int *binhead = (int *)malloc(sizeof(int) * (size_t)b.mbins);
  int *bins = (int *)malloc(sizeof(int) * (size_t)nall);
  int *atom2bin = (int *)malloc(sizeof(int) * (size_t)nall);
  for (int i = 0; i < b.mbins; i++) binhead[i] = -1;
  for (int i = nall - 1; i >= 0; i--) {
    int ibin = coord2bin(&b, &x[3 * i]);
    atom2bin[i] = ibin;
    bins[i] = binhead[ibin];
    binhead[ibin] = i;
  }

  long npairs = 0;
  int molecular = (special != NULL && nspecial != NULL);
  for (int i = 0; i < nlocal; i++) {
    int n = 0;
    firstoffset[i] = npairs;
    int itype = type[i];
    double xtmp = x[3 * i], ytmp = x[3 * i + 1], ztmp = x[3 * i + 2];
    /* two passes share the acceptance code: pass 0 = own bin, pass 1..nstencil = stencil bins */
    for (int k = -1; k < nstencil; k++) {
      int j = (k < 0) ? bins[i] : binhead[atom2bin[i] + stencil[k]];
      for (; j >= 0; j = bins[j]) {
        if (k < 0 && j >= nlocal) { /* src/npair_half_bin_newton.cpp:85-92 */
          if (x[3 * j + 2] < ztmp) continue;
          if (x[3 * j + 2] == ztmp) {
            if (x[3 * j + 1] < ytmp) continue;
            if (x[3 * j + 1] == ytmp && x[3 * j] < xtmp) continue;
          }
        }
        int jtype = type[j];
        double delx = xtmp - x[3 * j];
        double dely = ytmp - x[3 * j + 1];
        double delz = ztmp - x[3 * j + 2];
        double rsq = delx * delx + dely * dely + delz * delz;
        if (rsq <= cutneighsq[itype * n1 + jtype]) {
          int entry = -1;
          if (molecular) {
            int which = find_special(&special[(size_t)i * maxspecial], &nspecial[3 * i], tag[j],
                                     special_flag);
            if (which == 0) entry = j;
            else if ((periodic[0] && fabs(delx) > half[0]) || (periodic[1] && fabs(dely) > half[1]) ||
                     (periodic[2] && fabs(delz) > half[2])) /* src/domain.h:155-160 */
              entry = j;
            else if (which > 0) entry = j ^ (which << SBBITS);
          } else entry = j;
          if (entry != -1 || !molecular) {
            if (npairs + n >= maxpairs) {
              free(stencil); free(binhead); free(bins); free(atom2bin);
              return -2;
            }
            neigh[npairs + n] = entry;
            n++;
          }
        }
      }
    }
    numneigh[i] = n;
    npairs += n;
  }
  free(stencil);
  free(binhead);
  free(bins);
  free(atom2bin);
  return npairs;
}

/* ------------------------------------------------------------------------------------------ */
/* pieces shared by the literal and the row-gather forms                                       */
/* ------------------------------------------------------------------------------------------ */

/* one T_ij block in the orientation of the reference's matrix build
 * (src/pair_lj_cut_coul_long_polarization.cpp:1279-1306): xi = lower index, xj = higher index. */
static void t_block(const polref_params *p, const box_t *box, const double *xi, const double *xj,
                    double T[3][3], double *r2_out)
{
  double xjimage[3];
  closest_image(box, xi, xj, xjimage);
  double r2 = pow(xi[0] - xjimage[0], 2) + pow(xi[1] - xjimage[1], 2) + pow(xi[2] - xjimage[2], 2);
  double r = sqrt(r2), r3, r5;
  double damping_term1 = 1.0, damping_term2 = 1.0;
  double a = p->polar_damp;
  if (r == 0.0) r3 = r5 = DBL_MAX;
  else {
    r3 = 1.0 / (r * r * r);
    r5 = 1.0 / (r * r * r * r * r);
  }
  if (p->damping_type == POLREF_DAMP_EXPONENTIAL) {
    damping_term1 = 1.0 - exp(-a * r) * (0.5 * a * a * r2 + a * r + 1.0);
    damping_term2 = 1.0 - exp(-a * r) * (a * a * a * r2 * r / 6.0 + 0.5 * a * a * r2 + a * r + 1.0);
  }
  for (int pp = 0; pp < 3; pp++)
    for (int qq = 0; qq < 3; qq++) {
      T[pp][qq] = -3.0 * (xi[pp] - xjimage[pp]) * (xi[qq] - xjimage[qq]) * damping_term2 * r5;
      if (pp == qq) T[pp][qq] += damping_term1 * r3;
    }
  if (r2_out) *r2_out = r2;
}

/* stable descending order by metric == the reference's bubble sort with strict '<'
 * (src/pair_lj_cut_coul_long_polarization.cpp:1130-1143), realised as a merge sort */
static void stable_rank(int n, const double *metric, int *order)
{
  int *tmp = (int *)malloc(sizeof(int) * (size_t)n);
  for (int i = 0; i < n; i++) order[i] = i;
  for (int w = 1; w < n; w *= 2) {
    for (int lo = 0; lo < n; lo += 2 * w) {
      int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      int a = lo, b = mid, k = lo;
      while (a < mid && b < hi) {
        /* take from the right run only if strictly greater: keeps ties in original order */
        if (metric[order[b]] > metric[order[a]]) tmp[k++] = order[b++];
        else tmp[k++] = order[a++];
      }
      while (a < mid) tmp[k++] = order[a++];
      while (b < hi) tmp[k++] = order[b++];
    }
    memcpy(order, tmp, sizeof(int) * (size_t)n);
  }
  free(tmp);
}

static void bubble_rank(int n, const double *metric, int *order)
{
  /* src/pair_lj_cut_coul_long_polarization.cpp:1127-1143, literal */
  for (int i = 0; i < n; i++) order[i] = i;
  for (int i = 0; i < n; i++) {
    int sorted = 1;
    for (int j = 0; j < n - 1; j++) {
      if (metric[order[j]] < metric[order[j + 1]]) {
        sorted = 0;
        int t = order[j];
        order[j] = order[j + 1];
        order[j + 1] = t;
      }
    }
    if (sorted) break;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* literal compute()                                                                            */
/* ------------------------------------------------------------------------------------------ */

int polref_compute(const polref_params *p, int nlocal, int nghost, const double *x, const double *q,
                   const int *type, const int *molecule, const double *alpha, int inum,
                   const int *ilist, const int *numneigh, const long *firstoffset, const int *neigh,
                   double *mu, double *ef_static, double *f, int eflag, int vflag, int use_matrix,
                   polref_result *out, double *trace, int trace_max, int *ranked_out)
{
  int ntotal = nlocal + nghost;
  int n1 = p->ntypes + 1;
  box_t box;
  box_init(&box, p->boxlo, p->boxhi, p->periodic);
  memset(out, 0, sizeof(*out));
  double cut_coulsq = p->cut_coul * p->cut_coul;
  double polar_cutsq = (p->polar_cut > 0.0) ? p->polar_cut * p->polar_cut : -1.0;
  int newton_pair = 1;

  /* :150-156 */
  for (int i = 0; i < nlocal; i++) ef_static[3 * i] = ef_static[3 * i + 1] = ef_static[3 * i + 2] = 0;

  /* ev_setup (src/pair.cpp:752-833): eflag_global = eflag%2; vflag_global = vflag%4;
   * vflag_global==2 => virial via F.r after the force loops, pairwise tallies off */
  int eflag_global = eflag % 2;
  int vflag_global = vflag % 4;
  int vflag_fdotr = 0;
  if (vflag_global == 2) {
    vflag_fdotr = 1;
    vflag_global = 0;
  }
  double eng_vdwl = 0.0, eng_coul = 0.0;
  double virial[6] = {0, 0, 0, 0, 0, 0};

  double *rank_metric = (double *)calloc((size_t)(nlocal > 0 ? nlocal : 1), sizeof(double));
  int *ranked = (int *)malloc(sizeof(int) * (size_t)(nlocal > 0 ? nlocal : 1));
  double rmin = 1000.0;

  /* :192-227 rank metric over local+ghost raw coordinates */
  if (p->polar_gs_ranked) {
    for (int i = 0; i < nlocal; i++)
      for (int j = 0; j < ntotal; j++)
        if (i != j) {
          double r = sqrt(pow(x[3 * i] - x[3 * j], 2) + pow(x[3 * i + 1] - x[3 * j + 1], 2) +
                          pow(x[3 * i + 2] - x[3 * j + 2], 2));
          if (alpha[i] > 0 && alpha[j] > 0 && rmin > r &&
              ((molecule[i] != molecule[j]) || molecule[i] == 0))
            rmin = r;
        }
    for (int i = 0; i < nlocal; i++)
      for (int j = 0; j < ntotal; j++)
        if (i != j) {
          double r = sqrt(pow(x[3 * i] - x[3 * j], 2) + pow(x[3 * i + 1] - x[3 * j + 1], 2) +
                          pow(x[3 * i + 2] - x[3 * j + 2], 2));
          if (rmin * 1.5 > r && ((molecule[i] != molecule[j]) || molecule[i] == 0))
            rank_metric[i] += alpha[i] * alpha[j];
        }
  }
  out->rmin = rmin;

  /* :232-321 LJ + real-space Ewald over the half list */
  for (int ii = 0; ii < inum; ii++) {
    int i = ilist ? ilist[ii] : ii;
    double qtmp = q[i], xtmp = x[3 * i], ytmp = x[3 * i + 1], ztmp = x[3 * i + 2];
    int itype = type[i];
    const int *jlist = &neigh[firstoffset[i]];
    int jnum = numneigh[i];
    for (int jj = 0; jj < jnum; jj++) {
      int j = jlist[jj];
      int sb = (j >> SBBITS) & 3;
      double factor_lj = p->special_lj[sb];
      double factor_coul = p->special_coul[sb];
      j &= NEIGHMASK;
      double delx = xtmp - x[3 * j], dely = ytmp - x[3 * j + 1], delz = ztmp - x[3 * j + 2];
      double rsq = delx * delx + dely * dely + delz * delz;
      int jtype = type[j];
      if (rsq < p->cutsq[itype * n1 + jtype]) {
        double r2inv = 1.0 / rsq;
        double forcecoul, forcelj, r6inv = 0.0, prefactor = 0.0, erfc_ = 0.0, fraction = 0.0;
        int itable = 0;
        if (rsq < cut_coulsq) {
          if (!p->ncoultablebits || rsq <= p->tabinnersq) {
            double r = sqrt(rsq);
            double grij = p->g_ewald * r;
            double expm2 = exp(-grij * grij);
            double t = 1.0 / (1.0 + EWALD_P * grij);
            erfc_ = t * (A1 + t * (A2 + t * (A3 + t * (A4 + t * A5)))) * expm2;
            prefactor = p->qqrd2e * qtmp * q[j] / r;
            forcecoul = prefactor * (erfc_ + EWALD_F * grij * expm2);
            if (factor_coul < 1.0) forcecoul -= (1.0 - factor_coul) * prefactor;
          } else {
            int_float_t rsq_lookup;
            rsq_lookup.f = rsq;
            itable = rsq_lookup.i & p->ncoulmask;
            itable >>= p->ncoulshiftbits;
            fraction = (rsq_lookup.f - p->rtable[itable]) * p->drtable[itable];
            double table = p->ftable[itable] + fraction * p->dftable[itable];
            forcecoul = qtmp * q[j] * table;
            if (factor_coul < 1.0) {
              table = p->ctable[itable] + fraction * p->dctable[itable];
              prefactor = qtmp * q[j] * table;
              forcecoul -= (1.0 - factor_coul) * prefactor;
            }
          }
        } else forcecoul = 0.0;
        if (rsq < p->cut_ljsq[itype * n1 + jtype]) {
          r6inv = r2inv * r2inv * r2inv;
          forcelj = r6inv * (p->lj1[itype * n1 + jtype] * r6inv - p->lj2[itype * n1 + jtype]);
        } else forcelj = 0.0;
        double fpair = (forcecoul + factor_lj * forcelj) * r2inv;
        f[3 * i] += delx * fpair;
        f[3 * i + 1] += dely * fpair;
        f[3 * i + 2] += delz * fpair;
        if (newton_pair || j < nlocal) {
          f[3 * j] -= delx * fpair;
          f[3 * j + 1] -= dely * fpair;
          f[3 * j + 2] -= delz * fpair;
        }
        double ecoul = 0.0, evdwl = 0.0;
        if (eflag) {
          if (rsq < cut_coulsq) {
            if (!p->ncoultablebits || rsq <= p->tabinnersq) ecoul = prefactor * erfc_;
            else {
              double table = p->etable[itable] + fraction * p->detable[itable];
              ecoul = qtmp * q[j] * table;
            }
            if (factor_coul < 1.0) ecoul -= (1.0 - factor_coul) * prefactor;
          } else ecoul = 0.0;
          if (rsq < p->cut_ljsq[itype * n1 + jtype]) {
            evdwl = r6inv * (p->lj3[itype * n1 + jtype] * r6inv - p->lj4[itype * n1 + jtype]) -
                    p->offset[itype * n1 + jtype];
            evdwl *= factor_lj;
          } else evdwl = 0.0;
        }
        /* ev_tally, newton_pair on (src/pair.cpp:854-949) */
        if (eflag_global) {
          eng_vdwl += evdwl;
          eng_coul += ecoul;
        }
        if (vflag_global) {
          virial[0] += delx * delx * fpair;
          virial[1] += dely * dely * fpair;
          virial[2] += delz * delz * fpair;
          virial[3] += delx * dely * fpair;
          virial[4] += delx * delz * fpair;
          virial[5] += dely * delz * fpair;
        }
        if (p->eatom) { /* src/pair.cpp:888-892 */
          double epairhalf = 0.5 * (evdwl + ecoul);
          p->eatom[i] += epairhalf;
          p->eatom[j] += epairhalf;
        }
        if (p->vatom) { /* src/pair.cpp:894-947 */
          double v[6] = {delx * delx * fpair, dely * dely * fpair, delz * delz * fpair,
                         delx * dely * fpair, delx * delz * fpair, dely * delz * fpair};
          for (int k = 0; k < 6; k++) {
            p->vatom[6 * i + k] += 0.5 * v[k];
            p->vatom[6 * j + k] += 0.5 * v[k];
          }
        }
      }
    }
  }

  /* :324-361 static field, all i<j<nlocal minimum-image pairs */
  double f_shift = -1.0 / (p->cut_coul * p->cut_coul);
  double xjimage[3];
  for (int i = 0; i < nlocal; i++) {
    double qtmp = q[i], xtmp = x[3 * i], ytmp = x[3 * i + 1], ztmp = x[3 * i + 2];
    for (int j = i + 1; j < nlocal; j++) {
      closest_image(&box, &x[3 * i], &x[3 * j], xjimage);
      double delx = xtmp - xjimage[0], dely = ytmp - xjimage[1], delz = ztmp - xjimage[2];
      double rsq = delx * delx + dely * dely + delz * delz;
      if (rsq <= cut_coulsq) {
        if ((molecule[i] != molecule[j]) || molecule[i] == 0) {
          double r = sqrt(rsq);
          double dvdrr = 1.0 / rsq + f_shift;
          double ef_temp = dvdrr * 1.0 / r;
          ef_static[3 * i] += ef_temp * q[j] * delx;
          ef_static[3 * i + 1] += ef_temp * q[j] * dely;
          ef_static[3 * i + 2] += ef_temp * q[j] * delz;
          ef_static[3 * j] -= ef_temp * qtmp * delx;
          ef_static[3 * j + 1] -= ef_temp * qtmp * dely;
          ef_static[3 * j + 2] -= ef_temp * qtmp * delz;
        }
      }
    }
  }

  /* :367-386 */
  double kq = sqrt(p->qqrd2e);
  for (int i = 0; i < nlocal; i++) {
    ef_static[3 * i] = ef_static[3 * i] * kq;
    ef_static[3 * i + 1] = ef_static[3 * i + 1] * kq;
    ef_static[3 * i + 2] = ef_static[3 * i + 2] * kq;
    if (!p->use_previous) {
      for (int c = 0; c < 3; c++) {
        mu[3 * i + c] = alpha[i] * ef_static[3 * i + c];
        mu[3 * i + c] *= p->polar_gamma;
      }
    }
  }

  /* :389 DipoleSolverIterative (:1113-1238) */
  int iterations = 0;
  if (!p->zodid) {
    double *mu_new = (double *)malloc(sizeof(double) * 3 * (size_t)nlocal);
    double *mu_old = (double *)malloc(sizeof(double) * 3 * (size_t)nlocal);
    double *ef_ind = (double *)malloc(sizeof(double) * 3 * (size_t)nlocal);
    double *M = NULL;
    size_t ld = 3 * (size_t)nlocal;
    if (use_matrix) {
      /* build_dipole_field_matrix :1243-1316 */
      M = (double *)calloc(ld * ld, sizeof(double));
      for (int i = 0; i < nlocal; i++)
        for (int pp = 0; pp < 3; pp++)
          M[(3 * (size_t)i + pp) * ld + 3 * (size_t)i + pp] = (alpha[i] != 0.0) ? 1.0 / alpha[i] : DBL_MAX;
      for (int i = 0; i < nlocal - 1; i++)
        for (int j = i + 1; j < nlocal; j++) {
          double T[3][3], r2;
          t_block(p, &box, &x[3 * i], &x[3 * j], T, &r2);
          if (polar_cutsq > 0.0 && !(r2 < polar_cutsq)) continue;
          for (int pp = 0; pp < 3; pp++)
            for (int qq = 0; qq < 3; qq++) {
              M[(3 * (size_t)i + pp) * ld + 3 * (size_t)j + qq] = T[pp][qq];
              M[(3 * (size_t)j + pp) * ld + 3 * (size_t)i + qq] = T[pp][qq];
            }
        }
    }
    if (p->polar_gs_ranked) bubble_rank(nlocal, rank_metric, ranked);
    else
      for (int i = 0; i < nlocal; i++) ranked[i] = i;

    int gs = p->polar_gs || p->polar_gs_ranked;
    int nchunks = (gs && p->gs_chunks > 0) ? p->gs_chunks : 0;
    int keep_iterating = 1;
    while (keep_iterating) {
      for (int i = 0; i < 3 * nlocal; i++) {
        mu_old[i] = mu[i];
        ef_ind[i] = 0;
      }
      /* :1158-1180.  Blocks of ranked positions: Jacobi = one block that is never written back
       * here; reference Gauss-Seidel = blocks of one atom (immediate write-back, :1176-1177);
       * gs_chunks extension = nchunks blocks, written back when the block is complete. */
      int nblk = !gs ? 1 : (nchunks ? nchunks : nlocal);
      for (int c = 0; c < nblk; c++) {
        int beg = (int)(((long)c * nlocal) / nblk), end = (int)(((long)(c + 1) * nlocal) / nblk);
        for (int i = beg; i < end; i++) {
          int index = ranked[i];
          for (int j = 0; j < nlocal; j++) {
            if (index != j) {
              if (M) {
                for (int pp = 0; pp < 3; pp++)
                  for (int qq = 0; qq < 3; qq++)
                    ef_ind[3 * index + pp] -=
                        M[(3 * (size_t)index + pp) * ld + 3 * (size_t)j + qq] * mu[3 * j + qq];
              } else {
                int lo = index < j ? index : j, hi = index < j ? j : index;
                double T[3][3], r2;
                t_block(p, &box, &x[3 * lo], &x[3 * hi], T, &r2);
                if (polar_cutsq > 0.0 && !(r2 < polar_cutsq)) continue;
                for (int pp = 0; pp < 3; pp++)
                  for (int qq = 0; qq < 3; qq++) ef_ind[3 * index + pp] -= T[pp][qq] * mu[3 * j + qq];
              }
            }
          }
          for (int pp = 0; pp < 3; pp++)
            mu_new[3 * index + pp] = alpha[index] * (ef_static[3 * index + pp] + ef_ind[3 * index + pp]);
        }
        if (gs)
          for (int i = beg; i < end; i++)
            for (int pp = 0; pp < 3; pp++) mu[3 * ranked[i] + pp] = mu_new[3 * ranked[i] + pp];
      }
      if (trace && iterations < trace_max) {
        /* state the reference would print under `debug yes` (:1183-1191): mu after the sweep,
         * i.e. mu_new for Jacobi (pre-copy mu is still old there, so record mu_new instead) */
        double *dst = trace + (size_t)iterations * 3 * nlocal;
        for (int i = 0; i < 3 * nlocal; i++) dst[i] = gs ? mu[i] : mu_new[i];
      }
      if (p->fixed_iteration == 0) {
        keep_iterating = 0;
        double change = 0;
        for (int i = 0; i < nlocal; i++)
          for (int pp = 0; pp < 3; pp++)
            change += (mu_new[3 * i + pp] - mu_old[3 * i + pp]) * (mu_new[3 * i + pp] - mu_old[3 * i + pp]);
        change /= (double)(nlocal)*3.0;
        if (change > p->polar_precision * p->polar_precision) keep_iterating = 1;
      } else {
        if (iterations >= p->iterations_max) break; /* :1214 return before the copy */
      }
      for (int i = 0; i < 3 * nlocal; i++) mu[i] = mu_new[i];
      iterations++;
      if (iterations > p->iterations_max) { /* :1227-1235 */
        for (int i = 0; i < nlocal; i++)
          for (int pp = 0; pp < 3; pp++) mu[3 * i + pp] = alpha[i] * ef_static[3 * i + pp];
        out->diverged = 1;
        break;
      }
    }
    free(mu_new);
    free(mu_old);
    free(ef_ind);
    free(M);
  }
  out->iterations = iterations;
  if (ranked_out)
    for (int i = 0; i < nlocal; i++) ranked_out[i] = ranked[i];

  /* :406-631 dipole forces */
  double u_polar_self = 0.0, u_polar_ef = 0.0, u_polar_dd = 0.0;
  double a = p->polar_damp;
  for (int i = 0; i < nlocal; i++) {
    double qtmp = q[i], xtmp = x[3 * i], ytmp = x[3 * i + 1], ztmp = x[3 * i + 2];
    const double *mi = &mu[3 * i];
    if (eflag && alpha[i] != 0.0)
      u_polar_self += 0.5 * (mi[0] * mi[0] + mi[1] * mi[1] + mi[2] * mi[2]) / alpha[i];
    for (int j = i + 1; j < nlocal; j++) {
      const double *mj = &mu[3 * j];
      closest_image(&box, &x[3 * i], &x[3 * j], xjimage);
      double delx = xtmp - xjimage[0], dely = ytmp - xjimage[1], delz = ztmp - xjimage[2];
      double xsq = delx * delx, ysq = dely * dely, zsq = delz * delz;
      double rsq = xsq + ysq + zsq;
      double r2inv = 1.0 / rsq;
      double rinv = sqrt(r2inv);
      double r = 1.0 / rinv;
      double r3inv = r2inv * rinv;
      double forcecoulx = 0.0, forcecouly = 0.0, forcecoulz = 0.0;
      if (rsq < cut_coulsq) {
        if ((molecule[i] != molecule[j]) || molecule[i] == 0) {
          double dvdrr = 1.0 / rsq + f_shift;
          double ef_temp = dvdrr * 1.0 / r * kq;
          if (alpha[i] != 0.0 && q[j] != 0.0) {
            double common_factor = q[j] * kq * r3inv;
            forcecoulx += common_factor * (mi[0] * ((-2.0 * xsq + ysq + zsq) * r2inv + f_shift * (ysq + zsq)) +
                                           mi[1] * (-3.0 * delx * dely * r2inv - f_shift * delx * dely) +
                                           mi[2] * (-3.0 * delx * delz * r2inv - f_shift * delx * delz));
            forcecouly += common_factor * (mi[0] * (-3.0 * delx * dely * r2inv - f_shift * delx * dely) +
                                           mi[1] * ((-2.0 * ysq + xsq + zsq) * r2inv + f_shift * (xsq + zsq)) +
                                           mi[2] * (-3.0 * dely * delz * r2inv - f_shift * dely * delz));
            forcecoulz += common_factor * (mi[0] * (-3.0 * delx * delz * r2inv - f_shift * delx * delz) +
                                           mi[1] * (-3.0 * dely * delz * r2inv - f_shift * dely * delz) +
                                           mi[2] * ((-2.0 * zsq + xsq + ysq) * r2inv + f_shift * (xsq + ysq)));
            if (eflag) {
              double ef_0 = ef_temp * q[j] * delx, ef_1 = ef_temp * q[j] * dely, ef_2 = ef_temp * q[j] * delz;
              u_polar_ef -= mi[0] * ef_0 + mi[1] * ef_1 + mi[2] * ef_2;
            }
          }
          if (alpha[j] != 0.0 && qtmp != 0.0) {
            double common_factor = qtmp * kq * r3inv;
            forcecoulx -= common_factor * (mj[0] * ((-2.0 * xsq + ysq + zsq) * r2inv + f_shift * (ysq + zsq)) +
                                           mj[1] * (-3.0 * delx * dely * r2inv - f_shift * delx * dely) +
                                           mj[2] * (-3.0 * delx * delz * r2inv - f_shift * delx * delz));
            forcecouly -= common_factor * (mj[0] * (-3.0 * delx * dely * r2inv - f_shift * delx * dely) +
                                           mj[1] * ((-2.0 * ysq + xsq + zsq) * r2inv + f_shift * (xsq + zsq)) +
                                           mj[2] * (-3.0 * dely * delz * r2inv - f_shift * dely * delz));
            forcecoulz -= common_factor * (mj[0] * (-3.0 * delx * delz * r2inv - f_shift * delx * delz) +
                                           mj[1] * (-3.0 * dely * delz * r2inv - f_shift * dely * delz) +
                                           mj[2] * ((-2.0 * zsq + xsq + ysq) * r2inv + f_shift * (xsq + ysq)));
            if (eflag) {
              double ef_0 = ef_temp * qtmp * delx, ef_1 = ef_temp * qtmp * dely, ef_2 = ef_temp * qtmp * delz;
              u_polar_ef += mj[0] * ef_0 + mj[1] * ef_1 + mj[2] * ef_2;
            }
          }
        }
      }
      if (alpha[i] != 0.0 && alpha[j] != 0.0 && !(polar_cutsq > 0.0 && !(rsq < polar_cutsq))) {
        double r5inv = r3inv * r2inv;
        double r7inv = r5inv * r2inv;
        double pdotp = mi[0] * mj[0] + mi[1] * mj[1] + mi[2] * mj[2];
        double pidotr = mi[0] * delx + mi[1] * dely + mi[2] * delz;
        double pjdotr = mj[0] * delx + mj[1] * dely + mj[2] * delz;
        if (p->damping_type == POLREF_DAMP_EXPONENTIAL) {
          double term_1 = exp(-a * r);
          double term_2 = 1.0 + a * r + 0.5 * a * a * r * r;
          double term_3 = 1.0 + a * r + 0.5 * a * a * r * r + 1.0 / 6.0 * a * a * a * r * r * r;
          double pre1 = 3.0 * r5inv * pdotp * (1.0 - term_1 * term_2) -
                        15.0 * r7inv * pidotr * pjdotr * (1.0 - term_1 * term_3);
          double pre2 = 3.0 * r5inv * pjdotr * (1.0 - term_1 * term_3);
          double pre3 = 3.0 * r5inv * pidotr * (1.0 - term_1 * term_3);
          double pre4 = -pdotp * r3inv * (-term_1 * (a * rinv + a * a) + term_1 * a * term_2 * rinv);
          double pre5 = 3.0 * pidotr * pjdotr * r5inv *
                        (-term_1 * (a * rinv + a * a + 0.5 * r * a * a * a) + term_1 * a * term_3 * rinv);
          forcecoulx += pre1 * delx + pre2 * mi[0] + pre3 * mj[0] + pre4 * delx + pre5 * delx;
          forcecouly += pre1 * dely + pre2 * mi[1] + pre3 * mj[1] + pre4 * dely + pre5 * dely;
          forcecoulz += pre1 * delz + pre2 * mi[2] + pre3 * mj[2] + pre4 * delz + pre5 * delz;
          if (eflag)
            u_polar_dd += r3inv * pdotp * (1.0 - term_1 * term_2) -
                          3.0 * r5inv * pidotr * pjdotr * (1.0 - term_1 * term_3);
        } else {
          double pre1 = 3.0 * r5inv * pdotp - 15.0 * r7inv * pidotr * pjdotr;
          double pre2 = 3.0 * r5inv * pjdotr;
          double pre3 = 3.0 * r5inv * pidotr;
          forcecoulx += pre1 * delx + pre2 * mi[0] + pre3 * mj[0];
          forcecouly += pre1 * dely + pre2 * mi[1] + pre3 * mj[1];
          forcecoulz += pre1 * delz + pre2 * mi[2] + pre3 * mj[2];
          if (eflag) u_polar_dd += r3inv * pdotp - 3.0 * r5inv * pidotr * pjdotr;
        }
      }
      f[3 * i] += forcecoulx;
      f[3 * i + 1] += forcecouly;
      f[3 * i + 2] += forcecoulz;
      f[3 * j] -= forcecoulx;
      f[3 * j + 1] -= forcecouly;
      f[3 * j + 2] -= forcecoulz;
      /* ev_tally_xyz(i,j,nlocal,newton_pair,0,0,fx,fy,fz,delx,dely,delz): src/pair.cpp:1001-1089 */
      if (vflag_global) {
        virial[0] += delx * forcecoulx;
        virial[1] += dely * forcecouly;
        virial[2] += delz * forcecoulz;
        virial[3] += delx * forcecouly;
        virial[4] += delx * forcecoulz;
        virial[5] += dely * forcecoulz;
      }
      if (p->vatom) { /* src/pair.cpp:1041-1087: energies passed as 0, so eatom is untouched */
        double v[6] = {delx * forcecoulx, dely * forcecouly, delz * forcecoulz,
                       delx * forcecouly, delx * forcecoulz, dely * forcecoulz};
        for (int k = 0; k < 6; k++) {
          p->vatom[6 * i + k] += 0.5 * v[k];
          p->vatom[6 * j + k] += 0.5 * v[k];
        }
      }
    }
  }
  out->u_self = u_polar_self;
  out->u_ef = u_polar_ef;
  out->u_dd = u_polar_dd;
  /* :632,641: eng_pol = u_polar unconditionally (the partial sums are 0 when !eflag) */
  out->eng_pol = u_polar_self + u_polar_ef + u_polar_dd;
  out->eng_vdwl = eng_vdwl;
  out->eng_coul = eng_coul;

  /* :644 virial_fdotr_compute (src/pair.cpp:1495-1543): sum over local AND ghost atoms of f*x.
   * Like the reference this uses whatever is in f, so callers pass f zeroed (Verlet::force_clear). */
  if (vflag_fdotr) {
    for (int i = 0; i < ntotal; i++) {
      virial[0] += f[3 * i] * x[3 * i];
      virial[1] += f[3 * i + 1] * x[3 * i + 1];
      virial[2] += f[3 * i + 2] * x[3 * i + 2];
      virial[3] += f[3 * i + 1] * x[3 * i];
      virial[4] += f[3 * i + 2] * x[3 * i];
      virial[5] += f[3 * i + 2] * x[3 * i + 1];
    }
  }
  for (int k = 0; k < 6; k++) out->virial[k] = virial[k];
  free(rank_metric);
  free(ranked);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* row-gather form of the polarization part (large N / truncated mode / OpenMP)                */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  long *first; /* nlocal+1 */
  int *idx;    /* partners, ascending */
} partners_t;

static int cmp_int(const void *a, const void *b)
{
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}

/* minimum-image rsq in the reference's pair orientation (lower index first) */
static double pair_del(const box_t *box, const double *x, int i, int j, double del[3])
{
  int lo = i < j ? i : j, hi = i < j ? j : i;
  double img[3];
  closest_image(box, &x[3 * lo], &x[3 * hi], img);
  del[0] = x[3 * lo] - img[0];
  del[1] = x[3 * lo + 1] - img[1];
  del[2] = x[3 * lo + 2] - img[2];
  return del[0] * del[0] + del[1] * del[1] + del[2] * del[2];
}

/* partner lists for rows [row0,row1): all j != i whose minimum-image rsq <= cut^2 (cut<=0: all j) */
static int build_partners(const polref_params *p, const box_t *box, int nlocal, const double *x,
                          double cut, int row0, int row1, partners_t *pl)
{
  int nrows = row1 - row0;
  pl->first = (long *)malloc(sizeof(long) * (size_t)(nrows + 1));
  if (cut <= 0.0) {
    pl->idx = NULL; /* implicit: all j */
    for (int r = 0; r <= nrows; r++) pl->first[r] = (long)r * (nlocal - 1);
    return 0;
  }
  double cutsq = cut * cut;
  int nc[3];
  double cinv[3];
  for (int d = 0; d < 3; d++) {
    nc[d] = (int)(box->prd[d] / cut);
    if (nc[d] < 1) nc[d] = 1;
    if (!p->periodic[d]) nc[d] = nc[d] < 1 ? 1 : nc[d];
    cinv[d] = nc[d] / box->prd[d];
  }
  long ncell = (long)nc[0] * nc[1] * nc[2];
  int *cellof = (int *)malloc(sizeof(int) * (size_t)nlocal);
  long *cstart = (long *)calloc((size_t)(ncell + 1), sizeof(long));
  int *corder = (int *)malloc(sizeof(int) * (size_t)nlocal);
  for (int i = 0; i < nlocal; i++) {
    int c[3];
    for (int d = 0; d < 3; d++) {
      double s = (x[3 * i + d] - p->boxlo[d]) * cinv[d];
      int k = (int)floor(s);
      if (p->periodic[d]) {
        k %= nc[d];
        if (k < 0) k += nc[d];
      } else {
        if (k < 0) k = 0;
        if (k >= nc[d]) k = nc[d] - 1;
      }
      c[d] = k;
    }
    cellof[i] = (c[2] * nc[1] + c[1]) * nc[0] + c[0];
    cstart[cellof[i] + 1]++;
  }
  for (long c = 0; c < ncell; c++) cstart[c + 1] += cstart[c];
  long *fill = (long *)malloc(sizeof(long) * (size_t)ncell);
  memcpy(fill, cstart, sizeof(long) * (size_t)ncell);
  for (int i = 0; i < nlocal; i++) corder[fill[cellof[i]]++] = i;
  free(fill);

  /* count then fill */
  int **rowlists = (int **)malloc(sizeof(int *) * (size_t)nrows);
  int *rowcount = (int *)malloc(sizeof(int) * (size_t)nrows);
#pragma omp parallel for schedule(dynamic, 64)
  for (int r = 0; r < nrows; r++) {
    int i = row0 + r;
    int ci = cellof[i];
    int c0[3] = {ci % nc[0], (ci / nc[0]) % nc[1], ci / (nc[0] * nc[1])};
    int cap = 256, n = 0;
    int *lst = (int *)malloc(sizeof(int) * (size_t)cap);
    int lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
      if (nc[d] >= 3) {
        lo[d] = -1;
        hi[d] = 1;
      } else { /* fewer than 3 cells: visit every cell of this dimension exactly once */
        lo[d] = -c0[d];
        hi[d] = nc[d] - 1 - c0[d];
      }
    }
    for (int dz = lo[2]; dz <= hi[2]; dz++)
      for (int dy = lo[1]; dy <= hi[1]; dy++)
        for (int dx = lo[0]; dx <= hi[0]; dx++) {
          int c[3] = {c0[0] + dx, c0[1] + dy, c0[2] + dz};
          int skip = 0;
          for (int d = 0; d < 3; d++) {
            if (c[d] < 0 || c[d] >= nc[d]) {
              if (p->periodic[d]) c[d] = (c[d] + nc[d]) % nc[d];
              else skip = 1;
            }
          }
          if (skip) continue;
          long cc = ((long)c[2] * nc[1] + c[1]) * nc[0] + c[0];
          for (long k = cstart[cc]; k < cstart[cc + 1]; k++) {
            int j = corder[k];
            if (j == i) continue;
            double del[3];
            double rsq = pair_del(box, x, i, j, del);
            if (rsq <= cutsq) {
              if (n == cap) {
                cap *= 2;
                lst = (int *)realloc(lst, sizeof(int) * (size_t)cap);
              }
              lst[n++] = j;
            }
          }
        }
    qsort(lst, (size_t)n, sizeof(int), cmp_int);
    rowlists[r] = lst;
    rowcount[r] = n;
  }
  pl->first[0] = 0;
  for (int r = 0; r < nrows; r++) pl->first[r + 1] = pl->first[r] + rowcount[r];
  pl->idx = (int *)malloc(sizeof(int) * (size_t)(pl->first[nrows] > 0 ? pl->first[nrows] : 1));
  for (int r = 0; r < nrows; r++) {
    memcpy(pl->idx + pl->first[r], rowlists[r], sizeof(int) * (size_t)rowcount[r]);
    free(rowlists[r]);
  }
  free(rowlists);
  free(rowcount);
  free(cellof);
  free(cstart);
  free(corder);
  return 0;
}

static inline int partner_count(const partners_t *pl, int r) { return (int)(pl->first[r + 1] - pl->first[r]); }
static inline int partner_at(const partners_t *pl, int r, int i, int k)
{
  if (pl->idx) return pl->idx[pl->first[r] + k];
  return k < i ? k : k + 1; /* implicit all-j list, skipping i */
}

/* static-field row: the contributions the reference's i<j scatter loop (:329-361) delivers to atom i,
 * in the order it delivers them (ascending partner index) */
static void row_static(const polref_params *p, const box_t *box, const partners_t *pl, int r, int i,
                       const double *x, const double *q, const int *molecule, double f_shift,
                       double cut_coulsq, double *Ei)
{
  double e0 = 0, e1 = 0, e2 = 0;
  int n = partner_count(pl, r);
  for (int k = 0; k < n; k++) {
    int j = partner_at(pl, r, i, k);
    double del[3];
    double rsq = pair_del(box, x, i, j, del);
    if (rsq <= cut_coulsq && ((molecule[i] != molecule[j]) || molecule[i] == 0)) {
      double rr = sqrt(rsq);
      double dvdrr = 1.0 / rsq + f_shift;
      double ef_temp = dvdrr * 1.0 / rr;
      if (i < j) {
        e0 += ef_temp * q[j] * del[0];
        e1 += ef_temp * q[j] * del[1];
        e2 += ef_temp * q[j] * del[2];
      } else {
        e0 -= ef_temp * q[j] * del[0];
        e1 -= ef_temp * q[j] * del[1];
        e2 -= ef_temp * q[j] * del[2];
      }
    }
  }
  Ei[0] = e0;
  Ei[1] = e1;
  Ei[2] = e2;
}

/* induced-field row (:1161-1168) with on-the-fly T blocks; ov >= 0: partner ov contributes ovmu instead of its
 * entry of mu (the in-group Gauss-Seidel step of the group-coloured sweep) */
static void row_induced_ov(const polref_params *p, const box_t *box, const partners_t *pl, int r, int i,
                           const double *x, const double *mu, double polar_cutsq, int ov, const double *ovmu, double *Ei)
{
  double e[3] = {0, 0, 0};
  int n = partner_count(pl, r);
  for (int k = 0; k < n; k++) {
    int j = partner_at(pl, r, i, k);
    int lo = i < j ? i : j, hi = i < j ? j : i;
    double T[3][3], r2;
    t_block(p, box, &x[3 * lo], &x[3 * hi], T, &r2);
    if (polar_cutsq > 0.0 && !(r2 < polar_cutsq)) continue;
    const double *mj = (j == ov) ? ovmu : &mu[3 * j];
    for (int pp = 0; pp < 3; pp++)
      for (int qq = 0; qq < 3; qq++) e[pp] -= T[pp][qq] * mj[qq];
  }
  Ei[0] = e[0];
  Ei[1] = e[1];
  Ei[2] = e[2];
}

static void row_induced(const polref_params *p, const box_t *box, const partners_t *pl, int r, int i,
                        const double *x, const double *mu, double polar_cutsq, double *Ei)
{
  double e[3] = {0, 0, 0};
  int n = partner_count(pl, r);
  for (int k = 0; k < n; k++) {
    int j = partner_at(pl, r, i, k);
    int lo = i < j ? i : j, hi = i < j ? j : i;
    double T[3][3], r2;
    t_block(p, box, &x[3 * lo], &x[3 * hi], T, &r2);
    if (polar_cutsq > 0.0 && !(r2 < polar_cutsq)) continue;
    for (int pp = 0; pp < 3; pp++)
      for (int qq = 0; qq < 3; qq++) e[pp] -= T[pp][qq] * mu[3 * j + qq];
  }
  Ei[0] = e[0];
  Ei[1] = e[1];
  Ei[2] = e[2];
}

/* force/energy row: pair expressions of :435-631 evaluated in the (lo,hi) orientation; the force on
 * lo is added to row lo and subtracted from row hi; energies are tallied on the lo row only. */
static void row_force(const polref_params *p, const box_t *box, const partners_t *pl, int r, int i,
                      const double *x, const double *q, const int *molecule, const double *alpha,
                      const double *mu, double f_shift, double cut_coulsq, double polar_cutsq,
                      double kq, int eflag, double *Fi, double *u_ef, double *u_dd)
{
  double a = p->polar_damp;
  double fx = 0, fy = 0, fz = 0, uef = 0, udd = 0;
  int n = partner_count(pl, r);
  for (int k = 0; k < n; k++) {
    int jj = partner_at(pl, r, i, k);
    int lo = i < jj ? i : jj, hi = i < jj ? jj : i;
    const double *mi = &mu[3 * lo], *mj = &mu[3 * hi];
    double del[3];
    double rsq = pair_del(box, x, lo, hi, del);
    double delx = del[0], dely = del[1], delz = del[2];
    double xsq = delx * delx, ysq = dely * dely, zsq = delz * delz;
    rsq = xsq + ysq + zsq;
    double r2inv = 1.0 / rsq;
    double rinv = sqrt(r2inv);
    double rr = 1.0 / rinv;
    double r3inv = r2inv * rinv;
    double qlo = q[lo], qhi = q[hi];
    double cx = 0, cy = 0, cz = 0;
    if (rsq < cut_coulsq && ((molecule[lo] != molecule[hi]) || molecule[lo] == 0)) {
      double dvdrr = 1.0 / rsq + f_shift;
      double ef_temp = dvdrr * 1.0 / rr * kq;
      if (alpha[lo] != 0.0 && qhi != 0.0) {
        double cf = qhi * kq * r3inv;
        cx += cf * (mi[0] * ((-2.0 * xsq + ysq + zsq) * r2inv + f_shift * (ysq + zsq)) +
                    mi[1] * (-3.0 * delx * dely * r2inv - f_shift * delx * dely) +
                    mi[2] * (-3.0 * delx * delz * r2inv - f_shift * delx * delz));
        cy += cf * (mi[0] * (-3.0 * delx * dely * r2inv - f_shift * delx * dely) +
                    mi[1] * ((-2.0 * ysq + xsq + zsq) * r2inv + f_shift * (xsq + zsq)) +
                    mi[2] * (-3.0 * dely * delz * r2inv - f_shift * dely * delz));
        cz += cf * (mi[0] * (-3.0 * delx * delz * r2inv - f_shift * delx * delz) +
                    mi[1] * (-3.0 * dely * delz * r2inv - f_shift * dely * delz) +
                    mi[2] * ((-2.0 * zsq + xsq + ysq) * r2inv + f_shift * (xsq + ysq)));
        if (eflag && i == lo)
          uef -= mi[0] * (ef_temp * qhi * delx) + mi[1] * (ef_temp * qhi * dely) + mi[2] * (ef_temp * qhi * delz);
      }
      if (alpha[hi] != 0.0 && qlo != 0.0) {
        double cf = qlo * kq * r3inv;
        cx -= cf * (mj[0] * ((-2.0 * xsq + ysq + zsq) * r2inv + f_shift * (ysq + zsq)) +
                    mj[1] * (-3.0 * delx * dely * r2inv - f_shift * delx * dely) +
                    mj[2] * (-3.0 * delx * delz * r2inv - f_shift * delx * delz));
        cy -= cf * (mj[0] * (-3.0 * delx * dely * r2inv - f_shift * delx * dely) +
                    mj[1] * ((-2.0 * ysq + xsq + zsq) * r2inv + f_shift * (xsq + zsq)) +
                    mj[2] * (-3.0 * dely * delz * r2inv - f_shift * dely * delz));
        cz -= cf * (mj[0] * (-3.0 * delx * delz * r2inv - f_shift * delx * delz) +
                    mj[1] * (-3.0 * dely * delz * r2inv - f_shift * dely * delz) +
                    mj[2] * ((-2.0 * zsq + xsq + ysq) * r2inv + f_shift * (xsq + ysq)));
        if (eflag && i == lo)
          uef += mj[0] * (ef_temp * qlo * delx) + mj[1] * (ef_temp * qlo * dely) + mj[2] * (ef_temp * qlo * delz);
      }
    }
    if (alpha[lo] != 0.0 && alpha[hi] != 0.0 && !(polar_cutsq > 0.0 && !(rsq < polar_cutsq))) {
      double r5inv = r3inv * r2inv;
      double r7inv = r5inv * r2inv;
      double pdotp = mi[0] * mj[0] + mi[1] * mj[1] + mi[2] * mj[2];
      double pidotr = mi[0] * delx + mi[1] * dely + mi[2] * delz;
      double pjdotr = mj[0] * delx + mj[1] * dely + mj[2] * delz;
      if (p->damping_type == POLREF_DAMP_EXPONENTIAL) {
        double term_1 = exp(-a * rr);
        double term_2 = 1.0 + a * rr + 0.5 * a * a * rr * rr;
        double term_3 = 1.0 + a * rr + 0.5 * a * a * rr * rr + 1.0 / 6.0 * a * a * a * rr * rr * rr;
        double pre1 = 3.0 * r5inv * pdotp * (1.0 - term_1 * term_2) -
                      15.0 * r7inv * pidotr * pjdotr * (1.0 - term_1 * term_3);
        double pre2 = 3.0 * r5inv * pjdotr * (1.0 - term_1 * term_3);
        double pre3 = 3.0 * r5inv * pidotr * (1.0 - term_1 * term_3);
        double pre4 = -pdotp * r3inv * (-term_1 * (a * rinv + a * a) + term_1 * a * term_2 * rinv);
        double pre5 = 3.0 * pidotr * pjdotr * r5inv *
                      (-term_1 * (a * rinv + a * a + 0.5 * rr * a * a * a) + term_1 * a * term_3 * rinv);
        cx += pre1 * delx + pre2 * mi[0] + pre3 * mj[0] + pre4 * delx + pre5 * delx;
        cy += pre1 * dely + pre2 * mi[1] + pre3 * mj[1] + pre4 * dely + pre5 * dely;
        cz += pre1 * delz + pre2 * mi[2] + pre3 * mj[2] + pre4 * delz + pre5 * delz;
        if (eflag && i == lo)
          udd += r3inv * pdotp * (1.0 - term_1 * term_2) - 3.0 * r5inv * pidotr * pjdotr * (1.0 - term_1 * term_3);
      } else {
        double pre1 = 3.0 * r5inv * pdotp - 15.0 * r7inv * pidotr * pjdotr;
        double pre2 = 3.0 * r5inv * pjdotr;
        double pre3 = 3.0 * r5inv * pidotr;
        cx += pre1 * delx + pre2 * mi[0] + pre3 * mj[0];
        cy += pre1 * dely + pre2 * mi[1] + pre3 * mj[1];
        cz += pre1 * delz + pre2 * mi[2] + pre3 * mj[2];
        if (eflag && i == lo) udd += r3inv * pdotp - 3.0 * r5inv * pidotr * pjdotr;
      }
    }
    if (i == lo) {
      fx += cx;
      fy += cy;
      fz += cz;
    } else {
      fx -= cx;
      fy -= cy;
      fz -= cz;
    }
  }
  Fi[0] = fx;
  Fi[1] = fy;
  Fi[2] = fz;
  *u_ef = uef;
  *u_dd = udd;
}

static void set_threads(int nthreads)
{
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
}

int polref_polar_rows(const polref_params *p, int nlocal, const double *x, const double *q,
                      const int *molecule, const double *alpha, double *mu, double *ef_static,
                      double *f, int eflag, polref_result *out, double *trace, int trace_max,
                      int nthreads)
{
  set_threads(nthreads);
  box_t box;
  box_init(&box, p->boxlo, p->boxhi, p->periodic);
  memset(out, 0, sizeof(*out));
  double cut_coulsq = p->cut_coul * p->cut_coul;
  double polar_cutsq = (p->polar_cut > 0.0) ? p->polar_cut * p->polar_cut : -1.0;
  double f_shift = -1.0 / (p->cut_coul * p->cut_coul);
  double kq = sqrt(p->qqrd2e);
  double list_cut = (p->polar_cut > 0.0) ? (p->polar_cut > p->cut_coul ? p->polar_cut : p->cut_coul) : -1.0;
  partners_t pl;
  build_partners(p, &box, nlocal, x, list_cut, 0, nlocal, &pl);

  /* rank metric (:192-227) over minimum-image partners instead of raw ghost coordinates:
   * identical pair set whenever 1.5*rmin is below half the box and the ghost cutoff. */
  double *rank_metric = (double *)calloc((size_t)nlocal, sizeof(double));
  int *ranked = (int *)malloc(sizeof(int) * (size_t)nlocal);
  double rmin = 1000.0;
  if (p->polar_gs_ranked) {
    for (int i = 0; i < nlocal; i++) {
      int n = partner_count(&pl, i);
      for (int k = 0; k < n; k++) {
        int j = partner_at(&pl, i, i, k);
        double del[3];
        double r = sqrt(pair_del(&box, x, i, j, del));
        if (alpha[i] > 0 && alpha[j] > 0 && rmin > r && ((molecule[i] != molecule[j]) || molecule[i] == 0))
          rmin = r;
      }
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < nlocal; i++) {
      int n = partner_count(&pl, i);
      double m = 0;
      for (int k = 0; k < n; k++) {
        int j = partner_at(&pl, i, i, k);
        double del[3];
        double r = sqrt(pair_del(&box, x, i, j, del));
        if (rmin * 1.5 > r && ((molecule[i] != molecule[j]) || molecule[i] == 0)) m += alpha[i] * alpha[j];
      }
      rank_metric[i] = m;
    }
    stable_rank(nlocal, rank_metric, ranked);
  } else
    for (int i = 0; i < nlocal; i++) ranked[i] = i;
  out->rmin = rmin;

#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < nlocal; i++) {
    double E[3];
    row_static(p, &box, &pl, i, i, x, q, molecule, f_shift, cut_coulsq, E);
    for (int c = 0; c < 3; c++) {
      ef_static[3 * i + c] = E[c] * kq;
      if (!p->use_previous) {
        mu[3 * i + c] = alpha[i] * ef_static[3 * i + c];
        mu[3 * i + c] *= p->polar_gamma;
      }
    }
  }

  int iterations = 0;
  if (!p->zodid) {
    double *mu_new = (double *)malloc(sizeof(double) * 3 * (size_t)nlocal);
    double *mu_old = (double *)malloc(sizeof(double) * 3 * (size_t)nlocal);
    int gs = p->polar_gs || p->polar_gs_ranked;
    int nchunks = (gs && p->gs_chunks != 0) ? abs(p->gs_chunks) : 0;
    if (gs && p->gs_chunks < 0) {
      /* EXTENSION: interleaved colouring -- chunk c holds the ranked positions c, c+C, c+2C, ... so that atoms
       * that are neighbours in the ranked order (typically the sites of one molecule) fall into DIFFERENT chunks
       * and see each other's new dipoles within the sweep.  Realised by re-ordering the visiting order. */
      int *tmp = (int *)malloc(sizeof(int) * (size_t)nlocal);
      int k = 0;
      for (int c = 0; c < nchunks; c++)
        for (int pos = c; pos < nlocal; pos += nchunks) tmp[k++] = ranked[pos];
      memcpy(ranked, tmp, sizeof(int) * (size_t)nlocal);
      free(tmp);
    }
    int keep_iterating = 1;
    while (keep_iterating) {
      memcpy(mu_old, mu, sizeof(double) * 3 * (size_t)nlocal);
      int nblk = !gs ? 1 : (nchunks ? nchunks : nlocal);
      if (gs && p->gs_colour) {
        /* EXTENSION: explicit colouring (the group-coloured sweep of the CUDA list path).  Colours in turn; inside a
         * colour every atom is updated from the dipoles as they stand before the colour (Jacobi), except that an atom
         * with gs_after[i] = a >= 0 (the second member of a pair group) sees the NEW dipole of a. */
        for (int cb = 0; cb < p->gs_ncolours; cb++) {
#pragma omp parallel for schedule(dynamic, 64)
          for (int i = 0; i < nlocal; i++) {
            if (p->gs_colour[i] != cb || p->gs_after[i] >= 0) continue;
            double E[3];
            row_induced(p, &box, &pl, i, i, x, mu, polar_cutsq, E);
            for (int c = 0; c < 3; c++) mu_new[3 * i + c] = alpha[i] * (ef_static[3 * i + c] + E[c]);
          }
#pragma omp parallel for schedule(dynamic, 64)
          for (int i = 0; i < nlocal; i++) {
            if (p->gs_colour[i] != cb || p->gs_after[i] < 0) continue;
            int a = p->gs_after[i];
            double E[3];
            row_induced_ov(p, &box, &pl, i, i, x, mu, polar_cutsq, a, &mu_new[3 * a], E);
            for (int c = 0; c < 3; c++) mu_new[3 * i + c] = alpha[i] * (ef_static[3 * i + c] + E[c]);
          }
          for (int i = 0; i < nlocal; i++)
            if (p->gs_colour[i] == cb)
              for (int c = 0; c < 3; c++) mu[3 * i + c] = mu_new[3 * i + c];
        }
      } else if (nblk == nlocal) {
        for (int pos = 0; pos < nlocal; pos++) {
          int i = ranked[pos];
          double E[3];
          row_induced(p, &box, &pl, i, i, x, mu, polar_cutsq, E);
          for (int c = 0; c < 3; c++) {
            mu_new[3 * i + c] = alpha[i] * (ef_static[3 * i + c] + E[c]);
            mu[3 * i + c] = mu_new[3 * i + c];
          }
        }
      } else {
        for (int cb = 0; cb < nblk; cb++) {
          int beg = (int)(((long)cb * nlocal) / nblk), end = (int)(((long)(cb + 1) * nlocal) / nblk);
#pragma omp parallel for schedule(dynamic, 64)
          for (int pos = beg; pos < end; pos++) {
            int i = ranked[pos];
            double E[3];
            row_induced(p, &box, &pl, i, i, x, mu, polar_cutsq, E);
            for (int c = 0; c < 3; c++) mu_new[3 * i + c] = alpha[i] * (ef_static[3 * i + c] + E[c]);
          }
          if (gs)
            for (int pos = beg; pos < end; pos++)
              for (int c = 0; c < 3; c++) mu[3 * ranked[pos] + c] = mu_new[3 * ranked[pos] + c];
        }
      }
      if (trace && iterations < trace_max) {
        double *dst = trace + (size_t)iterations * 3 * nlocal;
        for (int i = 0; i < 3 * nlocal; i++) dst[i] = gs ? mu[i] : mu_new[i];
      }
      if (p->fixed_iteration == 0) {
        keep_iterating = 0;
        double change = 0;
        for (int i = 0; i < 3 * nlocal; i++) change += (mu_new[i] - mu_old[i]) * (mu_new[i] - mu_old[i]);
        change /= (double)(nlocal)*3.0;
        if (change > p->polar_precision * p->polar_precision) keep_iterating = 1;
      } else {
        if (iterations >= p->iterations_max) break;
      }
      memcpy(mu, mu_new, sizeof(double) * 3 * (size_t)nlocal);
      iterations++;
      if (iterations > p->iterations_max) {
        for (int i = 0; i < nlocal; i++)
          for (int c = 0; c < 3; c++) mu[3 * i + c] = alpha[i] * ef_static[3 * i + c];
        out->diverged = 1;
        break;
      }
    }
    free(mu_new);
    free(mu_old);
  }
  out->iterations = iterations;

  double *uef = (double *)calloc((size_t)nlocal, sizeof(double));
  double *udd = (double *)calloc((size_t)nlocal, sizeof(double));
  double *frow = (double *)malloc(sizeof(double) * 3 * (size_t)nlocal);
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < nlocal; i++)
    row_force(p, &box, &pl, i, i, x, q, molecule, alpha, mu, f_shift, cut_coulsq, polar_cutsq, kq, eflag,
              &frow[3 * i], &uef[i], &udd[i]);
  double u_self = 0, u_ef = 0, u_dd = 0;
  for (int i = 0; i < nlocal; i++) {
    f[3 * i] += frow[3 * i];
    f[3 * i + 1] += frow[3 * i + 1];
    f[3 * i + 2] += frow[3 * i + 2];
    if (eflag && alpha[i] != 0.0)
      u_self += 0.5 * (mu[3 * i] * mu[3 * i] + mu[3 * i + 1] * mu[3 * i + 1] + mu[3 * i + 2] * mu[3 * i + 2]) / alpha[i];
    u_ef += uef[i];
    u_dd += udd[i];
  }
  out->u_self = u_self;
  out->u_ef = u_ef;
  out->u_dd = u_dd;
  out->eng_pol = u_self + u_ef + u_dd;
  /* polarization part of the reference's F.r virial: forces act on local atoms only (H7) */
  for (int i = 0; i < nlocal; i++) {
    out->virial[0] += frow[3 * i] * x[3 * i];
    out->virial[1] += frow[3 * i + 1] * x[3 * i + 1];
    out->virial[2] += frow[3 * i + 2] * x[3 * i + 2];
    out->virial[3] += frow[3 * i + 1] * x[3 * i];
    out->virial[4] += frow[3 * i + 2] * x[3 * i];
    out->virial[5] += frow[3 * i + 2] * x[3 * i + 1];
  }
  free(uef);
  free(udd);
  free(frow);
  free(rank_metric);
  free(ranked);
  free(pl.first);
  free(pl.idx);
  return 0;
}

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* CPU-baseline kernel for bench.py: the polarization step (static field, nsweeps Jacobi sweeps,
 * force/energy pass) restricted to rows [row0,row1) of an nlocal-atom system.  Returns the seconds
 * spent in those three stages (partner-list construction is excluded and reported by the caller). */
double polref_bench_rows(const polref_params *p, int nlocal, const double *x, const double *q,
                         const int *molecule, const double *alpha, int row0, int row1, int nsweeps,
                         int nthreads, double *checksum)
{
  set_threads(nthreads);
  box_t box;
  box_init(&box, p->boxlo, p->boxhi, p->periodic);
  double cut_coulsq = p->cut_coul * p->cut_coul;
  double polar_cutsq = (p->polar_cut > 0.0) ? p->polar_cut * p->polar_cut : -1.0;
  double f_shift = -1.0 / (p->cut_coul * p->cut_coul);
  double kq = sqrt(p->qqrd2e);
  double list_cut = (p->polar_cut > 0.0) ? (p->polar_cut > p->cut_coul ? p->polar_cut : p->cut_coul) : -1.0;
  partners_t pl;
  build_partners(p, &box, nlocal, x, list_cut, row0, row1, &pl);
  int nrows = row1 - row0;
  double *ef = (double *)calloc(3 * (size_t)nlocal, sizeof(double));
  double *mu = (double *)calloc(3 * (size_t)nlocal, sizeof(double));
  double *mu_new = (double *)calloc(3 * (size_t)nlocal, sizeof(double));
  /* every atom needs a dipole for the gathers: seed non-sample atoms with a plausible value */
  for (int i = 0; i < nlocal; i++)
    for (int c = 0; c < 3; c++) mu[3 * i + c] = 1e-3 * alpha[i] * ((i * 3 + c) % 7 - 3);
  double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 16)
  for (int r = 0; r < nrows; r++) {
    int i = row0 + r;
    double E[3];
    row_static(p, &box, &pl, r, i, x, q, molecule, f_shift, cut_coulsq, E);
    for (int c = 0; c < 3; c++) ef[3 * i + c] = E[c] * kq;
  }
  for (int s = 0; s < nsweeps; s++) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int r = 0; r < nrows; r++) {
      int i = row0 + r;
      double E[3];
      row_induced(p, &box, &pl, r, i, x, mu, polar_cutsq, E);
      for (int c = 0; c < 3; c++) mu_new[3 * i + c] = alpha[i] * (ef[3 * i + c] + E[c]);
    }
    for (int r = 0; r < nrows; r++)
      for (int c = 0; c < 3; c++) mu[3 * (row0 + r) + c] = mu_new[3 * (row0 + r) + c];
  }
  double cs = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : cs)
  for (int r = 0; r < nrows; r++) {
    int i = row0 + r;
    double F[3], a, b;
    row_force(p, &box, &pl, r, i, x, q, molecule, alpha, mu, f_shift, cut_coulsq, polar_cutsq, kq, 1, F, &a, &b);
    cs += F[0] + F[1] + F[2] + a + b;
  }
  double t1 = now_s();
  if (checksum) *checksum = cs;
  free(ef);
  free(mu);
  free(mu_new);
  free(pl.first);
  free(pl.idx);
  return t1 - t0;
}


/* ===================================================================================================
 * KSpace: reciprocal-space Ewald
 * =================================================================================================== */

static double ewald_rms(int km, double prd, long natoms, double q2, double g)
{ /* ewald.cpp:343-351 */
  if (natoms == 0) natoms = 1;
  return 2.0 * q2 * g / prd * sqrt(1.0 / (M_PI * km * natoms)) * exp(-M_PI * M_PI * km * km / (g * g * prd * prd));
}

/* the half-space k set of Ewald::coeffs: the first non-zero component (x, then y, then z) is positive; axis
 * vectors run to kmax, the others to the per-dimension maxima; all inside |k|^2 <= gsqmx */
static int ewald_in_set(const polref_ewald_plan *p, const double unitk[3], int kx, int ky, int kz, double *sqk_out)
{
  if (kx < 0 || (kx == 0 && ky < 0) || (kx == 0 && ky == 0 && kz <= 0)) return 0;
  const int nz = (kx != 0) + (ky != 0) + (kz != 0);
  if (nz == 1) {
    if (abs(kx) > p->kmax || abs(ky) > p->kmax || abs(kz) > p->kmax) return 0;
  } else if (abs(kx) > p->kxmax || abs(ky) > p->kymax || abs(kz) > p->kzmax) return 0;
  const double sqk = (kx * unitk[0]) * (kx * unitk[0]) + (ky * unitk[1]) * (ky * unitk[1]) + (kz * unitk[2]) * (kz * unitk[2]);
  if (sqk > p->gsqmx) return 0;
  *sqk_out = sqk;
  return 1;
}

int polref_ewald_plan_make(double accuracy_relative, double qqrd2e, double two_charge_force, double qsqsum,
                           long natoms, double cutoff, const double prd[3], double g_ewald_in,
                           polref_ewald_plan *plan)
{
  const double accuracy = accuracy_relative * two_charge_force; /* ewald.cpp:131-132 */
  const double q2 = qsqsum * qqrd2e;                            /* kspace.cpp:293 */
  double g = g_ewald_in;
  if (!(g > 0.0)) { /* ewald.cpp:153-160 */
    if (accuracy <= 0.0 || q2 == 0.0) return 1;
    g = accuracy * sqrt(natoms * cutoff * prd[0] * prd[1] * prd[2]) / (2.0 * q2);
    if (g >= 1.0) g = (1.35 - 0.15 * log(accuracy)) / cutoff;
    else g = sqrt(-log(g)) / cutoff;
  }
  plan->g_ewald = g;
  int km[3];
  for (int d = 0; d < 3; d++) { /* ewald.cpp:241-258 */
    km[d] = 1;
    while (ewald_rms(km[d], prd[d], natoms, q2, g) > accuracy) km[d]++;
  }
  plan->kxmax = km[0]; plan->kymax = km[1]; plan->kzmax = km[2];
  plan->kmax = km[0] > km[1] ? km[0] : km[1];
  if (km[2] > plan->kmax) plan->kmax = km[2];
  double gs = 0.0;
  for (int d = 0; d < 3; d++) {
    const double u = 2.0 * M_PI / prd[d], v = u * u * km[d] * km[d];
    if (v > gs) gs = v;
  }
  plan->gsqmx = gs * 1.00001; /* ewald.cpp:311 */
  const double unitk[3] = {2.0 * M_PI / prd[0], 2.0 * M_PI / prd[1], 2.0 * M_PI / prd[2]};
  int cnt = 0;
  double sqk;
  for (int kx = 0; kx <= plan->kmax; kx++)
    for (int ky = -plan->kmax; ky <= plan->kmax; ky++)
      for (int kz = -plan->kmax; kz <= plan->kmax; kz++) cnt += ewald_in_set(plan, unitk, kx, ky, kz, &sqk);
  plan->kcount = cnt;
  return 0;
}

int polref_ewald_compute(const polref_ewald_plan *plan, int n, const double *x, const double *q,
                         const double prd[3], double qqrd2e, double *f, double *energy, double virial[6])
{
  const double g = plan->g_ewald, volume = prd[0] * prd[1] * prd[2];
  const double unitk[3] = {2.0 * M_PI / prd[0], 2.0 * M_PI / prd[1], 2.0 * M_PI / prd[2]};
  const double ginv2 = 1.0 / (g * g), preu = 4.0 * M_PI / volume;
  double qsum = 0.0, qsqsum = 0.0;
  for (int i = 0; i < n; i++) { qsum += q[i]; qsqsum += q[i] * q[i]; }
  double e = 0.0, v[6] = {0, 0, 0, 0, 0, 0};
  double *fk = (double *)calloc((size_t)3 * (n > 0 ? n : 1), sizeof(double));
  const int K = plan->kmax;
#pragma omp parallel for collapse(2) schedule(dynamic) reduction(+ : e)
  for (int kx = 0; kx <= K; kx++)
    for (int ky = -K; ky <= K; ky++)
      for (int kz = -K; kz <= K; kz++) {
        double sqk;
        if (!ewald_in_set(plan, unitk, kx, ky, kz, &sqk)) continue;
        const double kv[3] = {kx * unitk[0], ky * unitk[1], kz * unitk[2]};
        double sre = 0.0, sim = 0.0; /* structure factor, ewald.cpp:501-680 */
        for (int i = 0; i < n; i++) {
          const double ph = kv[0] * x[3 * i] + kv[1] * x[3 * i + 1] + kv[2] * x[3 * i + 2];
          sre += q[i] * cos(ph);
          sim += q[i] * sin(ph);
        }
        const double ug = preu * exp(-0.25 * sqk * ginv2) / sqk; /* ewald.cpp:776 */
        const double uk = ug * (sre * sre + sim * sim);
        e += uk;                                                 /* ewald.cpp:455-457 */
        const double vterm = -2.0 * (1.0 / sqk + 0.25 * ginv2);  /* ewald.cpp:781-788 */
        const double vk[6] = {1.0 + vterm * kv[0] * kv[0], 1.0 + vterm * kv[1] * kv[1], 1.0 + vterm * kv[2] * kv[2],
                              vterm * kv[0] * kv[1], vterm * kv[0] * kv[2], vterm * kv[1] * kv[2]};
#pragma omp critical
        {
          for (int a = 0; a < 6; a++) v[a] += uk * vk[a];
          for (int i = 0; i < n; i++) { /* field, ewald.cpp:417-431: partial = Im(e^{ikr} conj(S)) */
            const double ph = kv[0] * x[3 * i] + kv[1] * x[3 * i + 1] + kv[2] * x[3 * i + 2];
            const double partial = sin(ph) * sre - cos(ph) * sim;
            fk[3 * i] += partial * 2.0 * ug * kv[0];
            fk[3 * i + 1] += partial * 2.0 * ug * kv[1];
            fk[3 * i + 2] += partial * 2.0 * ug * kv[2];
          }
        }
      }
  for (int i = 0; i < n; i++) /* ewald.cpp:443-449 */
    for (int d = 0; d < 3; d++) f[3 * i + d] += qqrd2e * q[i] * fk[3 * i + d];
  free(fk);
  /* ewald.cpp:459-461: self energy and neutralising background */
  e -= g * qsqsum / sqrt(M_PI) + 0.5 * M_PI * qsum * qsum / (g * g * volume);
  *energy = e * qqrd2e;
  for (int a = 0; a < 6; a++) virial[a] = v[a] * qqrd2e;
  return 0;
}
