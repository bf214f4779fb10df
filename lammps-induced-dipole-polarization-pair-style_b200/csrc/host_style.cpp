// host_style.cpp -- see host_style.h.  Written against the behaviour of the reference, not its text:
// every block cites the lines whose observable behaviour (values, error strings, ordering rules) it keeps.
#include "host_style.h"

#include <cctype>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace polb200 {

namespace {

[[noreturn]] void fail(const char *msg, int code = POLB200_ERR_ARG) { throw StyleError{code, msg}; }

// Force::numeric / Force::inumeric (src/force.cpp:910-960): character-class validation, then atof/atoi
double numeric(const char *s)
{
  static const char *err = "Expected floating point parameter in input script or data file";
  if (!s || !*s) fail(err);
  for (const char *c = s; *c; ++c)
    if (!(isdigit((unsigned char)*c) || *c == '-' || *c == '+' || *c == '.' || *c == 'e' || *c == 'E')) fail(err);
  return atof(s);
}

int inumeric(const char *s)
{
  static const char *err = "Expected integer parameter in input script or data file";
  if (!s || !*s) fail(err);
  for (const char *c = s; *c; ++c)
    if (!(isdigit((unsigned char)*c) || *c == '-' || *c == '+')) fail(err);
  return atoi(s);
}

// Force::bounds (src/force.cpp:854-877): i, *, i*, *j, i*j with nmin = 1
void bounds(const char *s, int nmax, int &lo, int &hi)
{
  const char *star = strchr(s, '*');
  if (!star) lo = hi = atoi(s);
  else if (strlen(s) == 1) { lo = 1; hi = nmax; }
  else if (star == s) { lo = 1; hi = atoi(star + 1); }
  else if (star[1] == '\0') { lo = atoi(s); hi = nmax; }
  else { lo = atoi(s); hi = atoi(star + 1); }
  if (lo < 1 || hi > nmax || lo > hi) fail("Numeric index is out of bounds");
}

int yesno(const char *v, const char *errmsg)
{
  if (strcmp(v, "yes") == 0) return 1;
  if (strcmp(v, "no") == 0) return 0;
  fail(errmsg);
}

union IntFloat { int i; float f; };

}  // namespace

HostStyle::HostStyle() : tabinner(sqrt(2.0)) {}

void HostStyle::settings(int narg, const char *const *arg)
{
  static const char *illegal = "Illegal pair_style command";
  if (narg < 1) fail(illegal);
  cut_lj_global = numeric(arg[0]);
  cut_coul = (narg == 1) ? cut_lj_global : numeric(arg[1]);

  for (int k = 2; k < narg; k += 2) {
    if (k + 2 > narg) fail(illegal);  // every keyword takes exactly one value (pol.cpp:691)
    const char *key = arg[k], *val = arg[k + 1];
    if (!strcmp(key, "precision")) polar_precision = numeric(val);
    else if (!strcmp(key, "zodid")) {
      // checked BEFORE the value is parsed (pol.cpp:698): with defaults "zodid" is only legal after
      // "polar_gs_ranked no"
      if (polar_gs || polar_gs_ranked) fail("Zodid doesn't work with polar_gs or polar_gs_ranked");
      zodid = yesno(val, illegal);
    } else if (!strcmp(key, "fixed_iteration")) fixed_iteration = yesno(val, illegal);
    else if (!strcmp(key, "damp")) polar_damp = numeric(val);
    else if (!strcmp(key, "max_iterations")) iterations_max = inumeric(val);
    else if (!strcmp(key, "damp_type")) {
      if (!strcmp(val, "exponential")) damping_type = DAMP_EXPONENTIAL;
      else if (!strcmp(val, "none")) damping_type = DAMP_NONE;
      else fail(illegal);
    } else if (!strcmp(key, "polar_gs")) {
      if (polar_gs_ranked) fail("polar_gs and polar_gs_ranked are mutually exclusive");
      polar_gs = yesno(val, illegal);
    } else if (!strcmp(key, "polar_gs_ranked")) {
      if (polar_gs) fail("polar_gs and polar_gs_ranked are mutually exclusive");
      polar_gs_ranked = yesno(val, illegal);
    } else if (!strcmp(key, "polar_gamma")) polar_gamma = numeric(val);
    else if (!strcmp(key, "debug")) debug = yesno(val, illegal);
    else if (!strcmp(key, "use_previous")) use_previous = yesno(val, illegal);
    // ---- extensions (not in the reference) ----
    else if (!strcmp(key, "polar_cutoff")) {
      if (!strcmp(val, "none")) polar_cutoff = 0.0;
      else {
        polar_cutoff = numeric(val);
        if (polar_cutoff <= 0.0) fail(illegal);
      }
    } else if (!strcmp(key, "gs_chunks")) {
      gs_chunks = inumeric(val);  // n > 0: contiguous chunks of the ranked order; n < 0: |n| interleaved chunks; 0: default
    } else if (!strcmp(key, "restart_keywords")) restart_keywords = yesno(val, illegal);
    else fail(illegal);
  }

  // "reset cutoffs that have been explicitly set" (pol.cpp:760-765)
  if (allocated)
    for (int i = 1; i <= ntypes; i++)
      for (int j = i; j <= ntypes; j++)
        if (setflag[idx(i, j)]) cut_lj[idx(i, j)] = cut_lj_global;
  initialized = false;
}

void HostStyle::set_ntypes(int n)
{
  if (n < 1) fail("ntypes must be >= 1");
  if (allocated && n == ntypes) return;
  ntypes = n;
  size_t sz = (size_t)(n + 1) * (n + 1);
  setflag.assign(sz, 0);
  for (auto *v : {&epsilon, &sigma, &cut_lj, &cut_ljsq, &cutsq, &lj1, &lj2, &lj3, &lj4, &offset, &cut_pair,
                  &cutneighsq})
    v->assign(sz, 0.0);
  allocated = true;
  initialized = false;
}

void HostStyle::coeff(int narg, const char *const *arg)
{
  static const char *bad = "Incorrect args for pair coefficients";
  if (narg < 4 || narg > 5) fail(bad);
  if (!allocated) fail("polb200_set_ntypes must be called before polb200_coeff", POLB200_ERR_STATE);
  int ilo, ihi, jlo, jhi;
  bounds(arg[0], ntypes, ilo, ihi);
  bounds(arg[1], ntypes, jlo, jhi);
  double eps = numeric(arg[2]);
  double sig = numeric(arg[3]);
  double cut = (narg == 5) ? numeric(arg[4]) : cut_lj_global;
  int count = 0;
  for (int i = ilo; i <= ihi; i++)
    for (int j = (jlo > i ? jlo : i); j <= jhi; j++) {
      epsilon[idx(i, j)] = eps;
      sigma[idx(i, j)] = sig;
      cut_lj[idx(i, j)] = cut;
      setflag[idx(i, j)] = 1;
      count++;
    }
  if (count == 0) fail(bad);
  initialized = false;
}

void HostStyle::pair_modify(int narg, const char *const *arg)
{
  static const char *illegal = "Illegal pair_modify command";
  if (narg == 0) fail(illegal);
  for (int k = 0; k < narg; k += 2) {
    if (k + 2 > narg) fail(illegal);
    const char *key = arg[k], *val = arg[k + 1];
    if (!strcmp(key, "mix")) {
      if (!strcmp(val, "geometric")) mix_flag = MIX_GEOMETRIC;
      else if (!strcmp(val, "arithmetic")) mix_flag = MIX_ARITHMETIC;
      else if (!strcmp(val, "sixthpower")) mix_flag = MIX_SIXTHPOWER;
      else fail(illegal);
    } else if (!strcmp(key, "shift")) offset_flag = yesno(val, illegal);
    else if (!strcmp(key, "table")) {
      ncoultablebits = inumeric(val);
      if (ncoultablebits > (int)(sizeof(float) * CHAR_BIT)) fail("Too many total bits for bitmapped lookup table");
    } else if (!strcmp(key, "tabinner")) tabinner = numeric(val);
    else if (!strcmp(key, "tail")) {
      tail_flag = yesno(val, illegal);  // consumed by tail_correction(); the pair loops do not depend on it
    } else fail(illegal);
  }
  initialized = false;
}

static double mix_energy(int mix, double e1, double e2, double s1, double s2)
{
  if (mix == MIX_SIXTHPOWER)
    return 2.0 * sqrt(e1 * e2) * pow(s1, 3.0) * pow(s2, 3.0) / (pow(s1, 6.0) + pow(s2, 6.0));
  return sqrt(e1 * e2);
}

static double mix_distance(int mix, double s1, double s2)
{
  if (mix == MIX_GEOMETRIC) return sqrt(s1 * s2);
  if (mix == MIX_ARITHMETIC) return 0.5 * (s1 + s2);
  return pow(0.5 * (pow(s1, 6.0) + pow(s2, 6.0)), 1.0 / 6.0);
}

void HostStyle::init(const polb200_env &e)
{
  env = e;
  // init_style(), pol.cpp:806-852
  if (!env.q_flag) fail("Pair style lj/cut/coul/long requires atom attribute q");
  if (!env.polarizability_flag)
    fail("Pair style lj/cut/coul/long/polarization requires atom attribute polarizability");
  if (!env.kspace_present) fail("Pair style requires a KSpace style");
  // newton_pair off needs no special path: every owned atom computes all of its pairs itself (full list), which is
  // what the reference's newton-off half list adds up to; LAMMPS then asks for the pairwise virial (vflag = 1)
  // Pair::init(), src/pair.cpp:189-255
  if (offset_flag && tail_flag) fail("Cannot have both pair_modify shift and tail set to yes");
  if (!allocated) fail("All pair coeffs are not set");
  for (int i = 1; i <= ntypes; i++)
    if (!setflag[idx(i, i)]) fail("All pair coeffs are not set");

  cut_coulsq = cut_coul * cut_coul;
  cutforce = 0.0;
  for (int i = 1; i <= ntypes; i++)
    for (int j = i; j <= ntypes; j++) {
      int ij = idx(i, j), ji = idx(j, i), ii = idx(i, i), jj = idx(j, j);
      if (!setflag[ij]) {  // init_one(), pol.cpp:860-865
        epsilon[ij] = mix_energy(mix_flag, epsilon[ii], epsilon[jj], sigma[ii], sigma[jj]);
        sigma[ij] = mix_distance(mix_flag, sigma[ii], sigma[jj]);
        cut_lj[ij] = mix_distance(mix_flag, cut_lj[ii], cut_lj[jj]);
      }
      double cut = cut_lj[ij] > cut_coul ? cut_lj[ij] : cut_coul;  // qdist = 0 (pol.cpp:61,869)
      cut_ljsq[ij] = cut_lj[ij] * cut_lj[ij];
      double s6 = pow(sigma[ij], 6.0), s12 = pow(sigma[ij], 12.0);
      lj1[ij] = 48.0 * epsilon[ij] * s12;
      lj2[ij] = 24.0 * epsilon[ij] * s6;
      lj3[ij] = 4.0 * epsilon[ij] * s12;
      lj4[ij] = 4.0 * epsilon[ij] * s6;
      if (offset_flag && cut_lj[ij] > 0.0) {
        double ratio = sigma[ij] / cut_lj[ij];
        offset[ij] = 4.0 * epsilon[ij] * (pow(ratio, 12.0) - pow(ratio, 6.0));
      } else offset[ij] = 0.0;
      cut_ljsq[ji] = cut_ljsq[ij];
      lj1[ji] = lj1[ij]; lj2[ji] = lj2[ij]; lj3[ji] = lj3[ij]; lj4[ji] = lj4[ij];
      offset[ji] = offset[ij];
      cut_pair[ij] = cut_pair[ji] = cut;
      cutsq[ij] = cutsq[ji] = cut * cut;  // src/pair.cpp:243-246
      if (cut > cutforce) cutforce = cut;
    }

  // neighbor cutoffs, src/neighbor.cpp:293-320
  cutneighmax = 0.0;
  for (int i = 1; i <= ntypes; i++)
    for (int j = 1; j <= ntypes; j++) {
      double cutoff = sqrt(cutsq[idx(i, j)]);
      double cut = cutoff + (cutoff > 0.0 ? env.skin : 0.0);
      cutneighsq[idx(i, j)] = cut * cut;
      if (cut > cutneighmax) cutneighmax = cut;
    }
  // special_flag, src/neighbor.cpp:361-382: a KSpace style forces 2 for all three levels
  special_flag[0] = 0;
  special_flag[1] = special_flag[2] = special_flag[3] = 2;

  if (ncoultablebits) init_tables();
  else tab = CoulTables();
  initialized = true;
}

// Pair::init_bitmap + Pair::init_tables (src/pair.cpp:1676-1725, 313-520) for cut_respa == NULL and a
// non-MSM KSpace.  Must reproduce the reference tables bit for bit: the 1e-10 force parity target
// depends on it and tests/test_host_style.py compares against the reference's own arrays.
void HostStyle::init_tables()
{
  const double MY_ISPI4 = 1.12837916709551257389;  // src/math_const.h:29
  const double inner = tabinner, outer = cut_coul;
  const int nbits = ncoultablebits;

  // --- bitmap parameters ---
  if (nbits > (int)(sizeof(float) * CHAR_BIT)) fail("Too many total bits for bitmapped lookup table");
  int nlowermin = 1;
  for (;;) {
    double lo = pow(2.0, (double)nlowermin), hi = pow(2.0, (double)nlowermin + 1.0);
    if (lo <= inner * inner && hi > inner * inner) break;
    if (lo <= inner * inner) nlowermin++;
    else nlowermin--;
  }
  int nexpbits = 0;
  double required = outer * outer / pow(2.0, (double)nlowermin);
  for (double available = 2.0; available < required;) {
    nexpbits++;
    available = pow(2.0, pow(2.0, (double)nexpbits));
  }
  int nmantbits = nbits - nexpbits;
  if (nexpbits > (int)(sizeof(float) * CHAR_BIT) - FLT_MANT_DIG) fail("Too many exponent bits for lookup table");
  if (nmantbits + 1 > FLT_MANT_DIG) fail("Too many mantissa bits for lookup table");
  if (nmantbits < 3) fail("Too few bits for lookup table");
  const int shift = FLT_MANT_DIG - (nmantbits + 1);
  int mask = 1;
  for (int j = 0; j < nbits + shift; j++) mask *= 2;
  mask -= 1;
  IntFloat u;
  u.f = (float)(outer * outer);
  const int maskhi = u.i & ~mask;
  u.f = (float)(inner * inner);
  const int masklo = u.i & ~mask;

  // --- table values at the lower edge of every bin ---
  const int n = 1 << nbits;
  tab.nbits = nbits;
  tab.mask = mask;
  tab.shift = shift;
  for (auto *v : {&tab.r, &tab.dr, &tab.f, &tab.df, &tab.c, &tab.dc, &tab.e, &tab.de}) v->assign(n, 0.0);
  const double qq = env.qqrd2e, g = env.g_ewald;
  auto edge = [&](float rsqf, double &fv, double &cv, double &ev) {
    double r = sqrtf(rsqf);
    double grij = g * r;
    double expm2 = exp(-grij * grij);
    double derfc = erfc(grij);
    cv = qq / r;
    fv = qq / r * (derfc + MY_ISPI4 * grij * expm2);
    ev = qq / r * derfc;
  };
  const double innersq = inner * inner;
  IntFloat minrsq;
  minrsq.i = maskhi;
  for (int i = 0; i < n; i++) {
    IntFloat rl;
    rl.i = (i << shift) | masklo;
    if (rl.f < innersq) rl.i = (i << shift) | maskhi;
    tab.r[i] = rl.f;
    edge(rl.f, tab.f[i], tab.c[i], tab.e[i]);
    if (rl.f < minrsq.f) minrsq.f = rl.f;
  }
  tab.tabinnersq = minrsq.f;

  // --- deltas; the table is periodic in the index, then patched at the bin holding cut_coul^2 ---
  for (int i = 0; i < n; i++) {
    int nx = (i + 1) % n;
    tab.dr[i] = 1.0 / (tab.r[nx] - tab.r[i]);
    tab.df[i] = tab.f[nx] - tab.f[i];
    tab.dc[i] = tab.c[nx] - tab.c[i];
    tab.de[i] = tab.e[nx] - tab.e[i];
  }
  int itablemin = (minrsq.i & mask) >> shift;
  int itablemax = (itablemin == 0) ? n - 1 : itablemin - 1;
  IntFloat top;
  top.i = (itablemax << shift) | maskhi;
  if (top.f < cut_coulsq) {
    top.f = (float)cut_coulsq;
    double fv, cv, ev;
    edge(top.f, fv, cv, ev);
    tab.dr[itablemax] = 1.0 / (top.f - tab.r[itablemax]);
    tab.df[itablemax] = fv - tab.f[itablemax];
    tab.dc[itablemax] = cv - tab.c[itablemax];
    tab.de[itablemax] = ev - tab.e[itablemax];
  }
}

double HostStyle::single(int itype, int jtype, double qi, double qj, double rsq, double factor_coul,
                         double factor_lj, double &fforce) const
{
  const double EWALD_F = 1.12837917, EWALD_P = 0.3275911;  // pol.cpp:43-49
  const double A1 = 0.254829592, A2 = -0.284496736, A3 = 1.421413741, A4 = -1.453152027, A5 = 1.061405429;
  const int ij = idx(itype, jtype);
  double r2inv = 1.0 / rsq, forcecoul = 0.0, forcelj = 0.0, r6inv = 0.0;
  double prefactor = 0.0, erfcv = 0.0, fraction = 0.0;
  int itable = 0;
  const bool incoul = rsq < cut_coulsq;
  const bool analytic = !ncoultablebits || rsq <= tab.tabinnersq;
  if (incoul) {
    if (analytic) {
      double r = sqrt(rsq), grij = env.g_ewald * r, expm2 = exp(-grij * grij);
      double t = 1.0 / (1.0 + EWALD_P * grij);
      erfcv = t * (A1 + t * (A2 + t * (A3 + t * (A4 + t * A5)))) * expm2;
      prefactor = env.qqrd2e * qi * qj / r;
      forcecoul = prefactor * (erfcv + EWALD_F * grij * expm2);
      if (factor_coul < 1.0) forcecoul -= (1.0 - factor_coul) * prefactor;
    } else {
      IntFloat u;
      u.f = (float)rsq;
      itable = (u.i & tab.mask) >> tab.shift;
      fraction = ((double)u.f - tab.r[itable]) * tab.dr[itable];
      forcecoul = qi * qj * (tab.f[itable] + fraction * tab.df[itable]);
      if (factor_coul < 1.0) {
        prefactor = qi * qj * (tab.c[itable] + fraction * tab.dc[itable]);
        forcecoul -= (1.0 - factor_coul) * prefactor;
      }
    }
  }
  if (rsq < cut_ljsq[ij]) {
    r6inv = r2inv * r2inv * r2inv;
    forcelj = r6inv * (lj1[ij] * r6inv - lj2[ij]);
  }
  fforce = (forcecoul + factor_lj * forcelj) * r2inv;
  double eng = 0.0;
  if (incoul) {
    double phicoul = analytic ? prefactor * erfcv : qi * qj * (tab.e[itable] + fraction * tab.de[itable]);
    if (factor_coul < 1.0) phicoul -= (1.0 - factor_coul) * prefactor;
    eng += phicoul;
  }
  if (rsq < cut_ljsq[ij]) eng += factor_lj * (r6inv * (lj3[ij] * r6inv - lj4[ij]) - offset[ij]);
  return eng;
}

// Restart image = the bytes the reference fwrite()s, in its order: write_restart_settings
// Long-range Lennard-Jones tail correction of one type pair (init_one, pol.cpp:897-918): the integral of the
// LJ energy / virial over r > cut_lj for a uniform fluid, times the numbers of atoms of the two types.
void HostStyle::tail_correction(int i, int j, double count_i, double count_j, double &etail_ij, double &ptail_ij) const
{
  etail_ij = ptail_ij = 0.0;
  if (!tail_flag) return;
  const double pi = 3.14159265358979323846;
  const int ij = idx(i, j);
  const double sig2 = sigma[ij] * sigma[ij], sig6 = sig2 * sig2 * sig2;
  const double rc3 = cut_lj[ij] * cut_lj[ij] * cut_lj[ij], rc6 = rc3 * rc3, rc9 = rc3 * rc6;
  etail_ij = 8.0 * pi * count_i * count_j * epsilon[ij] * sig6 * (sig6 - 3.0 * rc6) / (9.0 * rc9);
  ptail_ij = 16.0 * pi * count_i * count_j * epsilon[ij] * sig6 * (2.0 * sig6 - 3.0 * rc6) / (9.0 * rc9);
}

// (cut_lj_global, cut_coul, offset_flag, mix_flag, tail_flag, ncoultablebits, tabinner; pol.cpp:976-985)
// then per i<=j: setflag and, if set, epsilon, sigma, cut_lj (pol.cpp:931-940).
// Keyword extension record.  The reference's settings block holds 7 fields and none of the polarization keywords
// (pol.cpp:976-985), and the restart format has no length prefix a foreign reader could use to skip unknown data, so the
// record is opt-in (`restart_keywords yes`): files written without it stay byte-identical to the reference's.  A reader
// recognises it by its magic -- what follows the settings in a reference-written file is an int 0/1 (the first setflag)
// or the next section of the file, never this byte string.
static const char RESTART_MAGIC[8] = {'P', 'O', 'L', 'B', '2', 'K', 'W', '1'};
static const long RESTART_SETTINGS_BYTES = 40, RESTART_EXT_BYTES = 88;

std::vector<char> HostStyle::restart_settings_image() const
{
  std::vector<char> out;
  auto put = [&](const void *p, size_t n) { out.insert(out.end(), (const char *)p, (const char *)p + n); };
  put(&cut_lj_global, 8); put(&cut_coul, 8);
  put(&offset_flag, 4); put(&mix_flag, 4); put(&tail_flag, 4); put(&ncoultablebits, 4);
  put(&tabinner, 8);
  if (restart_keywords) {
    put(RESTART_MAGIC, 8);
    put(&polar_precision, 8); put(&polar_damp, 8); put(&polar_gamma, 8); put(&polar_cutoff, 8);
    const int ints[12] = {iterations_max, damping_type, zodid, fixed_iteration, polar_gs, polar_gs_ranked,
                          use_previous, debug, gs_chunks, restart_keywords, 0, 0};
    put(ints, sizeof(ints));
  }
  return out;
}

long HostStyle::read_restart_settings_image(const void *buf, long nbytes)
{
  const char *p = (const char *)buf, *end = p + nbytes;
  auto get = [&](void *dst, size_t n) {
    if (p + n > end) fail("restart image truncated");
    memcpy(dst, p, n);
    p += n;
  };
  get(&cut_lj_global, 8); get(&cut_coul, 8);
  get(&offset_flag, 4); get(&mix_flag, 4); get(&tail_flag, 4); get(&ncoultablebits, 4);
  get(&tabinner, 8);
  if (end - p >= 8 && memcmp(p, RESTART_MAGIC, 8) == 0) {
    p += 8;
    get(&polar_precision, 8); get(&polar_damp, 8); get(&polar_gamma, 8); get(&polar_cutoff, 8);
    int ints[12];
    get(ints, sizeof(ints));
    iterations_max = ints[0]; damping_type = ints[1]; zodid = ints[2]; fixed_iteration = ints[3];
    polar_gs = ints[4]; polar_gs_ranked = ints[5]; use_previous = ints[6]; debug = ints[7];
    gs_chunks = ints[8]; restart_keywords = ints[9];
  }
  initialized = false;
  return (long)(p - (const char *)buf);
}

std::vector<char> HostStyle::restart_image() const
{
  std::vector<char> out = restart_settings_image();
  auto put = [&](const void *p, size_t n) { out.insert(out.end(), (const char *)p, (const char *)p + n); };
  for (int i = 1; i <= ntypes; i++)
    for (int j = i; j <= ntypes; j++) {
      put(&setflag[idx(i, j)], 4);
      if (setflag[idx(i, j)]) { put(&epsilon[idx(i, j)], 8); put(&sigma[idx(i, j)], 8); put(&cut_lj[idx(i, j)], 8); }
    }
  return out;
}

void HostStyle::read_restart_image(const void *buf, long nbytes)
{
  if (!allocated) fail("polb200_set_ntypes must be called before polb200_read_restart", POLB200_ERR_STATE);
  const long used = read_restart_settings_image(buf, nbytes);
  const char *p = (const char *)buf + used, *end = (const char *)buf + nbytes;
  auto get = [&](void *dst, size_t n) {
    if (p + n > end) fail("restart image truncated");
    memcpy(dst, p, n);
    p += n;
  };
  for (int i = 1; i <= ntypes; i++)
    for (int j = i; j <= ntypes; j++) {
      get(&setflag[idx(i, j)], 4);
      if (setflag[idx(i, j)]) { get(&epsilon[idx(i, j)], 8); get(&sigma[idx(i, j)], 8); get(&cut_lj[idx(i, j)], 8); }
    }
  initialized = false;
}

}  // namespace polb200
