/* ----------------------------------------------------------------------
   Shared device mirror of the per-atom arrays for the B200 drop-in styles (second half).

   The pair style, the KSpace styles and fix rigid/nve|nvt of this build all run on the GPU behind the C ABI of
   include/polb200.h, and every entry point takes device pointers (`on_device = 1`).  Without this mirror each style
   copies atom->x / v / f through the host on every step (the stock Verlet loop owns the host arrays,
   src/verlet.cpp:223-351).  With it the three styles work on ONE set of device arrays

       x v f   q alpha mu ef   type molecule tag mask nspecial special

   and the host copies are refreshed only where the stock loop or an output reads them:

     every step         x  -> host after fix initial_integrate (Neighbor::decide / check_distance read it)
     re-neighbor steps  v, mu, ef -> host in Fix::pre_exchange (before Domain::pbc / Atom::sort permute the host
                        arrays); everything host -> device again afterwards (the atoms may have a new order)
     output steps       v, f -> host after fix final_integrate; mu, ef -> host after Pair::compute
                        (thermo / dump / restart / the end of the run: update->ntimestep == output->next or laststep)

   The mirror is used for a run only when nothing else touches the per-atom state between the hooks (decided at
   init, `evaluate`): one MPI rank, run_style verlet, this build's pair style (+ ewald / pppm or no KSpace), no bonded
   styles, exactly one time-integration fix = this build's rigid/nve|nvt, and no other fix with a per-step hook.
   Anything else -- and POLB200_RESIDENT=0 -- keeps the host-buffer path.  Device memory comes from the library
   (polb200_dev_alloc / polb200_dev_copy): this file needs no CUDA tool chain.
------------------------------------------------------------------------- */

#ifndef LMP_DEVICE_ATOMS_B200_H
#define LMP_DEVICE_ATOMS_B200_H

#include <cstdlib>
#include <cstring>
#include "atom.h"
#include "comm.h"
#include "error.h"
#include "fix.h"
#include "force.h"
#include "lammps.h"
#include "modify.h"
#include "output.h"
#include "update.h"
#include "polb200.h"

namespace LAMMPS_NS {

class DeviceAtomsB200 {
 public:
  static DeviceAtomsB200 &instance()
  {
    static DeviceAtomsB200 one;
    return one;
  }

  bool resident;            // this run keeps the per-atom state on the device
  int device;
  int nlocal, cap, maxspecial, cap_special;
  double *x, *v, *f, *q, *alpha, *mu, *ef;
  int *type, *molecule, *tag, *mask, *nspecial, *special;
  bool xv_on_device;        // device x, v, f, tag hold the current state in the host's current atom order
  bool static_on_device;    // q, alpha, type, molecule, mask, special lists (and mu) uploaded since the last re-neighboring
  bool v_host_current, mu_host_current, f_host_current;
  long copies_h2d, copies_d2h;   // bytes moved, for tools/dropin_timing.py

  /* decide whether the coming run may keep its atoms on the device (called from the init() of the styles) */
  void evaluate(LAMMPS *lmp)
  {
    resident = false;
    xv_on_device = static_on_device = false;
    v_host_current = mu_host_current = f_host_current = true;
    const char *env = getenv("POLB200_RESIDENT");
    if (env && atoi(env) == 0) return;
    const char *dev = getenv("POLB200_DEVICE");
    device = dev ? atoi(dev) : 0;
    Force *force = lmp->force;
    Modify *modify = lmp->modify;
    Update *update = lmp->update;
    if (lmp->comm->nprocs != 1 || update->whichflag != 1) return;
    if (strcmp(update->integrate_style, "verlet") != 0) return;
    if (!force->pair || strcmp(force->pair_style, "lj/cut/coul/long/polarization") != 0) return;
    if (force->kspace && strcmp(force->kspace_style, "ewald") != 0 && strcmp(force->kspace_style, "pppm") != 0) return;
    if (force->bond || force->angle || force->dihedral || force->improper) return;
    const int per_step = FixConst::INITIAL_INTEGRATE | FixConst::POST_INTEGRATE | FixConst::PRE_EXCHANGE | FixConst::PRE_NEIGHBOR | FixConst::POST_NEIGHBOR | FixConst::PRE_FORCE |
                         FixConst::PRE_REVERSE | FixConst::POST_FORCE | FixConst::FINAL_INTEGRATE | FixConst::END_OF_STEP;
    int nrigid = 0;
    for (int i = 0; i < modify->nfix; i++) {
      const char *style = modify->fix[i]->style;
      if (strcmp(style, "rigid/nve") == 0 || strcmp(style, "rigid/nvt") == 0) nrigid++;
      else if (modify->fmask[i] & per_step) return;
    }
    if (nrigid != 1) return;
    resident = true;
  }

  bool output_step(LAMMPS *lmp) const
  {
    return lmp->update->ntimestep == lmp->output->next || lmp->update->ntimestep == lmp->update->laststep;
  }

  /* ---- host -> device ---- */
  void ensure_xv(LAMMPS *lmp)
  {
    if (xv_on_device) return;
    Atom *atom = lmp->atom;
    reserve(lmp, atom->nlocal, 0);
    const size_t n = (size_t) nlocal;
    if (n) {
      h2d(lmp, x, atom->x[0], 3 * n * sizeof(double));
      h2d(lmp, v, atom->v[0], 3 * n * sizeof(double));
      h2d(lmp, f, atom->f[0], 3 * n * sizeof(double));
      h2d(lmp, tag, atom->tag, n * sizeof(int));
    }
    xv_on_device = true;
    v_host_current = f_host_current = true;
  }

  void ensure_static(LAMMPS *lmp)
  {
    if (static_on_device) return;
    Atom *atom = lmp->atom;
    const int ms = (atom->molecular && atom->maxspecial > 0 && atom->special) ? atom->maxspecial : 0;
    reserve(lmp, atom->nlocal, ms);
    const size_t n = (size_t) nlocal;
    if (n) {
      h2d(lmp, q, atom->q, n * sizeof(double));
      h2d(lmp, alpha, atom->static_polarizability, n * sizeof(double));
      h2d(lmp, mu, atom->mu_induced[0], 3 * n * sizeof(double));
      h2d(lmp, ef, atom->ef_static[0], 3 * n * sizeof(double));
      h2d(lmp, type, atom->type, n * sizeof(int));
      h2d(lmp, molecule, atom->molecule, n * sizeof(int));
      h2d(lmp, mask, atom->mask, n * sizeof(int));
      if (ms) {
        h2d(lmp, nspecial, atom->nspecial[0], 3 * n * sizeof(int));
        h2d(lmp, special, atom->special[0], n * (size_t) ms * sizeof(int));
      }
    }
    maxspecial = ms;
    static_on_device = true;
    mu_host_current = true;
  }

  void upload_f(LAMMPS *lmp)
  {
    if (nlocal) h2d(lmp, f, lmp->atom->f[0], 3 * (size_t) nlocal * sizeof(double));
    f_host_current = true;
  }

  /* ---- device -> host ---- */
  void download_x(LAMMPS *lmp) { if (nlocal) d2h(lmp, lmp->atom->x[0], x, 3 * (size_t) nlocal * sizeof(double)); }
  void download_v(LAMMPS *lmp)
  {
    if (v_host_current || !xv_on_device) return;
    if (nlocal) d2h(lmp, lmp->atom->v[0], v, 3 * (size_t) nlocal * sizeof(double));
    v_host_current = true;
  }
  void download_f(LAMMPS *lmp)
  {
    if (f_host_current || !xv_on_device) return;
    if (nlocal) d2h(lmp, lmp->atom->f[0], f, 3 * (size_t) nlocal * sizeof(double));
    f_host_current = true;
  }
  void download_mu(LAMMPS *lmp)
  {
    if (mu_host_current || !static_on_device) return;
    if (nlocal) {
      d2h(lmp, lmp->atom->mu_induced[0], mu, 3 * (size_t) nlocal * sizeof(double));
      d2h(lmp, lmp->atom->ef_static[0], ef, 3 * (size_t) nlocal * sizeof(double));
    }
    mu_host_current = true;
  }

  /* the host is about to wrap / sort / exchange its atoms (Fix::pre_exchange): make it authoritative */
  void before_reneighbor(LAMMPS *lmp)
  {
    download_v(lmp);
    download_mu(lmp);
  }
  /* ... and did so (Fix::pre_neighbor): the device copies are in the old order */
  void after_reneighbor() { xv_on_device = static_on_device = false; }

  void zero_f(LAMMPS *lmp)
  {
    if (nlocal && polb200_dev_zero(device, f, 3 * (size_t) nlocal * sizeof(double)) != POLB200_OK)
      lmp->error->one(FLERR, "B200 device mirror: clearing the forces failed");
    f_host_current = false;
  }

 private:
  DeviceAtomsB200()
      : resident(false), device(0), nlocal(0), cap(0), maxspecial(0), cap_special(0), x(NULL), v(NULL), f(NULL), q(NULL),
        alpha(NULL), mu(NULL), ef(NULL), type(NULL), molecule(NULL), tag(NULL), mask(NULL), nspecial(NULL), special(NULL),
        xv_on_device(false), static_on_device(false), v_host_current(true), mu_host_current(true), f_host_current(true),
        copies_h2d(0), copies_d2h(0)
  {
  }

  template <class T>
  void grow(LAMMPS *lmp, T *&p, size_t count)
  {
    if (p) polb200_dev_free(device, p);
    p = static_cast<T *>(polb200_dev_alloc(device, count * sizeof(T)));
    if (!p) lmp->error->one(FLERR, "B200 device mirror: out of device memory");
  }

  void reserve(LAMMPS *lmp, int n, int ms)
  {
    if (n > cap) {
      const size_t c = (size_t) n + n / 8 + 64;
      grow(lmp, x, 3 * c); grow(lmp, v, 3 * c); grow(lmp, f, 3 * c); grow(lmp, mu, 3 * c); grow(lmp, ef, 3 * c);
      grow(lmp, q, c); grow(lmp, alpha, c);
      grow(lmp, type, c); grow(lmp, molecule, c); grow(lmp, tag, c); grow(lmp, mask, c); grow(lmp, nspecial, 3 * c);
      cap = (int) c;
      cap_special = 0;
      xv_on_device = static_on_device = false;
    }
    if (ms > 0 && (size_t) cap * ms > (size_t) cap_special) {
      grow(lmp, special, (size_t) cap * ms);
      cap_special = cap * ms;
    }
    nlocal = n;
  }

  void h2d(LAMMPS *lmp, void *dst, const void *src, size_t bytes)
  {
    if (polb200_dev_copy(device, dst, src, bytes, POLB200_COPY_H2D) != POLB200_OK)
      lmp->error->one(FLERR, "B200 device mirror: host to device copy failed");
    copies_h2d += (long) bytes;
  }
  void d2h(LAMMPS *lmp, void *dst, const void *src, size_t bytes)
  {
    if (polb200_dev_copy(device, dst, src, bytes, POLB200_COPY_D2H) != POLB200_OK)
      lmp->error->one(FLERR, "B200 device mirror: device to host copy failed");
    copies_d2h += (long) bytes;
  }
};

}  // namespace LAMMPS_NS

#endif
