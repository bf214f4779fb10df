/* ----------------------------------------------------------------------
   atom_style full + polarization arrays (see the header).  The stock AtomVecFull does the work for its own fields;
   this class adds the three arrays: as a TRAILER behind the stock exchange / restart records (the stock unpackers
   are handed a shortened record length, so fix extras keep working), and with its own border records.
------------------------------------------------------------------------- */

#include <cstdlib>
#include <cstring>
#include "atom_vec_full_polar_b200.h"
#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "fix.h"
#include "memory.h"
#include "modify.h"

using namespace LAMMPS_NS;

AtomVecFullPolar::AtomVecFullPolar(LAMMPS *lmp) : AtomVecFull(lmp), alpha(NULL), efield(NULL), dipole(NULL)
{
  size_border += 1;                       // static_polarizability travels with the ghosts
  atom->static_polarizability_flag = 1;   // what PairLJCutCoulLongPolarization::init_style checks (pol.cpp:812-813)
}

/* like the stock class, keep our own pointers to the arrays: `replicate` packs the atoms of the OLD Atom instance
   through its avec after lmp->atom already points to the new one */

void AtomVecFullPolar::grow(int n)
{
  AtomVecFull::grow(n);
  alpha = memory->grow(atom->static_polarizability,nmax,"atom:static_polarizability");
  efield = memory->grow(atom->ef_static,nmax,3,"atom:ef_static");
  dipole = memory->grow(atom->mu_induced,nmax,3,"atom:mu_induced");
}

void AtomVecFullPolar::grow_reset()
{
  AtomVecFull::grow_reset();
  alpha = atom->static_polarizability;
  efield = atom->ef_static;
  dipole = atom->mu_induced;
}

void AtomVecFullPolar::clear_polar(int i)
{
  alpha[i] = 0.0;
  for (int k = 0; k < 3; k++) efield[i][k] = dipole[i][k] = 0.0;
}

void AtomVecFullPolar::copy(int i, int j, int delflag)
{
  alpha[j] = alpha[i];
  for (int k = 0; k < 3; k++) {
    efield[j][k] = efield[i][k];
    dipole[j][k] = dipole[i][k];
  }
  AtomVecFull::copy(i,j,delflag);
}

/* ---- border records: x y z tag type mask q molecule alpha [vx vy vz] ---- */

template <int VEL>
int AtomVecFullPolar::pack_border_any(int n, int *list, double *buf, int pbc_flag, int *pbc)
{
  double shift[3] = {0.0,0.0,0.0}, dv[3] = {0.0,0.0,0.0};
  if (pbc_flag) {
    if (domain->triclinic == 0) {
      shift[0] = pbc[0]*domain->xprd; shift[1] = pbc[1]*domain->yprd; shift[2] = pbc[2]*domain->zprd;
    } else {
      shift[0] = pbc[0]; shift[1] = pbc[1]; shift[2] = pbc[2];
    }
    if (VEL && deform_vremap) {
      dv[0] = pbc[0]*h_rate[0] + pbc[5]*h_rate[5] + pbc[4]*h_rate[4];
      dv[1] = pbc[1]*h_rate[1] + pbc[3]*h_rate[3];
      dv[2] = pbc[2]*h_rate[2];
    }
  }
  int m = 0;
  for (int i = 0; i < n; i++) {
    const int j = list[i];
    for (int k = 0; k < 3; k++) buf[m++] = pbc_flag ? x[j][k] + shift[k] : x[j][k];
    buf[m++] = ubuf(tag[j]).d;
    buf[m++] = ubuf(type[j]).d;
    buf[m++] = ubuf(mask[j]).d;
    buf[m++] = q[j];
    buf[m++] = ubuf(molecule[j]).d;
    buf[m++] = alpha[j];
    if (VEL) {
      const int remap = pbc_flag && deform_vremap && (mask[j] & deform_groupbit);
      for (int k = 0; k < 3; k++) buf[m++] = remap ? v[j][k] + dv[k] : v[j][k];
    }
  }
  if (atom->nextra_border)
    for (int iextra = 0; iextra < atom->nextra_border; iextra++)
      m += modify->fix[atom->extra_border[iextra]]->pack_border(n,list,&buf[m]);
  return m;
}

template <int VEL>
void AtomVecFullPolar::unpack_border_any(int n, int first, double *buf)
{
  int m = 0;
  const int last = first + n;
  for (int i = first; i < last; i++) {
    if (i == nmax) grow(0);
    for (int k = 0; k < 3; k++) x[i][k] = buf[m++];
    tag[i] = (tagint) ubuf(buf[m++]).i;
    type[i] = (int) ubuf(buf[m++]).i;
    mask[i] = (int) ubuf(buf[m++]).i;
    q[i] = buf[m++];
    molecule[i] = (tagint) ubuf(buf[m++]).i;
    alpha[i] = buf[m++];
    if (VEL)
      for (int k = 0; k < 3; k++) v[i][k] = buf[m++];
  }
  if (atom->nextra_border)
    for (int iextra = 0; iextra < atom->nextra_border; iextra++)
      m += modify->fix[atom->extra_border[iextra]]->unpack_border(n,first,&buf[m]);
}

int AtomVecFullPolar::pack_border(int n, int *list, double *buf, int pbc_flag, int *pbc)
{
  return pack_border_any<0>(n,list,buf,pbc_flag,pbc);
}

int AtomVecFullPolar::pack_border_vel(int n, int *list, double *buf, int pbc_flag, int *pbc)
{
  return pack_border_any<1>(n,list,buf,pbc_flag,pbc);
}

void AtomVecFullPolar::unpack_border(int n, int first, double *buf)
{
  unpack_border_any<0>(n,first,buf);
}

void AtomVecFullPolar::unpack_border_vel(int n, int first, double *buf)
{
  unpack_border_any<1>(n,first,buf);
}

int AtomVecFullPolar::pack_border_hybrid(int n, int *list, double *buf)
{
  int m = AtomVecFull::pack_border_hybrid(n,list,buf);
  for (int i = 0; i < n; i++) buf[m++] = alpha[list[i]];
  return m;
}

int AtomVecFullPolar::unpack_border_hybrid(int n, int first, double *buf)
{
  int m = AtomVecFull::unpack_border_hybrid(n,first,buf);
  for (int i = first; i < first+n; i++) alpha[i] = buf[m++];
  return m;
}

/* ---- exchange: stock record + fix extras, then alpha mu[3] E[3]; buf[0] = total length ---- */

int AtomVecFullPolar::pack_exchange(int i, double *buf)
{
  int m = AtomVecFull::pack_exchange(i,buf);
  buf[m++] = alpha[i];
  for (int k = 0; k < 3; k++) buf[m++] = dipole[i][k];
  for (int k = 0; k < 3; k++) buf[m++] = efield[i][k];
  buf[0] = m;
  return m;
}

int AtomVecFullPolar::unpack_exchange(double *buf)
{
  const int total = static_cast<int> (buf[0]);
  int m = AtomVecFull::unpack_exchange(buf);       // increments atom->nlocal; may grow() the arrays
  if (m != total - NEXCHANGE) error->one(FLERR,"Atom style full (polar): corrupt exchange record");
  const int i = atom->nlocal - 1;
  alpha[i] = buf[m++];
  for (int k = 0; k < 3; k++) dipole[i][k] = buf[m++];
  for (int k = 0; k < 3; k++) efield[i][k] = buf[m++];
  return m;
}

/* ---- restart: stock record + fix extras, then alpha mu[3] ---- */

int AtomVecFullPolar::size_restart()
{
  return AtomVecFull::size_restart() + NRESTART * atom->nlocal;
}

int AtomVecFullPolar::pack_restart(int i, double *buf)
{
  int m = AtomVecFull::pack_restart(i,buf);
  buf[m++] = alpha[i];
  for (int k = 0; k < 3; k++) buf[m++] = dipole[i][k];
  buf[0] = m;
  return m;
}

int AtomVecFullPolar::unpack_restart(double *buf)
{
  const int total = static_cast<int> (buf[0]);
  buf[0] = total - NRESTART;                       // the stock reader takes "everything up to buf[0]" as fix extras
  AtomVecFull::unpack_restart(buf);                // increments atom->nlocal; may grow() the arrays
  buf[0] = total;
  const int i = atom->nlocal - 1;
  int m = total - NRESTART;
  alpha[i] = buf[m++];
  for (int k = 0; k < 3; k++) dipole[i][k] = buf[m++];
  for (int k = 0; k < 3; k++) efield[i][k] = 0.0;
  return total;
}

/* ---- new atoms start unpolarizable; `set ... static_polarizability` assigns the values ---- */

void AtomVecFullPolar::create_atom(int itype, double *coord)
{
  AtomVecFull::create_atom(itype,coord);
  clear_polar(atom->nlocal - 1);
}

void AtomVecFullPolar::data_atom(double *coord, imageint imagetmp, char **values)
{
  AtomVecFull::data_atom(coord,imagetmp,values);
  clear_polar(atom->nlocal - 1);
}

int AtomVecFullPolar::data_atom_hybrid(int nlocal, char **values)
{
  const int n = AtomVecFull::data_atom_hybrid(nlocal,values);
  clear_polar(nlocal);
  return n;
}

bigint AtomVecFullPolar::memory_usage()
{
  bigint bytes = AtomVecFull::memory_usage();
  if (atom->memcheck("static_polarizability")) bytes += memory->usage(atom->static_polarizability,nmax);
  if (atom->memcheck("ef_static")) bytes += memory->usage(atom->ef_static,nmax,3);
  if (atom->memcheck("mu_induced")) bytes += memory->usage(atom->mu_induced,nmax,3);
  return bytes;
}
