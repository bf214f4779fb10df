/* ----------------------------------------------------------------------
   Library-interface driver for the per-atom arrays of the polarization pair style:
   runs a LAMMPS input script through the C library interface (src/library.h) and prints, per atom,

       id  static_polarizability  mu_induced[3]  ef_static[3]

   as `lammps_extract_atom` returns them -- what a Python / C driver of LAMMPS sees.  The three names are added to
   Atom::extract (src/atom.cpp:2174) by the build scripts (three lines, the same idiom as the names around them).

     extract_driver in.script > table.txt

   With a second argument `exchange` it exercises the atom style's exchange record instead (what Comm::exchange sends when
   an atom changes MPI ranks, src/comm_brick.cpp:597-690 -- this build is single-rank, so nothing else reaches it): every
   atom is packed with AtomVec::pack_exchange and unpacked again as a NEW atom behind the owned ones; the copy must carry
   the same position, charge, molecule id, special list, polarizability, dipole and static field.  Prints the record length
   and the largest difference.
------------------------------------------------------------------------- */

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "library.h"
#include "lammps.h"
#include "atom.h"
#include "atom_vec.h"


int main(int argc, char **argv)
{
  if (argc < 2) {
    fprintf(stderr, "usage: extract_driver in.script\n");
    return 2;
  }
  char *args[] = {(char *) "extract_driver", (char *) "-log", (char *) "none", (char *) "-screen", (char *) "none"};
  void *lmp = NULL;
  lammps_open_no_mpi(5, args, &lmp);
  if (!lmp) return 3;
  lammps_file(lmp, argv[1]);
  const int n = lammps_get_natoms(lmp);
  if (argc > 2 && strcmp(argv[2], "exchange") == 0) {
    LAMMPS_NS::LAMMPS *l = (LAMMPS_NS::LAMMPS *) lmp;
    LAMMPS_NS::Atom *atom = l->atom;
    double buf[512];
    double worst = 0.0;
    int len_min = 1 << 30, len_max = 0;
    for (int i = 0; i < atom->nlocal; i++) {
      if (atom->nlocal == atom->nmax) atom->avec->grow(0);
      const int m = atom->avec->pack_exchange(i, buf);
      const int k = atom->nlocal;
      const int used = atom->avec->unpack_exchange(buf);   // appends the atom at index nlocal
      if (used != m || atom->nlocal != k + 1) {
        fprintf(stderr, "extract_driver: exchange record length mismatch (%d packed, %d unpacked)\n", m, used);
        return 5;
      }
      double d = fabs(atom->q[k] - atom->q[i]) + fabs(atom->static_polarizability[k] - atom->static_polarizability[i]);
      for (int c = 0; c < 3; c++)
        d += fabs(atom->x[k][c] - atom->x[i][c]) + fabs(atom->v[k][c] - atom->v[i][c]) +
             fabs(atom->mu_induced[k][c] - atom->mu_induced[i][c]) + fabs(atom->ef_static[k][c] - atom->ef_static[i][c]);
      d += (atom->tag[k] != atom->tag[i]) + (atom->type[k] != atom->type[i]) + (atom->molecule[k] != atom->molecule[i]) +
           (atom->image[k] != atom->image[i]) + (atom->mask[k] != atom->mask[i]);
      for (int c = 0; c < 3; c++) d += atom->nspecial[k][c] != atom->nspecial[i][c];
      for (int c = 0; c < atom->nspecial[i][2]; c++) d += atom->special[k][c] != atom->special[i][c];
      if (d > worst) worst = d;
      if (m < len_min) len_min = m;
      if (m > len_max) len_max = m;
      atom->nlocal = k;   // drop the copy again
    }
    printf("exchange atoms %d record_doubles %d %d max_difference %.17g\n", atom->nlocal, len_min, len_max, worst);
    lammps_close(lmp);
    return 0;
  }
  int *id = (int *) lammps_extract_atom(lmp, (char *) "id");
  double *alpha = (double *) lammps_extract_atom(lmp, (char *) "static_polarizability");
  double **mu = (double **) lammps_extract_atom(lmp, (char *) "mu_induced");
  double **ef = (double **) lammps_extract_atom(lmp, (char *) "ef_static");
  if (!id || !alpha || !mu || !ef) {
    fprintf(stderr, "extract_driver: lammps_extract_atom does not know the polarization arrays\n");
    lammps_close(lmp);
    return 4;
  }
  for (int i = 0; i < n; i++)
    printf("%d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", id[i], alpha[i], mu[i][0], mu[i][1], mu[i][2], ef[i][0], ef[i][1],
           ef[i][2]);
  lammps_close(lmp);
  return 0;
}
