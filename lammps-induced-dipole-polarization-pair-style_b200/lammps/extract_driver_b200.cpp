/* ----------------------------------------------------------------------
   Library-interface driver for the per-atom arrays of the polarization pair style:
   runs a LAMMPS input script through the C library interface (src/library.h) and prints, per atom,

       id  static_polarizability  mu_induced[3]  ef_static[3]

   as `lammps_extract_atom` returns them -- what a Python / C driver of LAMMPS sees.  The three names are added to
   Atom::extract (src/atom.cpp:2174) by the build scripts (three lines, the same idiom as the names around them).

     extract_driver in.script > table.txt
------------------------------------------------------------------------- */

#include <cstdio>
#include <cstdlib>
#include "library.h"


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
