/* polb200 oracle shim: upstream LAMMPS keeps this header out of git (src/.gitignore);
   the serial build needs nothing from it. */
