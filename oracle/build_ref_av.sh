#!/bin/bash
# CPU test vehicle for the committed atom style: oracle/_ref/lmp_serial_av =
#   the reference (scratch copy of $POLB200_REFERENCE/src, repaired with the shim headers like build_ref.sh, its OWN
#   CPU pair style untouched, NO AtomVecFull patch)
#   + lammps/atom_vec_full_polar_b200.{h,cpp} + lammps/compute_polarization_atom_b200.{h,cpp}  (host C++, no CUDA).
# tests/test_atom_style.py runs the shipped example through it and must get the reference's committed log; restart
# round trips and atom sorting are checked the same way -- all without a GPU.  Test infrastructure: never shipped.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
LMPDIR="$ROOT/lammps-induced-dipole-polarization-pair-style_b200/lammps"
REF="${POLB200_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
W="${POLB200_REFAV_SCRATCH:-/tmp/polb200_refavbuild}"
JOBS="${JOBS:-$(nproc)}"
if [ ! -d "$REF/src" ]; then
  echo "build_ref_av: $REF/src not present (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
if [ -x "$OUT/lmp_serial_av" ] && [ -x "$OUT/extract_driver_av" ] && [ -z "${POLB200_REF_REBUILD:-}" ]; then
  new=0
  for f in atom_vec_full_polar_b200.h atom_vec_full_polar_b200.cpp compute_polarization_atom_b200.h compute_polarization_atom_b200.cpp extract_driver_b200.cpp; do
    [ "$LMPDIR/$f" -nt "$OUT/lmp_serial_av" ] && new=1
  done
  if [ $new = 0 ]; then echo "build_ref_av: $OUT/lmp_serial_av is up to date"; exit 0; fi
fi
if [ ! -f "$W/src/Obj_serial/lammps.o" ]; then
  rm -rf "$W"; mkdir -p "$W"
  cp -r "$REF/src" "$W/src"
  chmod -R u+w "$W"
  cd "$W/src"
  rm -f STUBS/libmpi_stubs.a STUBS/*.o
  make yes-kspace yes-molecule yes-rigid > "$W/install.log" 2>&1
  (cd STUBS && make > "$W/stubs.log" 2>&1)
  for f in accelerator_kokkos.h accelerator_omp.h atom_vec_ellipsoid.h dihedral_hybrid.h improper_hybrid.h math_complex.h math_vector.h; do
    [ -f "$HERE/ref_shims/$f" ] && cp "$HERE/ref_shims/$f" .
  done
  rm -f compute_dihedral.* compute_improper.* fix_nve_sphere.* fix_nh_sphere.* fix_nvt_sphere.* \
        fix_npt_sphere.* fix_nph_sphere.* pair_lj_long_coul_long.* pair_buck_long_coul_long.* \
        pair_lj_long_tip4p_long.* ewald_disp.*
fi
cd "$W/src"
sed -i 's|^AtomStyle(full,AtomVecFull)|AtomStyle(full/stock,AtomVecFull)|' atom_vec_full.h
cp "$LMPDIR/atom_vec_full_polar_b200.h" "$LMPDIR/atom_vec_full_polar_b200.cpp" .
cp "$LMPDIR/compute_polarization_atom_b200.h" "$LMPDIR/compute_polarization_atom_b200.cpp" .
# Atom::extract learns the three per-atom arrays: what lammps_extract_atom / the Python module see
if ! grep -q '"mu_induced"' atom.cpp; then
  sed -i 's|^  if (strcmp(name,"mass") == 0) return (void \*) mass;|  if (strcmp(name,"static_polarizability") == 0) return (void *) static_polarizability;\n  if (strcmp(name,"mu_induced") == 0) return (void *) mu_induced;\n  if (strcmp(name,"ef_static") == 0) return (void *) ef_static;\n&|' atom.cpp
  grep -q '"mu_induced"' atom.cpp || { echo "Atom::extract patch did not apply"; exit 1; }
fi
make -j"$JOBS" serial > "$W/build.log" 2>&1 || { tail -40 "$W/build.log"; exit 1; }
cp lmp_serial "$OUT/lmp_serial_av"
LMPSRC="$LMPDIR"
# library-interface driver (tests): the same objects without main.o
g++ -g -O -I. -ISTUBS -c "$LMPSRC/extract_driver_b200.cpp" -o Obj_serial/extract_driver_b200.o
g++ -g -O Obj_serial/extract_driver_b200.o $(for f in *.cpp; do [ "$f" = main.cpp ] || echo "Obj_serial/${f%.cpp}.o"; done) -LSTUBS -lmpi_stubs  -o "$OUT/extract_driver_av"
strip "$OUT/extract_driver_av"

strip "$OUT/lmp_serial_av"
echo "build_ref_av: built $OUT/lmp_serial_av"
