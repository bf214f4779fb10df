#!/bin/bash
# Build a LAMMPS binary in which the reference's pair style is REPLACED by the B200 drop-in:
#   lammps/_build/lmp_b200  =  reference host framework (LAMMPS 16Mar2018, from a scratch copy of
#   $POLB200_REFERENCE/src, repaired exactly like the oracle build: oracle/build_ref.sh steps 1-3 without the
#   dump hooks)  +  pair_lj_cut_coul_long_polarization_b200.{h,cpp}  +  ewald_b200.{h,cpp}  +  pppm_b200.{h,cpp}  +
#   fix_rigid_nh_b200.{h,cpp}  +  atom_vec_full_polar_b200.{h,cpp}  +  compute_polarization_atom_b200.{h,cpp}  +
#   device_atoms_b200.h  +  libpolb200.so.
#   (the oracle's 12-line AtomVecFull patch is NOT used here: the committed atom style replaces it)
# An unchanged input script (polarization/examples/*) then drives the CUDA path.  Nothing of the reference is
# copied into this repository; the binary lands in lammps/_build/ (git-ignored, travels to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
PKG="$(dirname "$HERE")"
ROOT="$(dirname "$PKG")"
REF="${POLB200_REFERENCE:-/root/reference}"
OUT="$HERE/_build"
W="${POLB200_LMP_SCRATCH:-/tmp/polb200_lmpbuild}"
JOBS="${JOBS:-$(nproc)}"
if [ ! -d "$REF/src" ]; then
  echo "build_lmp_b200: $REF/src not present (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
[ -f "$PKG/libpolb200.so" ] || make -C "$PKG/csrc"
mkdir -p "$OUT"
if [ -x "$OUT/lmp_b200" ] && [ "$OUT/lmp_b200" -nt "$HERE/pair_lj_cut_coul_long_polarization_b200.cpp" ] \
   && [ "$OUT/lmp_b200" -nt "$HERE/pair_lj_cut_coul_long_polarization_b200.h" ] \
   && [ "$OUT/lmp_b200" -nt "$HERE/ewald_b200.cpp" ] && [ "$OUT/lmp_b200" -nt "$HERE/ewald_b200.h" ] \
   && [ "$OUT/lmp_b200" -nt "$HERE/atom_vec_full_polar_b200.cpp" ] && [ "$OUT/lmp_b200" -nt "$HERE/atom_vec_full_polar_b200.h" ] \
   && [ "$OUT/lmp_b200" -nt "$HERE/compute_polarization_atom_b200.cpp" ] && [ "$OUT/lmp_b200" -nt "$HERE/compute_polarization_atom_b200.h" ] \
   && [ "$OUT/lmp_b200" -nt "$HERE/pppm_b200.cpp" ] && [ "$OUT/lmp_b200" -nt "$HERE/pppm_b200.h" ] \
   && [ "$OUT/lmp_b200" -nt "$HERE/fix_rigid_nh_b200.cpp" ] && [ "$OUT/lmp_b200" -nt "$HERE/fix_rigid_nh_b200.h" ] \
   && [ "$OUT/lmp_b200" -nt "$HERE/device_atoms_b200.h" ] && [ -x "$OUT/extract_driver_b200" ] \
   && [ "$OUT/extract_driver_b200" -nt "$HERE/extract_driver_b200.cpp" ] \
   && [ "$OUT/lmp_b200" -nt "$ROOT/include/polb200.h" ] && [ -z "${POLB200_LMP_REBUILD:-}" ]; then
  echo "build_lmp_b200: $OUT/lmp_b200 is up to date"
  exit 0
fi
if [ ! -f "$W/src/Obj_serial/lammps.o" ]; then
  rm -rf "$W"; mkdir -p "$W"
  cp -r "$REF/src" "$W/src"
  chmod -R u+w "$W"
  cd "$W/src"
  rm -f STUBS/libmpi_stubs.a STUBS/*.o
  make yes-kspace yes-molecule yes-rigid > "$W/install.log" 2>&1
  (cd STUBS && make > "$W/stubs.log" 2>&1)
  for f in accelerator_kokkos.h accelerator_omp.h atom_vec_ellipsoid.h dihedral_hybrid.h improper_hybrid.h; do
    cp "$ROOT/oracle/ref_shims/$f" .
  done
  rm -f compute_dihedral.* compute_improper.* fix_nve_sphere.* fix_nh_sphere.* fix_nvt_sphere.* \
        fix_npt_sphere.* fix_nph_sphere.* pair_lj_long_coul_long.* pair_buck_long_coul_long.* \
        pair_lj_long_tip4p_long.* ewald_disp.*
fi
cd "$W/src"
# atom_style full that carries the polarization arrays: the stock class stays as the base and gives up
# its style name; our subclass registers as `full`.  compute polarization/atom exposes dipoles and fields.
cp "$REF/src/MOLECULE/atom_vec_full.h" "$REF/src/MOLECULE/atom_vec_full.cpp" .
chmod u+w atom_vec_full.h atom_vec_full.cpp
sed -i 's|^AtomStyle(full,AtomVecFull)|AtomStyle(full/stock,AtomVecFull)|' atom_vec_full.h
cp "$HERE/atom_vec_full_polar_b200.h" "$HERE/atom_vec_full_polar_b200.cpp" .
cp "$HERE/compute_polarization_atom_b200.h" "$HERE/compute_polarization_atom_b200.cpp" .
# the swap: the reference's implementation leaves, the drop-in takes its file names
cp "$HERE/pair_lj_cut_coul_long_polarization_b200.h" pair_lj_cut_coul_long_polarization.h
cp "$HERE/pair_lj_cut_coul_long_polarization_b200.cpp" pair_lj_cut_coul_long_polarization.cpp
# ... and so does the KSpace style every input of the pair style uses
cp "$HERE/ewald_b200.h" ewald.h
cp "$HERE/ewald_b200.cpp" ewald.cpp
cp "$HERE/pppm_b200.h" pppm.h
cp "$HERE/pppm_b200.cpp" pppm.cpp
rm -f pppm_cg.* pppm_stagger.* pppm_tip4p.*     # derived from the reference's class PPPM
# ... and the integrator of every shipped example, fix rigid/nve|nvt: one class under both style names
cp "$HERE/fix_rigid_nh_b200.h" fix_rigid_nve.h
cp "$HERE/fix_rigid_nh_b200.cpp" fix_rigid_nve.cpp
rm -f fix_rigid_nvt.h fix_rigid_nvt.cpp
# shared device mirror of atom->x / v / f / q / mu for the three styles above (second half)
cp "$HERE/device_atoms_b200.h" .
cp "$ROOT/include/polb200.h" .
# Atom::extract learns the three per-atom arrays: what lammps_extract_atom / the Python module see
if ! grep -q '"mu_induced"' atom.cpp; then
  sed -i 's|^  if (strcmp(name,"mass") == 0) return (void \*) mass;|  if (strcmp(name,"static_polarizability") == 0) return (void *) static_polarizability;\n  if (strcmp(name,"mu_induced") == 0) return (void *) mu_induced;\n  if (strcmp(name,"ef_static") == 0) return (void *) ef_static;\n&|' atom.cpp
  grep -q '"mu_induced"' atom.cpp || { echo "Atom::extract patch did not apply"; exit 1; }
fi
make -j"$JOBS" serial LIB="-L$PKG -lpolb200 -Wl,-rpath,'\$\$ORIGIN/../..'" > "$W/build.log" 2>&1 || { tail -40 "$W/build.log"; exit 1; }
cp lmp_serial "$OUT/lmp_b200"
strip "$OUT/lmp_b200"
LMPSRC="$HERE"
# library-interface driver (tests): the same objects without main.o
g++ -g -O -I. -ISTUBS -c "$LMPSRC/extract_driver_b200.cpp" -o Obj_serial/extract_driver_b200.o
g++ -g -O Obj_serial/extract_driver_b200.o $(for f in *.cpp; do [ "$f" = main.cpp ] || echo "Obj_serial/${f%.cpp}.o"; done) -LSTUBS -lmpi_stubs -L"$PKG" -lpolb200 -Wl,-rpath,'$ORIGIN/../..'  -o "$OUT/extract_driver_b200"
strip "$OUT/extract_driver_b200"

echo "build_lmp_b200: built $OUT/lmp_b200"
