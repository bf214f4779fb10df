#!/bin/bash
# Build the REPAIRED reference (LAMMPS 16Mar2018 fork + polarization pair style) as the parity
# oracle binary oracle/_ref/lmp_serial.  Test infrastructure only -- never linked into the product.
#
#   * the reference sources are compiled from a scratch COPY (the tree itself does not build as
#     shipped: 10 git-ignored upstream files are missing); nothing is copied
#     into this repository, outputs go only to oracle/_ref/ (git-ignored);
#   * we do not run the reference's build system beyond its plain `make serial` compile rules
#     (g++ over src/*.cpp with the MPI STUBS library) -- no cmake, no external libraries.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${POLB200_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
W="${POLB200_REF_SCRATCH:-/tmp/polb200_refbuild}"
JOBS="${JOBS:-$(nproc)}"

if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF/src not present (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
if [ -x "$OUT/lmp_serial" ] && [ -z "${POLB200_REF_REBUILD:-}" ]; then
  echo "build_ref: $OUT/lmp_serial exists (set POLB200_REF_REBUILD=1 to rebuild)"
  exit 0
fi
rm -rf "$W"; mkdir -p "$W"
cp -r "$REF/src" "$W/src"
chmod -R u+w "$W"
cd "$W/src"
rm -f STUBS/libmpi_stubs.a STUBS/*.o
make yes-kspace yes-molecule yes-rigid > "$W/install.log" 2>&1
(cd STUBS && make > "$W/stubs.log" 2>&1)
# 1. shims for upstream files that are git-ignored and therefore absent from the fork
cp "$HERE"/ref_shims/*.h .
# 2. translation units that need the absent math_vector.h / math_complex.h / *_hybrid.cpp
rm -f compute_dihedral.* compute_improper.* fix_nve_sphere.* fix_nh_sphere.* fix_nvt_sphere.* \
      fix_npt_sphere.* fix_nph_sphere.* pair_lj_long_coul_long.* pair_buck_long_coul_long.* \
      pair_lj_long_tip4p_long.* ewald_disp.*
# 3. atom style that carries the polarization arrays + dump hooks
python3 "$HERE/patch_ref.py" "$W/src"
make -j"$JOBS" serial > "$W/build.log" 2>&1 || { tail -40 "$W/build.log"; exit 1; }
cp lmp_serial "$OUT/lmp_serial"
strip "$OUT/lmp_serial"
echo "build_ref: built $OUT/lmp_serial"
