#!/usr/bin/env python3
"""Patch a SCRATCH COPY of the reference src/ so that it builds and can serve as oracle.

Usage: patch_ref.py <scratch_src_dir>

Never run on /root/reference (read-only).  Implements SURVEY.md Appendix A:
  * atom_vec_full.cpp (installed copy): allocate / copy / border-communicate the three
    polarization per-atom arrays that the fork declares in atom.h:160-163 but no atom style
    allocates (the author's modified file is git-ignored by src/.gitignore:183);
  * pair_lj_cut_coul_long_polarization.cpp: two one-line calls into polb200_dump.h (our code)
    so that per-atom dipoles / fields / forces, which the reference never exposes, can be
    written out for golden fixtures.  The numerics of compute() are untouched.
"""
import os
import re
import sys
from pathlib import Path


def sub_once(text, old, new, count=1, what=""):
    n = text.count(old)
    if n < count:
        raise SystemExit(f"patch anchor not found ({what}): {old!r} found {n}x, need {count}")
    return text.replace(old, new) if count == n else text.replace(old, new, count)


def patch_atom_vec_full(path: Path):
    t = path.read_text()
    t = sub_once(t, "size_border = 8;", "size_border = 9;", what="ctor size_border")
    t = sub_once(t, "atom->molecule_flag = atom->q_flag = 1;",
                 "atom->molecule_flag = atom->q_flag = 1;\n  atom->static_polarizability_flag = 1;",
                 what="ctor flag")
    t = sub_once(t, '  q = memory->grow(atom->q,nmax,"atom:q");\n',
                 '  q = memory->grow(atom->q,nmax,"atom:q");\n'
                 '  memory->grow(atom->static_polarizability,nmax,"atom:static_polarizability");\n'
                 '  memory->grow(atom->ef_static,nmax,3,"atom:ef_static");\n'
                 '  memory->grow(atom->mu_induced,nmax,3,"atom:mu_induced");\n',
                 what="grow")
    t = sub_once(t, "  q[j] = q[i];\n",
                 "  q[j] = q[i];\n"
                 "  atom->static_polarizability[j] = atom->static_polarizability[i];\n"
                 "  for (k = 0; k < 3; k++) {\n"
                 "    atom->ef_static[j][k] = atom->ef_static[i][k];\n"
                 "    atom->mu_induced[j][k] = atom->mu_induced[i][k];\n"
                 "  }\n", what="copy")
    # pack_border (two branches) and unpack_border only: the first two / first occurrences
    # inside those functions.  Split the file at function boundaries to stay precise.
    def patch_func(t, header, old, new, count):
        a = t.index(header)
        b = t.index("\n}\n", a) + 3
        body = t[a:b]
        if body.count(old) != count:
            raise SystemExit(f"{header}: expected {count} x {old!r}, got {body.count(old)}")
        return t[:a] + body.replace(old, new) + t[b:]
    t = patch_func(t, "int AtomVecFull::pack_border(int n, int *list, double *buf,",
                   "      buf[m++] = q[j];\n",
                   "      buf[m++] = q[j];\n      buf[m++] = atom->static_polarizability[j];\n", 2)
    t = patch_func(t, "void AtomVecFull::unpack_border(int n, int first, double *buf)",
                   "    q[i] = buf[m++];\n",
                   "    q[i] = buf[m++];\n    atom->static_polarizability[i] = buf[m++];\n", 1)
    zero = ("  atom->static_polarizability[nlocal] = 0.0;\n"
            "  atom->ef_static[nlocal][0] = atom->ef_static[nlocal][1] = atom->ef_static[nlocal][2] = 0.0;\n"
            "  atom->mu_induced[nlocal][0] = atom->mu_induced[nlocal][1] = atom->mu_induced[nlocal][2] = 0.0;\n")
    t = patch_func(t, "void AtomVecFull::create_atom(int itype, double *coord)",
                   "  q[nlocal] = 0.0;\n", "  q[nlocal] = 0.0;\n" + zero, 1)
    t = patch_func(t, "void AtomVecFull::data_atom(double *coord, imageint imagetmp, char **values)",
                   "  q[nlocal] = atof(values[3]);\n", "  q[nlocal] = atof(values[3]);\n" + zero, 1)
    path.write_text(t)


def patch_pair(path: Path):
    t = path.read_text()
    t = sub_once(t, '#include "domain.h"\n', '#include "domain.h"\n#include "polb200_dump.h"\n',
                 what="include")
    t = sub_once(t, "  double ef_temp;\n  /* end polarization stuff */\n",
                 "  double ef_temp;\n  polb200_dump::pre(atom);\n  /* end polarization stuff */\n",
                 what="pre hook")
    t = sub_once(t, "  if (vflag_fdotr) virial_fdotr_compute();\n}\n",
                 "  if (vflag_fdotr) virial_fdotr_compute();\n"
                 "  polb200_dump::post(atom,domain,force,list,eflag,vflag,iterations,eng_vdwl,eng_coul,\n"
                 "                     eng_pol,virial,g_ewald,cut_coul,tabinnersq,ncoultablebits,ncoulmask,\n"
                 "                     ncoulshiftbits,rtable,drtable,ftable,dftable,ctable,dctable,etable,\n"
                 "                     detable,cutsq,cut_ljsq,lj1,lj2,lj3,lj4,offset,neighbor->ago);\n}\n",
                 what="post hook")
    path.write_text(t)


def main():
    src = Path(sys.argv[1])
    if str(src.resolve()).startswith("/root/reference"):
        raise SystemExit("refusing to patch the read-only reference tree")
    patch_atom_vec_full(src / "atom_vec_full.cpp")
    if not os.environ.get("POLB200_PATCH_ATOMVEC_ONLY"):  # the drop-in build replaces the pair style instead
        patch_pair(src / "pair_lj_cut_coul_long_polarization.cpp")
    print("patched", src)


if __name__ == "__main__":
    main()
