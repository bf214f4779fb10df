#!/usr/bin/env python3
"""SURVEY §8(d) "CPU reference timing": Pair time per step of the reference binary (oracle/_ref/lmp_serial, one core --
the reference is serial by design) on the BASELINE config-2 fluid at growing N, with the fitted c*N^2 law and its
extrapolation to the bench size (32 000 atoms, where the reference's 3N x 3N matrix alone needs 74 GB).
Runs on the host CPU only; prints one JSON line.  usage: ref_cpu_scaling.py [ncell ...] (default 8 10 12 14)"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench

LMP = ROOT / "oracle" / "_ref" / "lmp_serial"
cells = [int(a) for a in sys.argv[1:]] or [8, 10, 12, 14]
rows = []
for nc in cells:
    sysm = bench.workloads().lj_charge_fluid(nc)
    work = Path(tempfile.mkdtemp(prefix="polb200_refscale_"))
    steps = 2
    bench.write_lammps_case(work, sysm, None, steps)
    t0 = time.time()
    r = subprocess.run([str(LMP), "-in", "in.fluid", "-echo", "none"], cwd=work, capture_output=True, text=True)
    wall = time.time() - t0
    if r.returncode != 0:
        print(r.stdout[-500:], file=sys.stderr)
        break
    log = (work / "log.lammps").read_text()
    pair = float(re.search(r"^Pair\s*\|\s*([0-9.eE+-]+)", log, flags=re.M).group(1)) / steps
    mem = re.search(r"Per MPI rank memory allocation \(min/avg/max\) = (\S+)", log)
    rows.append(dict(atoms=sysm.n, pair_s_per_step=pair, atom_steps_per_s=sysm.n / pair, wall_s=wall,
                     matrix_gb=72.0 * sysm.n ** 2 / 1e9, lammps_mbytes=float(mem.group(1)) if mem else None))
    print(f"N={sysm.n}: Pair {pair:.3f} s/step ({sysm.n / pair:.0f} atom-steps/s), 3Nx3N matrix {rows[-1]['matrix_gb']:.2f} GB", file=sys.stderr)
    shutil.rmtree(work, ignore_errors=True)
n = np.array([r["atoms"] for r in rows], dtype=float)
t = np.array([r["pair_s_per_step"] for r in rows])
c = float((t * n ** 2).sum() / (n ** 4).sum())      # least squares of t = c*N^2
out = dict(what="reference CPU scaling (lmp_serial, 1 core)", cpu=os.uname().nodename, cores_used=1, rows=rows, fit_c_s_per_atom2=c,
           fit_rel_residuals=[float(ti / (c * ni * ni) - 1.0) for ti, ni in zip(t, n)],
           extrapolated_32000=dict(pair_s_per_step=c * 32000.0 ** 2, atom_steps_per_s=32000.0 / (c * 32000.0 ** 2), matrix_gb=72.0 * 32000.0 ** 2 / 1e9,
                                   note="EXTRAPOLATED from the c*N^2 fit; the reference cannot run this size on a 62 GB host"))
print(json.dumps(out))
