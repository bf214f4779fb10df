#!/usr/bin/env python3
"""BASELINE config 3 as an unchanged LAMMPS script through the drop-in binary: rigid polarizable water box (3*W^3 atoms),
`fix rigid/nve molecule`, `kspace_style pppm`, the reference's default solver keywords -- pair style, KSpace and integrator on
the GPU, the stock Verlet loop on one host core.  Runs twice: device-resident atoms (default) and POLB200_RESIDENT=0 (every
style copies x / v / f through the host every step); prints LAMMPS' own timing breakdown and checks that both runs print
the same thermo.

  python tools/lammps_water.py [W=44] [steps=20]
"""
import os
import re
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench

W = int(sys.argv[1]) if len(sys.argv) > 1 else 44
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
s = bench.workloads().water_box(W)
n, L = s.n, float(s.boxhi[0])
work = Path(tempfile.mkdtemp(prefix="polb200_water_"))
rng = np.random.default_rng(5)
with open(work / "water.data", "w") as fh:
    fh.write(f"rigid polarizable water box\n\n{n} atoms\n2 atom types\n\n0.0 {L:.16g} xlo xhi\n0.0 {L:.16g} ylo yhi\n0.0 {L:.16g} zlo zhi\n\n")
    fh.write("Masses\n\n1 15.9994\n2 1.008\n\nAtoms\n\n")
    for i in range(n):
        fh.write("%d %d %d %.16g %.16g %.16g %.16g\n" % (i + 1, s.molecule[i], s.type[i], s.q[i], *s.x[i]))
(work / "in.water").write_text(f"""units real
boundary p p p
atom_style full
read_data water.data
set type 1 static_polarizability 0.837
set type 2 static_polarizability 0.496
kspace_style pppm 1.0e-4
pair_style lj/cut/coul/long/polarization 2.5 12.0 precision 1e-11 max_iterations 200 polar_gamma 1.03 damp_type exponential use_previous yes polar_cutoff 12.0
pair_coeff 1 1 0.155 3.166
pair_coeff 2 2 0.0 1.0
velocity all create 50.0 4928 loop geom
fix 1 all rigid/nve molecule
timestep 0.1
thermo_style custom step temp pe ke etotal epol press
thermo 10
run {steps}
""")
lmp = ROOT / "lammps-induced-dipole-polarization-pair-style_b200" / "lammps" / "_build" / "lmp_b200"
tables = {}
for label, env in (("host buffers (POLB200_RESIDENT=0), first run", {"POLB200_RESIDENT": "0"}), ("device-resident atoms", {}),
                   ("host buffers (POLB200_RESIDENT=0)", {"POLB200_RESIDENT": "0"})):
    r = subprocess.run([str(lmp), "-in", "in.water", "-echo", "none", "-log", f"log.{len(tables)}"], cwd=work, capture_output=True,
                       text=True, timeout=1800, env=dict(os.environ, **env))
    if r.returncode != 0:
        print(r.stdout[-3000:], r.stderr[-2000:])
        raise SystemExit(1)
    log = (work / f"log.{len(tables)}").read_text()
    loop = float(re.search(r"Loop time of ([0-9.eE+-]+)", log).group(1))
    print(f"lmp_b200, {n} atoms, {steps} steps, {label}: {loop / steps * 1e3:.2f} ms per step = {n * steps / loop:.3g} atom-steps/s")
    for name in ("Pair", "Kspace", "Neigh", "Comm", "Output", "Modify", "Other"):
        m = re.search(rf"^{name}\s*\|\s*([0-9.eE+-]*)\s*\|\s*([0-9.eE+-]+)\s*\|\s*([0-9.eE+-]*)\s*\|\s*[0-9.]*\s*\|\s*([0-9.]+)", log, flags=re.M)
        if m:
            print(f"  {name:7s} {float(m.group(2)) / steps * 1e3:8.3f} ms/step  {m.group(4)} %")
    rows = [l.split() for l in log.splitlines() if re.match(r"^\s+\d+\s+-?[0-9.]", l)]
    tables[label] = np.array([[float(v) for v in row] for row in rows])
    print("\n".join(l for l in log.splitlines() if l.startswith("Step") or re.match(r"^\s+\d+\s+-?[0-9.]", l))[:800])
_, a, b = list(tables.values())
err = np.abs(a - b).max() / np.abs(b).max()
print(f"thermo tables of the two runs agree to {err:.2e} (relative to the largest entry)")
assert a.shape == b.shape and err < 1e-9
