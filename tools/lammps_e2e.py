#!/usr/bin/env python3
"""End-to-end LAMMPS run of the bench workload through the drop-in binary (pair style + kspace_style ewald on the
GPU, everything else stock LAMMPS on one host core): the timing breakdown LAMMPS prints."""
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import bench

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
sysm = bench.workloads().lj_charge_fluid(ncell)
work = Path(tempfile.mkdtemp(prefix="polb200_e2e_"))
bench.write_lammps_case(work, sysm, None, steps)
text = (work / "in.fluid").read_text().replace(bench.STYLE_WORDS, bench.STYLE_WORDS + f" polar_cutoff {bench.CUT_COUL}")
text = text.replace("thermo 1\n", "thermo 10\n")
(work / "in.fluid").write_text(text)
lmp = ROOT / "lammps-induced-dipole-polarization-pair-style_b200" / "lammps" / "_build" / "lmp_b200"
r = subprocess.run([str(lmp), "-in", "in.fluid", "-echo", "none"], cwd=work, capture_output=True, text=True, timeout=900)
if r.returncode != 0:
    print(r.stdout[-3000:], r.stderr[-2000:])
    raise SystemExit(1)
log = (work / "log.lammps").read_text()
loop = float(re.search(r"Loop time of ([0-9.eE+-]+)", log).group(1))
print(f"lmp_b200, {sysm.n} atoms, {steps} steps: {loop / steps * 1e3:.2f} ms per step = {sysm.n * steps / loop:.3g} atom-steps/s (whole LAMMPS step)")
for name in ("Pair", "Kspace", "Neigh", "Comm", "Output", "Modify", "Other"):
    m = re.search(rf"^{name}\s*\|\s*([0-9.eE+-]*)\s*\|\s*([0-9.eE+-]+)\s*\|\s*([0-9.eE+-]*)\s*\|\s*[0-9.]*\s*\|\s*([0-9.]+)", log, flags=re.M)
    if m:
        print(f"  {name:7s} {float(m.group(2)) / steps * 1e3:8.3f} ms/step  {m.group(4)} %")
print("\n".join(l for l in log.splitlines() if l.startswith("Step") or re.match(r"^\s+\d+\s+-", l))[:600])
