#!/usr/bin/env python3
"""Time the reference binary and the drop-in binary on the same LAMMPS script (Bulk H2 example rebuilt from the
golden fixture, the reference's default keywords): Pair time per step from the LAMMPS timing breakdown."""
import re
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import polhelpers as H
import test_lammps_dropin as T

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for case, words in (("h2_default_step0", "precision 0.00000000001 max_iterations 100 damp_type exponential damp 2.1304 polar_gs_ranked yes debug no use_previous yes"),
                    ):
    fx = H.load_fixture(case)
    style = str(fx["pair_style"]) if words is None else "pair_style lj/cut/coul/long/polarization 2.5 10.797442 " + words
    work = Path(tempfile.mkdtemp())
    T.write_case(work, fx, style, steps)
    for name, binary in (("reference", T.LMP_REF), ("b200", T.LMP_B200)):
        tab, out = T.run(binary, work, name)
        log = (work / f"log.{name}").read_text()
        pair = float(re.search(r"^Pair\s*\|\s*([0-9.eE+-]+)", log, flags=re.M).group(1))
        loop = float(re.search(r"Loop time of ([0-9.eE+-]+)", log).group(1))
        print(f"{case} n={fx['x'].shape[0]} {name}: Pair {pair / steps * 1e3:.2f} ms/step, loop {loop / steps * 1e3:.2f} ms/step, E_pol(last) {tab[-1][5]:.8f}")

# the example AS SHIPPED (fix rigid/nve molecule, its masses and velocities): in lmp_b200 pair style, KSpace and the integrator
# all run on the device; LAMMPS' own timing breakdown per step
import lammps_cases as LC

work = Path(tempfile.mkdtemp())
fx = LC.write_h2_data(work)
(work / "in.case").write_text("\n".join(LC.h2_shipped_lines(fx) + LC.H2_DYNAMICS + [f"run {steps}"]) + "\n")
for name, binary in (("reference", T.LMP_REF), ("b200", T.LMP_B200)):
    cols, rows = LC.run_log(binary, work, name)
    log = (work / f"log.{name}").read_text()
    loop = float(re.search(r"Loop time of ([0-9.eE+-]+)", log).group(1))
    parts = {k: float(re.search(rf"^{k}\s*\|\s*([0-9.eE+-]+)", log, flags=re.M).group(1)) / steps * 1e3
             for k in ("Pair", "Kspace", "Neigh", "Comm", "Modify")}
    print(f"shipped h2.input ({fx['x'].shape[0]} atoms, rigid/nve molecule) {name}: loop {loop / steps * 1e3:.2f} ms/step  "
          + "  ".join(f"{k} {v:.3f}" for k, v in parts.items()) + f"  TotEng(last) {rows[-1][cols.index('TotEng')]:.6f}")
