#!/usr/bin/env python3
"""Golden fixtures for the LIBRARY-SIDE rebuild schedule (callers that pass ago < 0): tests/golden/ago_*.npz.

Runs HERE with the repaired reference binary (oracle/_ref/lmp_serial): a small hot LJ+charge fluid under `fix nve`, so that
`Neighbor::decide` / `check_distance` (src/neighbor.cpp:1923-2001) trigger rebuilds at irregular steps.  Per step the hooks of
oracle/ref_shims/polb200_dump.h give the positions the pair style saw and `neighbor->ago`; the fixture keeps the positions
of every step, ago of every step, and dipoles / forces of the first and the last step.

  python oracle/make_golden_ago.py
"""
import os
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle import polref as P  # noqa: E402
import polhelpers as H  # noqa: E402

LMP = ROOT / "oracle" / "_ref" / "lmp_serial"
OUT = ROOT / "tests" / "golden"
STYLE = "pair_style lj/cut/coul/long/polarization 2.5 6.0 polar_gs_ranked no fixed_iteration yes max_iterations 2 damp_type exponential"
CASES = {"ago_every1": "neigh_modify delay 0 every 1 check yes", "ago_delay4_every2": "neigh_modify delay 4 every 2 check yes",
         "ago_nocheck": "neigh_modify delay 0 every 3 check no"}
NSTEP = 40


def main():
    w = H._workloads().lj_charge_fluid(4, seed=77, rho=0.05)
    n = w.n
    L = float(w.boxhi[0])
    for name, neigh in CASES.items():
        work = Path(tempfile.mkdtemp(prefix=f"polgold_{name}_"))
        with open(work / "fluid.data", "w") as fh:
            fh.write("hot polarizable LJ+charge fluid\n\n%d atoms\n2 atom types\n\n" % n)
            fh.write("0.0 %.16g xlo xhi\n0.0 %.16g ylo yhi\n0.0 %.16g zlo zhi\n\nAtoms\n\n" % (L, L, L))
            for i in range(n):
                fh.write("%d 0 %d %.16g %.16g %.16g %.16g\n" % (i + 1, w.type[i], w.q[i], *w.x[i]))
        (work / "in.case").write_text(f"""units real
boundary p p p
atom_style full
atom_modify sort 0 0.0
read_data fluid.data
mass * 12.0
set type 1 static_polarizability 1.0
set type 2 static_polarizability 0.5
kspace_style ewald 1.0e-4
{STYLE}
pair_coeff 1 1 0.1 3.0
pair_coeff 2 2 0.1 3.0
{neigh}
velocity all create 6000.0 4928 loop geom
fix 1 all nve
timestep 2.0
thermo 10
run {NSTEP}
""")
        env = dict(os.environ, POLB200_DUMP=str(work / "dump"))
        r = subprocess.run([str(LMP), "-in", "in.case", "-echo", "none"], cwd=work, env=env, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout[-3000:])
            raise SystemExit(f"{name}: lmp_serial failed")
        g = None
        xs, agos, keep = [], [], {}
        for step in range(NSTEP + 1):
            d = P.read_refdump(work / f"dump.{step}.bin")
            nl = int(d["nlocal"][0])
            assert nl == n and np.array_equal(d["tag"][:nl], np.arange(1, n + 1))  # single rank, no sorting: caller order = tag order
            xs.append(d["x"].reshape(-1, 3)[:nl].copy())
            agos.append(int(d["neighbor_ago"][0]))
            g = float(d["g_ewald"][0])   # the exact double (the log prints six digits)
            if step in (0, NSTEP):
                tag = d["tag"]
                f_all = d["f"].reshape(-1, 3)
                f_own = f_all[:nl].copy()
                np.add.at(f_own, tag[nl:] - 1, f_all[nl:])   # ghost forces folded onto their owners (reverse_comm)
                keep[step] = dict(mu=d["mu_out"].reshape(-1, 3).copy(), f=f_own, eng_pol=float(d["eng_pol"][0]),
                                  eng_coul=float(d["eng_coul"][0]))
        agos = np.array(agos, dtype=np.int32)
        assert agos[0] == 0 and (agos[1:] == 0).sum() >= 2, agos
        np.savez_compressed(OUT / f"{name}.npz", x=np.array(xs), ago=agos, q=w.q, type=w.type, alpha=w.alpha, boxlo=w.boxlo,
                            boxhi=w.boxhi, g_ewald=g, pair_style=STYLE, neigh_modify=neigh, skin=2.0,
                            mu_first=keep[0]["mu"], f_first=keep[0]["f"], mu_last=keep[NSTEP]["mu"], f_last=keep[NSTEP]["f"],
                            eng_pol_last=keep[NSTEP]["eng_pol"], eng_coul_last=keep[NSTEP]["eng_coul"])
        print(name, "rebuild steps:", np.nonzero(agos == 0)[0].tolist())
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
