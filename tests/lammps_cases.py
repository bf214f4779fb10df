"""LAMMPS input cases shared by the drop-in tests (GPU: lmp_b200) and the atom-style tests (CPU: lmp_serial_av):
the reference's Bulk H2 example rebuilt from the committed golden fixture, so nothing is read from /root/reference at
run time."""
import subprocess

import numpy as np

import polhelpers as H

H2_MASS = ["mass 1 0.00001", "mass 2 1.00800", "mass 3 0.00001"]
H2_ALPHA = ["set type 1 static_polarizability 0.69380", "set type 2 static_polarizability 0.00044",
            "set type 3 static_polarizability 0.00000"]
H2_THERMO = "thermo_style custom step etotal ke pe evdwl ecoul elong epol temp press"


def write_h2_data(work, name="h2.data"):
    """h2.data of the shipped example from fixture h2_default_step0: every molecule unwrapped around its first atom
    (read_data wraps it again and sets the image flags fix rigid needs), bonds from the 1-2 special lists"""
    fx = H.load_fixture("h2_default_step0")
    n = fx["x"].shape[0]
    tag, mol, typ = fx["tag"], fx["molecule"], fx["type"]
    L = fx["boxhi"] - fx["boxlo"]
    order = np.argsort(tag)
    x = fx["x"].copy()
    first = {}
    for i in order:
        m = int(mol[i])
        if m not in first:
            first[m] = x[i].copy()
        x[i] = first[m] + (x[i] - first[m]) - L * np.rint((x[i] - first[m]) / L)
    nsp, sp = fx["nspecial"], fx["special"]
    bonds = sorted({(min(int(tag[i]), int(sp[i, k])), max(int(tag[i]), int(sp[i, k]))) for i in range(n)
                    for k in range(int(nsp[i, 0]))})
    with open(work / name, "w") as fh:
        fh.write(f"Bulk H2 from golden fixture\n\n{n} atoms\n3 atom types\n{len(bonds)} bonds\n1 bond types\n\n")
        for d, c in enumerate("xyz"):
            fh.write(f"{float(fx['boxlo'][d]):.17g} {float(fx['boxhi'][d]):.17g} {c}lo {c}hi\n")
        fh.write("\nAtoms\n\n")
        for i in order:
            fh.write(f"{int(tag[i])} {int(mol[i])} {int(typ[i])} {float(fx['q'][i]):.17g} {x[i, 0]:.17g} {x[i, 1]:.17g} {x[i, 2]:.17g}\n")
        fh.write("\nBonds\n\n")
        for k, (a, b) in enumerate(bonds):
            fh.write(f"{k + 1} 1 {a} {b}\n")
    return fx


def h2_shipped_lines(fx, pair_style=None):
    """the shipped h2.input up to (not including) velocity / fix / run, line for line where it matters
    (its `timestep 2` precedes `units real`, so dt = 1 fs)"""
    lines = ["timestep 2", "units real", "boundary p p p", "atom_style full", "read_data h2.data", "bond_style zero",
             "bond_coeff *"] + H2_MASS + H2_ALPHA
    lines += ["kspace_style ewald 1.0e-4", pair_style or str(fx["pair_style"])]
    lines += str(fx["pair_coeff"]).splitlines()
    lines += ["special_bonds lj/coul 0.0 0.0 0.0", H2_THERMO, "thermo 1"]
    return lines


H2_DYNAMICS = ["velocity all create 298.15 12345 rot yes mom yes dist gaussian", "fix rigid_nve all rigid/nve molecule"]


def thermo_rows(log):
    """last thermo table of a log as (columns, float rows)"""
    cols, rows, on = None, [], False
    for line in log.splitlines():
        if line.startswith("Step "):
            cols, rows, on = line.split(), [], True
            continue
        if on:
            if line.startswith("Loop time"):
                on = False
                continue
            t = line.split()
            if len(t) == len(cols):
                try:
                    rows.append([float(v) for v in t])
                except ValueError:
                    pass
    return cols, np.array(rows)


def run_log(binary, work, name, infile="in.case"):
    r = subprocess.run([str(binary), "-in", infile, "-echo", "none", "-log", f"log.{name}"], cwd=work,
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return thermo_rows((work / f"log.{name}").read_text())


def check_against_shipped_log(cols, rows, nrows=8, tol=3e-7):
    """the reference's committed log, polarization/examples/Bulk H2/log.lammps:92-100 (8 printed digits; it ends inside
    step 8, so 8 complete rows)"""
    ref = H.thermo_logs()["h2"]
    assert cols == ref["columns"] and rows.shape[0] >= nrows and len(ref["rows"]) >= nrows
    for r in range(nrows):
        for c, name in enumerate(cols):
            want = float(ref["rows"][r][name])
            assert abs(rows[r, c] - want) <= tol * max(abs(want), 1.0), (r, name, rows[r, c], want)
