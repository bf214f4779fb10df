"""Drop-in test (gpu marker): the SAME LAMMPS input script is run by
  * oracle/_ref/lmp_serial            -- the reference (its own CPU compute()), and
  * .../lammps/_build/lmp_b200        -- the reference's host framework with the pair style replaced by
                                         pair_lj_cut_coul_long_polarization_b200.cpp -> C ABI -> CUDA,
and the thermo tables (E_vdwl, E_coul, E_long, E_pol, PotEng, Press) must agree step by step.

The system is the reference's Bulk H2 example (750 atoms, 150 rigid 5-site molecules) rebuilt from the
committed golden fixture (positions, charges, types, molecule ids, bond topology from the special lists),
so nothing is read from /root/reference at run time.  Both binaries are built by __graft_entry__.build()
in the container that has the reference tree and travel to the GPU box.
"""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import polhelpers as H

ROOT = Path(__file__).resolve().parents[1]
LMP_REF = ROOT / "oracle" / "_ref" / "lmp_serial"
LMP_B200 = ROOT / "lammps-induced-dipole-polarization-pair-style_b200" / "lammps" / "_build" / "lmp_b200"

pytestmark = pytest.mark.gpu


def write_case(work, fx, style_words, steps, extra=()):
    n = fx["x"].shape[0]
    tag, mol, typ = fx["tag"], fx["molecule"], fx["type"]
    nsp, sp = fx["nspecial"], fx["special"]
    bonds = set()
    for i in range(n):
        for k in range(int(nsp[i, 0])):  # 1-2 partners
            a, b = int(tag[i]), int(sp[i, k])
            bonds.add((min(a, b), max(a, b)))
    bonds = sorted(bonds)
    with open(work / "sys.data", "w") as fh:
        fh.write(f"Bulk H2 from golden fixture\n\n{n} atoms\n{int(fx['ntypes'])} atom types\n{len(bonds)} bonds\n1 bond types\n\n")
        for d, c in enumerate("xyz"):
            fh.write(f"{float(fx['boxlo'][d]):.17g} {float(fx['boxhi'][d]):.17g} {c}lo {c}hi\n")
        fh.write("\nAtoms\n\n")
        for i in range(n):
            fh.write(f"{int(tag[i])} {int(mol[i])} {int(typ[i])} {float(fx['q'][i]):.17g} "
                     f"{fx['x'][i, 0]:.17g} {fx['x'][i, 1]:.17g} {fx['x'][i, 2]:.17g}\n")
        fh.write("\nBonds\n\n")
        for k, (a, b) in enumerate(bonds):
            fh.write(f"{k + 1} 1 {a} {b}\n")
    alpha_of_type = {int(t): float(np.unique(fx["alpha"][typ == t])[0]) for t in np.unique(typ)}
    lines = ["units real", "boundary p p p", "atom_style full", "read_data sys.data", "mass * 1.0",
             "bond_style zero", "bond_coeff *"]
    lines += [f"set type {t} static_polarizability {a:.17g}" for t, a in alpha_of_type.items()]
    lines += ["kspace_style ewald 1.0e-4", style_words]
    lines += str(fx["pair_coeff"]).splitlines()
    lines += ["special_bonds lj/coul 0.0 0.0 0.0"] + list(extra) + [
              "thermo_style custom step pe evdwl ecoul elong epol press",
              "thermo_modify format float %.12g", "thermo 1", "timestep 0.25", "fix 1 all nve", f"run {steps}"]
    (work / "in.case").write_text("\n".join(lines) + "\n")


def thermo_table(log):
    rows, on = [], False
    for line in log.splitlines():
        if line.startswith("Step "):
            on = True
            continue
        if on:
            if line.startswith("Loop time"):
                break
            t = line.split()
            if len(t) == 7 and re.fullmatch(r"\d+", t[0]):
                rows.append([float(v) for v in t])
    return np.array(rows)


def run(binary, work, name):
    r = subprocess.run([str(binary), "-in", "in.case", "-echo", "none", "-log", f"log.{name}"], cwd=work,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return thermo_table((work / f"log.{name}").read_text()), r.stdout


EXCLUDE = ("group gsite type 1", "group esite type 2", "neigh_modify exclude molecule/intra all exclude type 1 3",
           "neigh_modify exclude group gsite esite")


@pytest.mark.parametrize("words,tol,extra", [
    ("polar_gs_ranked no fixed_iteration yes max_iterations 20 damp_type exponential damp 2.1304", 1e-9, ()),
    ("precision 0.00000000001 max_iterations 100 damp_type exponential damp 2.1304 polar_gs_ranked yes debug no "
     "use_previous yes", 2e-8, ()),
    # neigh_modify exclude in the script: the Pair subclass forwards Neighbor's rules to the device list
    ("polar_gs_ranked no fixed_iteration yes max_iterations 20 damp_type exponential damp 2.1304", 1e-9, EXCLUDE),
])
def test_same_script_same_thermo(tmp_path, words, tol, extra):
    if not LMP_REF.exists() or not LMP_B200.exists():
        pytest.skip("LAMMPS binaries not built (need the reference tree at build time)")
    fx = H.load_fixture("h2_default_step0")
    style = "pair_style lj/cut/coul/long/polarization 2.5 10.797442 " + words
    write_case(tmp_path, fx, style, steps=4, extra=extra)
    ref, _ = run(LMP_REF, tmp_path, "ref")
    new, out = run(LMP_B200, tmp_path, "b200")
    assert ref.shape == new.shape and ref.shape[0] == 5
    for col, name in enumerate(["step", "pe", "evdwl", "ecoul", "elong", "epol", "press"]):
        scale = max(np.abs(ref[:, col]).max(), 1.0)
        assert np.abs(ref[:, col] - new[:, col]).max() <= tol * scale, (name, ref[:, col], new[:, col])


# ---- fix rigid/nve|nvt on the device -------------------------------------------------------

from lammps_cases import H2_DYNAMICS, check_against_shipped_log, h2_shipped_lines, run_log, write_h2_data  # noqa: E402


@pytest.mark.parametrize("fix_line,scalar,kspace", [
    ("fix rig all rigid/nve molecule", False, "ewald 1.0e-5"),
    ("fix rig all rigid/nvt molecule temp 298.15 250.0 100.0 tparam 50 1 3", True, "ewald 1.0e-5"),
    # kspace_style pppm (+ kspace_modify) on the device: polb200_pppm_* behind lammps/pppm_b200.{h,cpp}
    ("fix rig all rigid/nve molecule", False, "pppm 1.0e-5"),
    ("fix rig all rigid/nve molecule", False, "pppm 1.0e-4\nkspace_modify order 4 mesh 16 15 18 gewald 0.45"),
])
def test_rigid_fix_same_script_same_thermo(tmp_path, fix_line, scalar, kspace):
    """a rigid water box under the stock lj/cut/coul/long pair style: in lmp_b200 the fix (and Ewald) run on the GPU,
    in the reference binary on the host; temperature, energies, pressure and the fix's scalar agree step by step"""
    if not LMP_REF.exists() or not LMP_B200.exists():
        pytest.skip("LAMMPS binaries not built (need the reference tree at build time)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_rigid", ROOT / "oracle" / "make_golden_rigid.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    data, _ = mg.water_data(4)
    (tmp_path / "water.data").write_text(data)
    text = mg.water_input(fix_line, 8, 1.0).replace("kspace_style ewald 1.0e-5", "kspace_style " + kspace)
    assert "kspace_style " + kspace in text
    th = "thermo_style custom step temp ke pe elong press" + (" f_rig" if scalar else "")
    text += "\n".join([th, "thermo_modify format float %.14g", "thermo 1", "run 8"]) + "\n"
    (tmp_path / "in.case").write_text(text)
    cols_r, ref = run_log(LMP_REF, tmp_path, "ref")
    cols_n, new = run_log(LMP_B200, tmp_path, "b200")
    assert cols_r == cols_n and ref.shape == new.shape and ref.shape[0] == 9
    for c, name in enumerate(cols_r):
        scale = max(np.abs(ref[:, c]).max(), 1.0)
        assert np.abs(ref[:, c] - new[:, c]).max() <= 1e-8 * scale, (name, ref[:, c], new[:, c])


def test_shipped_h2_example_reproduces_the_committed_log(tmp_path):
    """The reference's Bulk H2 example AS SHIPPED -- polarization pair style, Ewald and `fix rigid/nve molecule` all
    replaced by their device counterparts in lmp_b200 -- against the thermo table of the reference's own committed
    log (polarization/examples/Bulk H2/log.lammps:92-100, 8 printed digits; tests/golden/thermo_logs.json)."""
    if not LMP_B200.exists():
        pytest.skip("lmp_b200 not built (needs the reference tree at build time)")
    fx = write_h2_data(tmp_path)
    (tmp_path / "in.case").write_text("\n".join(h2_shipped_lines(fx) + H2_DYNAMICS + ["run 7"]) + "\n")
    cols, new = run_log(LMP_B200, tmp_path, "b200")
    assert new.shape[0] == 8
    check_against_shipped_log(cols, new)


# ---- the committed atom style inside lmp_b200 ----------------------------------------------

def test_atom_style_cases_through_lmp_b200(tmp_path, monkeypatch):
    """tests/test_atom_style.py's cases (atom sorting, restart round trip without `set`, replicate, lammps_extract_atom) with every style on
    the device: the same atom style is built into lmp_b200, where pair style, Ewald and fix rigid read and write the
    arrays through the C ABI"""
    if not LMP_B200.exists():
        pytest.skip("lmp_b200 not built (needs the reference tree at build time)")
    import test_atom_style as TA
    monkeypatch.setattr(TA, "LMP_AV", LMP_B200)
    for k, case in enumerate([TA.test_atom_sorting_carries_the_arrays, TA.test_restart_round_trip_keeps_polarizabilities_and_dipoles,
                              TA.test_replicate_goes_through_the_restart_records,
                              TA.test_library_interface_extracts_the_arrays, TA.test_exchange_record_round_trips_every_atom]):
        work = tmp_path / f"case{k}"
        work.mkdir()
        case(work)
