"""The committed atom style: lammps/atom_vec_full_polar_b200.{h,cpp} = `atom_style full` + the
three per-atom arrays of the polarization pair style, and compute polarization/atom.  Host C++ only, so these tests run
on the CPU: oracle/_ref/lmp_serial_av (oracle/build_ref_av.sh) is the reference -- its own CPU pair style untouched --
with OUR atom style in place of the one the fork never shipped.  The yardsticks are the reference's committed log and
golden vectors dumped from the reference binary."""
from pathlib import Path

import numpy as np
import pytest

import polhelpers as H
from lammps_cases import H2_DYNAMICS, check_against_shipped_log, h2_shipped_lines, run_log, write_h2_data

LMP_AV = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "lmp_serial_av"

pytestmark = pytest.mark.skipif(not LMP_AV.exists(), reason="oracle/_ref/lmp_serial_av not built (needs the reference tree)")

FIXED3 = ("pair_style lj/cut/coul/long/polarization 2.5 10.797442 max_iterations 3 fixed_iteration yes damp_type exponential "
          "damp 2.1304 polar_gs_ranked no use_previous yes")


def test_shipped_example_runs_on_the_committed_atom_style(tmp_path):
    """as shipped the reference aborts in init_style ("requires atom attribute polarizability"); with the committed atom
    style the unchanged example reproduces the reference's committed log"""
    fx = write_h2_data(tmp_path)
    (tmp_path / "in.case").write_text("\n".join(h2_shipped_lines(fx) + H2_DYNAMICS + ["run 7"]) + "\n")
    cols, rows = run_log(LMP_AV, tmp_path, "av")
    check_against_shipped_log(cols, rows)


def test_atom_sorting_carries_the_arrays(tmp_path):
    """atom_modify sort every step (AtomVec::copy on every atom): polarizabilities and previous dipoles must move with
    their atoms -- same log"""
    fx = write_h2_data(tmp_path)
    lines = h2_shipped_lines(fx) + ["atom_modify sort 1 3.0"] + H2_DYNAMICS + ["run 7"]
    (tmp_path / "in.case").write_text("\n".join(lines) + "\n")
    cols, rows = run_log(LMP_AV, tmp_path, "sorted")
    check_against_shipped_log(cols, rows)


def test_restart_round_trip_keeps_polarizabilities_and_dipoles(tmp_path):
    """write_restart after 2 steps, read_restart WITHOUT any `set ... static_polarizability`, 2 more steps: the rows of
    steps 2-4 equal those of an uninterrupted 4-step run.  The pair style runs 3 fixed Jacobi iterations from the
    previous dipoles (use_previous yes), so E_pol after the restart depends on the restored mu_induced at the 1e-4
    level -- a lost array cannot hide behind SCF convergence."""
    fx = write_h2_data(tmp_path)
    # (the example's 1e-5 amu sites only make sense inside rigid bodies: keep the shipped integrator)
    head = h2_shipped_lines(fx, pair_style=FIXED3) + ["thermo_modify format float %.14g"] + H2_DYNAMICS
    (tmp_path / "in.full").write_text("\n".join(head + ["run 4"]) + "\n")
    (tmp_path / "in.first").write_text("\n".join(head + ["run 2", "write_restart half.restart"]) + "\n")
    second = ["read_restart half.restart", "bond_style zero", "bond_coeff *", "kspace_style ewald 1.0e-4", FIXED3]
    second += str(fx["pair_coeff"]).splitlines()
    second += ["special_bonds lj/coul 0.0 0.0 0.0", "thermo_style custom step etotal ke pe evdwl ecoul elong epol temp press",
               "thermo 1", "thermo_modify format float %.14g", H2_DYNAMICS[1], "run 2"]
    (tmp_path / "in.second").write_text("\n".join(second) + "\n")
    cols, full = run_log(LMP_AV, tmp_path, "full", "in.full")
    run_log(LMP_AV, tmp_path, "first", "in.first")
    cols2, cont = run_log(LMP_AV, tmp_path, "second", "in.second")
    assert cols == cols2 and full.shape[0] == 5 and cont.shape[0] == 3
    # energies and temperature to 2e-9 of the largest energy term (E_coul and E_long cancel to 4 digits).  Press is left
    # out: the reference's fix rigid restarts with its "2x set_v" virial guess (fix_rigid.cpp:882-888), not the real one
    scale = np.abs(full[:, 1:8]).max()
    for c, name in enumerate(cols):
        if name != "Press":
            assert np.abs(full[2:, c] - cont[:, c]).max() <= 2e-9 * scale, (name, full[2:, c], cont[:, c])
    assert abs(full[2, cols.index("E_pol")] - cont[0, cols.index("E_pol")]) < 1e-9 * abs(full[2, cols.index("E_pol")])
    # and the control: zeroing the dipoles before the continuation DOES change E_pol (the test can see a lost array)
    ctrl = second[:-1] + ["set group all static_polarizability 0.0", "run 0"]
    (tmp_path / "in.ctrl").write_text("\n".join(ctrl) + "\n")
    _, zero = run_log(LMP_AV, tmp_path, "ctrl", "in.ctrl")
    assert abs(zero[0, cols.index("E_pol")]) < 1e-12 and abs(cont[0, cols.index("E_pol")]) > 1e-3


def test_compute_polarization_atom_exposes_the_arrays(tmp_path):
    """dump custom ... c_pol[*] at step 0 of the shipped example against the per-atom arrays dumped from the reference
    binary (tests/golden/h2_default_step0.npz)"""
    fx = write_h2_data(tmp_path)
    lines = h2_shipped_lines(fx) + ["compute pol all polarization/atom",
                                     "dump d all custom 1 pol.dump id c_pol[1] c_pol[2] c_pol[3] c_pol[4] c_pol[5] c_pol[6] c_pol[7]",
                                     "dump_modify d sort id format float %.17g"] + H2_DYNAMICS + ["run 0"]
    (tmp_path / "in.case").write_text("\n".join(lines) + "\n")
    run_log(LMP_AV, tmp_path, "pol")
    rows = np.array([[float(t) for t in l.split()] for l in (tmp_path / "pol.dump").read_text().splitlines()[9:]])
    order = np.argsort(fx["tag"])
    assert np.array_equal(rows[:, 0].astype(int), fx["tag"][order])
    assert np.abs(rows[:, 1] - fx["alpha"][order]).max() == 0.0
    assert H.rel_err(rows[:, 2:5], fx["mu_out"][order]) < 1e-9
    assert H.rel_err(rows[:, 5:8], fx["ef_static"][order]) < 1e-9


def test_library_interface_extracts_the_arrays(tmp_path):
    """`lammps_extract_atom(lmp, "mu_induced" | "static_polarizability" | "ef_static")` -- what a C or Python driver of
    LAMMPS sees (src/library.cpp -> Atom::extract, three names added by the build): a driver that runs the shipped example
    through the library interface prints the arrays of the reference run"""
    import subprocess
    driver = LMP_AV.parent / LMP_AV.name.replace("lmp_serial_av", "extract_driver_av").replace("lmp_b200", "extract_driver_b200")
    if not driver.exists():
        pytest.skip(f"{driver.name} not built")
    fx = write_h2_data(tmp_path)
    (tmp_path / "in.case").write_text("\n".join(h2_shipped_lines(fx) + H2_DYNAMICS + ["run 0"]) + "\n")
    r = subprocess.run([str(driver), "in.case"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = np.array([[float(t) for t in l.split()] for l in r.stdout.splitlines() if l.strip()])
    assert rows.shape == (fx["x"].shape[0], 8)
    rows = rows[np.argsort(rows[:, 0])]
    order = np.argsort(fx["tag"])
    assert np.array_equal(rows[:, 0].astype(int), fx["tag"][order])
    assert np.abs(rows[:, 1] - fx["alpha"][order]).max() == 0.0
    assert H.rel_err(rows[:, 2:5], fx["mu_out"][order]) < 1e-9
    assert H.rel_err(rows[:, 5:8], fx["ef_static"][order]) < 1e-9


def test_exchange_record_round_trips_every_atom(tmp_path):
    """the record Comm::exchange ships when an atom changes MPI ranks (src/comm_brick.cpp:597-690) cannot occur in this
    single-rank build: the driver packs every atom with AtomVec::pack_exchange and unpacks it as a new atom -- the copy
    carries position, velocity, charge, ids, special list AND polarizability, dipole, static field (the trailer behind the
    stock record), and the unpacker consumes exactly what the packer wrote"""
    import subprocess
    driver = LMP_AV.parent / LMP_AV.name.replace("lmp_serial_av", "extract_driver_av").replace("lmp_b200", "extract_driver_b200")
    if not driver.exists():
        pytest.skip(f"{driver.name} not built")
    fx = write_h2_data(tmp_path)
    (tmp_path / "in.case").write_text("\n".join(h2_shipped_lines(fx) + H2_DYNAMICS + ["run 1"]) + "\n")
    r = subprocess.run([str(driver), "in.case", "exchange"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    w = r.stdout.split()
    assert w[0] == "exchange" and int(w[2]) == fx["x"].shape[0]
    # stock AtomVecFull record (1 length + x v tag type mask image q molecule + bonded topology) + 7 doubles of trailer
    assert int(w[4]) >= 18 + 7 and float(w[7]) == 0.0


def test_replicate_goes_through_the_restart_records(tmp_path):
    """`replicate` packs and unpacks every atom with pack_restart / unpack_restart: the copies keep their
    polarizabilities (sum over atoms doubles)"""
    fx = write_h2_data(tmp_path)
    lines = h2_shipped_lines(fx)
    i = lines.index("kspace_style ewald 1.0e-4")
    lines[i:i] = ["replicate 2 1 1"]
    lines = [l for l in lines if not l.startswith("thermo")]
    lines += ["compute pol all polarization/atom", "compute s all reduce sum c_pol[1]",
              "thermo_style custom step atoms c_s", "thermo_modify format float %.14g", H2_DYNAMICS[1], "run 0"]
    (tmp_path / "in.case").write_text("\n".join(lines) + "\n")
    cols, rows = run_log(LMP_AV, tmp_path, "rep")
    assert rows[0, cols.index("Atoms")] == 2 * fx["x"].shape[0]
    assert abs(rows[0, cols.index("c_s")] - 2.0 * fx["alpha"].sum()) < 1e-9 * fx["alpha"].sum()
