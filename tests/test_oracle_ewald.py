"""CPU tests: the oracle's reciprocal-space Ewald (oracle/polref.c: polref_ewald_*) against the reference's own
KSpace run (src/KSPACE/ewald.cpp).  Golden vectors (oracle/make_golden.py ewald): E_long from the thermo output,
KSpace forces and virial as the difference of two runs with and without `kspace_modify compute no`, the k-vector
statistics the reference prints at setup.  Three systems: Bulk H2 (cubic, 1e-4), MOF5+methane (cubic, 1e-6),
a non-cubic brick of the synthetic fluid (1e-5: kxmax != kymax != kzmax)."""
import numpy as np
import pytest

from oracle import polref as P
import polhelpers as H

CASES = ["ewald_h2", "ewald_methane", "ewald_brick"]


@pytest.mark.parametrize("case", CASES)
def test_ewald_oracle_matches_reference(case, golden_dir):
    g = np.load(golden_dir / f"{case}.npz")
    prd = g["boxhi"] - g["boxlo"]
    plan = P.ewald_plan(float(g["accuracy"]), g["q"], float(g["cut_coul"]), prd)
    assert plan.kcount == int(g["kcount"])                                  # k-vector set: exact
    assert (plan.kxmax, plan.kymax, plan.kzmax) == tuple(int(v) for v in g["kxyzmax"])
    assert abs(plan.g_ewald - float(g["g_ewald_printed"])) < 5e-6 * plan.g_ewald  # printed with 6 significant digits
    r = P.ewald_compute(plan, g["x"], g["q"], prd)
    assert abs(r["energy"] - float(g["elong"])) < 1e-12 * abs(float(g["elong"]))
    # forces / virial of the reference are differences of two runs: ~1e-12 relative noise
    assert np.abs(r["f"] - g["f_kspace"]).max() < 1e-10 * np.abs(g["f_kspace"]).max()
    assert H.rel_err(r["virial"], g["virial_kspace"]) < 1e-9


def test_ewald_translation_invariance(golden_dir):
    """size-independent property: shifting every atom by the same vector changes neither energy nor forces."""
    g = np.load(golden_dir / "ewald_brick.npz")
    prd = g["boxhi"] - g["boxlo"]
    plan = P.ewald_plan(float(g["accuracy"]), g["q"], float(g["cut_coul"]), prd)
    a = P.ewald_compute(plan, g["x"], g["q"], prd)
    b = P.ewald_compute(plan, g["x"] + np.array([1.234, -0.5, 7.0]), g["q"], prd)
    assert abs(a["energy"] - b["energy"]) < 1e-11 * abs(a["energy"])
    assert np.abs(a["f"] - b["f"]).max() < 1e-10 * np.abs(a["f"]).max()
    assert np.abs(a["f"].sum(0)).max() < 1e-10 * np.abs(a["f"]).max()       # no net force
