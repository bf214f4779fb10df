"""Pins the PPPM oracle (oracle/pppmref.py, restatement of src/KSPACE/pppm.cpp) against golden vectors from the
reference binary (tests/golden/pppm_*.npz, oracle/make_golden.py pppm): grid and g_ewald selection, E_long, per-atom
KSpace forces and the KSpace virial -- forces and virial as the difference of two reference runs with and without
`kspace_modify compute no`.  Stage 1 of the PPPM widening: no device code yet."""
import numpy as np
import pytest

import polhelpers as H
from oracle import pppmref as PP

CASES = ["pppm_h2", "pppm_methane", "pppm_brick", "pppm_brick_order4"]


def plan_for(g):
    km = str(g["kspace_modify"]).split()
    kw = {}
    if "order" in km:
        kw["order"] = int(km[km.index("order") + 1])
    if "mesh" in km:
        i = km.index("mesh")
        kw["mesh"] = [int(v) for v in km[i + 1:i + 4]]
    if "gewald" in km:
        kw["g_ewald"] = float(km[km.index("gewald") + 1])
    return PP.PPPMPlan(float(g["accuracy"]), g["q"], float(g["cut_coul"]), g["boxhi"] - g["boxlo"], **kw)


@pytest.mark.parametrize("case", CASES)
def test_pppm_oracle_matches_reference(case, golden_dir):
    g = np.load(golden_dir / f"{case}.npz")
    plan = plan_for(g)
    assert tuple(plan.n) == tuple(int(v) for v in g["grid"])               # grid selection: exact
    assert plan.order == int(g["order"])
    assert f"{plan.g_ewald:g}" == f"{float(g['g_ewald_printed']):g}"         # g_ewald to the 6 digits the log prints
    r = plan.compute(g["x"], g["q"], g["boxlo"])
    assert abs(r["energy"] - float(g["elong"])) < 1e-11 * abs(float(g["elong"]))
    assert np.abs(r["f"] - g["f_kspace"]).max() < 1e-10 * np.abs(g["f_kspace"]).max()
    assert H.rel_err(r["virial"], g["virial_kspace"]) < 1e-9


def test_pppm_and_ewald_agree_to_their_accuracy(golden_dir):
    """independent sanity: PPPM at 1e-6 and the (pinned) direct Ewald sum of the same system give the same forces to
    the accuracies they were asked for -- different g_ewald, so only real + reciprocal space together would agree
    exactly; here: the reciprocal-space energies differ by the known real-space complement"""
    gp, ge = np.load(golden_dir / "pppm_methane.npz"), np.load(golden_dir / "ewald_methane.npz")
    assert np.array_equal(gp["x"], ge["x"])
    # both fixtures asked for 1e-6 relative accuracy: the total Coulomb force is what agrees, which needs the pair part;
    # the k-space parts alone must at least be strongly correlated
    a, b = gp["f_kspace"].ravel(), ge["f_kspace"].ravel()
    assert np.corrcoef(a, b)[0, 1] > 0.99
