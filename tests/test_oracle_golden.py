"""CPU tests: the oracle restatement (oracle/polref.c) against the reference's own numbers.

  * golden npz fixtures = inputs/outputs of single compute() calls of the repaired reference binary
    (oracle/make_golden.py); the restatement must reproduce dipoles, fields, forces, energies, virial,
    iteration counts and the half neighbor list;
  * thermo tables committed by the reference authors (polarization/examples/*/log.lammps).
"""
import numpy as np
import pytest

from oracle import polref as P
import polhelpers as H

CASES = H.golden_cases()
TOL = 2e-12  # oracle vs reference binary: same algorithm, same order; differences are libm/ordering noise


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference_compute(case):
    fx = H.load_fixture(case)
    sysm = H.system_from_fixture(fx)
    st = H.style_from_fixture(fx)
    lists = H.reference_lists(fx, sysm, st, case)
    r = P.compute(sysm, st, mu_in=fx["mu_in"], eflag=int(fx["eflag"]), vflag=int(fx["vflag"]), use_matrix=True,
                  lists=lists)
    xall, owner, shift, numneigh, first, neigh = lists
    assert len(owner) == int(fx["nghost"])
    if "newtonoff" in case:
        # the reference's newton-off list holds every owned-ghost pair from both sides; the restatement keeps the
        # newton-on list (each pair once), which yields the same forces, energies and pairwise virial
        assert int(fx["npairs_half"]) > len(neigh)
    else:
        assert np.array_equal(numneigh, fx["numneigh_half"])      # neighbor list: bit exact
        assert len(neigh) == int(fx["npairs_half"])
    if "eatom" in fx:                                             # per-atom tallies (pe/atom, stress/atom)
        assert H.rel_err(r["eatom"], fx["eatom"]) < TOL and H.rel_err(r["vatom"], fx["vatom"]) < TOL
    assert r["iterations"] == int(fx["iterations"])               # iteration count: exact
    assert H.rel_err(r["ef_static"], fx["ef_static"]) < TOL
    assert H.rel_err(r["mu"], fx["mu_out"]) < TOL
    assert H.rel_err(r["f"], fx["f"]) < 1e-11
    for k in ("eng_vdwl", "eng_coul", "eng_pol"):
        assert abs(r[k] - float(fx[k])) <= 1e-12 * max(1.0, abs(float(fx[k])))
    assert H.rel_err(r["virial"], fx["virial"]) < 1e-10


def test_matrix_free_is_bitwise_the_matrix_form():
    fx = H.load_fixture("h2_default_step0")
    sysm, st = H.system_from_fixture(fx), H.style_from_fixture(fx)
    a = P.compute(sysm, st, mu_in=fx["mu_in"], use_matrix=True)
    b = P.compute(sysm, st, mu_in=fx["mu_in"], use_matrix=False, lists=a["lists"])
    assert a["iterations"] == b["iterations"]
    assert np.array_equal(a["mu"], b["mu"]) and np.array_equal(a["f_all"], b["f_all"])


def test_half_list_matches_reference_pair_by_pair():
    fx = H.load_fixture("h2_default_step0")
    sysm, st = H.system_from_fixture(fx), H.style_from_fixture(fx)
    xall, owner, shift = P.build_ghosts(sysm, st.cutneighmax)
    numneigh, first, neigh = P.build_half_list(sysm, st, xall, owner)
    n = sysm.n
    ii = np.repeat(np.arange(n), numneigh)
    j = neigh & 0x3FFFFFFF
    sb = (neigh >> 30) & 3
    allowner = np.concatenate([np.arange(n), owner])
    allshift = np.concatenate([np.zeros((n, 3), dtype=np.int32), shift])
    assert np.array_equal(ii, fx["half_i"].astype(np.int64))
    assert np.array_equal(allowner[j], fx["half_j"].astype(np.int64))
    assert np.array_equal(allshift[j], fx["half_shift"].astype(np.int32))
    assert np.array_equal(sb, fx["half_special"].astype(np.int64))


def test_tables_match_reference_bitwise():
    fx = H.load_fixture("h2_default_step0")
    st = H.style_from_fixture(fx)
    assert st.ncoulmask == int(fx["ncoulmask"]) and st.ncoulshiftbits == int(fx["ncoulshiftbits"])
    assert st.tabinnersq == float(fx["tabinnersq"])
    for k, v in st.tables.items():
        assert np.array_equal(v, fx["tab_" + k]), k


def test_ewald_g_matches_reference():
    fx = H.load_fixture("h2_default_step0")
    sysm = H.system_from_fixture(fx)
    g = P.ewald_g(1e-4, sysm.q, 10.797442, sysm.boxlo, sysm.boxhi)
    assert g == float(fx["g_ewald"])


@pytest.mark.parametrize("key,case", [("h2", "h2_default"), ("methane", "methane_default")])
def test_shipped_log_thermo_columns(key, case):
    """E_vdwl / E_coul / E_pol of the authors' committed logs, to the 8 digits they print."""
    logs = H.thermo_logs()
    rows = logs[key]["rows"]
    for step in range(3):
        fx = H.load_fixture(f"{case}_step{step}")
        sysm, st = H.system_from_fixture(fx), H.style_from_fixture(fx)
        r = P.compute(sysm, st, mu_in=fx["mu_in"], use_matrix=False,
                      lists=H.reference_lists(fx, sysm, st, f"{case}_step{step}"))
        row = rows[step]
        for col, val in (("E_vdwl", r["eng_vdwl"]), ("E_coul", r["eng_coul"]), ("E_pol", r["eng_pol"])):
            assert f"{val:.8g}" == f"{float(row[col]):.8g}", (step, col, val, row[col])


def test_rows_form_agrees_with_literal_form():
    fx = H.load_fixture("h2_jacobi_fixed3_step0")
    sysm, st = H.system_from_fixture(fx), H.style_from_fixture(fx)
    a = P.compute(sysm, st, mu_in=fx["mu_in"], use_matrix=False, trace_max=4)
    b = P.polar_rows(sysm, st, mu_in=fx["mu_in"], trace_max=4)
    assert a["iterations"] == b["iterations"] == 3
    assert np.array_equal(a["ef_static"], b["ef_static"])
    assert np.array_equal(a["trace"], b["trace"])            # Jacobi rows: bitwise the scatter loops
    assert np.array_equal(a["mu"], b["mu"])
    assert abs(a["eng_pol"] - b["eng_pol"]) < 1e-12 * abs(a["eng_pol"])
    # forces: literal f also holds LJ/Coulomb; compare the polarization increment
    st0 = H.style_from_fixture(fx, zodid=0)
    lj = P.compute(sysm, _nopol(fx), mu_in=fx["mu_in"], lists=a["lists"])
    assert H.rel_err(a["f"] - lj["f"], b["f"]) < 1e-11


def _nopol(fx):
    # alpha*E with all polarizabilities zeroed is handled by the caller; here: zodid + gamma 0
    st = H.style_from_fixture(fx, polar_gs_ranked=0, zodid=1, use_previous=0, polar_gamma=0.0)
    return st


def test_rows_form_truncated_equals_literal_truncated():
    fx = H.load_fixture("h2_jacobi_fixed3_step0")
    sysm = H.system_from_fixture(fx)
    st = H.style_from_fixture(fx, polar_cut=7.5)
    a = P.compute(sysm, st, mu_in=fx["mu_in"], use_matrix=False)
    b = P.polar_rows(sysm, st, mu_in=fx["mu_in"])
    assert np.array_equal(a["mu"], b["mu"])
    assert abs(a["eng_pol"] - b["eng_pol"]) < 1e-12 * abs(a["eng_pol"])


def test_chunked_gs_converges_to_sequential_gs():
    fx = H.load_fixture("h2_default_step0")
    sysm = H.system_from_fixture(fx)
    seq = P.polar_rows(sysm, H.style_from_fixture(fx), mu_in=fx["mu_in"])
    chk = P.polar_rows(sysm, H.style_from_fixture(fx, gs_chunks=8), mu_in=fx["mu_in"])
    # both stop when the rms change drops below precision=1e-11; fixed points agree to ~precision
    assert np.abs(seq["mu"] - chk["mu"]).max() < 50 * 1e-11
    assert abs(seq["eng_pol"] - chk["eng_pol"]) < 1e-9 * abs(seq["eng_pol"])


def test_interleaved_colouring_converges_like_sequential_gs():
    """oracle-level statement of the device default: 8 interleaved chunks reach the sequential sweep's fixed point in
    about as many iterations, contiguous chunks need several times more (rigid water box, strong intramolecular
    coupling)."""
    sysm = H.water_box(6)
    kw = dict(polar_cut=9.0, damp_type="exponential", polar_gs_ranked=1, precision=1e-11, max_iterations=300,
              polar_gamma=1.03)
    seq = P.polar_rows(sysm, H.water_style(sysm, 2.5, 9.0, gs_chunks=0, **kw))
    il = P.polar_rows(sysm, H.water_style(sysm, 2.5, 9.0, gs_chunks=-8, **kw))
    co = P.polar_rows(sysm, H.water_style(sysm, 2.5, 9.0, gs_chunks=8, **kw))
    assert not il["diverged"] and not co["diverged"]
    assert il["iterations"] <= seq["iterations"] + 4 and co["iterations"] >= il["iterations"] + 10
    assert np.abs(il["mu"] - seq["mu"]).max() < 1e-9 and np.abs(co["mu"] - seq["mu"]).max() < 1e-9
