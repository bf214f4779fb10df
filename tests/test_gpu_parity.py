"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI.

Bars (BASELINE.json north_star):
  * neighbor pair sets and iteration counts: exact;
  * Jacobi / identity-order Gauss-Seidel: dipoles, fields, forces, energies within 1e-10 relative of the
    reference's own compute() (golden fixtures dumped from the reference binary) and of the oracle;
  * ranked Gauss-Seidel precision modes: converged dipoles within 20*polar_precision absolute
    (the stopping rule bounds the rms change per sweep by polar_precision, pol.cpp:1205-1209),
    energies within 1e-8 relative.
"""
import numpy as np
import pytest

import polhelpers as H
from gpu_common import configure_from_fixture, pack_pairs, pb, run_fixture, run_system
from oracle import polref as P

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture()
def style():
    s = pb.PairStyle(device=0)
    yield s
    s.close()


def check_against_fixture(res, mu, ef, f, fx, mu_tol_abs=None, e_tol=TOL):
    assert H.rel_err(ef, fx["ef_static"]) < TOL
    if mu_tol_abs is None:
        assert H.rel_err(mu, fx["mu_out"]) < TOL
    else:
        assert np.abs(mu - fx["mu_out"]).max() < mu_tol_abs
    fscale = np.abs(fx["f"]).max()
    assert np.abs(f - fx["f"]).max() < (TOL if mu_tol_abs is None else 1e-7) * fscale
    for k in ("eng_vdwl", "eng_coul", "eng_pol"):
        ref = float(fx[k])
        assert abs(getattr(res, k) - ref) <= e_tol * max(1.0, abs(ref)), (k, getattr(res, k), ref)
    vir = np.array(res.virial[:])
    assert H.rel_err(vir, fx["virial"]) < (1e-9 if mu_tol_abs is None else 1e-7)


JACOBI_CASES = ["h2_jacobi_fixed3_step0", "h2_jacobi_fixed30_step0", "h2_jacobi_precision_step0",
                "h2_zodid_step0",
                # neigh_modify exclude molecule/intra | type + group + molecule/intra of a sub-group
                "h2_exclude_intra_step0", "h2_exclude_mixed_step0"]


@pytest.mark.parametrize("case", JACOBI_CASES)
def test_jacobi_and_zodid_match_reference(style, case):
    fx = H.load_fixture(case)
    configure_from_fixture(style, fx)
    res, mu, ef, f = run_fixture(style, fx)
    assert res.iterations == int(fx["iterations"])
    assert res.status & pb.STATUS_EXACT
    check_against_fixture(res, mu, ef, f, fx)


@pytest.mark.parametrize("case", ["h2_gs_step0", "h2_gs_fixed3_step0"])
def test_sequential_gauss_seidel_identity_order(style, case):
    fx = H.load_fixture(case)
    configure_from_fixture(style, fx)
    res, mu, ef, f = run_fixture(style, fx)
    assert res.iterations == int(fx["iterations"])
    check_against_fixture(res, mu, ef, f, fx)


@pytest.mark.parametrize("case", ["h2_default_step0", "h2_nodamp_step0", "h2_noprev_step0", "h2_notable_step0",
                                  "methane_default_step0", "co2_singlepoint_step0"])
def test_ranked_gauss_seidel_defaults(style, case):
    fx = H.load_fixture(case)
    configure_from_fixture(style, fx)
    res, mu, ef, f = run_fixture(style, fx)
    prec = 1e-15 if case.startswith("co2") else 1e-11
    assert abs(res.iterations - int(fx["iterations"])) <= 1
    check_against_fixture(res, mu, ef, f, fx, mu_tol_abs=max(20 * prec, 1e-12), e_tol=1e-8)


def test_divergence_path(style):
    fx = H.load_fixture("h2_diverge_step0")
    configure_from_fixture(style, fx)
    res, mu, ef, f = run_fixture(style, fx)
    assert res.status & pb.STATUS_DIVERGED and res.iterations == int(fx["iterations"]) == 4
    check_against_fixture(res, mu, ef, f, fx)       # mu = alpha*E exactly: tight tolerance


@pytest.mark.parametrize("case", ["h2_default", "methane_default", "h2_jacobi_fixed3"])
def test_multi_step_with_stale_lists_and_use_previous(style, case):
    """steps 0..k of one reference run: lists built at step 0 (ago=0) and reused (ago>0), dipoles carried."""
    fx0 = H.load_fixture(f"{case}_step0")
    configure_from_fixture(style, fx0)
    logs = H.thermo_logs()
    key = "h2" if case == "h2_default" else ("methane" if case == "methane_default" else None)
    step = 0
    mu_prev = None
    while (H.GOLDEN / f"{case}_step{step}.npz").exists():
        fx = H.load_fixture(f"{case}_step{step}")
        res, mu, ef, f = run_fixture(style, fx, ago=step, mu_in=fx["mu_in"] if mu_prev is None else mu_prev)
        jac = "jacobi" in case
        check_against_fixture(res, mu, ef, f, fx, mu_tol_abs=None if jac else 2e-10, e_tol=TOL if jac else 1e-8)
        if key:   # the reference authors' own log, 8 printed digits
            row = logs[key]["rows"][step]
            for col, val in (("E_vdwl", res.eng_vdwl), ("E_coul", res.eng_coul), ("E_pol", res.eng_pol)):
                ref = float(row[col])
                half_unit = 0.5 * 10.0 ** (np.floor(np.log10(abs(ref))) - 7)   # 8 printed digits
                assert abs(val - ref) <= 1.02 * half_unit, (step, col, val, row[col])
        mu_prev = mu
        step += 1
    assert step >= 2


def test_neighbor_pair_set_is_exactly_the_reference_list(style):
    fx = H.load_fixture("h2_default_step0")
    configure_from_fixture(style, fx)
    res, mu, ef, f = run_fixture(style, fx)
    n = fx["x"].shape[0]
    ng = res.nghost
    perm = style.debug_fetch("perm", np.int32, n)
    rowstart = style.debug_fetch("rowstart", np.uint64, n + 1).astype(np.int64)
    neigh = style.debug_fetch("neigh", np.int32, int(rowstart[-1]))
    gowner = style.debug_fetch("ghost_owner", np.int32, ng)
    gcode = style.debug_fetch("ghost_shift", np.int32, ng)
    gshift = np.stack([(gcode & 3) - 1, ((gcode >> 2) & 3) - 1, ((gcode >> 4) & 3) - 1], 1)
    assert rowstart[-1] == res.npairs_full == 2 * int(fx["npairs_half"])
    i_sorted = np.repeat(np.arange(n), np.diff(rowstart))
    j = neigh & 0x3FFFFFFF
    sb = (neigh >> 30) & 3
    owner_sorted = np.where(j < n, j, gowner[np.clip(j - n, 0, max(ng - 1, 0))])
    shift = np.where((j < n)[:, None], 0, gshift[np.clip(j - n, 0, max(ng - 1, 0))])
    dev = np.sort(pack_pairs(perm[i_sorted], perm[owner_sorted], shift, sb))
    # reference half list, symmetrised: (i, j, s) and (j, i, -s)
    hi, hj = fx["half_i"].astype(np.int64), fx["half_j"].astype(np.int64)
    hs, hb = fx["half_shift"].astype(np.int64), fx["half_special"].astype(np.int64)
    ref = np.sort(np.concatenate([pack_pairs(hi, hj, hs, hb), pack_pairs(hj, hi, -hs, hb)]))
    assert np.array_equal(dev, ref)


def test_list_mode_matches_oracle_truncated(style):
    """polar_cutoff extension on the synthetic LJ+charge fluid (BASELINE config 2 at reduced N)."""
    sysm = H.lj_charge_fluid(10)                       # 4000 atoms, L = 34.2 A
    kw = dict(fixed_iteration=1, max_iterations=5, damp_type="exponential", polar_gs_ranked=0)
    st = H.fluid_style(sysm, 2.5, 12.0, polar_cut=12.0, **kw)
    ref = P.polar_rows(sysm, st)
    lit = P.compute(sysm, H.fluid_style(sysm, 2.5, 12.0, polar_gs_ranked=0, zodid=1, polar_gamma=0.0))
    style.set_ntypes(2)
    style.command("pair_style lj/cut/coul/long/polarization 2.5 12.0 polar_gs_ranked no fixed_iteration yes "
                  "max_iterations 5 damp_type exponential polar_cutoff 12.0")
    style.command("pair_coeff 1 1 0.1 3.0")
    style.command("pair_coeff 2 2 0.1 3.0")
    style.init(g_ewald=st.g_ewald, molecular=0)
    style.set_box(sysm.boxlo, sysm.boxhi)
    res, mu, ef, f = run_system(style, sysm)
    assert not (res.status & pb.STATUS_EXACT) and res.iterations == 5
    assert H.rel_err(ef, ref["ef_static"]) < TOL
    assert H.rel_err(mu, ref["mu"]) < TOL
    assert abs(res.eng_pol - ref["eng_pol"]) < TOL * abs(ref["eng_pol"])
    assert abs(res.eng_vdwl - lit["eng_vdwl"]) < TOL * abs(lit["eng_vdwl"])
    assert abs(res.eng_coul - lit["eng_coul"]) < TOL * abs(lit["eng_coul"])
    ftot = lit["f"] + ref["f"]
    assert np.abs(f - ftot).max() < TOL * np.abs(ftot).max()
    vir = lit["virial"] + ref["virial"]
    assert H.rel_err(np.array(res.virial[:]), vir) < 1e-9


def test_list_mode_ranked_chunks_match_oracle_chunks(style):
    sysm = H.lj_charge_fluid(8)                        # 2048 atoms, L = 27.4 A
    st = H.fluid_style(sysm, 2.5, 12.0, polar_cut=12.0, damp_type="exponential", polar_gs_ranked=1,
                       gs_chunks=8, precision=1e-11, max_iterations=50)
    ref = P.polar_rows(sysm, st)
    style.set_ntypes(2)
    style.command("pair_style lj/cut/coul/long/polarization 2.5 12.0 damp_type exponential precision 1e-11 "
                  "polar_cutoff 12.0 gs_chunks 8")
    style.command("pair_coeff * * 0.1 3.0")
    style.init(g_ewald=st.g_ewald, molecular=0)
    style.set_box(sysm.boxlo, sysm.boxhi)
    res, mu, ef, f = run_system(style, sysm)
    assert abs(res.iterations - ref["iterations"]) <= 1
    assert np.abs(mu - ref["mu"]).max() < 20 * 1e-11
    assert abs(res.eng_pol - ref["eng_pol"]) < 1e-8 * abs(ref["eng_pol"])


def test_exact_mode_equals_list_mode_when_cutoff_covers_everything(style):
    """size-independent property: with polar_cutoff = cut_coul = L/2 - eps ... the two device paths agree
    on the static field and on LJ/Coulomb (identical pair sets), independent of the oracle."""
    sysm = H.lj_charge_fluid(6)                        # 864 atoms, L = 20.5
    L = float(sysm.boxhi[0])
    cc = 0.5 * L - 1e-6
    outs = []
    for extra in ("", f" polar_cutoff {cc}"):
        s = pb.PairStyle(device=0)
        s.set_ntypes(2)
        s.command(f"pair_style lj/cut/coul/long/polarization 2.5 {cc} polar_gs_ranked no zodid yes" + extra)
        s.command("pair_coeff * * 0.1 3.0")
        s.init(g_ewald=0.3, molecular=0)
        s.set_box(sysm.boxlo, sysm.boxhi)
        outs.append(run_system(s, sysm))
        s.close()
    (ra, mua, efa, fa), (rb, mub, efb, fb) = outs
    assert H.rel_err(efa, efb) < 1e-12 and H.rel_err(mua, mub) < 1e-12
    assert abs(ra.eng_coul - rb.eng_coul) < 1e-12 * abs(ra.eng_coul)


def test_config2_32k_list_mode_sampled_against_oracle(style):
    """BASELINE config 2 at full size (32 000 atoms, cut 2.5/12): device list path vs the OpenMP row oracle."""
    sysm = H.lj_charge_fluid(20)
    assert sysm.n == 32000
    st = H.fluid_style(sysm, 2.5, 12.0, polar_cut=12.0, fixed_iteration=1, max_iterations=2,
                       damp_type="exponential", polar_gs_ranked=0)
    ref = P.polar_rows(sysm, st)
    style.set_ntypes(2)
    style.command("pair_style lj/cut/coul/long/polarization 2.5 12.0 polar_gs_ranked no fixed_iteration yes "
                  "max_iterations 2 damp_type exponential polar_cutoff 12.0")
    style.command("pair_coeff * * 0.1 3.0")
    style.init(g_ewald=st.g_ewald, molecular=0)
    style.set_box(sysm.boxlo, sysm.boxhi)
    res, mu, ef, f = run_system(style, sysm)
    assert H.rel_err(ef, ref["ef_static"]) < TOL and H.rel_err(mu, ref["mu"]) < TOL
    assert abs(res.eng_pol - ref["eng_pol"]) < TOL * abs(ref["eng_pol"])
    # size-independent: Newton's third law, total pair force vanishes
    assert np.abs(f.sum(0)).max() < 1e-9 * np.abs(f).max() * np.sqrt(sysm.n)


@pytest.mark.parametrize("words", ["polar_gs_ranked yes fixed_iteration yes max_iterations 3",
                                   "polar_gs_ranked no polar_gs yes fixed_iteration yes max_iterations 2"])
def test_sequential_gs_is_iteration_for_iteration_the_oracle(style, words):
    """Same visiting order, same in-place updates: after a FIXED number of Gauss-Seidel sweeps the device
    dipoles equal the literal restatement's to 1e-10 (needs the rank order incl. tie-breaks to be identical)."""
    fx = dict(H.load_fixture("h2_default_step0"))
    base = "pair_style lj/cut/coul/long/polarization 2.5 10.797442 damp_type exponential damp 2.1304 use_previous yes "
    fx["pair_style"] = np.array(base + words)
    configure_from_fixture(style, fx)
    res, mu, ef, f = run_fixture(style, fx)
    sysm, st = H.system_from_fixture(fx), H.style_from_fixture(fx)
    ref = P.compute(sysm, st, mu_in=fx["mu_in"], use_matrix=False)
    assert res.iterations == ref["iterations"]
    assert H.rel_err(mu, ref["mu"]) < TOL
    assert abs(res.eng_pol - ref["eng_pol"]) < TOL * abs(ref["eng_pol"])
    assert np.abs(f - ref["f"]).max() < TOL * np.abs(ref["f"]).max()


@pytest.mark.parametrize("case", ["h2_default_step0", "methane_default_step0"])
def test_cluster_sweep_equals_the_launch_chain(case):
    """Exact-mode Gauss-Seidel: the 16-CTA cluster kernel that walks the blocks of a sweep in one launch (default) and the
    chain of one launch per block evaluate the same substitution -- same iteration count, dipoles and forces to rounding."""
    fx = H.load_fixture(case)
    out = []
    for cluster in (16, 8, 0):
        s = pb.PairStyle(device=0)
        configure_from_fixture(s, fx)
        s.set_option("gs_cluster", cluster)
        res, mu, ef, f = run_fixture(s, fx)
        out.append((res.iterations, mu, f, res.eng_pol))
        s.close()
    for it, mu, f, e in out[:2]:
        assert it == out[2][0]
        assert H.rel_err(mu, out[2][1]) < 1e-12
        assert np.abs(f - out[2][2]).max() < 1e-12 * np.abs(out[2][2]).max()
        assert abs(e - out[2][3]) < 1e-12 * abs(out[2][3])


def _fluid_style_on_device(style, sysm, g_ewald, words):
    style.set_ntypes(2)
    style.command("pair_style lj/cut/coul/long/polarization 2.5 12.0 " + words)
    style.command("pair_coeff * * 0.1 3.0")
    style.init(g_ewald=g_ewald, molecular=0)
    style.set_box(sysm.boxlo, sysm.boxhi)


@pytest.mark.parametrize("drop", [0, 1])
def test_every_sweep_kernel_gives_the_same_dipoles(drop):
    """The list-mode sweep exists in several realisations (first version, matrix-free, per-atom radial cache,
    pair groups with register prefetch, pair groups fed by TMA bulk copies = default).  All evaluate the same
    pair terms; they must agree to rounding, also with an odd atom count (a one-member group) and in
    precision mode (same iteration count)."""
    sysm = H.lj_charge_fluid(10)
    if drop:  # 3999 atoms: odd rows, single-member groups
        keep = np.arange(sysm.n) != 1234
        sysm = P.System(sysm.x[keep], sysm.q[keep], sysm.type[keep], sysm.molecule[keep], sysm.alpha[keep],
                        sysm.boxlo, sysm.boxhi, 2)
    g = P.ewald_g(1e-4, sysm.q, 12.0, sysm.boxlo, sysm.boxhi)
    ref = None
    for variant in (0, 6, 20, 31, 41):
        for words in ("polar_gs_ranked no fixed_iteration yes max_iterations 6 damp_type exponential polar_cutoff 12.0",
                      "polar_gs_ranked no precision 1e-9 max_iterations 80 damp_type exponential polar_cutoff 12.0"):
            s = pb.PairStyle(device=0)
            _fluid_style_on_device(s, sysm, g, words)
            s.set_option("sweep_variant", variant)
            res, mu, ef, f = run_system(s, sysm)
            s.close()
            key = words.split()[2]
            if ref is None:
                ref = {}
            if key not in ref:
                ref[key] = (res.iterations, mu, f, res.eng_pol)
                continue
            it0, mu0, f0, e0 = ref[key]
            assert res.iterations == it0, (variant, key)
            assert H.rel_err(mu, mu0) < 1e-12 and np.abs(f - f0).max() < 1e-12 * np.abs(f0).max(), (variant, key)
            assert abs(res.eng_pol - e0) < 1e-12 * abs(e0)


def _water_on_device(style, sysm, st, words):
    style.set_ntypes(2)
    style.command("pair_style lj/cut/coul/long/polarization 2.5 12.0 " + words)
    style.command("pair_coeff 1 1 0.155 3.166")
    style.command("pair_coeff 2 2 0.0 1.0")
    style.init(g_ewald=st.g_ewald, molecular=0)
    style.set_box(sysm.boxlo, sysm.boxhi)


def test_water_box_jacobi_matches_oracle(style):
    """BASELINE config 3 shape at reduced N (3000 atoms): molecule ids exclude intramolecular pairs from the
    static field and the charge-dipole terms, but not from dipole-dipole (reference semantics)."""
    sysm = H.water_box(10)
    st = H.water_style(sysm, 2.5, 12.0, polar_cut=12.0, damp_type="exponential", polar_gs_ranked=0,
                       fixed_iteration=1, max_iterations=10)
    ref = P.polar_rows(sysm, st)
    _water_on_device(style, sysm, st, "polar_gs_ranked no fixed_iteration yes max_iterations 10 damp_type exponential "
                                       "polar_cutoff 12.0")
    res, mu, ef, f = run_system(style, sysm)
    assert res.iterations == 10
    assert H.rel_err(ef, ref["ef_static"]) < TOL and H.rel_err(mu, ref["mu"]) < TOL
    assert abs(res.eng_pol - ref["eng_pol"]) < TOL * abs(ref["eng_pol"])


def test_water_box_ranked_colouring_sweep_converges_to_the_oracle(style):
    """config 3 keywords (precision 1e-11, polar_gs_ranked yes, polar_gamma 1.03): ranked colouring sweep vs the
    oracle's emulation of the same chunking -- same fixed point within 20*precision, iteration counts +-2; and
    the fixed point equals the strictly sequential Gauss-Seidel one (the reference's order)."""
    sysm = H.water_box(10)
    kw = dict(polar_cut=12.0, damp_type="exponential", polar_gs_ranked=1, precision=1e-11, max_iterations=200,
              polar_gamma=1.03)
    st = H.water_style(sysm, 2.5, 12.0, gs_chunks=8, **kw)
    ref = P.polar_rows(sysm, st)
    seq = P.polar_rows(sysm, H.water_style(sysm, 2.5, 12.0, gs_chunks=0, **kw))
    _water_on_device(style, sysm, st, "damp_type exponential precision 1e-11 max_iterations 200 polar_gamma 1.03 "
                                       "polar_cutoff 12.0 gs_chunks 8")
    res, mu, ef, f = run_system(style, sysm)
    assert not (res.status & pb.STATUS_DIVERGED)
    assert abs(res.iterations - ref["iterations"]) <= 2
    assert np.abs(mu - ref["mu"]).max() < 20 * 1e-11
    assert np.abs(mu - seq["mu"]).max() < 100 * 1e-11
    assert abs(res.eng_pol - seq["eng_pol"]) < 1e-8 * abs(seq["eng_pol"])


def test_per_atom_energy_and_virial_match_reference(style):
    """eflag & 2 / vflag & 4: Pair::ev_tally / ev_tally_xyz per-atom tallies (src/pair.cpp:854-949,1001-1089),
    golden arrays dumped from the reference binary run with compute pe/atom + stress/atom."""
    fx = H.load_fixture("h2_peratom_step0")
    assert int(fx["eflag"]) == 3 and int(fx["vflag"]) == 6
    configure_from_fixture(style, fx)
    pa = {}
    res, mu, ef, f = run_fixture(style, fx, peratom=pa)
    check_against_fixture(res, mu, ef, f, fx)
    assert np.abs(pa["eatom"] - fx["eatom"]).max() < TOL * np.abs(fx["eatom"]).max()
    assert np.abs(pa["vatom"] - fx["vatom"]).max() < TOL * np.abs(fx["vatom"]).max()
    # size-independent property: the per-atom energies add up to the global pair energies
    assert abs(pa["eatom"].sum() - (res.eng_vdwl + res.eng_coul)) < 1e-9 * abs(res.eng_coul)
    # requesting the tallies without arrays is refused
    with pytest.raises(pb.Polb200Error):
        run_fixture(style, fx)


def test_per_atom_tallies_list_mode_equal_exact_mode():
    """list-mode kernels (neighbor-list rows) against the all-pairs kernels with a cutoff that covers every pair."""
    sysm = H.lj_charge_fluid(6)
    cc = 0.5 * float(sysm.boxhi[0]) - 1e-6
    outs = []
    for extra in ("", f" polar_cutoff {cc}"):
        s = pb.PairStyle(device=0)
        s.set_ntypes(2)
        s.command(f"pair_style lj/cut/coul/long/polarization 2.5 {cc} polar_gs_ranked no fixed_iteration yes "
                  f"max_iterations 4 damp_type exponential" + extra)
        s.command("pair_coeff * * 0.1 3.0")
        s.init(g_ewald=0.3, molecular=0)
        s.set_box(sysm.boxlo, sysm.boxhi)
        n = sysm.n
        mu, f, ea, va = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n), np.zeros((n, 6))
        from gpu_common import c
        res = s.compute(c(sysm.x, np.float64), c(sysm.q, np.float64), c(sysm.type, np.int32), c(sysm.alpha, np.float64),
                        mu, f, eflag=3, vflag=5, eatom=ea, vatom=va)
        outs.append((res, ea, va, f))
        s.close()
    (ra, ea, va, fa), (rb, eb, vb, fb) = outs
    # LJ + Coulomb see the same pair set in both modes (dipole-dipole does not: exact mode has no cutoff)
    assert H.rel_err(ea, eb) < 1e-11
    # pairwise global virial (vflag & 3 == 1) equals the sum of the per-atom virials, in either mode
    assert H.rel_err(va.sum(0), np.array(ra.virial[:])) < 1e-10
    assert H.rel_err(vb.sum(0), np.array(rb.virial[:])) < 1e-10


def test_newton_pair_off_matches_reference(style):
    """`newton off`: the reference switches to the half/bin/newtoff list, skips ghost forces and tallies the virial
    pairwise (vflag = 1); the owner-computes device path needs no second code path to reproduce it."""
    fx = H.load_fixture("h2_newtonoff_step0")
    assert int(fx["vflag"]) == 1
    configure_from_fixture(style, fx, newton_pair=0)
    res, mu, ef, f = run_fixture(style, fx)
    check_against_fixture(res, mu, ef, f, fx)


@pytest.mark.parametrize("n", [1, 2, 5, 33])
def test_tiny_systems(n):
    """Edge sizes: fewer atoms than a warp, a single atom (no partner at all), odd group counts; exact mode against the
    literal oracle, list mode against the row oracle."""
    rng = np.random.default_rng(100 + n)
    L = 30.0
    x = rng.uniform(2.0, L - 2.0, size=(n, 3))
    typ = (np.arange(n) % 2 + 1).astype(np.int32)
    q = np.where(typ == 1, 0.4, -0.4) * (1.0 if n > 1 else 0.0)
    alpha = np.where(typ == 1, 1.0, 0.5)
    sysm = P.System(x, q, typ, np.zeros(n, dtype=np.int32), alpha, [0, 0, 0], [L, L, L], 2)
    kw = dict(fixed_iteration=1, max_iterations=4, damp_type="exponential", polar_gs_ranked=0)
    for polar_cut in (0.0, 12.0):
        st = H.fluid_style(sysm, 2.5, 12.0, polar_cut=polar_cut, **kw)
        lit = P.compute(sysm, st)
        s = pb.PairStyle(device=0)
        _fluid_style_on_device(s, sysm, st.g_ewald, "polar_gs_ranked no fixed_iteration yes max_iterations 4 damp_type exponential"
                               + (f" polar_cutoff {polar_cut}" if polar_cut else ""))
        res, mu, ef, f = run_system(s, sysm)
        s.close()
        scale = max(np.abs(lit["mu"]).max(), 1e-30)
        assert np.abs(mu - lit["mu"]).max() <= 1e-10 * scale
        assert np.abs(ef - lit["ef_static"]).max() <= 1e-10 * max(np.abs(lit["ef_static"]).max(), 1e-30)
        assert np.abs(f - lit["f"]).max() <= 1e-10 * max(np.abs(lit["f"]).max(), 1e-30)
        assert abs(res.eng_pol - lit["eng_pol"]) <= 1e-10 * max(abs(lit["eng_pol"]), 1e-30)
        assert abs(res.eng_coul - lit["eng_coul"]) <= 1e-10 * max(abs(lit["eng_coul"]), 1e-30)


@pytest.mark.parametrize("case", ["h2_default_step0", "methane_default_step0"])
def test_blocked_gauss_seidel_equals_atom_by_atom(case):
    """exact-mode Gauss-Seidel: the blocked forward substitution against the one-atom-at-a-time kernel (same operands,
    different order of additions inside a field sum): same iteration count, dipoles to rounding."""
    fx = H.load_fixture(case)
    outs = []
    for blocked in (0, 1):
        s = pb.PairStyle(device=0)
        configure_from_fixture(s, fx)
        s.set_option("gs_blocked", blocked)
        outs.append(run_fixture(s, fx))
        s.close()
    (ra, mua, efa, fa), (rb, mub, efb, fb) = outs
    assert ra.iterations == rb.iterations == int(fx["iterations"]) or abs(rb.iterations - int(fx["iterations"])) <= 1
    assert H.rel_err(mua, mub) < 1e-9 and H.rel_err(fa, fb) < 1e-9
    assert abs(ra.eng_pol - rb.eng_pol) < 1e-10 * abs(ra.eng_pol)


def test_interleaved_colouring_matches_oracle_and_converges_like_sequential(style):
    """per-atom interleaved colouring (`gs_chunks -8`: chunk c = ranked positions c, c+8, ...; the list-mode default of
    round 1, still the fallback without pair groups): against the oracle's emulation of the same colouring and against
    the strictly sequential sweep: same fixed point, and nearly the sequential iteration count (contiguous chunks need
    ~5x more on this system)."""
    sysm = H.water_box(10)
    kw = dict(polar_cut=12.0, damp_type="exponential", polar_gs_ranked=1, precision=1e-11, max_iterations=200,
              polar_gamma=1.03)
    ref = P.polar_rows(sysm, H.water_style(sysm, 2.5, 12.0, gs_chunks=-8, **kw))
    seq = P.polar_rows(sysm, H.water_style(sysm, 2.5, 12.0, gs_chunks=0, **kw))
    st = H.water_style(sysm, 2.5, 12.0, **kw)
    _water_on_device(style, sysm, st, "damp_type exponential precision 1e-11 max_iterations 200 polar_gamma 1.03 "
                                       "polar_cutoff 12.0 gs_chunks -8")
    res, mu, ef, f = run_system(style, sysm)
    assert abs(res.iterations - ref["iterations"]) <= 2
    assert res.iterations <= seq["iterations"] + 4
    assert np.abs(mu - ref["mu"]).max() < 20 * 1e-11 and np.abs(mu - seq["mu"]).max() < 100 * 1e-11
    assert abs(res.eng_pol - seq["eng_pol"]) < 1e-8 * abs(seq["eng_pol"])


def _colouring(style, n):
    raw = style.debug_fetch("gs_colouring", np.int32, 2 * n + 3)
    colour, after, (ncol, rounds, ngroups) = raw[:n], raw[n:2 * n], raw[2 * n:2 * n + 3]
    return colour, after, int(ncol), int(rounds), int(ngroups)


@pytest.mark.parametrize("system", ["water", "fluid"])
def test_group_coloured_gauss_seidel_is_iteration_for_iteration_the_oracle(style, system):
    """DEFAULT list-mode Gauss-Seidel (polar_gs_ranked yes, precision mode): the group-coloured sweep on the TMA pair-group
    kernel with the device-side stop flag.  The oracle replays the colouring the device chose (colour of every atom +
    in-group order): dipoles after 1, 2, 3 sweeps to 1e-10 relative, the converged run with the SAME iteration count and
    the dipoles to 1e-10; and against the reference's own order (strictly sequential ranked sweep): same fixed point
    within 20*precision (BASELINE.json's tolerance for the GS modes), no more iterations than the sequential sweep + 2."""
    if system == "water":
        sysm = H.water_box(10)
        mk, on_device = H.water_style, _water_on_device
    else:
        sysm = H.lj_charge_fluid(9)
        mk = H.fluid_style
        on_device = lambda s, sy, st, words: _fluid_style_on_device(s, sy, st.g_ewald, words)
    n = sysm.n
    kw = dict(polar_cut=12.0, damp_type="exponential", polar_gs_ranked=1, precision=1e-11, max_iterations=200,
              polar_gamma=1.03)
    st = mk(sysm, 2.5, 12.0, **kw)
    words = "damp_type exponential precision 1e-11 max_iterations 200 polar_gamma 1.03 polar_cutoff 12.0"
    on_device(style, sysm, st, words)
    res, mu, ef, f = run_system(style, sysm)
    assert not (res.status & pb.STATUS_DIVERGED)
    colour, after, ncol, rounds, ngroups = _colouring(style, n)
    # the colouring is a partition into ncol colours; second members follow a first member of the same colour
    assert colour.min() >= 0 and colour.max() < ncol and (n + 1) // 2 <= ngroups <= n
    second = after >= 0
    assert np.all(colour[second] == colour[after[second]]) and np.all(after[after[second]] == -1)
    ref = P.polar_rows(sysm, st, colouring=(colour, after, ncol))
    assert res.iterations == ref["iterations"], (res.iterations, ref["iterations"])
    assert H.rel_err(mu, ref["mu"]) < TOL
    assert abs(res.eng_pol - ref["eng_pol"]) < TOL * abs(ref["eng_pol"])
    seq = P.polar_rows(sysm, mk(sysm, 2.5, 12.0, gs_chunks=0, **kw))
    assert res.iterations <= seq["iterations"] + 2, (res.iterations, seq["iterations"])
    assert np.abs(mu - seq["mu"]).max() < 20 * 1e-11
    assert abs(res.eng_pol - seq["eng_pol"]) < 1e-8 * abs(seq["eng_pol"])
    # per-sweep parity: fixed number of sweeps (Gauss-Seidel keeps the extra sweep)
    for nsweeps in (1, 3):
        s2 = pb.PairStyle(device=0)
        on_device(s2, sysm, st, f"damp_type exponential fixed_iteration yes max_iterations {nsweeps} polar_gamma 1.03 polar_cutoff 12.0")
        r2, mu2, _, _ = run_system(s2, sysm)
        c2, a2, nc2, _, _ = _colouring(s2, n)
        s2.close()
        assert np.array_equal(c2, colour) and np.array_equal(a2, after)  # the colouring does not depend on the solver keywords
        kw2 = dict(kw, fixed_iteration=1, max_iterations=nsweeps)
        ref2 = P.polar_rows(sysm, mk(sysm, 2.5, 12.0, **kw2), colouring=(colour, after, ncol))
        assert r2.iterations == ref2["iterations"] == nsweeps
        assert H.rel_err(mu2, ref2["mu"]) < TOL


def test_mof_supercell_list_mode_matches_oracle(style):
    """BASELINE config 4's shape at reduced size: the reference's MOF-5 + CO2 example cell replicated 2 x 2 x 2 (7392 atoms,
    10 atom types, bond topology = special lists with special_lj/coul weights, 808 molecules), polar_cutoff = cut_coul,
    the reference's default solver (polar_gs_ranked yes, precision mode) on the group-coloured sweep.  LJ + Coulomb (special
    bonds, tables) against the literal oracle, polarization (static field, dipoles, forces, energy) against the row oracle
    replaying the device's colouring; fixed point against the strictly sequential ranked sweep."""
    from gpu_common import c
    W = H._workloads()
    fx = H.load_fixture("co2_singlepoint_step0")
    w, cut, coeff = W.mof_supercell(H.GOLDEN / "co2_singlepoint_step0.npz", 2)
    assert w.n == 7392 and int(w.molecule.max()) == 808
    sysm = P.System(w.x, w.q, w.type, w.molecule, w.alpha, w.boxlo, w.boxhi, w.ntypes, tag=w.tag, nspecial=w.nspecial,
                    special=w.special)
    kw = dict(polar_cut=cut, damp_type="exponential", damp=2.1304, polar_gs_ranked=1, precision=1e-11, max_iterations=200,
              polar_gamma=1.03, zodid=0, fixed_iteration=0, use_previous=0)
    st = H.style_from_fixture(fx, **kw)
    configure_from_fixture(style, fx, extra_words=[f"polar_cutoff {cut}", "precision 1e-11", "max_iterations 200",
                                                   "use_previous no"])
    style.set_box(sysm.boxlo, sysm.boxhi)
    res, mu, ef, f = run_system(style, sysm)
    assert not (res.status & (pb.STATUS_DIVERGED | pb.STATUS_EXACT))
    n = sysm.n
    colour, after, ncol, rounds, ngroups = _colouring(style, n)
    ref = P.polar_rows(sysm, st, colouring=(colour, after, ncol))
    assert res.iterations == ref["iterations"], (res.iterations, ref["iterations"])
    assert H.rel_err(ef, ref["ef_static"]) < TOL and H.rel_err(mu, ref["mu"]) < TOL
    assert abs(res.eng_pol - ref["eng_pol"]) < TOL * abs(ref["eng_pol"])
    # LJ + real-space Coulomb with the bond topology: literal oracle with the dipoles switched off
    lit = P.compute(sysm, H.style_from_fixture(fx, **dict(kw, polar_gs_ranked=0, zodid=1, polar_gamma=0.0)))
    assert abs(res.eng_vdwl - lit["eng_vdwl"]) < TOL * abs(lit["eng_vdwl"])
    assert abs(res.eng_coul - lit["eng_coul"]) < TOL * abs(lit["eng_coul"])
    ftot = lit["f"] + ref["f"]
    assert np.abs(f - ftot).max() < TOL * np.abs(ftot).max()
    # the reference's own order (strictly sequential ranked sweep): same fixed point within the GS tolerance
    seq = P.polar_rows(sysm, st)
    assert np.abs(mu - seq["mu"]).max() < 20 * 1e-11
    assert res.iterations <= seq["iterations"] + 2, (res.iterations, seq["iterations"])
    assert abs(res.rmin - seq["rmin"]) < 1e-12 * seq["rmin"]
    # second step on the same lists (no rebuild): rmin now comes out of the group cache, the colouring is kept
    res2, mu2, _, f2 = run_system(style, sysm, ago=1)
    assert res2.rmin == res.rmin and res2.iterations == res.iterations and np.array_equal(mu2, mu)


def test_group_coloured_sweep_stop_flag_and_lag_do_not_change_the_result():
    """the host looks at the device's stop flag one iteration late (the iteration enqueued meanwhile is skipped by every
    kernel): identical dipoles, iteration counts and energies with and without the lag, in Jacobi and Gauss-Seidel
    precision modes, and over repeated steps with use_previous."""
    sysm = H.water_box(8)
    outs = {}
    for words in ("polar_gs_ranked no precision 1e-9 max_iterations 300", "precision 1e-11 max_iterations 200 polar_gamma 1.03",
                  "polar_gs_ranked no precision 1e-9 max_iterations 5"):
        for lag in (0, 1):
            s = pb.PairStyle(device=0)
            _water_on_device(s, sysm, H.water_style(sysm, 2.5, 12.0), words + " damp_type exponential polar_cutoff 12.0 use_previous yes")
            s.set_option("scf_lag", lag)
            res, mu, ef, f = run_system(s, sysm)
            res2, mu2, _, f2 = run_system(s, sysm, mu_in=mu, ago=1)
            s.close()
            key = words
            if key in outs:
                r0, m0, r20, m20, f20 = outs[key]
                assert res.iterations == r0.iterations and res2.iterations == r20.iterations
                assert (res.status & pb.STATUS_DIVERGED) == (r0.status & pb.STATUS_DIVERGED)
                assert np.array_equal(mu, m0) and np.array_equal(mu2, m20) and np.array_equal(f2, f20)
                assert res.eng_pol == r0.eng_pol
            else:
                outs[key] = (res, mu, res2, mu2, f2)
    assert outs["polar_gs_ranked no precision 1e-9 max_iterations 5"][0].status & pb.STATUS_DIVERGED


@pytest.mark.parametrize("ruleset", ["mixed", "include", "many"])
def test_exclusions_in_list_mode_match_oracle(style, ruleset):
    """neigh_modify exclude / include with polar_cutoff (neighbor-list polarization): LJ / Coulomb drop the excluded pairs,
    static field, dipoles and dipole forces keep them -- type rule + group rule + molecule/intra rule on random group
    masks; `neigh_modify include g` (pairs of two atoms of g only) with a molecule/inter rule; 14 rules at once (the
    device holds up to 32)"""
    sysm = H.lj_charge_fluid(10)                       # 4000 atoms, L = 34.2 A
    rng = np.random.default_rng(17)
    mask = (1 | (2 * (rng.random(sysm.n) < 0.3)) | (4 * (rng.random(sysm.n) < 0.3)) | (8 * (rng.random(sysm.n) < 0.2))).astype(np.int32)
    rules = [("type", 1, 1), ("group", 2, 4), ("molecule/intra", 8)]
    if ruleset == "include":
        mask = (1 | (2 * (rng.random(sysm.n) < 0.7)) | (4 * (rng.random(sysm.n) < 0.5))).astype(np.int32)
        rules = [("include", 2), ("molecule/inter", 4)]
        sysm.molecule[:] = 1 + np.arange(sysm.n) // 7      # (molecule ids give molecule/inter something to separate)
    elif ruleset == "many":
        mask = np.ones(sysm.n, dtype=np.int32)
        for b in range(1, 15):
            mask |= ((rng.random(sysm.n) < 0.15) << b).astype(np.int32)
        rules = [("group", 1 << b, 1 << (b + 1)) for b in range(1, 13)] + [("type", 2, 2), ("molecule/intra", 1 << 14)]
        assert len(rules) == 14
    kw = dict(fixed_iteration=1, max_iterations=5, damp_type="exponential", polar_gs_ranked=0)
    st = H.fluid_style(sysm, 2.5, 12.0, polar_cut=12.0, **kw)
    ref = P.polar_rows(sysm, st)
    st0 = H.fluid_style(sysm, 2.5, 12.0, polar_gs_ranked=0, zodid=1, polar_gamma=0.0)
    xall, owner, shift = P.build_ghosts(sysm, st0.cutneighmax)
    nn, first, neigh = P.build_half_list(sysm, st0, xall, owner)
    full = P.compute(sysm, st0, lists=(xall, owner, shift, nn, first, neigh))
    lit = P.compute(sysm, st0, lists=P.apply_exclusions(sysm, (xall, owner, shift, nn, first, neigh), rules, mask))
    assert abs(lit["eng_coul"] - full["eng_coul"]) > 1e-3 * abs(full["eng_coul"])   # the rules bite
    style.set_ntypes(2)
    style.command("pair_style lj/cut/coul/long/polarization 2.5 12.0 polar_gs_ranked no fixed_iteration yes "
                  "max_iterations 5 damp_type exponential polar_cutoff 12.0")
    style.command("pair_coeff 1 1 0.1 3.0")
    style.command("pair_coeff 2 2 0.1 3.0")
    style.init(g_ewald=st.g_ewald, molecular=0)
    style.set_box(sysm.boxlo, sysm.boxhi)
    style.set_exclusions(rules)
    n = sysm.n
    mu, f, ef = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
    c = np.ascontiguousarray
    res = style.compute(c(sysm.x), c(sysm.q), c(sysm.type), c(sysm.alpha), mu, f, molecule=c(sysm.molecule), tag=c(sysm.tag),
                        ef_static=ef, mask=mask)
    assert not (res.status & pb.STATUS_EXACT)
    assert H.rel_err(ef, ref["ef_static"]) < TOL and H.rel_err(mu, ref["mu"]) < TOL
    assert abs(res.eng_pol - ref["eng_pol"]) < TOL * abs(ref["eng_pol"])
    assert abs(res.eng_vdwl - lit["eng_vdwl"]) < TOL * abs(lit["eng_vdwl"])
    assert abs(res.eng_coul - lit["eng_coul"]) < TOL * abs(lit["eng_coul"])
    ftot = lit["f"] + ref["f"]
    assert np.abs(f - ftot).max() < TOL * np.abs(ftot).max()
    # clearing the rules restores the full interaction
    style.set_exclusions([])
    mu[:], f[:] = 0.0, 0.0
    res2 = style.compute(c(sysm.x), c(sysm.q), c(sysm.type), c(sysm.alpha), mu, f, molecule=c(sysm.molecule), tag=c(sysm.tag),
                         ef_static=ef, mask=mask)
    assert abs(res2.eng_coul - full["eng_coul"]) < TOL * abs(full["eng_coul"])
