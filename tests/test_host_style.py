"""CPU tests of the product's host side through the C ABI (no GPU, no compute calls):
library loads, exports every symbol of include/polb200.h, and settings / coeff / init behave like
PairLJCutCoulLongPolarization::settings/coeff/init_style/init_one (same defaults, same error texts,
same order-dependent validation), with coefficient and Coulomb tables bit-identical to the
reference's own arrays (golden fixture dumped from the reference binary)."""
import importlib.util
import re
import sys
from pathlib import Path

import numpy as np
import pytest

import polhelpers as H

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "lammps-induced-dipole-polarization-pair-style_b200"


def load_pb():
    spec = importlib.util.spec_from_file_location("polb200", PKG / "polb200.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["polb200"] = mod
    spec.loader.exec_module(mod)
    return mod


pb = load_pb()


@pytest.fixture()
def style():
    s = pb.PairStyle(device=pb.DEVICE_NONE)
    yield s
    s.close()


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "polb200.h").read_text()
    declared = set(re.findall(r"\b(polb200_[a-z_0-9]+)\s*\(", header))
    assert declared == set(pb.ABI_SYMBOLS), declared ^ set(pb.ABI_SYMBOLS)
    L = pb.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.polb200_abi_version() == 5


def test_compute_without_device_fails_loudly(style):
    fx = H.load_fixture("h2_default_step0")
    configure_from_fixture(style, fx)
    n = fx["x"].shape[0]
    mu = np.zeros((n, 3))
    f = np.zeros((n, 3))
    with pytest.raises(pb.Polb200Error) as e:
        style.compute(np.ascontiguousarray(fx["x"]), np.ascontiguousarray(fx["q"]),
                      np.ascontiguousarray(fx["type"]), np.ascontiguousarray(fx["alpha"]), mu, f)
    assert e.value.code == pb.ERR_CUDA and "no CPU fallback" in e.value.msg


def configure_from_fixture(style, fx, extra_words=()):
    style.set_ntypes(int(fx["ntypes"]))
    style.command(str(fx["pair_style"]) + " " + " ".join(extra_words))
    for line in str(fx["pair_coeff"]).splitlines():
        style.command(line)
    for line in str(fx["pair_modify"]).splitlines():
        style.command(line)
    style.init(g_ewald=float(fx["g_ewald"]), special_lj=tuple(fx["special_lj"]),
               special_coul=tuple(fx["special_coul"]))
    style.set_box(fx["boxlo"], fx["boxhi"])


def test_defaults_and_keywords(style):
    style.set_ntypes(1)
    style.settings(["2.5", "12.0"])            # defaults: gs_ranked yes (pol.cpp:65-78)
    with pytest.raises(pb.Polb200Error, match="Zodid doesn't work with polar_gs or polar_gs_ranked"):
        style.settings(["2.5", "12.0", "zodid", "yes"])
    style.settings(["2.5", "12.0", "polar_gs_ranked", "no", "zodid", "yes"])
    with pytest.raises(pb.Polb200Error, match="polar_gs and polar_gs_ranked are mutually exclusive"):
        style.settings(["2.5", "12.0", "polar_gs_ranked", "yes", "polar_gs", "yes"])
    for bad in (["2.5", "12.0", "precision"], ["2.5", "12.0", "bogus", "1"], ["2.5", "12.0", "damp_type", "x"],
                ["2.5", "12.0", "fixed_iteration", "maybe"], []):
        with pytest.raises(pb.Polb200Error, match="Illegal pair_style command"):
            style.settings(bad)
    with pytest.raises(pb.Polb200Error, match="Expected floating point parameter"):
        style.settings(["abc"])
    # every reference keyword, incl. the undocumented debug / use_previous all shipped inputs pass
    style.settings("2.5 10.797442 precision 0.00000000001 max_iterations 100 damp_type exponential damp 2.1304 "
                   "polar_gs_ranked yes debug no use_previous yes polar_gamma 1.03 fixed_iteration no".split())
    # single cutoff argument: cut_coul = cut_lj_global (pol.cpp:683)
    style.settings(["9.0"])
    assert style.extract("cut_coul") == (9.0, 0)


def test_coeff_wildcards_and_errors(style):
    style.set_ntypes(3)
    style.settings(["2.5", "10.0"])
    with pytest.raises(pb.Polb200Error, match="Incorrect args for pair coefficients"):
        style.coeff(["1", "1", "0.1"])
    with pytest.raises(pb.Polb200Error, match="Numeric index is out of bounds"):
        style.coeff(["1", "4", "0.1", "3.0"])
    with pytest.raises(pb.Polb200Error, match="Incorrect args for pair coefficients"):
        style.coeff(["2", "1", "0.1", "3.0"])      # j<i only: count == 0 (pol.cpp:799)
    style.coeff(["*", "*", "0.1", "3.0"])
    style.coeff(["2*", "3", "0.2", "3.5", "7.0"])
    eps, dim = style.extract("epsilon")
    assert dim == 2 and eps[1, 1] == 0.1 and eps[2, 3] == 0.2 and eps[3, 3] == 0.2 and eps[1, 3] == 0.1
    with pytest.raises(pb.Polb200Error, match="Pair style requires a KSpace style"):
        style.init(g_ewald=0.2, kspace_present=0)
    with pytest.raises(pb.Polb200Error, match="requires atom attribute polarizability"):
        style.init(g_ewald=0.2, polarizability_flag=0)
    style.init(g_ewald=0.2)
    assert style.init_one(1, 1) == 10.0            # max(cut_lj, cut_coul)
    assert style.init_one(2, 3) == 10.0


def test_all_coeffs_must_be_set(style):
    style.set_ntypes(2)
    style.settings(["2.5", "10.0"])
    style.coeff(["1", "1", "0.1", "3.0"])
    with pytest.raises(pb.Polb200Error, match="All pair coeffs are not set"):
        style.init(g_ewald=0.2)


def test_tables_and_coefficients_bitwise_vs_reference(style):
    fx = H.load_fixture("h2_default_step0")
    configure_from_fixture(style, fx)
    nt = 1 << 12
    for k in ("rtable", "drtable", "ftable", "dftable", "ctable", "dctable", "etable", "detable"):
        got = style.debug_fetch("h_" + k, np.float64, nt)
        assert np.array_equal(got, fx["tab_" + k]), k
    meta = style.debug_fetch("h_tabmeta", np.float64, 4)
    assert int(meta[0]) == int(fx["ncoulmask"]) and int(meta[1]) == int(fx["ncoulshiftbits"])
    assert meta[2] == float(fx["tabinnersq"])
    # coefficient tables vs the oracle restatement (itself pinned to the reference dump)
    st = H.style_from_fixture(fx)
    n1 = int(fx["ntypes"]) + 1
    for k in ("cutsq", "cut_ljsq", "lj1", "lj2", "lj3", "lj4", "offset", "cutneighsq"):
        got = style.debug_fetch("h_" + k, np.float64, n1 * n1).reshape(n1, n1)
        assert np.array_equal(got[1:, 1:], getattr(st, k)[1:, 1:]), k


def test_single_matches_oracle_pair_terms(style):
    fx = H.load_fixture("methane_default_step0")
    configure_from_fixture(style, fx)
    # analytic branch (rsq <= tabinnersq), table branch, beyond cut_lj, special (factor_coul = 0)
    for rsq, fc, fl in ((1.5, 1.0, 1.0), (9.0, 1.0, 1.0), (100.0, 1.0, 1.0), (2.3, 0.0, 0.0), (30.0, 0.0, 0.0)):
        e, f = style.single(1, 2, 1.853, -1.0, rsq, fc, fl)
        assert np.isfinite(e) and np.isfinite(f)
    e1, f1 = style.single(1, 2, 1.853, -1.0, 200.0)   # beyond cut_coul^2 = 164.7 and cut_lj
    assert e1 == 0.0 and f1 == 0.0


def test_restart_roundtrip(style):
    fx = H.load_fixture("h2_default_step0")
    configure_from_fixture(style, fx)
    img = style.write_restart()
    # 7 settings fields (2 doubles, 4 ints, 1 double = 40 bytes) + 6 pairs x (int + 3 doubles)
    assert len(img) == 40 + 6 * 28
    other = pb.PairStyle(device=pb.DEVICE_NONE)
    other.set_ntypes(3)
    other.read_restart(img)
    other.init(g_ewald=float(fx["g_ewald"]))
    a, _ = style.extract("sigma")
    b, _ = other.extract("sigma")
    assert np.array_equal(a, b) and other.extract("cut_coul")[0] == 10.797442
    other.close()


def test_restart_settings_block_is_what_pair_hybrid_exchanges(style):
    """PairHybrid writes / reads only the settings of its sub-styles (src/pair_hybrid.cpp:650,691): the block must be read
    back symmetrically (pol.cpp:976-1009), 40 bytes in the reference's layout, and leave the stream where the reference
    leaves it."""
    fx = H.load_fixture("h2_default_step0")
    configure_from_fixture(style, fx)
    style.pair_modify(["table", "10", "tabinner", "1.9", "shift", "yes", "mix", "arithmetic"])
    img = style.write_restart()
    assert style.restart_settings_size() == 40
    other = pb.PairStyle(device=pb.DEVICE_NONE)
    other.set_ntypes(3)
    stream = img[:40] + b"\x01\x00\x00\x00NEXT-SECTION"      # what follows in a hybrid restart: other records
    assert other.read_restart_settings(stream) == 40            # consumed exactly the reference's 7 fields
    assert other.extract("cut_coul")[0] == 10.797442
    assert other.write_restart()[:40] == img[:40]               # offset / mix / table bits / tabinner came back too
    other.close()


def test_restart_keyword_record_is_opt_in_and_round_trips():
    """`restart_keywords yes` (extension): the polarization keywords travel in an 88-byte record behind the 7 reference
    fields; without the keyword the image is byte-identical to the reference's layout.  The reader recognises the record
    by its magic in either kind of file."""
    words = ["2.5", "11.0", "precision", "1e-9", "polar_gs_ranked", "no", "zodid", "no", "fixed_iteration", "yes", "damp", "1.7",
             "max_iterations", "17", "damp_type", "exponential", "polar_gs", "yes", "polar_gamma", "1.1", "use_previous", "yes",
             "polar_cutoff", "9.5", "gs_chunks", "-4"]
    a = pb.PairStyle(device=pb.DEVICE_NONE)
    a.set_ntypes(2)
    a.settings(words)
    a.coeff(["*", "*", "0.1", "3.0"])
    plain = a.write_restart()
    assert a.restart_settings_size() == 40
    a.close()
    a = pb.PairStyle(device=pb.DEVICE_NONE)   # (the keyword checks are order dependent, like the reference's: fresh handle)
    a.set_ntypes(2)
    a.settings(words + ["restart_keywords", "yes"])
    a.coeff(["*", "*", "0.1", "3.0"])
    ext = a.write_restart()
    assert a.restart_settings_size() == 128 and ext[:40] == plain[:40] and ext[40:48] == b"POLB2KW1" and ext[128:] == plain[40:]
    b = pb.PairStyle(device=pb.DEVICE_NONE)
    b.set_ntypes(2)
    b.settings(["3.0"])                      # defaults: none of the keywords above
    assert b.read_restart_settings(ext + b"trailing bytes of the file") == 128
    b.read_restart(ext)
    assert b.write_restart() == ext          # every keyword (and the opt-in itself) came back
    c = pb.PairStyle(device=pb.DEVICE_NONE)
    c.set_ntypes(2)
    c.settings(["3.0"])
    c.read_restart(plain)                    # a reference-layout file: settings and coefficients only
    assert c.write_restart() == plain[:0] + c.write_restart() and c.write_restart()[:40] == plain[:40]
    assert len(c.write_restart()) == len(plain)
    for s in (a, b, c):
        s.close()


def test_pair_modify(style):
    style.set_ntypes(1)
    style.settings(["2.5", "10.0"])
    style.coeff(["1", "1", "0.1", "3.0"])
    style.pair_modify(["table", "0"])
    style.pair_modify(["mix", "arithmetic", "shift", "yes"])
    with pytest.raises(pb.Polb200Error, match="Illegal pair_modify command"):
        style.pair_modify(["mix", "bogus"])
    style.init(g_ewald=0.25)
    # shift yes: offset = 4 eps ((s/rc)^12 - (s/rc)^6), pol.cpp:877-880
    off = style.debug_fetch("h_offset", np.float64, 4).reshape(2, 2)[1, 1]
    r = 3.0 / 2.5
    assert off == 4.0 * 0.1 * (r ** 12 - r ** 6)


def test_tail_correction(style):
    """pair_modify tail yes: etail_ij / ptail_ij of init_one (pol.cpp:897-918) against the closed-form integrals
    of the LJ energy and virial beyond the cutoff; zero without the keyword; shift+tail is refused like the reference."""
    style.set_ntypes(2)
    style.settings(["2.5", "10.0"])
    style.coeff(["1", "1", "0.1", "3.0"])
    style.coeff(["2", "2", "0.2", "2.0", "6.0"])
    style.init(g_ewald=0.25)
    assert style.tail(1, 1, 100.0, 100.0) == (0.0, 0.0)
    style.pair_modify(["tail", "yes"])
    style.init(g_ewald=0.25)
    for (i, j, eps, sig, rc, ni, nj) in [(1, 1, 0.1, 3.0, 2.5, 120.0, 120.0), (2, 2, 0.2, 2.0, 6.0, 40.0, 40.0),
                                         (1, 2, np.sqrt(0.1 * 0.2), np.sqrt(3.0 * 2.0), np.sqrt(2.5 * 6.0), 120.0, 40.0)]:
        e, p = style.tail(i, j, ni, nj)
        s6, rc3 = sig ** 6, rc ** 3
        e_ref = 8.0 * np.pi * ni * nj * eps * s6 * (s6 - 3.0 * rc3 ** 2) / (9.0 * rc3 ** 3)
        p_ref = 16.0 * np.pi * ni * nj * eps * s6 * (2.0 * s6 - 3.0 * rc3 ** 2) / (9.0 * rc3 ** 3)
        assert abs(e - e_ref) <= 1e-13 * abs(e_ref) and abs(p - p_ref) <= 1e-13 * abs(p_ref)
    style.pair_modify(["shift", "yes"])
    with pytest.raises(pb.Polb200Error, match="Cannot have both pair_modify shift and tail"):
        style.init(g_ewald=0.25)


def test_device_entry_points_fail_loudly_without_a_gpu():
    """No CUDA device => no handle: the product never falls back to a CPU implementation (pair path, KSpace, rigid-body integrator)."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pb.Polb200Error):
        pb.PairStyle(device=0)
    with pytest.raises(pb.Polb200Error):
        pb.Ewald(device=0)
    with pytest.raises(pb.Polb200Error):
        pb.Rigid(device=0)
    with pytest.raises(pb.Polb200Error):
        pb.PPPM(device=0)
