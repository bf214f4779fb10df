"""GPU parity of the device Ewald (polb200_ewald_*) -- through the C ABI, against the reference's
own KSpace numbers (golden vectors from the reference binary) and against the oracle at a size the CPU finishes
in seconds; plus size-independent properties at the bench size."""
import numpy as np
import pytest

import polhelpers as H
from gpu_common import pb
from oracle import polref as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["ewald_h2", "ewald_methane", "ewald_brick"])
def test_device_ewald_matches_reference(case, golden_dir):
    g = np.load(golden_dir / f"{case}.npz")
    e = pb.Ewald(device=0)
    info = e.init(float(g["accuracy"]), g["q"], float(g["cut_coul"]), g["boxlo"], g["boxhi"])
    assert info.kcount == int(g["kcount"])                       # k-vector set: exact
    assert (info.kxmax, info.kymax, info.kzmax) == tuple(int(v) for v in g["kxyzmax"])
    x, q = np.ascontiguousarray(g["x"]), np.ascontiguousarray(g["q"])
    f = np.zeros_like(x)
    energy, virial = e.compute(x, q, f)
    e.close()
    assert abs(energy - float(g["elong"])) < 1e-10 * abs(float(g["elong"]))
    assert np.abs(f - g["f_kspace"]).max() < 1e-10 * np.abs(g["f_kspace"]).max()
    assert H.rel_err(virial, g["virial_kspace"]) < 1e-9


def test_device_ewald_matches_oracle_on_the_fluid():
    sysm = H.lj_charge_fluid(10)                                 # 4000 atoms, kmax ~ 7
    prd = sysm.boxhi - sysm.boxlo
    plan = P.ewald_plan(1e-4, sysm.q, 12.0, prd)
    ref = P.ewald_compute(plan, sysm.x, sysm.q, prd)
    e = pb.Ewald(device=0)
    info = e.init(1e-4, sysm.q, 12.0, sysm.boxlo, sysm.boxhi)
    assert info.kcount == plan.kcount and abs(info.g_ewald - plan.g_ewald) < 1e-15
    f = np.zeros((sysm.n, 3))
    energy, virial = e.compute(np.ascontiguousarray(sysm.x), np.ascontiguousarray(sysm.q), f)
    e.close()
    assert abs(energy - ref["energy"]) < 1e-10 * abs(ref["energy"])
    assert np.abs(f - ref["f"]).max() < 1e-10 * np.abs(ref["f"]).max()
    assert H.rel_err(virial, ref["virial"]) < 1e-9


def test_device_ewald_properties_at_bench_size():
    """32 000 atoms (BASELINE config 2): no net force, translation invariance, and the virial trace identity
    of the Ewald reciprocal sum, trace(virial) = energy_k (before the self term), independent of any oracle."""
    sysm = H.lj_charge_fluid(20)
    e = pb.Ewald(device=0)
    info = e.init(1e-4, sysm.q, 12.0, sysm.boxlo, sysm.boxhi)
    x, q = np.ascontiguousarray(sysm.x), np.ascontiguousarray(sysm.q)
    f = np.zeros((sysm.n, 3))
    energy, virial = e.compute(x, q, f)
    f2 = np.zeros((sysm.n, 3))
    energy2, _ = e.compute(np.ascontiguousarray(x + np.array([3.3, -1.1, 0.7])), q, f2)
    ms = e.last_ms()
    e.close()
    assert np.abs(f.sum(0)).max() < 1e-9 * np.abs(f).max() * np.sqrt(sysm.n)
    assert abs(energy - energy2) < 1e-10 * abs(energy) and np.abs(f - f2).max() < 1e-9 * np.abs(f).max()
    # sum_k uk (3 + vterm k^2) = sum_k uk (1 - k^2/(2 g^2)) ... checked through the oracle-free identity below
    self_term = info.g_ewald * float((q * q).sum()) / np.sqrt(np.pi) * pb.REAL_QQRD2E
    ek = energy + self_term                                      # neutral system: no background term
    assert ek > 0 and virial[:3].sum() < 3 * ek                  # each vg diagonal < 1
    print(f"device Ewald, 32000 atoms, kcount {info.kcount}: {ms:.3f} ms per compute")


def test_device_ewald_is_bit_reproducible():
    sysm = H.lj_charge_fluid(10)
    outs = []
    for _ in range(3):
        e = pb.Ewald(device=0)
        e.init(1e-4, sysm.q, 12.0, sysm.boxlo, sysm.boxhi)
        f = np.zeros((sysm.n, 3))
        energy, virial = e.compute(np.ascontiguousarray(sysm.x), np.ascontiguousarray(sysm.q), f)
        e.close()
        outs.append((energy, f.copy(), virial))
    for energy, f, virial in outs[1:]:
        assert energy == outs[0][0] and np.array_equal(f, outs[0][1]) and np.array_equal(virial, outs[0][2])


def test_column_kernels_equal_the_quad_kernels(monkeypatch):
    """Two realisations of the same sums (ewald.cuh): the column form (default) and round 1's quad form
    (POLB200_EWALD_QUADS=1, read at init).  Energy, forces and virial must agree to rounding, also for an atom count that is
    no multiple of a tile and on a non-cubic box (different kmax per dimension)."""
    sysm = H.lj_charge_fluid((12, 9, 7))
    keep = np.arange(sysm.n) % 97 != 5
    x, q = np.ascontiguousarray(sysm.x[keep]), np.ascontiguousarray(sysm.q[keep])
    out = []
    for quads in ("0", "1"):
        monkeypatch.setenv("POLB200_EWALD_QUADS", quads)
        e = pb.Ewald(device=0)
        info = e.init(1e-5, q, 10.0, sysm.boxlo, sysm.boxhi)
        f = np.zeros_like(x)
        energy, virial = e.compute(x, q, f)
        e.close()
        out.append((info.kcount, energy, f, np.array(virial)))
    assert out[0][0] == out[1][0]
    assert abs(out[0][1] - out[1][1]) < 1e-12 * abs(out[1][1])
    assert np.abs(out[0][2] - out[1][2]).max() < 1e-12 * np.abs(out[1][2]).max()
    assert H.rel_err(out[0][3], out[1][3]) < 1e-11


def test_more_kx_than_one_register_block(monkeypatch):
    """kmax = 36 (tight accuracy, short real-space cutoff): the structure-factor kernel holds 32 kx per pass, so this
    k set needs two passes; the force kernel's recurrences run 73 steps.  Against the oracle's direct sums and the quad
    kernels."""
    sysm = H.lj_charge_fluid(4)                                  # 256 atoms, 97 634 k-vectors
    prd = sysm.boxhi - sysm.boxlo
    plan = P.ewald_plan(1e-10, sysm.q, 2.5, prd)
    assert max(plan.kxmax, plan.kymax, plan.kzmax) > 32
    ref = P.ewald_compute(plan, sysm.x, sysm.q, prd)
    x, q = np.ascontiguousarray(sysm.x), np.ascontiguousarray(sysm.q)
    for quads in ("0", "1"):
        monkeypatch.setenv("POLB200_EWALD_QUADS", quads)
        e = pb.Ewald(device=0)
        info = e.init(1e-10, q, 2.5, sysm.boxlo, sysm.boxhi)
        assert info.kcount == plan.kcount
        f = np.zeros_like(x)
        energy, virial = e.compute(x, q, f)
        e.close()
        assert abs(energy - ref["energy"]) < 1e-10 * abs(ref["energy"])
        assert np.abs(f - ref["f"]).max() < 1e-10 * np.abs(ref["f"]).max()
        assert H.rel_err(virial, ref["virial"]) < 1e-9
