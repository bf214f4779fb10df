"""GPU parity of the device PPPM (polb200_pppm_*, SURVEY §8f rank 1, second half) -- through the C ABI, against the
reference's own KSpace numbers (golden vectors from the reference binary) and against the oracle; plus the agreement
of PPPM with the device Ewald sum on total (real + reciprocal is not available here, so: reciprocal parts at equal
g_ewald) at a size where Ewald is exact enough to serve as the yardstick."""
import numpy as np
import pytest

import polhelpers as H
from gpu_common import pb
from oracle import pppmref as PP
from test_pppm_oracle import CASES, plan_for

pytestmark = pytest.mark.gpu


def device_kwargs(g):
    km = str(g["kspace_modify"]).split()
    kw = {}
    if "order" in km:
        kw["order"] = int(km[km.index("order") + 1])
    if "mesh" in km:
        i = km.index("mesh")
        kw["mesh"] = [int(v) for v in km[i + 1:i + 4]]
    if "gewald" in km:
        kw["g_ewald"] = float(km[km.index("gewald") + 1])
    return kw


@pytest.mark.parametrize("case", CASES)
def test_device_pppm_matches_reference(case, golden_dir):
    g = np.load(golden_dir / f"{case}.npz")
    p = pb.PPPM(device=0)
    info = p.init(float(g["accuracy"]), g["q"], float(g["cut_coul"]), g["boxlo"], g["boxhi"], **device_kwargs(g))
    assert (info.nx, info.ny, info.nz) == tuple(int(v) for v in g["grid"]) and info.order == int(g["order"])
    assert f"{info.g_ewald:g}" == f"{float(g['g_ewald_printed']):g}"
    ref_plan = plan_for(g)
    assert abs(info.g_ewald - ref_plan.g_ewald) < 1e-14
    x, q = np.ascontiguousarray(g["x"]), np.ascontiguousarray(g["q"])
    f = np.zeros_like(x)
    energy, virial = p.compute(x, q, f)
    p.close()
    assert abs(energy - float(g["elong"])) < 1e-10 * abs(float(g["elong"]))
    assert np.abs(f - g["f_kspace"]).max() < 1e-9 * np.abs(g["f_kspace"]).max()
    assert H.rel_err(virial, g["virial_kspace"]) < 1e-8


def test_device_pppm_matches_oracle_on_the_fluid():
    sysm = H.lj_charge_fluid(10)                                 # 4000 atoms
    prd = sysm.boxhi - sysm.boxlo
    plan = PP.PPPMPlan(1e-4, sysm.q, 12.0, prd)
    ref = plan.compute(sysm.x, sysm.q, sysm.boxlo)
    p = pb.PPPM(device=0)
    info = p.init(1e-4, sysm.q, 12.0, sysm.boxlo, sysm.boxhi)
    assert (info.nx, info.ny, info.nz) == tuple(int(v) for v in plan.n) and abs(info.g_ewald - plan.g_ewald) < 1e-14
    f = np.zeros((sysm.n, 3))
    energy, virial = p.compute(np.ascontiguousarray(sysm.x), np.ascontiguousarray(sysm.q), f)
    assert abs(energy - ref["energy"]) < 1e-10 * abs(ref["energy"])
    assert np.abs(f - ref["f"]).max() < 1e-9 * np.abs(ref["f"]).max()
    assert H.rel_err(virial, ref["virial"]) < 1e-8
    # device pointers = host buffers
    import torch
    xt, qt = torch.from_numpy(np.ascontiguousarray(sysm.x)).cuda(), torch.from_numpy(np.ascontiguousarray(sysm.q)).cuda()
    ft = torch.zeros((sysm.n, 3), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    e2, v2 = p.compute_device(sysm.n, xt.data_ptr(), qt.data_ptr(), ft.data_ptr())
    assert abs(e2 - energy) < 1e-12 * abs(energy) and np.abs(ft.cpu().numpy() - f).max() < 1e-11 * np.abs(f).max()
    print(f"device PPPM: {sysm.n} atoms, grid {info.nx}x{info.ny}x{info.nz}: {p.last_ms():.3f} ms")
    p.close()


def test_device_pppm_agrees_with_device_ewald_at_equal_g():
    """independent of the reference and the oracle: with the same g_ewald, PPPM (1e-6) and the exact Ewald sum (1e-8)
    compute the same reciprocal-space quantity"""
    sysm = H.lj_charge_fluid(8)                                  # 2048 atoms
    x, q = np.ascontiguousarray(sysm.x), np.ascontiguousarray(sysm.q)
    p = pb.PPPM(device=0)
    info = p.init(1e-6, q, 10.0, sysm.boxlo, sysm.boxhi)
    fp = np.zeros((sysm.n, 3))
    ep, _ = p.compute(x, q, fp)
    e = pb.Ewald(device=0)
    e.init(1e-8, q, 10.0, sysm.boxlo, sysm.boxhi, g_ewald=info.g_ewald)
    fe = np.zeros((sysm.n, 3))
    ee, _ = e.compute(x, q, fe)
    assert abs(ep - ee) < 1e-3 * abs(ee)
    assert np.abs(fp - fe).max() < 1e-3 * np.abs(fe).max()
    p.close(), e.close()


def test_charge_assignment_is_bit_reproducible(monkeypatch):
    """make_rho without atomics (atoms sorted by grid cell, one thread per grid point, fixed summation order): repeated
    computes give bit-identical forces, energy and virial -- and the same numbers as the first version (order^3 FP64
    atomicAdd per atom, POLB200_PPPM_ATOMICS=1) to rounding"""
    sysm = H.water_box(16)                                       # 12288 atoms, several atoms per grid cell
    x, q = np.ascontiguousarray(sysm.x), np.ascontiguousarray(sysm.q)
    outs = []
    for atomics in ("0", "0", "1"):
        monkeypatch.setenv("POLB200_PPPM_ATOMICS", atomics)
        p = pb.PPPM(device=0)
        p.init(1e-5, q, 10.0, sysm.boxlo, sysm.boxhi)
        runs = []
        for _ in range(3):
            f = np.zeros((sysm.n, 3))
            e, v = p.compute(x, q, f)
            runs.append((e, np.array(v), f))
        p.close()
        outs.append(runs)
    for runs in outs[:2]:
        for e, v, f in runs[1:]:
            assert e == runs[0][0] and np.array_equal(v, runs[0][1]) and np.array_equal(f, runs[0][2])
    assert outs[0][0][0] == outs[1][0][0] and np.array_equal(outs[0][0][2], outs[1][0][2])      # a second handle: same bits
    e0, v0, f0 = outs[0][0]
    e1, v1, f1 = outs[2][0]
    assert abs(e0 - e1) < 1e-12 * abs(e1) and np.abs(f0 - f1).max() < 1e-11 * np.abs(f1).max()
