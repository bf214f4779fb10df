"""Shared test helpers: load golden fixtures, build oracle objects from reference-style text."""
import json
from pathlib import Path

import numpy as np

from oracle import polref as P

GOLDEN = Path(__file__).resolve().parent / "golden"


def golden_cases():
    """fixtures of single compute() calls of the pair style (the ewald_* fixtures belong to the KSpace tests)"""
    return sorted(p.stem for p in GOLDEN.glob("*.npz") if not p.stem.startswith(("ewald_", "rigid_", "pppm_", "ago_")))


def load_fixture(name):
    d = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    return {k: d[k] for k in d.files}


def thermo_logs():
    return json.loads((GOLDEN / "thermo_logs.json").read_text())


class StyleError(Exception):
    pass


def parse_pair_style(line):
    """Oracle-side restatement of settings() (src/pair_lj_cut_coul_long_polarization.cpp:678-766):
    returns (cut_lj_global, cut_coul, keyword dict) with the reference's defaults (:65-78)."""
    arg = line.split()
    assert arg[0] == "pair_style" and arg[1] == "lj/cut/coul/long/polarization"
    arg = arg[2:]
    if len(arg) < 1:
        raise StyleError("Illegal pair_style command")
    cut_lj = float(arg[0])
    cut_coul = cut_lj if len(arg) == 1 else float(arg[1])
    kw = dict(precision=1e-11, zodid=0, fixed_iteration=0, damp=2.1304, damp_type="none", max_iterations=50,
              polar_gs=0, polar_gs_ranked=1, polar_gamma=1.03, use_previous=0, debug=0)
    yn = {"yes": 1, "no": 0}
    i = 2
    while i < len(arg):
        if i + 2 > len(arg):
            raise StyleError("Illegal pair_style command")
        k, v = arg[i], arg[i + 1]
        if k == "precision":
            kw["precision"] = float(v)
        elif k == "zodid":
            if kw["polar_gs"] or kw["polar_gs_ranked"]:
                raise StyleError("Zodid doesn't work with polar_gs or polar_gs_ranked")
            kw["zodid"] = yn[v]
        elif k == "fixed_iteration":
            kw["fixed_iteration"] = yn[v]
        elif k == "damp":
            kw["damp"] = float(v)
        elif k == "max_iterations":
            kw["max_iterations"] = int(v)
        elif k == "damp_type":
            if v not in ("exponential", "none"):
                raise StyleError("Illegal pair_style command")
            kw["damp_type"] = v
        elif k == "polar_gs":
            if kw["polar_gs_ranked"]:
                raise StyleError("polar_gs and polar_gs_ranked are mutually exclusive")
            kw["polar_gs"] = yn[v]
        elif k == "polar_gs_ranked":
            if kw["polar_gs"]:
                raise StyleError("polar_gs and polar_gs_ranked are mutually exclusive")
            kw["polar_gs_ranked"] = yn[v]
        elif k == "polar_gamma":
            kw["polar_gamma"] = float(v)
        elif k == "debug":
            kw["debug"] = yn[v]
        elif k == "use_previous":
            kw["use_previous"] = yn[v]
        else:
            raise StyleError("Illegal pair_style command")
        i += 2
    return cut_lj, cut_coul, kw


def system_from_fixture(fx):
    return P.System(fx["x"], fx["q"], fx["type"], fx["molecule"], fx["alpha"], fx["boxlo"], fx["boxhi"],
                    int(fx["ntypes"]), tag=fx["tag"],
                    nspecial=fx["nspecial"] if "nspecial" in fx else None,
                    special=fx["special"] if "special" in fx else None)


def style_from_fixture(fx, **override):
    cut_lj, cut_coul, kw = parse_pair_style(str(fx["pair_style"]))
    kw.pop("debug")
    kw.update(override)
    ncoultablebits = 12
    for l in str(fx["pair_modify"]).splitlines():
        t = l.split()
        if len(t) >= 3 and t[1] == "table":
            ncoultablebits = int(t[2])
    st = P.Style(int(fx["ntypes"]), cut_lj, cut_coul, g_ewald=float(fx["g_ewald"]),
                 special_lj=tuple(fx["special_lj"]), special_coul=tuple(fx["special_coul"]),
                 ncoultablebits=ncoultablebits, **kw)
    for l in str(fx["pair_coeff"]).splitlines():
        t = l.split()
        st.coeff(int(t[1]), int(t[2]), float(t[3]), float(t[4]), float(t[5]) if len(t) > 5 else None)
    st.init()
    return st


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.abs(b).max()), 1e-300)
    return float(np.abs(a - b).max()) / scale


def _workloads():
    import importlib.util
    import sys
    if "polb200_workloads" in sys.modules:
        return sys.modules["polb200_workloads"]
    path = Path(__file__).resolve().parents[1] / "lammps-induced-dipole-polarization-pair-style_b200" / "workloads.py"
    spec = importlib.util.spec_from_file_location("polb200_workloads", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["polb200_workloads"] = mod
    spec.loader.exec_module(mod)
    return mod


def _as_system(w):
    return P.System(w.x, w.q, w.type, w.molecule, w.alpha, w.boxlo, w.boxhi, w.ntypes)


def lj_charge_fluid(ncell, seed=12345, rho=0.1, jitter=0.3):
    """BASELINE config 2 generator (lammps-..._b200/workloads.py) as an oracle System."""
    return _as_system(_workloads().lj_charge_fluid(ncell, seed, rho, jitter))


def water_box(nmol_side, seed=2, rho=0.1):
    """BASELINE config 3 generator (lammps-..._b200/workloads.py) as an oracle System."""
    return _as_system(_workloads().water_box(nmol_side, seed, rho))


def water_style(sysm, cut_lj=2.5, cut_coul=12.0, **kw):
    g = P.ewald_g(1e-4, sysm.q, cut_coul, sysm.boxlo, sysm.boxhi)
    st = P.Style(2, cut_lj, cut_coul, g_ewald=g, **kw)
    st.coeff(1, 1, 0.155, 3.166)
    st.coeff(2, 2, 0.0, 1.0)
    st.init()
    return st


def fluid_style(sysm, cut_lj=2.5, cut_coul=12.0, **kw):
    g = P.ewald_g(1e-4, sysm.q, cut_coul, sysm.boxlo, sysm.boxhi)
    st = P.Style(2, cut_lj, cut_coul, g_ewald=g, **kw)
    st.coeff(1, 1, 0.1, 3.0)
    st.coeff(2, 2, 0.1, 3.0)
    st.init()
    return st


def exclusion_rules(fx):
    """rule tuples of the fixture's `neigh_modify exclude` lines (groups by name -> bitmask)"""
    bits = {str(k): int(v) for k, v in zip(fx["group_names"], fx["group_bits"])}
    return P.parse_exclusions(str(fx["neigh_modify"]), bits)


def reference_lists(fx, sysm, st, case):
    """Ghosts + half list as the reference holds them at this step: built from the positions of the
    last reneighboring (step 0 of the same run: delay 10, src/neighbor.cpp:1923-1937), ghost
    coordinates refreshed from the current owners (CommBrick::forward_comm, src/comm_brick.cpp:463-524)."""
    step = int(fx["step"])
    if step == 0:
        xall, owner, shift = P.build_ghosts(sysm, st.cutneighmax)
        numneigh, first, neigh = P.build_half_list(sysm, st, xall, owner)
        lists = (xall, owner, shift, numneigh, first, neigh)
        if "neigh_modify" in fx:
            lists = P.apply_exclusions(sysm, lists, exclusion_rules(fx), fx["mask"])
        return lists
    fx0 = load_fixture(case.rsplit("_step", 1)[0] + "_step0")
    sys0 = system_from_fixture(fx0)
    xall0, owner, shift = P.build_ghosts(sys0, st.cutneighmax)
    numneigh, first, neigh = P.build_half_list(sys0, st, xall0, owner)
    prd = sysm.boxhi - sysm.boxlo
    xg = sysm.x[owner].copy()
    for d in range(3):
        xg[:, d] = np.where(shift[:, d] != 0, sysm.x[owner, d] + shift[:, d] * prd[d], sysm.x[owner, d])
    xall = np.concatenate([sysm.x, xg])
    return xall, owner, shift, numneigh, first, neigh
