"""Helpers for the -m gpu tests: drive the product through its C ABI from golden fixtures / oracle systems."""
import importlib.util
import sys
from pathlib import Path

import numpy as np

import polhelpers as H

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "lammps-induced-dipole-polarization-pair-style_b200"


def load_pb():
    if "polb200" in sys.modules:
        return sys.modules["polb200"]
    spec = importlib.util.spec_from_file_location("polb200", PKG / "polb200.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["polb200"] = mod
    spec.loader.exec_module(mod)
    return mod


pb = load_pb()


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def configure_from_fixture(style, fx, extra_words=(), **init_kw):
    style.set_ntypes(int(fx["ntypes"]))
    style.command(str(fx["pair_style"]) + " " + " ".join(extra_words))
    for line in str(fx["pair_coeff"]).splitlines():
        style.command(line)
    for line in str(fx["pair_modify"]).splitlines():
        style.command(line)
    style.init(g_ewald=float(fx["g_ewald"]), special_lj=tuple(fx["special_lj"]),
               special_coul=tuple(fx["special_coul"]), **init_kw)
    style.set_box(fx["boxlo"], fx["boxhi"])
    if "neigh_modify" in fx:   # neigh_modify exclude rules of the reference run
        style.set_exclusions(H.exclusion_rules(fx))


def c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def run_fixture(style, fx, ago=0, mu_in=None, peratom=None):
    """One compute() of the product on the fixture's inputs; returns (Result, mu, ef, f).
    peratom: dict that receives the per-atom tallies "eatom" / "vatom" when the fixture's flags ask for them."""
    n = fx["x"].shape[0]
    mu = c(fx["mu_in"] if mu_in is None else mu_in, np.float64).copy()
    f = np.zeros((n, 3))
    ef = np.zeros((n, 3))
    kw = {}
    if peratom is not None:
        if int(fx["eflag"]) // 2:
            kw["eatom"] = peratom["eatom"] = np.zeros(n)
        if int(fx["vflag"]) // 4:
            kw["vatom"] = peratom["vatom"] = np.zeros((n, 6))
    if "mask" in fx:
        kw["mask"] = c(fx["mask"], np.int32)
    res = style.compute(c(fx["x"], np.float64), c(fx["q"], np.float64), c(fx["type"], np.int32),
                        c(fx["alpha"], np.float64), mu, f, molecule=c(fx["molecule"], np.int32),
                        tag=c(fx["tag"], np.int32), ef_static=ef,
                        nspecial=c(fx["nspecial"], np.int32) if "nspecial" in fx else None,
                        special=c(fx["special"], np.int32) if "special" in fx else None,
                        eflag=int(fx["eflag"]), vflag=int(fx["vflag"]), ago=ago, **kw)
    return res, mu, ef, f


def run_system(style, sysm, mu_in=None, eflag=1, vflag=2, ago=0):
    n = sysm.n
    mu = np.zeros((n, 3)) if mu_in is None else c(mu_in, np.float64).copy()
    f = np.zeros((n, 3))
    ef = np.zeros((n, 3))
    res = style.compute(c(sysm.x, np.float64), c(sysm.q, np.float64), c(sysm.type, np.int32),
                        c(sysm.alpha, np.float64), mu, f, molecule=c(sysm.molecule, np.int32),
                        tag=c(sysm.tag, np.int32), ef_static=ef, nspecial=sysm.nspecial, special=sysm.special,
                        eflag=eflag, vflag=vflag, ago=ago)
    return res, mu, ef, f


def canonical_device_list(style, n):
    """Device full list as a set of (i_caller, j_caller, sx, sy, sz, special) tuples packed in int64."""
    res_n = n
    perm = style.debug_fetch("perm", np.int32, res_n)
    rowstart = style.debug_fetch("rowstart", np.uint64, res_n + 1)
    npairs = int(rowstart[-1])
    neigh = style.debug_fetch("neigh", np.int32, npairs)
    # ghosts
    # number of ghosts is not known a priori: fetch generously via xq size = (n+ng)*4 doubles
    return perm, rowstart, neigh


def pack_pairs(i, j, shift, sb):
    return (((i.astype(np.int64) * 100003 + j.astype(np.int64)) * 4 + (shift[:, 0] + 1)) * 4 +
            (shift[:, 1] + 1)) * 16 + (shift[:, 2] + 1) * 4 + sb
