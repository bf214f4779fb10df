"""ctypes bindings for the CPU oracle (oracle/libpolref.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIB = None

DAMP_EXPONENTIAL, DAMP_NONE = 0, 1
MIX_GEOMETRIC, MIX_ARITHMETIC, MIX_SIXTHPOWER = 0, 1, 2

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
lp = C.POINTER(C.c_long)


class Params(C.Structure):
    _fields_ = [
        ("ntypes", C.c_int),
        ("cutsq", dp), ("cut_ljsq", dp), ("lj1", dp), ("lj2", dp), ("lj3", dp), ("lj4", dp), ("offset", dp),
        ("cut_coul", C.c_double), ("g_ewald", C.c_double), ("qqrd2e", C.c_double),
        ("special_lj", C.c_double * 4), ("special_coul", C.c_double * 4),
        ("ncoultablebits", C.c_int), ("ncoulmask", C.c_int), ("ncoulshiftbits", C.c_int),
        ("tabinnersq", C.c_double),
        ("rtable", dp), ("drtable", dp), ("ftable", dp), ("dftable", dp),
        ("ctable", dp), ("dctable", dp), ("etable", dp), ("detable", dp),
        ("iterations_max", C.c_int), ("damping_type", C.c_int), ("zodid", C.c_int),
        ("fixed_iteration", C.c_int), ("polar_gs", C.c_int), ("polar_gs_ranked", C.c_int),
        ("use_previous", C.c_int),
        ("polar_damp", C.c_double), ("polar_precision", C.c_double), ("polar_gamma", C.c_double),
        ("polar_cut", C.c_double), ("gs_chunks", C.c_int),
        ("boxlo", C.c_double * 3), ("boxhi", C.c_double * 3), ("periodic", C.c_int * 3),
        ("eatom", dp), ("vatom", dp),
        ("gs_colour", ip), ("gs_after", ip), ("gs_ncolours", C.c_int),
    ]


class Result(C.Structure):
    _fields_ = [
        ("eng_vdwl", C.c_double), ("eng_coul", C.c_double), ("eng_pol", C.c_double),
        ("virial", C.c_double * 6),
        ("u_self", C.c_double), ("u_ef", C.c_double), ("u_dd", C.c_double),
        ("rmin", C.c_double), ("iterations", C.c_int), ("diverged", C.c_int),
    ]


def build(force=False):
    so = HERE / "libpolref.so"
    src = HERE / "polref.c"
    if force or not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, (HERE / "polref.h").stat().st_mtime):
        subprocess.run(["make", "-C", str(HERE), "libpolref.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(str(so))
        L.polref_ewald_g.restype = C.c_double
        L.polref_bench_rows.restype = C.c_double
        L.polref_build_half_list.restype = C.c_long
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(dp)


def _i(a):
    return a.ctypes.data_as(ip)


def _l(a):
    return a.ctypes.data_as(lp)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


REAL_QQRD2E = 332.06371  # src/update.cpp:157 (units real)


class System:
    """Everything the hot path needs about one configuration, as flat numpy arrays (host)."""

    def __init__(self, x, q, type_, molecule, alpha, boxlo, boxhi, ntypes, tag=None,
                 nspecial=None, special=None, periodic=(1, 1, 1)):
        self.x = f64(x).reshape(-1, 3)
        self.n = self.x.shape[0]
        self.q = f64(q)
        self.type = i32(type_)
        self.molecule = i32(molecule)
        self.alpha = f64(alpha)
        self.boxlo = f64(boxlo)
        self.boxhi = f64(boxhi)
        self.periodic = i32(periodic)
        self.ntypes = int(ntypes)
        self.tag = i32(tag) if tag is not None else np.arange(1, self.n + 1, dtype=np.int32)
        self.nspecial = i32(nspecial).reshape(-1, 3) if nspecial is not None else None
        self.special = i32(special) if special is not None else None
        if self.special is not None:
            self.special = self.special.reshape(self.n, -1)


class Style:
    """Host-side settings/coeff/init state of the pair style (oracle flavour).

    Mirrors PairLJCutCoulLongPolarization::settings/coeff/init_style/init_one
    (src/pair_lj_cut_coul_long_polarization.cpp:678-921) through the C restatement.
    """

    def __init__(self, ntypes, cut_lj_global, cut_coul, *, precision=1e-11, zodid=0, fixed_iteration=0,
                 damp=2.1304, damp_type="none", max_iterations=50, polar_gs=0, polar_gs_ranked=1,
                 polar_gamma=1.03, use_previous=0, polar_cut=0.0, gs_chunks=0, g_ewald=0.0,
                 qqrd2e=REAL_QQRD2E, special_lj=(1.0, 0.0, 0.0, 0.0), special_coul=(1.0, 0.0, 0.0, 0.0),
                 ncoultablebits=12, tabinner=np.sqrt(2.0), mix=MIX_GEOMETRIC, offset_flag=0, skin=2.0):
        self.ntypes = ntypes
        self.cut_lj_global = cut_lj_global
        self.cut_coul = cut_coul
        self.kw = dict(precision=precision, zodid=zodid, fixed_iteration=fixed_iteration, damp=damp,
                       damp_type=damp_type, max_iterations=max_iterations, polar_gs=polar_gs,
                       polar_gs_ranked=polar_gs_ranked, polar_gamma=polar_gamma, use_previous=use_previous,
                       polar_cut=polar_cut, gs_chunks=gs_chunks)
        self.g_ewald = g_ewald
        self.qqrd2e = qqrd2e
        self.special_lj = tuple(special_lj)
        self.special_coul = tuple(special_coul)
        self.ncoultablebits = ncoultablebits
        self.tabinner = float(tabinner)
        self.mix = mix
        self.offset_flag = offset_flag
        self.skin = skin
        n1 = ntypes + 1
        self.epsilon = np.zeros((n1, n1))
        self.sigma = np.zeros((n1, n1))
        self.cut_lj = np.zeros((n1, n1))
        self.setflag = np.zeros((n1, n1), dtype=np.int32)
        self._init = False

    def coeff(self, i, j, eps, sigma, cut_lj=None):
        if j < i:
            i, j = j, i
        self.epsilon[i, j] = eps
        self.sigma[i, j] = sigma
        self.cut_lj[i, j] = self.cut_lj_global if cut_lj is None else cut_lj
        self.setflag[i, j] = 1
        self._init = False

    def init(self):
        L = lib()
        n1 = self.ntypes + 1
        for name in ("cutsq", "cut_ljsq", "lj1", "lj2", "lj3", "lj4", "offset"):
            setattr(self, name, np.zeros((n1, n1)))
        rc = L.polref_init_coeffs(self.ntypes, _d(self.epsilon), _d(self.sigma), _d(self.cut_lj),
                                  _i(self.setflag), self.mix, self.offset_flag, C.c_double(self.cut_coul),
                                  _d(self.cutsq), _d(self.cut_ljsq), _d(self.lj1), _d(self.lj2),
                                  _d(self.lj3), _d(self.lj4), _d(self.offset))
        if rc:
            raise RuntimeError("All pair coeffs are not set")
        self.ncoulmask = self.ncoulshiftbits = 0
        self.tabinnersq = self.tabinner ** 2
        if self.ncoultablebits:
            nt = 1 << self.ncoultablebits
            self.tables = {k: np.zeros(nt) for k in
                           ("rtable", "drtable", "ftable", "dftable", "ctable", "dctable", "etable", "detable")}
            m, s, t = C.c_int(), C.c_int(), C.c_double()
            rc = L.polref_init_tables(C.c_double(self.cut_coul), C.c_double(self.g_ewald),
                                      C.c_double(self.qqrd2e), self.ncoultablebits, C.c_double(self.tabinner),
                                      C.byref(m), C.byref(s), C.byref(t),
                                      *[_d(self.tables[k]) for k in
                                        ("rtable", "drtable", "ftable", "dftable", "ctable", "dctable",
                                         "etable", "detable")])
            if rc:
                raise RuntimeError(f"init_tables failed {rc}")
            self.ncoulmask, self.ncoulshiftbits, self.tabinnersq = m.value, s.value, t.value
        else:
            self.tables = {k: np.zeros(1) for k in
                           ("rtable", "drtable", "ftable", "dftable", "ctable", "dctable", "etable", "detable")}
        self.cutneighsq = np.zeros((n1, n1))
        cuts = np.sqrt(self.cutsq[1:, 1:])
        cn = np.where(cuts > 0.0, cuts + self.skin, 0.0)  # src/neighbor.cpp:301-309
        self.cutneighsq[1:, 1:] = cn * cn
        self.cutneighmax = float(cn.max())
        self._init = True

    def special_flag(self):
        # src/neighbor.cpp:361-382 with a KSpace style present: all 2
        return i32([0, 2, 2, 2])

    def params(self, sysm):
        if not self._init:
            self.init()
        p = Params()
        p.ntypes = self.ntypes
        for name in ("cutsq", "cut_ljsq", "lj1", "lj2", "lj3", "lj4", "offset"):
            setattr(p, name, _d(getattr(self, name)))
        p.cut_coul, p.g_ewald, p.qqrd2e = self.cut_coul, self.g_ewald, self.qqrd2e
        p.special_lj = (C.c_double * 4)(*self.special_lj)
        p.special_coul = (C.c_double * 4)(*self.special_coul)
        p.ncoultablebits, p.ncoulmask, p.ncoulshiftbits = self.ncoultablebits, self.ncoulmask, self.ncoulshiftbits
        p.tabinnersq = self.tabinnersq
        for k, v in self.tables.items():
            setattr(p, k, _d(v))
        kw = self.kw
        p.iterations_max = kw["max_iterations"]
        p.damping_type = DAMP_EXPONENTIAL if kw["damp_type"] == "exponential" else DAMP_NONE
        p.zodid, p.fixed_iteration = kw["zodid"], kw["fixed_iteration"]
        p.polar_gs, p.polar_gs_ranked, p.use_previous = kw["polar_gs"], kw["polar_gs_ranked"], kw["use_previous"]
        p.polar_damp, p.polar_precision, p.polar_gamma = kw["damp"], kw["precision"], kw["polar_gamma"]
        p.polar_cut, p.gs_chunks = kw["polar_cut"], kw["gs_chunks"]
        p.boxlo = (C.c_double * 3)(*sysm.boxlo)
        p.boxhi = (C.c_double * 3)(*sysm.boxhi)
        p.periodic = (C.c_int * 3)(*[int(v) for v in sysm.periodic])
        return p


def ewald_g(accuracy, q, cut_coul, boxlo, boxhi, qqrd2e=REAL_QQRD2E, two_charge_force=REAL_QQRD2E):
    """g_ewald of `kspace_style ewald <accuracy>` (src/KSPACE/ewald.cpp:133-162; real units:
    two_charge_force = qqr2e*qelectron^2/angstrom^2 = 332.06371, src/kspace.cpp:79-81)."""
    prd = f64(boxhi) - f64(boxlo)
    q = f64(q)
    return float(lib().polref_ewald_g(C.c_double(accuracy), C.c_double(qqrd2e), C.c_double(two_charge_force),
                                      C.c_double(float(np.cumsum(q * q)[-1])), C.c_long(len(q)), C.c_double(cut_coul),
                                      C.c_double(prd[0]), C.c_double(prd[1]), C.c_double(prd[2])))


class EwaldPlan(C.Structure):
    _fields_ = [("g_ewald", C.c_double), ("kxmax", C.c_int), ("kymax", C.c_int), ("kzmax", C.c_int), ("kmax", C.c_int),
                ("gsqmx", C.c_double), ("kcount", C.c_int)]


def ewald_plan(accuracy, q, cutoff, prd, g_ewald=0.0, qqrd2e=REAL_QQRD2E, two_charge_force=REAL_QQRD2E):
    """Ewald::init/setup of the reference (g_ewald, kmax per dimension, |k|^2 cutoff, k count)."""
    q = f64(q)
    p = EwaldPlan()
    prd3 = (C.c_double * 3)(*[float(v) for v in prd])
    rc = lib().polref_ewald_plan_make(C.c_double(accuracy), C.c_double(qqrd2e), C.c_double(two_charge_force),
                                      C.c_double(float(np.cumsum(q * q)[-1])), C.c_long(len(q)), C.c_double(cutoff), prd3,
                                      C.c_double(g_ewald), C.byref(p))
    if rc:
        raise RuntimeError("polref_ewald_plan_make failed")
    return p


def ewald_compute(plan, x, q, prd, qqrd2e=REAL_QQRD2E):
    """Reciprocal-space Ewald energy, forces, virial (the reference's KSpace::compute for kspace_style ewald)."""
    x, q = f64(x).reshape(-1, 3), f64(q)
    n = len(q)
    f = np.zeros((n, 3))
    e = C.c_double()
    v = (C.c_double * 6)()
    prd3 = (C.c_double * 3)(*[float(t) for t in prd])
    rc = lib().polref_ewald_compute(C.byref(plan), n, _d(x), _d(q), prd3, C.c_double(qqrd2e), _d(f), C.byref(e), v)
    if rc:
        raise RuntimeError("polref_ewald_compute failed")
    return dict(energy=e.value, f=f, virial=np.array(v[:]))


def build_ghosts(sysm, cutghost):
    L = lib()
    n = sysm.n
    maxghost = max(1024, 40 * n)
    while True:
        xall = np.zeros((n + maxghost, 3))
        owner = np.zeros(maxghost, dtype=np.int32)
        shift = np.zeros((maxghost, 3), dtype=np.int32)
        ng = L.polref_build_ghosts(n, _d(sysm.x), _d(sysm.boxlo), _d(sysm.boxhi), _i(sysm.periodic),
                                   C.c_double(cutghost), maxghost, _d(xall), _i(owner), _i(shift))
        if ng >= 0:
            break
        maxghost = -ng + 16
    return xall[: n + ng].copy(), owner[:ng].copy(), shift[:ng].copy()


def build_half_list(sysm, style, xall, owner):
    """LAMMPS half/bin/newton list over local+ghost atoms; returns (numneigh, firstoffset, neigh)."""
    L = lib()
    if not style._init:
        style.init()
    n = sysm.n
    ng = len(owner)
    allidx = np.concatenate([np.arange(n, dtype=np.int32), owner])
    type_all = i32(sysm.type[allidx])
    tag_all = i32(sysm.tag[allidx])
    cutghost = style.cutneighmax
    maxpairs = 1 << 20
    maxspecial = sysm.special.shape[1] if sysm.special is not None else 0
    while True:
        numneigh = np.zeros(n, dtype=np.int32)
        first = np.zeros(n, dtype=np.int64)
        neigh = np.zeros(maxpairs, dtype=np.int32)
        np_ = L.polref_build_half_list(
            n, ng, _d(f64(xall)), _i(type_all), _i(tag_all), _d(sysm.boxlo), _d(sysm.boxhi),
            _i(sysm.periodic), sysm.ntypes, _d(style.cutneighsq), C.c_double(style.cutneighmax),
            C.c_double(cutghost),
            _i(sysm.nspecial) if sysm.nspecial is not None else None,
            _i(sysm.special) if sysm.special is not None else None, maxspecial,
            _i(style.special_flag()), C.c_long(maxpairs), _i(numneigh), _l(first), _i(neigh))
        if np_ == -2:
            maxpairs *= 4
            continue
        if np_ < 0:
            raise RuntimeError("half list build failed")
        return numneigh, first, neigh[:np_].copy()


def parse_exclusions(text, group_bits):
    """`neigh_modify exclude ...` lines (src/neighbor.cpp:2276-2333) -> rule tuples; group_bits maps group names to bitmasks"""
    rules = []
    for line in str(text).splitlines():
        w = line.split()
        if len(w) < 3 or w[0] != "neigh_modify":
            continue
        i = 1
        while i < len(w):
            if w[i] == "include":   # src/neighbor.cpp:2264-2274: the list is built over the atoms of the group only
                rules.append(("include", group_bits[w[i + 1]]))
                i += 2
                continue
            if w[i] != "exclude":
                raise ValueError("only neigh_modify exclude / include are restated: " + line)
            kind = w[i + 1]
            if kind == "type":
                rules.append(("type", int(w[i + 2]), int(w[i + 3])))
                i += 4
            elif kind == "group":
                rules.append(("group", group_bits[w[i + 2]], group_bits[w[i + 3]]))
                i += 4
            elif kind in ("molecule/intra", "molecule/inter"):
                rules.append((kind, group_bits[w[i + 2]]))
                i += 3
            elif kind == "none":
                rules = []
                i += 2
            else:
                raise ValueError("Illegal neigh_modify command")
    return rules


def excluded_pairs(rules, itype, jtype, imask, jmask, imol, jmol):
    """NPair::exclusion (src/npair.cpp:173-203), vectorised over pairs"""
    ex = np.zeros(len(itype), dtype=bool)
    for r in rules:
        if r[0] == "type":   # ex_type is symmetric (src/neighbor.cpp:448-452)
            ex |= ((itype == r[1]) & (jtype == r[2])) | ((itype == r[2]) & (jtype == r[1]))
        elif r[0] == "group":
            ex |= ((imask & r[1]) != 0) & ((jmask & r[2]) != 0)
            ex |= ((imask & r[2]) != 0) & ((jmask & r[1]) != 0)
        elif r[0] == "molecule/intra":
            ex |= ((imask & r[1]) != 0) & ((jmask & r[1]) != 0) & (imol == jmol)
        elif r[0] == "molecule/inter":
            ex |= ((imask & r[1]) != 0) & ((jmask & r[1]) != 0) & (imol != jmol)
        elif r[0] == "include":   # src/nbin_standard.cpp:209-223 (only group atoms are binned), npair_half_bin_newton.cpp:51
            ex |= ((imask & r[1]) == 0) | ((jmask & r[1]) == 0)
    return ex


def apply_exclusions(sysm, lists, rules, mask):
    """Drop the excluded pairs from a half list: the reference tests `exclusion()` before the distance test while it
    builds the list (src/npair_half_bin_newton.cpp:94), which leaves exactly this list."""
    xall, owner, shift, numneigh, first, neigh = lists
    n = sysm.n
    allidx = np.concatenate([np.arange(n, dtype=np.int64), owner])
    ii = np.repeat(np.arange(n), numneigh)
    jj = allidx[neigh & 0x3FFFFFFF]
    mask = np.asarray(mask)
    ex = excluded_pairs(rules, sysm.type[ii], sysm.type[jj], mask[ii], mask[jj], sysm.molecule[ii], sysm.molecule[jj])
    keep = ~ex
    numneigh2 = np.bincount(ii[keep], minlength=n).astype(np.int32)
    first2 = np.concatenate([[0], np.cumsum(numneigh2)[:-1]]).astype(np.int64)
    return xall, owner, shift, numneigh2, first2, np.ascontiguousarray(neigh[keep])


def compute(sysm, style, mu_in=None, eflag=1, vflag=2, use_matrix=False, trace_max=0, lists=None):
    """Full literal compute() on one configuration.  Returns dict of outputs (forces folded onto owners)."""
    L = lib()
    if not style._init:
        style.init()
    n = sysm.n
    if lists is None:
        xall, owner, shift = build_ghosts(sysm, style.cutneighmax)
        numneigh, first, neigh = build_half_list(sysm, style, xall, owner)
    else:
        xall, owner, shift, numneigh, first, neigh = lists
    ng = len(owner)
    allidx = np.concatenate([np.arange(n, dtype=np.int32), owner])
    p = style.params(sysm)
    mu = np.zeros((n, 3)) if mu_in is None else f64(mu_in).reshape(n, 3).copy()
    ef = np.zeros((n, 3))
    f = np.zeros((n + ng, 3))
    res = Result()
    trace = np.zeros((max(trace_max, 1), n, 3))
    ranked = np.zeros(n, dtype=np.int32)
    q_all, type_all = f64(sysm.q[allidx]), i32(sysm.type[allidx])
    mol_all, alpha_all = i32(sysm.molecule[allidx]), f64(sysm.alpha[allidx])
    eatom = np.zeros(n + ng) if (eflag // 2) else None   # per-atom tallies (eflag_atom / vflag_atom)
    vatom = np.zeros((n + ng, 6)) if (vflag // 4) else None
    if eatom is not None:
        p.eatom = _d(eatom)
    if vatom is not None:
        p.vatom = _d(vatom)
    rc = L.polref_compute(C.byref(p), n, ng, _d(f64(xall)), _d(q_all), _i(type_all), _i(mol_all),
                          _d(alpha_all), n, None, _i(numneigh), _l(first), _i(neigh), _d(mu), _d(ef), _d(f),
                          eflag, vflag, int(use_matrix), C.byref(res), _d(trace), trace_max, _i(ranked))
    if rc:
        raise RuntimeError("polref_compute failed")
    f_owner = f[:n].copy()
    np.add.at(f_owner, owner, f[n:])
    extra = {}
    if eatom is not None:
        e_own = eatom[:n].copy()
        np.add.at(e_own, owner, eatom[n:])
        extra["eatom"] = e_own
    if vatom is not None:
        v_own = vatom[:n].copy()
        np.add.at(v_own, owner, vatom[n:])
        extra["vatom"] = v_own
    return dict(**extra, mu=mu, ef_static=ef, f=f_owner, f_all=f, eng_vdwl=res.eng_vdwl, eng_coul=res.eng_coul,
                eng_pol=res.eng_pol, virial=np.array(res.virial[:]), iterations=res.iterations,
                diverged=res.diverged, u_self=res.u_self, u_ef=res.u_ef, u_dd=res.u_dd, rmin=res.rmin,
                trace=trace[: min(trace_max, res.iterations + 1)], ranked=ranked,
                lists=(xall, owner, shift, numneigh, first, neigh))


def polar_rows(sysm, style, mu_in=None, eflag=1, trace_max=0, nthreads=0, colouring=None):
    """Row-gather polarization part (static field, SCF, dipole forces) with OpenMP.
    colouring = (colour[n], after[n], ncolours): replay the device's group-coloured Gauss-Seidel sweep."""
    L = lib()
    if not style._init:
        style.init()
    n = sysm.n
    p = style.params(sysm)
    if colouring is not None:
        col, aft = i32(colouring[0]), i32(colouring[1])
        assert col.shape == (n,) and aft.shape == (n,)
        p.gs_colour, p.gs_after, p.gs_ncolours = _i(col), _i(aft), int(colouring[2])
    mu = np.zeros((n, 3)) if mu_in is None else f64(mu_in).reshape(n, 3).copy()
    ef = np.zeros((n, 3))
    f = np.zeros((n, 3))
    res = Result()
    trace = np.zeros((max(trace_max, 1), n, 3))
    rc = L.polref_polar_rows(C.byref(p), n, _d(sysm.x), _d(sysm.q), _i(sysm.molecule), _d(sysm.alpha),
                             _d(mu), _d(ef), _d(f), eflag, C.byref(res), _d(trace), trace_max, nthreads)
    if rc:
        raise RuntimeError("polref_polar_rows failed")
    return dict(mu=mu, ef_static=ef, f=f, eng_pol=res.eng_pol, virial=np.array(res.virial[:]),
                iterations=res.iterations, diverged=res.diverged, u_self=res.u_self, u_ef=res.u_ef,
                u_dd=res.u_dd, rmin=res.rmin, trace=trace[: min(trace_max, res.iterations + 1)])


def bench_rows(sysm, style, row0, row1, nsweeps, nthreads=0):
    L = lib()
    p = style.params(sysm)
    cs = C.c_double()
    t = L.polref_bench_rows(C.byref(p), sysm.n, _d(sysm.x), _d(sysm.q), _i(sysm.molecule), _d(sysm.alpha),
                            row0, row1, nsweeps, nthreads, C.byref(cs))
    return float(t), cs.value


def read_refdump(path):
    """Reader for the records written by oracle/ref_shims/polb200_dump.h."""
    out = {}
    data = Path(path).read_bytes()
    pos = 0
    while pos < len(data):
        name = data[pos:pos + 16].split(b"\0", 1)[0].decode()
        dtype = chr(data[pos + 16])
        n = int(np.frombuffer(data, dtype=np.int64, count=1, offset=pos + 17)[0])
        pos += 25
        if dtype == "i":
            arr = np.frombuffer(data, dtype=np.int32, count=n, offset=pos).copy()
            pos += 4 * n
        else:
            arr = np.frombuffer(data, dtype=np.float64, count=n, offset=pos).copy()
            pos += 8 * n
        out[name] = arr
    return out
