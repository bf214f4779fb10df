"""ctypes binding of the C ABI (include/polb200.h) -- plumbing for tests, bench.py and smoke().

The product is libpolb200.so (hand-written CUDA behind a C ABI); this module only marshals numpy /
torch buffers into polb200_atoms and mirrors the reference's script-level interface
(pair_style / pair_coeff / pair_modify lines) so that tests read like LAMMPS input.
It never falls back to a CPU implementation: if the library or a CUDA device is missing, the calls
raise.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libpolb200.so"

OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED, ERR_OVERFLOW, ERR_NAN = range(7)
STATUS_DIVERGED, STATUS_REBUILT, STATUS_EXACT = 1, 2, 4
DEVICE_NONE = -1

ABI_SYMBOLS = [
    "polb200_abi_version", "polb200_create", "polb200_destroy", "polb200_last_error", "polb200_settings",
    "polb200_set_ntypes", "polb200_coeff", "polb200_pair_modify", "polb200_init", "polb200_init_one",
    "polb200_extract", "polb200_single", "polb200_restart_size", "polb200_write_restart",
    "polb200_read_restart", "polb200_restart_settings_size", "polb200_read_restart_settings", "polb200_set_box", "polb200_compute", "polb200_comm_id_size",
    "polb200_comm_create_id", "polb200_comm_init", "polb200_subdomain", "polb200_debug_fetch",
    "polb200_launch_count", "polb200_set_option", "polb200_dev_alloc", "polb200_dev_free", "polb200_dev_copy",
    "polb200_dev_zero", "polb200_host_register", "polb200_host_unregister", "polb200_decomp_plan", "polb200_tail", "polb200_set_exclusions",
    "polb200_ewald_create", "polb200_ewald_destroy", "polb200_ewald_last_error", "polb200_ewald_init",
    "polb200_ewald_compute", "polb200_ewald_last_ms", "polb200_ewald_comm_init", "polb200_pppm_comm_init",
    "polb200_pppm_create", "polb200_pppm_destroy", "polb200_pppm_last_error", "polb200_pppm_init", "polb200_pppm_compute",
    "polb200_pppm_last_ms",
    "polb200_rigid_create", "polb200_rigid_destroy", "polb200_rigid_last_error", "polb200_rigid_init",
    "polb200_rigid_dof", "polb200_rigid_setup", "polb200_rigid_initial_integrate", "polb200_rigid_final_integrate",
    "polb200_rigid_pre_neighbor", "polb200_rigid_virial", "polb200_rigid_scalar", "polb200_rigid_reset_dt",
    "polb200_rigid_get_chain", "polb200_rigid_set_chain", "polb200_rigid_comm_init", "polb200_comm_init_replicated",
    "polb200_rigid_fetch", "polb200_rigid_launch_count", "polb200_rigid_last_ms",
]


class Env(C.Structure):
    _fields_ = [("g_ewald", C.c_double), ("qqrd2e", C.c_double), ("special_lj", C.c_double * 4),
                ("special_coul", C.c_double * 4), ("newton_pair", C.c_int), ("skin", C.c_double),
                ("neigh_every", C.c_int), ("neigh_delay", C.c_int), ("neigh_check", C.c_int),
                ("kspace_present", C.c_int), ("q_flag", C.c_int), ("polarizability_flag", C.c_int),
                ("molecular", C.c_int)]


class Atoms(C.Structure):
    _fields_ = [("nlocal", C.c_int), ("x", C.c_void_p), ("q", C.c_void_p), ("type", C.c_void_p),
                ("molecule", C.c_void_p), ("tag", C.c_void_p), ("alpha", C.c_void_p), ("mu", C.c_void_p),
                ("ef_static", C.c_void_p), ("f", C.c_void_p), ("nspecial", C.c_void_p),
                ("special", C.c_void_p), ("maxspecial", C.c_int), ("on_device", C.c_int),
                ("eatom", C.c_void_p), ("vatom", C.c_void_p), ("mask", C.c_void_p)]


class Exclusion(C.Structure):
    _fields_ = [("kind", C.c_int), ("a", C.c_int), ("b", C.c_int)]


EXCL_KINDS = {"type": 0, "group": 1, "molecule/intra": 2, "molecule/inter": 3, "include": 4}


class Result(C.Structure):
    _fields_ = [("eng_vdwl", C.c_double), ("eng_coul", C.c_double), ("eng_pol", C.c_double),
                ("virial", C.c_double * 6), ("u_self", C.c_double), ("u_ef", C.c_double),
                ("u_dd", C.c_double), ("rmin", C.c_double), ("iterations", C.c_int), ("status", C.c_int),
                ("npairs_full", C.c_long), ("nghost", C.c_int), ("ms_neigh", C.c_float),
                ("ms_pair", C.c_float), ("ms_scf", C.c_float), ("ms_force", C.c_float),
                ("ms_total", C.c_float)]


class EwaldSetup(C.Structure):
    _fields_ = [("accuracy_relative", C.c_double), ("g_ewald", C.c_double), ("qqrd2e", C.c_double),
                ("two_charge_force", C.c_double), ("qsum", C.c_double), ("qsqsum", C.c_double), ("natoms", C.c_long),
                ("cutoff", C.c_double), ("boxlo", C.c_double * 3), ("boxhi", C.c_double * 3), ("periodic", C.c_int * 3)]


class EwaldInfo(C.Structure):
    _fields_ = [("g_ewald", C.c_double), ("gsqmx", C.c_double), ("kxmax", C.c_int), ("kymax", C.c_int),
                ("kzmax", C.c_int), ("kmax", C.c_int), ("kcount", C.c_int)]


class PppmSetup(C.Structure):
    _fields_ = [("accuracy_relative", C.c_double), ("g_ewald", C.c_double), ("order", C.c_int), ("mesh", C.c_int * 3),
                ("qqrd2e", C.c_double), ("two_charge_force", C.c_double), ("qsum", C.c_double), ("qsqsum", C.c_double),
                ("natoms", C.c_long), ("cutoff", C.c_double), ("boxlo", C.c_double * 3), ("boxhi", C.c_double * 3),
                ("periodic", C.c_int * 3)]


class PppmInfo(C.Structure):
    _fields_ = [("g_ewald", C.c_double), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("order", C.c_int)]


class RigidParams(C.Structure):
    _fields_ = [("thermostat", C.c_int), ("t_start", C.c_double), ("t_stop", C.c_double), ("t_period", C.c_double),
                ("t_chain", C.c_int), ("t_iter", C.c_int), ("t_order", C.c_int), ("dt", C.c_double),
                ("ftm2v", C.c_double), ("mvv2e", C.c_double), ("boltz", C.c_double), ("boxlo", C.c_double * 3),
                ("boxhi", C.c_double * 3), ("periodic", C.c_int * 3)]


class RigidInfo(C.Structure):
    _fields_ = [("nbody", C.c_int), ("nlinear", C.c_int), ("nf_t", C.c_int), ("nf_r", C.c_int), ("maxmembers", C.c_int)]


class RigidAtoms(C.Structure):
    _fields_ = [("nlocal", C.c_int), ("tag", C.c_void_p), ("x", C.c_void_p), ("v", C.c_void_p), ("f", C.c_void_p),
                ("on_device", C.c_int)]


class Polb200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"polb200 error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib = None


def build(verbose=False):
    """Compile libpolb200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(HERE / "csrc")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libpolb200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (no CPU fallback exists)")
        L = C.CDLL(str(LIB_PATH))
        L.polb200_last_error.restype = C.c_char_p
        L.polb200_extract.restype = C.c_void_p
        L.polb200_debug_fetch.restype = C.c_long
        L.polb200_launch_count.restype = C.c_long
        L.polb200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.polb200_destroy.argtypes = [C.c_void_p]
        L.polb200_last_error.argtypes = [C.c_void_p]
        for name in ("polb200_settings", "polb200_coeff", "polb200_pair_modify"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p)]
        L.polb200_set_ntypes.argtypes = [C.c_void_p, C.c_int]
        L.polb200_init.argtypes = [C.c_void_p, C.POINTER(Env)]
        L.polb200_init_one.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.polb200_extract.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.polb200_single.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                     C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.polb200_tail.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double),
                                   C.POINTER(C.c_double)]
        L.polb200_restart_size.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
        L.polb200_write_restart.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
        L.polb200_read_restart.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
        L.polb200_restart_settings_size.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
        L.polb200_read_restart_settings.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long)]
        L.polb200_dev_alloc.argtypes = [C.c_int, C.c_size_t]
        L.polb200_dev_alloc.restype = C.c_void_p
        L.polb200_dev_free.argtypes = [C.c_int, C.c_void_p]
        L.polb200_dev_copy.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.polb200_dev_zero.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
        L.polb200_ewald_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.polb200_pppm_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.polb200_rigid_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.polb200_comm_init_replicated.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.polb200_set_box.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.polb200_compute.argtypes = [C.c_void_p, C.POINTER(Atoms), C.c_int, C.c_int, C.c_int, C.POINTER(Result)]
        L.polb200_debug_fetch.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long]
        L.polb200_launch_count.argtypes = [C.c_void_p, C.c_int]
        L.polb200_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.polb200_set_exclusions.argtypes = [C.c_void_p, C.c_int, C.POINTER(Exclusion)]
        L.polb200_comm_create_id.argtypes = [C.c_void_p]
        L.polb200_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.polb200_subdomain.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.polb200_ewald_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.polb200_ewald_destroy.argtypes = [C.c_void_p]
        L.polb200_ewald_last_error.argtypes = [C.c_void_p]
        L.polb200_ewald_last_error.restype = C.c_char_p
        L.polb200_ewald_init.argtypes = [C.c_void_p, C.POINTER(EwaldSetup), C.POINTER(EwaldInfo)]
        L.polb200_ewald_compute.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                            C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.polb200_ewald_last_ms.argtypes = [C.c_void_p]
        L.polb200_ewald_last_ms.restype = C.c_double
        L.polb200_pppm_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.polb200_pppm_destroy.argtypes = [C.c_void_p]
        L.polb200_pppm_last_error.argtypes = [C.c_void_p]
        L.polb200_pppm_last_error.restype = C.c_char_p
        L.polb200_pppm_init.argtypes = [C.c_void_p, C.POINTER(PppmSetup), C.POINTER(PppmInfo)]
        L.polb200_pppm_compute.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                           C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.polb200_pppm_last_ms.argtypes = [C.c_void_p]
        L.polb200_pppm_last_ms.restype = C.c_double
        L.polb200_rigid_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.polb200_rigid_destroy.argtypes = [C.c_void_p]
        L.polb200_rigid_last_error.argtypes = [C.c_void_p]
        L.polb200_rigid_last_error.restype = C.c_char_p
        L.polb200_rigid_init.argtypes = [C.c_void_p, C.POINTER(RigidParams), C.c_int] + [C.c_void_p] * 7 + [C.POINTER(RigidInfo)]
        L.polb200_rigid_dof.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.polb200_rigid_setup.argtypes = [C.c_void_p, C.POINTER(RigidAtoms), C.c_int]
        L.polb200_rigid_initial_integrate.argtypes = [C.c_void_p, C.POINTER(RigidAtoms), C.c_int, C.c_double]
        L.polb200_rigid_final_integrate.argtypes = [C.c_void_p, C.POINTER(RigidAtoms)]
        L.polb200_rigid_pre_neighbor.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.polb200_rigid_virial.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.polb200_rigid_scalar.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.polb200_rigid_reset_dt.argtypes = [C.c_void_p, C.c_double]
        L.polb200_rigid_get_chain.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.polb200_rigid_set_chain.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.polb200_rigid_fetch.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long]
        L.polb200_rigid_fetch.restype = C.c_long
        L.polb200_rigid_launch_count.argtypes = [C.c_void_p, C.c_int]
        L.polb200_rigid_launch_count.restype = C.c_long
        L.polb200_rigid_last_ms.argtypes = [C.c_void_p]
        L.polb200_rigid_last_ms.restype = C.c_double
        L.polb200_decomp_plan.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int),
                                          C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                          C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _argv(words):
    arr = (C.c_char_p * len(words))(*[w.encode() for w in words])
    return len(words), arr


REAL_QQRD2E = 332.06371  # units real, src/update.cpp:157


def decomp_plan(nranks, rank, procgrid, periodic, boxlo, boxhi):
    """Host-side plan of the brick decomposition (no device needed): dict with dest[27], src[27],
    wrap[27,3], sublo[3], subhi[3]; direction d = (dz+1)*9 + (dy+1)*3 + (dx+1)."""
    dest = (C.c_int * 27)()
    src = (C.c_int * 27)()
    wrap = (C.c_int * 81)()
    sublo = (C.c_double * 3)()
    subhi = (C.c_double * 3)()
    rc = lib().polb200_decomp_plan(nranks, rank, (C.c_int * 3)(*[int(v) for v in procgrid]),
                                   (C.c_int * 3)(*[int(v) for v in periodic]),
                                   (C.c_double * 3)(*[float(v) for v in boxlo]),
                                   (C.c_double * 3)(*[float(v) for v in boxhi]), dest, src, wrap, sublo, subhi)
    if rc != OK:
        raise Polb200Error(rc, "Bad grid of processors")
    return dict(dest=np.array(dest[:]), src=np.array(src[:]), wrap=np.array(wrap[:]).reshape(27, 3),
                sublo=np.array(sublo[:]), subhi=np.array(subhi[:]))


def comm_create_id():
    """Opaque NCCL unique id (bytes) that rank 0 creates and every rank passes to PairStyle.comm_init."""
    n = lib().polb200_comm_id_size()
    buf = (C.c_char * n)()
    rc = lib().polb200_comm_create_id(buf)
    if rc != OK:
        raise Polb200Error(rc, "polb200_comm_create_id failed (NCCL not loadable?)")
    return bytes(buf)


class PairStyle:
    """One instance of pair style lj/cut/coul/long/polarization on one GPU.

    Method names follow the reference's Pair interface (settings / coeff / init_style+init_one /
    compute / single / extract / write_restart ...), arguments are the words of the input script.
    """

    STYLE = "lj/cut/coul/long/polarization"

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().polb200_create(C.byref(self._h), device)
        if rc != OK:
            raise Polb200Error(rc, "polb200_create failed (no CUDA device? there is no CPU fallback)")
        self.device = device
        self.ntypes = 0

    def close(self):
        if self._h:
            lib().polb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise Polb200Error(rc, lib().polb200_last_error(self._h).decode())

    # ---- script-level interface ----
    def command(self, line):
        """Feed one input-script line: pair_style / pair_coeff / pair_modify."""
        w = line.split("#", 1)[0].split()
        if not w:
            return
        if w[0] == "pair_style":
            if w[1] != self.STYLE:
                raise Polb200Error(ERR_ARG, f"Unknown pair style {w[1]}")
            self.settings(w[2:])
        elif w[0] == "pair_coeff":
            self.coeff(w[1:])
        elif w[0] == "pair_modify":
            self.pair_modify(w[1:])
        else:
            raise Polb200Error(ERR_ARG, f"Unknown command: {w[0]}")

    def settings(self, words):
        self._check(lib().polb200_settings(self._h, *_argv(list(words))))

    def set_ntypes(self, n):
        self._check(lib().polb200_set_ntypes(self._h, n))
        self.ntypes = n

    def coeff(self, words):
        self._check(lib().polb200_coeff(self._h, *_argv([str(w) for w in words])))

    def pair_modify(self, words):
        self._check(lib().polb200_pair_modify(self._h, *_argv(list(words))))

    def init(self, g_ewald, qqrd2e=REAL_QQRD2E, special_lj=(1.0, 0.0, 0.0, 0.0),
             special_coul=(1.0, 0.0, 0.0, 0.0), newton_pair=1, skin=2.0, neigh_every=1, neigh_delay=10,
             neigh_check=1, kspace_present=1, q_flag=1, polarizability_flag=1, molecular=1):
        e = Env()
        e.g_ewald, e.qqrd2e = g_ewald, qqrd2e
        e.special_lj = (C.c_double * 4)(*special_lj)
        e.special_coul = (C.c_double * 4)(*special_coul)
        e.newton_pair, e.skin = newton_pair, skin
        e.neigh_every, e.neigh_delay, e.neigh_check = neigh_every, neigh_delay, neigh_check
        e.kspace_present, e.q_flag, e.polarizability_flag, e.molecular = (kspace_present, q_flag,
                                                                          polarizability_flag, molecular)
        self._check(lib().polb200_init(self._h, C.byref(e)))

    def init_one(self, i, j):
        cut = C.c_double()
        self._check(lib().polb200_init_one(self._h, i, j, C.byref(cut)))
        return cut.value

    def tail(self, i, j, count_i, count_j):
        e, p = C.c_double(), C.c_double()
        self._check(lib().polb200_tail(self._h, i, j, count_i, count_j, C.byref(e), C.byref(p)))
        return e.value, p.value

    def extract(self, name):
        dim = C.c_int()
        p = lib().polb200_extract(self._h, name.encode(), C.byref(dim))
        if not p:
            return None, dim.value
        if dim.value == 0:
            return C.cast(p, C.POINTER(C.c_double))[0], 0
        n1 = self.ntypes + 1
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n1, n1)).copy(), 2

    def single(self, itype, jtype, qi, qj, rsq, factor_coul=1.0, factor_lj=1.0):
        f, e = C.c_double(), C.c_double()
        self._check(lib().polb200_single(self._h, itype, jtype, qi, qj, rsq, factor_coul, factor_lj,
                                         C.byref(f), C.byref(e)))
        return e.value, f.value

    def write_restart(self):
        n = C.c_long()
        self._check(lib().polb200_restart_size(self._h, C.byref(n)))
        buf = (C.c_char * n.value)()
        self._check(lib().polb200_write_restart(self._h, buf, n.value))
        return bytes(buf)

    def read_restart(self, image):
        buf = C.create_string_buffer(image, len(image))
        self._check(lib().polb200_read_restart(self._h, buf, len(image)))

    def restart_settings_size(self):
        n = C.c_long()
        self._check(lib().polb200_restart_settings_size(self._h, C.byref(n)))
        return n.value

    def read_restart_settings(self, image):
        """the settings block alone (what PairHybrid hands its sub-styles); returns the bytes consumed"""
        buf = C.create_string_buffer(image, len(image))
        used = C.c_long()
        self._check(lib().polb200_read_restart_settings(self._h, buf, len(image), C.byref(used)))
        return used.value

    def set_box(self, boxlo, boxhi, periodic=(1, 1, 1)):
        lo = (C.c_double * 3)(*[float(v) for v in boxlo])
        hi = (C.c_double * 3)(*[float(v) for v in boxhi])
        per = (C.c_int * 3)(*[int(v) for v in periodic])
        self._check(lib().polb200_set_box(self._h, lo, hi, per))

    # ---- multi-GPU: one process per GPU, bricks of the box ----
    def comm_init(self, rank, nranks, id_bytes, procgrid):
        buf = C.create_string_buffer(id_bytes, len(id_bytes))
        self._check(lib().polb200_comm_init(self._h, rank, nranks, buf, (C.c_int * 3)(*[int(v) for v in procgrid])))

    def comm_init_replicated(self, rank, nranks, id_bytes):
        """all-pairs (exact) mode shared by rows: every process passes the whole system and gets the whole result"""
        buf = C.create_string_buffer(id_bytes, len(id_bytes))
        self._check(lib().polb200_comm_init_replicated(self._h, rank, nranks, buf))

    def subdomain(self):
        lo = (C.c_double * 3)()
        hi = (C.c_double * 3)()
        self._check(lib().polb200_subdomain(self._h, lo, hi))
        return np.array(lo[:]), np.array(hi[:])

    def set_exclusions(self, rules):
        """`neigh_modify exclude` rules as tuples: ("type", i, j), ("group", bit1, bit2), ("molecule/intra", bit),
        ("molecule/inter", bit), ("include", bit) for `neigh_modify include`; [] clears them"""
        arr = (Exclusion * max(len(rules), 1))()
        for k, r in enumerate(rules):
            arr[k].kind, arr[k].a, arr[k].b = EXCL_KINDS[r[0]], int(r[1]), int(r[2]) if len(r) > 2 else 0
        self._check(lib().polb200_set_exclusions(self._h, len(rules), arr))

    def set_option(self, name, value):
        self._check(lib().polb200_set_option(self._h, name.encode(), float(value)))

    # ---- hot path ----
    def compute(self, x, q, type_, alpha, mu, f, molecule=None, tag=None, ef_static=None, nspecial=None,
                special=None, eflag=1, vflag=2, ago=0, eatom=None, vatom=None, mask=None):
        """One compute() call on HOST numpy buffers (mu and f updated in place).  Returns Result."""
        n = x.shape[0]
        a = Atoms()
        a.nlocal = n
        keep = []

        def ptr(arr, dtype):
            if arr is None:
                return None
            assert arr.dtype == dtype and arr.flags["C_CONTIGUOUS"], "pass contiguous arrays of the ABI dtype"
            keep.append(arr)
            return arr.ctypes.data

        a.x, a.q, a.alpha = ptr(x, np.float64), ptr(q, np.float64), ptr(alpha, np.float64)
        a.type, a.molecule, a.tag = ptr(type_, np.int32), ptr(molecule, np.int32), ptr(tag, np.int32)
        a.mu, a.f, a.ef_static = ptr(mu, np.float64), ptr(f, np.float64), ptr(ef_static, np.float64)
        a.nspecial, a.special = ptr(nspecial, np.int32), ptr(special, np.int32)
        a.maxspecial = special.shape[1] if special is not None else 0
        a.eatom, a.vatom = ptr(eatom, np.float64), ptr(vatom, np.float64)
        a.mask = ptr(mask, np.int32)
        a.on_device = 0
        res = Result()
        self._check(lib().polb200_compute(self._h, C.byref(a), eflag, vflag, ago, C.byref(res)))
        return res

    def compute_device(self, n, ptrs, eflag=1, vflag=2, ago=0, maxspecial=0):
        """compute() on DEVICE pointers (dict name -> int address), e.g. torch tensors' data_ptr()."""
        a = Atoms()
        a.nlocal = n
        for k in ("x", "q", "type", "molecule", "tag", "alpha", "mu", "ef_static", "f", "nspecial", "special", "eatom",
                  "vatom", "mask"):
            setattr(a, k, ptrs.get(k))
        a.maxspecial = maxspecial
        a.on_device = 1
        res = Result()
        self._check(lib().polb200_compute(self._h, C.byref(a), eflag, vflag, ago, C.byref(res)))
        return res

    def debug_fetch(self, name, dtype, count_hint):
        buf = np.zeros(count_hint, dtype=dtype)
        n = lib().polb200_debug_fetch(self._h, name.encode(), buf.ctypes.data, buf.nbytes)
        if n < 0:
            raise Polb200Error(ERR_ARG, lib().polb200_last_error(self._h).decode())
        return buf

    def launch_count(self, reset=False):
        return lib().polb200_launch_count(self._h, int(reset))


class Ewald:
    """`kspace_style ewald <accuracy>` on one GPU: the device counterpart of the reference's class Ewald
    (src/KSPACE/ewald.cpp).  init() = Ewald::init + setup, compute() = Ewald::compute."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().polb200_ewald_create(C.byref(self._h), device)
        if rc != OK:
            raise Polb200Error(rc, "polb200_ewald_create failed (no CUDA device? there is no CPU fallback)")

    def close(self):
        if self._h:
            lib().polb200_ewald_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise Polb200Error(rc, lib().polb200_ewald_last_error(self._h).decode())

    def init(self, accuracy, q, cutoff, boxlo, boxhi, g_ewald=0.0, qqrd2e=REAL_QQRD2E, two_charge_force=REAL_QQRD2E,
             periodic=(1, 1, 1), natoms=None):
        q = np.asarray(q, dtype=np.float64)
        s = EwaldSetup()
        s.accuracy_relative, s.g_ewald, s.qqrd2e, s.two_charge_force = accuracy, g_ewald, qqrd2e, two_charge_force
        s.qsum, s.qsqsum = float(np.cumsum(q)[-1]) if len(q) else 0.0, float(np.cumsum(q * q)[-1]) if len(q) else 0.0
        s.natoms = len(q) if natoms is None else natoms
        s.cutoff = cutoff
        s.boxlo = (C.c_double * 3)(*[float(v) for v in boxlo])
        s.boxhi = (C.c_double * 3)(*[float(v) for v in boxhi])
        s.periodic = (C.c_int * 3)(*[int(v) for v in periodic])
        info = EwaldInfo()
        self._check(lib().polb200_ewald_init(self._h, C.byref(s), C.byref(info)))
        return info

    def compute(self, x, q, f, eflag=1, vflag=1):
        """f (n,3) += KSpace forces; returns (energy, virial[6]).  Host numpy buffers."""
        n = x.shape[0]
        for a in (x, q, f):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        e = C.c_double()
        v = (C.c_double * 6)()
        self._check(lib().polb200_ewald_compute(self._h, n, x.ctypes.data, q.ctypes.data, f.ctypes.data, eflag, vflag, 0,
                                                C.byref(e), v))
        return e.value, np.array(v[:])

    def compute_device(self, n, x_ptr, q_ptr, f_ptr, eflag=1, vflag=1):
        e = C.c_double()
        v = (C.c_double * 6)()
        self._check(lib().polb200_ewald_compute(self._h, n, x_ptr, q_ptr, f_ptr, eflag, vflag, 1, C.byref(e), v))
        return e.value, np.array(v[:])

    def last_ms(self):
        return lib().polb200_ewald_last_ms(self._h)

    def comm_init(self, rank, nranks, id_bytes):
        """multi-GPU: every rank passes its own atoms to compute(); init() takes the GLOBAL charges"""
        buf = C.create_string_buffer(id_bytes, len(id_bytes))
        self._check(lib().polb200_ewald_comm_init(self._h, rank, nranks, buf))


class PPPM:
    """`kspace_style pppm <accuracy>` on one GPU: the device counterpart of the reference's class PPPM
    (src/KSPACE/pppm.cpp).  init() = PPPM::init + setup, compute() = PPPM::compute."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().polb200_pppm_create(C.byref(self._h), device)
        if rc != OK:
            raise Polb200Error(rc, "polb200_pppm_create failed (no CUDA device? there is no CPU fallback)")

    def close(self):
        if self._h:
            lib().polb200_pppm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise Polb200Error(rc, lib().polb200_pppm_last_error(self._h).decode())

    def init(self, accuracy, q, cutoff, boxlo, boxhi, g_ewald=0.0, order=0, mesh=(0, 0, 0), qqrd2e=REAL_QQRD2E,
             two_charge_force=REAL_QQRD2E, periodic=(1, 1, 1), natoms=None):
        q = np.asarray(q, dtype=np.float64)
        s = PppmSetup()
        s.accuracy_relative, s.g_ewald, s.order = accuracy, g_ewald, order
        s.mesh = (C.c_int * 3)(*[int(v) for v in mesh])
        s.qqrd2e, s.two_charge_force = qqrd2e, two_charge_force
        s.qsum, s.qsqsum = float(np.cumsum(q)[-1]) if len(q) else 0.0, float(np.cumsum(q * q)[-1]) if len(q) else 0.0
        s.natoms = len(q) if natoms is None else natoms
        s.cutoff = cutoff
        s.boxlo = (C.c_double * 3)(*[float(v) for v in boxlo])
        s.boxhi = (C.c_double * 3)(*[float(v) for v in boxhi])
        s.periodic = (C.c_int * 3)(*[int(v) for v in periodic])
        info = PppmInfo()
        self._check(lib().polb200_pppm_init(self._h, C.byref(s), C.byref(info)))
        return info

    def compute(self, x, q, f, eflag=1, vflag=1):
        """f (n,3) += KSpace forces; returns (energy, virial[6]).  Host numpy buffers."""
        for a in (x, q, f):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        e = C.c_double()
        v = (C.c_double * 6)()
        self._check(lib().polb200_pppm_compute(self._h, x.shape[0], x.ctypes.data, q.ctypes.data, f.ctypes.data, eflag, vflag, 0,
                                               C.byref(e), v))
        return e.value, np.array(v[:])

    def compute_device(self, n, x_ptr, q_ptr, f_ptr, eflag=1, vflag=1):
        e = C.c_double()
        v = (C.c_double * 6)()
        self._check(lib().polb200_pppm_compute(self._h, n, x_ptr, q_ptr, f_ptr, eflag, vflag, 1, C.byref(e), v))
        return e.value, np.array(v[:])

    def last_ms(self):
        return lib().polb200_pppm_last_ms(self._h)

    def comm_init(self, rank, nranks, id_bytes):
        """multi-GPU: every rank passes its own atoms to compute(); init() takes the GLOBAL charges"""
        buf = C.create_string_buffer(id_bytes, len(id_bytes))
        self._check(lib().polb200_pppm_comm_init(self._h, rank, nranks, buf))


# real units (src/update.cpp:150-170)
REAL_FTM2V = 1.0 / 48.88821291 / 48.88821291
REAL_MVV2E = 48.88821291 * 48.88821291
REAL_BOLTZ = 0.0019872067
REAL_NKTV2P = 68568.415


def pack_image(image3):
    """(n,3) integer image flags -> LAMMPS' packed 32-bit imageint (src/lmptype.h:96-103)"""
    im = np.asarray(image3, dtype=np.int64) + 512
    return ((im[:, 2] << 20) | (im[:, 1] << 10) | im[:, 0]).astype(np.int32)


class Rigid:
    """`fix ID group rigid/nve molecule` / `rigid/nvt molecule temp ... tparam ...` on one GPU: the device counterpart
    of the reference's FixRigidNH (src/RIGID/fix_rigid_nh.cpp).  Methods mirror the Fix interface: init, dof, setup,
    initial_integrate, final_integrate, pre_neighbor, compute_scalar."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().polb200_rigid_create(C.byref(self._h), device)
        if rc != OK:
            raise Polb200Error(rc, "polb200_rigid_create failed (no CUDA device? there is no CPU fallback)")
        self.info = None

    def close(self):
        if self._h:
            lib().polb200_rigid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise Polb200Error(rc, lib().polb200_rigid_last_error(self._h).decode())

    def init(self, tag, molecule, mass, image, x, v, boxlo, boxhi, dt, ingroup=None, temp=None, tparam=(10, 1, 3),
             ftm2v=REAL_FTM2V, mvv2e=REAL_MVV2E, boltz=REAL_BOLTZ, periodic=(1, 1, 1)):
        """image: packed imageint [n] or (n,3) flags.  temp = (Tstart, Tstop, Tdamp) selects rigid/nvt."""
        p = RigidParams()
        p.thermostat = 0 if temp is None else 1
        if temp is not None:
            p.t_start, p.t_stop, p.t_period = [float(t) for t in temp]
        p.t_chain, p.t_iter, p.t_order = [int(t) for t in tparam]
        p.dt, p.ftm2v, p.mvv2e, p.boltz = float(dt), ftm2v, mvv2e, boltz
        p.boxlo = (C.c_double * 3)(*[float(t) for t in boxlo])
        p.boxhi = (C.c_double * 3)(*[float(t) for t in boxhi])
        p.periodic = (C.c_int * 3)(*[int(t) for t in periodic])
        image = np.asarray(image)
        if image.ndim == 2:
            image = pack_image(image)
        arrs = [np.ascontiguousarray(tag, dtype=np.int32), np.ascontiguousarray(molecule, dtype=np.int32),
                None if ingroup is None else np.ascontiguousarray(ingroup, dtype=np.int32),
                np.ascontiguousarray(mass, dtype=np.float64), np.ascontiguousarray(image, dtype=np.int32),
                np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(v, dtype=np.float64)]
        info = RigidInfo()
        self._check(lib().polb200_rigid_init(self._h, C.byref(p), len(arrs[0]),
                                             *[None if a is None else a.ctypes.data for a in arrs], C.byref(info)))
        self.info = info
        return info

    def comm_init(self, rank, nranks, id_bytes):
        """more than one process (before init): bodies replicated, every call takes this process's own atoms"""
        buf = C.create_string_buffer(id_bytes, len(id_bytes))
        self._check(lib().polb200_rigid_comm_init(self._h, rank, nranks, buf))

    def dof(self, tag, tgroup=None):
        tag = np.ascontiguousarray(tag, dtype=np.int32)
        tg = None if tgroup is None else np.ascontiguousarray(tgroup, dtype=np.int32)
        out = C.c_int()
        self._check(lib().polb200_rigid_dof(self._h, len(tag), tag.ctypes.data, None if tg is None else tg.ctypes.data,
                                            C.byref(out)))
        return out.value

    @staticmethod
    def _atoms(tag, x, v, f):
        for a in (x, v, f):
            assert a is None or (a.dtype == np.float64 and a.flags["C_CONTIGUOUS"])
        assert tag.dtype == np.int32 and tag.flags["C_CONTIGUOUS"]
        return RigidAtoms(len(tag), tag.ctypes.data, None if x is None else x.ctypes.data, v.ctypes.data,
                          None if f is None else f.ctypes.data, 0)

    def setup(self, tag, x, v, f, vflag=1):
        """host numpy buffers; v is updated in place"""
        a = self._atoms(tag, x, v, f)
        self._check(lib().polb200_rigid_setup(self._h, C.byref(a), vflag))

    def initial_integrate(self, tag, x, v, f, vflag=1, run_fraction=0.0):
        a = self._atoms(tag, x, v, f)
        self._check(lib().polb200_rigid_initial_integrate(self._h, C.byref(a), vflag, run_fraction))

    def final_integrate(self, tag, x, v, f):
        a = self._atoms(tag, x, v, f)
        self._check(lib().polb200_rigid_final_integrate(self._h, C.byref(a)))

    # device-resident variants: integer addresses of device buffers (e.g. torch.Tensor.data_ptr())
    def setup_device(self, n, tag_ptr, x_ptr, v_ptr, f_ptr, vflag=1):
        a = RigidAtoms(n, tag_ptr, x_ptr, v_ptr, f_ptr, 1)
        self._check(lib().polb200_rigid_setup(self._h, C.byref(a), vflag))

    def initial_integrate_device(self, n, tag_ptr, x_ptr, v_ptr, f_ptr, vflag=1, run_fraction=0.0):
        a = RigidAtoms(n, tag_ptr, x_ptr, v_ptr, f_ptr, 1)
        self._check(lib().polb200_rigid_initial_integrate(self._h, C.byref(a), vflag, run_fraction))

    def final_integrate_device(self, n, tag_ptr, x_ptr, v_ptr, f_ptr):
        a = RigidAtoms(n, tag_ptr, x_ptr, v_ptr, f_ptr, 1)
        self._check(lib().polb200_rigid_final_integrate(self._h, C.byref(a)))

    def pre_neighbor(self, tag, image):
        tag = np.ascontiguousarray(tag, dtype=np.int32)
        image = np.asarray(image)
        if image.ndim == 2:
            image = pack_image(image)
        image = np.ascontiguousarray(image, dtype=np.int32)
        self._check(lib().polb200_rigid_pre_neighbor(self._h, len(tag), tag.ctypes.data, image.ctypes.data, 0))

    def pre_neighbor_device(self, n, tag_ptr, image_ptr):
        """device int32 arrays: atom ids and packed image flags"""
        self._check(lib().polb200_rigid_pre_neighbor(self._h, n, tag_ptr, image_ptr, 1))

    def virial(self):
        v = (C.c_double * 6)()
        self._check(lib().polb200_rigid_virial(self._h, v))
        return np.array(v[:])

    def scalars(self):
        """(compute_scalar, translational KE, rotational KE) -- KE in mass*velocity^2 units (multiply by mvv2e)"""
        s, kt, kr = C.c_double(), C.c_double(), C.c_double()
        self._check(lib().polb200_rigid_scalar(self._h, C.byref(s), C.byref(kt), C.byref(kr)))
        return s.value, kt.value, kr.value

    def reset_dt(self, dt):
        self._check(lib().polb200_rigid_reset_dt(self._h, float(dt)))

    def get_chain(self):
        """thermostat state as FixRigidNH::write_restart stores it: (t_chain, 4) = eta_t, eta_r, eta_dot_t, eta_dot_r"""
        buf = np.zeros(4 * 64)
        nc = C.c_int()
        self._check(lib().polb200_rigid_get_chain(self._h, buf.ctypes.data, buf.size, C.byref(nc)))
        return buf[:4 * nc.value].reshape(nc.value, 4).copy()

    def set_chain(self, state):
        state = np.ascontiguousarray(state, dtype=np.float64)
        self._check(lib().polb200_rigid_set_chain(self._h, state.ctypes.data, state.shape[0]))

    def fetch(self, name):
        width = 4 if name in ("quat", "conjqm") else 1 if name == "masstotal" else 3
        out = np.zeros((self.info.nbody, width))
        n = lib().polb200_rigid_fetch(self._h, name.encode(), out.ctypes.data, out.size)
        if n < 0:
            raise Polb200Error(ERR_ARG, f"polb200_rigid_fetch({name}) = {n}")
        return out

    def launch_count(self, reset=False):
        return lib().polb200_rigid_launch_count(self._h, 1 if reset else 0)

    def last_ms(self):
        return lib().polb200_rigid_last_ms(self._h)
