#!/usr/bin/env python3
"""bench.py -- atom-steps/s (incl. the dipole SCF) of the lj/cut/coul/long/polarization hot path.

Workloads = BASELINE.json's configs, selected with --config:

  5 (default, every N): weak-scaled polarizable LJ+charge fluid, 1 000 188 atoms PER GPU (8.0 M on 8 GPUs), rho 0.1/A^3,
       pair_style lj/cut/coul/long/polarization 2.5 12 fixed_iteration yes max_iterations 30 damp_type exponential
       polar_gs_ranked no, dipole-dipole over the neighbor list (polar_cutoff 12).  One periodic system cut into N
       bricks (one process per GPU): ghost positions once per step, ghost dipoles once per sweep.  This is the series
       the metric "atom-steps/s at 1/2/4/8 B200" is quoted on; the same per-GPU workload at every N makes the driver's
       v_N / (N v_1) a true weak-scaling efficiency.
  2: the same fluid and keywords at 32 000 atoms per GPU (round-1 headline; at N = 1 it is ALSO measured by the default
       run and reported under "also.config2", with its roofline, for continuity)
  3: rigid water-like box, 255 552 atoms in total (strong scaling over N), precision 1e-11, polar_gs_ranked yes (the
       reference's default solver: group-coloured Gauss-Seidel sweep), polar_gamma 1.03; at N = 1 also measured by the
       default run ("also.config3")
  4: MOF-5 + CO2 supercell (the reference's example cell x 10^3 = 924 000 atoms, 10 atom types, bond topology),
       damp 2.1304, polar_gs_ranked yes precision 1e-11 -- strong scaling over N

A "step" = one compute() call = one pass of the whole hot path (neighbor refresh, LJ+Coulomb+static field, the dipole
iterations, polarization forces, reductions); the device structures are rebuilt every REBUILD_EVERY steps inside the timed
region, as LAMMPS' delay-10 schedule would.

  value  : whole-job atom-steps/s with inputs resident in HBM (device pointers through the C ABI)
  e2e    : same metric through the C ABI with HOST buffers (pinned), H2D of x/mu and D2H of f/mu/E inside
  roofline: dominant kernel k_sweep_group_tma (one dipole iteration; the Gauss-Seidel configs launch it once per colour,
            8 launches = one iteration): bytes it must move through HBM per iteration (36 B per pair-group row entry
            streamed + the atom records once) over the CUDA-event duration of an iteration, against
            MEASURED_PEAKS.json hbm_gbs; SURVEY §8d's every-gather-from-HBM model beside it
  cpu_baseline: the reference binary (oracle/_ref/lmp_serial, 1 core: it is serial by design) on a bounded
            4000-atom sample of the same fluid (all-pairs dipole semantics: it has no dipole cutoff), with the GPU run of
            EXACTLY that sample and those semantics beside it ("same_work"); plus the oracle port of the list algorithm
  N > 1  : before the timed run, a decomposed 32 000-atoms-per-GPU system is compared atom by atom with a single-GPU run
            of the same global system on rank 0 -> check.mgpu_max_rel_err (must be < 1e-10)
  --impl reference: times that CPU reference arm alone.
"""
import argparse
import importlib.util
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "atom-steps/s incl. dipole SCF"
UNIT = "atom-steps/s"
REBUILD_EVERY = 10
NCELL = 20  # config 2: 4*20^3 = 32000 atoms per GPU
NCELL5 = 63  # config 5: 4*63^3 = 1 000 188 atoms per GPU
CUT_LJ, CUT_COUL, ITER = 2.5, 12.0, 30
SAMPLE_NCELL = 10  # 4000-atom sample for the serial reference binary
STYLE_WORDS = (f"{CUT_LJ} {CUT_COUL} polar_gs_ranked no fixed_iteration yes max_iterations {ITER} "
               "damp_type exponential")
RANKED_WORDS = "precision 1e-11 max_iterations 200 polar_gamma 1.03 damp_type exponential"
MOF_CELL = ROOT / "tests" / "golden" / "co2_singlepoint_step0.npz"

PKG = ROOT / "lammps-induced-dipole-polarization-pair-style_b200"


def _load(name, filename):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, PKG / filename)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_pb():
    return _load("polb200", "polb200.py")


def workloads():
    """the synthetic systems (numpy only): the GPU arm never touches oracle/"""
    return _load("polb200_workloads", "workloads.py")


GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def make_config(cfg, world):
    """-> namespace(system, style lines, scaling, description) of BASELINE config `cfg` on `world` GPUs"""
    W = workloads()
    pg = GRIDS[world]
    c = SimpleNamespace(id=cfg, pg=pg, molecular=0, coeff=["pair_coeff * * 0.1 3.0"], cut=CUT_COUL, sweeps=None)
    if cfg in (2, 5):
        nc = NCELL if cfg == 2 else NCELL5
        c.sys = W.lj_charge_fluid(nc if world == 1 else tuple(nc * np.array(pg)), seed=12345)
        c.words = f"{STYLE_WORDS} polar_cutoff {CUT_COUL}"
        c.scaling = "weak"
        c.sweeps = ITER
        c.desc = (f"BASELINE config {cfg}: synthetic polarizable LJ+charge fluid, {4 * nc ** 3} atoms per GPU ({c.sys.n} in total), "
                  f"rho 0.1/A^3, cut {CUT_LJ}/{CUT_COUL}, fixed_iteration yes max_iterations {ITER}, damp_type exponential, "
                  f"polar_gs_ranked no, polar_cutoff {CUT_COUL} (neighbor-list dipole sweep)")
    elif cfg == 3:
        c.sys = W.water_box(44)
        c.words = f"{CUT_LJ} {CUT_COUL} {RANKED_WORDS} polar_cutoff {CUT_COUL}"
        c.coeff = ["pair_coeff 1 1 0.155 3.166", "pair_coeff 2 2 0.0 1.0"]
        c.scaling = "strong"
        c.desc = (f"BASELINE config 3: rigid 3-site water-like box, {c.sys.n} atoms in total, rho 0.1/A^3, cut {CUT_LJ}/{CUT_COUL}, "
                  f"precision 1e-11, polar_gs_ranked yes (group-coloured Gauss-Seidel sweep), polar_gamma 1.03, damp_type exponential, "
                  f"polar_cutoff {CUT_COUL}")
    elif cfg == 4:
        c.sys, c.cut, c.coeff = W.mof_supercell(MOF_CELL, 10)
        c.words = f"{CUT_LJ} {c.cut} {RANKED_WORDS} damp 2.1304 polar_cutoff {c.cut}"
        c.molecular = 1
        c.scaling = "strong"
        c.desc = (f"BASELINE config 4: MOF-5 + CO2 supercell (reference example cell x 10^3), {c.sys.n} atoms in total, 10 atom types, "
                  f"bond topology, cut {CUT_LJ}/{c.cut}, zodid no, damp 2.1304, damp_type exponential, polar_gs_ranked yes, "
                  f"precision 1e-11, polar_gamma 1.03, polar_cutoff {c.cut}")
    else:
        raise SystemExit(f"bench.py: unknown --config {cfg}")
    return c


def workload_config(c, n_gpus, parallelism):
    return {"workload": c.desc, "baseline_config": c.id, "atoms_per_gpu": c.sys.n // n_gpus, "atoms_total": c.sys.n,
            "sweeps_per_step": c.sweeps if c.sweeps else "to convergence", "rebuild_every": REBUILD_EVERY,
            "parallelism": parallelism,
            "l2_policy": "working set (neighbor list + radial cache, ~19 KB per atom) exceeds the 126 MB L2"}


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].startswith("Active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def nvlink_counters(index):
    """{tx, rx} bytes moved over all NVLinks of GPU `index` so far (nvidia-smi nvlink -gt d), None when unavailable"""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        return None
    tx = [int(v) for v in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out)]
    rx = [int(v) for v in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out)]
    if not tx and not rx:
        return None
    return {"tx": 1024 * sum(tx), "rx": 1024 * sum(rx)}


# ----------------------------------------------------------------------------------------------------
# CPU reference arm
# ----------------------------------------------------------------------------------------------------
def write_lammps_case(work, sysm, g_ewald_unused, steps):
    n = sysm.n
    L = float(sysm.boxhi[0])
    with open(work / "fluid.data", "w") as fh:
        fh.write("synthetic polarizable LJ+charge fluid\n\n%d atoms\n2 atom types\n\n" % n)
        fh.write("0.0 %.16g xlo xhi\n0.0 %.16g ylo yhi\n0.0 %.16g zlo zhi\n\nAtoms\n\n" % (L, L, L))
        for i in range(n):
            fh.write("%d 0 %d %.16g %.16g %.16g %.16g\n" % (i + 1, sysm.type[i], sysm.q[i], *sysm.x[i]))
    (work / "in.fluid").write_text(f"""units real
boundary p p p
atom_style full
read_data fluid.data
mass * 12.0
set type 1 static_polarizability 1.0
set type 2 static_polarizability 0.5
kspace_style ewald 1.0e-4
pair_style lj/cut/coul/long/polarization {STYLE_WORDS}
pair_coeff 1 1 0.1 3.0
pair_coeff 2 2 0.1 3.0
thermo_style custom step pe evdwl ecoul epol
thermo 1
fix 1 all nve
timestep 0.5
run {steps}
""")


def sample_description(n):
    return (f"{n}-atom sample of the BASELINE config 2/5 fluid (same density 0.1/A^3, cut {CUT_LJ}/{CUT_COUL}, fixed_iteration yes "
            f"max_iterations {ITER}, damp_type exponential, polar_gs_ranked no), dipole-dipole over ALL minimum-image pairs: the "
            "reference has no dipole cutoff and is O(N^2) with a dense 3Nx3N matrix (72 N^2 bytes), so it cannot run the GPU arm's "
            "sizes")


def run_reference_binary(steps, warmup, ncell=SAMPLE_NCELL):
    """Times oracle/_ref/lmp_serial (the repaired, otherwise unmodified reference) on the sample."""
    lmp = ROOT / "oracle" / "_ref" / "lmp_serial"
    if not lmp.exists():
        return None
    sysm = workloads().lj_charge_fluid(ncell)
    work = Path(tempfile.mkdtemp(prefix="polb200_refarm_"))
    try:
        write_lammps_case(work, sysm, None, steps + warmup)
        t0 = time.time()
        r = subprocess.run([str(lmp), "-in", "in.fluid", "-echo", "none"], cwd=work, capture_output=True, text=True)
        wall = time.time() - t0
        if r.returncode != 0:
            return {"error": r.stdout[-400:]}
        log = (work / "log.lammps").read_text()
        m = re.search(r"^Pair\s*\|\s*([0-9.eE+-]+)", log, flags=re.M)
        loop = re.search(r"Loop time of ([0-9.eE+-]+) on", log)
        pair_s = float(m.group(1)) if m else float(loop.group(1))
        nsteps = steps + warmup
        # run N = N+1 force evaluations (setup + N steps); the Pair timer covers the N steps
        per_step = pair_s / max(nsteps, 1)
        epol = None
        rows = re.findall(r"^\s*(\d+)\s+(-?[0-9.eE+-]+)\s+(-?[0-9.eE+-]+)\s+(-?[0-9.eE+-]+)\s+(-?[0-9.eE+-]+)\s*$", log, flags=re.M)
        if rows:
            epol = float(rows[0][4])  # step 0
        return {"atoms": sysm.n, "s_per_step": per_step, "value": sysm.n / per_step, "wall_s": wall,
                "steps": nsteps, "epol_step0": epol}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def run_port_sample(nthreads=0, target_s=10.0):
    """Oracle port (truncated list algorithm, OpenMP) on a row sample of the 32k-atom fluid."""
    import polhelpers as H
    from oracle import polref as P
    sysm = H.lj_charge_fluid(NCELL)
    st = H.fluid_style(sysm, CUT_LJ, CUT_COUL, polar_cut=CUT_COUL, fixed_iteration=1, max_iterations=ITER,
                       damp_type="exponential", polar_gs_ranked=0)
    rows = 64
    t, _ = P.bench_rows(sysm, st, 0, rows, ITER, nthreads)
    rows = int(min(sysm.n, max(64, rows * target_s / max(t, 1e-3))))
    t, _ = P.bench_rows(sysm, st, 0, rows, ITER, nthreads)
    return {"rows": rows, "seconds": t, "value": rows / t, "threads": nthreads or os.cpu_count()}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    ref = run_reference_binary(args.steps, args.warmup)
    if ref and "value" in ref:
        kind, cores, value = "reference", 1, ref["value"]
        natoms = ref["atoms"]
        sample = (f"oracle/_ref/lmp_serial (repaired reference, serial by design), {ref['steps']} MD steps, Pair timer, on a "
                  + sample_description(natoms))
        ms = ref["s_per_step"] * 1e3
    else:
        port = run_port_sample()
        kind, cores, value = "port", port["threads"], port["value"]
        natoms = port["rows"]
        sample = (f"oracle port (list algorithm with polar_cutoff {CUT_COUL}, OpenMP), rows sample: {port['rows']} of the 32000 atoms "
                  "of BASELINE config 2, polarization stages only")
        ms = port["seconds"] * 1e3
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": sample, "atoms_total": natoms, "atoms_per_gpu": None, "parallelism": "cpu, 1 core",
                       "same_config_as_gpu_arm": False,
                       "note": "the GPU arm's headline runs 1 000 188 atoms per GPU with the polar_cutoff extension; its line "
                               "carries cpu_baseline.same_work = the GPU timed on exactly this sample with these semantics"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def make_style(pb, c, device, words=None):
    # g_ewald as `kspace_style ewald 1e-4` would set it: the product's own Ewald::init (polb200_ewald_init)
    ew = pb.Ewald(device=device)
    g = ew.init(1e-4, c.sys.q, c.cut, c.sys.boxlo, c.sys.boxhi).g_ewald
    ew.close()
    s = pb.PairStyle(device=device)
    s.set_ntypes(int(c.sys.ntypes))
    s.command("pair_style lj/cut/coul/long/polarization " + (words or c.words))
    for line in c.coeff:
        s.command(line)
    s.init(g_ewald=g, molecular=c.molecular)
    s.set_box(c.sys.boxlo, c.sys.boxhi)
    return s


class Run:
    """one workload on this rank: owned atoms, device-resident and pinned host copies"""

    def __init__(self, pb, torch, dist, c, rank, world, local, comm=True):
        self.pb, self.torch, self.dist, self.c = pb, torch, dist, c
        self.rank, self.world, self.local = rank, world, local
        g = c.sys
        self.style = make_style(pb, c, local)
        dev = torch.device("cuda", local)
        if world > 1 and comm:
            box = [pb.comm_create_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            self.style.comm_init(rank, world, box[0], c.pg)
            lo, hi = self.style.subdomain()
            own = np.nonzero(np.all((g.x >= lo) & (g.x < hi), axis=1))[0]
        else:
            own = np.arange(g.n)
        self.own = own
        n = self.n = len(own)
        take = lambda a: np.ascontiguousarray(a[own])
        self.h = dict(x=take(g.x), q=take(g.q), type=take(g.type), alpha=take(g.alpha), tag=take(g.tag),
                      molecule=take(g.molecule))
        self.maxspecial = 0
        if getattr(g, "special", None) is not None:
            self.h["nspecial"], self.h["special"] = take(g.nspecial), take(g.special)
            self.maxspecial = g.special.shape[1]
        # device-resident inputs
        f64, i32 = torch.float64, torch.int32
        t = self.t = {k: torch.tensor(v, dtype=f64 if v.dtype == np.float64 else i32, device=dev).contiguous()
                      for k, v in self.h.items()}
        for k in ("mu", "f", "ef_static"):
            t[k] = torch.zeros((n, 3), dtype=f64, device=dev)
        self.ptrs = {k: v.data_ptr() for k, v in t.items()}
        torch.cuda.synchronize()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def step_dev(self, k):
        self.t["f"].zero_()
        self.torch.cuda.synchronize()
        return self.style.compute_device(self.n, self.ptrs, eflag=1, vflag=2, ago=k % REBUILD_EVERY, maxspecial=self.maxspecial)

    def pin_host(self):
        torch = self.torch
        self.p = {k: torch.tensor(self.h["x"] if k == "x" else np.zeros((self.n, 3)), dtype=torch.float64).pin_memory().numpy()
                  for k in ("x", "mu", "f", "ef")}

    def step_host(self, k):
        p, h = self.p, self.h
        p["f"][:] = 0.0
        return self.style.compute(p["x"], h["q"], h["type"], h["alpha"], p["mu"], p["f"], molecule=h["molecule"], tag=h["tag"],
                                  ef_static=p["ef"], nspecial=h.get("nspecial"), special=h.get("special"), eflag=1, vflag=2,
                                  ago=k % REBUILD_EVERY)

    def measure(self, steps, warmup, clocks=False, e2e=True):
        """-> dict of this rank's measurements (times are max-reduced over the ranks by the caller)"""
        pb, style = self.pb, self.style
        for k in range(warmup):
            res = self.step_dev(k)
        polar_pairs = int(style.debug_fetch("polar_pairs", np.uint64, 1)[0])
        group_stats = style.debug_fetch("group_stats", np.float64, 4)
        comm_stats = style.debug_fetch("comm_stats", np.float64, 5)
        style.set_option("time_sweeps", 1.0)
        style.launch_count(reset=True)
        sampler = ClockSampler(self.local) if clocks and self.rank == 0 else None
        if sampler:
            sampler.start()
        bar0 = style.debug_fetch("barrier_stats", np.float64, 2) if self.world > 1 else None
        nvl0 = nvlink_counters(self.local) if self.world > 1 and clocks else None
        self.barrier()
        t0 = time.perf_counter()
        dev_ms, stage, iters = 0.0, np.zeros(5), 0
        rebuild_ms = []
        for k in range(steps):
            res = self.step_dev(k)
            dev_ms += res.ms_total
            iters += res.iterations
            stage += [res.ms_neigh, res.ms_pair, res.ms_scf, res.ms_force, res.ms_total]
            if res.status & pb.STATUS_REBUILT:
                rebuild_ms.append(res.ms_neigh)
        self.barrier()
        wall = time.perf_counter() - t0
        nvl1 = nvlink_counters(self.local) if nvl0 else None
        bar1 = style.debug_fetch("barrier_stats", np.float64, 2) if self.world > 1 else None
        out = SimpleNamespace(wall=wall, launches=style.launch_count(), sweep=style.debug_fetch("sweep_timing", np.float64, 2),
                              clocks=sampler.stop() if sampler else None, eng_pol=res.eng_pol, res=res, stage=stage / steps,
                              dev_ms=dev_ms / steps, iterations=iters / steps, polar_pairs=polar_pairs, group_stats=group_stats,
                              comm_stats=comm_stats, wall_e2e=None, rebuild_ms=rebuild_ms,
                              nvlink={k: nvl1[k] - nvl0[k] for k in nvl0} if nvl0 and nvl1 else None,
                              barrier_ms=(bar1[0] - bar0[0]) * 1e-6 if bar0 is not None else None,
                              barriers=int(bar1[1] - bar0[1]) if bar0 is not None else None)
        style.set_option("time_sweeps", 0.0)
        if e2e:
            # ---- end to end through the C ABI with host (pinned) buffers ----
            self.pin_host()
            for k in range(max(1, warmup // 2)):
                self.step_host(k)
            self.barrier()
            t1 = time.perf_counter()
            for k in range(steps):
                r2 = self.step_host(k)
            self.barrier()
            out.wall_e2e = time.perf_counter() - t1
            assert abs(r2.eng_pol - out.eng_pol) <= 1e-9 * abs(out.eng_pol), (r2.eng_pol, out.eng_pol)
        return out

    def reduce(self, m):
        """times: max over ranks; atoms / energy / pairs: sums (per-rank partials by LAMMPS convention)"""
        torch, dist = self.torch, self.dist
        dev = torch.device("cuda", self.local)
        times = torch.tensor([m.wall, m.wall_e2e or 0.0], dtype=torch.float64, device=dev)
        sums = torch.tensor([float(self.n), m.eng_pol, float(m.polar_pairs)], dtype=torch.float64, device=dev)
        if self.world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
            dist.all_reduce(sums)
        m.wall, m.wall_e2e = float(times[0]), float(times[1]) or None
        m.total_atoms, m.eng_pol_total = int(sums[0]), float(sums[1])
        m.per_rank = None
        if self.world > 1:
            # per brick: owned atoms, ghosts, mean sweep-kernel time, time inside inter-GPU barriers, SCF stage
            mine = torch.tensor([float(self.n), float(m.res.nghost), float(m.sweep[0]) / max(float(m.sweep[1]), 1.0) * 1e3,
                                 m.barrier_ms or 0.0, float(m.barriers or 0), float(m.stage[2]), float(m.stage[4])],
                                dtype=torch.float64, device=dev)
            allr = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(allr, mine)
            keys = ("owned", "ghosts", "sweep_us", "barrier_ms_total", "barriers", "scf_ms_per_step", "device_ms_per_step")
            m.per_rank = [dict(zip(keys, [float(v) for v in t.cpu()])) for t in allr]
        return m

    def roofline(self, m, peaks, traffic_file=None):
        n, world = self.n, self.world
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        sweep_ms = float(m.sweep[0]) / max(float(m.sweep[1]), 1.0)
        # Bytes one iteration of the dominant kernel must move through HBM.  k_sweep_group_tma: one warp per
        # pair group (two cell-row neighbours); per group-row entry it streams, exactly once, 4 B of neighbour index + 32 B
        # of cached radial scalars {s1a,s2a,s1b,s2b}; the 32-B position and dipole records of the owned+ghost atoms are
        # gathered hundreds of times each but from L2/L1, so they count once; plus E_static in and the new dipole out per
        # owned atom.  (Per-atom fallback kernel k_sweep_cached: 20 B per pair instead.)
        nghost = int(m.res.nghost)
        grouped = bool(m.group_stats[3]) and m.group_stats[1] > 0
        entries = float(m.group_stats[1]) if grouped else float(m.polar_pairs)
        alg_bytes = (36.0 if grouped else 20.0) * entries + 64.0 * (n + nghost) + 64.0 * n
        achieved = alg_bytes / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None
        kernel = "k_sweep_group_tma" if grouped else "k_sweep_cached"
        gs = self.c.sweeps is None
        traffic = None
        if traffic_file is not None and traffic_file.exists() and world == 1:
            t = json.loads(traffic_file.read_text())
            for e in (t if isinstance(t, list) else [t]):
                if e.get("kernel") == kernel and e.get("atoms") == n and bool(e.get("gauss_seidel", False)) == gs:
                    traffic = e["dram_bytes_per_launch"] * (e.get("launches_per_iteration", 1))
        survey_bytes = 52.0 * m.polar_pairs + 104.0 * n  # every gather charged to HBM
        return {"bound": "hbm",
                "kernel": kernel + (" (one dipole iteration = one launch per colour of the group-coloured Gauss-Seidel sweep + commits)"
                                    if gs else " (one dipole iteration over the neighbor list = one launch)"),
                "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs if achieved else None,
                "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                "algorithmic_bytes_per_launch": alg_bytes,
                "bytes_model": ("36 B per group-row entry streamed (index + cached radial scalars of both members)"
                                if grouped else "20 B/pair streamed (index + cached radial scalars)") +
                               " + 64 B per owned+ghost atom (position and dipole records, read once) + 64 B per "
                               "owned atom (E_static in, dipole out)",
                "group_row_entries": entries if grouped else None,
                "pairs_in_cutoff": m.polar_pairs, "launch_ms": sweep_ms,
                "bytes_per_pair": alg_bytes / max(m.polar_pairs, 1),
                "gpairs_per_s": m.polar_pairs / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None,
                "survey_8d_model": {"bytes": survey_bytes,
                                    "achieved": survey_bytes / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None,
                                    "note": "52 B/pair + 104 B/atom with every neighbour gather charged to HBM; "
                                            "exceeds the HBM peak because the gathers are L2/L1 hits"}}

    def close(self):
        self.style.close()


def summary(run, m, steps, peaks):
    """compact record of a secondary workload measured in the same process"""
    return {"workload": run.c.desc, "value": m.total_atoms * steps / m.wall, "unit": UNIT, "ms_per_step": m.wall / steps * 1e3,
            "e2e_value": m.total_atoms * steps / m.wall_e2e if m.wall_e2e else None,
            "iterations_per_step": m.iterations, "us_per_dipole_iteration": float(m.sweep[0]) / max(float(m.sweep[1]), 1.0) * 1e3,
            "stage_ms": {"neigh_refresh": m.stage[0], "pair_field": m.stage[1], "scf": m.stage[2], "pol_force": m.stage[3]},
            "gpu_launches": int(m.launches), "roofline": run.roofline(m, peaks, ROOT / "profiles" / "ncu_traffic.json"),
            "check": {"eng_pol": m.eng_pol_total}}


def mgpu_parity(pb, torch, dist, rank, world, local):
    """decomposed run of a 32 000-atoms-per-GPU fluid (converged Jacobi + one fixed-sweep run) against the single-GPU run
    of the same global system on rank 0: max relative error of dipoles, static fields and forces over ALL atoms"""
    c = make_config(2, world)
    worst = 0.0
    detail = {}
    for label, words in (("fixed30", c.words), ("ranked", f"{CUT_LJ} {CUT_COUL} {RANKED_WORDS} polar_cutoff {CUT_COUL}")):
        c.words = words
        run = Run(pb, torch, dist, c, rank, world, local)
        run.pin_host()
        for k in range(2):  # a rebuild step and a step on stale lists
            r = run.step_host(k)
        parts = dict(own=run.own, mu=run.p["mu"].copy(), ef=run.p["ef"].copy(), f=run.p["f"].copy(), it=r.iterations)
        if label == "fixed30":
            # the same decomposed system timed (device-resident inputs): the 32 000-atoms-per-GPU weak-scaling point of round 1
            m2 = run.reduce(run.measure(20, 5, e2e=False))
            if rank == 0:
                detail["config2_weak"] = {"atoms_per_gpu": 4 * NCELL ** 3, "atoms_total": m2.total_atoms, "steps": 20,
                                          "ms_per_step": m2.wall / 20 * 1e3, "value": m2.total_atoms * 20 / m2.wall,
                                          "us_per_dipole_iteration": float(m2.sweep[0]) / max(float(m2.sweep[1]), 1.0) * 1e3,
                                          "barrier_ms_per_step_rank0": (m2.barrier_ms or 0.0) / 20}
        allp = [None] * world if rank == 0 else None
        dist.gather_object(parts, allp, dst=0)
        run.close()
        if rank == 0:
            g = c.sys
            MU, EF, F = np.zeros((g.n, 3)), np.zeros((g.n, 3)), np.zeros((g.n, 3))
            for p_ in allp:
                MU[p_["own"]], EF[p_["own"]], F[p_["own"]] = p_["mu"], p_["ef"], p_["f"]
            one = Run(pb, torch, dist, c, 0, 1, local, comm=False)
            one.pin_host()
            for k in range(2):
                r1 = one.step_host(k)
            rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
            e = dict(mu=rel(MU, one.p["mu"]), ef=rel(EF, one.p["ef"]), f=rel(F, one.p["f"]),
                     iterations=[int(allp[0]["it"]), int(r1.iterations)])
            one.close()
            detail[label] = e
            if label == "fixed30":  # Jacobi: per-iteration parity; the Gauss-Seidel colouring differs between decompositions
                worst = max(worst, e["mu"], e["ef"], e["f"])
                assert e["iterations"][0] == e["iterations"][1]
            else:
                worst = max(worst, e["ef"])
                detail[label]["mu_abs"] = float(np.abs(MU - one.p["mu"]).max())
                # tolerance parity of the GS modes (BASELINE.json): the two runs use different colourings (colours are
                # chosen per brick) and each stops within ~10*precision of the common fixed point
                assert detail[label]["mu_abs"] < 100 * 1e-11, detail
        dist.barrier()
    if rank == 0:
        assert worst < 1e-10, f"multi-GPU parity failed: {detail}"
    return worst, detail


def same_work_on_gpu(pb, local, ref):
    """the reference arm's sample with the reference's semantics (all-pairs dipoles, no polar_cutoff) on the GPU"""
    c = make_config(2, 1)
    c.sys = workloads().lj_charge_fluid(SAMPLE_NCELL)
    s = make_style(pb, c, local, words=STYLE_WORDS)
    n = c.sys.n
    mu, f, ef = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
    h = [np.ascontiguousarray(a) for a in (c.sys.x, c.sys.q, c.sys.type, c.sys.alpha)]
    for k in range(3):
        s.compute(h[0], h[1], h[2], h[3], mu, f, ef_static=ef, ago=0)
    t0 = time.perf_counter()
    nrep = 10
    for k in range(nrep):
        mu[:] = 0.0
        r = s.compute(h[0], h[1], h[2], h[3], mu, f, ef_static=ef, ago=0 if k % REBUILD_EVERY == 0 else 1)
    dt = (time.perf_counter() - t0) / nrep
    s.close()
    out = {"atoms": n, "gpu_ms_per_step": dt * 1e3, "gpu_value": n / dt, "semantics": "exact mode: all minimum-image pairs, host buffers",
           "gpu_eng_pol": r.eng_pol}
    if ref:
        out["speedup_vs_reference"] = (n / dt) / ref["value"]
        if ref.get("epol_step0") is not None:
            out["reference_eng_pol_step0"] = ref["epol_step0"]
    return out


def gpu_arm(args):
    import torch
    import torch.distributed as dist

    pb = load_pb()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world not in GRIDS:
        raise SystemExit(f"bench.py: no brick grid defined for {world} GPUs (use 1, 2, 4 or 8)")
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())

    # multi-GPU parity before the timed run (every rank takes part)
    mgpu = None
    if world > 1 and not args.no_parity:
        mgpu = mgpu_parity(pb, torch, dist, rank, world, local)

    c = make_config(args.config, world)
    run = Run(pb, torch, dist, c, rank, world, local)
    n = run.n
    m = run.reduce(run.measure(args.steps, args.warmup, clocks=True))
    style = run.style
    comm_stats = m.comm_stats

    also = {}
    if world == 1 and args.config == 5 and not args.no_also:
        run.close()
        run_alive = False
        for cfg in (2, 3):
            r2 = Run(pb, torch, dist, make_config(cfg, 1), rank, 1, local)
            m2 = r2.reduce(r2.measure(args.steps, args.warmup))
            also[f"config{cfg}"] = summary(r2, m2, args.steps, peaks)
            r2.close()
    else:
        run_alive = True

    if rank == 0:
        assert m.total_atoms == c.sys.n
        value = m.total_atoms * args.steps / m.wall
        e2e_value = m.total_atoms * args.steps / m.wall_e2e
        roof = run.roofline(m, peaks, ROOT / "profiles" / "ncu_traffic.json")
        # CPU baseline on a bounded sample (rank 0, N=1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ref = run_reference_binary(1, 1)
            if ref and "value" in ref:
                cpu = {"value": ref["value"], "unit": UNIT, "cores": 1, "kind": "reference",
                       "sample": f"oracle/_ref/lmp_serial, {ref['steps']} steps, {ref['s_per_step']:.3f} s/step, on a "
                                 + sample_description(ref["atoms"])}
                cpu["same_work"] = same_work_on_gpu(pb, local, ref)
            port = run_port_sample(target_s=8.0)
            cpu_port = {"value": port["value"], "unit": UNIT, "cores": port["threads"], "kind": "port",
                        "sample": f"oracle port (list algorithm with polar_cutoff {CUT_COUL}, OpenMP), {port['rows']} of the 32000 rows "
                                  "of BASELINE config 2, polarization stages"}
            if cpu is None:
                cpu = cpu_port
            else:
                cpu["port"] = cpu_port
        pg = c.pg
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m.wall / args.steps * 1e3, "higher_is_better": True,
            "scaling": c.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(c, world, "single" if world == 1 else
                                      f"spatial decomposition, {pg[0]}x{pg[1]}x{pg[2]} bricks, one process per GPU; ghost positions "
                                      f"once per step, ghost dipoles once per sweep ("
                                      + ("stored by the sweep kernel into peer memory over NVLink + signal/wait barrier"
                                         if int(comm_stats[3]) else "NCCL send/recv") + ")"),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 64 * n, "d2h_bytes_per_step": 72 * n + 192,
                    "ms_per_step": m.wall_e2e / args.steps * 1e3},
            "gpu_launches": int(m.launches),
            "us_per_dipole_iteration": roof["launch_ms"] * 1e3,
            "iterations_per_step": m.iterations,
            "device_ms_per_step": m.dev_ms,
            "stage_ms": {"neigh_refresh": m.stage[0], "pair_field": m.stage[1], "scf": m.stage[2], "pol_force": m.stage[3]},
            "rebuild_steps_in_timed_region": len(m.rebuild_ms),
            "rebuild_ms_rank0": float(np.mean(m.rebuild_ms)) if m.rebuild_ms else None,
            "roofline": roof,
            "cpu_baseline": cpu, "clocks": m.clocks,
            "halo": None if world == 1 else {"rank0_owned": n, "rank0_send_slots": int(comm_stats[0]),
                                             "rank0_ghosts": int(comm_stats[1]), "bytes_per_sweep_rank0": 32 * int(comm_stats[0]),
                                             "model_24B_x_ghosts": 24 * int(comm_stats[1]),
                                             "peer_push": bool(int(comm_stats[3])),
                                             # NVLink bytes of rank 0's GPU over the timed region per dipole sweep (nvidia-smi
                                             # nvlink counters; includes the once-per-step position halo and NCCL traffic)
                                             "nvlink_bytes_per_sweep_rank0": None if not m.nvlink else
                                             {k: v / max(float(m.sweep[1]), 1.0) for k, v in m.nvlink.items()},
                                             "per_rank": m.per_rank},
            "check": {"eng_pol": m.eng_pol_total, "iterations": m.res.iterations,
                      "mgpu_max_rel_err": mgpu[0] if mgpu else None, "mgpu_detail": mgpu[1] if mgpu else None},
            "also": (also or None) if world == 1 else ({"config2_weak": mgpu[1].get("config2_weak")} if mgpu else None),
        }
        print(json.dumps(line))
    if run_alive:
        run.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=5, choices=[2, 3, 4, 5])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU parity stage (N > 1)")
    ap.add_argument("--no-also", action="store_true", help="N = 1: skip the secondary measurements of configs 2 and 3")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 behind Python's back
    # (NCCL prints its version banner there when the first communicator comes up) are sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(json_fd, "w", buffering=1)
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
