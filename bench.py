#!/usr/bin/env python3
"""bench.py -- atom-steps/s (incl. the dipole SCF) of the lj/cut/coul/long/polarization hot path.

Workload (BASELINE.json configs[1]): synthetic polarizable LJ+charge fluid, 32 000 atoms, rho = 0.1 /A^3,
pair_style lj/cut/coul/long/polarization 2.5 12 fixed_iteration yes max_iterations 30
damp_type exponential polar_gs_ranked no, dipole-dipole over the neighbor list (polar_cutoff 12).
A "step" = one compute() call = one pass of the whole hot path (neighbor refresh, LJ+Coulomb+static field,
30 Jacobi dipole sweeps, polarization forces, reductions); the device structures are rebuilt every
REBUILD_EVERY steps inside the timed region, as LAMMPS' delay-10 schedule would.

  value  : whole-job atom-steps/s with inputs resident in HBM (device pointers through the C ABI)
  e2e    : same metric through the C ABI with HOST buffers (pinned), H2D of x/mu and D2H of f/mu/E inside
  roofline: dominant kernel k_sweep_cached (one dipole iteration): bytes it must move through HBM per launch
            (20 B per pair streamed + the atom records once) over the CUDA-event duration of its
            launches, against MEASURED_PEAKS.json hbm_gbs; SURVEY §8d's every-gather-from-HBM model beside it
  cpu_baseline: the reference binary (oracle/_ref/lmp_serial, 1 core: it is serial by design) on a bounded
            2048-atom sample of the same fluid, else the oracle port on all host threads
  N > 1  : ONE periodic system of 32000*N atoms, spatially decomposed into N bricks (one process per GPU):
            ghost positions once per step, ghost dipoles once per SCF sweep (weak scaling)
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

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "atom-steps/s incl. dipole SCF"
UNIT = "atom-steps/s"
REBUILD_EVERY = 10
NCELL = 20  # 4*20^3 = 32000 atoms
CUT_LJ, CUT_COUL, ITER = 2.5, 12.0, 30
SAMPLE_NCELL = 8  # 2048-atom sample for the serial reference binary
STYLE_WORDS = (f"{CUT_LJ} {CUT_COUL} polar_gs_ranked no fixed_iteration yes max_iterations {ITER} "
               "damp_type exponential")


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


def workload_config(n_gpus, parallelism):
    return {"workload": f"synthetic polarizable LJ+charge fluid, {4 * NCELL ** 3} atoms per GPU "
                        f"({4 * NCELL ** 3 * n_gpus} in total), rho 0.1/A^3, "
                        f"cut {CUT_LJ}/{CUT_COUL}, fixed_iteration yes max_iterations {ITER}, damp_type exponential, "
                        f"polar_gs_ranked no, polar_cutoff {CUT_COUL} (neighbor-list dipole sweep)",
            "atoms_per_gpu": 4 * NCELL ** 3, "atoms_total": 4 * NCELL ** 3 * n_gpus, "sweeps_per_step": ITER,
            "rebuild_every": REBUILD_EVERY, "parallelism": parallelism,
            "l2_policy": "working set (neighbor list + radial cache ~600 MB per GPU) exceeds the 126 MB L2"}


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


def run_reference_binary(steps, warmup):
    """Times oracle/_ref/lmp_serial (the repaired, otherwise unmodified reference) on the 2048-atom sample."""
    lmp = ROOT / "oracle" / "_ref" / "lmp_serial"
    if not lmp.exists():
        return None
    sysm = workloads().lj_charge_fluid(SAMPLE_NCELL)
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
        return {"atoms": sysm.n, "s_per_step": per_step, "value": sysm.n / per_step, "wall_s": wall,
                "steps": nsteps}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def run_port_sample(nthreads=0, target_s=10.0):
    """Oracle port (truncated list algorithm, OpenMP) on a row sample of the full 32k workload."""
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
        sample = (f"oracle/_ref/lmp_serial (repaired reference, serial by design) on a {ref['atoms']}-atom sample of the "
                  f"same fluid (same density/cutoffs/keywords), {ref['steps']} MD steps, Pair timer; the reference is "
                  f"O(N^2) with a dense 3Nx3N matrix, so its per-atom cost at 32000 atoms would be ~{(4 * NCELL ** 3 / ref['atoms']):.0f}x higher "
                  "and needs 74 GB")
        ms = ref["s_per_step"] * 1e3
    else:
        port = run_port_sample()
        kind, cores, value = "port", port["threads"], port["value"]
        sample = f"oracle port rows sample: {port['rows']} of 32000 atoms, polarization stages only"
        ms = port["seconds"] * 1e3
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus, "cpu"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def make_style(pb, sysm, device):
    # g_ewald as `kspace_style ewald 1e-4` would set it: the product's own Ewald::init (polb200_ewald_init)
    ew = pb.Ewald(device=device)
    g = ew.init(1e-4, sysm.q, CUT_COUL, sysm.boxlo, sysm.boxhi).g_ewald
    ew.close()
    s = pb.PairStyle(device=device)
    s.set_ntypes(2)
    s.command(f"pair_style lj/cut/coul/long/polarization {STYLE_WORDS} polar_cutoff {CUT_COUL}")
    s.command("pair_coeff 1 1 0.1 3.0")
    s.command("pair_coeff 2 2 0.1 3.0")
    s.init(g_ewald=g, molecular=0)
    s.set_box(sysm.boxlo, sysm.boxhi)
    return s


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

    # N GPUs: ONE periodic system of 32000*N atoms cut into N bricks (spatial decomposition, weak scaling);
    # every rank generates the same global system and keeps the atoms of its own brick
    pg = GRIDS.get(world)
    if pg is None:
        raise SystemExit(f"bench.py: no brick grid defined for {world} GPUs (use 1, 2, 4 or 8)")
    gsys = workloads().lj_charge_fluid(NCELL if world == 1 else tuple(NCELL * np.array(pg)), seed=12345)
    style = make_style(pb, gsys, local)
    dev = torch.device("cuda", local)
    if world > 1:
        box = [pb.comm_create_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        style.comm_init(rank, world, box[0], pg)
        lo, hi = style.subdomain()
        own = np.nonzero(np.all((gsys.x >= lo) & (gsys.x < hi), axis=1))[0]
    else:
        own = np.arange(gsys.n)

    class Owned:
        pass
    sysm = Owned()
    sysm.x, sysm.q, sysm.type, sysm.alpha, sysm.tag = (np.ascontiguousarray(gsys.x[own]), np.ascontiguousarray(gsys.q[own]),
                                                       np.ascontiguousarray(gsys.type[own]), np.ascontiguousarray(gsys.alpha[own]),
                                                       np.ascontiguousarray(gsys.tag[own]))
    n = sysm.n = len(own)

    # device-resident inputs
    t_x = torch.tensor(sysm.x, dtype=torch.float64, device=dev).contiguous()
    t_q = torch.tensor(sysm.q, dtype=torch.float64, device=dev)
    t_type = torch.tensor(sysm.type, dtype=torch.int32, device=dev)
    t_alpha = torch.tensor(sysm.alpha, dtype=torch.float64, device=dev)
    t_tag = torch.tensor(sysm.tag, dtype=torch.int32, device=dev)
    t_mu = torch.zeros((n, 3), dtype=torch.float64, device=dev)
    t_f = torch.zeros((n, 3), dtype=torch.float64, device=dev)
    t_ef = torch.zeros((n, 3), dtype=torch.float64, device=dev)
    ptrs = dict(x=t_x.data_ptr(), q=t_q.data_ptr(), type=t_type.data_ptr(), alpha=t_alpha.data_ptr(),
                tag=t_tag.data_ptr(), mu=t_mu.data_ptr(), f=t_f.data_ptr(), ef_static=t_ef.data_ptr())
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev(k):
        t_f.zero_()
        torch.cuda.synchronize()
        return style.compute_device(n, ptrs, eflag=1, vflag=2, ago=k % REBUILD_EVERY)

    # ---- resident-in-HBM throughput ----
    for k in range(args.warmup):
        res = step_dev(k)
    polar_pairs = int(style.debug_fetch("polar_pairs", np.uint64, 1)[0])
    group_stats = style.debug_fetch("group_stats", np.float64, 4)
    comm_stats = style.debug_fetch("comm_stats", np.float64, 5)
    pb.lib().polb200_set_option(style._h, b"time_sweeps", 1.0)
    style.launch_count(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    stage = np.zeros(5)
    for k in range(args.steps):
        res = step_dev(k)
        dev_ms += res.ms_total
        stage += [res.ms_neigh, res.ms_pair, res.ms_scf, res.ms_force, res.ms_total]
    barrier()
    wall = time.perf_counter() - t0
    launches = style.launch_count()
    sweep = style.debug_fetch("sweep_timing", np.float64, 2)
    pb.lib().polb200_set_option(style._h, b"time_sweeps", 0.0)
    clocks = sampler.stop() if rank == 0 else None
    eng_pol = res.eng_pol

    # ---- end to end through the C ABI with host (pinned) buffers ----
    h_x = torch.tensor(sysm.x, dtype=torch.float64).pin_memory().numpy()
    h_mu = torch.zeros((n, 3), dtype=torch.float64).pin_memory().numpy()
    h_f = torch.zeros((n, 3), dtype=torch.float64).pin_memory().numpy()
    h_ef = torch.zeros((n, 3), dtype=torch.float64).pin_memory().numpy()
    h_q = np.ascontiguousarray(sysm.q)
    h_type = np.ascontiguousarray(sysm.type)
    h_alpha = np.ascontiguousarray(sysm.alpha)
    h_tag = np.ascontiguousarray(sysm.tag)

    def step_host(k):
        h_f[:] = 0.0
        return style.compute(h_x, h_q, h_type, h_alpha, h_mu, h_f, tag=h_tag, ef_static=h_ef, eflag=1, vflag=2,
                             ago=k % REBUILD_EVERY)

    for k in range(max(1, args.warmup // 2)):
        step_host(k)
    barrier()
    t1 = time.perf_counter()
    for k in range(args.steps):
        r2 = step_host(k)
    barrier()
    wall_e2e = time.perf_counter() - t1
    assert abs(r2.eng_pol - eng_pol) <= 1e-9 * abs(eng_pol)

    # max over ranks (times), sum over ranks (atoms, energy: per-rank partials by LAMMPS convention)
    times = torch.tensor([wall, wall_e2e], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(n), eng_pol, float(polar_pairs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums)
    wall, wall_e2e = float(times[0]), float(times[1])
    eng_pol_total = float(sums[1])

    if rank == 0:
        total_atoms = int(sums[0])
        assert total_atoms == gsys.n
        value = total_atoms * args.steps / wall
        e2e_value = total_atoms * args.steps / wall_e2e
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        sweep_ms = float(sweep[0]) / max(float(sweep[1]), 1.0)
        # Bytes one launch of the dominant kernel must move through HBM.  Default kernel
        # k_sweep_group_tma: one warp per pair group (two cell-row neighbours); per group-row entry it streams,
        # exactly once, 4 B of neighbour index + 32 B of cached radial scalars {s1a,s2a,s1b,s2b}; the 32-B position
        # and dipole records of the owned+ghost atoms are gathered hundreds of times each but from L2/L1, so they
        # count once; plus E_static in and the new dipole out per owned atom.  (Per-atom fallback kernel
        # k_sweep_cached: 20 B per pair instead.)
        nghost = int(res.nghost)
        grouped = bool(group_stats[3]) and group_stats[1] > 0
        entries = float(group_stats[1]) if grouped else float(polar_pairs)
        alg_bytes = (36.0 if grouped else 20.0) * entries + 64.0 * (n + nghost) + 64.0 * n
        achieved = alg_bytes / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None
        # measured DRAM traffic of the same kernel on the same workload (one `ncu --set full` capture, per launch)
        traffic = None
        tf = ROOT / "profiles" / "ncu_traffic.json"
        if tf.exists() and world == 1:
            t = json.loads(tf.read_text())
            if t.get("kernel") == ("k_sweep_group_tma" if grouped else "k_sweep_cached") and t.get("atoms") == n:
                traffic = t["dram_bytes_per_launch"]
        # SURVEY §8d's matrix-free model (every gather charged to HBM): 52 B per pair + 104 B per atom
        survey_bytes = 52.0 * polar_pairs + 104.0 * n
        # CPU baseline on a bounded sample (rank 0, N=1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ref = run_reference_binary(1, 1)
            if ref and "value" in ref:
                cpu = {"value": ref["value"], "unit": UNIT, "cores": 1, "kind": "reference",
                       "sample": f"oracle/_ref/lmp_serial on a {ref['atoms']}-atom sample of the same fluid, "
                                 f"{ref['steps']} steps, {ref['s_per_step']:.3f} s/step (reference is O(N^2): ~"
                                 f"{4 * NCELL ** 3 // ref['atoms']}x more per atom at 32000 atoms)"}
            port = run_port_sample(target_s=8.0)
            cpu_port = {"value": port["value"], "unit": UNIT, "cores": port["threads"], "kind": "port",
                        "sample": f"oracle port (list algorithm, OpenMP), {port['rows']} of 32000 rows, polarization stages"}
            if cpu is None:
                cpu = cpu_port
            else:
                cpu["port"] = cpu_port
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, "single" if world == 1 else
                                      f"spatial decomposition, {pg[0]}x{pg[1]}x{pg[2]} bricks, one process per GPU; ghost positions "
                                      f"once per step (NCCL), ghost dipoles once per sweep ("
                                      + ("stored by the sweep kernel into peer memory over NVLink + signal/wait barrier"
                                         if int(comm_stats[3]) else "NCCL send/recv") + ")"),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 48 * n, "d2h_bytes_per_step": 72 * n + 192,
                    "ms_per_step": wall_e2e / args.steps * 1e3},
            "gpu_launches": int(launches),
            "us_per_dipole_iteration": sweep_ms * 1e3,
            "device_ms_per_step": dev_ms / args.steps,
            "stage_ms": {"neigh_refresh": stage[0] / args.steps, "pair_field": stage[1] / args.steps,
                         "scf": stage[2] / args.steps, "pol_force": stage[3] / args.steps},
            "roofline": {"bound": "hbm", "kernel": ("k_sweep_group_tma" if grouped else "k_sweep_cached") +
                                                    " (one dipole iteration over the neighbor list)",
                         "achieved": achieved, "peak": peak_gbs,
                         "unit": "GB/s", "frac": achieved / peak_gbs if achieved else None, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "bytes_model": ("36 B per group-row entry streamed (index + cached radial scalars of both members)"
                                         if grouped else "20 B/pair streamed (index + cached radial scalars)") +
                                        " + 64 B per owned+ghost atom (position and dipole records, read once) + 64 B per "
                                        "owned atom (E_static in, dipole out)",
                         "group_row_entries": entries if grouped else None,
                         "pairs_in_cutoff": polar_pairs, "launch_ms": sweep_ms,
                         "gpairs_per_s": polar_pairs / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None,
                         "survey_8d_model": {"bytes": survey_bytes,
                                             "achieved": survey_bytes / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None,
                                             "note": "52 B/pair + 104 B/atom with every neighbour gather charged to HBM; "
                                                     "exceeds the HBM peak because the gathers are L2/L1 hits"}},
            "cpu_baseline": cpu, "clocks": clocks,
            "halo": None if world == 1 else {"rank0_owned": n, "rank0_send_slots": int(comm_stats[0]),
                                             "rank0_ghosts": int(comm_stats[1]), "bytes_per_sweep_rank0": 32 * int(comm_stats[0]),
                                             "peer_push": bool(int(comm_stats[3]))},
            "check": {"eng_pol": eng_pol_total, "iterations": res.iterations},
        }
        print(json.dumps(line))
    style.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
