#!/usr/bin/env python3
"""A whole MD step resident on one B200: rigid-body integrator + polarization pair style + reciprocal-space Ewald,
all through the C ABI with DEVICE pointers (on_device = 1) -- positions, velocities, forces and induced dipoles never
leave HBM between steps.  This is SURVEY §8f rank 2's point: with the integrator on the device the per-step host
round trip of x and f (24 N bytes each way per consumer) disappears.

Workload: BASELINE config 3's rigid polarizable water box (workloads.water_box), `fix rigid/nve molecule`,
pair_style lj/cut/coul/long/polarization (precision 1e-11, polar_gs_ranked, polar_gamma 1.03, use_previous),
kspace_style ewald 1e-4.  Prints one JSON line: ms per step by stage (CUDA events of each library call), the same
loop through HOST buffers (what a LAMMPS Fix/Pair/KSpace triple pays) for comparison, and the conserved energy
KE_trans + KE_rot + E_vdwl + E_coul + E_long + E_pol over the run (an NVE trajectory must hold it).

usage: md_resident.py [nside=44] [steps=40] [dt=1.0] [--host-too] [--pppm]
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import bench

pb = bench.load_pb()
CUT = 12.0
REBUILD = 10


def water_topology(n):
    """special lists of 3-site molecules O H H (ids 3m+1, 3m+2, 3m+3): O: 1-2 = {H,H}; H: 1-2 = {O}, 1-3 = {other H}"""
    nspecial = np.zeros((n, 3), dtype=np.int32)
    special = np.zeros((n, 2), dtype=np.int32)
    ids = np.arange(1, n + 1, dtype=np.int32).reshape(-1, 3)
    o, h1, h2 = ids[:, 0], ids[:, 1], ids[:, 2]
    nspecial[0::3] = (2, 2, 2)
    special[0::3, 0], special[0::3, 1] = h1, h2
    nspecial[1::3] = (1, 2, 2)
    special[1::3, 0], special[1::3, 1] = o, h2
    nspecial[2::3] = (1, 2, 2)
    special[2::3, 0], special[2::3, 1] = o, h1
    return nspecial, special


def build(nside):
    s = bench.workloads().water_box(nside)
    n = s.n
    L = s.boxhi - s.boxlo
    first = s.x[0::3].repeat(3, axis=0)
    image = -np.rint((s.x - first) / L).astype(np.int64)   # molecules were wrapped atom by atom
    mass = np.where(s.type == 1, 15.9994, 1.008)
    rng = np.random.default_rng(99)
    v = rng.normal(size=(n, 3)) * np.sqrt(pb.REAL_BOLTZ * 298.15 / (mass[:, None] * pb.REAL_MVV2E))
    return s, image, mass, v


def make_pair(s, g_ewald):
    p = pb.PairStyle(device=0)
    p.set_ntypes(2)
    p.command(f"pair_style lj/cut/coul/long/polarization 2.5 {CUT} precision 1e-11 max_iterations 200 polar_gamma 1.03 "
              f"damp_type exponential use_previous yes polar_cutoff {CUT}")
    p.command("pair_coeff 1 1 0.155 3.166 10.0")   # a physical O-O LJ cutoff (the bench configs' 2.5 A leaves no repulsion)
    p.command("pair_coeff 2 2 0.0 1.0")
    p.init(g_ewald=g_ewald, special_lj=(1.0, 0.0, 0.0, 0.0), special_coul=(1.0, 0.0, 0.0, 0.0), molecular=1)
    p.set_box(s.boxlo, s.boxhi)
    return p


def run(nside, steps, dt, resident, budget_s=150.0, kspace="ewald"):
    t_start = time.time()
    s, image, mass, v0 = build(nside)
    n = s.n
    L = s.boxhi - s.boxlo
    nspecial, special = water_topology(n)
    ew = pb.PPPM(device=0) if kspace == "pppm" else pb.Ewald(device=0)   # same interface: init -> g_ewald, compute(_device)
    kinfo = ew.init(1e-4, s.q, CUT, s.boxlo, s.boxhi)
    g = kinfo.g_ewald
    pair = make_pair(s, g)
    rig = pb.Rigid(device=0)
    info = rig.init(s.tag, s.molecule, mass, image, s.x, v0, s.boxlo, s.boxhi, dt)
    dev = torch.device("cuda", 0)
    host = dict(x=np.ascontiguousarray(s.x), v=np.ascontiguousarray(v0), f=np.zeros((n, 3)), mu=np.zeros((n, 3)),
                q=np.ascontiguousarray(s.q), type=np.ascontiguousarray(s.type), molecule=np.ascontiguousarray(s.molecule),
                tag=np.ascontiguousarray(s.tag), alpha=np.ascontiguousarray(s.alpha), nspecial=nspecial, special=special)
    img = image.copy()
    if resident:
        d = {k: torch.from_numpy(a).to(dev) for k, a in host.items()}
        dimg = torch.from_numpy(img).to(dev)
        lo, prd = torch.tensor(s.boxlo, device=dev), torch.tensor(L, device=dev)
        ptrs = {k: d[k].data_ptr() for k in ("x", "q", "type", "molecule", "tag", "alpha", "mu", "f", "nspecial", "special")}
        torch.cuda.synchronize()

    def forces(k):
        ago = k % REBUILD
        if resident:
            if ago == 0:  # Domain::pbc + FixRigid::pre_neighbor: wrap the atoms, hand the new image flags to the integrator
                shift = torch.floor((d["x"] - lo) / prd)
                d["x"] -= shift * prd
                dimg.add_(shift.to(torch.int64))
                packed = (((dimg[:, 2] + 512) << 20) | ((dimg[:, 1] + 512) << 10) | (dimg[:, 0] + 512)).to(torch.int32).contiguous()
                torch.cuda.synchronize()   # the library calls run on their own streams
                rig.pre_neighbor_device(n, d["tag"].data_ptr(), packed.data_ptr())
            d["f"].zero_()
            torch.cuda.synchronize()
            r = pair.compute_device(n, ptrs, eflag=1, vflag=0, ago=ago, maxspecial=2)
            elong, _ = ew.compute_device(n, d["x"].data_ptr(), d["q"].data_ptr(), d["f"].data_ptr(), eflag=1, vflag=0)
        else:
            if ago == 0:
                shift = np.floor((host["x"] - s.boxlo) / L)
                host["x"] -= shift * L
                img[:] += shift.astype(np.int64)
                rig.pre_neighbor(host["tag"], img)
            host["f"][:] = 0.0
            r = pair.compute(host["x"], host["q"], host["type"], host["alpha"], host["mu"], host["f"], molecule=host["molecule"],
                             tag=host["tag"], nspecial=nspecial, special=special, eflag=1, vflag=0, ago=ago)
            elong, _ = ew.compute(host["x"], host["q"], host["f"], eflag=1, vflag=0)
        return r, elong, pair_ms(r), ew.last_ms()

    def pair_ms(r):
        return r.ms_total

    def rigid_call(which):
        if resident:
            a = (n, d["tag"].data_ptr(), d["x"].data_ptr(), d["v"].data_ptr(), d["f"].data_ptr())
            if which == "setup":
                rig.setup_device(*a, vflag=0)
            elif which == "initial":
                rig.initial_integrate_device(*a, vflag=0)
            else:
                rig.final_integrate_device(*a)
        else:
            a = (host["tag"], host["x"], host["v"], host["f"])
            if which == "setup":
                rig.setup(*a, vflag=0)
            elif which == "initial":
                rig.initial_integrate(*a, vflag=0)
            else:
                rig.final_integrate(*a)
        return rig.last_ms()

    def energy(r, elong):
        _, ket, ker = rig.scalars()
        ke = (ket + ker) * pb.REAL_MVV2E
        return dict(ke=ke, evdwl=r.eng_vdwl, ecoul=r.eng_coul, elong=elong, epol=r.eng_pol,
                    etotal=ke + r.eng_vdwl + r.eng_coul + elong + r.eng_pol, iterations=r.iterations)

    print(f"md_resident: {n} atoms, g_ewald {g:.5f}, setting up ({'device' if resident else 'host'} buffers)", file=sys.stderr, flush=True)
    r, elong, _, _ = forces(0)
    print(f"first force call: {r.ms_total:.1f} ms pair ({r.iterations} SCF iterations), {ew.last_ms():.1f} ms {kspace}", file=sys.stderr, flush=True)
    rigid_call("setup")
    rows = [energy(r, elong)]
    stage = dict(rigid=[], pair=[], kspace=[], wall=[])   # (profiles written before the PPPM path call the third stage "ewald")
    for k in range(1, steps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ms_r = rigid_call("initial")
        r, elong, ms_p, ms_e = forces(k)
        ms_r += rigid_call("final")
        torch.cuda.synchronize()
        stage["wall"].append((time.perf_counter() - t0) * 1e3)
        stage["rigid"].append(ms_r), stage["pair"].append(ms_p), stage["kspace"].append(ms_e)
        rows.append(energy(r, elong))
        if not np.isfinite(rows[-1]["etotal"]) or time.time() - t_start > budget_s:
            print(f"md_resident: stopping at step {k}: etotal {rows[-1]['etotal']}, {time.time() - t_start:.0f} s", file=sys.stderr)
            steps = k
            break
        if k % 10 == 0:
            print(f"step {k}: etotal {rows[-1]['etotal']:.6f} ke {rows[-1]['ke']:.4f} it {rows[-1]['iterations']} "
                  f"wall {stage['wall'][-1]:.1f} ms", file=sys.stderr, flush=True)
    out = dict(resident=resident, kspace=kspace + (f" grid {kinfo.nx}x{kinfo.ny}x{kinfo.nz} order {kinfo.order}" if kspace == "pppm"
                                                   else f" kcount {kinfo.kcount}"), atoms=n, bodies=info.nbody, steps=steps, dt_fs=dt, rebuild_every=REBUILD,
               ms_per_step={k: float(np.mean(v[2:])) for k, v in stage.items()},
               ms_per_step_no_rebuild={k: float(np.mean([t for i, t in enumerate(v, start=1) if i % REBUILD])) for k, v in stage.items()},
               atom_steps_per_s=n / (float(np.mean(stage["wall"][2:])) * 1e-3),
               etotal_first=rows[0]["etotal"], etotal_last=rows[-1]["etotal"], ke_first=rows[0]["ke"], ke_last=rows[-1]["ke"],
               etotal_max_abs_drift=float(max(abs(rw["etotal"] - rows[0]["etotal"]) for rw in rows)),
               scf_iterations=[rw["iterations"] for rw in rows], rigid_launches=rig.launch_count(),
               energies=[{k: float(v) for k, v in rw.items()} for rw in rows[:: max(1, steps // 8)]])
    pair.close(), ew.close(), rig.close()
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    nside = int(args[0]) if len(args) > 0 else 44
    steps = int(args[1]) if len(args) > 1 else 40
    dt = float(args[2]) if len(args) > 2 else 1.0
    kspace = "pppm" if "--pppm" in sys.argv else "ewald"
    res = run(nside, steps, dt, resident=True, kspace=kspace)
    line = dict(what="md_resident", workload=f"rigid polarizable water box, {res['atoms']} atoms, rigid/nve + polarization pair "
                f"style (precision 1e-11 GS-ranked) + {kspace} 1e-4, all device-resident through the C ABI", resident=res)
    if "--host-too" in sys.argv:
        line["host_buffers"] = run(nside, max(12, steps // 3), dt, resident=False, kspace=kspace)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
