#!/usr/bin/env python3
"""Time a complete decomposed MD step (rigid integrator + polarization pair style + PPPM / Ewald) on N GPUs, device-resident.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/mgpu_md_timing.py [nside] [steps] [--ewald]

Workload: workloads.water_box(nside) (3 * nside^3 atoms in a cube), cut into the px x py x pz bricks of bench.py; every
rank keeps the atoms of its brick on its GPU and calls the three decomposed handles with device pointers
(polb200_comm_init, polb200_pppm_comm_init / polb200_ewald_comm_init, polb200_rigid_comm_init).  Step 0 builds the lists;
steps 1..steps-1 re-use them (`ago` = step, no atom changes brick inside the 2 A skin over such a short run), so the timed
region holds no re-neighboring.  Rank 0 prints one JSON line: ms per step by stage (CUDA-event times of each library call,
max over ranks) and wall ms per step (barrier to barrier).  N = 1 runs the same loop through undecomposed handles.
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import torch
import torch.distributed as dist
import bench
from md_resident import water_topology

pb = bench.load_pb()
CUT = 12.0
GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    nside = int(args[0]) if args else 44
    steps = int(args[1]) if len(args) > 1 else 10
    kspace = "ewald" if "--ewald" in sys.argv else "pppm"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    pg = GRIDS[world]
    s = bench.workloads().water_box(nside)
    n = s.n
    L = s.boxhi - s.boxlo
    first = s.x[0::3].repeat(3, axis=0)
    image = -np.rint((s.x - first) / L).astype(np.int64)
    mass = np.where(s.type == 1, 15.9994, 1.008)
    v0 = np.random.default_rng(99).normal(size=(n, 3)) * np.sqrt(pb.REAL_BOLTZ * 298.15 / (mass[:, None] * pb.REAL_MVV2E))
    nspecial, special = water_topology(n)

    def fresh_ids(k):
        box = [[pb.comm_create_id() for _ in range(k)] if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    ks = pb.PPPM(device=local) if kspace == "pppm" else pb.Ewald(device=local)
    pair = pb.PairStyle(device=local)
    rig = pb.Rigid(device=local)
    if world > 1:
        ids = fresh_ids(3)
        ks.comm_init(rank, world, ids[1])
        rig.comm_init(rank, world, ids[2])
    kinfo = ks.init(1e-4, s.q, CUT, s.boxlo, s.boxhi)
    pair.set_ntypes(2)
    pair.command(f"pair_style lj/cut/coul/long/polarization 2.5 {CUT} precision 1e-11 max_iterations 200 polar_gamma 1.03 "
                 f"damp_type exponential use_previous yes polar_cutoff {CUT}")
    pair.command("pair_coeff 1 1 0.155 3.166 10.0")
    pair.command("pair_coeff 2 2 0.0 1.0")
    pair.init(g_ewald=kinfo.g_ewald, special_lj=(1.0, 0.0, 0.0, 0.0), special_coul=(1.0, 0.0, 0.0, 0.0), molecular=1)
    pair.set_box(s.boxlo, s.boxhi)
    if world > 1:
        pair.comm_init(rank, world, ids[0], pg)
        lo, hi = pair.subdomain()
        idx = np.nonzero(np.all((s.x >= lo) & (s.x < hi), axis=1))[0]
    else:
        idx = np.arange(n)
    nl = len(idx)
    info = rig.init(s.tag[idx], s.molecule[idx], mass[idx], image[idx], s.x[idx], v0[idx], s.boxlo, s.boxhi, 1.0)
    host = dict(x=s.x[idx], v=v0[idx], f=np.zeros((nl, 3)), mu=np.zeros((nl, 3)), q=s.q[idx], type=s.type[idx], molecule=s.molecule[idx],
                tag=s.tag[idx], alpha=s.alpha[idx], nspecial=nspecial[idx], special=special[idx])
    d = {k: torch.from_numpy(np.ascontiguousarray(a)).to(dev) for k, a in host.items()}
    ptrs = {k: d[k].data_ptr() for k in ("x", "q", "type", "molecule", "tag", "alpha", "mu", "f", "nspecial", "special")}
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def forces(ago):
        d["f"].zero_()
        torch.cuda.synchronize()
        r = pair.compute_device(nl, ptrs, eflag=1, vflag=0, ago=ago, maxspecial=2)
        t_pair = r.ms_total
        ks.compute_device(nl, d["x"].data_ptr(), d["q"].data_ptr(), d["f"].data_ptr(), eflag=1, vflag=0)
        return r, t_pair, ks.last_ms()

    a = (nl, d["tag"].data_ptr(), d["x"].data_ptr(), d["v"].data_ptr(), d["f"].data_ptr())
    r, _, _ = forces(0)
    rig.setup_device(*a, vflag=0)
    rows = []
    for k in range(1, steps):
        barrier()
        t0 = time.perf_counter()
        rig.initial_integrate_device(*a, vflag=0)
        t_r = rig.last_ms()
        r, t_p, t_k = forces(k)
        rig.final_integrate_device(*a)
        t_r += rig.last_ms()
        barrier()
        rows.append((t_r, t_p, t_k, (time.perf_counter() - t0) * 1e3, r.iterations))
    m = np.array(rows[1:])[:, :4].mean(axis=0)    # the first step after the build still warms up (graphs, clocks)
    t = torch.tensor(m, dtype=torch.float64, device=dev)
    cnt = torch.tensor([nl, nl], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mn = cnt[:1].clone()
        dist.all_reduce(cnt[1:], op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        cnt[0] = mn[0]
    t = t.cpu().numpy()
    if rank == 0:
        print(json.dumps(dict(what="mgpu_md_timing", n_gpus=world, procgrid=pg, atoms=n, bodies=info.nbody, atoms_per_gpu=[int(cnt[0]), int(cnt[1])],
                              kspace=kspace + (f" grid {kinfo.nx}x{kinfo.ny}x{kinfo.nz} order {kinfo.order}" if kspace == "pppm" else f" kcount {kinfo.kcount}"),
                              steps_timed=len(rows) - 1, scf_iterations=[int(x[4]) for x in rows],
                              ms_per_step=dict(rigid=float(t[0]), pair=float(t[1]), kspace=float(t[2]), wall=float(t[3])),
                              atom_steps_per_s=n / (float(t[3]) * 1e-3))), flush=True)
    pair.close(), ks.close(), rig.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
