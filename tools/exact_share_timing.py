#!/usr/bin/env python3
"""All-pairs (exact) mode shared by rows over N GPUs (polb200_comm_init_replicated): ms per step against one GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/exact_share_timing.py [ncell=17] [iterations=10]
Workload: the bench fluid without a dipole cutoff (every minimum-image pair, the reference's semantics), Jacobi sweeps,
fixed_iteration.  Every rank passes the whole system (host buffers) and gets the whole result; rank 0 prints one line."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
import bench

pb = bench.load_pb()


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    ncell = int(args[0]) if args else 17
    iters = int(args[1]) if len(args) > 1 else 10
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    s = bench.workloads().lj_charge_fluid(ncell)
    ew = pb.Ewald(device=local)
    g = ew.init(1e-4, s.q, 12.0, s.boxlo, s.boxhi).g_ewald
    ew.close()
    out = {}
    for shared in ([False, True] if world > 1 else [False]):
        p = pb.PairStyle(device=local)
        p.set_ntypes(2)
        p.command(f"pair_style lj/cut/coul/long/polarization 2.5 12.0 polar_gs_ranked no fixed_iteration yes max_iterations {iters} damp_type exponential")
        p.command("pair_coeff 1 1 0.1 3.0")
        p.command("pair_coeff 2 2 0.1 3.0")
        p.init(g_ewald=g, molecular=0)
        p.set_box(s.boxlo, s.boxhi)
        if shared:
            box = [pb.comm_create_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            p.comm_init_replicated(rank, world, box[0])
        x, q, t, al = (np.ascontiguousarray(a) for a in (s.x, s.q, s.type, s.alpha))
        tag = np.arange(1, s.n + 1, dtype=np.int32)
        mu = np.zeros((s.n, 3))
        ms = []
        for k in range(5):
            f = np.zeros((s.n, 3))
            r = p.compute(x, q, t, al, mu, f, tag=tag, eflag=1, vflag=2, ago=0 if k == 0 else k)
            ms.append((r.ms_total, r.ms_pair, r.ms_scf, r.ms_force))
        out["shared" if shared else "one_gpu"] = dict(zip(("ms_total", "ms_pair_field", "ms_scf", "ms_force"), np.array(ms[2:]).mean(0).round(3).tolist()),
                                                      eng_pol=r.eng_pol)
        p.close()
    if rank == 0:
        print(json.dumps(dict(what="exact_share_timing", n_gpus=world, atoms=s.n, jacobi_sweeps=iters, **out)), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
