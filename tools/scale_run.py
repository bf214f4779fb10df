#!/usr/bin/env python3
"""Large-system runs of the decomposed path (BASELINE configs 3-5 shapes) with an INDEPENDENT full-size check.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/scale_run.py --ncell 32 --steps 3 --mode precision

One periodic LJ+charge fluid of 4*ncell^3 atoms PER GPU (ncell 32 -> 131 072 per GPU, 1.05 M on 8 GPUs;
ncell 63 -> 1.0 M per GPU, 8.0 M on 8), cut 2.5/12, polar_cutoff 12, SCF either converged
(`--mode precision`: Jacobi to 1e-8; `--mode ranked`: default polar_gs_ranked, precision 1e-11, polar_gamma 1.03)
or fixed (`--mode fixed`: 30 Jacobi sweeps).

Check at full size without the O(N^2) reference: rank 0 gathers positions, dipoles, static fields and forces of all
bricks and, for a sample of atoms, recomputes in numpy (minimum image over ALL atoms) the static field and the
SCF residual  mu_i - alpha_i (E_i - sum_j T_ij mu_j); plus Newton's third law (total pair force = 0).
Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist

import bench  # product-side helpers only: C-ABI binding, workloads (no oracle)

pb = bench.load_pb()

GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
CUT = 12.0
DAMP = 2.1304
KQ = np.sqrt(332.06371)


def sample_check(x, q, alpha, mu, ef, boxlen, samples, rng, mol=None):
    """numpy recomputation for `samples` random atoms against all atoms (minimum image)."""
    n = x.shape[0]
    idx = rng.choice(n, samples, replace=False)
    worst_ef = worst_res = 0.0
    for i in idx:
        d = x[i] - x
        d -= boxlen * np.round(d / boxlen)
        r2 = (d * d).sum(1)
        m = (r2 < CUT * CUT) & (r2 > 0)
        dd, rr2 = d[m], r2[m]
        r = np.sqrt(rr2)
        # static field: k * sum q_j (1/r^2 - 1/rc^2)/r * del   (r <= rc; inter-molecular pairs only)
        w = q[m] if mol is None or mol[i] == 0 else np.where(mol[m] != mol[i], q[m], 0.0)
        e_s = KQ * (((1.0 / rr2 - 1.0 / (CUT * CUT)) / r * w)[:, None] * dd).sum(0)
        ar = DAMP * r
        ex = np.exp(-ar)
        d1 = 1.0 - ex * (1.0 + ar + 0.5 * ar * ar)
        d2 = d1 - ex * ar ** 3 / 6.0
        s1, s2 = d1 / r ** 3, -3.0 * d2 / r ** 5
        mj = mu[m]
        e_ind = -((s1[:, None] * mj).sum(0) + ((s2 * (dd * mj).sum(1))[:, None] * dd).sum(0))
        worst_ef = max(worst_ef, np.abs(e_s - ef[i]).max() / max(np.abs(ef[i]).max(), 1e-300))
        res = mu[i] - alpha[i] * (ef[i] + e_ind)
        worst_res = max(worst_res, np.abs(res).max())
    return worst_ef, worst_res


def mof_supercell(R):
    """BASELINE config 4 shape (workloads.mof_supercell)"""
    sysm, cut, coeff = bench.workloads().mof_supercell(ROOT / "tests" / "golden" / "co2_singlepoint_step0.npz", R)
    return dict(sys=sysm, cut=cut, pair_coeff=coeff)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncell", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--mode", default="precision", choices=["precision", "ranked", "fixed"])
    ap.add_argument("--samples", type=int, default=48)
    ap.add_argument("--global-ncell", type=int, default=0, help="strong scaling: 4*G^3 atoms in total, whatever the GPU count")
    ap.add_argument("--water", type=int, default=0, help="BASELINE config 3: rigid water box of 3*W^3 atoms instead of the fluid")
    ap.add_argument("--mof", type=int, default=0, help="BASELINE config 4: the MOF-5 + CO2 cell of the golden fixture (924 atoms, "
                    "bond topology, 10 atom types) replicated R^3 times")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pg = GRIDS[world]
    t0 = time.time()
    cells = (args.global_ncell,) * 3 if args.global_ncell else tuple(args.ncell * np.array(pg))
    global CUT
    mof = None
    if args.mof:
        mof = mof_supercell(args.mof)
        gsys = mof["sys"]
        CUT = mof["cut"]
    else:
        gsys = bench.workloads().water_box(args.water) if args.water else bench.workloads().lj_charge_fluid(cells, seed=4242)
    words = {"precision": "polar_gs_ranked no precision 1e-8 max_iterations 200 damp_type exponential",
             "ranked": "precision 1e-11 max_iterations 200 polar_gamma 1.03 damp_type exponential",
             "fixed": "polar_gs_ranked no fixed_iteration yes max_iterations 30 damp_type exponential"}[args.mode]
    ew = pb.Ewald(device=local)  # g_ewald as `kspace_style ewald 1e-4` sets it
    g = ew.init(1e-4, gsys.q, CUT, gsys.boxlo, gsys.boxhi).g_ewald
    ew.close()
    s = pb.PairStyle(device=local)
    s.set_ntypes(int(gsys.ntypes))
    s.command(f"pair_style lj/cut/coul/long/polarization 2.5 {CUT} {words} polar_cutoff {CUT}")
    if mof:
        for line in mof["pair_coeff"]:
            s.command(line)
    elif args.water:
        s.command("pair_coeff 1 1 0.155 3.166")
        s.command("pair_coeff 2 2 0.0 1.0")
    else:
        s.command("pair_coeff * * 0.1 3.0")
    s.init(g_ewald=g, molecular=1 if mof else 0)
    s.set_box(gsys.boxlo, gsys.boxhi)
    if world > 1:
        box = [pb.comm_create_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        s.comm_init(rank, world, box[0], pg)
        lo, hi = s.subdomain()
        own = np.nonzero(np.all((gsys.x >= lo) & (gsys.x < hi), axis=1))[0]
    else:
        own = np.arange(gsys.n)
    n = len(own)
    x = np.ascontiguousarray(gsys.x[own]); q = np.ascontiguousarray(gsys.q[own])
    ty = np.ascontiguousarray(gsys.type[own]); al = np.ascontiguousarray(gsys.alpha[own])
    tag = np.ascontiguousarray(gsys.tag[own]); molecule = np.ascontiguousarray(gsys.molecule[own])
    nspecial = np.ascontiguousarray(gsys.nspecial[own]) if gsys.nspecial is not None else None
    special = np.ascontiguousarray(gsys.special[own]) if gsys.special is not None else None
    mu = np.zeros((n, 3)); f = np.zeros((n, 3)); ef = np.zeros((n, 3))
    t_setup = time.time() - t0
    rows = []
    for k in range(args.steps):
        f[:] = 0.0
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t1 = time.perf_counter()
        r = s.compute(x, q, ty, al, mu, f, molecule=molecule, tag=tag, ef_static=ef, eflag=1, vflag=2, ago=k,
                      nspecial=nspecial, special=special)
        torch.cuda.synchronize()
        rows.append(dict(step=k, ms=(time.perf_counter() - t1) * 1e3, iterations=r.iterations, ms_neigh=r.ms_neigh,
                         ms_pair=r.ms_pair, ms_scf=r.ms_scf, ms_force=r.ms_force, diverged=bool(r.status & pb.STATUS_DIVERGED),
                         eng_pol=r.eng_pol, nghost=r.nghost, npairs_full=r.npairs_full))
    free_b, total_b = torch.cuda.mem_get_info()
    # gather everything on rank 0 for the independent check
    parts = dict(own=own, mu=mu, ef=ef, f=f)
    if world > 1:
        allp = [None] * world if rank == 0 else None
        dist.gather_object(parts, allp, dst=0)
        times = torch.tensor([rw["ms"] for rw in rows], dtype=torch.float64, device="cuda")
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        epol = torch.tensor([rows[-1]["eng_pol"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(epol)
        times, epol = times.cpu().numpy(), float(epol)
    else:
        allp, times, epol = [parts], np.array([rw["ms"] for rw in rows]), rows[-1]["eng_pol"]
    if rank == 0:
        N = gsys.n
        MU = np.zeros((N, 3)); EF = np.zeros((N, 3)); F = np.zeros((N, 3))
        seen = np.zeros(N, dtype=np.int64)
        for p_ in allp:
            MU[p_["own"]] = p_["mu"]; EF[p_["own"]] = p_["ef"]; F[p_["own"]] = p_["f"]
            seen[p_["own"]] += 1
        rng = np.random.default_rng(11)
        tc = time.time()
        worst_ef, worst_res = sample_check(gsys.x, gsys.q, gsys.alpha, MU, EF, gsys.boxhi - gsys.boxlo, args.samples, rng,
                                           gsys.molecule if (args.water or mof) else None)
        fsum = np.abs(F.sum(0)).max() / (np.abs(F).max() * np.sqrt(N))
        line = dict(what="scale_run", system="mof5+co2" if mof else ("water" if args.water else "fluid"), n_gpus=world, grid=pg, atoms_total=N, atoms_per_gpu=N // world, mode=args.mode,
                    pair_style=words + f" polar_cutoff {CUT}", steps=rows, ms_per_step_max_over_ranks=list(map(float, times)),
                    atom_steps_per_s_last=N / (times[-1] * 1e-3), eng_pol_total=epol, every_atom_owned_once=bool(np.all(seen == 1)),
                    check=dict(samples=args.samples, static_field_rel_err=worst_ef, scf_residual_abs=worst_res,
                               mu_scale=float(np.abs(MU).max()), total_force_over_sqrtN_fmax=float(fsum), seconds=time.time() - tc),
                    hbm_used_gb_rank0=(total_b - free_b) / 2 ** 30, setup_s=t_setup)
        print(json.dumps(line))
    s.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
