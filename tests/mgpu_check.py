"""Multi-GPU parity check of the spatial decomposition.  Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_check.py
Every rank computes the WHOLE system on its own GPU with a plain single-GPU handle (the already
parity-checked path) and its brick of the same system through the decomposed path; owned dipoles,
fields and forces must agree, energies and virial after summing over ranks.

Cases: Jacobi fixed-iteration (NCCL halo and fused peer push), Jacobi precision mode (iteration counts
equal), ranked colouring sweep (tolerance), a step without rebuild after moving atoms, and a rebuild; exclusion rules and
atom_slack across bricks; decomposed Ewald / PPPM; the all-pairs (exact) mode shared by rows (bit-identical to one GPU);
collective error agreement.
Prints one line per case and exits non-zero on failure.  Used by tests/test_multi_gpu.py (gpu marker).
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist

import polhelpers as H
from gpu_common import c, pb
from oracle import polref as P  # only ewald_g (host setup arithmetic)

GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def make_style(device, sysm, words, cut_coul):
    g = P.ewald_g(1e-4, sysm.q, cut_coul, sysm.boxlo, sysm.boxhi)
    s = pb.PairStyle(device=device)
    s.set_ntypes(2)
    s.command(f"pair_style lj/cut/coul/long/polarization 2.5 {cut_coul} {words} polar_cutoff {cut_coul}")
    s.command("pair_coeff 1 1 0.1 3.0")
    s.command("pair_coeff 2 2 0.1 3.0")
    s.init(g_ewald=g, molecular=0)
    s.set_box(sysm.boxlo, sysm.boxhi)
    return s


def run(style, x, q, typ, alpha, tag, mu, ago, mask=None):
    n = x.shape[0]
    f = np.zeros((n, 3))
    ef = np.zeros((n, 3))
    mu = mu.copy()
    res = style.compute(c(x, np.float64), c(q, np.float64), c(typ, np.int32), c(alpha, np.float64), mu, f,
                        tag=c(tag, np.int32), ef_static=ef, eflag=1, vflag=2, ago=ago,
                        mask=None if mask is None else c(mask, np.int32))
    return res, mu, ef, f


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pg = GRIDS[world]
    ncell = int(os.environ.get("MGPU_NCELL", "8"))
    cut = float(os.environ.get("MGPU_CUT", "8.0"))
    sysm = H.lj_charge_fluid(tuple(ncell * np.array(pg)), seed=777)
    failures = []

    def fresh_id():
        # one NCCL id per communicator; it travels through torch.distributed (the caller's own transport,
        # MPI_Bcast in LAMMPS)
        box = [pb.comm_create_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def owned_mask(x, lo, hi):
        return np.all((x >= lo) & (x < hi), axis=1)

    cases = [("jacobi_fixed_nccl", "polar_gs_ranked no fixed_iteration yes max_iterations 12 damp_type exponential", 0, 1e-12),
             ("jacobi_fixed_push", "polar_gs_ranked no fixed_iteration yes max_iterations 12 damp_type exponential", 1, 1e-12),
             ("jacobi_precision_push", "polar_gs_ranked no precision 1e-9 max_iterations 60 damp_type exponential use_previous yes", 1, 1e-12),
             ("jacobi_precision_nccl", "polar_gs_ranked no precision 1e-9 max_iterations 60 damp_type exponential", 0, 1e-12),
             ("gs_ranked_chunks", "precision 1e-11 max_iterations 60 damp_type exponential", 1, 5e-9),
             # neigh_modify exclude / include rules across bricks: the ghosts carry their owners' membership bits
             ("exclusions_push", "polar_gs_ranked no fixed_iteration yes max_iterations 6 damp_type exponential", 1, 1e-12),
             # option atom_slack: a caller that keeps whole molecules together hands a brick atoms up to `slack` outside it
             ("atom_slack_push", "polar_gs_ranked no fixed_iteration yes max_iterations 6 damp_type exponential", 1, 1e-12)]
    off = np.random.default_rng(6).uniform(-0.8, 0.8, size=(sysm.n, 3))   # |off| <= 1.39 < the 1.5 A slack below
    gmask = (1 | (2 * (np.random.default_rng(3).random(sysm.n) < 0.4)) | (4 * (np.random.default_rng(4).random(sysm.n) < 0.6))).astype(np.int32)
    for name, words, push, tol in cases:
        ref = make_style(local, sysm, words, cut)
        dec = make_style(local, sysm, words, cut)
        dec.comm_init(rank, world, fresh_id(), pg)
        dec.set_option("p2p_push", push)
        mask = None
        if name.startswith("exclusions"):
            rules = [("type", 1, 1), ("group", 2, 4), ("include", 4)]
            ref.set_exclusions(rules)
            dec.set_exclusions(rules)
            mask = gmask
        lo, hi = dec.subdomain()
        slack = name.startswith("atom_slack")
        if slack:
            dec.set_option("atom_slack", 1.5)
        x = sysm.x.copy()
        mu_g = np.zeros((sysm.n, 3))
        rng = np.random.default_rng(5)
        ok = True
        msgs = []
        for step, ago in enumerate([0, 1, 2, 0]):
            if step > 0 and ago == 0:
                # rebuild: atoms may have changed bricks (wrap first, like Domain::pbc + Comm::exchange)
                x = sysm.boxlo + np.mod(x - sysm.boxlo, sysm.boxhi - sysm.boxlo)
            if ago == 0:
                where = sysm.boxlo + np.mod(x + off - sysm.boxlo, sysm.boxhi - sysm.boxlo) if slack else x
                mine = owned_mask(where, lo, hi)
                idx = np.nonzero(mine)[0]
                # the periodic image next to the brick (an atom at boxhi - 0.3 that belongs to a molecule of the brick at boxlo)
                image_shift = (where - off - x) if slack else np.zeros_like(x)
            r0, mu0, ef0, f0 = run(ref, x, sysm.q, sysm.type, sysm.alpha, sysm.tag, mu_g, ago, mask)
            r1, mu1, ef1, f1 = run(dec, (x + image_shift)[idx], sysm.q[idx], sysm.type[idx], sysm.alpha[idx], sysm.tag[idx], mu_g[idx], ago,
                                   None if mask is None else mask[idx])
            stats = dec.debug_fetch("comm_stats", np.float64, 5)
            # the polarization virial is F.r at the STORED coordinates of the owned atoms (the reference's fdotr form,
            # SURVEY H7), so it follows the periodic image an atom is stored at: left out where the images differ on purpose
            nv = 0 if slack else 6
            e = torch.tensor([r1.eng_vdwl, r1.eng_coul, r1.eng_pol] + list(r1.virial[:nv]), dtype=torch.float64, device="cuda")
            dist.all_reduce(e)
            e = e.cpu().numpy()
            e0 = np.array([r0.eng_vdwl, r0.eng_coul, r0.eng_pol] + list(r0.virial[:nv]))
            errs = dict(mu=H.rel_err(mu1, mu0[idx]), ef=H.rel_err(ef1, ef0[idx]),
                        f=float(np.abs(f1 - f0[idx]).max() / np.abs(f0).max()),
                        e=float(np.abs(e - e0).max() / np.abs(e0).max()))
            # the colouring chunks are cut from each brick's own ranked order: same fixed point, iteration counts may differ
            it_ok = (r1.iterations == r0.iterations if "gs_ranked" not in name
                     else abs(r1.iterations - r0.iterations) <= max(3, r0.iterations // 3))
            good = all(v < tol for v in errs.values()) and it_ok and int(stats[2]) == sysm.n
            if push and "gs" not in name:
                good = good and int(stats[3]) == 1
            ok = ok and good
            msgs.append(f"step{step}(ago={ago}) it {r1.iterations}/{r0.iterations} " +
                        " ".join(f"{k}={v:.1e}" for k, v in errs.items()) + f" push={int(stats[3])}")
            # move atoms a little (within the half-skin) for the next step; persistent dipoles follow
            x = x + rng.uniform(-0.15, 0.15, size=x.shape)
            mu_g = mu0
        flag = torch.tensor([0 if ok else 1], device="cuda")
        dist.all_reduce(flag)
        if rank == 0:
            print(f"[mgpu {world} ranks grid {pg}] {name}: {'OK' if int(flag) == 0 else 'FAIL'} | " + " | ".join(msgs), flush=True)
        if int(flag):
            failures.append(name)
        ref.close()
        dec.close()
    # KSpace on the bricks: every rank computes with ITS atoms (structure factors / charge grid all-reduced), forces of the
    # owned atoms and the summed energy / virial against the single-GPU compute of the whole system
    lo, hi = None, None
    probe = make_style(local, sysm, cases[0][1], cut)
    probe.comm_init(rank, world, fresh_id(), pg)
    lo, hi = probe.subdomain()
    probe.close()
    idx = np.nonzero(owned_mask(sysm.x, lo, hi))[0]
    for name, cls, acc in (("kspace_ewald", pb.Ewald, 1e-5), ("kspace_pppm", pb.PPPM, 1e-5)):
        one = cls(device=local)
        one.init(acc, sysm.q, cut, sysm.boxlo, sysm.boxhi)
        f0 = np.zeros((sysm.n, 3))
        e0, v0 = one.compute(c(sysm.x, np.float64), c(sysm.q, np.float64), f0)
        one.close()
        dec = cls(device=local)
        dec.comm_init(rank, world, fresh_id())
        dec.init(acc, sysm.q, cut, sysm.boxlo, sysm.boxhi)          # global charges: qsum, qsqsum, natoms of the whole system
        f1 = np.zeros((len(idx), 3))
        e1, v1 = dec.compute(c(sysm.x[idx], np.float64), c(sysm.q[idx], np.float64), f1)
        dec.close()
        ev = torch.tensor([e1] + list(v1), dtype=torch.float64, device="cuda")
        dist.all_reduce(ev)
        ev = ev.cpu().numpy()
        errs = dict(f=float(np.abs(f1 - f0[idx]).max() / np.abs(f0).max()), e=abs(ev[0] - e0) / abs(e0),
                    v=float(np.abs(ev[1:] - v0).max() / np.abs(v0).max()))
        good = all(v < 1e-11 for v in errs.values())
        flag = torch.tensor([0 if good else 1], device="cuda")
        dist.all_reduce(flag)
        if rank == 0:
            print(f"[mgpu {world} ranks] {name}: {'OK' if int(flag) == 0 else 'FAIL'} | " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()), flush=True)
        if int(flag):
            failures.append(name)

    # the all-pairs (exact) mode shared by rows (polb200_comm_init_replicated): every rank holds the whole system; the result
    # must equal the single-GPU one BIT FOR BIT (same per-block partial sums, same fixed-order reductions), on every rank
    small = H.lj_charge_fluid(6, seed=31)
    keep = np.arange(small.n) != 17          # 863 atoms: the last rank's chunk is short, the last row block partial
    xs, qs, ts, als, tags = small.x[keep], small.q[keep], small.type[keep], small.alpha[keep], np.arange(1, keep.sum() + 1)
    g_small = P.ewald_g(1e-4, qs, 9.0, small.boxlo, small.boxhi)
    full = (xs, qs, ts, als, tags)
    for name, words in (("exact_jacobi_fixed", "polar_gs_ranked no fixed_iteration yes max_iterations 9 damp_type exponential"),
                        ("exact_jacobi_precision", "polar_gs_ranked no precision 1e-10 max_iterations 80 damp_type exponential"),
                        ("exact_gs_ranked", "precision 1e-11 max_iterations 60 damp_type exponential"),
                        # five atoms: fewer row blocks than processes, the last processes own no row at all
                        ("exact_tiny", "polar_gs_ranked no precision 1e-10 max_iterations 80 damp_type exponential")):
        xs, qs, ts, als, tags = (a[:5] for a in full) if name == "exact_tiny" else full
        outs = []
        for shared in (False, True):
            s = pb.PairStyle(device=local)
            s.set_ntypes(2)
            s.command(f"pair_style lj/cut/coul/long/polarization 2.5 9.0 {words}")
            s.command("pair_coeff 1 1 0.1 3.0")
            s.command("pair_coeff 2 2 0.1 3.0")
            s.init(g_ewald=g_small, molecular=0)
            s.set_box(small.boxlo, small.boxhi)
            if shared:
                s.comm_init_replicated(rank, world, fresh_id())
            mu = np.zeros((len(xs), 3))
            res = None
            for step in range(2):            # second step: use of the previous dipoles, lists re-used
                res, mu, ef, f = run(s, xs + 0.01 * step, qs, ts, als, tags, mu, step)
            outs.append((res, mu, ef, f))
            s.close()
        (r0, mu0, ef0, f0), (r1, mu1, ef1, f1) = outs
        good = (r0.iterations == r1.iterations and np.array_equal(mu0, mu1) and np.array_equal(ef0, ef1) and np.array_equal(f0, f1)
                and r0.eng_pol == r1.eng_pol and r0.eng_coul == r1.eng_coul and list(r0.virial[:]) == list(r1.virial[:]))
        flag = torch.tensor([0 if good else 1], device="cuda")
        dist.all_reduce(flag)
        if rank == 0:
            print(f"[mgpu {world} ranks] {name}: {'OK' if int(flag) == 0 else 'FAIL'} | {len(xs)} atoms, iterations {r1.iterations}/{r0.iterations}, "
                  f"max |dmu| {np.abs(mu1 - mu0).max():.1e}, max |df| {np.abs(f1 - f0).max():.1e}, E_pol {r1.eng_pol:.10f} / {r0.eng_pol:.10f}", flush=True)
        if int(flag):
            failures.append(name)

    # a brick that fails alone must take the others with it (same error on every rank, no hang): rank 0 is handed no
    # atoms at a rebuild step
    dec = make_style(local, sysm, cases[0][1], cut)
    dec.comm_init(rank, world, fresh_id(), pg)
    lo, hi = dec.subdomain()
    idx = np.nonzero(owned_mask(sysm.x, lo, hi))[0]
    if rank == 0:
        idx = idx[:0]
    try:
        run(dec, sysm.x[idx], sysm.q[idx], sysm.type[idx], sysm.alpha[idx], sysm.tag[idx], np.zeros((sysm.n, 3))[idx], 0)
        msg = "no error"
    except pb.Polb200Error as e:
        msg = str(e)
    agreed = "owns no atoms" in msg
    flag = torch.tensor([0 if agreed else 1], device="cuda")
    dist.all_reduce(flag)
    if rank == 0:
        print(f"[mgpu {world} ranks] empty brick -> every rank: {msg!r}: {'OK' if int(flag) == 0 else 'FAIL'}", flush=True)
    if int(flag):
        failures.append("error_agreement")
    dec.close()
    dist.barrier()
    dist.destroy_process_group()
    if failures:
        print("FAILED:", failures)
        sys.exit(1)
    if rank == 0:
        print("mgpu_check: all cases passed")


if __name__ == "__main__":
    main()
