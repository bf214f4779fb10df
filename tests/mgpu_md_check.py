"""A complete MD step on more than one GPU (SURVEY §8e + §8f ranks 1-2): rigid-body integrator + polarization pair
style + reciprocal-space solver, every one of them decomposed, against the same trajectory on a single GPU.  Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_md_check.py

Workload: the rigid polarizable water box of BASELINE config 3 (workloads.water_box), `fix rigid/nve molecule` (and one
rigid/nvt case), pair_style lj/cut/coul/long/polarization with special bonds, kspace ewald / pppm.  Atoms are owned
atom by atom by the brick they sit in (a molecule's atoms may live on different GPUs, as under the reference's MPI
decomposition, where `fix rigid` keeps every body on every rank and all-reduces the body forces and torques,
fix_rigid.cpp:782-855):

   pair style   polb200_comm_init      bricks + halo exchange (ghost dipoles every SCF sweep)
   KSpace       polb200_*_comm_init    every rank spreads / sums its own charges, S(k) or the charge grid is all-reduced
   rigid        polb200_rigid_comm_init   bodies replicated, 6 * nbody force / torque sums all-reduced

Every rank also advances the WHOLE system on its GPU with plain single-GPU handles (the already parity-checked path);
after every step the owned positions and velocities, the summed energies and the rigid-body kinetic energy must agree.
Re-neighboring steps wrap the atoms and move them to their new owner (the test's stand-in for Comm::exchange); the run
must see atoms change owner.  The SCF is the Jacobi sweep (same arithmetic in both copies, so the trajectories agree to
rounding); started from the generator's lattice it runs into the 100-sweep cap on every step, in both copies alike.
Prints one line per case, exits non-zero on failure.  Used by tests/test_multi_gpu.py (gpu marker).
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
from gpu_common import c as _c, pb


def c(a, dt=np.float64):
    return _c(a, dt)

GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
CUT = 8.0
REBUILD = 4


def water_topology(n):
    """special lists of 3-site molecules O H H (ids 3m+1..3m+3): O: 1-2 = {H,H}; H: 1-2 = {O}, 1-3 = {other H}"""
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


class Md:
    """one copy of the MD loop: `idx` = the atoms this copy owns (all of them for the single-GPU copy)"""

    def __init__(self, s, v0, img0, mass, dt, kspace, thermostat, comm=None):
        self.s, self.mass, self.comm = s, mass, comm
        self.L = s.boxhi - s.boxlo
        dev = torch.cuda.current_device()
        self.nspecial, self.special = water_topology(s.n)
        self.ks = pb.PPPM(device=dev) if kspace == "pppm" else pb.Ewald(device=dev)
        self.pair = pb.PairStyle(device=dev)
        self.rig = pb.Rigid(device=dev)
        if comm is not None:
            rank, world, ids, pg = comm
            self.pair_comm = (rank, world, ids[0], pg)
            self.ks.comm_init(rank, world, ids[1])
            self.rig.comm_init(rank, world, ids[2])
        kinfo = self.ks.init(1e-4, s.q, CUT, s.boxlo, s.boxhi)   # the GLOBAL charges (qsqsum, g_ewald)
        p = self.pair
        p.set_ntypes(2)
        p.command(f"pair_style lj/cut/coul/long/polarization 2.5 {CUT} polar_gs_ranked no precision 1e-10 max_iterations 100 "
                  f"damp_type exponential use_previous yes polar_cutoff {CUT}")
        p.command("pair_coeff 1 1 0.155 3.166 8.0")
        p.command("pair_coeff 2 2 0.0 1.0")
        p.init(g_ewald=kinfo.g_ewald, special_lj=(1.0, 0.0, 0.0, 0.0), special_coul=(1.0, 0.0, 0.0, 0.0), molecular=1)
        p.set_box(s.boxlo, s.boxhi)
        if comm is not None:
            p.comm_init(*self.pair_comm)
            self.lo, self.hi = p.subdomain()
        self.x, self.v, self.img = s.x.copy(), v0.copy(), img0.copy()     # global bookkeeping (valid for owned atoms)
        self.mu = np.zeros((s.n, 3))
        self.f = np.zeros((s.n, 3))
        self.own(first=True)
        idx = self.idx
        temp = (300.0, 300.0, 100.0) if thermostat else None
        self.info = self.rig.init(c(s.tag[idx], np.int32), c(s.molecule[idx], np.int32), c(mass[idx]), self.img[idx],
                                  c(self.x[idx]), c(self.v[idx]), s.boxlo, s.boxhi, dt, temp=temp)

    def own(self, first=False):
        s = self.s
        if self.comm is None:
            self.idx = np.arange(s.n)
        else:
            self.idx = np.nonzero(np.all((self.x >= self.lo) & (self.x < self.hi), axis=1))[0]

    def wrap(self):
        s = self.s
        shift = np.floor((self.x - s.boxlo) / self.L)
        self.x -= shift * self.L
        self.img += shift.astype(np.int64)

    def forces(self, ago):
        s, idx = self.s, self.idx
        x = c(self.x[idx])
        f = np.zeros((len(idx), 3))
        mu = c(self.mu[idx])
        r = self.pair.compute(x, c(s.q[idx]), c(s.type[idx], np.int32), c(s.alpha[idx]), mu, f, molecule=c(s.molecule[idx], np.int32),
                              tag=c(s.tag[idx], np.int32), nspecial=c(self.nspecial[idx], np.int32), special=c(self.special[idx], np.int32),
                              eflag=1, vflag=0, ago=ago)
        elong, _ = self.ks.compute(x, c(s.q[idx]), f, eflag=1, vflag=0)
        self.mu[idx] = mu
        self.f[idx] = f
        return r, elong

    def rigid(self, which):
        s, idx = self.s, self.idx
        tag, x, v, f = c(s.tag[idx], np.int32), c(self.x[idx]), c(self.v[idx]), c(self.f[idx])
        if which == "setup":
            self.rig.setup(tag, x, v, f, vflag=0)
        elif which == "initial":
            self.rig.initial_integrate(tag, x, v, f, vflag=0)
        else:
            self.rig.final_integrate(tag, x, v, f)
        self.x[idx], self.v[idx] = x, v

    def pre_neighbor(self):
        idx = self.idx
        self.rig.pre_neighbor(c(self.s.tag[idx], np.int32), self.img[idx])

    def close(self):
        self.pair.close(), self.ks.close(), self.rig.close()


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pg = GRIDS[world]
    nside = int(os.environ.get("MGPU_MD_NSIDE", "13"))   # odd: the brick faces cut through a layer of molecules
    steps = int(os.environ.get("MGPU_MD_STEPS", "10"))
    dt = 1.0
    wl = H._workloads().water_box(nside)
    n = wl.n
    L = wl.boxhi - wl.boxlo
    first = wl.x[0::3].repeat(3, axis=0)
    img0 = -np.rint((wl.x - first) / L).astype(np.int64)     # the generator wrapped the molecules atom by atom
    mass = np.where(wl.type == 1, 15.9994, 1.008)
    v0 = np.random.default_rng(99).normal(size=(n, 3)) * np.sqrt(pb.REAL_BOLTZ * 298.15 / (mass[:, None] * pb.REAL_MVV2E))
    failures = []

    def fresh_ids(k):
        box = [[pb.comm_create_id() for _ in range(k)] if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def gather_owned(md):
        """every rank learns the current x, v, mu, image of all atoms from their owners (the stand-in for Comm::exchange)"""
        own = torch.zeros(n, dtype=torch.float64, device="cuda")
        own[torch.from_numpy(md.idx).cuda()] = 1.0
        packed = np.zeros((n, 12))
        packed[md.idx, 0:3], packed[md.idx, 3:6], packed[md.idx, 6:9] = md.x[md.idx], md.v[md.idx], md.mu[md.idx]
        packed[md.idx, 9:12] = md.img[md.idx]
        t = torch.from_numpy(packed).cuda()
        dist.all_reduce(t)
        dist.all_reduce(own)
        assert bool((own == 1.0).all()), "every atom must have exactly one owner"
        a = t.cpu().numpy()
        md.x, md.v, md.mu, md.img = a[:, 0:3].copy(), a[:, 3:6].copy(), a[:, 6:9].copy(), np.rint(a[:, 9:12]).astype(np.int64)

    for name, kspace, thermostat in (("rigid_nve_ewald", "ewald", False), ("rigid_nvt_pppm", "pppm", True)):
        ref = Md(wl, v0, img0, mass, dt, kspace, thermostat)
        dec = Md(wl, v0, img0, mass, dt, kspace, thermostat, comm=(rank, world, fresh_ids(3), pg))
        assert dec.info.nbody == ref.info.nbody == n // 3
        worst = dict(x=0.0, v=0.0, e=0.0, ke=0.0)
        moved = 0
        iters = []
        for k in range(steps + 1):
            ago = k % REBUILD
            for md in (ref, dec):
                if k > 0:
                    md.rigid("initial")
                if ago == 0 and k > 0:
                    if md is dec:
                        gather_owned(md)
                    md.wrap()
                    before = set(md.idx.tolist())
                    md.own()
                    if md is dec:
                        moved += len(set(md.idx.tolist()) - before)
                        if os.environ.get("MGPU_MD_DEBUG"):
                            print(f"rank {rank} step {k}: owned {len(before)} -> {len(md.idx)}, arrivals {len(set(md.idx.tolist()) - before)}, "
                                  f"lo {md.lo} hi {md.hi}", flush=True)
                    md.pre_neighbor()
            r0, el0 = ref.forces(ago)
            r1, el1 = dec.forces(ago)
            for md in (ref, dec):
                md.rigid("setup" if k == 0 else "final")
            e1 = torch.tensor([r1.eng_vdwl, r1.eng_coul, r1.eng_pol, el1], dtype=torch.float64, device="cuda")
            dist.all_reduce(e1)
            e1 = e1.cpu().numpy()
            e0 = np.array([r0.eng_vdwl, r0.eng_coul, r0.eng_pol, el0])
            idx = dec.idx
            ke0, ke1 = ref.rig.scalars(), dec.rig.scalars()
            worst["x"] = max(worst["x"], float(np.abs(dec.x[idx] - ref.x[idx]).max()))
            worst["v"] = max(worst["v"], float(np.abs(dec.v[idx] - ref.v[idx]).max() / np.abs(ref.v).max()))
            worst["e"] = max(worst["e"], float(np.abs(e1 - e0).max() / np.abs(e0).max()))
            worst["ke"] = max(worst["ke"], abs(ke1[0] - ke0[0]) / abs(ke0[0]))
            iters.append((r1.iterations, r0.iterations))
        un0 = ref.x + ref.img * L
        disp = float(np.abs(un0 - (wl.x + img0 * L)).max())
        ok = worst["x"] < 1e-8 and worst["v"] < 1e-8 and worst["e"] < 1e-9 and worst["ke"] < 1e-9
        ok = ok and all(a == b for a, b in iters)
        t = torch.tensor([0 if ok else 1, moved], device="cuda")
        dist.all_reduce(t)
        if int(t[1]) == 0:   # the run must exercise migration between bricks
            t[0] += 1
        if rank == 0:
            print(f"[mgpu md {world} ranks grid {pg}] {name}: {'OK' if int(t[0]) == 0 else 'FAIL'} | {n} atoms {n // 3} bodies, {steps} steps, "
                  f"rebuild every {REBUILD}, {int(t[1])} atoms changed owner | max |dx| {worst['x']:.1e} A, dv {worst['v']:.1e}, "
                  f"energies {worst['e']:.1e}, rigid KE {worst['ke']:.1e} | SCF iterations {iters[0]}..{iters[-1]} | "
                  f"E_pol {e0[2]:.3f} E_coul {e0[1]:.3f} E_long {e0[3]:.3f} KE {ke0[0]:.4f}, largest move {disp:.3f} A", flush=True)
        if int(t[0]):
            failures.append(name)
        ref.close(), dec.close()
    dist.barrier()
    dist.destroy_process_group()
    if failures:
        if rank == 0:
            print("FAILED: " + ", ".join(failures), flush=True)
        sys.exit(1)
    if rank == 0:
        print("all md cases passed", flush=True)


if __name__ == "__main__":
    main()
