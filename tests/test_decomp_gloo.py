"""CPU test (gloo, world_size 2) of the host-side logic of the multi-GPU path: the
decomposition plan exported by the C ABI (polb200_decomp_plan: who receives which boundary shell, with
which periodic image shift) drives a two-process emulation of the halo -- ghost positions once, ghost
dipoles once per SCF sweep -- whose converged-for-K-sweeps dipoles must equal the oracle's on the whole
periodic system.  No GPU and no compute call of the product: the per-pair arithmetic here is plain numpy.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
for _p in (str(ROOT), str(ROOT / "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

CUT, SKIN, SWEEPS, DAMP, GAMMA = 6.0, 2.0, 6, 2.1304, 1.03


def _sweep_field(xi, xall, mu_all, cut, self_index):
    """-sum_j T_ij mu_j over ext atoms within cut (exponential damping), dense numpy."""
    d = xi[:, None, :] - xall[None, :, :]
    r2 = (d * d).sum(-1)
    mask = r2 < cut * cut
    mask[np.arange(xi.shape[0]), self_index] = False
    r2 = np.where(mask, r2, 1.0)
    r = np.sqrt(r2)
    ar = DAMP * r
    e = np.exp(-ar)
    d1 = 1.0 - e * (1.0 + ar + 0.5 * ar * ar)
    d2 = d1 - e * ar ** 3 / 6.0
    s1 = np.where(mask, d1 / r ** 3, 0.0)
    s2 = np.where(mask, -3.0 * d2 / r ** 5, 0.0)
    dm = (d * mu_all[None, :, :]).sum(-1)
    return -(s1 @ mu_all + ((s2 * dm)[:, :, None] * d).sum(1))


def _static_field(xi, qall, xall, cut, self_index, kq):
    d = xi[:, None, :] - xall[None, :, :]
    r2 = (d * d).sum(-1)
    mask = r2 <= cut * cut
    mask[np.arange(xi.shape[0]), self_index] = False
    r2 = np.where(mask, r2, 1.0)
    sc = np.where(mask, (1.0 / r2 - 1.0 / (cut * cut)) / np.sqrt(r2), 0.0) * qall[None, :]
    return kq * (sc[:, :, None] * d).sum(1)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    import polhelpers as H
    from gpu_common import pb

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pg = (2, 1, 1)
    sysm = H.lj_charge_fluid((8, 4, 4), seed=99)
    prd = sysm.boxhi - sysm.boxlo
    plan = pb.decomp_plan(world, rank, pg, (1, 1, 1), sysm.boxlo, sysm.boxhi)
    lo, hi = plan["sublo"], plan["subhi"]
    own = np.nonzero(np.all((sysm.x >= lo) & (sysm.x < hi), axis=1))[0]
    x, q, alpha = sysm.x[own], sysm.q[own], sysm.alpha[own]
    cutghost = CUT + SKIN

    def send_lists():
        out = {}
        for d in range(27):
            if d == 13 or plan["dest"][d] < 0:
                continue
            v = (d % 3 - 1, (d // 3) % 3 - 1, d // 9 - 1)
            m = np.ones(len(own), dtype=bool)
            for k in range(3):
                if v[k] < 0:
                    m &= x[:, k] <= lo[k] + cutghost
                elif v[k] > 0:
                    m &= x[:, k] >= hi[k] - cutghost
            out[d] = np.nonzero(m)[0]
        return out

    sl = send_lists()

    def exchange(payload_of):
        """payload_of(d, idx) -> array; returns the concatenation of what arrives, direction-major."""
        mine = {d: payload_of(d, idx) for d, idx in sl.items()}
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        got = []
        for d in range(27):
            s = plan["src"][d]
            if d == 13 or s < 0:
                continue
            got.append(everyone[s][d])
        return np.concatenate(got)

    gx = exchange(lambda d, idx: x[idx] + plan["wrap"][d] * prd)
    gq = exchange(lambda d, idx: q[idx])
    xall = np.concatenate([x, gx])
    qall = np.concatenate([q, gq])
    self_index = np.arange(len(own))
    kq = np.sqrt(332.06371)
    ef = _static_field(x, qall, xall, CUT, self_index, kq)
    mu = GAMMA * alpha[:, None] * ef
    for _ in range(SWEEPS):
        gmu = exchange(lambda d, idx: mu[idx])
        mu = alpha[:, None] * (ef + _sweep_field(x, xall, np.concatenate([mu, gmu]), CUT, self_index))
    np.savez(Path(out_dir) / f"rank{rank}.npz", own=own, mu=mu, ef=ef, nghost=len(gx))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_halo_matches_oracle(tmp_path):
    import torch.multiprocessing as mp

    import polhelpers as H
    from oracle import polref as P

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sysm = H.lj_charge_fluid((8, 4, 4), seed=99)
    st = H.fluid_style(sysm, 2.5, CUT, polar_cut=CUT, fixed_iteration=1, max_iterations=SWEEPS, damp_type="exponential",
                       polar_gs_ranked=0)
    ref = P.polar_rows(sysm, st)
    seen = np.zeros(sysm.n, dtype=int)
    for r in range(2):
        z = np.load(tmp_path / f"rank{r}.npz")
        own = z["own"]
        seen[own] += 1
        assert z["nghost"] > 0
        assert H.rel_err(z["ef"], ref["ef_static"][own]) < 1e-11
        assert H.rel_err(z["mu"], ref["mu"][own]) < 1e-11
    assert np.all(seen == 1)  # the bricks partition the atoms


def test_plan_is_consistent_on_every_grid():
    from gpu_common import pb
    for world, pg in [(1, (1, 1, 1)), (2, (2, 1, 1)), (4, (2, 2, 1)), (8, (2, 2, 2)), (6, (3, 2, 1)), (12, (3, 2, 2))]:
        plans = [pb.decomp_plan(world, r, pg, (1, 1, 1), (0, 0, 0), (30, 20, 10)) for r in range(world)]
        for r, p in enumerate(plans):
            for d in range(27):
                if d == 13:
                    assert p["dest"][d] == -1 and p["src"][d] == -1
                    continue
                t = p["dest"][d]
                assert 0 <= t < world
                assert plans[t]["src"][d] == r  # what r sends towards d arrives at t as direction d
                # shift applied by the sender puts the image next to the receiver's brick
                lo_t, hi_t = plans[t]["sublo"], plans[t]["subhi"]
                v = np.array([d % 3 - 1, (d // 3) % 3 - 1, d // 9 - 1])
                face = np.where(v > 0, p["subhi"], np.where(v < 0, p["sublo"], 0.5 * (p["sublo"] + p["subhi"])))
                img = face + p["wrap"][d] * np.array([30.0, 20.0, 10.0])
                assert np.all(img >= lo_t - 1e-9) and np.all(img <= hi_t + 1e-9)
    # non-periodic dimension: no neighbour across the open faces
    p = pb.decomp_plan(2, 0, (2, 1, 1), (0, 1, 1), (0, 0, 0), (30, 20, 10))
    assert p["dest"][12] == -1 and p["dest"][14] == 1 and p["src"][14] == -1 and p["src"][12] == 1
    with pytest.raises(pb.Polb200Error):
        pb.decomp_plan(4, 0, (3, 1, 1), (1, 1, 1), (0, 0, 0), (1, 1, 1))
