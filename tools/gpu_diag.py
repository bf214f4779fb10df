#!/usr/bin/env python3
"""Stage-by-stage diagnostics on a GPU box: prints errors instead of asserting (first-light tool)."""
import sys
import time
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import polhelpers as H  # noqa: E402
from gpu_common import configure_from_fixture, pb, run_fixture, run_system  # noqa: E402
from oracle import polref as P  # noqa: E402


def report(tag, res, mu, ef, f, fx):
    fs = np.abs(fx["f"]).max()
    print(f"[{tag}] iters dev/ref {res.iterations}/{int(fx['iterations'])} status {res.status} "
          f"nghost {res.nghost} npairs {res.npairs_full} (ref half*2 {2 * int(fx['npairs_half'])})")
    print(f"    ef rel {H.rel_err(ef, fx['ef_static']):.3e}  mu rel {H.rel_err(mu, fx['mu_out']):.3e} "
          f"mu abs {np.abs(mu - fx['mu_out']).max():.3e}  f rel {np.abs(f - fx['f']).max() / fs:.3e}")
    for k in ("eng_vdwl", "eng_coul", "eng_pol"):
        r = float(fx[k])
        print(f"    {k} {getattr(res, k):.12g} ref {r:.12g} rel {abs(getattr(res, k) - r) / max(abs(r), 1e-300):.3e}")
    print(f"    virial rel {H.rel_err(np.array(res.virial[:]), fx['virial']):.3e}")
    print(f"    ms neigh {res.ms_neigh:.3f} pair {res.ms_pair:.3f} scf {res.ms_scf:.3f} force {res.ms_force:.3f} "
          f"total {res.ms_total:.3f}")


def main():
    cases = sys.argv[1:] or [p.stem for p in sorted(H.GOLDEN.glob("*_step0.npz"))]
    for case in cases:
        try:
            fx = H.load_fixture(case)
            s = pb.PairStyle(device=0)
            configure_from_fixture(s, fx)
            res, mu, ef, f = run_fixture(s, fx)
            report(case, res, mu, ef, f, fx)
            s.close()
        except Exception:
            print(f"[{case}] FAILED")
            traceback.print_exc()
    # list mode, synthetic
    try:
        for ncell, iters in ((10, 5), (20, 30)):
            sysm = H.lj_charge_fluid(ncell)
            st = H.fluid_style(sysm, 2.5, 12.0, polar_cut=12.0, fixed_iteration=1, max_iterations=iters,
                               damp_type="exponential", polar_gs_ranked=0)
            s = pb.PairStyle(device=0)
            s.set_ntypes(2)
            s.command("pair_style lj/cut/coul/long/polarization 2.5 12.0 polar_gs_ranked no fixed_iteration yes "
                      f"max_iterations {iters} damp_type exponential polar_cutoff 12.0")
            s.command("pair_coeff * * 0.1 3.0")
            s.init(g_ewald=st.g_ewald, molecular=0)
            s.set_box(sysm.boxlo, sysm.boxhi)
            res, mu, ef, f = run_system(s, sysm)
            t0 = time.time()
            for _ in range(3):
                res, mu, ef, f = run_system(s, sysm, ago=1)
            dt = (time.time() - t0) / 3
            print(f"[fluid N={sysm.n}] iters {res.iterations} nghost {res.nghost} npairs {res.npairs_full} "
                  f"ms neigh {res.ms_neigh:.3f} pair {res.ms_pair:.3f} scf {res.ms_scf:.3f} "
                  f"force {res.ms_force:.3f} total {res.ms_total:.3f} wall {dt * 1e3:.2f}")
            if ncell == 10:
                ref = P.polar_rows(sysm, st)
                print(f"    ef rel {H.rel_err(ef, ref['ef_static']):.3e} mu rel {H.rel_err(mu, ref['mu']):.3e} "
                      f"epol {res.eng_pol:.12g} ref {ref['eng_pol']:.12g}")
            s.close()
    except Exception:
        print("[fluid] FAILED")
        traceback.print_exc()


if __name__ == "__main__":
    main()
