#!/usr/bin/env python3
"""A/B timing of kernel variants on the bench workload (GPU box)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402


def main():
    import torch
    pb = bench.load_pb()
    ncell = int(sys.argv[1]) if len(sys.argv) > 1 else bench.NCELL
    sysm = bench.workloads().lj_charge_fluid(ncell)
    n = sysm.n
    s = bench.make_style(pb, sysm, 0)
    x = np.ascontiguousarray(sysm.x); q = np.ascontiguousarray(sysm.q)
    ty = np.ascontiguousarray(sysm.type); al = np.ascontiguousarray(sysm.alpha)
    ref_mu = None
    import os
    spec = os.environ.get("VB_OPTS", "41:4:10:1,41:4:10:0")
    opts = [tuple(float(t) for t in o.split(":")) for o in spec.split(",")]
    for variant, tight, xb, alt, ev in opts:
        pb.lib().polb200_set_option(s._h, b"sweep_variant", float(variant))
        pb.lib().polb200_set_option(s._h, b"bin_div", float(tight))
        pb.lib().polb200_set_option(s._h, b"xsort_bits", float(xb))
        pb.lib().polb200_set_option(s._h, b"alternate", float(alt))
        pb.lib().polb200_set_option(s._h, b"l2_evict_first", float(ev))
        mu = np.zeros((n, 3)); f = np.zeros((n, 3))
        s.compute(x, q, ty, al, mu, f, ago=0)
        pb.lib().polb200_set_option(s._h, b"time_sweeps", 1.0)
        acc = np.zeros(5)
        reps = 5
        for k in range(reps):
            f[:] = 0
            r = s.compute(x, q, ty, al, mu, f, ago=1 + k)
            acc += [r.ms_neigh, r.ms_pair, r.ms_scf, r.ms_force, r.ms_total]
        sw = s.debug_fetch("sweep_timing", np.float64, 2)
        pb.lib().polb200_set_option(s._h, b"time_sweeps", 0.0)
        acc /= reps
        if ref_mu is None:
            ref_mu = mu.copy()
        err = np.abs(mu - ref_mu).max() / np.abs(ref_mu).max()
        if os.environ.get("VB_FLAGS"):
            print("   debug flags:", s.debug_fetch("flags", np.int32, 8))
        print(f"variant {variant:.0f} bin_div {tight} xsort_bits {xb:.0f} alternate {alt:.0f} evict_first {ev:.0f}: sweep {sw[0] / sw[1] * 1e3:8.1f} us  neigh {acc[0]:.3f} pair {acc[1]:.3f} "
              f"scf {acc[2]:.3f} force {acc[3]:.3f} total {acc[4]:.3f} ms   mu dev vs v0 {err:.2e}  E_pol {r.eng_pol:.10f}")
    s.close()


if __name__ == "__main__":
    main()
