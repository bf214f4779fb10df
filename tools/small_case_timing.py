#!/usr/bin/env python3
"""Stage times (CUDA events inside the library) of the reference's Bulk H2 example (750 atoms, exact all-pairs mode, the
reference's default solver) over repeated steps, host buffers: where the small-system step goes."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import polhelpers as H
from gpu_common import pb, configure_from_fixture, run_fixture

fx = H.load_fixture(sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "h2_default_step0")
for opts in (({},) if "--one" in sys.argv else ({}, {"gs_cluster": 8}, {"gs_cluster": 0}, {"gs_cluster": 0, "use_graphs": 0}, {"scf_lag": 0})):
    s = pb.PairStyle(device=0)
    configure_from_fixture(s, fx)
    for k, v in opts.items():
        s.set_option(k, v)
    mu = None
    rows = []
    import time
    for step in range(12):
        t0 = time.perf_counter()
        res, mu, ef, f = run_fixture(s, fx, ago=0 if step % 10 == 0 else 1, mu_in=mu)
        wall = (time.perf_counter() - t0) * 1e3
        rows.append((wall, res.ms_total, res.ms_neigh, res.ms_pair, res.ms_scf, res.ms_force, res.iterations))
    r = np.array(rows[2:10])
    print(opts, "wall %.3f  device total %.3f  neigh %.3f  pair+field %.3f  scf %.3f  force %.3f  iterations %.1f  launches/step %.0f" %
          (*r.mean(0), s.launch_count() / 12))
    s.close()
