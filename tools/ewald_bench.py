#!/usr/bin/env python3
"""Device Ewald on the bench workload (32 000-atom fluid, accuracy 1e-4): CUDA-event time per compute with the
inputs resident in HBM, and the oracle on the host cores beside it."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import bench

pb = bench.load_pb()

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sysm = bench.workloads().lj_charge_fluid(ncell)
e = pb.Ewald(device=0)
info = e.init(1e-4, sysm.q, 12.0, sysm.boxlo, sysm.boxhi)
dev = torch.device("cuda", 0)
x = torch.tensor(sysm.x, dtype=torch.float64, device=dev)
q = torch.tensor(sysm.q, dtype=torch.float64, device=dev)
f = torch.zeros((sysm.n, 3), dtype=torch.float64, device=dev)
ms = []
for k in range(reps + 3):
    f.zero_()
    torch.cuda.synchronize()
    energy, virial = e.compute_device(sysm.n, x.data_ptr(), q.data_ptr(), f.data_ptr())
    if k >= 3:
        ms.append(e.last_ms())
print(f"device Ewald: {sysm.n} atoms, kcount {info.kcount}, kmax {info.kmax}, g {info.g_ewald:.6f}: "
      f"{np.mean(ms):.3f} ms per compute (min {np.min(ms):.3f}), E_long {energy:.10f}")
if len(sys.argv) > 3:  # optional: the oracle's direct sums on the host cores beside it
    from oracle import polref as P
    prd = sysm.boxhi - sysm.boxlo
    plan = P.ewald_plan(1e-4, sysm.q, 12.0, prd)
    t0 = time.time()
    ref = P.ewald_compute(plan, sysm.x, sysm.q, prd)
    print(f"oracle (OpenMP, direct sums): {time.time() - t0:.2f} s, E_long {ref['energy']:.10f}, "
          f"force err {np.abs(f.cpu().numpy() - ref['f']).max() / np.abs(ref['f']).max():.2e}")
e.close()
