#!/usr/bin/env python3
"""Device PPPM vs device Ewald, inputs resident in HBM: CUDA-event time per compute on BASELINE config 2 (32 000-atom
fluid) and config 3 (255 552-atom water box), accuracy 1e-4, cutoff 12 A."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import bench

pb = bench.load_pb()
W = bench.workloads()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda", 0)
for label, sysm in (("config 2 fluid", W.lj_charge_fluid(20)), ("config 3 water", W.water_box(44))):
    x = torch.tensor(sysm.x, dtype=torch.float64, device=dev)
    q = torch.tensor(sysm.q, dtype=torch.float64, device=dev)
    f = torch.zeros((sysm.n, 3), dtype=torch.float64, device=dev)
    out = {}
    for name, cls in (("pppm", pb.PPPM), ("ewald", pb.Ewald)):
        k = cls(device=0)
        info = k.init(1e-4, sysm.q, 12.0, sysm.boxlo, sysm.boxhi)
        ms = []
        for r in range(reps + 3 if name == "pppm" else 5):
            f.zero_()
            torch.cuda.synchronize()
            e, v = k.compute_device(sysm.n, x.data_ptr(), q.data_ptr(), f.data_ptr())
            if r >= 2:
                ms.append(k.last_ms())
        torch.cuda.synchronize()
        out[name] = (float(np.mean(ms)), float(np.min(ms)), e, f.clone(), info)
        k.close()
    ip = out["pppm"][4]
    df = (out["pppm"][3] - out["ewald"][3]).abs().max().item() / out["ewald"][3].abs().max().item()
    print(f"{label}: {sysm.n} atoms | PPPM grid {ip.nx}x{ip.ny}x{ip.nz} order {ip.order} g {ip.g_ewald:.5f}: {out['pppm'][0]:.3f} ms "
          f"(min {out['pppm'][1]:.3f}) | Ewald kcount {out['ewald'][4].kcount} g {out['ewald'][4].g_ewald:.5f}: {out['ewald'][0]:.3f} ms | "
          f"E_long {out['pppm'][2]:.6f} vs {out['ewald'][2]:.6f} (different g_ewald: only real + reciprocal agree), max force diff {df:.2e}")
