#!/bin/bash
# strong scaling of ONE 1 048 576-atom system (fixed 30 Jacobi sweeps) over 1, 2, 4, 8 GPUs
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python tools/scale_run.py --global-ncell 64 --steps 3 --mode fixed --samples 8 2>/dev/null | grep '^{' > gpurun_out/strong_${n}.json
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2980$n tools/scale_run.py --global-ncell 64 --steps 3 --mode fixed --samples 8 2>/dev/null | grep '^{' > gpurun_out/strong_${n}.json; fi
  python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/strong_{n}.json"))
    print(n, "GPUs:", d["atoms_total"], "atoms, ms/step", [round(v, 1) for v in d["ms_per_step_max_over_ranks"]], "atom-steps/s", f"{d['atom_steps_per_s_last']:.3g}", "hbm GB", round(d["hbm_used_gb_rank0"], 1), "E_pol", d["eng_pol_total"])
except Exception as e:
    print(n, "FAILED", e)
PY
done
