#!/bin/bash
# 8-GPU evidence runs (BASELINE configs 4/5 shapes + the headline bench); outputs in gpurun_out/
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29701 tools/scale_run.py --ncell 32 --steps 3 --mode precision 2> gpurun_out/s8_prec.err | grep '^{' > gpurun_out/scale_8gpu_1M_precision.json
$TR --master-port 29702 tools/scale_run.py --ncell 32 --steps 2 --mode ranked 2> gpurun_out/s8_rank.err | grep '^{' > gpurun_out/scale_8gpu_1M_ranked.json
$TR --master-port 29703 bench.py --gpus 8 --steps 20 --warmup 5 2> gpurun_out/b8.err | grep '^{' > gpurun_out/b8_tma.json
$TR --master-port 29704 tools/scale_run.py --ncell 63 --steps 3 --mode fixed --samples 16 2> gpurun_out/s8_8M.err | grep '^{' > gpurun_out/scale_8gpu_8M_fixed.json
for f in gpurun_out/scale_8gpu_1M_precision.json gpurun_out/scale_8gpu_1M_ranked.json gpurun_out/scale_8gpu_8M_fixed.json; do
python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], d["atoms_total"], d["mode"], "ms/step", [round(v, 1) for v in d["ms_per_step_max_over_ranks"]],
          "iters", [s["iterations"] for s in d["steps"]], "check", d["check"], "hbm_gb", round(d["hbm_used_gb_rank0"], 1))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
python tools/show_bench.py gpurun_out/b8_tma.json
tail -3 gpurun_out/s8_8M.err
