#!/bin/bash
# final 8-GPU evidence of round 2: decomposed parity (pair style, KSpace, rigid MD step), weak-scaling bench, MD step timing
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29801 tests/mgpu_md_check.py > $O/final8_md_check.log 2>&1; echo "md_check rc=$?"; grep "mgpu md\|all md\|FAILED" $O/final8_md_check.log
MGPU_NCELL=8 MGPU_CUT=8.0 timeout 400 $TR --master-port 29802 tests/mgpu_check.py > $O/final8_mgpu_check.log 2>&1; echo "mgpu_check rc=$?"; grep "OK\|FAIL\|passed" $O/final8_mgpu_check.log | cut -c1-150
timeout 400 $TR --master-port 29803 bench.py --gpus 8 --steps 12 --warmup 3 > $O/final8_bench.json 2> $O/final8_bench.err; echo "bench rc=$?"; python tools/show_bench.py $O/final8_bench.json
timeout 300 $TR --master-port 29804 tools/mgpu_md_timing.py 88 8 > $O/final8_md_timing.log 2>&1; echo "md_timing rc=$?"; grep mgpu_md_timing $O/final8_md_timing.log
