#!/bin/bash
# ncu evidence of round 2 (one GPU; every command first runs WITHOUT ncu and must exit 0).  The reports are summarised on
# the box (tools/ncu_summary.py); only the summaries and the two sweep reports travel back (gpurun_out is capped at 64 MiB).
set -x
O=gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also"
B2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --config 2"
$B --config 5 > $O/ncu_plain5.json 2> $O/ncu_plain5.err || exit 1
$B --config 3 > $O/ncu_plain3.json 2> $O/ncu_plain3.err || exit 1
$B2 > $O/ncu_plain2.json 2> $O/ncu_plain2.err || exit 1
# launch list of the headline command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_config5.csv $B --config 5 > $O/ncu_l5.log 2>&1
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip --launch-count $cnt -f -o $O/$name "$@" > $O/ncu_$name.log 2>&1
  python tools/ncu_summary.py $O/$name.ncu-rep $O/$name.md > /dev/null 2>&1
}
cap r2_sweep_tma_1M k_sweep_group_tma 70 1 $B --config 5
cap r2_sweep_tma_gs_256k k_sweep_group_tma 150 1 $B --config 3
cap r2_pair_force_group_256k 'k_polforce_group|k_pair_group|k_group_cache' 3 3 $B --config 3
rm -f $O/r2_pair_force_group_256k.ncu-rep
cap r2_sweep_tma_32k k_sweep_group_tma 70 1 $B2
rm -f $O/r2_sweep_tma_32k.ncu-rep
cap r2_pair_force_group_32k 'k_polforce_group|k_pair_group|k_group_cache' 3 3 $B2
rm -f $O/r2_pair_force_group_32k.ncu-rep
ls -la $O/
