for cfg in 2 3; do
for env in "POLB200_GROUP_PAIRS=0" "POLB200_GPF_MINB=4" "POLB200_GPF_MINB=5"; do
  echo "== config $cfg $env"
  env $env python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms'].items()}, round(d['us_per_dipole_iteration'],1), d['check']['eng_pol'])"
done; done
