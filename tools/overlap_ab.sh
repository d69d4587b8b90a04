# A/B of the optimiser-update overlap (UDH_OVERLAP_UPDATE), its Adam grid and the SM reservation of the conv4_x backward at N = 1
for v in "UDH_OVERLAP_UPDATE=0" "UDH_SIDE_ADAM_GRID=148 UDH_SM_RESERVE_TOP=0" "UDH_SIDE_ADAM_GRID=296 UDH_SM_RESERVE_TOP=0" "UDH_SIDE_ADAM_GRID=74 UDH_SM_RESERVE_TOP=0" "UDH_SIDE_ADAM_GRID=148 UDH_SM_RESERVE_TOP=16" "UDH_SIDE_ADAM_GRID=592 UDH_SM_RESERVE_TOP=0" "UDH_OVERLAP_UPDATE=0"; do
  echo "== $v"
  env $v timeout 300 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{\"metric\"'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['clocks']['reasons'], {k:v for k,v in list(d['phases_ms_per_step'].items()) if k in ('adam','conv4_2.dgrad','conv4_2.wgrad','conv4_1.dgrad','conv4_1.wgrad','pool.bwd','conv3_2.dgrad','conv3_2.wgrad')})
"
done
