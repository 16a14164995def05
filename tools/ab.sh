#!/bin/bash
# A/B of experimental builds on the GPU box: tools/ab.sh "<variants>" "<bench args>" — prints us per launch and roofline frac.
for v in $1; do
  so=""; [ "$v" != base ] && so=$PWD/vec-ode_b200/variants/libvecode_b200_$v.so
  for ar in fast strict; do
    VECODE_B200_SO=$so python bench.py $2 --arith $ar --no-cpu --no-also --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import sys,json
try:
    j=json.loads(sys.stdin.read()); print('$v $ar', round(j['roofline']['kernel_us'],2),'us frac',round(j['roofline']['frac'],3), 'clk', j['clocks']['sm_mhz'], j['clocks']['reasons'])
except Exception as e: print('$v $ar FAILED', e)"
  done
done
