#!/bin/bash
# A/B of the whole-step heat kernel inside ONE gpurun call (clocks differ from box to box under the power cap)
for rep in 1 2; do for v in $1; do
  so=""; [ "$v" != base ] && so=$PWD/vec-ode_b200/variants/libvecode_b200_$v.so
  VECODE_B200_SO=$so python bench.py --workload heat_rk4_fused --arith fast --steps 300 --warmup 20 --no-cpu --no-also --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('$v rep$rep', round(j['ms_per_step'],4),'ms', j['clocks']['sm_mhz'], j['clocks']['reasons'])"
done; done
