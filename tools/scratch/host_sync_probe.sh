#!/bin/bash
# e2e of the headline workload under the three host wait policies, with all hardware threads and pinned to two (5 host threads on 2
# hardware threads = the oversubscription of 8 ranks x 5 threads on a 16-thread box). Prints policy, cpus, ms per e2e solve.
for cpus in all 0-1; do
  for mode in spin yield blocking; do
    if [ "$cpus" = all ]; then pre=""; else pre="taskset -c $cpus"; fi
    VECODE_BENCH_HOST_SYNC=$mode $pre python bench.py --steps 20 --warmup 5 --no-also --no-cpu --e2e-steps 12 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$mode', '$cpus', d['config'].get('host_sync'), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))
"
  done
done
