#!/bin/bash
# Launch time against trajectories per launch (fixed overhead vs streaming rate): tools/nsweep.sh <workload> <arith> "<n...>"
for n in $3; do
  python bench.py --workload $1 --arith $2 --n-traj $n --steps 3000 --warmup 100 --no-cpu --no-also --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('$1 $2 n=$n', round(j['roofline']['kernel_us'],2),'us frac',round(j['roofline']['frac'],3))"
done
