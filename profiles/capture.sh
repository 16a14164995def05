#!/bin/bash
# ncu captures for one workload (run under gpurun, ONE GPU). Usage: profiles/capture.sh <workload> <tag> [kernel-regex]
# 1) plain run must exit 0; 2) launch list with per-launch device time; 3) one --set full capture of the dominant kernel.
set -u
WL=$1; TAG=$2; KRE=${3:-rk_small_kernel}
CMD="python bench.py --workload $WL ${EXTRA:-} --steps ${STEPS:-200} --warmup ${WARM:-20} --no-cpu --no-also --e2e-steps 1"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-40} -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s ${SKIPF:-30} -c ${CNT:-3} -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "capture $TAG rc=$?"
