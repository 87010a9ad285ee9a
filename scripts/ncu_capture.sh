#!/bin/bash
# ncu evidence for one round tag (run under gpurun, 1 GPU):  bash scripts/ncu_capture.sh r02b
# 1) launch list of the default bench command (device time per launch; cold-cache, serialised: compare shares)
# 2) --set full captures of the kernels of the hot path (conditioner with fused spline, local sweep, total energy,
#    global accept, the two feature / unconditional-spline kernels)
# Every ncu run is preceded by the same command without ncu (B200_PROFILING.md).
set -u
TAG=${1:-r02f}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
OUT=gpurun_out
$CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_l_$TAG.log 2>&1
echo "launch list rc=$?"
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 $CMD > $OUT/ncu_f_${TAG}_$1.log 2>&1
  echo "$1 rc=$?"
}
cap cond "tc_conditioner_kernel" 62 6
cap sweep "local_sweep_fast_kernel" 3 1
cap energy "energy_total_kernel" 3 1
cap accept "accept_global_kernel" 3 1
cap prep "prep_v3" 40 2
ls -la $OUT/*$TAG*
