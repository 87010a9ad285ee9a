#!/bin/bash
# extra ncu evidence (run under gpurun, 1 GPU): the layer-parallel conditioner launch of alg1_n32 and the training kernels
set -u
OUT=gpurun_out
TAG=${1:-r02f}
CMD="python bench.py --workload alg1_n32 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
$CMD > $OUT/plain_${TAG}_n32.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:tc_conditioner_kernel" -s 6 -c 2 -f -o $OUT/prof_${TAG}_cond_lp $CMD > $OUT/ncu_f_${TAG}_cond_lp.log 2>&1; echo "cond_lp rc=$?"
T="python scripts/train_time.py --native-only"
$T > $OUT/plain_${TAG}_train.log 2>&1 || { echo plain train failed; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:chain_lanes|gemm_nn|gemm_tn|gemm_nt|adam_apply" -s 40 -c 24 -f -o $OUT/prof_${TAG}_train $T > $OUT/ncu_f_${TAG}_train.log 2>&1; echo "train rc=$?"
