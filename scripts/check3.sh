set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_drivers.py -x -q -m gpu > $OUT/t3.log 2>&1; tail -12 $OUT/t3.log
python scripts/train_time.py --native-only
FS_TRAIN_CHAIN_SCALAR=1 python scripts/train_time.py --native-only
timeout 600 python bench.py --workload alg2_n64 --steps 10 --warmup 4 --no-cpu-baseline --no-secondary > $OUT/b_alg2.json 2> $OUT/b_alg2.err; tail -3 $OUT/b_alg2.err
python -c "
import json; d=json.load(open('$OUT/b_alg2.json')); print(d['value'], d['ms_per_step'], d['phases_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:bn_bwd|bn_stats|chain_|features_|gather_uncond|gemm_n|gemm_t|loss_kernel|scatter_uncond|adam_" --launch-skip 200 -c 80 --csv --log-file $OUT/launches_train.csv python scripts/train_time.py --native-only > $OUT/ncu_train.log 2>&1
