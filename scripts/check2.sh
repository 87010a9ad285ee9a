set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_global.py tests/test_gpu_drivers.py tests/test_gpu_target.py -x -q -m gpu > $OUT/t2.log 2>&1; tail -12 $OUT/t2.log
timeout 600 python -m pytest tests/test_gpu_flow.py -x -q -m gpu -k "sampling_pass" -s 2>&1 | grep -E "log q|passed|failed|Error"
python scripts/train_time.py --native-only
timeout 600 python bench.py --workload alg2_n64 --steps 10 --warmup 4 --no-cpu-baseline --no-secondary > $OUT/b_alg2.json 2> $OUT/b_alg2.err; tail -3 $OUT/b_alg2.err
python -c "
import json; d=json.load(open('$OUT/b_alg2.json')); print(d['value'], d['ms_per_step'], d['phases_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 6000 -c 200 --csv --log-file $OUT/launches_train.csv python scripts/train_time.py --native-only > $OUT/ncu_train.log 2>&1; echo ncu rc=$?
