set -u
OUT=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=15 > $OUT/r2f_gpu_tests.log 2>&1; tail -25 $OUT/r2f_gpu_tests.log
timeout 900 python bench.py > $OUT/r2f_bench.json 2> $OUT/r2f_bench.err; echo bench rc=$?; tail -2 $OUT/r2f_bench.err
python -c "
import json
d=json.load(open('$OUT/r2f_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['launch_ms'], d['cpu_baseline'], d.get('parity_checked'))
for k,v in d['secondary'].items():
    if k!='energy_sweep': print(k, v['value'], v['ms_per_step'], v['roofline']['frac'], v['roofline']['launch_ms'], v.get('phases_ms'))
    else: print(k, [(e['N'], round(e['frac'],3)) for e in v] if isinstance(v,list) else v)
"
