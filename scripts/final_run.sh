set -u
OUT=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/r2_gpu_tests_final.log 2>&1; tail -3 $OUT/r2_gpu_tests_final.log
timeout 900 python bench.py > $OUT/r2_final_bench.json 2> $OUT/r2_final_bench.err; echo bench rc=$?; tail -2 $OUT/r2_final_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r2_final_ref.json 2> $OUT/r2_final_ref.err; echo ref rc=$?; tail -2 $OUT/r2_final_ref.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
$CMD > $OUT/plain_r02c.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_r02c.csv $CMD > $OUT/ncu_l_r02c.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on -k "regex:tc_conditioner_kernel" -s 75 -c 8 -f -o $OUT/prof_r02c_cond $CMD > $OUT/ncu_f_r02c_cond.log 2>&1; echo cond rc=$?
python -c "
import json
d=json.load(open('$OUT/r2_final_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['launch_ms'], d['cpu_baseline'], d.get('parity_checked'))
for k,v in d['secondary'].items():
    if k!='energy_sweep': print(k, v['value'], v['ms_per_step'], v['roofline']['frac'], v['roofline']['launch_ms'])
    else: print(k, [(e['N'], round(e['frac'],3)) for e in v] if isinstance(v,list) else v)
print(open('$OUT/r2_final_ref.json').read()[:600])
"
