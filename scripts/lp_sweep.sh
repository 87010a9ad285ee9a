#!/bin/bash
# development: layer-parallel schedule against one launch per layer, per workload (run under gpurun)
OUT=gpurun_out
for mt in "" 0 96; do
  export FS_LP_MAX_TILES=$mt; [ -z "$mt" ] && unset FS_LP_MAX_TILES
  echo "== FS_LP_MAX_TILES=${mt:-unset}"
  FS_BENCH_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 20 > $OUT/lp_${mt:-all}.json 2> $OUT/lp_${mt:-all}.err
  grep "\[bench\]" $OUT/lp_${mt:-all}.err | cut -c1-260
  python -c "
import json
d=json.load(open('$OUT/lp_${mt:-all}.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phases_ms'])
"
done
