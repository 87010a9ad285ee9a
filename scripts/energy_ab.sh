set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_energy.py tests/test_gpu_global.py tests/test_gpu_sweep.py -x -q -m gpu > $OUT/e_tests.log 2>&1; tail -8 $OUT/e_tests.log
echo "== default"; timeout 300 python scripts/energy_sweep.py 2>&1 | tee $OUT/e_fold.log | cut -c1-150
for G in 16 32; do echo "== G=$G"; FS_ENERGY_G=$G timeout 300 python scripts/energy_sweep.py 2>&1 | head -3 | cut -c1-150; done
