#!/bin/bash
# two rows per thread (colour_sweep_fast2): parity, then rates against the one-row kernel
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
t0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 -k "config5 or many_rows or gaussian_torus or sharding or bit_exact" > gpurun_out/t_fast2.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/t_fast2.log
PIQMC_ROWS_PER_BLOCK=256 timeout 900 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 -k "many_rows or gaussian_torus or sharding or qa_colour_bit_exact or level_colouring" > gpurun_out/t_fast2b.log 2>&1
echo "tests (256-row units) rc=$?"; tail -2 gpurun_out/t_fast2b.log
PIQMC_FORCE_GENERIC_FN=1 PIQMC_ROWS_PER_BLOCK=256 timeout 900 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 -k "many_rows or gaussian_torus" > gpurun_out/t_fast2c.log 2>&1
echo "tests (generic functions) rc=$?"; tail -2 gpurun_out/t_fast2c.log
run() {
  name=$1; rep=$2; steps=$3; shift 3
  env "$@" timeout 300 python bench.py --steps $steps --warmup 3 --no-cpu --replicas $rep > gpurun_out/fast2_$name.json 2> gpurun_out/fast2_$name.err
  python -c "
import json
d = json.loads(open('gpurun_out/fast2_$name.json').read().strip().splitlines()[-1])
print('$name: value %.3e ms/sweep %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
}
run r4096_one 4096 20 PIQMC_FAST2=0
run r4096_two6 4096 20 X=1
run r4096_two5 4096 20 PIQMC_MINB2=5
run r4096_two7 4096 20 PIQMC_MINB2=7
run r4096_two6_rpb256 4096 20 PIQMC_ROWS_PER_BLOCK=256
run r4096_two6_rpb1024 4096 20 PIQMC_ROWS_PER_BLOCK=1024
run r2048_one 2048 50 PIQMC_FAST2=0
run r2048_two 2048 50 X=1
run r1024_one 1024 50 PIQMC_FAST2=0
run r1024_two 1024 50 X=1
run r512_one 512 50 X=1
run r512_two_rpb256 512 50 PIQMC_ROWS_PER_BLOCK=256
run r512_two5_rpb256 512 50 PIQMC_ROWS_PER_BLOCK=256 PIQMC_MINB2=5
echo "elapsed $(( $(date +%s) - t0 )) s"
