#!/bin/bash
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
timeout 600 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 -k "many_rows or config5_shard or staggered" > gpurun_out/t_tabs4.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/t_tabs4.log
run() {
  name=$1; rep=$2; steps=$3; shift 3
  env "$@" timeout 300 python bench.py --steps $steps --warmup 3 --no-cpu --replicas $rep > gpurun_out/tabs4_$name.json 2> gpurun_out/tabs4_$name.err
  python -c "
import json
d = json.loads(open('gpurun_out/tabs4_$name.json').read().strip().splitlines()[-1])
print('$name: value %.3e ms/sweep %.3f e2e %.3e launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))"
}
run r4096_s20 4096 20 X=1
run r4096_s50 4096 50 X=1
run r4096_s50_minb8 4096 50 PIQMC_MINB=8
run r4096_s50_rpb256 4096 50 PIQMC_ROWS_PER_BLOCK=256
run r2048_s50 2048 50 X=1
run r1024_s50 1024 50 X=1
run r512_s50 512 50 X=1
run r512_s50_a0 512 50 PIQMC_TAB_AHEAD=4000000000
