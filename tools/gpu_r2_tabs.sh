#!/bin/bash
# decision tables from their own kernel (once per (sweep, spin) instead of once per unit): parity, then rates
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
t0=$(date +%s)
timeout 1200 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 > gpurun_out/t_tabs.log 2>&1
echo "colour tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -4 gpurun_out/t_tabs.log
PIQMC_FORCE_GENERIC_FN=1 timeout 600 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 -k "bit_exact and not resident" > gpurun_out/t_tabs_generic.log 2>&1
echo "generic-fn tests rc=$?"; tail -2 gpurun_out/t_tabs_generic.log
run() {  # name, replicas, steps, env...
  name=$1; rep=$2; steps=$3; shift 3
  env "$@" timeout 300 python bench.py --steps $steps --warmup 3 --no-cpu --replicas $rep > gpurun_out/tabs_$name.json 2> gpurun_out/tabs_$name.err
  python -c "
import json
d = json.loads(open('gpurun_out/tabs_$name.json').read().strip().splitlines()[-1])
print('$name: value %.3e ms/sweep %.3f e2e %.3e launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))"
}
run r4096_s20 4096 20 X=1
run r4096_s50 4096 50 X=1
run r4096_s50_rpb256 4096 50 PIQMC_ROWS_PER_BLOCK=256
run r4096_s50_rpb1024 4096 50 PIQMC_ROWS_PER_BLOCK=1024
run r4096_s50_minb8 4096 50 PIQMC_MINB=8
run r2048_s50 2048 50 X=1
run r1024_s50 1024 50 X=1
run r512_s50 512 50 X=1
run r512_s50_rpb256 512 50 PIQMC_ROWS_PER_BLOCK=256
run r512_s50_minb8 512 50 PIQMC_MINB=8
echo "elapsed $(( $(date +%s) - t0 )) s"
