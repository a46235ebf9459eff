#!/bin/bash
# single-class fast path ahead of the class loop (all lanes of the warp in Trotter class 0): parity, then rates
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
timeout 1200 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 > gpurun_out/t_class0.log 2>&1
echo "colour tests rc=$?"; tail -2 gpurun_out/t_class0.log
PIQMC_FORCE_GENERIC_FN=1 timeout 600 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 -k "bit_exact and not resident" > gpurun_out/t_class0_generic.log 2>&1
echo "generic-fn tests rc=$?"; tail -1 gpurun_out/t_class0_generic.log
for rep in 4096 2048 1024 512; do
  python bench.py --steps 50 --warmup 3 --no-cpu --replicas $rep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rows $rep: value %.3e ms/sweep %.3f' % (d['value'], d['ms_per_step']))"
done
python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('steps 20: value %.3e e2e %.3e' % (d['value'], d['e2e']['value']))"
