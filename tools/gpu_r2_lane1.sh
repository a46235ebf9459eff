#!/bin/bash
# slice 1 decided from the class-0 / class-1 functions (evaluated for all lanes, class 0 reused by the other lanes)
# instead of per-word truth-table look-ups: parity, then rates (the experiment is described in
# profiles/r2_e2e_overlap.md section 8; measured slower, not kept)
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
timeout 900 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 > gpurun_out/t_lane1.log 2>&1
echo "colour tests rc=$?"; tail -1 gpurun_out/t_lane1.log
PIQMC_FORCE_GENERIC_FN=1 timeout 600 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 -k "many_rows or gaussian_torus or qa_colour_bit_exact" > gpurun_out/t_lane1_generic.log 2>&1
echo "generic-fn tests rc=$?"; tail -1 gpurun_out/t_lane1_generic.log
for rep in 4096 512; do
  python bench.py --steps 20 --warmup 3 --no-cpu --replicas $rep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rows $rep: value %.3e ms/sweep %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
