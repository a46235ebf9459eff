#!/bin/bash
# dependency poll with relaxed loads + one acquire fence (PIQMC_POLL_RELAXED=1) against acquire loads (an acquire
# load is followed by an L1 invalidate, CCTL.IVALL, at every poll)
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
for mode in 0 1; do
  export PIQMC_POLL_RELAXED=$mode
  for rep in 512 1024 2048 4096; do
    python bench.py --steps 50 --warmup 3 --no-cpu --replicas $rep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('relaxed $mode rows $rep: value %.3e ms/sweep %.3f' % (d['value'], d['ms_per_step']))"
  done
done
PIQMC_POLL_RELAXED=1 timeout 600 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 -k "many_rows or config5 or sharding" 2>&1 | tail -2
