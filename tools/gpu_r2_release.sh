#!/bin/bash
# where the release fence sits at the end of a unit: 0 thread 0 after the barrier (committed), 1 every thread before
# the barrier as well, 2 every thread before the barrier and a relaxed flag store.  The experiment was three lines at the
# end of colour_sweep_fast (not kept):   if (a.release_mode != 0) __threadfence();   before the closing __syncthreads(),
# and for mode 2   st.relaxed.gpu.global.u32   instead of st_release for the flag; a.release_mode = PIQMC_RELEASE_MODE.
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
for mode in 0 1 2; do
  export PIQMC_RELEASE_MODE=$mode
  for rep in 512 1024 4096; do
    python bench.py --steps 50 --warmup 3 --no-cpu --replicas $rep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mode $mode rows $rep: value %.3e ms/sweep %.3f' % (d['value'], d['ms_per_step']))"
  done
done
PIQMC_RELEASE_MODE=2 timeout 600 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 -k "many_rows or config5 or sharding" 2>&1 | tail -2
