#!/bin/bash
# the overlapped anneal + results path where the library arms it by itself: 16384 rows on one GPU (8.6 GB state)
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=60000
for mode in off auto; do
  if [ $mode = off ]; then export PIQMC_PIPE=0; else unset PIQMC_PIPE; export PIQMC_PIPE_TRACE=1; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --replicas 16384 > gpurun_out/pipe_big_$mode.json 2> gpurun_out/pipe_big_$mode.err
  echo "== $mode rc=$?"; grep "piqmc pipe\]" gpurun_out/pipe_big_$mode.err | head -3 | cut -c1-150
  python -c "
import json
d = json.loads(open('gpurun_out/pipe_big_$mode.json').read().strip().splitlines()[-1])
print('value %.3e e2e %.3e (%.1f ms) overlapped %s' % (d['value'], d['e2e']['value'], 1e3 * d['e2e']['seconds'], d['e2e']['overlapped_download']), {k: round(1e3 * v, 1) for k, v in d['e2e']['breakdown_s'].items()})"
done
