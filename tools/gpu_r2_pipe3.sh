#!/bin/bash
# compact tickets for the staggered chunks: parity, then e2e against the stagger
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
timeout 900 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 -k "staggered or many_rows or config5_shard" > gpurun_out/t_pipe3.log 2>&1
echo "pipe tests rc=$?"; tail -3 gpurun_out/t_pipe3.log
export PIQMC_PIPE_TRACE=1
for lag in 0 16 24 32 40 48 56 64 73; do
  export PIQMC_PIPE_LAG16=$lag
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/pipe3_$lag.json 2> gpurun_out/pipe3_$lag.err
  echo "== lag16 $lag"; grep "piqmc pipe" gpurun_out/pipe3_$lag.err | tail -9 | cut -c1-110
  python -c "
import json
d = json.loads(open('gpurun_out/pipe3_$lag.json').read().strip().splitlines()[-1])
print('value %.3e e2e %.3e (%.1f ms)' % (d['value'], d['e2e']['value'], 1e3 * d['e2e']['seconds']), {k: round(1e3 * v, 1) for k, v in d['e2e']['breakdown_s'].items()})"
done
