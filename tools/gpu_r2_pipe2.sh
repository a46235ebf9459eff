#!/bin/bash
# trace of the overlapped anneal + results call: when chunks finish, when their downloads / energy reductions run
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000 PIQMC_PIPE_TRACE=1
for cfg in "24 chunk" "24 end" "0 end" "40 end"; do
  set -- $cfg
  export PIQMC_PIPE_LAG16=$1 PIQMC_PIPE_ENERGY=$2
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/pipe2_$1_$2.json 2> gpurun_out/pipe2_$1_$2.err
  echo "== lag16 $1 energy $2"; grep "piqmc pipe" gpurun_out/pipe2_$1_$2.err | tail -9
  python -c "
import json
d = json.loads(open('gpurun_out/pipe2_$1_$2.json').read().strip().splitlines()[-1])
print('value %.3e e2e %.3e (%.1f ms)' % (d['value'], d['e2e']['value'], 1e3 * d['e2e']['seconds']))"
done
