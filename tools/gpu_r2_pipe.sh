#!/bin/bash
# round 2, session 4: the overlapped anneal + results call (staggered row chunks) -- parity, then e2e at several staggers
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
t0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 \
  -k "staggered or config5 or many_rows" > gpurun_out/t_pipe.log 2>&1
echo "pipe tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -5 gpurun_out/t_pipe.log
for lag in off auto 8 16 24 32 48; do
  if [ $lag = off ]; then export PIQMC_PIPE=0; unset PIQMC_PIPE_LAG16
  elif [ $lag = auto ]; then unset PIQMC_PIPE; unset PIQMC_PIPE_LAG16
  else unset PIQMC_PIPE; export PIQMC_PIPE_LAG16=$lag; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/pipe_s20_$lag.json 2> gpurun_out/pipe_s20_$lag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/pipe_s20_$lag.json").read().strip().splitlines()[-1])
    print("lag $lag: value %.3e e2e %.3e (%.1f ms) %s" % (d["value"], d["e2e"]["value"], 1e3 * d["e2e"]["seconds"], {k: round(1e3 * v, 1) for k, v in d["e2e"]["breakdown_s"].items()}))
except Exception as e:
    print("lag $lag failed", e)
PY
done
unset PIQMC_PIPE PIQMC_PIPE_LAG16
timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu > gpurun_out/pipe_s100_auto.json 2> gpurun_out/pipe_s100_auto.err
python -c "
import json
d = json.loads(open('gpurun_out/pipe_s100_auto.json').read().strip().splitlines()[-1])
print('steps 100 auto: value %.3e e2e %.3e' % (d['value'], d['e2e']['value']), d['e2e']['breakdown_s'])"
echo "elapsed $(( $(date +%s) - t0 )) s"
