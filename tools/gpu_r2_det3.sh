#!/bin/bash
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
timeout 1200 python -m pytest tests/test_gpu_det.py tests/test_gpu_reference_suite.py -q -x --timeout 600 > gpurun_out/t_det3.log 2>&1
echo "det tests rc=$?"; tail -3 gpurun_out/t_det3.log
python tools/bench_det.py 1000 4096 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 -k "staggered or many_rows or config5_shard" > gpurun_out/t_pipe4.log 2>&1
echo "pipe tests rc=$?"; tail -2 gpurun_out/t_pipe4.log
for i in 1 2; do
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/b20_$i.json 2>/dev/null
python -c "
import json
d = json.loads(open('gpurun_out/b20_$i.json').read().strip().splitlines()[-1])
print('value %.3e e2e %.3e' % (d['value'], d['e2e']['value']), {k: round(1e3 * v, 1) for k, v in d['e2e']['breakdown_s'].items()})"
done
