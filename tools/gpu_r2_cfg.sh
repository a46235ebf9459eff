#!/bin/bash
export PIQMC_WATCHDOG_MS=5000
mkdir -p gpurun_out
for R in 512 4096; do
timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu --replicas $R > gpurun_out/p.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/p.log").read().strip().splitlines()[-1])
print("R=$R 100 steps chain: timed ms/step %7.3f  e2e sweeps ms/step %7.3f" % (d["ms_per_step"], 1e3*d["e2e"]["breakdown_s"]["sweeps"]/d["steps"]))
PY
done
echo "--- chain"; timeout 600 python tools/bench_configs.py 8192 2>&1 | grep natural
echo "--- dataflow"; PIQMC_CHAIN=0 timeout 600 python tools/bench_configs.py 8192 2>&1 | grep natural
echo "--- chain R=1000"; timeout 600 python tools/bench_configs.py 1000 2>&1 | grep natural
echo "--- dataflow R=1000"; PIQMC_CHAIN=0 timeout 600 python tools/bench_configs.py 1000 2>&1 | grep natural
echo "--- chain R=8192 rpt1"; PIQMC_CHAIN_RPT=1 timeout 600 python tools/bench_configs.py 8192 2>&1 | grep natural
