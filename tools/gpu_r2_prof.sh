#!/bin/bash
export PIQMC_WATCHDOG_MS=5000
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_chain.py -q --timeout 300 > gpurun_out/t_chain.log 2>&1
echo "chain tests rc=$?"; tail -3 gpurun_out/t_chain.log
for R in 512 4096; do
echo "=== R=$R"
PIQMC_CHAIN_PROF=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --replicas $R > gpurun_out/p.log 2>&1; grep "chain prof" gpurun_out/p.log | tail -11
python - <<PY
import json
d=json.loads(open("gpurun_out/p.log").read().strip().splitlines()[-1])
print("timed ms/step %7.3f   e2e sweeps ms/step %7.3f" % (d["ms_per_step"], 1e3*d["e2e"]["breakdown_s"]["sweeps"]/d["steps"]))
PY
done
echo "=== R=512 RPT=1"
PIQMC_CHAIN_RPT=1 PIQMC_CHAIN_PROF=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --replicas 512 > gpurun_out/p.log 2>&1; grep "chain prof" gpurun_out/p.log | tail -11
for R in 512 4096; do
timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu --replicas $R > gpurun_out/p.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/p.log").read().strip().splitlines()[-1])
print("R=$R 100 steps, no prof: timed ms/step %7.3f value %.3e  e2e sweeps ms/step %7.3f" % (d["ms_per_step"], d["value"], 1e3*d["e2e"]["breakdown_s"]["sweeps"]/d["steps"]))
PY
done
