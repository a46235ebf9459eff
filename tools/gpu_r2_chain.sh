#!/bin/bash
# chain pipeline on the GPU: parity, then quick throughput at 4096 and 512 rows
export PIQMC_WATCHDOG_MS=5000
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_chain.py -q -x --timeout 300 > gpurun_out/t_chain.log 2>&1
echo "chain tests rc=$?" >> gpurun_out/t_chain.log
tail -5 gpurun_out/t_chain.log
for R in 4096 512; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --replicas $R > gpurun_out/b_chain_R$R.log 2>&1
  echo "R=$R rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/b_chain_R$R.log").read().strip().splitlines()[-1])
    print("value %.3e ms/step %.3f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["e2e"].get("breakdown_s"))
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/b_chain_R$R.log").read()[-1500:])
PY
done
