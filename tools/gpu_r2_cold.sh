#!/bin/bash
export PIQMC_WATCHDOG_MS=8000
mkdir -p gpurun_out
run() { name=$1; R=$2; shift 2
  env "$@" timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --replicas $R > gpurun_out/m_$name.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/m_$name.log").read().strip().splitlines()[-1])
    print("%-22s R=%-5s e2e sweeps ms/step %.3f  timed %.3f" % ("$name", "$R", 1e3*d["e2e"]["breakdown_s"]["sweeps"]/d["steps"], d["ms_per_step"]))
except Exception as e:
    print("$name failed", e, open("gpurun_out/m_$name.log").read()[-300:])
PY
}
run chain_T01 512 PIQMC_CHAIN=1
run chain_T001 512 PIQMC_CHAIN=1 PIQMC_BENCH_TEMP=0.001
run chain_T0001 512 PIQMC_CHAIN=1 PIQMC_BENCH_TEMP=0.0001
run flow_T01 512 A=1
run flow_T001 512 PIQMC_BENCH_TEMP=0.001
run flow_T0001 512 PIQMC_BENCH_TEMP=0.0001
run chain4096_T0001 4096 PIQMC_CHAIN=1 PIQMC_BENCH_TEMP=0.0001
run flow4096_T0001 4096 PIQMC_BENCH_TEMP=0.0001
run flow4096_T01 4096 A=1
