#!/bin/bash
export PIQMC_WATCHDOG_MS=8000
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_level.py -q -x --timeout 600 > gpurun_out/t_level.log 2>&1
echo "level tests rc=$?"; tail -8 gpurun_out/t_level.log
run() { name=$1; R=$2; shift 2
  env "$@" timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --replicas $R > gpurun_out/m_$name.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/m_$name.log").read().strip().splitlines()[-1])
    print("%-22s R=%-5s value %.3e ms/step %.3f  e2e sweeps ms/step %.3f" % ("$name", "$R", d["value"], d["ms_per_step"], 1e3*d["e2e"]["breakdown_s"]["sweeps"]/d["steps"]))
except Exception as e:
    print("$name failed", e, open("gpurun_out/m_$name.log").read()[-300:])
PY
}
export PIQMC_LEVEL=1
run lvx_512 512 A=1
run lvx_512_cold 512 PIQMC_BENCH_TEMP=0.0001
run lvx_1024 1024 A=1
run lvx_2048 2048 A=1
run lvx_4096 4096 A=1
run lvx_64 64 A=1
