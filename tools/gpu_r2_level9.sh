#!/bin/bash
export PIQMC_WATCHDOG_MS=8000
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_level.py -q -x --timeout 600 > gpurun_out/t_level.log 2>&1
echo "level tests rc=$?"; tail -12 gpurun_out/t_level.log
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
run lvt_4096 4096 A=1
run lvt_4096_cold 4096 PIQMC_BENCH_TEMP=0.0001
run lvt_2048 2048 A=1
run lvt_1024 1024 A=1
run lvt_512_k8 512 PIQMC_LEVEL_STAGED=0
run lvt_512_k4 512 PIQMC_LEVEL_STAGED=0 PIQMC_LEVEL_K=4
run lvt_8192 8192 A=1
