#!/bin/bash
# level-synchronous kernel: parity tests, then the bench shape at 4096 / 512 rows for several geometries
export PIQMC_WATCHDOG_MS=8000
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_level.py -q -x --timeout 600 > gpurun_out/t_level.log 2>&1
echo "level tests rc=$?"; tail -5 gpurun_out/t_level.log
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
run flow_512 512 A=1
run lv_512_k8 512 PIQMC_LEVEL=1
run lv_512_k4 512 PIQMC_LEVEL=1 PIQMC_LEVEL_K=4
run lv_512_k8w16 512 PIQMC_LEVEL=1 PIQMC_LEVEL_K=8 PIQMC_LEVEL_WARPS=16
run lv_512_k16w16 512 PIQMC_LEVEL=1 PIQMC_LEVEL_K=16 PIQMC_LEVEL_WARPS=16
run flow_4096 4096 A=1
run lv_4096 4096 PIQMC_LEVEL=1
run lv_4096_w16 4096 PIQMC_LEVEL=1 PIQMC_LEVEL_WARPS=16
run lv_4096_k2 4096 PIQMC_LEVEL=1 PIQMC_LEVEL_K=2
run lv_1024 1024 PIQMC_LEVEL=1
run flow_1024 1024 A=1
run lv_2048 2048 PIQMC_LEVEL=1
