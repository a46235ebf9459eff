#!/bin/bash
export PIQMC_WATCHDOG_MS=8000
mkdir -p gpurun_out
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
run dry_k8_s0 512 PIQMC_LEVEL_DRY=1 PIQMC_LEVEL_SYNC=0
run dry_k8_s1 512 PIQMC_LEVEL_DRY=1 PIQMC_LEVEL_SYNC=1
run dry_k8_s2 512 PIQMC_LEVEL_DRY=1 PIQMC_LEVEL_SYNC=2
run dry_k1 4096 PIQMC_LEVEL_DRY=1
run dry_k8_s1_w8 512 PIQMC_LEVEL_DRY=1 PIQMC_LEVEL_SYNC=1 PIQMC_LEVEL_WARPS=8
run lv_512_s1 512 PIQMC_LEVEL_SYNC=1
run lv_512_s2 512 PIQMC_LEVEL_SYNC=2
run lv_512_s1_k4 512 PIQMC_LEVEL_SYNC=1 PIQMC_LEVEL_K=4
run lv_512_s1_k16w16 512 PIQMC_LEVEL_SYNC=1 PIQMC_LEVEL_K=16 PIQMC_LEVEL_WARPS=16
