#!/bin/bash
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
run dry1_k8 512 PIQMC_LEVEL_DRY=1
run dry2_k8 512 PIQMC_LEVEL_DRY=2
run dry3_k8 512 PIQMC_LEVEL_DRY=3
run dry3_k8_w8 512 PIQMC_LEVEL_DRY=3 PIQMC_LEVEL_WARPS=8
run dry3_k1_512 512 PIQMC_LEVEL_DRY=3 PIQMC_LEVEL_K=1
run dry3_k1_4096 4096 PIQMC_LEVEL_DRY=3
run dry2_k1_4096 4096 PIQMC_LEVEL_DRY=2
