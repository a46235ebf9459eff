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
run pf0_512 512 A=1
run pf256_512 512 PIQMC_PF_DIST=256
run pf512_512 512 PIQMC_PF_DIST=512
run pf1024_512 512 PIQMC_PF_DIST=1024
run pf2048_512 512 PIQMC_PF_DIST=2048
run pf0_1024 1024 A=1
run pf512_1024 1024 PIQMC_PF_DIST=512
run pf0_4096 4096 A=1
run pf512_4096 4096 PIQMC_PF_DIST=512
run pf1024_4096 4096 PIQMC_PF_DIST=1024
