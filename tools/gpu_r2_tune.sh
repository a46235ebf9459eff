#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; R=$2; shift 2
  env "$@" timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --replicas $R > gpurun_out/m_$name.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/m_$name.log").read().strip().splitlines()[-1])
    print("%-22s R=%-5s value %.3e ms/step %.3f" % ("$name", "$R", d["value"], d["ms_per_step"]))
except Exception as e:
    print("$name failed", e, open("gpurun_out/m_$name.log").read()[-300:])
PY
}
run base 4096 A=1
run rpb1024 4096 PIQMC_ROWS_PER_BLOCK=1024
run rpb2048 4096 PIQMC_ROWS_PER_BLOCK=2048
run rpb256 4096 PIQMC_ROWS_PER_BLOCK=256
run minb10 4096 PIQMC_MINB=10
run minb8 4096 PIQMC_MINB=8
run rpb1024_minb10 4096 PIQMC_ROWS_PER_BLOCK=1024 PIQMC_MINB=10
run b512_rpb256 512 PIQMC_ROWS_PER_BLOCK=256
run b512_minb10 512 PIQMC_MINB=10
