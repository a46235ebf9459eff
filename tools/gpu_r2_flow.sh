#!/bin/bash
export PIQMC_WATCHDOG_MS=8000
mkdir -p gpurun_out
PIQMC_FAST_RPT=2 timeout 1500 python -m pytest tests/test_gpu_colour.py tests/test_gpu_chain.py -q -x --timeout 600 > gpurun_out/t_colour_rpt2.log 2>&1
echo "colour tests (RPT=2 forced) rc=$?"; tail -4 gpurun_out/t_colour_rpt2.log
run() { name=$1; R=$2; shift 2
  env "$@" timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu --replicas $R > gpurun_out/m_$name.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/m_$name.log").read().strip().splitlines()[-1])
    print("%-22s R=%-5s value %.3e ms/step %.3f  e2e sweeps ms/step %.3f" % ("$name", "$R", d["value"], d["ms_per_step"], 1e3*d["e2e"]["breakdown_s"]["sweeps"]/d["steps"]))
except Exception as e:
    print("$name failed", e, open("gpurun_out/m_$name.log").read()[-300:])
PY
}
run rpt1_4096 4096 PIQMC_FAST_RPT=1
run rpt2_4096 4096 A=1
run rpt2_mb5 4096 PIQMC_MINB2=5
run rpt2_mb7 4096 PIQMC_MINB2=7
run rpt2_mb8 4096 PIQMC_MINB2=8
run rpt2_rpb1024 4096 PIQMC_ROWS_PER_BLOCK=1024
run rpt2_rpb256 4096 PIQMC_ROWS_PER_BLOCK=256
run rpt1_512 512 PIQMC_FAST_RPT=1
run rpt2_512 512 A=1
run rpt2_512_rpb256 512 PIQMC_ROWS_PER_BLOCK=256
run rpt2_512_mb8 512 PIQMC_MINB2=8
