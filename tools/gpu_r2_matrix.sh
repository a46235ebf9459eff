#!/bin/bash
# experiment matrix for the chain pipeline (one B200)
export PIQMC_WATCHDOG_MS=5000
mkdir -p gpurun_out
run() { # name, replicas, env...
  name=$1; R=$2; shift 2
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --replicas $R > gpurun_out/m_$name.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/m_$name.log").read().strip().splitlines()[-1])
    print("%-28s R=%-5s value %.3e ms/step %.3f" % ("$name", "$R", d["value"], d["ms_per_step"]))
except Exception as e:
    print("$name failed", e, open("gpurun_out/m_$name.log").read()[-300:])
PY
}
timeout 300 python tools/debug_chain.py > gpurun_out/debug_chain.log 2>&1; grep -c "mismatching (replica, spin): 0 of" gpurun_out/debug_chain.log; grep "mismatching" gpurun_out/debug_chain.log | grep -v ": 0 of" | head -3
run base512 512 A=1
run base512b 512 A=1
run bands16 512 PIQMC_CHAIN_BANDS=16
run bands17 512 PIQMC_CHAIN_BANDS=17
run bands20 512 PIQMC_CHAIN_BANDS=20
run bands32 512 PIQMC_CHAIN_BANDS=32
run rpt1 512 PIQMC_CHAIN_RPT=1
run rpt1b32 512 PIQMC_CHAIN_RPT=1 PIQMC_CHAIN_BANDS=32
run base4096 4096 A=1
run cw15_4096 4096 PIQMC_CHAIN_CW=15
run cw8_4096 4096 PIQMC_CHAIN_CW=8
run rpt1_4096 4096 PIQMC_CHAIN_RPT=1
timeout 1200 python -m pytest tests/test_gpu_chain.py -q --timeout 300 > gpurun_out/t_chain.log 2>&1
echo "chain tests rc=$?"; tail -8 gpurun_out/t_chain.log
