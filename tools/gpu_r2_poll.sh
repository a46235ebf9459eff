#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
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
run poll0 4096 A=1
run poll50 4096 PIQMC_POLL_NS=50
run poll100 4096 PIQMC_POLL_NS=100
run poll200 4096 PIQMC_POLL_NS=200
run poll500 4096 PIQMC_POLL_NS=500
run poll1000 4096 PIQMC_POLL_NS=1000
run b512_poll100 512 PIQMC_POLL_NS=100
run b512_poll300 512 PIQMC_POLL_NS=300
