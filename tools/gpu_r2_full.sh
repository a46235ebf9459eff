#!/bin/bash
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
timeout 3000 python -m pytest tests -q -x -m gpu --timeout 900 > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$?"; tail -6 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1_steps20.json 2> gpurun_out/bench_n1_steps20.err; tail -c 600 gpurun_out/bench_n1_steps20.json
