#!/bin/bash
# hot SA: transposed pattern lookup in the per-thread draw path
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
timeout 900 python -m pytest tests/test_gpu_colour.py tests/test_gpu_chain.py tests/test_gpu_level.py -q -x -m gpu --timeout 600 -k "sa" > gpurun_out/t_sa.log 2>&1
echo "SA tests rc=$?"; tail -2 gpurun_out/t_sa.log
python tools/bench_sa.py 2>&1 | tail -4
