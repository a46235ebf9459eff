#!/bin/bash
# on-chip replay kernels (one warp per replica, replica in shared memory): parity, then config 2 timings
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1200 python -m pytest tests/test_gpu_det.py tests/test_gpu_reference_suite.py -q -x --timeout 600 > gpurun_out/t_det2.log 2>&1
echo "det tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/t_det2.log
echo "== on chip"; python tools/bench_det.py 1000 4096 2>&1 | tail -5
echo "== thread per replica"; PIQMC_DET_ONCHIP=0 python tools/bench_det.py 1000 2>&1 | tail -4
