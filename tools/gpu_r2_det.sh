#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_det.py tests/test_gpu_reference_suite.py -q -x --timeout 600 > gpurun_out/t_det.log 2>&1
echo "det tests rc=$?"; tail -3 gpurun_out/t_det.log
python tools/bench_det.py 1000 16384 2>&1 | tail -5
