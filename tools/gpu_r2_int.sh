#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_colour.py tests/test_gpu_det.py -q -x --timeout 900 -m gpu > gpurun_out/t_colour.log 2>&1
echo "colour+det tests rc=$?"; tail -8 gpurun_out/t_colour.log
python tools/bench_config4.py 65536 2>&1 | grep bipartite
python tools/bench_config4.py 4194304 2>&1 | grep "bipartite.*variant 5"
python tools/bench_det.py 1000 2>&1 | tail -4
