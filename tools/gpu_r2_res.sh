#!/bin/bash
mkdir -p gpurun_out
python tools/bench_config4.py 65536 2>&1 | tail -8
python tools/bench_config4.py 4194304 2>&1 | grep "variant 5"
