#!/bin/bash
# needs level_kernels.o built with -DLV_PROFILE
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=8000
timeout 1500 python -m pytest tests/test_gpu_level.py -q -x --timeout 600 > gpurun_out/t_level.log 2>&1
echo "level tests rc=$?"; tail -8 gpurun_out/t_level.log
export PIQMC_LEVEL=1 PIQMC_LEVEL_PROF=1
for cfg in "512 A=1" "512 PIQMC_BENCH_TEMP=0.0001" "64 A=1" "4096 A=1"; do
  set -- $cfg; R=$1; shift
  echo "=== R=$R $@"
  env "$@" timeout 300 python bench.py --steps 20 --warmup 1 --no-cpu --replicas $R > gpurun_out/p.log 2>&1; grep "level prof" gpurun_out/p.log | tail -8
done
