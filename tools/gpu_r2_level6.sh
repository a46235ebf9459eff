#!/bin/bash
# needs level_kernels.o built with -DLV_PROFILE
mkdir -p gpurun_out
export PIQMC_LEVEL=1 PIQMC_LEVEL_PROF=1
for cfg in "512 A=1" "512 PIQMC_LEVEL_DRY=2" "512 PIQMC_BENCH_TEMP=0.0001" "512 PIQMC_LEVEL_K=16 PIQMC_LEVEL_WARPS=16" "4096 A=1"; do
  set -- $cfg; R=$1; shift
  echo "=== R=$R $@"
  env "$@" timeout 300 python bench.py --steps 20 --warmup 1 --no-cpu --replicas $R > gpurun_out/p.log 2>&1; grep "level prof" gpurun_out/p.log | tail -9
done
