#!/bin/bash
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=60000
timeout 900 ncu --set full --import-source on --clock-control none -k regex:colour_sweep_fast -s 1 -c 1 -f -o gpurun_out/prof_r2_final_512 python bench.py --steps 10 --warmup 3 --no-cpu --replicas 512 > gpurun_out/prof_r2_final_512.log 2>&1
echo rc=$?
