#!/bin/bash
export PIQMC_WATCHDOG_MS=5000
mkdir -p gpurun_out
timeout 300 python tools/debug_chain.py > gpurun_out/debug_chain.log 2>&1
cut -c1-200 gpurun_out/debug_chain.log
