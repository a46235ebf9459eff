#!/bin/bash
# Tuning sweep of the fast kernel's launch knobs on one B200 (run under gpurun):
#   PIQMC_MINB            resident blocks per SM the kernel is compiled for (7, 8, 9, 10)
#   PIQMC_ROWS_PER_BLOCK  rows per work unit (multiple of 128)
#   PIQMC_POLL_NS         back-off between polls of a completion flag
# usage: bash tools/sweep.sh            prints attempts/s per setting at 4096 and 512 rows
run() { echo "$@"; env "$@" timeout 120 python bench.py --no-cpu --steps 50 $EXTRA 2>&1 | grep -o '"value": [0-9.e+]*' | head -1; }
for R in 4096 512; do
    EXTRA="--replicas $R"
    echo "== $R rows"
    for m in 8 9 10; do run PIQMC_MINB=$m; done
    for r in 128 256 512 1024; do run PIQMC_ROWS_PER_BLOCK=$r; done
done
