#!/bin/bash
export PIQMC_BENCH_ORDERS=natural
for R in 8192 32768; do
echo "== flow auto R=$R"; python tools/bench_configs.py $R 2>&1 | grep attempts
echo "== level per_word=3 R=$R"; PIQMC_LEVEL=1 PIQMC_BENCH_PER_WORD=3 python tools/bench_configs.py $R 2>&1 | grep attempts
echo "== level per_word=1 R=$R"; PIQMC_LEVEL=1 PIQMC_BENCH_PER_WORD=1 python tools/bench_configs.py $R 2>&1 | grep attempts
echo "== level staged forced per_word=3 R=$R"; PIQMC_LEVEL=1 PIQMC_LEVEL_STAGED=2 PIQMC_BENCH_PER_WORD=3 python tools/bench_configs.py $R 2>&1 | grep attempts
done
