#!/bin/bash
# back-off between dependency polls (PIQMC_POLL_NS) in the issue-bound regime: the polling warp executes ~14
# instructions per poll, 2.9 polls per word pass at 4096 rows
for ns in 0 100 250 500 1000; do
  for rep in 4096 512; do
    PIQMC_POLL_NS=$ns python bench.py --steps 50 --warmup 3 --no-cpu --replicas $rep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('poll_ns $ns rows $rep: value %.3e ms/sweep %.3f' % (d['value'], d['ms_per_step']))"
  done
done
