#!/bin/bash
# launch knobs again after the instruction count fell to 321 per 32 words (issue slots 67 %)
run() { name=$1; rep=$2; shift 2; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu --replicas $rep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name rows $rep: value %.3e ms/sweep %.3f' % (d['value'], d['ms_per_step']))"; }
run default 4096 X=1
run minb8 4096 PIQMC_MINB=8
run minb10 4096 PIQMC_MINB=10
run rpb256 4096 PIQMC_ROWS_PER_BLOCK=256
run rpb1024 4096 PIQMC_ROWS_PER_BLOCK=1024
