#!/bin/bash
# tables from their own kernel, thresholds staged in shared memory: rates + launch list (kernel durations)
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
timeout 600 python -m pytest tests/test_gpu_colour.py -q -x -m gpu --timeout 600 -k "many_rows or config5_shard or staggered or per_word" > gpurun_out/t_tabs2.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/t_tabs2.log
run() {
  name=$1; rep=$2; steps=$3; shift 3
  env "$@" timeout 300 python bench.py --steps $steps --warmup 3 --no-cpu --replicas $rep > gpurun_out/tabs2_$name.json 2> gpurun_out/tabs2_$name.err
  python -c "
import json
d = json.loads(open('gpurun_out/tabs2_$name.json').read().strip().splitlines()[-1])
print('$name: value %.3e ms/sweep %.3f e2e %.3e launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))"
}
run r4096_s20 4096 20 X=1
run r4096_s50 4096 50 X=1
run r4096_s50_minb8 4096 50 PIQMC_MINB=8
run r512_s50 512 50 X=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/tabs2_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/tabs2_ncu.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/tabs2_launches.csv')) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
for r in rows[1:]:
    if 'fast' in r[ki]:
        print(r[ki][:60], r[vi])
PY
