#!/bin/bash
# end-of-round validation on one B200: all GPU tests, smoke, the bench line (both arms), launch list + full capture
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
t0=$(date +%s)
timeout 3000 python -m pytest tests -q -x -m gpu --timeout 900 > gpurun_out/t_all_r2.log 2>&1
echo "all gpu tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -4 gpurun_out/t_all_r2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 400 gpurun_out/r2_bench_n1.json; echo
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; tail -c 300 gpurun_out/r2_bench_reference_arm.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps10.csv python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_launches.log 2>&1
echo "launch list rc=$?"; grep -c "colour_sweep_fast" gpurun_out/r2_launches_bench_steps10.csv
timeout 900 ncu --set full --import-source on --clock-control none -k regex:colour_sweep_fast -s 1 -c 1 -f -o gpurun_out/prof_r2_final python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/prof_r2_final.log 2>&1
echo "full capture rc=$?"
echo "elapsed $(( $(date +%s) - t0 )) s"
