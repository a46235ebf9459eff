#!/bin/bash
# round-2 profiles: launch list and full capture of the bench kernel, full captures of the resident kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps10.csv python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:colour_sweep_fast -s 1 -c 1 -f -o gpurun_out/prof_r2_flow_4096 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/ncu_flow.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:resident_sweeps_int -s 2 -c 1 -f -o gpurun_out/prof_r2_resident_int python tools/bench_config4.py 4194304 > gpurun_out/ncu_res_int.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:resident_sweeps -s 5 -c 1 -f -o gpurun_out/prof_r2_resident python tools/bench_config4.py 4194304 > gpurun_out/ncu_res.log 2>&1
tail -2 gpurun_out/ncu_res.log; ls -la gpurun_out/*.ncu-rep
