#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/en_run.py <<'PY'
import sys
sys.path.insert(0, "pathintegral-qmc_b200")
import piqmc.tools as T
from piqmc import device
nbs, _ = T.GaussianTorusNeighbors(256, 2024)
dev = device.Device(0)
dev.set_graph(nbs, T.TorusNaturalLevels(256))
dev.state_alloc(4096, 64)
dev.state_init_random(1, 0, tile=False)
dev.energy(download=False); dev.energy(download=False); dev.synchronize()
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:energy_partial -s 1 -c 1 -f -o gpurun_out/prof_r2_energy python /tmp/en_run.py > gpurun_out/prof_r2_energy.log 2>&1
echo rc=$?
ncu -i gpurun_out/prof_r2_energy.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
h, u, v = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active', 'dram__bytes_read.sum', 'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum']
for k, uu, vv in zip(h, u, v):
    if k in want or 'fp64' in k: print(k, vv, uu)
"
