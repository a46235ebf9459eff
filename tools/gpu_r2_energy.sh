#!/bin/bash
# energy reduction with per-spin entry lists and two spins in flight: parity, then its time at the bench size
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
timeout 900 python -m pytest tests -q -x -m gpu --timeout 600 -k "energy or config5 or sharding or handover or histogram or santoro or boixo or reference_suite" > gpurun_out/t_energy.log 2>&1
echo "energy tests rc=$?"; tail -2 gpurun_out/t_energy.log
python - <<'PY'
import sys, time
sys.path.insert(0, "pathintegral-qmc_b200")
import numpy as np
import piqmc.tools as T
from piqmc import device
L = 256
nbs, _ = T.GaussianTorusNeighbors(L, 2024)
dev = device.Device(0)
dev.set_graph(nbs, T.TorusNaturalLevels(L))
for rows in (512, 4096):
    dev.state_alloc(rows, 64)
    dev.state_init_random(1, 0, tile=False)
    dev.energy(download=False); dev.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dev.energy(download=False)
    dev.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("energy reduction, 256x256, %d rows x 64 lanes: %.2f ms (%.0f GB/s of packed state, %.2e float64 adds/s)"
          % (rows, 1e3 * dt, rows * L * L * 8 / dt / 1e9, rows * 64.0 * L * L * 2 / dt))
PY
