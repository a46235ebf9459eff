#!/bin/bash
for lpt in 8 16 32; do
PIQMC_ENERGY_LPT=$lpt python - <<'PY'
import os, sys, time
sys.path.insert(0, "pathintegral-qmc_b200")
import piqmc.tools as T
from piqmc import device
L = 256
nbs, _ = T.GaussianTorusNeighbors(L, 2024)
dev = device.Device(0)
dev.set_graph(nbs, T.TorusNaturalLevels(L))
for rows in (512, 4096):
    dev.state_alloc(rows, 64)
    dev.state_init_random(1, 0, tile=False)
    e0 = dev.energy(); dev.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dev.energy(download=False)
    dev.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("lanes per thread %s, %d rows: %.2f ms   checksum %.10e" % (os.environ["PIQMC_ENERGY_LPT"], rows, 1e3 * dt, float(e0.sum())))
PY
done
