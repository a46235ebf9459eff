#!/usr/bin/env python
"""Throughput of the production SA path (sa.Anneal rules, 64 replicas per word) on the 256x256
Gaussian torus: R = 65536 replicas (1024 rows), natural-order level colouring, SWEEPS (default 100) sweeps per
schedule.  Hot schedules draw a uniform for about half of all attempts, cold ones for almost none.

    python tools/bench_sa.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pathintegral-qmc_b200"))
import piqmc.tools as tools  # noqa: E402
from piqmc import device  # noqa: E402

L, ROWS = 256, 1024
SWEEPS = int(os.environ.get("SWEEPS", "100"))
nbs, _ = tools.GaussianTorusNeighbors(L, 2024)
dev = device.Device(0)
dev.set_graph(nbs, tools.TorusNaturalLevels(L))
dev.state_alloc(ROWS, 64)
for name, lo, hi in (("hot 3.0->1.0", 3.0, 1.0), ("cold 0.3->0.01", 0.3, 0.01), ("frozen 0.01", 0.01, 0.01),
                     ("full 3.0->0.01", 3.0, 0.01)):
    sched = np.linspace(lo, hi, SWEEPS)
    dev.state_init_random(1, 0, tile=False)
    dev.sa_colour(sched[:2], 1, 1)                     # warm-up
    dev.state_init_random(1, 0, tile=False)
    dev.synchronize()
    t0 = time.perf_counter()
    dev.sa_colour(sched, 1, 1)
    dev.synchronize()
    dt = time.perf_counter() - t0
    print("SA fast kernel, %dx%d, R=%d (%d rows), %s: %.3f s -> %.3e attempts/s"
          % (L, L, 64 * ROWS, ROWS, name, dt, 64.0 * ROWS * L * L * SWEEPS / dt))
