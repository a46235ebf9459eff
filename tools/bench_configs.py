#!/usr/bin/env python
"""Throughput of the production PIQMC path on the reference's own configurations (BASELINE.json
configs[1] and configs[2]): inst_0_32x32 and santoro_80x80, P = 20 slices, T = 0.01,
Gamma 1.5 -> 1e-8 in 100 steps, R replicas at once; natural order (the reference's statistics) and
checkerboard order.  Prints attempts/s and the residual energy per spin.

    python tools/bench_configs.py [R]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pathintegral-qmc_b200"))
sys.path.insert(0, os.path.join(ROOT, "examples"))
import _instances  # noqa: E402
import piqmc.qmc as qmc  # noqa: E402
import piqmc.tools as tools  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
PER_WORD = os.environ.get("PIQMC_BENCH_PER_WORD", "auto")       # experiments: force the replicas per word
PER_WORD = PER_WORD if PER_WORD == "auto" else int(PER_WORD)
ORDERS = os.environ.get("PIQMC_BENCH_ORDERS", "natural,checkerboard").split(",")
P, T, steps = 20, 0.01, 100
sched = np.linspace(1.5, 1e-8, steps)
for name, n in (("inst_0_32x32", 1024), ("santoro_80x80", 6400)):
    J = _instances.load(name, n)
    _, gs = _instances.ground_state(name)
    nbs = tools.GenerateNeighbors(n, J, 4)
    for order in ORDERS:
        qmc.QuantumAnnealReplicas(sched[:2], 1, P, T, n, None, nbs, 1, order=order, nreplicas=R, per_word=PER_WORD)   # warm-up
        t0 = time.perf_counter()
        out = qmc.QuantumAnnealReplicas(sched, 1, P, T, n, None, nbs, 1, order=order, nreplicas=R, per_word=PER_WORD)
        dt = out["seconds"]["sweeps"]
        en = out["energies"]
        print("%-14s P=%d R=%d per_word=%d %-12s sweeps %.4f s -> %.3e attempts/s   residual/spin %.4f (wall %.3f s)"
              % (name, P, R, out["per_word"], order, dt, float(R) * P * n * steps / dt, (en.mean() - gs) / n,
                 time.perf_counter() - t0))
