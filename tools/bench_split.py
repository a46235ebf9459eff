#!/usr/bin/env python
"""Where along the anneal does which kernel win?  BASELINE configs[4] at R rows on one GPU, K schedule steps:
the first s steps through the dataflow kernel (variant 2), the rest through the chain pipeline (variant 3), each
part timed on its own (both are the same sequential sweep, bit for bit, so they can be mixed freely).

    python tools/bench_split.py [rows] [K]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pathintegral-qmc_b200"))
import piqmc.tools as tools  # noqa: E402
from piqmc import device  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512
K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
L, P, T = 256, 64, 0.01
nbs, _ = tools.GaussianTorusNeighbors(L, 2024)
color = tools.TorusNaturalLevels(L)
dev = device.default_device(0)
dev.set_graph(nbs, color)
dev.state_alloc(rows, P)
sched = np.linspace(1.5, 1e-8, K)


def part(variant, a, b):
    if a >= b:
        return 0.0
    dev.set_variant(variant)
    dev.synchronize()
    t0 = time.perf_counter()
    dev.qa_colour(sched[a:b], 1, T, 2024, sweep0=a)
    dev.synchronize()
    dev.set_variant(0)
    return time.perf_counter() - t0


for warm in range(2):
    for s in ([0, K] if warm == 0 else list(range(0, K + 1, max(1, K // 10)))):
        dev.state_init_random(2024, 0, tile=True)
        t1 = part(2, 0, s)
        t2 = part(3, s, K)
        if warm:
            print("rows %d: first %3d steps dataflow %.2f ms (%.3f/step), last %3d steps chain %.2f ms (%.3f/step): total %.2f ms = %.3f ms/step"
                  % (rows, s, 1e3 * t1, 1e3 * t1 / max(s, 1), K - s, 1e3 * t2, 1e3 * t2 / max(K - s, 1), 1e3 * (t1 + t2),
                     1e3 * (t1 + t2) / K))
