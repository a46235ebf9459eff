#!/usr/bin/env python
"""Production throughput on DENSE instances (the graphs of qmc.QuantumAnneal_dense / sa.Anneal_dense,
piqmc/qmc.pyx:141-242, piqmc/sa.pyx:126-187): fully connected K_N with Gaussian couplings and fields, whose
neighbour table is the dense row (maxnb = N).  These run through the resident kernel (the state of a replica row
in shared memory, the whole run one launch, sequential visiting order = the reference's), SA with 64 replicas
per word and PIQMC with P slices; the bit-exact replays of the `_dense` signatures are separate
(piqmc_qa_dense_det / piqmc_sa_dense_det).  Prints attempts/s of the sweep phase and, for one small case, checks
the device energies against sa.ClassicalIsingEnergy restated on the host.

    python tools/bench_dense.py [R]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pathintegral-qmc_b200"))
import piqmc.tools as tools  # noqa: E402
from piqmc import device  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = device.default_device(0)
for n in (16, 32, 64, 128):
    rng = np.random.RandomState(n)
    J = np.triu(rng.randn(n, n))                                    # upper triangle + diagonal (fields), as the _dense variants read it
    nbs = np.zeros((n, n, 2))
    for i in range(n):                                              # dense row: the diagonal entry first, then every other spin
        cols = [i] + [j for j in range(n) if j != i]
        nbs[i, :, 0] = cols
        nbs[i, :, 1] = [J[min(i, j), max(i, j)] for j in cols]
    color = np.arange(n, dtype=np.int32)                            # K_N: every spin its own class = the sequential sweep
    dev.set_graph(nbs, color)
    for kind in ("sa", "qa"):
        sweeps = 50
        if kind == "sa":
            rows, P = (R + 63) // 64, 64
            dev.state_alloc(rows, 64)
            sched = np.linspace(3.0 * np.sqrt(n), 0.05, sweeps)
            run = lambda: dev.sa_colour(sched, 1, 7)
        else:
            rows, P = max(64, R // 16), 16
            dev.state_alloc(rows, P)
            sched = np.linspace(3.0, 1e-8, sweeps)
            run = lambda: dev.qa_colour(sched, 1, 0.05 * np.sqrt(n), 7)
        attempts = float(rows) * P * n * sweeps
        dev.state_init_random(7, 0, tile=(kind == "qa"))
        run()
        dev.synchronize()
        dev.state_init_random(7, 0, tile=(kind == "qa"))
        dev.synchronize()
        t0 = time.perf_counter()
        run()
        dev.synchronize()
        dt = time.perf_counter() - t0
        en = dev.energy()
        w = dev.state_download_words()
        s = 1.0 - 2.0 * ((w[0] >> np.uint64(0)) & np.uint64(1)).astype(np.float64)   # row 0, lane 0
        ref = -(s @ (np.triu(J, 1) @ s)) - np.dot(np.diag(J), s)
        ok = abs(en[0, 0] - ref) <= 1e-9 * max(1.0, abs(ref))
        print("K_%d %s  %d rows x %d lanes, %d sweeps: %.4f s -> %.3e attempts/s (%.1f coupling terms per attempt; "
              "energy check %s, mean E/N %.3f)" % (n, kind.upper(), rows, P, sweeps, dt, attempts / dt, n, "ok" if ok else "FAILED",
                                                   en.mean() / n))
