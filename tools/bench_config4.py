#!/usr/bin/env python
"""Throughput of BASELINE.json configs[3] (the multispin-coded instances: bipartite8 with J, h = +-1 and
maxnb 5, hopfield8 with maxnb 8; examples/bipartite8.py:60, examples/hopfield8.py:22-33,98 of the
reference) on one B200: SA with 64 replicas per word (sa.Anneal_multispin's coding) and PIQMC with P = 10
slices, through the per-class generic kernel (variant 1) and the resident kernel (variant 5: the state of a
row in shared memory, the whole run one launch).  Prints attempts/s of the sweep phase.

    python tools/bench_config4.py [R]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pathintegral-qmc_b200"))
from piqmc import device  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
g4 = np.load(os.path.join(ROOT, "tests", "golden", "ref_config4.npz"))
vec = np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))
dev = device.default_device(0)
for inst in ("bipartite8", "hopfield8"):
    nbs = vec["nbs_" + inst]
    n = nbs.shape[0]
    sched, mcsteps = g4["sched_" + inst], int(g4["mcsteps_" + inst])
    color = np.arange(n, dtype=np.int32) if inst == "hopfield8" else None
    import piqmc.tools as tools
    if color is None:
        color = tools.ColourGraph(nbs, "natural")
    for kind in ("sa", "qa"):
        for variant in (1, 5):
            dev.set_graph(nbs, color)
            dev.set_variant(variant)
            try:
                if kind == "sa":
                    rows, P = (R + 63) // 64, 64
                    dev.state_alloc(rows, 64)
                    run = lambda: dev.sa_colour(sched, mcsteps, 11)
                    attempts = float(rows) * 64 * n * sched.size * mcsteps
                else:
                    rows, P = R, 10
                    dev.state_alloc(rows, P)
                    qsched = np.linspace(3.0, 1e-8, sched.size)
                    run = lambda: dev.qa_colour(qsched, mcsteps, 0.05, 11)
                    attempts = float(rows) * P * n * sched.size * mcsteps
                dev.state_init_random(11, 0, tile=(kind == "qa"))
                run()
                dev.synchronize()
                dev.state_init_random(11, 0, tile=(kind == "qa"))
                dev.synchronize()
                t0 = time.perf_counter()
                run()
                dev.synchronize()
                dt = time.perf_counter() - t0
            finally:
                dev.set_variant(0)
            print("%-10s %s  R=%d (%d rows)  %d schedule steps x %d sweeps  variant %d  %.4f s -> %.3e attempts/s"
                  % (inst, kind.upper(), R, rows, sched.size, mcsteps, variant, dt, attempts / dt))

# ---- the +-J instance (bipartite8) in the regime multispin coding is made for: long, cold runs (500 sweeps;
#      SA T 0.3 -> 0.01, PIQMC T = 0.05 with P = 64 slices), per-lane resident kernel vs bit-sliced resident kernel
nbs = vec["nbs_bipartite8"]
color = tools.ColourGraph(nbs, "natural")
for kind in ("sa", "qa"):
    for env, label in (({"PIQMC_NO_INT_KERNEL": "1"}, "per-lane"), ({"PIQMC_FORCE_INT_KERNEL": "1"}, "bit-sliced")):
        os.environ.update(env)
        try:
            dev.set_graph(nbs, color)
            dev.set_variant(5)
            if kind == "sa":
                rows, P = (R + 63) // 64, 64
                dev.state_alloc(rows, 64)
                ssched = np.linspace(0.3, 0.01, 5)
                run = lambda: dev.sa_colour(ssched, 100, 11)
            else:
                rows, P = R // 8, 64
                dev.state_alloc(rows, P)
                qsched = np.linspace(3.0, 1e-8, 5)
                run = lambda: dev.qa_colour(qsched, 100, 0.05, 11)
            attempts = float(rows) * P * 8 * 500
            dev.state_init_random(11, 0, tile=(kind == "qa"))
            run()
            dev.synchronize()
            dev.state_init_random(11, 0, tile=(kind == "qa"))
            dev.synchronize()
            t0 = time.perf_counter()
            run()
            dev.synchronize()
            dt = time.perf_counter() - t0
        finally:
            dev.set_variant(0)
            for k in env:
                del os.environ[k]
        print("bipartite8 %s cold, 500 sweeps, %d rows x %d lanes, %-10s kernel: %.4f s -> %.3e attempts/s"
              % (kind.upper(), rows, P, label, dt, attempts / dt))
