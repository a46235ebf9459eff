#!/usr/bin/env python
"""The deterministic (bit-exact replay) path on one B200: one call of the drop-in qmc.QuantumAnneal on
BASELINE.json configs[1] (inst_0_32x32, P = 20, 100 schedule steps: 2.05e6 attempts), next to the same call of
the reference's compiled Cython on a host core when oracle/_ref is present, and the batched entry
qmc.QuantumAnnealBatch (R independent replays in one launch, one CUDA thread each).

    python tools/bench_det.py [R ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pathintegral-qmc_b200"))
import piqmc.qmc as qmc  # noqa: E402

vec = np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))
nbs = vec["nbs_inst_0_32x32"]
n, P, T = 1024, 20, 0.01
sched = np.linspace(1.5, 1e-8, 100)
attempts = float(n) * P * sched.size


def start(r):
    rng = np.random.RandomState(r)
    sv = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
    return np.tile(sv, (P, 1)).T.copy(), rng


for rep in range(2):                                             # the first call pays the library start-up
    confs, rng = start(0)
    t0 = time.perf_counter()
    qmc.QuantumAnneal(sched, 1, P, T, n, confs, nbs, rng)
    dt = time.perf_counter() - t0
print("qmc.QuantumAnneal (GPU drop-in, one call): %.3f s -> %.3e attempts/s" % (dt, attempts / dt))
try:
    from oracle import oracle as O
    if O.ref() is not None:
        from piqmc_ref import qmc as rqmc
        confs, rng = start(0)
        t0 = time.perf_counter()
        rqmc.QuantumAnneal(sched, 1, P, T, n, confs, nbs, rng)
        dt = time.perf_counter() - t0
        print("qmc.QuantumAnneal (reference Cython, one host core): %.3f s -> %.3e attempts/s" % (dt, attempts / dt))
except Exception as e:                                           # the compiled reference does not travel everywhere
    print("reference not available here:", e)
for R in [int(a) for a in sys.argv[1:]] or [1000, 16384]:
    inits, rngs = [], []
    for r in range(R):
        c, g = start(r % 64)
        inits.append(c)
        rngs.append(g)
    from piqmc import device
    d = device.default_device()
    d.set_graph(nbs)
    spins = np.ascontiguousarray(np.array(inits), dtype=np.int8)
    t0 = time.perf_counter()
    perms = np.stack([qmc._draw_perms(rng, n, sched.size) for rng in rngs])     # MT19937 on the host, as the reference
    th = time.perf_counter() - t0
    st = device.rand_states(list(range(R)))
    t0 = time.perf_counter()
    d.qa_det(sched, 1, P, T, spins, perms, rstates=st)
    dt = time.perf_counter() - t0
    print("qmc.QuantumAnnealBatch R=%d: device call %.3f s -> %.3e attempts/s (host: %.2f s drawing the %d x %d permutations)"
          % (R, dt, attempts * R / dt, th, R, sched.size + 1))
