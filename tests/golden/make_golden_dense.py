#!/usr/bin/env python
"""Golden fixtures for the dense variants, from the REFERENCE ITSELF: qmc.QuantumAnneal_dense
(piqmc/qmc.pyx:141-242) and sa.Anneal_dense (piqmc/sa.pyx:126-187), as compiled by
oracle/build_ref.py into oracle/_ref/piqmc_ref.  Build container only; the output
tests/golden/ref_dense.npz is committed and is all the tests read.

    python tests/golden/make_golden_dense.py

Seeding as in make_golden.py: np.random.RandomState(rng_seed) for the initial state and the spin
orders, libc srand(srand_seed) for the Metropolis uniforms.  The dense matrices are the upper
triangles (+ diagonal = fields) of the golden instances, plus one full random K12 matrix whose
lower triangle holds different values (the reference never reads it, qmc.pyx:203-211).
"""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

libc = ctypes.CDLL("libc.so.6")


def dense_from_triples(ijv, n):
    J = np.zeros((n, n))
    for i, j, v in ijv:
        i, j = int(i) - 1, int(j) - 1
        J[min(i, j), max(i, j)] = v
    return J


def main():
    assert O.ref() is not None, "oracle/_ref not built (python oracle/build_ref.py)"
    from piqmc_ref import qmc, sa
    inst = np.load(os.path.join(HERE, "instances.npz"))
    mats = {}
    for name, n in (("boixo", 8), ("bipartite8", 8), ("hopfield8", 8), ("boixo16", 16)):
        mats[name] = dense_from_triples(inst["inst_" + name], n)
    r = np.random.RandomState(77)
    K = r.uniform(-1, 1, size=(12, 12))                 # lower triangle != upper: never read
    mats["k12"] = K
    sub = dense_from_triples(inst["inst_inst_0_32x32"], 1024)[:48, :48].copy()
    mats["torus48"] = sub                               # sparse-in-dense, 48 spins

    cases, out = [], {}
    for name, J in mats.items():
        out["J_" + name] = J
    qa = [("boixo", 5, 0.01, (0.5, 1e-8, 10), 3, 1), ("boixo", 20, 0.01, (0.5, 1e-8, 10), 1, 2),
          ("hopfield8", 10, 0.01, (8.0, 1e-8, 5), 20, 3), ("bipartite8", 10, 0.3, (3.0, 1e-8, 6), 4, 4),
          ("boixo16", 8, 0.05, (1.0, 1e-8, 12), 2, 5), ("k12", 6, 0.5, (2.0, 1e-3, 8), 3, 6),
          ("torus48", 12, 0.2, (1.5, 1e-8, 6), 2, 7)]
    for inst_name, P, T, sch, mcsteps, seed in qa:
        J = mats[inst_name]
        n = J.shape[0]
        rng = np.random.RandomState(seed)
        sv = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
        confs = np.tile(sv, (P, 1)).T.copy()
        libc.srand(seed + 100)
        sched = np.linspace(sch[0], sch[1], int(sch[2]))
        qmc.QuantumAnneal_dense(sched, mcsteps, P, T, n, confs, J, rng)
        name = "qad_%s_P%d_s%d" % (inst_name, P, seed)
        out[name + "__init"] = sv.astype(np.int8)
        out[name + "__final"] = confs.astype(np.int8)
        out[name + "__libc_next"] = np.array([libc.rand() for _ in range(4)], dtype=np.int64)
        out[name + "__rng_next"] = rng.randint(1 << 30, size=4).astype(np.int64)
        cases.append(dict(name=name, kind="qa_dense", inst=inst_name, P=P, T=T, sched=list(sch),
                          mcsteps=mcsteps, rng_seed=seed, srand_seed=seed + 100))
    sa_cases = [("boixo", (3.0, 0.01, 20), 2, 11), ("hopfield8", (8.0, 0.01, 10), 5, 12),
                ("bipartite8", (3.0, 0.01, 10), 3, 13), ("k12", (2.0, 0.05, 15), 2, 14),
                ("torus48", (3.0, 0.01, 12), 1, 15)]
    for inst_name, sch, mcsteps, seed in sa_cases:
        J = mats[inst_name]
        n = J.shape[0]
        rng = np.random.RandomState(seed)
        sv = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
        init = sv.copy()
        libc.srand(seed + 100)
        sched = np.linspace(sch[0], sch[1], int(sch[2]))
        sa.Anneal_dense(sched, mcsteps, sv, J, rng)
        name = "sad_%s_s%d" % (inst_name, seed)
        out[name + "__init"] = init.astype(np.int8)
        out[name + "__final"] = sv.astype(np.int8)
        out[name + "__libc_next"] = np.array([libc.rand() for _ in range(4)], dtype=np.int64)
        out[name + "__rng_next"] = rng.randint(1 << 30, size=4).astype(np.int64)
        cases.append(dict(name=name, kind="sa_dense", inst=inst_name, sched=list(sch), mcsteps=mcsteps,
                          rng_seed=seed, srand_seed=seed + 100))
    out["cases_json"] = np.array(json.dumps(cases))
    np.savez_compressed(os.path.join(HERE, "ref_dense.npz"), **out)
    print("wrote ref_dense.npz: %d cases" % len(cases))


if __name__ == "__main__":
    main()
