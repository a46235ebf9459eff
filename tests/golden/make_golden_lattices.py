#!/usr/bin/env python
"""Golden fixtures for tools.Generate2DLattice / tools.GenerateKblockLattice from the REFERENCE
ITSELF (piqmc/tools.pyx:132-272, compiled by oracle/build_ref.py).  Build container only.

    python tests/golden/make_golden_lattices.py   ->  tests/golden/ref_lattices.npz

Per case: the dense matrix, the DOK key order (which decides the row order of GenerateNeighbors),
and the next draw of the shared RandomState (the generators must consume exactly the same stream).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

CASES = [("lattice", 4, 5, 0), ("lattice", 4, 5, 1), ("lattice", 3, 3, 1), ("lattice", 1, 6, 1),
         ("lattice", 6, 1, 0), ("lattice", 2, 2, 1), ("kblock", 5, 6, 1), ("kblock", 5, 6, 2),
         ("kblock", 3, 3, 3), ("kblock", 4, 7, 2), ("kblock", 1, 5, 2), ("kblock", 7, 7, 3)]


def main():
    assert O.ref() is not None, "oracle/_ref not built (python oracle/build_ref.py)"
    from piqmc_ref import tools
    out, meta = {}, []
    for k, (kind, a, b, c) in enumerate(CASES):
        rng = np.random.RandomState(100 + k)
        J = tools.Generate2DLattice(a, b, rng, c) if kind == "lattice" else tools.GenerateKblockLattice(a, b, rng, c)
        out["J%d" % k] = J.toarray()
        out["keys%d" % k] = np.array(list(J.keys()), dtype=np.int64).reshape(-1, 2)
        out["next%d" % k] = np.array([rng.randint(1 << 30)], dtype=np.int64)
        meta.append(dict(kind=kind, nrows=a, ncols=b, arg=c, seed=100 + k))
    out["cases_json"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "ref_lattices.npz"), **out)
    print("wrote ref_lattices.npz: %d cases" % len(meta))


if __name__ == "__main__":
    main()
