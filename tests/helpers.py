"""Shared helpers for the parity tests: golden-case loading and input reconstruction."""
import json

import numpy as np

NSPINS = {"boixo": 8, "boixo16": 16, "bipartite8": 8, "hopfield8": 8, "inst_0_32x32": 1024,
          "santoro_80x80": 6400}
GS_ENERGY = {"inst_0_32x32": -1591.9166416866, "santoro_80x80": -10115.3067314770}


def cases(vec, kinds=None):
    out = json.loads(str(vec["cases_json"]))
    if kinds is not None:
        out = [c for c in out if c["kind"] in kinds]
    return out


def case_inputs(case, vec):
    """(sched, nbs, rng positioned after the initial-state draws, initial state) for a golden case,
    drawn exactly as tests/golden/make_golden.py drew them."""
    n = NSPINS[case["inst"]]
    a, b, k = case["sched"]
    sched = np.linspace(a, b, int(k))
    nbs = vec["nbs_" + case["inst"]]
    rng = np.random.RandomState(case["rng_seed"])
    if case["kind"] == "multispin":
        init = np.array([[rng.randint(2) for _ in range(n)] for _ in range(64)], dtype=np.float64)
    else:
        init = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
    assert np.array_equal(init.astype(np.int8), vec[case["name"] + "__init"])
    return sched, nbs, rng, init


def ks_2samp_p(a, b):
    from scipy.stats import ks_2samp
    return ks_2samp(a, b).pvalue
