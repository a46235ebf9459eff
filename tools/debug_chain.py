"""debug: where does the chain kernel differ from the oracle? (GPU)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pathintegral-qmc_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O
import piqmc.tools as tools
from piqmc import device

def run(L, P, R, nsw, T, env, seed=5):
    for k, v in env.items():
        os.environ[k] = v
    nbs, _ = tools.GaussianTorusNeighbors(L, seed)
    idx, J32 = O.nbs_to_ell(nbs)
    color = tools.TorusNaturalLevels(L)
    n = L * L
    sched = np.linspace(1.5, 1e-8, nsw)
    init = O.colour_init_spins(seed, 0, R, n)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    O.qa_colour(sched, 1, P, T, idx, J32, color, want, seed, 0, 0, 0)
    dev = device.default_device(0)
    dev.set_graph(nbs, color)
    dev.set_variant(3)
    dev.state_alloc(R, P)
    print("chain_info", dev.chain_info(), "L", L, "P", P, "R", R, "sweeps", nsw, env)
    dev.state_init_random(seed, 0, tile=True)
    try:
        dev.qa_colour(sched, 1, T, seed, replica0=0, sweep0=0)
    except Exception as e:
        print("ERROR", e)
        return
    got = np.transpose(tools.UnpackWords(dev.state_download_words(), P), (0, 2, 1))
    bad = np.argwhere((got != want).any(axis=2))
    print("  mismatching (replica, spin):", len(bad), "of", R * n, " changed from init:", int((got[:, :, 0] != init).sum()))
    if len(bad):
        sp = np.unique(bad[:, 1])
        print("  first spins", sp[:40], " rows(y) hist", np.bincount(sp // L, minlength=L)[:L], " cols(x) hist", np.bincount(sp % L, minlength=L)[:L])
        print("  replicas", np.unique(bad[:, 0])[:20])
        badset = set(map(tuple, bad))
        shown = 0
        for r, i in bad:
            lower = [j for j in ((i - 1) if i % L else None, i - L if i >= L else None) if j is not None]
            if any((r, j) in badset for j in lower):
                continue
            d = (got[r, i] != want[r, i])
            print("    root: replica %d spin %d (y %d x %d): differing slices %s  want %s got %s" % (
                r, i, i // L, i % L, np.flatnonzero(d)[:12], want[r, i][:8], got[r, i][:8]))
            shown += 1
            if shown >= 12:
                break
    for k in env:
        del os.environ[k]

E = {"PIQMC_CHAIN_CW": "16", "PIQMC_CHAIN_BANDS": "1"}
run(16, 64, 45, 1, 0.01, E)
run(16, 64, 64, 1, 0.01, E)
run(16, 64, 64, 1, 0.01, dict(E, PIQMC_CHAIN_RPT="1"))
run(16, 20, 64, 1, 0.01, E)
