"""GPU parity tests of the level-synchronous kernel (csrc/level_kernels.cu): every group of 32 replica
rows walks the dependency levels of the visiting order on its own thread-block cluster -- in two variants:
the plain one (any colouring, any geometry) and the staged one (static colourings with one member per warp
and step and an even number of rows: exchange buffers in distributed shared memory, cp.async staging; the
default when it applies and its clusters fit the device) and the streamed one (static colourings, an even
number of rows, any number of members per warp and step: per-warp record streams, bulk copies (TMA) into
staging slots; the default for the rest).  All must equal,
bit for bit, the CPU statement of the sequential sweep (oracle/piqmc_oracle.c part 3; the order of the
reference's per-spin-reset variant, piqmc/qmc.pyx:320-357) -- for every geometry (1, 2, 4, 8 blocks per
cluster, few and many warps, several members per warp and step), QA and SA, one and several replicas per
word, static colourings (natural order, checkerboard) and per-sweep visiting orders (sa.Anneal's fresh
permutation per sweep, piqmc/sa.pyx:120), graphs that are not lattices, and BASELINE.json's full size.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
import piqmc.tools as tools
from helpers import NSPINS

pytestmark = pytest.mark.gpu


def _inst(golden, inst):
    nbs = golden["vec"]["nbs_" + inst]
    idx, J32 = O.nbs_to_ell(nbs)
    return nbs, idx, J32, tools.OrderLevels(nbs)               # level colouring of the natural order


def _torus(L, seed):
    nbs, checker = tools.GaussianTorusNeighbors(L, seed)
    idx, J32 = O.nbs_to_ell(nbs)
    return nbs, idx, J32, tools.TorusNaturalLevels(L), checker


class _Env:
    def __init__(self, env):
        self.env = env

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.env}
        os.environ.update(self.env)

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


def _run_qa(dev, nbs, color, sched, mcsteps, P, T, R, seed, replica0, sweep0, env, orders=None):
    dev.set_graph(nbs, color)
    dev.set_variant(4)
    try:
        with _Env(env):
            dev.state_alloc(R, P)
            dev.state_init_random(seed, replica0, tile=True)
            l0 = dev.launch_count
            dev.qa_colour(sched, mcsteps, T, seed, replica0=replica0, sweep0=sweep0, orders=orders)
            assert dev.launch_count - l0 == 2                  # decision tables + the sweeps: one launch each
            got = tools.UnpackWords(dev.state_download_words(), P)
    finally:
        dev.set_variant(0)
    return np.ascontiguousarray(np.transpose(got, (0, 2, 1)))


def K(k, w=None):
    """a forced geometry of the plain kernel (the staged and streamed variants are switched off)"""
    return dict({"PIQMC_LEVEL_K": str(k), "PIQMC_LEVEL_STAGED": "0", "PIQMC_LEVEL_STREAM": "0"},
                **({"PIQMC_LEVEL_WARPS": str(w)} if w else {}))


def KS(k):
    """the streamed variant with k blocks per cluster"""
    return {"PIQMC_LEVEL_K": str(k), "PIQMC_LEVEL_STAGED": "0"}


PLAIN = {"PIQMC_LEVEL_STAGED": "0", "PIQMC_LEVEL_STREAM": "0"}
STAGED = {"PIQMC_LEVEL_STAGED": "2"}

CASES = [
    # inst, P, T, sched, mcsteps, R, geometries
    ("inst_0_32x32", 20, 0.01, (1.5, 1e-8, 25), 1, 6, ({}, PLAIN, KS(1), KS(2), K(2), K(8, 4))),                 # config 2 shape
    ("inst_0_32x32", 64, 0.01, (1.5, 1e-8, 8), 1, 35, ({}, K(4, 8))),                       # full words, 2 groups (one ragged)
    ("inst_0_32x32", 64, 0.01, (1.5, 1e-8, 8), 1, 130, ({}, KS(1), K(1, 3), K(2, 32), K(8, 1))),       # 5 groups; odd warp counts
    ("inst_0_32x32", 33, 0.3, (1.5, 1e-8, 6), 2, 3, ({}, K(4))),                            # odd lanes, hot (many draws)
    ("inst_0_32x32", 33, 0.3, (1.5, 1e-8, 6), 2, 66, ({}, KS(1), K(1, 5))),
    ("inst_0_32x32", 2, 0.5, (1.5, 1e-3, 6), 2, 70, ({}, KS(1), PLAIN)),                                 # minimum slices
    ("santoro_80x80", 20, 0.01, (1.5, 1e-8, 4), 1, 2, ({}, K(8))),                          # config 3 shape
    ("santoro_80x80", 20, 0.01, (1.5, 1e-8, 3), 1, 64, ({}, KS(1), KS(2), K(2, 16))),
    ("santoro_80x80", 64, 0.3, (1.5, 1e-8, 3), 2, 34, ({}, KS(1), KS(4))),                               # hot, clusters of 4, ragged group
    ("boixo", 5, 0.05, (0.5, 1e-8, 10), 3, 40, ({}, PLAIN, KS(1), K(2, 2))),                              # config 1: 8 spins, not a lattice
]


@pytest.mark.parametrize("inst,P,T,sch,mcsteps,R,geoms", CASES)
def test_level_qa_bit_exact(golden, dev, inst, P, T, sch, mcsteps, R, geoms):
    nbs, idx, J32, color = _inst(golden, inst)
    n = NSPINS[inst]
    sched = np.linspace(*sch[:2], int(sch[2]))
    seed, r0, s0 = 0xC0FFEE + P, 11, 5
    init = O.colour_init_spins(seed, r0, R, n)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    O.qa_colour(sched, mcsteps, P, T, idx, J32, color, want, seed, r0, s0, 0)
    for env in geoms:
        got = _run_qa(dev, nbs, color, sched, mcsteps, P, T, R, seed, r0, s0, env)
        assert np.array_equal(want, got), "geometry %r" % (env,)


@pytest.mark.parametrize("order", ["natural", "checkerboard", "permutation"])
@pytest.mark.parametrize("R,P,env", [(300, 8, {}), (300, 8, PLAIN), (300, 8, KS(1)), (130, 64, K(2)), (130, 64, KS(2)), (33, 20, K(4, 2))])
def test_level_orders_bit_exact(dev, order, R, P, env):
    """All three visiting orders (static level colourings and a fresh permutation per sweep, level-coloured
    on the host per sweep) against the sequential CPU statement; mcsteps = 2 so that sweeps and schedule
    steps differ."""
    nbs, idx, J32, natural, checker = _torus(8, 11)
    n = 64
    sched = np.linspace(1.5, 1e-8, 6)
    seed = 1000 + R + P
    init = O.colour_init_spins(seed, 5, R, n)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    orders = None
    if order == "checkerboard":
        O.qa_colour(sched, 2, P, 0.03, idx, J32, checker, want, seed, replica0=5, sweep0=3)
        color = checker
    elif order == "natural":
        O.qa_colour(sched, 2, P, 0.03, idx, J32, checker, want, seed, replica0=5, sweep0=3,
                    orders=np.tile(np.arange(n, dtype=np.int32), (12, 1)))
        color = tools.ColourGraph(nbs, "natural")
    else:
        prng = np.random.RandomState(R)
        orders = np.stack([prng.permutation(n) for _ in range(12)]).astype(np.int32)
        O.qa_colour(sched, 2, P, 0.03, idx, J32, checker, want, seed, replica0=5, sweep0=3, orders=orders)
        color = checker
    got = _run_qa(dev, nbs, color, sched, 2, P, 0.03, R, seed, 5, 3, env, orders=orders)
    assert np.array_equal(want, got)


def test_level_generic_function_path(golden, dev):
    """PIQMC_FORCE_GENERIC_FN: every decision function through the truth-table fallback."""
    nbs, idx, J32, color = _inst(golden, "inst_0_32x32")
    sched = np.linspace(1.5, 1e-8, 5)
    init = O.colour_init_spins(3, 0, 4, 1024)
    want = np.repeat(init[:, :, None], 20, axis=2).copy()
    O.qa_colour(sched, 1, 20, 0.2, idx, J32, color, want, 3, 0, 0, 0)
    got = _run_qa(dev, nbs, color, sched, 1, 20, 0.2, 4, 3, 0, 0, {"PIQMC_FORCE_GENERIC_FN": "1"})
    assert np.array_equal(want, got)


@pytest.mark.parametrize("inst,sch,mcsteps,R,geoms", [
    ("inst_0_32x32", (3.0, 0.01, 12), 1, 130, ({}, K(2, 8))),
    ("inst_0_32x32", (3.0, 1.0, 4), 2, 2100, ({},)),            # hot: every thread draws (33 rows)
    ("inst_0_32x32", (3.0, 1.0, 3), 2, 4100, ({}, KS(1), K(1, 7))),    # hot, 66 rows (staged, streamed, plain)
    ("santoro_80x80", (3.0, 0.01, 3), 1, 129, ({}, PLAIN)),
    ("santoro_80x80", (3.0, 0.01, 3), 1, 128, (KS(1), KS(2))),
])
def test_level_sa_bit_exact(golden, dev, inst, sch, mcsteps, R, geoms):
    import piqmc.sa as sa
    nbs, idx, J32, color = _inst(golden, inst)
    n = NSPINS[inst]
    sched = np.linspace(*sch[:2], int(sch[2]))
    rng = np.random.RandomState(R)
    init = (2 * rng.randint(2, size=(R, n)) - 1).astype(np.int8)
    want = init.copy()
    O.sa_colour(sched, mcsteps, idx, J32, color, want, seed=31337, row0=2)
    for env in geoms:
        dev.set_graph(nbs, color)
        dev.set_variant(4)
        try:
            with _Env(env):
                out = sa.AnnealReplicas(sched, mcsteps, init, nbs, 31337, color=color, row0=2, device=dev)
        finally:
            dev.set_variant(0)
        assert np.array_equal(out["spins"], want), "geometry %r" % (env,)


def test_level_sa_permutation_orders(golden, dev):
    """sa.Anneal's visiting order (a fresh permutation per sweep, piqmc/sa.pyx:92,120) through the level kernel."""
    nbs, idx, J32, color = _inst(golden, "inst_0_32x32")
    n, R = 1024, 200
    sched = np.linspace(3.0, 0.05, 5)
    rng = np.random.RandomState(4)
    init = (2 * rng.randint(2, size=(R, n)) - 1).astype(np.int8)
    orders = np.stack([rng.permutation(n) for _ in range(10)]).astype(np.int32)
    want = init.copy()
    O.sa_colour(sched, 2, idx, J32, color, want, seed=5, row0=0, orders=orders)
    import piqmc.sa as sa
    dev.set_variant(4)
    try:
        out = sa.AnnealReplicas(sched, 2, init, nbs, 5, order=orders, device=dev)
    finally:
        dev.set_variant(0)
    assert np.array_equal(out["spins"], want)


@pytest.mark.parametrize("P,R,T", [(20, 100, 0.01), (20, 97, 0.3), (16, 130, 0.05), (4, 520, 0.2), (32, 70, 0.3)])
def test_level_replicas_per_word_bit_exact(dev, P, R, T):
    """P <= 32: floor(64/P) replicas per word through the level kernel == the CPU statement replica by
    replica (same Philox keys as with one replica per word)."""
    import piqmc.qmc as qmc
    nbs, idx, J32, color, _ = _torus(8, 21)
    n = 64
    sched = np.linspace(1.5, 1e-8, 7)
    seed, r0 = 77 + P, 5
    init = O.colour_init_spins(seed, r0, R, n)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    O.qa_colour(sched, 2, P, T, idx, J32, color, want, seed, replica0=r0)
    dev.set_variant(4)
    try:
        many = qmc.QuantumAnnealReplicas(sched, 2, P, T, n, None, nbs, seed, color=color, nreplicas=R, replica0=r0,
                                         device=dev)
        one = qmc.QuantumAnnealReplicas(sched, 2, P, T, n, None, nbs, seed, color=color, nreplicas=R, replica0=r0,
                                        device=dev, per_word=1)
    finally:
        dev.set_variant(0)
    assert many["per_word"] == 64 // P and one["per_word"] == 1
    got = np.transpose(tools.UnpackWords(many["words"], P), (0, 2, 1))
    assert np.array_equal(got, want)
    assert np.array_equal(many["words"], one["words"]) and np.array_equal(many["energies"], one["energies"])


def test_level_staged_is_chosen(dev):
    """Geometry of the staged variant (reported through the launch count: one launch for the tables, one
    for the sweeps) and the cases it leaves to the plain kernel."""
    nbs, idx, J32, color, checker = _torus(16, 3)
    sched = np.linspace(1.5, 1e-8, 4)
    for R, col in ((64, color), (63, color), (64, checker)):
        init = O.colour_init_spins(9, 0, R, 256)
        want = np.repeat(init[:, :, None], 16, axis=2).copy()
        O.qa_colour(sched, 1, 16, 0.05, idx, J32, col, want, 9, 0, 0, 0)
        assert np.array_equal(want, _run_qa(dev, nbs, col, sched, 1, 16, 0.05, R, 9, 0, 0, {}))


@pytest.mark.parametrize("rows,env", [(4096, {}), (512, {}), (512, KS(8)), (512, PLAIN)])
def test_level_config5_full_size_bit_exact_sampled_replicas(dev, rows, env):
    """BASELINE configs[4] in the shape bench.py times (256x256 Gaussian torus, P = 64, natural order; 4096
    rows on one GPU = 128 groups on one block each, 512 rows = the 8-GPU shard = 16 groups on clusters of 8),
    compared with the CPU statement for sampled replicas: the Philox key carries the global replica id, so
    one replica can be re-run on its own."""
    L, P, steps, seed = 256, 64, 6, 2024
    n = L * L
    nbs, idx, J32, color, _ = _torus(L, seed)
    sched = np.linspace(1.5, 1e-8, steps)
    replica0 = 0 if rows == 4096 else 3584                      # rank 7 of 8
    dev.set_graph(nbs, color)
    dev.set_variant(4)
    try:
        with _Env(env):
            dev.state_alloc(rows, P)
            dev.state_init_random(seed, replica0, tile=True)
            dev.qa_colour(sched, 1, 0.01, seed, replica0=replica0)
            words = dev.state_download_words()                  # [rows, n]
    finally:
        dev.set_variant(0)
    for r in (0, rows // 8 - 1, rows // 2, rows - 1):
        init = O.colour_init_spins(seed, replica0 + r, 1, n)
        want = np.repeat(init[:, :, None], P, axis=2).copy()
        O.qa_colour(sched, 1, P, 0.01, idx, J32, color, want, seed, replica0 + r, 0, 0)
        got = np.transpose(tools.UnpackWords(words[r:r + 1], P), (0, 2, 1))
        assert np.array_equal(want, got), "replica %d" % (replica0 + r)
