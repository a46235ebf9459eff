"""GPU parity tests of the chain pipeline (csrc/chain_kernels.cu): the natural-order sweep of the
reference's per-spin-reset variant (piqmc/qmc.pyx:320-357) run as concurrent chains must equal, bit
for bit, the CPU statement of the sequential sweep (oracle/piqmc_oracle.c part 3) -- for every
geometry (one or two rows per thread, one band or several ragged bands per ring, partial and multiple
rings), QA and SA, one and several replicas per word, and at BASELINE.json's full size for sampled
replicas.
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
    nbs, _ = tools.GaussianTorusNeighbors(L, seed)
    idx, J32 = O.nbs_to_ell(nbs)
    return nbs, idx, J32, tools.TorusNaturalLevels(L)


class _Env:
    """geometry knobs of the chain pipeline, read by the library at launch"""

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


def _run_qa(dev, nbs, color, sched, mcsteps, P, T, R, seed, replica0, sweep0, env):
    dev.set_graph(nbs, color)
    dev.set_variant(3)
    try:
        with _Env(env):
            dev.state_alloc(R, P)
            C, nch, per, sel = dev.chain_info()
            assert C >= 8 and sel, "no chain plan: the test would not exercise the chain kernel"
            dev.state_init_random(seed, replica0, tile=True)
            dev.qa_colour(sched, mcsteps, T, seed, replica0=replica0, sweep0=sweep0)
            got = tools.UnpackWords(dev.state_download_words(), P)
    finally:
        dev.set_variant(0)
    return np.ascontiguousarray(np.transpose(got, (0, 2, 1))), C


CASES = [
    # inst, P, T, sched, mcsteps, R, geometries
    ("inst_0_32x32", 20, 0.01, (1.5, 1e-8, 25), 1, 6, ({}, {"PIQMC_CHAIN_CW": "5"})),            # config 2 shape; 7 ragged bands
    ("inst_0_32x32", 64, 0.01, (1.5, 1e-8, 8), 1, 35, ({},)),                                      # full words, odd rows: 2 rings, 1 row per thread
    ("inst_0_32x32", 64, 0.01, (1.5, 1e-8, 8), 1, 130, ({}, {"PIQMC_CHAIN_RPT": "1"}, {"PIQMC_CHAIN_BANDS": "32"})),   # 2 rows per thread, 3 rings; one chain per block
    ("inst_0_32x32", 33, 0.3, (1.5, 1e-8, 6), 2, 3, ({},)),                                        # odd lanes, hot (many draws)
    ("inst_0_32x32", 33, 0.3, (1.5, 1e-8, 6), 2, 66, ({"PIQMC_CHAIN_CW": "3"},)),                  # hot, 2 rows per thread
    ("inst_0_32x32", 2, 0.5, (1.5, 1e-3, 6), 2, 70, ({"PIQMC_CHAIN_G": "1"},)),                    # minimum slices, 2 rings
    ("santoro_80x80", 20, 0.01, (1.5, 1e-8, 4), 1, 2, ({},)),                                      # config 3 shape, 5 bands
    ("santoro_80x80", 20, 0.01, (1.5, 1e-8, 3), 1, 64, ({"PIQMC_CHAIN_CW": "7"},)),
]


@pytest.mark.parametrize("inst,P,T,sch,mcsteps,R,geoms", CASES)
def test_chain_qa_bit_exact(golden, dev, inst, P, T, sch, mcsteps, R, geoms):
    nbs, idx, J32, color = _inst(golden, inst)
    n = NSPINS[inst]
    sched = np.linspace(*sch[:2], int(sch[2]))
    seed, r0, s0 = 0xC0FFEE + P, 11, 5
    init = O.colour_init_spins(seed, r0, R, n)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    O.qa_colour(sched, mcsteps, P, T, idx, J32, color, want, seed, r0, s0, 0)
    for env in geoms:
        got, used = _run_qa(dev, nbs, color, sched, mcsteps, P, T, R, seed, r0, s0, env)
        assert np.array_equal(want, got), "geometry %r" % (env,)


def test_chain_rejects_what_it_cannot_run(golden, dev):
    """Graphs that are not lattices of rows (boixo: 8 spins, K_4,4 cell) have no plan and run through the
    other kernels; the result is the same sequential sweep."""
    nbs = golden["vec"]["nbs_boixo"]
    dev.set_graph(nbs, tools.OrderLevels(nbs))
    assert dev.chain_info()[0] == 0 and not dev.chain_info()[3]


def test_chain_qa_gaussian_torus_rows_and_selection(dev):
    """Config 5's family (Gaussian torus, P = 64): the plan is one chain per lattice row; the pipeline is
    opt-in (variant 3), the default variant runs the dataflow kernel."""
    nbs, idx, J32, color = _torus(16, 2024)
    sched = np.linspace(1.5, 1e-8, 10)
    R, seed = 45, 2024
    init = O.colour_init_spins(seed, 4091, R, 256)
    want = np.repeat(init[:, :, None], 64, axis=2).copy()
    O.qa_colour(sched, 1, 64, 0.01, idx, J32, color, want, seed, 4091, 17, 0)
    dev.set_graph(nbs, color)
    dev.set_variant(0)
    C, nch, per, sel = dev.chain_info()
    assert (C, nch) == (16, 16) and not sel and per < 40
    dev.set_variant(3)
    try:
        assert dev.chain_info()[3]
        dev.state_alloc(R, 64)
        dev.state_init_random(seed, 4091, tile=True)
        l0 = dev.launch_count
        dev.qa_colour(sched, 1, 0.01, seed, replica0=4091, sweep0=17)
        assert dev.launch_count - l0 == 2                      # decision tables + the sweeps: one launch each
        got = np.transpose(tools.UnpackWords(dev.state_download_words(), 64), (0, 2, 1))
    finally:
        dev.set_variant(0)
    assert np.array_equal(want, got)
    # the dataflow kernel on the same colouring gives the same state (both equal the sequential sweep)
    dev.set_variant(2)
    try:
        dev.state_init_random(seed, 4091, tile=True)
        dev.qa_colour(sched, 1, 0.01, seed, replica0=4091, sweep0=17)
        got2 = np.transpose(tools.UnpackWords(dev.state_download_words(), 64), (0, 2, 1))
    finally:
        dev.set_variant(0)
    assert np.array_equal(want, got2)


def test_chain_generic_function_path(golden, dev):
    """PIQMC_FORCE_GENERIC_FN: every decision function through the truth-table fallback."""
    nbs, idx, J32, color = _inst(golden, "inst_0_32x32")
    sched = np.linspace(1.5, 1e-8, 5)
    init = O.colour_init_spins(3, 0, 4, 1024)
    want = np.repeat(init[:, :, None], 20, axis=2).copy()
    O.qa_colour(sched, 1, 20, 0.2, idx, J32, color, want, 3, 0, 0, 0)
    os.environ["PIQMC_FORCE_GENERIC_FN"] = "1"
    try:
        got, _ = _run_qa(dev, nbs, color, sched, 1, 20, 0.2, 4, 3, 0, 0, {})
    finally:
        del os.environ["PIQMC_FORCE_GENERIC_FN"]
    assert np.array_equal(want, got)


@pytest.mark.parametrize("inst,sch,mcsteps,R,chains", [
    ("inst_0_32x32", (3.0, 0.01, 12), 1, 130, ({}, {"PIQMC_CHAIN_CW": "6"})),
    ("inst_0_32x32", (3.0, 1.0, 4), 2, 2100, ({},)),            # hot: every thread draws (33 rows)
    ("inst_0_32x32", (3.0, 1.0, 3), 2, 4100, ({"PIQMC_CHAIN_RPT": "2"},)),   # hot, 66 rows, 2 rows per thread
    ("santoro_80x80", (3.0, 0.01, 3), 1, 65, ({},)),
])
def test_chain_sa_bit_exact(golden, dev, inst, sch, mcsteps, R, chains):
    import piqmc.sa as sa
    nbs, idx, J32, color = _inst(golden, inst)
    n = NSPINS[inst]
    sched = np.linspace(*sch[:2], int(sch[2]))
    rng = np.random.RandomState(R)
    init = (2 * rng.randint(2, size=(R, n)) - 1).astype(np.int8)
    want = init.copy()
    O.sa_colour(sched, mcsteps, idx, J32, color, want, seed=31337, row0=2)
    for env in chains:
        dev.set_graph(nbs, color)
        dev.set_variant(3)
        try:
            with _Env(env):
                assert dev.chain_info()[0] >= 8
                l0 = dev.launch_count
                out = sa.AnnealReplicas(sched, mcsteps, init, nbs, 31337, color=color, row0=2, device=dev)
        finally:
            dev.set_variant(0)
        assert np.array_equal(out["spins"], want), "geometry %r" % (env,)


@pytest.mark.parametrize("P,R,T", [(20, 100, 0.01), (20, 97, 0.3), (16, 130, 0.05), (4, 520, 0.2), (32, 70, 0.3)])
def test_chain_replicas_per_word_bit_exact(dev, P, R, T):
    """P <= 32: floor(64/P) replicas per word through the chain kernel == the CPU statement replica by
    replica (same Philox keys as with one replica per word)."""
    import piqmc.qmc as qmc
    nbs, idx, J32, color = _torus(8, 21)
    n = 64
    sched = np.linspace(1.5, 1e-8, 7)
    seed, r0 = 77 + P, 5
    init = O.colour_init_spins(seed, r0, R, n)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    O.qa_colour(sched, 2, P, T, idx, J32, color, want, seed, replica0=r0)
    dev.set_variant(3)
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


@pytest.mark.parametrize("variant", [2, 3])
@pytest.mark.parametrize("rows", [4096, 512])
def test_config5_full_size_bit_exact_sampled_replicas(dev, rows, variant):
    """BASELINE configs[4] in the shape bench.py times (256x256 Gaussian torus, P = 64, natural order;
    4096 rows on one GPU, 512 rows = the 8-GPU shard), through the dataflow kernel bench.py measures
    (variant 2: 8 row chunks x 4 passes per block at 4096 rows, 511 levels, period-major tickets) and
    through the chain pipeline (variant 3), compared with the CPU statement for sampled replicas: the
    Philox key carries the global replica id, so one replica can be re-run on its own."""
    L, P, steps, seed = 256, 64, 6, 2024
    n = L * L
    nbs, idx, J32, color = _torus(L, seed)
    sched = np.linspace(1.5, 1e-8, steps)
    replica0 = 0 if rows == 4096 else 3584                      # rank 7 of 8
    dev.set_graph(nbs, color)
    dev.set_variant(variant)
    try:
        dev.state_alloc(rows, P)
        assert dev.chain_info()[3] == (variant == 3)
        dev.state_init_random(seed, replica0, tile=True)
        dev.qa_colour(sched, 1, 0.01, seed, replica0=replica0)
        words = dev.state_download_words()                      # [rows, n]
    finally:
        dev.set_variant(0)
    for r in (0, rows // 8 - 1, rows // 2, rows - 1):
        init = O.colour_init_spins(seed, replica0 + r, 1, n)
        want = np.repeat(init[:, :, None], P, axis=2).copy()
        O.qa_colour(sched, 1, P, 0.01, idx, J32, color, want, seed, replica0 + r, 0, 0)
        got = np.transpose(tools.UnpackWords(words[r:r + 1], P), (0, 2, 1))
        assert np.array_equal(want, got), "replica %d" % (replica0 + r)
