"""GPU parity tests, production colour-parallel paths.

Two levels (BASELINE.json north_star):
  1. bit-exact: the CUDA kernels on bit-packed state against the CPU statement of the same
     semantics (oracle/piqmc_oracle.c part 3) -- spins identical, energies <= 1e-12 relative;
  2. statistical: residual-energy distribution over >= 1000 replicas against the REFERENCE's own
     per-spin-reset variant qmc.QuantumAnneal_parallel (golden distributions produced by the
     compiled reference), two-sample KS at p > 0.01.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
import piqmc.tools as tools
from helpers import GS_ENERGY, NSPINS, ks_2samp_p
from test_oracle import _J

pytestmark = pytest.mark.gpu


def _graph(golden, inst):
    import piqmc.tools as T
    nbs = golden["vec"]["nbs_" + inst]
    idx, J32 = O.nbs_to_ell(nbs)
    return nbs, idx, J32, T.ColourGraph(nbs)


def _torus(L, seed):
    import piqmc.tools as T
    nbs, color = T.GaussianTorusNeighbors(L, seed)
    idx, J32 = O.nbs_to_ell(nbs)
    return nbs, idx, J32, color


def _qa_both(dev, nbs, idx, J32, color, sched, mcsteps, P, T, R, seed, replica0=0, sweep0=0,
             trotter=0, variant=0):
    """Run oracle and device from the same Philox start; return (oracle int8[R,N,P], device same)."""
    import piqmc.tools as tools
    n = nbs.shape[0]
    init = O.colour_init_spins(seed, replica0, R, n)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    O.qa_colour(sched, mcsteps, P, T, idx, J32, color, want, seed, replica0, sweep0, trotter)
    dev.set_graph(nbs, color)
    dev.set_variant(variant)
    dev.state_alloc(R, P)
    dev.state_init_random(seed, replica0, tile=True)
    w0 = dev.state_download_words()
    assert np.array_equal(tools.UnpackWords(w0, P)[:, 0, :], init)       # device init == oracle init
    dev.qa_colour(sched, mcsteps, T, seed, replica0=replica0, sweep0=sweep0, trotter=trotter)
    got = tools.UnpackWords(dev.state_download_words(), P)               # [R,P,N]
    dev.set_variant(0)
    return want, np.ascontiguousarray(np.transpose(got, (0, 2, 1)))


QA_CASES = [
    # inst, P, T, sched, mcsteps, R
    ("boixo", 5, 0.01, (0.5, 1e-8, 10), 3, 9),
    ("boixo", 20, 0.01, (0.5, 1e-8, 10), 3, 5),
    ("boixo16", 8, 0.05, (1.0, 1e-8, 12), 2, 7),
    ("bipartite8", 10, 0.01, (8.0, 1e-8, 5), 20, 6),          # maxnb 5, fields, pad rows
    ("hopfield8", 10, 0.01, (8.0, 1e-8, 5), 20, 6),           # K8: 8 colours, maxnb 8
    ("inst_0_32x32", 20, 0.01, (1.5, 1e-8, 25), 1, 6),        # config 2 shape
    ("inst_0_32x32", 64, 0.01, (1.5, 1e-8, 8), 1, 3),         # full 64-lane words
    ("inst_0_32x32", 33, 0.3, (1.5, 1e-8, 6), 2, 3),          # odd lane count, hot (many draws)
    ("inst_0_32x32", 2, 0.5, (1.5, 1e-3, 6), 2, 4),           # minimum slices
    ("santoro_80x80", 20, 0.01, (1.5, 1e-8, 4), 1, 2),        # config 3 shape
]


@pytest.mark.parametrize("variant", [1, 2, 5])
@pytest.mark.parametrize("inst,P,T,sch,mcsteps,R", QA_CASES)
def test_qa_colour_bit_exact(golden, dev, inst, P, T, sch, mcsteps, R, variant):
    nbs, idx, J32, color = _graph(golden, inst)
    sched = np.linspace(*sch[:2], int(sch[2]))
    want, got = _qa_both(dev, nbs, idx, J32, color, sched, mcsteps, P, T, R, seed=0xC0FFEE + P,
                         variant=variant)
    assert np.array_equal(want, got)
    # device energies of the packed state == ClassicalIsingEnergy per (replica, slice)
    en = dev.energy()
    J = _J(golden, inst)
    ref = np.array([[O.energy(J, got[r, :, k].astype(np.float64)) for k in range(P)] for r in range(R)])
    np.testing.assert_allclose(en, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("variant", [1, 2])
def test_qa_colour_gaussian_torus_and_offsets(dev, variant):
    """Config 5's instance family at a size the oracle finishes in seconds; non-zero replica0
    and sweep0 (sharding / continuation must not change the streams)."""
    nbs, idx, J32, color = _torus(16, 2024)
    sched = np.linspace(1.5, 1e-8, 10)
    want, got = _qa_both(dev, nbs, idx, J32, color, sched, 1, 64, 0.01, 5, seed=2024, replica0=4091,
                         sweep0=17, variant=variant)
    assert np.array_equal(want, got)


@pytest.mark.parametrize("variant", [1, 2])
def test_qa_colour_periodic_trotter(golden, dev, variant):
    nbs, idx, J32, color = _graph(golden, "inst_0_32x32")
    sched = np.linspace(1.5, 1e-8, 8)
    for P in (20, 7):
        want, got = _qa_both(dev, nbs, idx, J32, color, sched, 1, P, 0.05, 3, seed=5, trotter=1,
                             variant=variant)
        assert np.array_equal(want, got)


@pytest.mark.parametrize("inst", ["boixo16", "hopfield8"])
def test_resident_kernel_periodic_trotter_and_orders(golden, dev, inst):
    """The resident kernel (state of a row in shared memory, sequential visits): periodic Trotter neighbours
    with even and odd slice counts, and per-sweep visiting orders (the orders themselves are the sweep)."""
    nbs, idx, J32, color = _graph(golden, inst)
    n = NSPINS[inst]
    sched = np.linspace(2.0, 1e-8, 6)
    for P in (12, 7, 64):
        want, got = _qa_both(dev, nbs, idx, J32, color, sched, 2, P, 0.3, 37, seed=5 + P, trotter=1, variant=5)
        assert np.array_equal(want, got)
    prng = np.random.RandomState(3)
    orders = np.stack([prng.permutation(n) for _ in range(12)]).astype(np.int32)
    R, P = 45, 10
    init = O.colour_init_spins(8, 2, R, n)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    O.qa_colour(sched, 2, P, 0.3, idx, J32, color, want, 8, replica0=2, sweep0=4, orders=orders)
    dev.set_graph(nbs, color)
    dev.set_variant(5)
    try:
        dev.state_alloc(R, P)
        dev.state_init_random(8, 2, tile=True)
        l0 = dev.launch_count
        dev.qa_colour(sched, 2, 0.3, 8, replica0=2, sweep0=4, orders=orders)
        assert dev.launch_count - l0 == 1                      # the whole run: one launch
        got = np.transpose(tools.UnpackWords(dev.state_download_words(), P), (0, 2, 1))
    finally:
        dev.set_variant(0)
    assert np.array_equal(want, got)


def test_qa_colour_sharding_invariance(dev):
    """Replicas [0,8) in one state == replicas [0,3) and [3,8) run separately with replica0."""
    import piqmc.qmc as qmc
    nbs, idx, J32, color = _torus(12, 7)
    sched = np.linspace(1.5, 1e-8, 6)
    full = qmc.QuantumAnnealReplicas(sched, 1, 16, 0.01, 144, None, nbs, 99, color=color, nreplicas=8)
    a = qmc.QuantumAnnealReplicas(sched, 1, 16, 0.01, 144, None, nbs, 99, color=color, nreplicas=3)
    b = qmc.QuantumAnnealReplicas(sched, 1, 16, 0.01, 144, None, nbs, 99, color=color, nreplicas=5,
                                  replica0=3)
    assert np.array_equal(full["words"], np.concatenate([a["words"], b["words"]]))
    assert np.array_equal(full["energies"], np.concatenate([a["energies"], b["energies"]]))


@pytest.mark.parametrize("variant", [1, 2, 5])
@pytest.mark.parametrize("inst,sch,mcsteps,R", [
    ("boixo", (1.0, 0.01, 10), 3, 70),
    ("boixo16", (2.0, 0.5, 6), 2, 1000),                        # hot: most lanes draw
    ("bipartite8", (3.0, 0.01, 10), 2, 64),
    ("hopfield8", (8.0, 1e-8, 5), 10, 130),
    ("inst_0_32x32", (3.0, 0.01, 12), 1, 130),
    ("santoro_80x80", (3.0, 0.01, 3), 1, 65),
])
def test_sa_colour_bit_exact(golden, dev, inst, sch, mcsteps, R, variant):
    import piqmc.sa as sa
    nbs, idx, J32, color = _graph(golden, inst)
    n = NSPINS[inst]
    sched = np.linspace(*sch[:2], int(sch[2]))
    rng = np.random.RandomState(R)
    init = (2 * rng.randint(2, size=(R, n)) - 1).astype(np.int8)
    want = init.copy()
    O.sa_colour(sched, mcsteps, idx, J32, color, want, seed=31337, row0=2)
    dev.set_variant(variant)
    out = sa.AnnealReplicas(sched, mcsteps, init, nbs, 31337, color=color, row0=2, device=dev)
    dev.set_variant(0)
    assert np.array_equal(out["spins"], want)
    J = _J(golden, inst)
    ref = np.array([O.energy(J, s.astype(np.float64)) for s in want])
    np.testing.assert_allclose(out["energies"], ref, rtol=1e-12, atol=1e-12)


def test_sa_colour_random_start_matches_oracle_init(dev):
    import piqmc.sa as sa
    nbs, idx, J32, color = _torus(8, 1)
    R = 100
    init = np.concatenate([O.colour_init_spins(5, 64 * g + l, 1, 64) for g in range(2) for l in range(64)])[:R]
    want = init.copy()
    sched = np.linspace(3.0, 0.01, 5)
    O.sa_colour(sched, 1, idx, J32, color, want, seed=5)
    out = sa.AnnealReplicas(sched, 1, None, nbs, 5, color=color, nreplicas=R, device=dev)
    assert np.array_equal(out["spins"], want)


@pytest.mark.parametrize("R,P", [(100, 20), (64, 64), (7, 5)])
def test_resident_handover_sa_to_qa_and_results(dev, R, P):
    """SA pre-anneal -> PIQMC hand-over on the device (examples/spinglass32.py:124-127, np.tile of
    the annealed vector over the slices) == download + tiled upload through the host; results()
    (energies + words in one call) == energy() + state_download_words()."""
    import piqmc.qmc as qmc
    import piqmc.sa as sa
    nbs, idx, J32, color = _torus(8, 3)
    ssched, qsched = np.linspace(3.0, 0.01, 6), np.linspace(1.5, 1e-8, 5)
    pre = sa.AnnealReplicas(ssched, 1, None, nbs, 11, color=color, nreplicas=R, device=dev)
    host = qmc.QuantumAnnealReplicas(qsched, 1, P, 0.01, 64, pre["spins"], nbs, 12, color=color, device=dev)
    # the pre-anneal runs under another colouring (fresh permutation per sweep): the resident state
    # survives the change of colouring
    pre2 = sa.AnnealReplicas(ssched, 1, None, nbs, 11, order="permutation", nreplicas=R, device=dev)
    host2 = qmc.QuantumAnnealReplicas(qsched, 1, P, 0.01, 64, pre2["spins"], nbs, 12, color=color, device=dev)
    sa.AnnealReplicas(ssched, 1, None, nbs, 11, order="permutation", nreplicas=R, device=dev, download=False)
    res2 = qmc.QuantumAnnealReplicas(qsched, 1, P, 0.01, 64, "resident", nbs, 12, color=color, nreplicas=R,
                                     device=dev)
    assert np.array_equal(res2["words"], host2["words"])
    sa.AnnealReplicas(ssched, 1, None, nbs, 11, color=color, nreplicas=R, device=dev, download=False)
    # (at P = 20 the hand-over also changes the layout: 3 replicas per word, the last word padded)
    res = qmc.QuantumAnnealReplicas(qsched, 1, P, 0.01, 64, "resident", nbs, 12, color=color, nreplicas=R,
                                    device=dev, per_word=3 if P == 20 else 1)
    assert res["per_word"] == (3 if P == 20 else 1) and host["per_word"] == 1
    assert np.array_equal(res["words"], host["words"])
    assert np.array_equal(res["energies"], host["energies"])
    # the separate calls agree with the combined one
    assert np.array_equal(dev.energy()[:R], res["energies"])          # [:R]: a packed state pads its last word
    assert np.array_equal(dev.state_download_words()[:R], res["words"])
    # and with the CPU statement of the whole chain
    want = pre["spins"].copy()
    ref = np.repeat(want[:, :, None], P, axis=2).copy()
    O.qa_colour(qsched, 1, P, 0.01, idx, J32, color, ref, 12)
    got = np.transpose(tools.UnpackWords(res["words"], P), (0, 2, 1))
    assert np.array_equal(got, ref)


# ------------------------------------------------------------------- order-equivalent colourings
@pytest.mark.parametrize("inst", ["boixo", "hopfield8", "inst_0_32x32"])
def test_level_colouring_equals_sequential_sweep_qa(golden, dev, inst):
    """order="natural": the colour-class sweep over the dependency levels of the natural order is
    bit-identical to the plain sequential natural-order sweep (the order of the reference's
    QuantumAnneal_parallel, qmc.pyx:320); same for explicit per-sweep permutations."""
    import piqmc.qmc as qmc
    import piqmc.tools as T
    from piqmc.device import order_levels
    nbs, idx, J32, _ = _graph(golden, inst)
    n, P, R = NSPINS[inst], 12, 4
    sched = np.linspace(1.5, 1e-8, 7)
    lev = T.ColourGraph(nbs, "natural")
    assert np.array_equal(lev, order_levels(nbs))                   # C helper == Python statement
    init = O.colour_init_spins(11, 0, R, n)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    O.qa_colour(sched, 2, P, 0.02, idx, J32, lev, want, 11, orders=np.tile(np.arange(n, dtype=np.int32), (14, 1)))
    out = qmc.QuantumAnnealReplicas(sched, 2, P, 0.02, n, None, nbs, 11, order="natural", nreplicas=R, device=dev)
    got = np.transpose(T.UnpackWords(out["words"], P), (0, 2, 1))
    assert np.array_equal(want, got)
    # per-sweep random permutations
    prng = np.random.RandomState(3)
    orders = np.stack([prng.permutation(n) for _ in range(14)]).astype(np.int32)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    O.qa_colour(sched, 2, P, 0.02, idx, J32, lev, want, 11, orders=orders)
    out = qmc.QuantumAnnealReplicas(sched, 2, P, 0.02, n, None, nbs, 11, order=orders, nreplicas=R, device=dev)
    got = np.transpose(T.UnpackWords(out["words"], P), (0, 2, 1))
    assert np.array_equal(want, got)


def test_level_colouring_equals_sequential_sweep_sa(golden, dev):
    import piqmc.sa as sa
    nbs, idx, J32, color = _graph(golden, "inst_0_32x32")
    n, R = 1024, 70
    sched = np.linspace(3.0, 0.01, 9)
    prng = np.random.RandomState(8)
    init = (2 * prng.randint(2, size=(R, n)) - 1).astype(np.int8)
    # order="permutation" draws RandomState(seed).permutation(n) once per sweep
    seed = 4242
    orng = np.random.RandomState(seed)
    orders = np.stack([orng.permutation(n) for _ in range(9)]).astype(np.int32)
    want = init.copy()
    O.sa_colour(sched, 1, idx, J32, color, want, seed, orders=orders)
    out = sa.AnnealReplicas(sched, 1, init, nbs, seed, order="permutation", device=dev)
    assert np.array_equal(out["spins"], want)
    want = init.copy()
    O.sa_colour(sched, 1, idx, J32, color, want, seed, orders=np.tile(np.arange(n, dtype=np.int32), (9, 1)))
    out = sa.AnnealReplicas(sched, 1, init, nbs, seed, order="natural", device=dev)
    assert np.array_equal(out["spins"], want)


def test_torus_natural_levels_closed_form():
    import piqmc.tools as T
    nbs, _ = T.GaussianTorusNeighbors(10, 1)
    assert np.array_equal(T.TorusNaturalLevels(10), T.OrderLevels(nbs))


def test_parallel_signatures_in_place(golden, dev):
    """qmc.QuantumAnneal_parallel / sa.Anneal_parallel keep the reference's signatures and act in place."""
    import ctypes
    import piqmc.qmc as qmc
    import piqmc.sa as sa
    nbs = golden["vec"]["nbs_boixo"]
    J = _J(golden, "boixo")
    rng = np.random.RandomState(1)
    ctypes.CDLL(None).srand(3)
    confs = np.tile(np.array([2 * rng.randint(2) - 1 for _ in range(8)], dtype=np.float64), (5, 1)).T
    assert qmc.QuantumAnneal_parallel(np.linspace(0.5, 1e-8, 10), 3, 5, 0.01, 8, confs, nbs, 4) is None
    assert all(sa.ClassicalIsingEnergy(confs[:, k], J) == -8.0 for k in range(5))
    sv = np.array([2 * rng.randint(2) - 1 for _ in range(8)], dtype=np.float64)
    assert sa.Anneal_parallel(np.linspace(1.0, 0.01, 10), 3, sv, nbs, 4) is None
    assert set(np.unique(sv)) <= {-1.0, 1.0}


# ------------------------------------------------------------------------------------ statistics
def _residual(en, inst):
    return (np.asarray(en) - GS_ENERGY[inst]) / NSPINS[inst]


@pytest.mark.parametrize("nsteps", [100, 30])
def test_qa_colour_residual_energy_distribution_vs_reference(golden, dev, nsteps):
    """Config 2 (examples/spinglass32.py:58-68): N=1024, P=20, T=0.01, Gamma 1.5->1e-8 in `nsteps`
    steps, 1024 replicas.  Reference sample: qmc.QuantumAnneal_parallel(nthreads=1), the
    reference's own per-spin-reset variant (golden, made by the compiled reference).  The
    production kernel runs the natural-order level colouring (order="natural"), i.e. the same
    sequential sweep.  Statistic: per-replica residual energy per spin, slice-averaged and
    best-slice.  KS p > 0.01.  (order="checkerboard" fails this test at T=0.01: a zero-temperature
    quench depends on the visiting order -- 0.276 vs 0.246 residual per spin; see DESIGN.md.)"""
    import piqmc.qmc as qmc
    nbs = golden["vec"]["nbs_inst_0_32x32"]
    ref = golden["dist"]["qa_par_%d" % nsteps]
    R = ref.shape[0]
    assert R >= 1000
    out = qmc.QuantumAnnealReplicas(np.linspace(1.5, 1e-8, nsteps), 1, 20, 0.01, 1024, None, nbs,
                                    seed=20240 + nsteps, order="natural", nreplicas=R, device=dev)
    mine = out["energies"]
    p_mean = ks_2samp_p(_residual(mine.mean(axis=1), "inst_0_32x32"), _residual(ref.mean(axis=1), "inst_0_32x32"))
    p_min = ks_2samp_p(_residual(mine.min(axis=1), "inst_0_32x32"), _residual(ref.min(axis=1), "inst_0_32x32"))
    print("nsteps", nsteps, "residual/spin mine %.4f ref %.4f  KS p(mean) %.3f p(min) %.3f"
          % (_residual(mine.mean(), "inst_0_32x32"), _residual(ref.mean(), "inst_0_32x32"), p_mean, p_min))
    assert p_mean > 0.01 and p_min > 0.01


@pytest.mark.parametrize("nsteps", [100, 30])
def test_sa_colour_residual_energy_distribution_vs_reference(golden, dev, nsteps):
    """sa.Anneal on config 2's pre-anneal schedule shape (3.0 -> 0.01), 1024 replicas, against
    the reference's sa.Anneal sample.  KS p > 0.01."""
    import piqmc.sa as sa
    nbs = golden["vec"]["nbs_inst_0_32x32"]
    ref = golden["dist"]["sa_%d" % nsteps][:, 0]
    out = sa.AnnealReplicas(np.linspace(3.0, 0.01, nsteps), 1, None, nbs, seed=777 + nsteps,
                            order="permutation", nreplicas=ref.shape[0], device=dev)
    p = ks_2samp_p(_residual(out["energies"], "inst_0_32x32"), _residual(ref, "inst_0_32x32"))
    print("nsteps", nsteps, "residual/spin mine %.4f ref %.4f KS p %.3f"
          % (_residual(out["energies"].mean(), "inst_0_32x32"), _residual(ref.mean(), "inst_0_32x32"), p))
    assert p > 0.01


# ------------------------------------------------------------------------------------ config 4
@pytest.mark.parametrize("inst", ["bipartite8", "hopfield8"])
def test_config4_state_populations_vs_reference(golden, dev, inst):
    """BASELINE configs[3]: bipartite8 / hopfield8 through the multispin-coded SA kernel (64
    replicas per word).  Final-state populations over the 2^8 states and final energies of 4096
    replicas against 4096 runs of the reference's sa.Anneal (golden): chi-square on the state
    histogram (rare states pooled) and KS on the energies, p > 0.01."""
    import os
    from scipy.stats import chi2_contingency
    import piqmc.sa as sa
    g4 = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_config4.npz"))
    nbs = golden["vec"]["nbs_" + inst]
    sched, mcsteps = g4["sched_" + inst], int(g4["mcsteps_" + inst])
    ref_state, ref_en = g4["sa_state_" + inst], g4["sa_energy_" + inst]
    R = ref_state.size
    out = sa.AnnealReplicas(sched, mcsteps, None, nbs, seed=404, order="permutation", nreplicas=R, device=dev)
    mine_state = ((out["spins"] < 0) * (1 << np.arange(8))).sum(axis=1)
    a = np.bincount(mine_state, minlength=256).astype(float)
    b = np.bincount(ref_state, minlength=256).astype(float)
    big = (a + b) >= 20
    table = np.array([np.append(a[big], a[~big].sum()), np.append(b[big], b[~big].sum())])
    table = table[:, table.sum(axis=0) > 0]
    p_chi = chi2_contingency(table)[1]
    # 8-spin instances have a handful of discrete energy levels: compare them as levels (the device
    # reduction and the reference's dense matvec sum in different orders and differ in the last ulp)
    p_ks = ks_2samp_p(np.round(out["energies"], 9), np.round(ref_en, 9))
    print(inst, "states used", int(big.sum()), "chi2 p %.3f  KS(energy) p %.3f  mean E mine %.3f ref %.3f"
          % (p_chi, p_ks, out["energies"].mean(), ref_en.mean()))
    assert p_chi > 0.01 and p_ks > 0.01


# ------------------------------------------------------------------------------------ config 3
@pytest.mark.parametrize("tau", [10, 30, 100, 300, 1000])
def test_santoro_residual_energy_vs_tau(golden, dev, tau):
    """BASELINE configs[2] (examples/santoro80.py:23-33): the 80x80 Martonak-Santoro-Tosatti
    instance, P=20, T=0.01, Gamma 1.5 -> 1e-8 in tau steps; residual energy above the known ground
    state (examples/ising_instances/santoro_80x80_answer.txt:24).  1024 replicas on the GPU against
    256 runs of the reference's QuantumAnneal_parallel (golden): KS p > 0.01."""
    import os
    import piqmc.qmc as qmc
    # tau <= 100: 256 reference runs each; tau = 300 / 1000 (the long end of examples/santoro80.py:290-323): 256 / 128
    fn = "ref_santoro.npz" if tau <= 100 else "ref_santoro_long.npz"
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", fn))["qa_par_%d" % tau]
    nbs = golden["vec"]["nbs_santoro_80x80"]
    out = qmc.QuantumAnnealReplicas(np.linspace(1.5, 1e-8, tau), 1, 20, 0.01, 6400, None, nbs, seed=8080 + tau,
                                    order="natural", nreplicas=1024, device=dev)
    mine = out["energies"]
    p = ks_2samp_p(_residual(mine.mean(axis=1), "santoro_80x80"), _residual(ref.mean(axis=1), "santoro_80x80"))
    print("tau", tau, "residual/spin mine %.4f ref %.4f KS p %.3f"
          % (_residual(mine.mean(), "santoro_80x80"), _residual(ref.mean(), "santoro_80x80"), p))
    assert p > 0.01
    assert _residual(mine.min(), "santoro_80x80") > 0.0          # never below the exact ground state


# ------------------------------------------------------------------- fast kernel, larger states
@pytest.mark.parametrize("order", ["natural", "checkerboard", "permutation"])
@pytest.mark.parametrize("R,P", [(300, 8), (130, 64), (33, 20)])
def test_fast_kernel_many_rows_bit_exact(dev, order, R, P):
    """The dataflow kernel with several row chunks / several passes per block, non-multiple-of-32
    row counts and all three visiting orders, against the sequential CPU statement."""
    import piqmc.qmc as qmc
    import piqmc.tools as T
    nbs, idx, J32, checker = _torus(8, 11)
    n = 64
    sched = np.linspace(1.5, 1e-8, 6)
    seed = 1000 + R + P
    init = O.colour_init_spins(seed, 5, R, n)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    if order == "checkerboard":
        O.qa_colour(sched, 2, P, 0.03, idx, J32, checker, want, seed, replica0=5, sweep0=3)
        kw = dict(color=checker)
    elif order == "natural":
        O.qa_colour(sched, 2, P, 0.03, idx, J32, checker, want, seed, replica0=5, sweep0=3,
                    orders=np.tile(np.arange(n, dtype=np.int32), (12, 1)))
        kw = dict(color=T.ColourGraph(nbs, "natural"))
    else:
        prng = np.random.RandomState(R)
        orders = np.stack([prng.permutation(n) for _ in range(12)]).astype(np.int32)
        O.qa_colour(sched, 2, P, 0.03, idx, J32, checker, want, seed, replica0=5, sweep0=3, orders=orders)
        kw = dict(order=orders)
    dev.set_variant(2)
    if "color" in kw:
        dev.set_graph(nbs, kw["color"])
    else:
        dev.set_graph(nbs, checker)
    dev.state_alloc(R, P)
    dev.state_init_random(seed, 5, tile=True)
    dev.qa_colour(sched, 2, 0.03, seed, replica0=5, sweep0=3, orders=kw.get("order"))
    got = np.transpose(T.UnpackWords(dev.state_download_words(), P), (0, 2, 1))
    dev.set_variant(0)
    assert np.array_equal(want, got)


def test_fast_kernel_sa_many_rows_bit_exact(golden, dev):
    """SA through the dataflow kernel with 40 rows (2560 replicas), hot and cold temperatures."""
    import piqmc.sa as sa
    nbs, idx, J32, color = _torus(8, 12)
    R, n = 2560, 64
    rng = np.random.RandomState(1)
    init = (2 * rng.randint(2, size=(R, n)) - 1).astype(np.int8)
    for sch in ((3.0, 0.5, 4), (0.3, 0.01, 4)):
        sched = np.linspace(*sch[:2], sch[2])
        want = init.copy()
        O.sa_colour(sched, 2, idx, J32, color, want, seed=77, row0=9)
        out = sa.AnnealReplicas(sched, 2, init, nbs, 77, color=color, row0=9, device=dev)
        assert np.array_equal(out["spins"], want)


# ------------------------------------------------------------------- world-line (global) moves
@pytest.mark.parametrize("trotter", [0, 1])
@pytest.mark.parametrize("inst,P,T", [("boixo", 6, 0.3), ("inst_0_32x32", 20, 0.2), ("inst_0_32x32", 64, 0.01)])
def test_world_line_moves_bit_exact(golden, dev, inst, P, T, trotter):
    """World-line moves (a north_star capability the reference does not have): the kernel against
    this repository's CPU statement of them, bit for bit, and a sanity check that they do act."""
    import piqmc.tools as tools
    nbs, idx, J32, color = _graph(golden, inst)
    n, R, seed = NSPINS[inst], 40, 4711
    sched = np.linspace(1.5, 0.2, 5)
    init = O.colour_init_spins(seed, 0, R, n)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    O.qa_colour(sched, 1, P, T, idx, J32, color, want, seed, trotter=trotter, global_moves=True)
    plain = np.repeat(init[:, :, None], P, axis=2).copy()
    O.qa_colour(sched, 1, P, T, idx, J32, color, plain, seed, trotter=trotter, global_moves=False)
    dev.set_graph(nbs, color)
    dev.set_variant(2)
    dev.set_global_moves(True)
    try:
        dev.state_alloc(R, P)
        dev.state_init_random(seed, 0, tile=True)
        dev.qa_colour(sched, 1, T, seed, trotter=trotter)
        got = np.transpose(tools.UnpackWords(dev.state_download_words(), P), (0, 2, 1))
    finally:
        dev.set_global_moves(False)
        dev.set_variant(0)
    assert np.array_equal(want, got)
    if T > 0.1:
        assert not np.array_equal(want, plain)          # at these temperatures some global moves are accepted


# --------------------------------------------------------------------- BASELINE.json full size
def test_config5_full_size_properties(dev):
    """BASELINE configs[4] at full size (256x256 Gaussian torus, P = 64, R = 4096; the oracle would
    need hours), through properties that do not depend on size:
      * reproducibility: the same seed gives the same 2.1 GB state, word for word;
      * sharding invariance: replicas [0, R) in one state == [0, R/2) and [R/2, R) run on their own
        with the global replica id in the Philox key (what the multi-GPU path relies on);
      * energy consistency: device energies == ClassicalIsingEnergy recomputed on the host from the
        downloaded words, for sampled (replica, slice) pairs, to 1e-12 relative;
      * every energy lies above the trivial bound -sum|J| and the anneal lowers the mean energy."""
    import piqmc.qmc as qmc
    import piqmc.tools as T
    L, P, R, steps = 256, 64, 4096, 6
    n = L * L
    nbs, _ = T.GaussianTorusNeighbors(L, 2024)
    color = T.TorusNaturalLevels(L)
    sched = np.linspace(1.5, 1e-8, steps)
    full = qmc.QuantumAnnealReplicas(sched, 1, P, 0.01, n, None, nbs, 2024, color=color, nreplicas=R, device=dev)
    w_full = full["words"].copy()
    e_full = full["energies"].copy()
    again = qmc.QuantumAnnealReplicas(sched, 1, P, 0.01, n, None, nbs, 2024, color=color, nreplicas=R, device=dev)
    assert np.array_equal(again["words"], w_full) and np.array_equal(again["energies"], e_full)
    del again
    half = R // 2
    for r0 in (0, half):
        part = qmc.QuantumAnnealReplicas(sched, 1, P, 0.01, n, None, nbs, 2024, color=color, nreplicas=half,
                                         replica0=r0, device=dev)
        assert np.array_equal(part["words"], w_full[r0:r0 + half])
        assert np.array_equal(part["energies"], e_full[r0:r0 + half])
        del part
    # host recomputation of the energy: bonds (i, right), (i, down), each once
    idx = nbs[:, :, 0].astype(np.int64)
    Jv = nbs[:, :, 1]
    ii = np.repeat(np.arange(n), 4).reshape(n, 4)
    once = idx > ii
    bi, bj, bJ = ii[once], idx[once], Jv[once]
    rng = np.random.RandomState(0)
    for r, k in zip(rng.randint(R, size=6), rng.randint(P, size=6)):
        s = 1.0 - 2.0 * ((w_full[r] >> np.uint64(k)) & np.uint64(1)).astype(np.float64)
        ref = -np.sum(bJ * s[bi] * s[bj])
        np.testing.assert_allclose(e_full[r, k], ref, rtol=1e-12)
    assert e_full.min() > -np.abs(bJ).sum()
    start = qmc.QuantumAnnealReplicas(sched[:0], 1, P, 0.01, n, None, nbs, 2024, color=color, nreplicas=64,
                                      device=dev)
    assert e_full.mean() < start["energies"].mean() - 0.5 * n
    # bit-exact against the oracle at the bench shape (4096 rows: 8 row chunks x 4 passes per block, 511 levels,
    # two sweeps in flight): the Philox key carries the global replica id, so single replicas of the full state
    # can be restated on the CPU in isolation -- the reference's natural-order sequential sweep
    # (piqmc/qmc.pyx:296-357), about 0.3 s per replica.  First / last row of a chunk, of a pass, of the state.
    idx32, J32 = O.nbs_to_ell(nbs)
    natural = np.tile(np.arange(n, dtype=np.int32), (steps, 1))
    for r in (0, 127, 128, 511, 512, 2048, 3583, 4095):
        _check_replica_against_oracle(w_full[r], sched, P, 0.01, idx32, J32, color, 2024, r, natural)
    del full, w_full


def _check_replica_against_oracle(words_r, sched, P, temp, idx32, J32, color, seed, r, orders):
    import piqmc.tools as T
    n = idx32.shape[0]
    init = O.colour_init_spins(seed, r, 1, n)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    O.qa_colour(sched, 1, P, temp, idx32, J32, color, want, seed, replica0=r, orders=orders)
    got = np.transpose(T.UnpackWords(words_r[None, :], P), (0, 2, 1))
    assert np.array_equal(got, want), "replica %d differs from the oracle" % r


@pytest.mark.parametrize("rows,replica0", [(512, 3584), (1024, 1024)])
def test_config5_shard_shapes_bit_exact_vs_oracle(dev, rows, replica0):
    """The per-GPU shapes of the 8- and 4-GPU runs of BASELINE configs[4] (512 / 1024 rows of the 256x256, P = 64
    state: 4 row chunks of one / two passes per block), with the shard's global replica offset: sampled replicas
    against the oracle's sequential natural-order sweep, bit for bit (piqmc/qmc.pyx:296-357)."""
    import piqmc.qmc as qmc
    import piqmc.tools as T
    L, P, steps = 256, 64, 6
    n = L * L
    nbs, _ = T.GaussianTorusNeighbors(L, 2024)
    color = T.TorusNaturalLevels(L)
    sched = np.linspace(1.5, 1e-8, steps)
    out = qmc.QuantumAnnealReplicas(sched, 1, P, 0.01, n, None, nbs, 2024, color=color, nreplicas=rows,
                                    replica0=replica0, device=dev)
    idx32, J32 = O.nbs_to_ell(nbs)
    natural = np.tile(np.arange(n, dtype=np.int32), (steps, 1))
    for k in (0, 127, 128, rows // 2 + 1, rows - 1):
        _check_replica_against_oracle(out["words"][k], sched, P, 0.01, idx32, J32, color, 2024, replica0 + k, natural)


# ------------------------------------------------------------------- several replicas per word
@pytest.mark.parametrize("P,R,T", [(20, 100, 0.01), (20, 97, 0.3), (16, 130, 0.05), (8, 300, 0.5), (4, 520, 0.2),
                                   (32, 70, 0.3), (12, 161, 1.0)])
def test_replicas_per_word_bit_exact(dev, P, R, T):
    """P <= 32 slices: floor(64/P) replicas share a word (3 at P = 20).  The result is, replica by
    replica, the one of the one-replica-per-word layout and of the CPU statement -- same Philox keys
    -- for cold and hot (many draws) anneals, replica counts that do not fill the last word, a
    non-zero first replica id, and both ways of starting (Philox start, uploaded spins)."""
    import piqmc.qmc as qmc
    nbs, idx, J32, color = _torus(8, 21)
    n = 64
    sched = np.linspace(1.5, 1e-8, 7)
    seed, r0 = 77 + P, 5
    init = O.colour_init_spins(seed, r0, R, n)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    O.qa_colour(sched, 2, P, T, idx, J32, color, want, seed, replica0=r0)
    dev.set_variant(2)
    try:
        one = qmc.QuantumAnnealReplicas(sched, 2, P, T, n, None, nbs, seed, color=color, nreplicas=R, replica0=r0,
                                        device=dev, per_word=1)
        many = qmc.QuantumAnnealReplicas(sched, 2, P, T, n, None, nbs, seed, color=color, nreplicas=R, replica0=r0,
                                         device=dev)
        up = qmc.QuantumAnnealReplicas(sched, 2, P, T, n, init, nbs, seed, color=color, replica0=r0, device=dev)
    finally:
        dev.set_variant(0)
    assert one["per_word"] == 1 and many["per_word"] == 64 // P and up["per_word"] == 64 // P
    got = np.transpose(tools.UnpackWords(many["words"], P), (0, 2, 1))
    assert np.array_equal(got, want)
    assert np.array_equal(many["words"], one["words"]) and np.array_equal(up["words"], one["words"])
    assert np.array_equal(many["energies"], one["energies"]) and np.array_equal(up["energies"], one["energies"])


def test_path_graph_natural_order_exceeds_one_grid(dev):
    """A path graph in natural order has N levels and a level gap of 1: the period-major order of the
    dataflow kernel would need (1 + N/2) * N units -- more than one grid holds at N = 65536.  The call
    must not truncate the grid silently: it runs class by class and still equals the sequential sweep."""
    n, R, P = 65536, 3, 8
    rng = np.random.RandomState(5)
    nbs = np.zeros((n, 2, 2))
    Jb = rng.randn(n - 1)
    for i in range(n - 1):
        nbs[i, 1] = (i + 1, Jb[i])
        nbs[i + 1, 0] = (i, Jb[i])
    nbs[0, 0] = (0, 0.0)
    nbs[n - 1, 1] = (n - 1, 0.0)
    idx, J32 = O.nbs_to_ell(nbs)
    color = tools.OrderLevels(nbs)
    assert color.max() == n - 1
    sched = np.array([0.7])
    init = O.colour_init_spins(3, 0, R, n)
    want = np.repeat(init[:, :, None], P, axis=2).copy()
    O.qa_colour(sched, 1, P, 0.3, idx, J32, color, want, 3, 0, 0, 0)
    dev.set_graph(nbs, color)
    dev.set_variant(2)
    try:
        dev.state_alloc(R, P)
        dev.state_init_random(3, 0, tile=True)
        l0 = dev.launch_count
        dev.qa_colour(sched, 1, 0.3, 3)
        assert dev.launch_count - l0 == n                      # one launch per class
        got = np.transpose(tools.UnpackWords(dev.state_download_words(), P), (0, 2, 1))
    finally:
        dev.set_variant(0)
    assert np.array_equal(want, got)


def _integer_instance(n, seed, unit=0.25):
    """Random graph with couplings and fields in {+-1, +-2, +-3} * unit, at most 6 bonds + 1 field per spin."""
    rng = np.random.RandomState(seed)
    rows = [[] for _ in range(n)]
    for _ in range(3 * n):
        i, j = rng.randint(n, size=2)
        if i == j or len(rows[i]) >= 6 or len(rows[j]) >= 6 or any(e[0] == j for e in rows[i]):
            continue
        v = unit * rng.choice([-3, -2, -1, 1, 2, 3])
        rows[i].append((j, v))
        rows[j].append((i, v))
    nbs = np.zeros((n, 7, 2))
    for i in range(n):
        ent = rows[i] + ([(i, unit * rng.choice([-2, -1, 1, 3]))] if rng.rand() < 0.7 else [])
        for k, (j, v) in enumerate(ent):
            nbs[i, k] = (j, v)
    return nbs


@pytest.mark.parametrize("kind,P,T,R", [("qa", 7, 0.3, 70), ("qa", 64, 0.05, 33), ("qa", 20, 2.0, 40),
                                        ("sa", 64, None, 200), ("sa", 64, None, 4100)])
def test_resident_integer_kernel_bit_exact(dev, kind, P, T, R):
    """Couplings that are small integer multiples of a power of two run through the bit-sliced resident kernel
    (n in five bit planes, thresholds per (sweep, spin) by the warp): equal to the CPU statement, and to the
    per-lane resident kernel (PIQMC_NO_INT_KERNEL), cold and hot, static colouring and per-sweep orders."""
    import piqmc.sa as sa
    n = 24
    nbs = _integer_instance(n, 11 + P)
    idx, J32 = O.nbs_to_ell(nbs)
    color = tools.OrderLevels(nbs)
    prng = np.random.RandomState(R)
    for orders in (None, np.stack([prng.permutation(n) for _ in range(12)]).astype(np.int32)):
        got = []
        for env in ({"PIQMC_FORCE_INT_KERNEL": "1"}, {}, {"PIQMC_NO_INT_KERNEL": "1"}):   # bit-sliced | hot head per lane, cold tail bit-sliced | per lane
            os.environ.update(env)
            try:
                dev.set_graph(nbs, color)
                dev.set_variant(5)
                if kind == "qa":
                    sched = np.linspace(2.0, 1e-8, 6)
                    init = O.colour_init_spins(3, 1, R, n)
                    want = np.repeat(init[:, :, None], P, axis=2).copy()
                    O.qa_colour(sched, 2, P, T, idx, J32, color, want, 3, replica0=1, sweep0=2, orders=orders)
                    dev.state_alloc(R, P)
                    dev.state_init_random(3, 1, tile=True)
                    dev.qa_colour(sched, 2, T, 3, replica0=1, sweep0=2, orders=orders)
                    got.append(np.transpose(tools.UnpackWords(dev.state_download_words(), P), (0, 2, 1)))
                else:
                    sched = np.linspace(3.0, 0.05, 6)
                    init = (2 * np.random.RandomState(7).randint(2, size=(R, n)) - 1).astype(np.int8)
                    want = init.copy()
                    O.sa_colour(sched, 2, idx, J32, color, want, seed=5, row0=1, orders=orders)
                    out = sa.AnnealReplicas(sched, 2, init, nbs, 5, color=color if orders is None else None,
                                            order=orders if orders is not None else "natural", row0=1, device=dev)
                    got.append(out["spins"])
            finally:
                dev.set_variant(0)
                for k in env:
                    del os.environ[k]
            assert np.array_equal(got[-1], want), "env %r" % (env,)


def test_energy_histogram_on_device(golden, dev):
    """piqmc_energy_histogram == numpy on the downloaded energies, for the three reductions."""
    import piqmc.qmc as qmc
    nbs = golden["vec"]["nbs_inst_0_32x32"]
    out = qmc.QuantumAnnealReplicas(np.linspace(1.5, 1e-8, 20), 1, 20, 0.01, 1024, None, nbs, 5, order="natural",
                                    nreplicas=301, device=dev)
    en = out["energies"]
    gs, n = -1591.9166416866, 1024
    for reduce, val in (("mean", en.mean(axis=1)), ("min", en.min(axis=1)), ("all", en.reshape(-1))):
        v = (val - gs) / n
        lo, hi = 0.2, 0.3
        h = dev.energy_histogram(e0=gs, scale=1.0 / n, lo=lo, hi=hi, nbins=25, reduce=reduce)
        ref = np.histogram(v[(v >= lo) & (v < hi)], bins=25, range=(lo, hi))[0]
        assert np.array_equal(h["counts"], ref.astype(np.uint64))
        assert h["below"] == int((v < lo).sum()) and h["above"] == int((v >= hi).sum())
        np.testing.assert_allclose([h["mean"], h["min"], h["max"]], [v.mean(), v.min(), v.max()], rtol=1e-12)


# ------------------------------------------------------------ the as-shipped semantics at production speed
@pytest.mark.parametrize("inst,P,T,R,mcsteps", [("inst_0_32x32", 20, 0.01, 5, 1), ("inst_0_32x32", 64, 0.3, 3, 2),
                                               ("inst_0_32x32", 33, 0.05, 4, 1), ("hopfield8", 10, 0.2, 70, 3),
                                               ("boixo", 2, 0.5, 9, 2), ("santoro_80x80", 20, 0.01, 2, 1)])
def test_qa_carry_bit_exact(golden, dev, inst, P, T, R, mcsteps):
    """piqmc_qa_carry (one warp per replica, the slices of a replica in lockstep over the visiting order, slice 1
    first) == the CPU statement of the as-shipped loop nest (slices outermost, energy difference carried over a
    slice sweep, piqmc/qmc.pyx:98-136), for permutation orders, the natural order, non-zero replica and sweep
    offsets, P below and above 32."""
    nbs, idx, J32, color = _graph(golden, inst)
    n = NSPINS[inst]
    sched = np.linspace(1.5, 1e-8, 5)
    prng = np.random.RandomState(P)
    for orders in (np.stack([prng.permutation(n) for _ in range(5 * mcsteps)]).astype(np.int32), None):
        init = O.colour_init_spins(21, 6, R, n)
        want = np.repeat(init[:, :, None], P, axis=2).copy()
        O.qa_carry(sched, mcsteps, P, T, idx, J32, want, 21, replica0=6, sweep0=9, orders=orders)
        dev.set_graph(nbs, color)
        dev.state_alloc(R, P)
        for env in ({"PIQMC_CARRY_RESIDENT": "1"}, {"PIQMC_CARRY_GLOBAL": "1"}):   # words in shared memory | read through L2
            os.environ.update(env)
            try:
                dev.state_init_random(21, 6, tile=True)
                dev.qa_carry(sched, mcsteps, T, 21, replica0=6, sweep0=9, orders=orders)
            finally:
                for k in env:
                    del os.environ[k]
            got = np.transpose(tools.UnpackWords(dev.state_download_words(), P), (0, 2, 1))
            assert np.array_equal(want, got)


def test_qa_carry_residual_energy_distribution_vs_reference(golden, dev):
    """Config 2 with the reference's DEFAULT function: 1024 runs of qmc.QuantumAnneal (as shipped: the energy
    difference carried over a slice sweep; golden qa_100, made by the compiled reference) against 1024 replicas
    of QuantumAnnealReplicas(semantics="reference", order="permutation") -- in 16 calls of 64 replicas with
    their own seeds, because the replicas of one call share their permutations (every reference run draws its
    own) and the residual depends on them a little.  KS p > 0.01 on the slice-averaged and the best-slice
    residual energy per spin; and the carry statistics are far from the per-spin-reset ones (about 1.0 against
    0.25 per spin)."""
    import piqmc.qmc as qmc
    nbs = golden["vec"]["nbs_inst_0_32x32"]
    ref = golden["dist"]["qa_100"]
    mine = np.concatenate([
        qmc.QuantumAnnealReplicas(np.linspace(1.5, 1e-8, 100), 1, 20, 0.01, 1024, None, nbs, seed=4242 + 977 * b,
                                  order="permutation", semantics="reference", nreplicas=64, replica0=64 * b,
                                  device=dev)["energies"] for b in range(ref.shape[0] // 64)])
    for name, a, b in (("mean", mine.mean(axis=1), ref.mean(axis=1)), ("best", mine.min(axis=1), ref.min(axis=1))):
        p = ks_2samp_p(_residual(a, "inst_0_32x32"), _residual(b, "inst_0_32x32"))
        print("as-shipped semantics, %s slice: residual/spin mine %.4f ref %.4f KS p %.3f"
              % (name, _residual(a.mean(), "inst_0_32x32"), _residual(b.mean(), "inst_0_32x32"), p))
        assert p > 0.01
    assert _residual(mine.mean(), "inst_0_32x32") > 0.6


# ------------------------------------------------------------------- anneal + results in one call
@pytest.mark.parametrize("lag16", [None, 0, 8, 40, 100])
@pytest.mark.parametrize("R,P,L", [(300, 8, 8), (1100, 64, 16)])
def test_qa_colour_results_staggered_chunks(dev, monkeypatch, R, P, L, lag16):
    """piqmc_qa_colour_results: the row chunks of the dataflow launch run staggered (chunk c lags c * lag16 / 16
    sweeps), report to the host when they are final and are downloaded / reduced to energies while the others
    still sweep.  Words and energies must equal the plain sequence qa_colour -> results bit for bit, for any
    stagger (0, a fraction of a sweep, several sweeps, more than the run), and equal the oracle."""
    import piqmc.tools as T
    nbs, idx, J32, checker = _torus(L, 21)
    n = L * L
    steps = 5
    sched = np.linspace(1.5, 1e-8, steps)
    seed = 4242 + R
    color = T.ColourGraph(nbs, "natural")
    dev.set_variant(2)
    try:
        dev.set_graph(nbs, color)
        dev.state_alloc(R, P)
        dev.state_init_random(seed, 3, tile=True)
        dev.qa_colour(sched, 1, 0.05, seed, replica0=3)
        e_plain, w_plain = dev.results()
        w_plain = w_plain.copy()
        if lag16 is None:                                  # the library's own choice: the plain sequence (the stagger is opt-in)
            monkeypatch.delenv("PIQMC_PIPE_LAG16", raising=False)
        else:
            monkeypatch.setenv("PIQMC_PIPE_LAG16", str(lag16))
        dev.state_init_random(seed, 3, tile=True)
        runs0 = dev.pipelined_runs
        from piqmc import device as D
        buf = D.pinned_empty((n, R), np.uint64)
        e_pipe, w_pipe = dev.qa_colour_results(sched, 1, 0.05, seed, replica0=3, words_out=buf)
        assert dev.pipelined_runs == runs0 + (0 if lag16 is None else 1), "the overlapped path was (not) taken"
        runs0 = dev.pipelined_runs - 1
        assert np.array_equal(w_pipe, w_plain)
        assert np.array_equal(e_pipe, e_plain)
        # and with the switch off: the same call runs the steps one after the other
        monkeypatch.setenv("PIQMC_PIPE", "0")
        dev.state_init_random(seed, 3, tile=True)
        e_seq, w_seq = dev.qa_colour_results(sched, 1, 0.05, seed, replica0=3)
        assert dev.pipelined_runs == runs0 + 1
        assert np.array_equal(w_seq, w_plain) and np.array_equal(e_seq, e_plain)
    finally:
        dev.set_variant(0)
    if lag16 in (None, 40):
        for r in (0, R // 2, R - 1):
            init = O.colour_init_spins(seed, 3 + r, 1, n)
            want = np.repeat(init[:, :, None], P, axis=2).copy()
            O.qa_colour(sched, 1, P, 0.05, idx, J32, checker, want, seed, replica0=3 + r,
                        orders=np.tile(np.arange(n, dtype=np.int32), (steps, 1)))
            got = np.transpose(T.UnpackWords(w_plain[r][None, :], P), (0, 2, 1))
            assert np.array_equal(got, want)


# ------------------------------------------------------------------- SA at the size of tools/bench_sa.py
@pytest.mark.parametrize("hi,lo", [(3.0, 1.0), (0.3, 0.01)])
def test_sa_bench_shape_bit_exact_vs_oracle(dev, hi, lo):
    """The production SA path at the shape tools/bench_sa.py measures (256x256 Gaussian torus, 1024 rows = 65 536
    replicas, natural-order levels), hot (about half of all lanes draw a uniform: the per-thread draw path with
    the transposed pattern lookup, full warps) and cold: sampled word rows (64 replicas each) against the oracle's
    sa.Anneal rules (piqmc/sa.pyx:95-120), bit for bit.  The oracle starts from the downloaded initial rows."""
    import piqmc.tools as T
    L, rows, steps = 256, 1024, 4
    n = L * L
    nbs, _ = T.GaussianTorusNeighbors(L, 2024)
    color = T.TorusNaturalLevels(L)
    idx32, J32 = O.nbs_to_ell(nbs)
    sched = np.linspace(hi, lo, steps)
    dev.set_graph(nbs, color)
    dev.state_alloc(rows, 64)
    dev.state_init_random(99, 0, tile=False)
    start = dev.state_download_words().copy()                      # [rows, n]
    dev.sa_colour(sched, 1, 99, row0=0)
    final = dev.state_download_words()
    lanes = np.arange(64, dtype=np.uint64)
    natural = np.tile(np.arange(n, dtype=np.int32), (steps, 1))
    for r in (0, 255, 256, 1023):
        unpack = lambda w: (1 - 2 * ((w[None, :] >> lanes[:, None]) & np.uint64(1)).astype(np.int8)).astype(np.int8)
        want = np.ascontiguousarray(unpack(start[r]))              # [64 replicas, n]
        O.sa_colour(sched, 1, idx32, J32, color, want, seed=99, row0=r, orders=natural)
        assert np.array_equal(unpack(final[r]), want), "row %d differs from the oracle" % r
    assert not np.array_equal(start[0], final[0])
