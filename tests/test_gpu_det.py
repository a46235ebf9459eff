"""GPU parity tests, deterministic paths: the CUDA kernels (through the C ABI and the drop-in
Python mirror) against the reference's golden vectors and the oracle.  Bar: spin configurations
bit-exact, consumed-uniform counts and random-stream positions identical, energies <= 1e-12 rel."""
import ctypes

import numpy as np
import pytest

from oracle import oracle as O
from helpers import NSPINS, case_inputs, cases
from test_oracle import _J, _ids, multispin_streams

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", _ids(("qa",)))
def test_quantumanneal_dropin_golden(golden, dev, name):
    """piqmc.qmc.QuantumAnneal called exactly like the reference (F-strided confs view, shared
    rng, process-global libc stream) reproduces the reference's output and stream positions."""
    import piqmc.qmc as qmc
    import piqmc.sa as sa
    vec = golden["vec"]
    case = [c for c in cases(vec) if c["name"] == name][0]
    sched, nbs, rng, init = case_inputs(case, vec)
    n, P = NSPINS[case["inst"]], case["P"]
    confs = np.tile(init, (P, 1)).T
    libc = ctypes.CDLL(None)
    libc.srand(case["srand_seed"])
    assert qmc.QuantumAnneal(sched, case["mcsteps"], P, case["T"], n, confs, nbs, rng) is None
    assert np.array_equal(confs.astype(np.int8), vec[name + "__out"])
    assert libc.rand() == int(vec[name + "__libc_next"])          # libc left where the reference leaves it
    assert rng.randint(1 << 30) == int(vec[name + "__rng_next"])   # and so is the NumPy generator
    J = _J(golden, case["inst"])
    en = np.array([sa.ClassicalIsingEnergy(confs[:, k], J) for k in range(P)])
    np.testing.assert_allclose(en, vec[name + "__energy"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", _ids(("sa",)))
def test_anneal_dropin_golden(golden, dev, name):
    import piqmc.sa as sa
    vec = golden["vec"]
    case = [c for c in cases(vec) if c["name"] == name][0]
    sched, nbs, rng, init = case_inputs(case, vec)
    sv = init.copy()
    libc = ctypes.CDLL(None)
    libc.srand(case["srand_seed"])
    assert sa.Anneal(sched, case["mcsteps"], sv, nbs, rng) is None
    assert np.array_equal(sv.astype(np.int8), vec[name + "__out"])
    assert libc.rand() == int(vec[name + "__libc_next"])
    assert rng.randint(1 << 30) == int(vec[name + "__rng_next"])
    e = sa.ClassicalIsingEnergy(sv, _J(golden, case["inst"]))
    np.testing.assert_allclose(e, vec[name + "__energy"][0], rtol=1e-12, atol=1e-12)


def _dense_cases(kind):
    import json
    import os
    vec = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_dense.npz"))
    return vec, [c for c in json.loads(str(vec["cases_json"])) if c["kind"] == kind]


@pytest.mark.parametrize("k", range(7))
def test_quantumanneal_dense_dropin_golden(dev, k):
    """piqmc.qmc.QuantumAnneal_dense called like the reference's (qmc.pyx:141-242; F-strided confs
    view, shared rng, process-global libc stream) reproduces the reference's output and leaves
    both random streams where the reference leaves them."""
    import piqmc.qmc as qmc
    vec, cs = _dense_cases("qa_dense")
    case = cs[k]
    J = vec["J_" + case["inst"]]
    n, P = J.shape[0], case["P"]
    rng = np.random.RandomState(case["rng_seed"])
    init = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
    confs = np.tile(init, (P, 1)).T
    sched = np.linspace(case["sched"][0], case["sched"][1], int(case["sched"][2]))
    libc = ctypes.CDLL(None)
    libc.srand(case["srand_seed"])
    assert qmc.QuantumAnneal_dense(sched, case["mcsteps"], P, case["T"], n, confs, J, rng) is None
    assert np.array_equal(confs.astype(np.int8), vec[case["name"] + "__final"])
    assert [libc.rand() for _ in range(4)] == list(vec[case["name"] + "__libc_next"])
    assert list(rng.randint(1 << 30, size=4)) == list(vec[case["name"] + "__rng_next"])


@pytest.mark.parametrize("k", range(5))
def test_anneal_dense_dropin_golden(dev, k):
    """piqmc.sa.Anneal_dense == the reference's sa.Anneal_dense (sa.pyx:126-187)."""
    import piqmc.sa as sa
    vec, cs = _dense_cases("sa_dense")
    case = cs[k]
    J = vec["J_" + case["inst"]]
    n = J.shape[0]
    rng = np.random.RandomState(case["rng_seed"])
    sv = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
    sched = np.linspace(case["sched"][0], case["sched"][1], int(case["sched"][2]))
    libc = ctypes.CDLL(None)
    libc.srand(case["srand_seed"])
    assert sa.Anneal_dense(sched, case["mcsteps"], sv, J, rng) is None
    assert np.array_equal(sv.astype(np.int8), vec[case["name"] + "__final"])
    assert [libc.rand() for _ in range(4)] == list(vec[case["name"] + "__libc_next"])
    assert list(rng.randint(1 << 30, size=4)) == list(vec[case["name"] + "__rng_next"])


def test_dense_equals_sparse_on_float32_couplings(golden, dev):
    """With couplings exactly representable in float32 the dense and the sparse QA replays agree
    (they differ only in where the coupling is narrowed), replica by replica in one launch."""
    vec = golden["vec"]
    nbs = vec["nbs_boixo16"]
    n, P, R = 16, 6, 5
    J = np.zeros((n, n))
    for i in range(n):
        for j, v in nbs[i]:
            j = int(j)
            if v != 0.0:
                J[min(i, j), max(i, j)] = np.float32(v)
    sched = np.linspace(1.0, 1e-8, 7)
    rng0 = np.random.RandomState(3)
    spins = (2 * rng0.randint(2, size=(R, n, 1)) - 1).astype(np.int8).repeat(P, axis=2)
    perms = np.stack([O.make_perms(np.random.RandomState(r), n, sched.size * 2) for r in range(R)])
    from piqmc import device
    a, b = spins.copy(), spins.copy()
    dev.set_graph(nbs)
    ca = dev.qa_det(sched, 2, P, 0.05, a, perms, rstates=device.rand_states(range(R)))
    cb = dev.qa_dense_det(sched, 2, P, 0.05, b, J, perms, rstates=device.rand_states(range(R)))
    assert np.array_equal(a, b) and np.array_equal(ca, cb)


def test_qa_batch_matches_oracle_per_replica(golden, dev):
    """R replicas in one launch (replica r: RandomState(r), srand(r)) == R oracle runs."""
    import piqmc.qmc as qmc
    vec = golden["vec"]
    nbs = vec["nbs_inst_0_32x32"]
    n, P, T, R = 1024, 20, 0.01, 48
    sched = np.linspace(1.5, 1e-8, 12)
    inits, want, cons = [], [], []
    for r in range(R):
        rng = np.random.RandomState(r)
        sv = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
        c = np.tile(sv, (P, 1)).T.copy()
        g = O.glibc_state(r)
        cons.append(O.qa_reference(sched, 1, P, T, n, c, nbs, O.make_perms(rng, n, sched.size), gstate=g))
        inits.append(np.tile(sv, (P, 1)).T)
        want.append(c.astype(np.int8))
    rngs = []
    for r in range(R):
        rng = np.random.RandomState(r)
        [rng.randint(2) for _ in range(n)]
        rngs.append(rng)
    got, consumed = qmc.QuantumAnnealBatch(sched, 1, P, T, n, np.array(inits), nbs, rngs, list(range(R)))
    assert np.array_equal(got, np.array(want))
    assert np.array_equal(consumed, np.array(cons, dtype=np.uint64))
    # golden replicas 0 and 1 of config 2 start the same way: first 12 steps differ in schedule, so
    # only check against the oracle here; the full 100-step goldens are covered above.


def test_sa_batch_matches_oracle_per_replica(golden, dev):
    import piqmc.sa as sa
    vec = golden["vec"]
    nbs = vec["nbs_santoro_80x80"]
    n, R = 6400, 40
    sched = np.linspace(3.0, 0.01, 6)
    inits, want, cons, rngs = [], [], [], []
    for r in range(R):
        rng = np.random.RandomState(100 + r)
        sv = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
        inits.append(sv.copy())
        rng2 = np.random.RandomState(100 + r)
        [rng2.randint(2) for _ in range(n)]
        rngs.append(rng2)
        g = O.glibc_state(100 + r)
        cons.append(O.sa_reference(sched, 2, sv, nbs, O.make_perms(rng, n, 12), gstate=g))
        want.append(sv.astype(np.int8))
    got, consumed = sa.AnnealBatch(sched, 2, np.array(inits), nbs, rngs, [100 + r for r in range(R)])
    assert np.array_equal(got, np.array(want))
    assert np.array_equal(consumed, np.array(cons, dtype=np.uint64))


def test_uniform_table_source_and_exhaustion(golden, dev):
    """The C ABI also accepts an explicit uniform table; consumption is lazy and reported."""
    vec = golden["vec"]
    case = [c for c in cases(vec) if c["name"] == "boixo_qa_p5"][0]
    sched, nbs, rng, init = case_inputs(case, vec)
    perms = O.make_perms(rng, 8, 30)
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(case["srand_seed"])
    table = np.array([libc.rand() / 2147483647.0 for _ in range(8 * 5 * 30)])
    spins = np.ascontiguousarray(np.tile(init, (5, 1)).T.astype(np.int8))[None].copy()
    dev.set_graph(nbs)
    consumed = dev.qa_det(sched, 3, 5, 0.01, spins, perms[None].copy(), uniforms=table[None])
    assert np.array_equal(spins[0], vec["boixo_qa_p5__out"])
    assert int(consumed[0]) == int(vec["boixo_qa_p5__consumed"])


def test_multispin_dropin_golden(golden, dev):
    import piqmc.sa as sa
    vec = golden["vec"]
    for case in cases(vec, ("multispin",)):
        sched, nbs, rng, init = case_inputs(case, vec)
        bits = init.copy()
        assert sa.Anneal_multispin(sched, case["mcsteps"], bits, nbs, rng) is None
        want = vec[case["name"] + "__out"]
        # column 0 of rows 1..63 of the reference's output is corrupted by its unpack overrun
        assert np.array_equal(bits[:, 1:].astype(np.int8), want[:, 1:])
        assert bits[0, 0] == want[0, 0]
        assert rng.randint(1 << 30) == int(vec[case["name"] + "__rng_next"])
        # and the oracle agrees on the column the reference corrupts
        sched, nbs, rng, init = case_inputs(case, vec)
        perms, rands = multispin_streams(rng, NSPINS[case["inst"]], sched.size * case["mcsteps"])
        ob = init.copy()
        O.sa_multispin(sched, case["mcsteps"], ob, nbs, perms, rands)
        assert np.array_equal(bits, ob)


def test_zero_division_and_errors(golden, dev):
    import piqmc.qmc as qmc
    nbs = golden["vec"]["nbs_boixo"]
    confs = np.ones((8, 5))
    with pytest.raises(ZeroDivisionError):              # piqmc/qmc.c:2065-2074
        qmc.QuantumAnneal(np.linspace(0.5, 0.1, 3), 1, 5, 0.0, 8, confs, nbs, np.random.RandomState(0))
    with pytest.raises(ValueError):
        qmc.QuantumAnneal(np.linspace(0.5, 0.1, 3), 1, 5, 0.01, 9, np.ones((9, 5)), nbs, np.random.RandomState(0))
    with pytest.raises(ValueError):
        qmc.QuantumAnneal(np.linspace(0.5, 0.1, 3), 1, 5, 0.01, 8, 0.5 * confs, nbs, np.random.RandomState(0))


def test_energy_known_answers(golden, dev):
    """testing/test_boixo.py:60-70 and the Spin-Glass-Server ground states."""
    import piqmc.sa as sa
    J = _J(golden, "boixo")
    svecs = np.array([[-1, 1, -1, 1, 1, -1, -1, 1], [1, 1, 1, 1, 1, -1, 1, -1],
                      [1, -1, 1, 1, -1, -1, -1, -1], [1, -1, -1, 1, 1, -1, 1, 1]])
    for v, en in zip(svecs, [4.0, -8.0, -4.0, 0.0]):
        assert sa.ClassicalIsingEnergy(v, J) == en
    for inst in ("inst_0_32x32", "santoro_80x80"):
        e = sa.ClassicalIsingEnergy(golden["inst"]["gs_" + inst], _J(golden, inst))
        np.testing.assert_allclose(e, float(golden["inst"]["gs_energy_cie_" + inst]), rtol=1e-12)
    vec = golden["vec"]
    for inst in ("boixo", "bipartite8", "hopfield8", "inst_0_32x32"):
        J = _J(golden, inst)
        got = [sa.ClassicalIsingEnergy(s, J) for s in vec["energy_probe_spins_" + inst]]
        np.testing.assert_allclose(got, vec["energy_probe_" + inst], rtol=1e-12, atol=1e-12)


# ------------------------------------------------------------------- on-chip replay kernels
@pytest.mark.parametrize("inst,P,T,hot", [("inst_0_32x32", 20, 0.01, False), ("boixo16", 6, 0.5, True),
                                          ("santoro_80x80", 20, 0.3, True)])
def test_onchip_replay_equals_thread_per_replica(golden, dev, monkeypatch, inst, P, T, hot):
    """The deterministic replays run one warp per replica with the replica in shared memory whenever it fits
    (det_onchip_kernel); PIQMC_DET_ONCHIP=0 keeps the round-1 kernels (one thread per replica, global memory).
    Same spins, same number of uniforms consumed, same libc generator state afterwards -- QA and SA, cold
    (few draws) and hot (a draw at most attempts), with the libc generator and with a uniform table."""
    from piqmc import device
    vec = golden["vec"]
    nbs = vec["nbs_" + inst]
    n, R, steps = nbs.shape[0], 5, 4
    sched = np.linspace(1.5, 1e-8, steps)
    rng0 = np.random.RandomState(11)
    spins = (2 * rng0.randint(2, size=(R, n, 1)) - 1).astype(np.int8).repeat(P, axis=2)
    perms = np.stack([O.make_perms(np.random.RandomState(100 + r), n, steps * 2) for r in range(R)])
    dev.set_graph(nbs)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("PIQMC_DET_ONCHIP", mode)
        a = spins.copy()
        st = device.rand_states(range(R))
        ca = dev.qa_det(sched, 2, P, T, a, perms, rstates=st)
        b = spins[:, :, 0].copy()
        st2 = device.rand_states(range(R))
        temps = np.linspace(3.0 if hot else 0.3, 0.05, steps)
        cb = dev.sa_det(temps, 2, b, perms, rstates=st2)
        u = np.random.RandomState(5).rand(R, 2 * steps * n * P)
        c = spins.copy()
        cc = dev.qa_det(sched, 2, P, T, c, perms, uniforms=u)
        out[mode] = (a, np.array(ca), np.frombuffer(bytes(st), dtype=np.uint8).copy(), b, np.array(cb),
                     np.frombuffer(bytes(st2), dtype=np.uint8).copy(), c, np.array(cc))
    for x, y in zip(out["1"], out["0"]):
        assert np.array_equal(x, y)
    assert out["1"][1].sum() > 0 and (not hot or out["1"][4].sum() > n)     # draws did happen
