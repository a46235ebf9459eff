"""CPU tests: the oracle (oracle/piqmc_oracle.c) against the reference's golden vectors.

The goldens in tests/golden/ were produced by the reference's own compiled Cython modules
(tests/golden/make_golden.py).  These tests pin the C restatement to them bit for bit and restate
the reference's own test-suite (testing/test_core.py, testing/test_boixo.py) on the oracle.
"""
import ctypes
import os

import numpy as np
import pytest

from oracle import oracle as O
from helpers import NSPINS, case_inputs, cases


def _ids(kinds):
    import json, os
    v = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz"))
    return [c["name"] for c in json.loads(str(v["cases_json"])) if c["kind"] in kinds]


# ----------------------------------------------------------------------------- third-party streams
def test_glibc_rand_restatement_matches_live_libc():
    libc = ctypes.CDLL("libc.so.6")
    for seed in (0, 1, 5, 123, 2 ** 31 + 7, 4294967295):
        libc.srand(seed)
        live = [libc.rand() for _ in range(2000)]
        g = O.glibc_state(seed)
        mine = [O.lib().oracle_glibc_rand(ctypes.byref(g)) for _ in range(2000)]
        assert live == mine


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert [hex(x) for x in O.philox([0, 0, 0, 0], [0, 0])] == \
        ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in O.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == \
        ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in O.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344],
                                     [0xA4093822, 0x299F31D0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_colour_threshold_is_exp():
    assert O.colour_thresh(0.0) == 0xFFFFFFFF
    xs = np.linspace(-21.9, -1e-3, 2000).astype(np.float32)
    got = np.array([O.colour_thresh(float(x)) for x in xs], dtype=np.float64) / 2.0 ** 32
    want = np.exp(xs.astype(np.float64))
    assert np.all(np.abs(got - want) <= 5e-6 * want + 2.0 ** -32)      # floor() costs up to one unit
    mono = np.array([O.colour_thresh(float(x)) for x in np.linspace(-22, 0, 5000).astype(np.float32)])
    assert np.all(np.diff(mono.astype(np.int64)) >= 0)


def test_jperp_matches_formula(golden):
    for gamma in (1.5, 0.5, 1e-3, 1e-8):
        for P, T in ((20, 0.01), (5, 0.01), (64, 0.01), (10, 0.5)):
            j = O.jperp(gamma, P, T)
            t32 = float(np.float32(T))
            pt = float(np.float32(P) * np.float32(T))
            ref = np.float32(-0.5 * P * t32 * np.log(np.tanh(gamma / pt)))
            assert j == ref


# ----------------------------------------------------------------------------- golden replays
@pytest.mark.parametrize("name", _ids(("qa",)))
def test_qa_reference_golden(golden, name):
    vec = golden["vec"]
    case = [c for c in cases(vec) if c["name"] == name][0]
    sched, nbs, rng, init = case_inputs(case, vec)
    n, P = NSPINS[case["inst"]], case["P"]
    confs = np.tile(init, (P, 1)).T                     # F-strided view, like the examples
    perms = O.make_perms(rng, n, sched.size * case["mcsteps"])
    g = O.glibc_state(case["srand_seed"])
    consumed = O.qa_reference(sched, case["mcsteps"], P, case["T"], n, confs, nbs, perms, gstate=g)
    assert np.array_equal(confs.astype(np.int8), vec[name + "__out"])
    assert consumed == int(vec[name + "__consumed"])
    assert O.lib().oracle_glibc_rand(ctypes.byref(g)) == int(vec[name + "__libc_next"])
    assert rng.randint(1 << 30) == int(vec[name + "__rng_next"])
    J = _J(golden, case["inst"])
    en = np.array([O.energy(J, confs[:, k]) for k in range(P)])
    np.testing.assert_allclose(en, vec[name + "__energy"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", _ids(("sa",)))
def test_sa_reference_golden(golden, name):
    vec = golden["vec"]
    case = [c for c in cases(vec) if c["name"] == name][0]
    sched, nbs, rng, init = case_inputs(case, vec)
    n = NSPINS[case["inst"]]
    sv = init.copy()
    perms = O.make_perms(rng, n, sched.size * case["mcsteps"])
    g = O.glibc_state(case["srand_seed"])
    consumed = O.sa_reference(sched, case["mcsteps"], sv, nbs, perms, gstate=g)
    assert np.array_equal(sv.astype(np.int8), vec[name + "__out"])
    assert consumed == int(vec[name + "__consumed"])
    assert O.lib().oracle_glibc_rand(ctypes.byref(g)) == int(vec[name + "__libc_next"])
    assert rng.randint(1 << 30) == int(vec[name + "__rng_next"])


def _dense_cases(kind):
    import json
    vec = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_dense.npz"))
    return vec, [c for c in json.loads(str(vec["cases_json"])) if c["kind"] == kind]


@pytest.mark.parametrize("k", range(7))
def test_qa_dense_golden(k):
    """oracle_qa_dense == the reference's qmc.QuantumAnneal_dense (qmc.pyx:141-242): spins and the
    positions both random streams are left at."""
    vec, cs = _dense_cases("qa_dense")
    case = cs[k]
    J = vec["J_" + case["inst"]]
    n, P = J.shape[0], case["P"]
    rng = np.random.RandomState(case["rng_seed"])
    init = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
    assert np.array_equal(init.astype(np.int8), vec[case["name"] + "__init"])
    confs = np.tile(init, (P, 1)).T
    sched = np.linspace(case["sched"][0], case["sched"][1], int(case["sched"][2]))
    perms = O.make_perms(rng, n, sched.size * case["mcsteps"])
    g = O.glibc_state(case["srand_seed"])
    O.qa_dense(sched, case["mcsteps"], P, case["T"], n, confs, J, perms, gstate=g)
    assert np.array_equal(confs.astype(np.int8), vec[case["name"] + "__final"])
    assert [O.lib().oracle_glibc_rand(ctypes.byref(g)) for _ in range(4)] == list(vec[case["name"] + "__libc_next"])
    assert list(rng.randint(1 << 30, size=4)) == list(vec[case["name"] + "__rng_next"])


@pytest.mark.parametrize("k", range(5))
def test_sa_dense_golden(k):
    """oracle_sa_dense == the reference's sa.Anneal_dense (sa.pyx:126-187)."""
    vec, cs = _dense_cases("sa_dense")
    case = cs[k]
    J = vec["J_" + case["inst"]]
    n = J.shape[0]
    rng = np.random.RandomState(case["rng_seed"])
    sv = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
    sched = np.linspace(case["sched"][0], case["sched"][1], int(case["sched"][2]))
    perms = O.make_perms(rng, n, sched.size * case["mcsteps"])
    g = O.glibc_state(case["srand_seed"])
    O.sa_dense(sched, case["mcsteps"], sv, J, perms, gstate=g)
    assert np.array_equal(sv.astype(np.int8), vec[case["name"] + "__final"])
    assert [O.lib().oracle_glibc_rand(ctypes.byref(g)) for _ in range(4)] == list(vec[case["name"] + "__libc_next"])
    assert list(rng.randint(1 << 30, size=4)) == list(vec[case["name"] + "__rng_next"])


def test_parallel1_variants_golden(golden):
    vec = golden["vec"]
    for case in cases(vec, ("qa_par", "sa_par")):
        sched, nbs, rng, init = case_inputs(case, vec)
        n = NSPINS[case["inst"]]
        g = O.glibc_state(case["srand_seed"])
        if case["kind"] == "qa_par":
            c = np.tile(init, (case["P"], 1)).T.copy()
            cnt = O.qa_parallel1(sched, case["mcsteps"], case["P"], case["T"], n, c, nbs, gstate=g)
        else:
            c = init.copy()
            cnt = O.sa_parallel1(sched, case["mcsteps"], c, nbs, gstate=g)
        assert np.array_equal(c.astype(np.int8), vec[case["name"] + "__out"])
        assert cnt == int(vec[case["name"] + "__consumed"])


def multispin_streams(rng, n, nsweeps):
    """perms int32[nsweeps,N] and rands float64[nsweeps*N,64] in the reference's draw order
    (sa.pyx:331,336-337,398,400)."""
    carry = rng.rand(64)
    perm = rng.permutation(range(n))
    perms = np.empty((nsweeps, n), dtype=np.int32)
    rands = np.empty((nsweeps, n, 64))
    for s in range(nsweeps):
        perms[s] = perm
        blocks = rng.rand(n * 64).reshape(n, 64)
        rands[s, 0] = carry
        rands[s, 1:] = blocks[:-1]
        carry = blocks[-1]
        perm = rng.permutation(perm)
    return perms, rands.reshape(nsweeps * n, 64)


def test_multispin_golden(golden):
    vec = golden["vec"]
    for case in cases(vec, ("multispin",)):
        sched, nbs, rng, init = case_inputs(case, vec)
        n = NSPINS[case["inst"]]
        perms, rands = multispin_streams(rng, n, sched.size * case["mcsteps"])
        bits = init.copy()
        O.sa_multispin(sched, case["mcsteps"], bits, nbs, perms, rands)
        want = vec[case["name"] + "__out"]
        # the reference's unpack overrun corrupts column 0 of rows 1..63 (sa.pyx:402-405)
        assert np.array_equal(bits[:, 1:].astype(np.int8), want[:, 1:])
        assert bits[0, 0] == want[0, 0]
        assert rng.randint(1 << 30) == int(vec[case["name"] + "__rng_next"])


def _J(golden, inst):
    import scipy.sparse as sps
    ijv = golden["inst"]["inst_" + inst]
    n = NSPINS[inst]
    J = sps.dok_matrix((n, n))
    for i, j, v in ijv:
        J[int(i) - 1, int(j) - 1] = v
    return J


def test_energy_probes(golden):
    vec = golden["vec"]
    for inst in ("boixo", "bipartite8", "hopfield8", "inst_0_32x32"):
        J = _J(golden, inst)
        sv = vec["energy_probe_spins_" + inst].astype(np.float64)
        got = np.array([O.energy(J, s) for s in sv])
        np.testing.assert_allclose(got, vec["energy_probe_" + inst], rtol=1e-12, atol=1e-12)


def test_known_ground_states(golden):
    """Spin-Glass-Server ground states shipped in examples/spinglass32.py:76-79 and
    examples/santoro80.py:41-263, energies in spinglass32.py:33 / santoro_80x80_answer.txt:24."""
    for inst, e_per_spin in (("inst_0_32x32", -1.55460615142578), ("santoro_80x80", -1.58051667679)):
        J = _J(golden, inst)
        gs = golden["inst"]["gs_" + inst].astype(np.float64)
        e = O.energy(J, gs)
        assert abs(e - float(golden["inst"]["gs_energy_" + inst])) < 1e-9
        assert abs(e / NSPINS[inst] - e_per_spin) < 1e-6


# ----------------------------------------------------------------------------- reference test-suite, restated
def test_reference_boixo_suite(golden):
    """testing/test_boixo.py:60-114 on the oracle."""
    J = _J(golden, "boixo")
    nbs = golden["vec"]["nbs_boixo"]
    svecs = np.array([[-1, 1, -1, 1, 1, -1, -1, 1], [1, 1, 1, 1, 1, -1, 1, -1],
                      [1, -1, 1, 1, -1, -1, -1, -1], [1, -1, -1, 1, 1, -1, 1, 1]], dtype=np.float64)
    for v, en in zip(svecs, [4.0, -8.0, -4.0, 0.0]):
        assert O.energy(J, v) == en
    rng = np.random.RandomState(123)
    ctypes.CDLL("libc.so.6").srand(1)
    for _ in range(5):
        sv = np.array([2 * rng.randint(2) - 1 for _ in range(8)], dtype=np.float64)
        sched = np.linspace(1.0, 0.01, 10)
        O.sa_reference(sched, 3, sv, nbs, O.make_perms(rng, 8, 30))
        assert np.sum(sv[:4]) == 4 or np.sum(sv) == -8
        assert O.energy(J, sv) == -8.0
    confs = np.tile(np.array([2 * rng.randint(2) - 1 for _ in range(8)], dtype=np.float64), (5, 1)).T
    sched = np.linspace(0.5, 1e-8, 10)
    O.qa_reference(sched, 3, 5, 0.01, 8, confs, nbs, O.make_perms(rng, 8, 30))
    en = np.array([O.energy(J, confs[:, k]) for k in range(5)])
    assert np.sum(en) == -8.0 * 5
    for k in range(5):
        assert np.sum(confs[:4, k]) == 4.0 or np.sum(confs[:, k]) == -8.0


# ----------------------------------------------------------------------------- colour semantics sanity
def test_colour_oracle_reaches_boixo_ground_state(golden):
    nbs = golden["vec"]["nbs_boixo"]
    idx, J32 = O.nbs_to_ell(nbs)
    import piqmc.tools as T
    color = T.ColourGraph(nbs)
    J = _J(golden, "boixo")
    R, P = 8, 5
    spins = np.repeat(O.colour_init_spins(42, 0, R, 8)[:, :, None], P, axis=2).copy()
    O.qa_colour(np.linspace(0.5, 1e-8, 10), 3, P, 0.01, idx, J32, color, spins, seed=42)
    for r in range(R):
        for k in range(P):
            assert O.energy(J, spins[r, :, k].astype(np.float64)) == -8.0
    s2 = O.colour_init_spins(43, 0, 70, 8)
    O.sa_colour(np.linspace(1.0, 0.01, 10), 3, idx, J32, color, s2, seed=43)
    en = np.array([O.energy(J, s.astype(np.float64)) for s in s2])
    assert np.mean(en == -8.0) > 0.9
