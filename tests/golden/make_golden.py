#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the REFERENCE ITSELF.

Runs only in the build container: it imports the reference's own Cython modules compiled by
oracle/build_ref.py (oracle/_ref/piqmc_ref) and reads the instance files under
/root/reference/examples/ising_instances.  Nothing at test time reads /root/reference: the
instances (data, not code) and the reference's outputs are stored in the .npz files here.

    python tests/golden/make_golden.py            # vectors (fast, ~1 min)
    python tests/golden/make_golden.py --dist     # + residual-energy distributions (~10 min)

Seeding convention: np.random.RandomState(rng_seed) for the spin
order / initial state, libc srand(srand_seed) for the Metropolis uniforms.
"""
import ctypes
import json
import multiprocessing as mp
import os
import re
import sys

import numpy as np
import scipy.sparse as sps

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"
INST = os.path.join(REF, "examples", "ising_instances")
libc = ctypes.CDLL("libc.so.6")


def load_ijv(name):
    return np.loadtxt(os.path.join(INST, name + ".txt"))


def dok_from_ijv(ijv, n):
    J = sps.dok_matrix((n, n))
    for i, j, v in ijv:
        J[int(i) - 1, int(j) - 1] = v
    return J


def hopfield8_ijv():
    """examples/hopfield8.py:40-45,76-81 rebuilt with NumPy (the script's scipy aliases are gone)."""
    mem = np.array([[-1, -1, -1, -1, 1, 1, 1, 1],
                    [-1, 1, -1, 1, -1, 1, -1, 1],
                    [-1, -1, 1, 1, -1, -1, 1, 1]], dtype=np.float64).T
    vinput = np.array([-1, -1, -1, 1, 1, 1, 1, 1], dtype=np.float64)
    Jd = np.triu(mem @ np.linalg.pinv(mem)) + 0.024 * np.diag(vinput)
    rows = []
    for i in range(8):
        for j in range(8):
            if Jd[i, j] != 0.0:
                rows.append((i + 1, j + 1, Jd[i, j]))
    return np.array(rows)


def ground_state(script, n):
    src = open(os.path.join(REF, "examples", script)).read()
    m = re.search(r"gsspinups\s*=\s*np\.array\(\s*\[([0-9,\s]*)\]", src)
    ups = np.array([int(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()]) - 1
    gs = -np.ones(n, dtype=np.int8)
    gs[ups] = 1
    return gs


def instances():
    out = {}
    for name, n in (("boixo", 8), ("boixo16", 16), ("bipartite8", 8), ("inst_0_32x32", 1024),
                    ("santoro_80x80", 6400)):
        out[name] = (load_ijv(name), n)
    out["hopfield8"] = (hopfield8_ijv(), 8)
    return out


MAXNB = {"boixo": 4, "boixo16": 6, "bipartite8": 5, "inst_0_32x32": 4, "santoro_80x80": 4,
         "hopfield8": 8}

CASES = [
    # the reference's own test configuration (testing/test_boixo.py:22-31,90-105)
    dict(name="boixo_qa_p5", kind="qa", inst="boixo", P=5, T=0.01, sched=[0.5, 1e-8, 10], mcsteps=3,
         rng_seed=123, srand_seed=1),
    # BASELINE.json configs[0] quotes P=20
    dict(name="boixo_qa_p20", kind="qa", inst="boixo", P=20, T=0.01, sched=[0.5, 1e-8, 10], mcsteps=3,
         rng_seed=5, srand_seed=5),
    dict(name="boixo_sa", kind="sa", inst="boixo", sched=[1.0, 0.01, 10], mcsteps=3,
         rng_seed=123, srand_seed=2),
    dict(name="boixo16_qa", kind="qa", inst="boixo16", P=8, T=0.05, sched=[1.0, 1e-8, 12], mcsteps=2,
         rng_seed=16, srand_seed=16),
    # config 2 (examples/spinglass32.py:58-68): full 100-step anneal, three replicas
    dict(name="sg32_qa_r0", kind="qa", inst="inst_0_32x32", P=20, T=0.01, sched=[1.5, 1e-8, 100],
         mcsteps=1, rng_seed=0, srand_seed=0),
    dict(name="sg32_qa_r1", kind="qa", inst="inst_0_32x32", P=20, T=0.01, sched=[1.5, 1e-8, 100],
         mcsteps=1, rng_seed=1, srand_seed=1),
    dict(name="sg32_qa_hot", kind="qa", inst="inst_0_32x32", P=20, T=0.5, sched=[1.5, 1e-8, 20],
         mcsteps=2, rng_seed=2, srand_seed=2),
    dict(name="sg32_qa_p64", kind="qa", inst="inst_0_32x32", P=64, T=0.01, sched=[1.5, 1e-8, 10],
         mcsteps=1, rng_seed=3, srand_seed=3),
    dict(name="sg32_sa_pre", kind="sa", inst="inst_0_32x32", sched=[3.0, 0.01, 100], mcsteps=1,
         rng_seed=0, srand_seed=0),
    dict(name="sg32_sa_r4", kind="sa", inst="inst_0_32x32", sched=[3.0, 0.01, 30], mcsteps=3,
         rng_seed=4, srand_seed=4),
    # config 3 (examples/santoro80.py:23-33), shortened schedule
    dict(name="santoro_qa", kind="qa", inst="santoro_80x80", P=20, T=0.01, sched=[1.5, 1e-8, 10],
         mcsteps=1, rng_seed=80, srand_seed=80),
    dict(name="santoro_sa", kind="sa", inst="santoro_80x80", sched=[3.0, 0.01, 10], mcsteps=1,
         rng_seed=81, srand_seed=81),
    # config 4 instances through the ELL path (maxnb 5 and 8, diagonal fields, pad rows)
    dict(name="bipartite8_sa", kind="sa", inst="bipartite8", sched=[3.0, 0.01, 10], mcsteps=1,
         rng_seed=8, srand_seed=8),
    dict(name="bipartite8_qa", kind="qa", inst="bipartite8", P=10, T=0.01, sched=[8.0, 1e-8, 5],
         mcsteps=20, rng_seed=9, srand_seed=9),
    dict(name="hopfield8_sa", kind="sa", inst="hopfield8", sched=[8.0, 1e-8, 5], mcsteps=100,
         rng_seed=10, srand_seed=10),
    dict(name="hopfield8_qa", kind="qa", inst="hopfield8", P=10, T=0.01, sched=[8.0, 1e-8, 5],
         mcsteps=100, rng_seed=11, srand_seed=11),
    # OpenMP variants run with nthreads=1 (qmc.pyx:247-357, sa.pyx:193-265)
    dict(name="sg32_qapar", kind="qa_par", inst="inst_0_32x32", P=20, T=0.01, sched=[1.5, 1e-8, 20],
         mcsteps=1, rng_seed=6, srand_seed=6),
    dict(name="sg32_sapar", kind="sa_par", inst="inst_0_32x32", sched=[3.0, 0.01, 20], mcsteps=1,
         rng_seed=7, srand_seed=7),
    # multispin (sa.pyx:282-405; examples/spinglass32_multispin.py:57-66 with a short schedule)
    dict(name="sg32_multispin", kind="multispin", inst="inst_0_32x32", sched=[3.0, 0.01, 12], mcsteps=1,
         rng_seed=1234),
    dict(name="bipartite8_multispin", kind="multispin", inst="bipartite8", sched=[3.0, 0.01, 10],
         mcsteps=2, rng_seed=88),
]


def run_case(ref, case, J, nbs, n):
    """Run one case through the compiled reference; return dict of arrays."""
    rng = np.random.RandomState(case["rng_seed"])
    a, b, k = case["sched"]
    sched = np.linspace(a, b, int(k))
    kind = case["kind"]
    out = {}
    if kind == "multispin":
        # The reference's unpack loop runs one column too far (sa.pyx:402-405, bounds checks off):
        # it rewrites svec_mat[1:,0] with the bits of whatever follows svec[] in memory and
        # writes one element past the end.  Give it a 65th row to scribble on; rows 1..63 of
        # column 0 of its output are garbage and are excluded from every comparison.
        buf = np.zeros((65, n), dtype=np.float64)
        bits = buf[:64]
        bits[:] = np.array([[rng.randint(2) for _ in range(n)] for _ in range(64)], dtype=np.float64)
        out["init"] = bits.astype(np.int8)
        ref.sa.Anneal_multispin(sched, case["mcsteps"], bits, nbs, rng)
        out["out"] = bits.astype(np.int8)
        out["rng_next"] = np.array(rng.randint(1 << 30))
        return out
    sv = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
    out["init"] = sv.astype(np.int8)
    libc.srand(case["srand_seed"])
    if kind == "qa":
        confs = np.tile(sv, (case["P"], 1)).T          # F-strided view, as the examples pass it
        ref.qmc.QuantumAnneal(sched, case["mcsteps"], case["P"], case["T"], n, confs, nbs, rng)
        res = confs
    elif kind == "qa_par":
        confs = np.tile(sv, (case["P"], 1)).T.copy()
        ref.qmc.QuantumAnneal_parallel(sched, case["mcsteps"], case["P"], case["T"], n, confs, nbs, 1)
        res = confs
    elif kind == "sa":
        ref.sa.Anneal(sched, case["mcsteps"], sv, nbs, rng)
        res = sv
    elif kind == "sa_par":
        ref.sa.Anneal_parallel(sched, case["mcsteps"], sv, nbs, 1)
        res = sv
    out["libc_next"] = np.array(libc.rand())
    out["rng_next"] = np.array(rng.randint(1 << 30))
    out["out"] = np.ascontiguousarray(res).astype(np.int8)
    if res.ndim == 2:
        out["energy"] = np.array([ref.sa.ClassicalIsingEnergy(res[:, k], J) for k in range(res.shape[1])])
    else:
        out["energy"] = np.array([ref.sa.ClassicalIsingEnergy(res, J)])
    return out


def make_vectors():
    ref = O.ref()
    from piqmc_ref import qmc, sa, tools  # noqa: F401
    ref.qmc, ref.sa, ref.tools = qmc, sa, tools
    inst = instances()
    data = {}
    # ---- instances + known answers
    for name, (ijv, n) in inst.items():
        data["inst_" + name] = ijv
    data["gs_inst_0_32x32"] = ground_state("spinglass32.py", 1024)
    data["gs_santoro_80x80"] = ground_state("santoro80.py", 6400)
    # energies via the reference's convention E = g.(-J g)   (examples/spinglass32.py:86)
    for name in ("inst_0_32x32", "santoro_80x80"):
        ijv, n = inst[name]
        J = dok_from_ijv(ijv, n)
        g = data["gs_" + name].astype(np.float64)
        data["gs_energy_" + name] = np.array(np.dot(g, -J.dot(g)))
        data["gs_energy_cie_" + name] = np.array(sa.ClassicalIsingEnergy(g, J))
    np.savez_compressed(os.path.join(HERE, "instances.npz"), **data)

    vec = {}
    # ---- GenerateNeighbors outputs (tools.pyx:74-96; Py3 DOK insertion order)
    Js, nbss = {}, {}
    for name, (ijv, n) in inst.items():
        J = dok_from_ijv(ijv, n)
        nbs = np.asarray(tools.GenerateNeighbors(n, J, MAXNB[name]))
        Js[name], nbss[name] = J, nbs
        vec["nbs_" + name] = nbs
    # the 5-spin graph of testing/test_core.py:30-52
    J5 = sps.dok_matrix((5, 5), dtype=np.float64)
    for (i, j) in ((0, 1), (1, 2), (2, 3), (3, 4), (0, 3), (1, 4), (0, 4), (0, 0), (1, 1), (2, 2), (3, 3), (4, 4)):
        J5[i, j] = 1
    vec["nbs_core5"] = np.asarray(tools.GenerateNeighbors(nspins=5, J=J5, maxnb=4, savepath=None))
    # ---- Generate2DIsingInstance (tools.pyx:98-130)
    for L, seed in ((4, 123), (6, 7)):
        Jg = tools.Generate2DIsingInstance(L, np.random.RandomState(seed)).tocoo()
        vec["gen2d_%d_%d" % (L, seed)] = np.stack([Jg.row, Jg.col, Jg.data], axis=1)
    # ---- ClassicalIsingEnergy known answers (testing/test_boixo.py:60-70) + random probes
    prng = np.random.RandomState(99)
    for name in ("boixo", "bipartite8", "hopfield8", "inst_0_32x32"):
        n = inst[name][1]
        sv = (2 * prng.randint(2, size=(6, n)) - 1).astype(np.float64)
        vec["energy_probe_spins_" + name] = sv.astype(np.int8)
        vec["energy_probe_" + name] = np.array([sa.ClassicalIsingEnergy(v, Js[name]) for v in sv])
    # ---- annealing cases
    for case in CASES:
        name = case["inst"]
        res = run_case(ref, case, Js[name], nbss[name], inst[name][1])
        # count the uniforms the reference consumed, via the pinned restatement
        if case["kind"] in ("qa", "sa", "qa_par", "sa_par"):
            g = O.glibc_state(case["srand_seed"])
            rng = np.random.RandomState(case["rng_seed"])
            n = inst[name][1]
            sv = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
            a, b, k = case["sched"]
            sched = np.linspace(a, b, int(k))
            nsw = sched.size * case["mcsteps"]
            if case["kind"] == "qa":
                c = np.tile(sv, (case["P"], 1)).T.copy()
                cnt = O.qa_reference(sched, case["mcsteps"], case["P"], case["T"], n, c, nbss[name],
                                     O.make_perms(rng, n, nsw), gstate=g)
            elif case["kind"] == "qa_par":
                c = np.tile(sv, (case["P"], 1)).T.copy()
                cnt = O.qa_parallel1(sched, case["mcsteps"], case["P"], case["T"], n, c, nbss[name], gstate=g)
            elif case["kind"] == "sa":
                c = sv.copy()
                cnt = O.sa_reference(sched, case["mcsteps"], c, nbss[name], O.make_perms(rng, n, nsw), gstate=g)
            else:
                c = sv.copy()
                cnt = O.sa_parallel1(sched, case["mcsteps"], c, nbss[name], gstate=g)
            assert np.array_equal(c.astype(np.int8), res["out"]), case["name"]
            assert O.lib().oracle_glibc_rand(ctypes.byref(g)) == int(res["libc_next"]), case["name"]
            res["consumed"] = np.array(cnt, dtype=np.uint64)
        for k, v in res.items():
            vec[case["name"] + "__" + k] = v
        print(case["name"], "ok", {k: (v.shape if v.ndim else v.item()) for k, v in res.items() if k != "out"})
    vec["cases_json"] = np.array(json.dumps(CASES))
    vec["maxnb_json"] = np.array(json.dumps(MAXNB))
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **vec)


# --------------------------------------------------------------------- distributions (KS tests)
_G = {}


def _dist_init():
    ref = O.ref()
    from piqmc_ref import qmc, sa, tools
    ijv = load_ijv("inst_0_32x32")
    J = dok_from_ijv(ijv, 1024)
    _G.update(qmc=qmc, sa=sa, J=J.tocsr(), nbs=np.asarray(tools.GenerateNeighbors(1024, J, 4)))


def _energies(confs):
    """ClassicalIsingEnergy per column (no diagonal in this instance): -s.J.s"""
    J = _G["J"]
    return -np.einsum("ik,ik->k", confs, J.dot(confs))


def _dist_job(args):
    kind, r, nsteps = args
    N, P, T = 1024, 20, 0.01
    rng = np.random.RandomState(r)
    libc.srand(r)
    sv = np.array([2 * rng.randint(2) - 1 for _ in range(N)], dtype=np.float64)
    if kind == "sa":
        _G["sa"].Anneal(np.linspace(3.0, 0.01, nsteps), 1, sv, _G["nbs"], rng)
        return _energies(sv[:, None])
    confs = np.tile(sv, (P, 1)).T.copy()
    sched = np.linspace(1.5, 1e-8, nsteps)
    if kind == "qa":
        _G["qmc"].QuantumAnneal(sched, 1, P, T, N, confs, _G["nbs"], rng)
    else:
        _G["qmc"].QuantumAnneal_parallel(sched, 1, P, T, N, confs, _G["nbs"], 1)
    return _energies(confs)


def make_distributions(nrep=1024):
    """Config 2 (examples/spinglass32.py:58-68): N=1024, P=20, T=0.01, Gamma 1.5->1e-8.
    Replica r uses RandomState(r) and srand(r).  Stored: final energy of every slice."""
    out = {}
    with mp.Pool(min(8, os.cpu_count()), initializer=_dist_init) as pool:
        for kind, nsteps in (("qa_par", 100), ("qa_par", 30), ("qa", 100), ("sa", 100), ("sa", 30)):
            res = pool.map(_dist_job, [(kind, r, nsteps) for r in range(nrep)], chunksize=8)
            out["%s_%d" % (kind, nsteps)] = np.array(res)
            print(kind, nsteps, "mean residual/spin",
                  (np.array(res).mean() + 1591.9166416866) / 1024.0)
    np.savez_compressed(os.path.join(HERE, "ref_distributions.npz"), **out)


def _santoro_init():
    ref = O.ref()
    from piqmc_ref import qmc, tools
    ijv = load_ijv("santoro_80x80")
    J = dok_from_ijv(ijv, 6400)
    _G.update(qmc=qmc, J=J.tocsr(), nbs=np.asarray(tools.GenerateNeighbors(6400, J, 4)))


def _santoro_job(args):
    r, nsteps = args
    N, P, T = 6400, 20, 0.01
    rng = np.random.RandomState(r)
    libc.srand(r)
    sv = np.array([2 * rng.randint(2) - 1 for _ in range(N)], dtype=np.float64)
    confs = np.tile(sv, (P, 1)).T.copy()
    _G["qmc"].QuantumAnneal_parallel(np.linspace(1.5, 1e-8, nsteps), 1, P, T, N, confs, _G["nbs"], 1)
    return _energies(confs)


def make_santoro(nrep=256):
    """Config 3 (examples/santoro80.py:23-33): 80x80, P=20, T=0.01, Gamma 1.5->1e-8 in tau steps,
    residual energy vs tau with the reference's per-spin-reset variant (nthreads=1)."""
    out = {}
    with mp.Pool(min(8, os.cpu_count()), initializer=_santoro_init) as pool:
        for tau in (10, 30, 100):
            res = np.array(pool.map(_santoro_job, [(r, tau) for r in range(nrep)], chunksize=4))
            out["qa_par_%d" % tau] = res
            print("santoro tau", tau, "mean residual/spin", (res.mean() + 10115.3067314770) / 6400.0)
    np.savez_compressed(os.path.join(HERE, "ref_santoro.npz"), **out)


def make_santoro_long():
    """The long end of the Santoro protocol (examples/santoro80.py:290-323 sweeps tau upwards): tau = 300 and
    tau = 1000, fewer replicas (the reference takes 0.4 s and 1.3 s of one core per replica and step count)."""
    out = {}
    with mp.Pool(min(8, os.cpu_count()), initializer=_santoro_init) as pool:
        for tau, nrep in ((300, 256), (1000, 128)):
            res = np.array(pool.map(_santoro_job, [(r, tau) for r in range(nrep)], chunksize=2))
            out["qa_par_%d" % tau] = res
            print("santoro tau", tau, "mean residual/spin", (res.mean() + 10115.3067314770) / 6400.0)
    np.savez_compressed(os.path.join(HERE, "ref_santoro_long.npz"), **out)


def make_config4(nsamples=4096):
    """BASELINE configs[3]: bipartite8 (examples/bipartite8.py:20-27,60-66) and hopfield8
    (examples/hopfield8.py:22-45,98) annealed with the reference's sa.Anneal and
    sa.Anneal_multispin from random starts; stored: final state index (8 bits, bit i = spin i down)
    and final energy of every sample."""
    ref = O.ref()
    from piqmc_ref import sa, tools
    inst = instances()
    out = {}
    cfgs = {"bipartite8": (np.linspace(3.0, 0.01, 10), 1), "hopfield8": (8.0 * (1e-8 / 8.0) ** (np.arange(5) / 5.0), 100)}
    for name, (sched, mcsteps) in cfgs.items():
        ijv, n = inst[name]
        J = dok_from_ijv(ijv, n)
        nbs = np.asarray(tools.GenerateNeighbors(n, J, MAXNB[name]))
        out["sched_" + name] = sched
        out["mcsteps_" + name] = np.array(mcsteps)
        rng = np.random.RandomState(4)
        libc.srand(4)
        states, en = [], []
        for _ in range(nsamples):
            sv = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
            sa.Anneal(sched, mcsteps, sv, nbs, rng)
            states.append(int(np.sum((sv < 0) * (1 << np.arange(n)))))
            en.append(sa.ClassicalIsingEnergy(sv, J))
        out["sa_state_" + name] = np.array(states, dtype=np.int32)
        out["sa_energy_" + name] = np.array(en)
        states, en = [], []
        rng = np.random.RandomState(5)
        for _ in range(nsamples // 64):
            buf = np.zeros((65, n))
            bits = buf[:64]
            bits[:] = rng.randint(2, size=(64, n))
            sa.Anneal_multispin(sched, mcsteps, bits, nbs, rng)
            # column 0 of rows 1..63 is corrupted by the reference's unpack overrun: use row 0 only
            # for spin 0 ... instead keep all rows but mark spin 0 as unknown for rows >= 1
            for k in range(64):
                sv = 1.0 - 2.0 * bits[k]
                if k > 0:
                    continue
                states.append(int(np.sum((sv < 0) * (1 << np.arange(n)))))
                en.append(sa.ClassicalIsingEnergy(sv, J))
        out["ms_state_" + name] = np.array(states, dtype=np.int32)
        out["ms_energy_" + name] = np.array(en)
        print(name, "sa mean E", np.mean(out["sa_energy_" + name]), "multispin(row 0) mean E", np.mean(en))
    np.savez_compressed(os.path.join(HERE, "ref_config4.npz"), **out)


if __name__ == "__main__":
    if "--config4" in sys.argv:
        make_config4()
        sys.exit(0)
    if "--santoro" in sys.argv:
        make_santoro()
        sys.exit(0)
    if "--santoro-long" in sys.argv:
        make_santoro_long()
        sys.exit(0)
    make_vectors()
    if "--dist" in sys.argv:
        make_distributions()
        make_config4()
        make_santoro()
