"""CPU tests of the host-side mirror piqmc.tools against the reference's goldens, plus the
reference's own testing/test_core.py restated."""
import numpy as np
import pytest
import scipy.sparse as sps

import piqmc.tools as tools
from helpers import NSPINS


def test_spinbitconversion():
    # testing/test_core.py:16-22
    a = np.array([0., 1., 0., 1., 0., 1., 0., 1., 0., 1.])
    b = tools.bits2spins(a)
    c = tools.spins2bits(b)
    assert np.all(a == [0, 1, 0, 1, 0, 1, 0, 1, 0, 1])
    assert np.all(b == [1, -1, 1, -1, 1, -1, 1, -1, 1, -1])
    assert np.all(c == a)


def test_generateising():
    # testing/test_core.py:24-28
    J = tools.Generate2DIsingInstance(4, np.random.RandomState(123))
    assert (J - sps.triu(J)).nnz == 0


@pytest.mark.parametrize("L,seed", [(4, 123), (6, 7)])
def test_generateising_matches_reference_draws(golden, L, seed):
    rng = np.random.RandomState(seed)
    J = tools.Generate2DIsingInstance(L, rng).tocoo()
    want = golden["vec"]["gen2d_%d_%d" % (L, seed)]
    got = np.stack([J.row, J.col, J.data], axis=1)
    assert np.array_equal(got, want)


def _core5():
    J = sps.dok_matrix((5, 5), dtype=np.float64)
    for (i, j) in ((0, 1), (1, 2), (2, 3), (3, 4), (0, 3), (1, 4), (0, 4), (0, 0), (1, 1), (2, 2), (3, 3), (4, 4)):
        J[i, j] = 1
    return J


def test_generateneighbors_core5(golden):
    """testing/test_core.py:30-81.  The literal in that test encodes Python-2 dict order; on
    Python 3 the reference itself produces insertion order (stored in the golden).  Check both:
    identical to the reference's Py3 output, and equal to the Py2 literal row by row as sets."""
    nb = tools.GenerateNeighbors(nspins=5, J=_core5(), maxnb=4, savepath=None)
    assert np.array_equal(nb, golden["vec"]["nbs_core5"])
    true_nb = np.array(
        [[[1., 1.], [0., 1.], [4., 1.], [3., 1.]],
         [[0., 1.], [2., 1.], [4., 1.], [1., 1.]],
         [[1., 1.], [2., 1.], [3., 1.], [0., 0.]],
         [[3., 1.], [2., 1.], [0., 1.], [4., 1.]],
         [[4., 1.], [1., 1.], [0., 1.], [3., 1.]]])
    for i in range(5):
        assert sorted(map(tuple, nb[i])) == sorted(map(tuple, true_nb[i]))
    assert np.array_equal(nb[2, 3], [0., 0.])           # pad row stays [0, 0] and comes last


@pytest.mark.parametrize("inst", sorted(NSPINS))
def test_generateneighbors_instances(golden, inst):
    J = tools.IsingFromTriples(golden["inst"]["inst_" + inst], NSPINS[inst])
    want = golden["vec"]["nbs_" + inst]
    nb = tools.GenerateNeighbors(NSPINS[inst], J, want.shape[1])
    assert nb.dtype == np.float64 and np.array_equal(nb, want)


def test_generateneighbors_overflow_and_save(tmp_path, golden):
    J = tools.IsingFromTriples(golden["inst"]["inst_boixo16"], 16)
    with pytest.raises(IndexError):
        tools.GenerateNeighbors(16, J, 4)               # degree 6 > maxnb
    p = str(tmp_path / "nb.npy")
    nb = tools.GenerateNeighbors(16, J, 6, p)
    assert np.array_equal(np.load(p), nb)


def test_generateneighbors_other_formats_and_lower_triangle():
    # santoro_80x80.txt stores some bonds with i > j; CSR input goes through todok()
    J = sps.dok_matrix((4, 4))
    J[2, 0] = 0.5
    J[0, 1] = -1.0
    J[3, 3] = 0.25
    nb = tools.GenerateNeighbors(4, J, 2)
    assert np.array_equal(nb[0], [[2, 0.5], [1, -1.0]])
    assert np.array_equal(nb[2], [[0, 0.5], [0, 0]])
    assert np.array_equal(nb[3], [[3, 0.25], [0, 0]])
    nb2 = tools.GenerateNeighbors(4, J.tocsr(), 2)
    assert sorted(map(tuple, nb2[0])) == sorted(map(tuple, nb[0]))


@pytest.mark.parametrize("inst", sorted(NSPINS))
def test_colouring_is_proper(golden, inst):
    nbs = golden["vec"]["nbs_" + inst]
    color = tools.ColourGraph(nbs)
    n = nbs.shape[0]
    for i in range(n):
        for j, v in nbs[i]:
            if int(j) != i and v != 0.0:
                assert color[int(j)] != color[i]
    ncol = color.max() + 1
    assert ncol == (8 if inst == "hopfield8" else 2)


def test_gaussian_torus_matches_generateneighbors():
    L = 6
    nbs, color = tools.GaussianTorusNeighbors(L, 3)
    rng = np.random.RandomState(3)
    Jb = rng.standard_normal(size=(L * L, 2)).astype(np.float32).astype(np.float64)
    J = sps.dok_matrix((L * L, L * L))
    for i in range(L * L):
        y, x = divmod(i, L)
        J[i, y * L + (x + 1) % L] = Jb[i, 0]
        J[i, ((y + 1) % L) * L + x] = Jb[i, 1]
    assert np.array_equal(nbs, tools.GenerateNeighbors(L * L, J, 4))
    assert set(color) == {0, 1}
    for i in range(L * L):
        assert all(color[int(j)] != color[i] for j in nbs[i, :, 0])


def test_pack_unpack_words():
    rng = np.random.RandomState(0)
    s = (2 * rng.randint(2, size=(3, 20, 17)) - 1).astype(np.int8)
    w = tools.PackWords(s)
    assert w.shape == (3, 17) and w.dtype == np.uint64
    assert np.array_equal(tools.UnpackWords(w, 20), s)
    assert (int(w[1, 5]) >> 7) & 1 == (1 if s[1, 7, 5] < 0 else 0)


def test_lattice_generators_match_reference():
    """tools.Generate2DLattice / GenerateKblockLattice (tools.pyx:132-272): same matrix, same key
    order and same number of draws as the reference's own functions (tests/golden/ref_lattices.npz)."""
    import json
    import os
    vec = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_lattices.npz"))
    for k, c in enumerate(json.loads(str(vec["cases_json"]))):
        rng = np.random.RandomState(c["seed"])
        if c["kind"] == "lattice":
            J = tools.Generate2DLattice(c["nrows"], c["ncols"], rng, c["arg"])
        else:
            J = tools.GenerateKblockLattice(c["nrows"], c["ncols"], rng, c["arg"])
        assert np.array_equal(J.toarray(), vec["J%d" % k]), c
        assert np.array_equal(np.array(list(J.keys()), dtype=np.int64).reshape(-1, 2), vec["keys%d" % k]), c
        assert rng.randint(1 << 30) == int(vec["next%d" % k][0]), c
        assert (J - sps.triu(J)).nnz == 0


def test_generateneighbors_with_colouring(golden):
    """GenerateNeighbors(..., colouring=...) hands back the table the reference returns plus the
    colour classes the production kernels sweep by (a proper colouring either way)."""
    J = tools.IsingFromTriples(golden["inst"]["inst_inst_0_32x32"], 1024)
    nbs = tools.GenerateNeighbors(1024, J, 4)
    for mode, nclasses in (("checkerboard", 2), ("natural", 63)):
        nbs2, color = tools.GenerateNeighbors(1024, J, 4, colouring=mode)
        assert np.array_equal(nbs, nbs2) and color.shape == (1024,) and color.max() + 1 == nclasses
        for i in range(1024):
            for j, v in nbs[i]:
                if v != 0.0 and int(j) != i:
                    assert color[int(j)] != color[i]


@pytest.mark.parametrize("seed", range(6))
def test_level_colouring_equals_sequential_sweep_on_random_graphs(seed):
    """The claim the production path rests on: sweeping the dependency levels
    of a visiting order, a whole level at once, IS the sequential sweep in that order.  Checked on
    the CPU statement of the colour semantics, on random sparse graphs with fields, random orders,
    QA (both Trotter modes) and SA, hot enough that most attempts draw a uniform."""
    from oracle import oracle as O
    rng = np.random.RandomState(1000 + seed)
    n, maxnb = 24 + 3 * seed, 5
    J = sps.dok_matrix((n, n))
    deg = np.zeros(n, dtype=int)
    for _ in range(3 * n):
        a, b = rng.randint(n, size=2)
        if a != b and (min(a, b), max(a, b)) not in J and deg[a] < maxnb - 1 and deg[b] < maxnb - 1:
            J[min(a, b), max(a, b)] = rng.uniform(-2, 2)
            deg[a] += 1
            deg[b] += 1
    for i in rng.choice(n, n // 3, replace=False):
        J[i, i] = rng.uniform(-1, 1)
    nbs = tools.GenerateNeighbors(n, J, maxnb)
    idx, J32 = O.nbs_to_ell(nbs)
    order = rng.permutation(n).astype(np.int32)
    levels = tools.OrderLevels(nbs, order)
    assert np.array_equal(levels, tools.ColourGraph(nbs, order))
    nsweeps, P, R = 4, 6, 3
    sched = np.linspace(1.5, 0.2, nsweeps)
    init = (2 * rng.randint(2, size=(R, n)) - 1).astype(np.int8)
    for trotter in (0, 1):
        a = np.repeat(init[:, :, None], P, axis=2).copy()
        b = a.copy()
        O.qa_colour(sched, 1, P, 0.7, idx, J32, levels, a, 5, trotter=trotter)
        O.qa_colour(sched, 1, P, 0.7, idx, J32, levels, b, 5, trotter=trotter, orders=np.tile(order, (nsweeps, 1)))
        assert np.array_equal(a, b)
    a, b = init.copy(), init.copy()
    O.sa_colour(np.linspace(2.0, 0.3, nsweeps), 1, idx, J32, levels, a, 9)
    O.sa_colour(np.linspace(2.0, 0.3, nsweeps), 1, idx, J32, levels, b, 9, orders=np.tile(order, (nsweeps, 1)))
    assert np.array_equal(a, b)


def test_csr_form_of_the_neighbour_table(golden):
    """GenerateNeighbors(format="csr") / NeighborsToCSR / CSRToNeighbors: rows in table order, local fields as
    self entries, every bond in both rows; the round trip gives the table back (pads are zeros)."""
    for inst in ("boixo", "bipartite8", "hopfield8", "inst_0_32x32"):
        nbs = golden["vec"]["nbs_" + inst]
        indptr, indices, data = tools.NeighborsToCSR(nbs)
        n = nbs.shape[0]
        assert indptr.shape == (n + 1,) and indptr[-1] == indices.size == data.size == np.count_nonzero(nbs[:, :, 1])
        for i in (0, n // 2, n - 1):
            row = nbs[i][nbs[i, :, 1] != 0.0]
            assert np.array_equal(indices[indptr[i]:indptr[i + 1]], row[:, 0].astype(np.int32))
            assert np.array_equal(data[indptr[i]:indptr[i + 1]], row[:, 1])
        back = tools.CSRToNeighbors(indptr, indices, data, maxnb=nbs.shape[1])
        live = nbs[:, :, 1] != 0.0
        # the table may keep zero-valued entries between live ones; compare the live entries row by row
        for i in range(n):
            assert np.array_equal(back[i][back[i, :, 1] != 0.0], nbs[i][live[i]])
    with pytest.raises(IndexError):
        tools.CSRToNeighbors(indptr, indices, data, maxnb=1)
