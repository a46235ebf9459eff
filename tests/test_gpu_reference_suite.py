"""The reference's own test-suite (testing/test_boixo.py, the energy and annealing parts of
testing/test_core.py), restated in Python 3 and run unchanged in substance against the drop-in
package on the GPU: same calls, same parameters, same assertions.  (The reference seeds nothing but
`rng`; libc rand() is seeded here so that the run is reproducible.)"""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sps

pytestmark = pytest.mark.gpu


class TestBoixo:
    """testing/test_boixo.py:20-114 -- the 8-qubit diamond graph of Boixo et al. (arXiv:1212.1739)."""

    def setup_method(self, method):
        import piqmc.tools as tools
        self.nspins = 8
        self.preannealingtemp = 1.0
        self.annealingtemp = 0.01
        self.annealingsteps = 10
        self.annealingmcsteps = 3
        self.trotterslices = 5
        self.fieldstart = 0.5
        self.fieldend = 1e-8
        self.rng = np.random.RandomState(123)
        ctypes.CDLL(None).srand(123)
        kvp = [[1, 2, 1.0], [1, 4, 1.0], [1, 5, 1.0], [2, 3, 1.0], [2, 6, 1.0], [3, 4, 1.0], [3, 7, 1.0],
               [4, 8, 1.0], [1, 1, 1.0], [2, 2, 1.0], [3, 3, 1.0], [4, 4, 1.0], [5, 5, -1.0], [6, 6, -1.0],
               [7, 7, -1.0], [8, 8, -1.0]]
        self.isingJ = sps.dok_matrix((self.nspins, self.nspins))
        for i, j, val in kvp:
            self.isingJ[i - 1, j - 1] = val
        self.nbs = tools.GenerateNeighbors(self.nspins, self.isingJ, 4, None)

    def test_classicalisingenergy(self, dev):
        import piqmc.sa as sa
        svecs = np.array([[-1, 1, -1, 1, 1, -1, -1, 1],      # E = 4.0
                          [1, 1, 1, 1, 1, -1, 1, -1],        # E = -8.0
                          [1, -1, 1, 1, -1, -1, -1, -1],     # E = -4.0
                          [1, -1, -1, 1, 1, -1, 1, 1]])      # E = 0.0
        for vec, en in zip(svecs, [4.0, -8.0, -4.0, 0.0]):
            assert sa.ClassicalIsingEnergy(vec, self.isingJ) == en

    def test_sa(self, dev):
        import piqmc.sa as sa
        for i in range(self.trotterslices):
            spinVector = np.array([2 * self.rng.randint(2) - 1 for k in range(self.nspins)], dtype=np.float64)
            tannealingsched = np.linspace(self.preannealingtemp, self.annealingtemp, self.annealingsteps)
            sa.Anneal(tannealingsched, self.annealingmcsteps, spinVector, self.nbs, self.rng)
            # with the given parameters, they have surely found a ground state
            assert np.sum(spinVector[:4]) == 4 or np.sum(spinVector) == -8
            assert sa.ClassicalIsingEnergy(spinVector, self.isingJ) == -8.0

    def test_qmc(self, dev):
        import piqmc.qmc as qmc
        import piqmc.sa as sa
        configurations = np.tile(np.array([2 * self.rng.randint(2) - 1 for k in range(self.nspins)],
                                          dtype=np.float64), (self.trotterslices, 1)).T
        annealingsched = np.linspace(self.fieldstart, self.fieldend, self.annealingsteps)
        qmc.QuantumAnneal(annealingsched, self.annealingmcsteps, self.trotterslices, self.annealingtemp,
                          self.nspins, configurations, self.nbs, self.rng)
        energies = np.array([sa.ClassicalIsingEnergy(configurations[:, k], self.isingJ)
                             for k in range(self.trotterslices)])
        assert np.sum(energies) == -8.0 * self.trotterslices
        for k in range(self.trotterslices):
            assert np.sum(configurations[:4, k]) == 4.0 or np.sum(configurations[:, k]) == -8.0

    def test_dense_variants_reach_the_ground_state(self, dev):
        """The same protocol through sa.Anneal_dense / qmc.QuantumAnneal_dense (the reference ships
        them without tests; examples/spinglass32_dense.py is their only caller)."""
        import piqmc.qmc as qmc
        import piqmc.sa as sa
        J = np.asarray(self.isingJ.todense())
        spinVector = np.array([2 * self.rng.randint(2) - 1 for k in range(self.nspins)], dtype=np.float64)
        sa.Anneal_dense(np.linspace(self.preannealingtemp, self.annealingtemp, self.annealingsteps),
                        self.annealingmcsteps, spinVector, J, self.rng)
        assert sa.ClassicalIsingEnergy(spinVector, self.isingJ) == -8.0
        configurations = np.tile(spinVector, (self.trotterslices, 1)).T
        qmc.QuantumAnneal_dense(np.linspace(self.fieldstart, self.fieldend, self.annealingsteps),
                                self.annealingmcsteps, self.trotterslices, self.annealingtemp, self.nspins,
                                configurations, J, self.rng)
        for k in range(self.trotterslices):
            assert sa.ClassicalIsingEnergy(configurations[:, k], self.isingJ) == -8.0

    def test_production_replicas_reach_the_ground_state(self, dev):
        """... and through the production path: 256 replicas at once, every slice of every replica."""
        import piqmc.qmc as qmc
        import piqmc.sa as sa
        pre = sa.AnnealReplicas(np.linspace(self.preannealingtemp, self.annealingtemp, self.annealingsteps),
                                self.annealingmcsteps, None, self.nbs, seed=123, nreplicas=256, device=dev)
        assert np.all(pre["energies"] == -8.0)
        out = qmc.QuantumAnnealReplicas(np.linspace(self.fieldstart, self.fieldend, self.annealingsteps),
                                        self.annealingmcsteps, self.trotterslices, self.annealingtemp, self.nspins,
                                        pre["spins"], self.nbs, seed=124, device=dev)
        assert np.all(out["energies"] == -8.0)
