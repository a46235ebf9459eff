"""The 8-qubit Boixo et al. problem (arXiv:1212.1739) -- Py3 restatement of the reference's
examples/boixo.py on the B200 library: SA pre-anneal + PIQMC, drop-in (bit-exact) calls."""
import numpy as np

import _instances  # noqa: F401  (sets sys.path)
import piqmc.qmc as qmc
import piqmc.sa as sa
import piqmc.tools as tools

nspins, trotterslices, annealingtemp = 8, 20, 0.01
rng = np.random.RandomState(123)
isingJ = _instances.load("boixo", nspins)
neighbors = tools.GenerateNeighbors(nspins, isingJ, 4)

spinVector = np.array([2 * rng.randint(2) - 1 for _ in range(nspins)], dtype=np.float64)
sa.Anneal(np.linspace(1.0, 0.01, 10), 1, spinVector, neighbors, rng)
print("SA energy:", sa.ClassicalIsingEnergy(spinVector, isingJ), tools.spins2bits(spinVector))

configurations = np.tile(spinVector, (trotterslices, 1)).T
qmc.QuantumAnneal(np.linspace(0.5, 1e-8, 10), 1, trotterslices, annealingtemp, nspins, configurations,
                  neighbors, rng)
energies = [sa.ClassicalIsingEnergy(configurations[:, k], isingJ) for k in range(trotterslices)]
print("PIQMC slice energies:", energies)
