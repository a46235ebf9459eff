"""Reference examples/spinglass32_dense.py restated: the dense drop-in calls sa.Anneal_dense and
qmc.QuantumAnneal_dense (bit-exact replays of piqmc/sa.pyx:126-187 and piqmc/qmc.pyx:141-242) on a
64-spin corner of the 32x32 instance (a dense replay is one GPU thread walking an N x N matrix per
attempt: it is the parity path, not the fast one)."""
import numpy as np

import _instances
import piqmc.qmc as qmc
import piqmc.sa as sa

n, P, T = 64, 20, 0.01
rng = np.random.RandomState(0)
isingJ = _instances.load("inst_0_32x32", 1024).tocsr()[:n, :n].todok()
J = np.asarray(isingJ.todense())

spinVector = np.array([2 * rng.randint(2) - 1 for _ in range(n)], dtype=np.float64)
print("Initial state energy:", sa.ClassicalIsingEnergy(spinVector, isingJ))
sa.Anneal_dense(np.linspace(3.0, 0.01, 100), 1, spinVector, J, rng)
print("Final SA energy:     ", sa.ClassicalIsingEnergy(spinVector, isingJ))

configurations = np.tile(spinVector, (P, 1)).T
qmc.QuantumAnneal_dense(np.linspace(1.5, 1e-8, 100), 1, P, T, n, configurations, J, rng)
print("Final PIQMC energies:", sorted(sa.ClassicalIsingEnergy(configurations[:, k], isingJ) for k in range(P))[:3])
