"""Residual energy vs annealing time on the 80x80 Martonak-Santoro-Tosatti instance
(reference examples/santoro80.py; PRB 66, 094203 protocol), 256 replicas per point."""
import numpy as np

import _instances
import piqmc.qmc as qmc
import piqmc.sa as sa
import piqmc.tools as tools

nspins, P, T, R = 6400, 20, 0.01, 256
isingJ = _instances.load("santoro_80x80", nspins)
gs, gs_energy = _instances.ground_state("santoro_80x80")
neighbors = tools.GenerateNeighbors(nspins, isingJ, 4)
print("ground state energy %.6f (%.8f per spin)" % (gs_energy, gs_energy / nspins))
print("%8s %14s %14s" % ("tau", "SA residual", "PIQMC residual"))
for tau in (10, 30, 100, 300, 1000):
    s = sa.AnnealReplicas(np.linspace(3.0, 0.01, tau), 1, None, neighbors, seed=tau, nreplicas=R)
    q = qmc.QuantumAnnealReplicas(np.linspace(1.5, 1e-8, tau), 1, P, T, nspins, None, neighbors, seed=tau,
                                  nreplicas=R)
    print("%8d %14.5f %14.5f" % (tau, (s["energies"].mean() - gs_energy) / nspins,
                                 (q["energies"].min(axis=1).mean() - gs_energy) / nspins))
