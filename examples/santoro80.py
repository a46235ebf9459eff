"""Residual energy vs annealing time on the 80x80 Martonak-Santoro-Tosatti instance
(reference examples/santoro80.py; PRB 66, 094203 protocol), 256 replicas per point."""
import numpy as np

import _instances
import piqmc.qmc as qmc
import piqmc.sa as sa
import piqmc.tools as tools
from piqmc import device

nspins, P, T, R = 6400, 20, 0.01, 256
isingJ = _instances.load("santoro_80x80", nspins)
gs, gs_energy = _instances.ground_state("santoro_80x80")
neighbors = tools.GenerateNeighbors(nspins, isingJ, 4)
print("ground state energy %.6f (%.8f per spin)" % (gs_energy, gs_energy / nspins))
print("%8s %14s %14s   %s" % ("tau", "SA residual", "PIQMC residual", "PIQMC best-slice residual per spin: histogram over the replicas"))
dev = device.default_device()
for tau in (10, 30, 100, 300, 1000):
    s = sa.AnnealReplicas(np.linspace(3.0, 0.01, tau), 1, None, neighbors, seed=tau, nreplicas=R)
    # the anneal leaves state and energies on the device; the statistics are reduced there too
    qmc.QuantumAnnealReplicas(np.linspace(1.5, 1e-8, tau), 1, P, T, nspins, None, neighbors, seed=tau, nreplicas=R,
                              download=False)
    h = dev.energy_histogram(e0=gs_energy, scale=1.0 / nspins, lo=0.20, hi=0.30, nbins=10, reduce="min")
    print("%8d %14.5f %14.5f   %s" % (tau, (s["energies"].mean() - gs_energy) / nspins, h["mean"],
                                      " ".join("%d" % c for c in h["counts"])))
