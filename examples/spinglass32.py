"""32x32 random Ising spin glass (reference examples/spinglass32.py), 1000 replicas at once on
the production path: SA pre-anneal (64 replicas per word) handing its replicas to PIQMC (P=20
slices) on the device -- the np.tile(spinVector, (P, 1)).T of the reference without a host trip."""
import numpy as np

import _instances
import piqmc.qmc as qmc
import piqmc.sa as sa
import piqmc.tools as tools

nspins, P, T, R = 1024, 20, 0.01, 1000
isingJ = _instances.load("inst_0_32x32", nspins)
gs, gs_energy = _instances.ground_state("inst_0_32x32")
neighbors = tools.GenerateNeighbors(nspins, isingJ, 4)

pre = sa.AnnealReplicas(np.linspace(3.0, 0.01, 100), 1, None, neighbors, seed=1, nreplicas=R, download=False)
print("SA  residual energy per spin: %.4f" % ((pre["energies"].mean() - gs_energy) / nspins))

out = qmc.QuantumAnnealReplicas(np.linspace(1.5, 1e-8, 100), 1, P, T, nspins, "resident", neighbors,
                                seed=2, order="natural", nreplicas=R)
best = out["energies"].min(axis=1)
print("QA  residual energy per spin: %.4f (slice mean), %.4f (best slice)"
      % ((out["energies"].mean() - gs_energy) / nspins, (best.mean() - gs_energy) / nspins))

# The reference's default function (qmc.QuantumAnneal, what examples/spinglass32.py of the reference calls)
# carries its energy difference over a whole slice sweep (qmc.pyx:134-135); semantics="reference" runs exactly
# that rule, a fresh permutation per sweep, from the same kind of random start -- markedly different statistics.
ref = qmc.QuantumAnnealReplicas(np.linspace(1.5, 1e-8, 100), 1, P, T, nspins, None, neighbors, seed=3,
                                order="permutation", semantics="reference", nreplicas=R)
print("QA, as-shipped rule: residual energy per spin %.4f (slice mean)" % ((ref["energies"].mean() - gs_energy) / nspins))
