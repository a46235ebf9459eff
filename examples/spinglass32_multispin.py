"""Reference examples/spinglass32_multispin.py restated: 64 simultaneous anneals of the 32x32
instance, multispin-coded.  sa.Anneal_multispin is the bit-exact replay (bits in, bits out, replica
k in bit 63-k); sa.AnnealReplicas is the production path (any number of replicas, 64 per word)."""
import numpy as np

import _instances
import piqmc.sa as sa
import piqmc.tools as tools

nspins = 1024
rng = np.random.RandomState(1234)
isingJ = _instances.load("inst_0_32x32", nspins)
gs, gs_energy = _instances.ground_state("inst_0_32x32")
neighbors = tools.GenerateNeighbors(nspins, isingJ, 4)
sched = np.linspace(3.0, 0.01, 200)

bits = np.array([[rng.randint(2) for _ in range(nspins)] for _ in range(64)], dtype=np.float64)
sa.Anneal_multispin(sched, 1, bits, neighbors, rng)
en = [sa.ClassicalIsingEnergy(tools.bits2spins(b), isingJ) for b in bits]
print("Anneal_multispin (replay):  residual/spin %.4f" % ((np.mean(en) - gs_energy) / nspins))

out = sa.AnnealReplicas(sched, 1, None, neighbors, seed=1234, nreplicas=4096)
print("AnnealReplicas (4096 reps): residual/spin %.4f" % ((out["energies"].mean() - gs_energy) / nspins))
