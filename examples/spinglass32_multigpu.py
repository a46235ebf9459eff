"""Reference examples/spinglass32_mpi.py restated for one box of GPUs: independent replicas per
rank, no communication during the anneal, one gather of the energies at the end.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        examples/spinglass32_multigpu.py
"""
import os

import numpy as np
import torch
import torch.distributed as dist

import _instances
import piqmc.qmc as qmc
import piqmc.tools as tools
from piqmc import device
from piqmc.shard import gather_energies, shard_replicas

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

nspins, P, T, R = 1024, 20, 0.01, 4096
isingJ = _instances.load("inst_0_32x32", nspins)
_, gs_energy = _instances.ground_state("inst_0_32x32")
neighbors = tools.GenerateNeighbors(nspins, isingJ, 4)
dev = device.Device(local)
replica0, count = shard_replicas(R, world, rank)
qmc.QuantumAnnealReplicas(np.linspace(1.5, 1e-8, 100), 1, P, T, nspins, None, neighbors, seed=7, order="natural",
                          nreplicas=count, replica0=replica0, device=dev, download=False)
en = gather_energies(dev, R).cpu().numpy() if world > 1 else dev.energy()
if rank == 0:
    print("%d replicas on %d GPU(s): residual/spin %.4f (best slice %.4f)"
          % (R, world, (en.mean() - gs_energy) / nspins, (en.min(axis=1).mean() - gs_energy) / nspins))
if world > 1:
    dist.destroy_process_group()
