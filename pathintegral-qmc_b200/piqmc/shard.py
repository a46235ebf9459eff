"""Replica sharding across the GPUs of one box (one process per GPU, torch.distributed).

Replicas never interact -- in the reference they are separate calls or separate MPI ranks that
never communicate (examples/spinglass32_mpi.py:20-26,74-80) -- so the sweep path needs no
collective: rank g anneals the contiguous block of replicas [replica0, replica0 + count) and
the Philox key carries the GLOBAL replica id, which makes the result independent of the number of
ranks.  The only communication is one gather of the final energies (and, optionally, of the
bit-packed configurations) at the end.
"""
import numpy as np


def shard_replicas(nreplicas, world_size, rank):
    """Contiguous shard of `nreplicas` for `rank`: (replica0, count).  The first
    nreplicas % world_size ranks take one extra replica."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(int(nreplicas), int(world_size))
    count = base + (1 if rank < extra else 0)
    replica0 = rank * base + min(rank, extra)
    return replica0, count


def bind_to_gpu_numa(gpu_index):
    """Pin this process to the CPU cores next to GPU `gpu_index` (NVML's ideal affinity) before it allocates
    its pinned host buffers: with one process per GPU the eight result downloads of a box then land in the
    memory of the socket their GPU hangs on instead of all crossing to one node.  Returns the number of cores
    bound to, or 0 when NVML or the affinity call is not available (nothing changes then)."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w in range(words) for b in range(64) if (int(mask[w]) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def gather_rows(local, nreplicas, group=None):
    """All-gather a per-replica tensor (first dimension = this rank's replicas, sharded with
    shard_replicas) into the full [nreplicas, ...] tensor, in global replica order, on every
    rank.  Works on CUDA tensors (NCCL) and CPU tensors (gloo); uneven shards are padded."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = [shard_replicas(nreplicas, world, r)[1] for r in range(world)]
    cmax = max(counts)
    if local.shape[0] != counts[dist.get_rank(group)]:
        raise ValueError("local tensor does not match this rank's shard")
    pad = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world, cmax) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view((world * cmax,) + tuple(local.shape[1:])), pad, group=group)
    return torch.cat([out[r, :counts[r]] for r in range(world)], dim=0)


def _my_count(nreplicas, group):
    import torch.distributed as dist
    return shard_replicas(nreplicas, dist.get_world_size(group), dist.get_rank(group))[1]


def gather_energies(dev, nreplicas, group=None):
    """Final-energy gather for a Device whose resident rows are this rank's replica shard:
    runs the device energy reduction and all-gathers float64[nreplicas, slices] over NCCL without
    staging through the host.  States with several replicas per word (per_word > 1) carry padding
    replicas in their last word: only this rank's replicas are sent."""
    import torch
    dev.energy(download=False)
    dev.synchronize()
    local = torch.as_tensor(dev.energy_device_array(), device="cuda:%d" % dev.index)
    return gather_rows(local[:_my_count(nreplicas, group)], nreplicas, group)


def gather_words(dev, nreplicas, group=None):
    """Gather of the bit-packed final configurations: uint64[nreplicas, nspins] on every rank (bit k of a
    word = slice k of that replica, whatever the number of replicas per word on the device)."""
    import torch
    dev.synchronize()
    # device layout is [nspins, nrows]; gather along replicas needs [nrows, nspins].  NCCL has no
    # uint64: reinterpret as int64.
    local = torch.as_tensor(dev.state_device_array(), device="cuda:%d" % dev.index).view(torch.int64)
    local = local.t().contiguous()
    S = dev.per_word
    if S > 1:                                      # segments of `slices` lanes -> one word per replica
        P = dev.lanes // S
        mask = (1 << P) - 1
        local = torch.stack([(local >> (g * P)) & mask for g in range(S)], dim=1).reshape(-1, local.shape[1])
    return gather_rows(local[:_my_count(nreplicas, group)].contiguous(), nreplicas, group)
