"""CPU tests of the multi-GPU host logic with gloo, world_size 2 (and an uneven world of 3):
replica sharding, the final gather, and -- through the CPU statement of the colour semantics --
that annealing shards [replica0, replica0+count) separately gives exactly the unsharded result
(the Philox key carries the global replica id)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from piqmc.shard import gather_rows, shard_replicas


def test_shard_arithmetic():
    for R in (1, 7, 64, 4096, 1000):
        for W in (1, 2, 3, 4, 8):
            spans = [shard_replicas(R, W, r) for r in range(W)]
            assert spans[0][0] == 0
            assert sum(c for _, c in spans) == R
            for (a, ca), (b, _) in zip(spans, spans[1:]):
                assert b == a + ca
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert shard_replicas(4096, 8, 3) == (1536, 512)
    with pytest.raises(ValueError):
        shard_replicas(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, R, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "pathintegral-qmc_b200"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    import piqmc.tools as T
    nbs, color = T.GaussianTorusNeighbors(8, 5)
    idx, J32 = O.nbs_to_ell(nbs)
    n, P = 64, 8
    r0, cnt = shard_replicas(R, world, rank)
    spins = np.repeat(O.colour_init_spins(77, r0, cnt, n)[:, :, None], P, axis=2).copy()
    O.qa_colour(np.linspace(1.5, 1e-8, 5), 1, P, 0.05, idx, J32, color, spins, 77, replica0=r0)
    en = np.array([[O.energy_ell(idx, J32, np.ascontiguousarray(spins[r, :, k])) for k in range(P)]
                   for r in range(cnt)]).reshape(cnt, P)
    words = T.PackWords(np.transpose(spins, (0, 2, 1))).astype(np.int64)
    g_en = gather_rows(torch.from_numpy(en), R)
    g_w = gather_rows(torch.from_numpy(words), R)
    if rank == 0:
        np.save(os.path.join(out_dir, "en.npy"), g_en.numpy())
        np.save(os.path.join(out_dir, "w.npy"), g_w.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,R", [(2, 6), (3, 7)])
def test_gloo_sharded_anneal_equals_unsharded(tmp_path, world, R):
    from oracle import oracle as O
    import piqmc.tools as T
    port = _free_port()
    mp.spawn(_worker, args=(world, port, R, str(tmp_path)), nprocs=world, join=True)
    nbs, color = T.GaussianTorusNeighbors(8, 5)
    idx, J32 = O.nbs_to_ell(nbs)
    n, P = 64, 8
    spins = np.repeat(O.colour_init_spins(77, 0, R, n)[:, :, None], P, axis=2).copy()
    O.qa_colour(np.linspace(1.5, 1e-8, 5), 1, P, 0.05, idx, J32, color, spins, 77)
    en = np.array([[O.energy_ell(idx, J32, np.ascontiguousarray(spins[r, :, k])) for k in range(P)]
                   for r in range(R)])
    words = T.PackWords(np.transpose(spins, (0, 2, 1))).astype(np.int64)
    assert np.array_equal(np.load(tmp_path / "w.npy"), words)
    np.testing.assert_allclose(np.load(tmp_path / "en.npy"), en, rtol=0, atol=0)


def _gpu_worker(rank, world, port, R, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "pathintegral-qmc_b200"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", 0))
    import piqmc.qmc as qmc
    import piqmc.tools as T
    from piqmc import device
    from piqmc.shard import gather_energies, gather_words
    nbs, color = T.GaussianTorusNeighbors(8, 5)
    dev = device.Device(0)
    sched = np.linspace(1.5, 1e-8, 5)
    out = {}
    for S in (1, 3):                                 # one replica per word, then three (P = 20)
        qmc.QuantumAnnealReplicas(sched, 1, 20, 0.05, 64, None, nbs, 77, color=color, nreplicas=R, device=dev,
                                  per_word=S, energies=False, download=False)
        out[S] = (gather_energies(dev, R).cpu().numpy(), gather_words(dev, R).cpu().numpy())
    np.save(os.path.join(out_dir, "en1.npy"), out[1][0])
    np.save(os.path.join(out_dir, "en3.npy"), out[3][0])
    np.save(os.path.join(out_dir, "w1.npy"), out[1][1])
    np.save(os.path.join(out_dir, "w3.npy"), out[3][1])
    dist.destroy_process_group()


@pytest.mark.gpu
def test_gather_of_states_with_several_replicas_per_word(tmp_path):
    """gather_energies / gather_words on a state with three replicas per word (P = 20; 100 replicas do not fill
    the last word) give, replica by replica, what the one-replica-per-word state gives."""
    R = 100
    mp.spawn(_gpu_worker, args=(1, _free_port(), R, str(tmp_path)), nprocs=1, join=True)
    en1, en3 = np.load(tmp_path / "en1.npy"), np.load(tmp_path / "en3.npy")
    w1, w3 = np.load(tmp_path / "w1.npy"), np.load(tmp_path / "w3.npy")
    assert en1.shape == (R, 20) and w1.shape == (R, 64)
    assert np.array_equal(en1, en3) and np.array_equal(w1, w3)
