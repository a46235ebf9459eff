"""piqmc.qmc -- path-integral quantum annealing (Martonak-Santoro-Tosatti, PRB 66, 094203) on B200.

Mirror of the reference module piqmc/qmc.pyx (same names, argument order, in-place behaviour);
the sweeps run in libpiqmc_b200.so on the GPU.  There is no CPU fallback.

  QuantumAnneal           drop-in, bit-exact replay of qmc.QuantumAnneal (qmc.pyx:30-136)
  QuantumAnneal_parallel  same signature as the OpenMP variant (qmc.pyx:247-357); runs the
                          colour-parallel kernel (per-spin ediff, like that variant)
  QuantumAnnealBatch      R replicas of QuantumAnneal in one launch (deterministic)
  QuantumAnnealReplicas   production path: R replicas x P slices, bit-packed, colour-parallel,
                          Philox; the path bench.py measures
"""
import ctypes

import numpy as np

from . import device as _dev
from . import tools as _tools
from ._lib import RandState, lib
from .sa import _draw_perms, _f64, _resolve_order, _spins_i8

__all__ = ["QuantumAnneal", "QuantumAnneal_parallel", "QuantumAnnealBatch", "QuantumAnnealReplicas",
           "JPerp"]

TROTTER = {"reference": 0, "periodic": 1, 0: 0, 1: 1}


def JPerp(gamma, slices, temp):
    """J_perp = -PT/2 log(tanh(Gamma/PT)) with the reference's float32/float64 cast order
    (piqmc/qmc.pyx:95)."""
    return float(lib.piqmc_jperp(float(gamma), int(slices), ctypes.c_float(temp)))


def _check_confs(confs, nspins, slices):
    confs = _f64(confs, 2, "confs")
    if confs.shape != (nspins, slices):
        raise ValueError("confs must have shape (nspins, slices) = (%d, %d), got %s"
                         % (nspins, slices, confs.shape))
    return confs


def QuantumAnneal(sched, mcsteps, slices, temp, nspins, confs, nbs, rng, device=None):
    """Path-integral quantum annealing; @confs (float64[nspins, slices] of +-1, any strides) is
    updated in place.  Bit-exact replay of the reference (piqmc/qmc.pyx:30-136) including its
    as-shipped behaviour: float32 running energy difference reset once per slice (:134-135),
    Trotter neighbours `slices-1` and `1` for every slice (:115-117), libc rand() consumed
    lazily (:130-133; the process-global libc generator is left where the reference would leave
    it), spin order from @rng.permutation drawn exactly as the reference draws it.
    Raises ZeroDivisionError when slices*temp == 0, like the reference.  Returns None."""
    sched = _f64(sched, 1, "sched")
    nspins, slices, mcsteps = int(nspins), int(slices), int(mcsteps)
    confs = _check_confs(confs, nspins, slices)
    d = device or _dev.default_device()
    d.set_graph(nbs)
    if nspins != d.nspins:
        raise ValueError("nspins=%d but nbs describes %d spins" % (nspins, d.nspins))
    perms = _draw_perms(rng, nspins, sched.size * mcsteps)[None]
    spins = np.ascontiguousarray(_spins_i8(confs, "confs"))[None].copy()
    st = (RandState * 1)()
    st[0] = _dev.capture_libc_rand()
    d.qa_det(sched, mcsteps, slices, temp, spins, np.ascontiguousarray(perms), rstates=st)
    _dev.restore_libc_rand(st[0])
    confs[:, :] = spins[0]
    return None


def QuantumAnneal_dense(sched, mcsteps, slices, temp, nspins, confs, J, rng, device=None):
    """qmc.QuantumAnneal with a dense coupling matrix @J (float64[nspins, nspins]; off-diagonals are
    couplings, of which only the upper triangle is read, the diagonal holds the local fields).
    Bit-exact replay of the reference (piqmc/qmc.pyx:141-242), same as-shipped behaviour as
    QuantumAnneal (per-slice reset of the float32 energy difference, Trotter neighbours `slices-1`
    and `1`, lazy libc rand()); the couplings stay float64.  @confs is updated in place.  Returns None."""
    sched = _f64(sched, 1, "sched")
    J = _f64(J, 2, "J")
    nspins, slices, mcsteps = int(nspins), int(slices), int(mcsteps)
    confs = _check_confs(confs, nspins, slices)
    if J.shape != (nspins, nspins):
        raise ValueError("J must have shape (nspins, nspins) = (%d, %d), got %s" % (nspins, nspins, J.shape))
    d = device or _dev.default_device()
    perms = _draw_perms(rng, nspins, sched.size * mcsteps)[None]
    spins = np.ascontiguousarray(_spins_i8(confs, "confs"))[None].copy()
    st = (RandState * 1)()
    st[0] = _dev.capture_libc_rand()
    d.qa_dense_det(sched, mcsteps, slices, temp, spins, J, np.ascontiguousarray(perms), rstates=st)
    _dev.restore_libc_rand(st[0])
    confs[:, :] = spins[0]
    return None


def QuantumAnnealBatch(sched, mcsteps, slices, temp, nspins, confs, nbs, rngs, srand_seeds, device=None):
    """R independent QuantumAnneal runs in one launch (deterministic, bit-exact per replica).
    confs: [R, nspins, slices] (+-1, float64 or int8); rngs: one RandomState per replica;
    srand_seeds: the libc seed each replica's own process would have used.
    Returns (final int8[R,nspins,slices], consumed uint64[R])."""
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    d = device or _dev.default_device()
    d.set_graph(nbs)
    spins = np.ascontiguousarray(_spins_i8(np.asarray(confs), "confs")).copy()
    R = spins.shape[0]
    if spins.shape != (R, nspins, slices):
        raise ValueError("confs must have shape (R, nspins, slices)")
    perms = np.stack([_draw_perms(rng, nspins, sched.size * int(mcsteps)) for rng in rngs])
    st = _dev.rand_states(srand_seeds)
    consumed = d.qa_det(sched, int(mcsteps), int(slices), temp, spins, perms, rstates=st)
    return spins, consumed


def QuantumAnneal_parallel(sched, mcsteps, slices, temp, nspins, confs, nbs, nthreads=1, device=None,
                           trotter="reference"):
    """Same signature as the reference's OpenMP variant (piqmc/qmc.pyx:247-357): per-spin energy
    difference, no rng argument (uniforms ultimately come from the process-global libc stream:
    two rand() calls seed Philox), Trotter neighbours slices-1 and 1.  The reference runs spins
    of a slice concurrently with data races; here the same per-spin rule is applied colour class
    by colour class on the GPU, race-free.  @nthreads is accepted and ignored.  In place."""
    sched = _f64(sched, 1, "sched")
    confs = _check_confs(confs, int(nspins), int(slices))
    libc = ctypes.CDLL(None)
    seed = (libc.rand() << 31) | libc.rand()
    out = QuantumAnnealReplicas(sched, mcsteps, slices, temp, nspins,
                                np.ascontiguousarray(_spins_i8(confs, "confs").T)[None], nbs, seed,
                                device=device, trotter=trotter, energies=False, tile=False)
    confs[:, :] = _tools.UnpackWords(out["words"], int(slices))[0].T
    return None


def QuantumAnnealReplicas(sched, mcsteps, slices, temp, nspins, spins0, nbs, seed, order="natural",
                          color=None, replica0=0, trotter="reference", device=None, energies=True,
                          tile=True, nreplicas=None, download=True, words_out=None, global_moves=False,
                          per_word="auto", semantics="parallel"):
    """Production PIQMC: R replicas x `slices` Trotter slices x nspins, one uint64 word per
    (replica, spin) holding all slices, colour-class Metropolis sweeps with Philox4x32-10 keyed by
    (seed; spin, slice, sweep, replica0 + r), J_perp recomputed per schedule step.

    order: the sequential sweep each colour-class sweep is equivalent to --
        "natural"      spins 0..N-1 in every slice: the order of the reference's per-spin-reset
                       variant qmc.QuantumAnneal_parallel (qmc.pyx:320), whose residual-energy
                       statistics this mode reproduces;
        "permutation"  a fresh random permutation per sweep (qmc.pyx:100,136), from RandomState(seed);
        "checkerboard" fewest classes (fastest; at T << J its quench statistics differ measurably
                       from the natural-order sweep);
        int32[N] / int32[nsweeps,N]  explicit visiting order(s).
    color: explicit colour classes (overrides order).
    semantics: "parallel" (default) resets the energy difference per spin, like qmc.QuantumAnneal_parallel;
        "reference" is the AS-SHIPPED qmc.QuantumAnneal (qmc.pyx:98-136): the energy difference is reset once
        per slice sweep and carried over all spins in between -- markedly different statistics (residual energy
        1.00 per spin on inst_0_32x32 against 0.25).  It runs sequentially in the visiting order (order=
        "permutation" is the reference's; "natural" or explicit orders also work), one warp per replica, and
        takes one replica per word, the reference Trotter neighbours and no world-line moves.
    global_moves: also attempt a world-line move (flip the spin in all slices at once) after the
        local moves of every spin -- not in the reference; off by default.

    spins0: "resident" = take the replicas of the SA state resident on the device (after
            sa.AnnealReplicas(..., download=False); give nreplicas), or int8[R, nspins] copied to every slice (tile=True, the reference's
            np.tile(spinVector, (P,1)).T start), or int8[R, slices, nspins] (tile=False), or None
            for a Philox-generated random start (give nreplicas).
    words_out: optional C-contiguous uint64[nspins, R] host buffer (e.g. device.pinned_empty) that
            receives the packed state in device layout (one replica per word only).
    per_word: replicas per 64-bit word.  "auto": floor(64/slices) when slices <= 32 is a multiple of 4,
            the Trotter mode is the reference's, there are no world-line moves and the table kernel
            applies (maxnb <= 4) -- e.g. 3 replicas per word at P = 20; results are the same replica
            by replica as with one replica per word.
    Returns dict(words=uint64[R,nspins] (bit k <-> slice k, set <-> spin -1; a transposed view of
                 the spin-major buffer), energies=float64[R,slices] ClassicalIsingEnergy per slice)."""
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    slices = int(slices)
    if not 2 <= slices <= 64:
        raise ValueError("the packed colour path supports 2 <= slices <= 64 (got %d)" % slices)
    d = device or _dev.default_device()
    if semantics not in ("parallel", "reference"):
        raise ValueError("semantics must be 'parallel' or 'reference'")
    carry = semantics == "reference"
    if carry and (global_moves or TROTTER[trotter] != 0 or color is not None or per_word not in ("auto", 1)):
        raise ValueError("semantics='reference' takes a visiting order, the reference Trotter neighbours, no "
                         "world-line moves and one replica per word")
    orders = None
    if color is None:
        color, orders = _resolve_order(order, nbs, sched.size * int(mcsteps), int(nspins),
                                       int(seed) & 0xFFFFFFFF)
    if carry:
        if orders is None and not (isinstance(order, str) and order == "natural"):
            raise ValueError("semantics='reference' needs order='natural', 'permutation' or explicit per-sweep orders")
        per_word = 1
    import time
    t = [time.perf_counter()]
    d.set_graph(nbs, color)
    if int(nspins) != d.nspins:
        raise ValueError("nspins=%d but nbs describes %d spins" % (nspins, d.nspins))
    resident = isinstance(spins0, str) and spins0 == "resident"
    R = int(nreplicas) if (spins0 is None or resident) else int(np.asarray(spins0).shape[0])
    S = 1
    if per_word == "auto":
        if (slices <= 32 and slices % 4 == 0 and TROTTER[trotter] == 0 and not global_moves
                and d.maxnb <= 4 and d.variant != 1):
            S = 64 // slices
            # Chain pipeline (natural-order colourings): a warp walks a chain for 32 words whatever they
            # hold, so packing S replicas per word is S times the work at the same cost -- always pack.
            # Dataflow kernel: packing divides the rows by S.  That pays when a wavefront step (one colour
            # class, all rows) still offers several times the words the GPU holds in flight (~170k); with
            # many small classes the sweep is bound by the dependency chain and fewer rows only make it
            # worse (measured: tools/bench_configs.py).
            chain = orders is None and d.chain_info()[3]
            nclasses = (int(np.max(color)) + 1) if orders is None else 16
            per_step = max(1, int(nspins) // nclasses)
            while (S > 1 and not chain and d.variant < 2
                   and ((R + S - 1) // S < 32 or per_step * ((R + S - 1) // S) < 350000)):
                S -= 1
    else:
        S = int(per_word)
    rows = (R + S - 1) // S
    if not resident:
        d.state_alloc(rows, slices, S)
    t.append(time.perf_counter())
    if resident:
        d.state_replicas_to_slices(R, slices, S)
    elif spins0 is None:
        d.state_init_random(seed, replica0, tile=True)
    else:
        sp = np.asarray(spins0, dtype=np.int8)
        if rows * S != R:                               # pad the last word with +1 replicas (discarded below)
            sp = np.concatenate([sp, np.ones((rows * S - R,) + sp.shape[1:], dtype=np.int8)])
        d.state_upload_spins(sp, tile=tile)
    t.append(time.perf_counter())
    d.set_global_moves(global_moves)
    out = {"energies": None, "words": None}
    fused = bool(energies and download and not carry)      # anneal + energies + download as one library call
    try:
        if carry:
            d.qa_carry(sched, int(mcsteps), temp, seed, replica0=replica0, orders=orders)
        elif fused:
            out["energies"], out["words"] = d.qa_colour_results(
                sched, int(mcsteps), temp, seed, replica0=replica0, trotter=TROTTER[trotter], orders=orders,
                words_out=words_out if S == 1 else None)
        else:
            d.qa_colour(sched, int(mcsteps), temp, seed, replica0=replica0, trotter=TROTTER[trotter],
                        orders=orders)
    finally:
        d.set_global_moves(False)
    t.append(time.perf_counter())
    if fused:                                           # one library call: its own split of the time
        sw, res = d.last_phase_seconds()
        t[-1] = t[-2] + sw
        t += [t[-1], t[-1] + res]
    elif energies and download:
        out["energies"], out["words"] = d.results(words_out if S == 1 else None)
        t.append(time.perf_counter())
    else:
        if energies:
            out["energies"] = d.energy(download=download)
        t.append(time.perf_counter())
        if download:
            out["words"] = d.state_download_words(out=words_out if S == 1 else None)
        else:
            d.synchronize()
    t.append(time.perf_counter())
    if rows * S != R:                                   # drop the padding replicas of the last word
        for key in ("energies", "words"):
            if isinstance(out[key], np.ndarray):
                out[key] = out[key][:R]
    out["per_word"] = S
    out["seconds"] = dict(zip(("graph+alloc", "upload", "sweeps", "energy", "download"), np.diff(t).tolist()))
    if fused:                                           # energies and download overlap: one figure for both
        out["seconds"]["energy+download"] = out["seconds"].pop("download")
        out["seconds"].pop("energy")
    return out
