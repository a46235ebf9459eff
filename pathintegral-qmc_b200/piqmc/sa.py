"""piqmc.sa -- thermal annealing of Ising models on B200.

Mirror of the reference module piqmc/sa.pyx (same names, argument order and in-place
behaviour); the sweeps run in libpiqmc_b200.so on the GPU.  There is no CPU fallback.

  Anneal            drop-in, bit-exact replay of sa.Anneal (sa.pyx:50-120)
  Anneal_multispin  drop-in, bit-exact replay of sa.Anneal_multispin (sa.pyx:282-405)
  Anneal_parallel   same signature as the OpenMP variant (sa.pyx:193-265); runs the
                    colour-parallel kernel (statistically equivalent, not bit-identical)
  AnnealBatch       R replicas of Anneal in one launch (deterministic)
  AnnealReplicas    production path: colour-parallel, 64 replicas per word, Philox
  ClassicalIsingEnergy   device reduction (sa.pyx:25-44)
"""
import ctypes

import numpy as np

from . import device as _dev
from . import tools as _tools
from ._lib import RandState, lib

__all__ = ["ClassicalIsingEnergy", "Anneal", "Anneal_dense", "Anneal_parallel", "Anneal_multispin", "AnnealBatch",
           "AnnealReplicas"]


def _f64(a, ndim, name):
    """The reference takes Cython memoryviews np.float_t[:...]: float64 only, any strides."""
    a = np.asarray(a) if not isinstance(a, np.ndarray) else a
    if a.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'float_t' but got '%s' (%s)" % (a.dtype, name))
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d) (%s)"
                         % (ndim, a.ndim, name))
    return a


def _spins_i8(a, name):
    s = np.rint(a).astype(np.int8)
    if not np.all((s == 1) | (s == -1)):
        raise ValueError("%s must hold +-1 spins" % name)
    return s


def _draw_perms(rng, nspins, nsweeps):
    """The reference's calls (sa.pyx:91-92,120): rng.permutation(range(N)) once, then
    rng.permutation(previous) after every sweep -- nsweeps+1 calls, the last result unused."""
    out = np.empty((nsweeps, nspins), dtype=np.int32)
    p = rng.permutation(range(nspins))
    for s in range(nsweeps):
        out[s] = p
        p = rng.permutation(p)
    return out


def ClassicalIsingEnergy(spins, J, device=None):
    """Energy of configuration @spins (+-1) in the Ising system @J (scipy sparse; off-diagonals
    are couplings stored once, diagonal holds local fields):  -s.(J_off s) - sum_i J_ii s_i.
    Reference: piqmc/sa.pyx:25-44 (which densifies J: O(N^2)); here an O(nnz) device reduction
    in float64."""
    d = device or _dev.default_device()
    coo = J.tocoo()
    s = _spins_i8(np.asarray(spins, dtype=np.float64).reshape(1, -1), "spins")
    return float(d.energy_coo(coo.shape[0], coo.row, coo.col, coo.data, s)[0])


def Anneal(sched, mcsteps, svec, nbs, rng, device=None):
    """Thermal annealing according to @sched with @mcsteps sweeps per step; @svec (float64 +-1)
    is updated in place.  Bit-exact replay of the reference (piqmc/sa.pyx:50-120): spin order from
    @rng.permutation exactly as the reference draws it, Metropolis uniforms from the process-global
    libc rand() stream (consumed lazily; the libc generator is left exactly where the reference
    would leave it).  Returns None."""
    sched = _f64(sched, 1, "sched")
    svec = _f64(svec, 1, "svec")
    d = device or _dev.default_device()
    d.set_graph(nbs)
    n = svec.size
    if n != d.nspins:
        raise ValueError("svec has %d spins but nbs describes %d" % (n, d.nspins))
    nsweeps = sched.size * int(mcsteps)
    perms = _draw_perms(rng, n, nsweeps)[None]
    spins = _spins_i8(svec, "svec")[None].copy()
    st = (RandState * 1)()
    st[0] = _dev.capture_libc_rand()
    d.sa_det(sched, int(mcsteps), spins, np.ascontiguousarray(perms), rstates=st)
    _dev.restore_libc_rand(st[0])
    svec[:] = spins[0]
    return None


def Anneal_dense(sched, mcsteps, svec, J, rng, device=None):
    """sa.Anneal with a dense coupling matrix @J (float64[N, N]; upper triangle + diagonal = local
    fields).  Bit-exact replay of the reference (piqmc/sa.pyx:126-187); note that this variant
    accepts on ediff > 0 where sa.Anneal accepts on ediff >= 0, and keeps the couplings in
    float64.  @svec is updated in place.  Returns None."""
    sched = _f64(sched, 1, "sched")
    svec = _f64(svec, 1, "svec")
    J = _f64(J, 2, "J")
    n = svec.size
    if J.shape != (n, n):
        raise ValueError("J must have shape (%d, %d), got %s" % (n, n, J.shape))
    d = device or _dev.default_device()
    perms = _draw_perms(rng, n, sched.size * int(mcsteps))[None]
    spins = _spins_i8(svec, "svec")[None].copy()
    st = (RandState * 1)()
    st[0] = _dev.capture_libc_rand()
    d.sa_dense_det(sched, int(mcsteps), spins, J, np.ascontiguousarray(perms), rstates=st)
    _dev.restore_libc_rand(st[0])
    svec[:] = spins[0]
    return None


def AnnealBatch(sched, mcsteps, svecs, nbs, rngs, srand_seeds, device=None):
    """R independent sa.Anneal runs in one launch.  svecs float64/int8 [R,N] (+-1), rngs: one
    RandomState per replica, srand_seeds: the libc seed each replica would have called srand()
    with in its own process.  Returns (final int8[R,N], consumed uint64[R])."""
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    d = device or _dev.default_device()
    d.set_graph(nbs)
    spins = _spins_i8(np.asarray(svecs), "svecs").copy()
    R, n = spins.shape
    nsweeps = sched.size * int(mcsteps)
    perms = np.stack([_draw_perms(rng, n, nsweeps) for rng in rngs])
    st = _dev.rand_states(srand_seeds)
    consumed = d.sa_det(sched, int(mcsteps), spins, perms, rstates=st)
    return spins, consumed


def Anneal_parallel(sched, mcsteps, svec, nbs, nthreads=1, device=None):
    """Same signature as the reference's OpenMP variant (piqmc/sa.pyx:193-265), which updates all
    spins of a sweep concurrently without any ordering.  Here the sweep is colour-parallel on the
    GPU (race-free); @nthreads is accepted and ignored.  Like the reference it takes its
    randomness from the process-global libc stream: two rand() calls seed the Philox generator.
    Metropolis shortcut is `> 0` in the reference's variant and `>= 0` in sa.Anneal; this uses the
    sa.Anneal rule.  Not bit-identical to the reference (which is racy by design)."""
    sched = _f64(sched, 1, "sched")
    svec = _f64(svec, 1, "svec")
    libc = ctypes.CDLL(None)
    seed = (libc.rand() << 31) | libc.rand()
    out = AnnealReplicas(sched, mcsteps, _spins_i8(svec, "svec")[None], nbs, seed, order="natural",
                         device=device, energies=False)
    svec[:] = out["spins"][0]
    return None


def Anneal_multispin(sched, mcsteps, svec_mat, nbs, rng, device=None):
    """64 simultaneous anneals, multispin-coded.  @svec_mat is float64[64, N] of BITS (0 <-> +1,
    1 <-> -1), updated in place.  Bit-exact replay of piqmc/sa.pyx:282-405: replica k in bit 63-k,
    float64 energy differences, acceptance exp(ediff/temp) > rng.rand(64) with no ediff>0 shortcut,
    rng consumed in the reference's order (rand(64), permutation, then rand(64) after every
    attempt and permutation after every sweep).  The reference's unpack loop runs one column too
    far (sa.pyx:402) and corrupts column 0 of rows 1..63; this does not."""
    sched = _f64(sched, 1, "sched")
    svec_mat = _f64(svec_mat, 2, "svec_mat")
    if svec_mat.shape[0] != 64:
        raise ValueError("svec_mat must have 64 rows")
    d = device or _dev.default_device()
    d.set_graph(nbs)
    n = svec_mat.shape[1]
    if n != d.nspins:
        raise ValueError("svec_mat has %d spins but nbs describes %d" % (n, d.nspins))
    mcsteps = int(mcsteps)
    bits = (svec_mat != 0.0).astype(np.uint64)
    shifts = (63 - np.arange(64, dtype=np.uint64)).reshape(64, 1)
    words = np.bitwise_or.reduce(bits << shifts, axis=0).reshape(1, n)
    carry = rng.rand(64)                                   # sa.pyx:331
    perm = rng.permutation(range(n))                       # sa.pyx:336-337
    # sweeps are shipped in chunks so the uniform table stays bounded (64 doubles per attempt)
    chunk = max(1, (64 << 20) // (n * 64 * 8))
    nsweeps = sched.size * mcsteps
    done = 0
    while done < nsweeps:
        m = min(chunk, nsweeps - done)
        perms = np.empty((1, m, n), dtype=np.int32)
        rands = np.empty((m, n, 64), dtype=np.float64)
        for s in range(m):
            perms[0, s] = perm
            blocks = rng.rand(n * 64).reshape(n, 64)       # sa.pyx:398, one block per attempt
            rands[s, 0] = carry
            rands[s, 1:] = blocks[:-1]
            carry = blocks[-1]
            perm = rng.permutation(perm)                   # sa.pyx:400
        # the chunk may start/end inside a temperature step: expand to one temperature per sweep
        temps = np.repeat(sched, mcsteps)[done:done + m]
        d.sa_multispin_det(temps, 1, words, perms, rands)
        done += m
    out = ((words[0][None, :] >> shifts) & np.uint64(1)).astype(np.float64)
    svec_mat[:, :] = out
    return None


def _resolve_order(order, nbs, nsweeps, nspins, order_seed):
    """-> (color or None, orders or None) for the `order` option of the production paths."""
    if isinstance(order, str):
        if order == "permutation":
            prng = np.random.RandomState(order_seed)
            return None, np.stack([prng.permutation(nspins) for _ in range(nsweeps)]).astype(np.int32)
        if order in ("natural", "checkerboard"):
            return _tools.ColourGraph(nbs, order), None
        raise ValueError("order must be 'natural', 'checkerboard', 'permutation' or an int array")
    o = np.asarray(order)
    if o.ndim == 1:
        return _tools.ColourGraph(nbs, o), None
    return None, np.ascontiguousarray(o, dtype=np.int32)


def AnnealReplicas(sched, mcsteps, spins0, nbs, seed, order="permutation", color=None, row0=0,
                   device=None, energies=True, nreplicas=None, download=True):
    """Production SA: R independent replicas, 64 per uint64 word, colour-class Metropolis with
    Philox4x32-10 uniforms keyed by (seed; spin, replica, sweep).  sa.Anneal rules (float32 local
    field in table order, `>= 0` shortcut, temperature schedule).

    order: the sequential sweep each colour-class sweep is equivalent to --
        "permutation"  a fresh random permutation per sweep, as sa.Anneal (sa.pyx:100,120); the
                       permutations come from RandomState(seed) and are shared by all replicas;
        "natural"      0..N-1, as sa.Anneal_parallel with one thread (sa.pyx:248);
        "checkerboard" fewest classes (fastest; not equivalent to a reference order);
        int32[N] / int32[nsweeps,N]  explicit visiting order(s).
    color: explicit colour classes (overrides order).
    spins0: int8[R,N] (+-1) or None for a Philox-generated random start (then give nreplicas).
    download=False leaves the result on the device only (for qmc.QuantumAnnealReplicas(spins0=
    "resident"), the SA pre-anneal -> PIQMC hand-over of examples/spinglass32.py:124-127).
    Returns dict(spins=int8[R,N], energies=float64[R] or None, words=uint64[G,N])."""
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    d = device or _dev.default_device()
    nsweeps = sched.size * int(mcsteps)
    nspins = np.asarray(nbs).shape[0]
    orders = None
    if color is None:
        color, orders = _resolve_order(order, nbs, nsweeps, nspins, int(seed) & 0xFFFFFFFF)
    d.set_graph(nbs, color)
    n = d.nspins
    R = int(nreplicas) if spins0 is None else int(np.asarray(spins0).shape[0])
    G = (R + 63) // 64
    d.state_alloc(G, 64)
    if spins0 is None:
        d.state_init_random(seed, row0, tile=False)
    else:
        s = np.ones((G * 64, n), dtype=np.int8)
        s[:R] = _spins_i8(np.asarray(spins0), "spins0")
        d.state_upload_spins(s.reshape(G, 64, n), tile=False)
    d.sa_colour(sched, int(mcsteps), seed, row0=row0, orders=orders)
    en = d.energy().reshape(-1)[:R] if energies else None
    if not download:
        d.synchronize()
        return {"spins": None, "energies": en, "words": None}
    words = d.state_download_words()
    spins = _tools.UnpackWords(words, 64).reshape(G * 64, n)[:R]
    return {"spins": spins, "energies": en, "words": words}
