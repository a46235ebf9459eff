"""ctypes front-end of the CPU oracle (oracle/piqmc_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libpiqmc_oracle.so")

_lib = None

c_dp = ctypes.POINTER(ctypes.c_double)
c_fp = ctypes.POINTER(ctypes.c_float)
c_ip = ctypes.POINTER(ctypes.c_int32)
c_bp = ctypes.POINTER(ctypes.c_int8)
c_up = ctypes.POINTER(ctypes.c_uint32)


class GlibcRandState(ctypes.Structure):
    _fields_ = [("r", ctypes.c_uint32 * 31), ("f", ctypes.c_int32), ("b", ctypes.c_int32)]


def build(force=False):
    src = os.path.join(HERE, "piqmc_oracle.c")
    if (not force and os.path.exists(SO)
            and (not os.path.exists(src) or os.path.getmtime(SO) >= os.path.getmtime(src))):
        return SO
    subprocess.check_call(["make", "-s", "-C", HERE, "-B", "_build/libpiqmc_oracle.so"])
    return SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(SO)
        L.oracle_jperp.restype = ctypes.c_float
        L.oracle_jperp.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_float]
        usrc = [ctypes.c_int, c_dp, ctypes.c_uint64, ctypes.POINTER(GlibcRandState)]
        L.oracle_qa_reference.restype = ctypes.c_uint64
        L.oracle_qa_reference.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                          ctypes.c_int, c_dp, ctypes.c_long, ctypes.c_long,
                                          c_dp, ctypes.c_int, c_ip] + usrc
        L.oracle_sa_reference.restype = ctypes.c_uint64
        L.oracle_sa_reference.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          c_dp, ctypes.c_long, c_dp, ctypes.c_int, c_ip] + usrc
        L.oracle_qa_dense.restype = ctypes.c_uint64
        L.oracle_qa_dense.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                      ctypes.c_int, c_dp, ctypes.c_long, ctypes.c_long,
                                      c_dp, ctypes.c_long, ctypes.c_long, c_ip] + usrc
        L.oracle_sa_dense.restype = ctypes.c_uint64
        L.oracle_sa_dense.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      c_dp, ctypes.c_long, c_dp, ctypes.c_long, ctypes.c_long, c_ip] + usrc
        L.oracle_qa_parallel1.restype = ctypes.c_uint64
        L.oracle_qa_parallel1.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                          ctypes.c_int, c_dp, ctypes.c_long, ctypes.c_long,
                                          c_dp, ctypes.c_int] + usrc
        L.oracle_sa_parallel1.restype = ctypes.c_uint64
        L.oracle_sa_parallel1.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          c_dp, ctypes.c_long, c_dp, ctypes.c_int] + usrc
        L.oracle_sa_multispin.restype = None
        L.oracle_sa_multispin.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          c_dp, ctypes.c_long, ctypes.c_long, c_dp, ctypes.c_int,
                                          c_ip, c_dp]
        L.oracle_energy.restype = ctypes.c_double
        L.oracle_energy.argtypes = [ctypes.c_int, c_ip, c_ip, c_dp, c_dp, ctypes.c_long]
        L.oracle_glibc_srand.restype = None
        L.oracle_glibc_srand.argtypes = [ctypes.POINTER(GlibcRandState), ctypes.c_uint]
        L.oracle_glibc_rand.restype = ctypes.c_int32
        L.oracle_glibc_rand.argtypes = [ctypes.POINTER(GlibcRandState)]
        L.oracle_philox4x32_10.restype = None
        L.oracle_philox4x32_10.argtypes = [c_up, c_up, c_up]
        L.oracle_colour_u32.restype = ctypes.c_uint32
        L.oracle_colour_u32.argtypes = [ctypes.c_uint64] + [ctypes.c_uint32] * 4
        L.oracle_colour_initbit.restype = ctypes.c_uint32
        L.oracle_colour_initbit.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32]
        L.oracle_colour_thresh.restype = ctypes.c_uint32
        L.oracle_colour_thresh.argtypes = [ctypes.c_float]
        L.oracle_qa_carry.restype = None
        L.oracle_qa_carry.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                      ctypes.c_int, c_ip, c_fp, ctypes.c_int, c_bp, ctypes.c_uint64, ctypes.c_uint32,
                                      ctypes.c_uint32, c_ip]
        L.oracle_qa_colour.restype = None
        L.oracle_qa_colour.argtypes = [c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                       ctypes.c_int, ctypes.c_int, c_ip, c_fp, ctypes.c_int, c_ip,
                                       ctypes.c_int, c_bp, ctypes.c_uint64, ctypes.c_uint32,
                                       ctypes.c_uint32, ctypes.c_int, c_ip]
        L.oracle_sa_colour.restype = None
        L.oracle_sa_colour.argtypes = [c_dp, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, c_ip, c_fp, ctypes.c_int, c_ip,
                                       ctypes.c_int, c_bp, ctypes.c_uint64, ctypes.c_uint32,
                                       ctypes.c_uint32, c_ip]
        L.oracle_energy_ell.restype = ctypes.c_double
        L.oracle_energy_ell.argtypes = [ctypes.c_int, ctypes.c_int, c_ip, c_fp, c_bp, ctypes.c_long]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def _usrc(uniforms, gstate):
    """(kind, table, ntable, gstate) for the C uniform source."""
    if uniforms is not None:
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        return 1, u, _p(u, c_dp), u.size, None
    if gstate is not None:
        return 2, None, None, 0, ctypes.byref(gstate)
    return 0, None, None, 0, None


# ------------------------------------------------------------------ helpers shared by tests
def make_perms(rng, nspins, nsweeps):
    """The permutations the reference draws (qmc.pyx:90-91,136 / sa.pyx:91-92,120): one
    rng.permutation(range(N)) up front, then rng.permutation(prev) after every sweep -- so
    nsweeps+1 calls in total, the last result unused.  Returns int32[nsweeps, N]."""
    out = np.empty((nsweeps, nspins), dtype=np.int32)
    p = rng.permutation(range(nspins))
    for s in range(nsweeps):
        out[s] = p
        p = rng.permutation(p)
    return out


def nbs_to_ell(nbs):
    """(idx int32[N,maxnb], J float32[N,maxnb]) with the casts the kernels apply
    (qmc.pyx:104,106: int(...) and C-float narrowing)."""
    nbs = np.asarray(nbs, dtype=np.float64)
    return (np.ascontiguousarray(nbs[:, :, 0].astype(np.int32)),
            np.ascontiguousarray(nbs[:, :, 1].astype(np.float32)))


def glibc_state(seed):
    g = GlibcRandState()
    lib().oracle_glibc_srand(ctypes.byref(g), seed)
    return g


# ------------------------------------------------------------------ reference semantics
def jperp(gamma, slices, temp):
    return lib().oracle_jperp(float(gamma), int(slices), ctypes.c_float(temp))


def qa_reference(sched, mcsteps, slices, temp, nspins, confs, nbs, perms, uniforms=None, gstate=None):
    """In-place on confs (float64 [N,P], any strides).  Returns #uniforms consumed."""
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    nbs = np.ascontiguousarray(nbs, dtype=np.float64)
    perms = np.ascontiguousarray(perms, dtype=np.int32)
    assert confs.dtype == np.float64 and confs.shape == (nspins, slices)
    assert perms.shape == (sched.size * mcsteps, nspins)
    kind, keep, tab, ntab, g = _usrc(uniforms, gstate)
    return lib().oracle_qa_reference(_p(sched, c_dp), sched.size, mcsteps, slices, ctypes.c_float(temp),
                                     nspins, _p(confs, c_dp), confs.strides[0] // 8, confs.strides[1] // 8,
                                     _p(nbs, c_dp), nbs.shape[1], _p(perms, c_ip), kind, tab, ntab, g)


def sa_reference(sched, mcsteps, svec, nbs, perms, uniforms=None, gstate=None):
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    nbs = np.ascontiguousarray(nbs, dtype=np.float64)
    perms = np.ascontiguousarray(perms, dtype=np.int32)
    assert svec.dtype == np.float64 and svec.ndim == 1
    kind, keep, tab, ntab, g = _usrc(uniforms, gstate)
    return lib().oracle_sa_reference(_p(sched, c_dp), sched.size, mcsteps, svec.size,
                                     _p(svec, c_dp), svec.strides[0] // 8, _p(nbs, c_dp), nbs.shape[1],
                                     _p(perms, c_ip), kind, tab, ntab, g)


def qa_dense(sched, mcsteps, slices, temp, nspins, confs, J, perms, uniforms=None, gstate=None):
    """qmc.QuantumAnneal_dense (qmc.pyx:141-242), in place on confs float64[N,P]; J float64[N,N]."""
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    J = np.asarray(J, dtype=np.float64)
    perms = np.ascontiguousarray(perms, dtype=np.int32)
    assert confs.dtype == np.float64 and confs.shape == (nspins, slices) and J.shape == (nspins, nspins)
    assert perms.shape == (sched.size * mcsteps, nspins)
    kind, keep, tab, ntab, g = _usrc(uniforms, gstate)
    return lib().oracle_qa_dense(_p(sched, c_dp), sched.size, mcsteps, slices, ctypes.c_float(temp),
                                 nspins, _p(confs, c_dp), confs.strides[0] // 8, confs.strides[1] // 8,
                                 _p(J, c_dp), J.strides[0] // 8, J.strides[1] // 8, _p(perms, c_ip),
                                 kind, tab, ntab, g)


def sa_dense(sched, mcsteps, svec, J, perms, uniforms=None, gstate=None):
    """sa.Anneal_dense (sa.pyx:126-187), in place on svec float64[N]; J float64[N,N]."""
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    J = np.asarray(J, dtype=np.float64)
    perms = np.ascontiguousarray(perms, dtype=np.int32)
    assert svec.dtype == np.float64 and svec.ndim == 1 and J.shape == (svec.size, svec.size)
    kind, keep, tab, ntab, g = _usrc(uniforms, gstate)
    return lib().oracle_sa_dense(_p(sched, c_dp), sched.size, mcsteps, svec.size,
                                 _p(svec, c_dp), svec.strides[0] // 8, _p(J, c_dp), J.strides[0] // 8,
                                 J.strides[1] // 8, _p(perms, c_ip), kind, tab, ntab, g)


def qa_parallel1(sched, mcsteps, slices, temp, nspins, confs, nbs, uniforms=None, gstate=None):
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    nbs = np.ascontiguousarray(nbs, dtype=np.float64)
    assert confs.dtype == np.float64 and confs.shape == (nspins, slices)
    kind, keep, tab, ntab, g = _usrc(uniforms, gstate)
    return lib().oracle_qa_parallel1(_p(sched, c_dp), sched.size, mcsteps, slices, ctypes.c_float(temp),
                                     nspins, _p(confs, c_dp), confs.strides[0] // 8, confs.strides[1] // 8,
                                     _p(nbs, c_dp), nbs.shape[1], kind, tab, ntab, g)


def sa_parallel1(sched, mcsteps, svec, nbs, uniforms=None, gstate=None):
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    nbs = np.ascontiguousarray(nbs, dtype=np.float64)
    kind, keep, tab, ntab, g = _usrc(uniforms, gstate)
    return lib().oracle_sa_parallel1(_p(sched, c_dp), sched.size, mcsteps, svec.size,
                                     _p(svec, c_dp), svec.strides[0] // 8, _p(nbs, c_dp), nbs.shape[1],
                                     kind, tab, ntab, g)


def sa_multispin(sched, mcsteps, bits, nbs, perms, rands):
    """bits float64[64,N] of 0/1, in place.  rands float64[nsweeps*N, 64]."""
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    nbs = np.ascontiguousarray(nbs, dtype=np.float64)
    perms = np.ascontiguousarray(perms, dtype=np.int32)
    rands = np.ascontiguousarray(rands, dtype=np.float64)
    assert bits.dtype == np.float64 and bits.shape[0] == 64
    n = bits.shape[1]
    assert rands.size == sched.size * mcsteps * n * 64
    lib().oracle_sa_multispin(_p(sched, c_dp), sched.size, mcsteps, n, _p(bits, c_dp),
                              bits.strides[0] // 8, bits.strides[1] // 8, _p(nbs, c_dp), nbs.shape[1],
                              _p(perms, c_ip), _p(rands, c_dp))


def energy(J, spins):
    """ClassicalIsingEnergy(spins, J) for a scipy sparse J holding each bond once."""
    coo = J.tocoo()
    row = np.ascontiguousarray(coo.row, dtype=np.int32)
    col = np.ascontiguousarray(coo.col, dtype=np.int32)
    val = np.ascontiguousarray(coo.data, dtype=np.float64)
    s = np.ascontiguousarray(spins, dtype=np.float64)
    return lib().oracle_energy(val.size, _p(row, c_ip), _p(col, c_ip), _p(val, c_dp), _p(s, c_dp), 1)


# ------------------------------------------------------------------ colour semantics
def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(_p(c, c_up), _p(k, c_up), _p(o, c_up))
    return o


def colour_thresh(x):
    return lib().oracle_colour_thresh(ctypes.c_float(x))


def colour_init_spins(seed, replica0, nreplicas, nspins):
    """int8[R,N] of +-1: the device-side initial state of the colour path."""
    L = lib()
    out = np.empty((nreplicas, nspins), dtype=np.int8)
    for r in range(nreplicas):
        for i in range(nspins):
            out[r, i] = -1 if L.oracle_colour_initbit(seed, replica0 + r, i) else 1
    return out


def _orders(orders, nsweeps, n):
    if orders is None:
        return None, None
    o = np.ascontiguousarray(orders, dtype=np.int32)
    assert o.shape == (nsweeps, n)
    return o, _p(o, c_ip)


def qa_colour(sched, mcsteps, slices, temp, idx, J, color, spins, seed, replica0=0, sweep0=0, trotter=0,
              orders=None, global_moves=False):
    """spins int8[R,N,P] in place.  orders int32[nsweeps,N]: sequential sweeps in those visiting
    orders instead of colour classes."""
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    J = np.ascontiguousarray(J, dtype=np.float32)
    color = np.ascontiguousarray(color, dtype=np.int32)
    assert spins.dtype == np.int8 and spins.flags.c_contiguous and spins.ndim == 3
    R, N, P = spins.shape
    assert P == slices and idx.shape == J.shape == (N, idx.shape[1])
    ctypes.c_int.in_dll(lib(), "oracle_colour_global_moves").value = 1 if global_moves else 0
    lib().oracle_qa_colour(_p(sched, c_dp), sched.size, mcsteps, slices, ctypes.c_float(temp),
                           N, idx.shape[1], _p(idx, c_ip), _p(J, c_fp), int(color.max()) + 1,
                           _p(color, c_ip), R, _p(spins, c_bp), seed, replica0, sweep0, trotter,
                           _orders(orders, sched.size * mcsteps, N)[1])


def qa_carry(sched, mcsteps, slices, temp, idx, J, spins, seed, replica0=0, sweep0=0, orders=None):
    """The as-shipped QuantumAnneal semantics (per-slice energy carry) with Philox uniforms; spins int8[R,N,P]
    in place; orders int32[nsweeps,N] or None (0..N-1)."""
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    J = np.ascontiguousarray(J, dtype=np.float32)
    assert spins.dtype == np.int8 and spins.flags.c_contiguous and spins.ndim == 3
    R, N, P = spins.shape
    assert P == slices and idx.shape == J.shape == (N, idx.shape[1])
    lib().oracle_qa_carry(_p(sched, c_dp), sched.size, mcsteps, slices, ctypes.c_float(temp), N, idx.shape[1],
                          _p(idx, c_ip), _p(J, c_fp), R, _p(spins, c_bp), seed, replica0, sweep0,
                          _orders(orders, sched.size * mcsteps, N)[1])


def sa_colour(sched, mcsteps, idx, J, color, spins, seed, row0=0, sweep0=0, orders=None):
    """spins int8[R,N] in place."""
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    J = np.ascontiguousarray(J, dtype=np.float32)
    color = np.ascontiguousarray(color, dtype=np.int32)
    assert spins.dtype == np.int8 and spins.flags.c_contiguous and spins.ndim == 2
    R, N = spins.shape
    lib().oracle_sa_colour(_p(sched, c_dp), sched.size, mcsteps, N, idx.shape[1], _p(idx, c_ip),
                           _p(J, c_fp), int(color.max()) + 1, _p(color, c_ip), R, _p(spins, c_bp),
                           seed, row0, sweep0, _orders(orders, sched.size * mcsteps, N)[1])


def energy_ell(idx, J, s):
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    J = np.ascontiguousarray(J, dtype=np.float32)
    s = np.ascontiguousarray(s, dtype=np.int8)
    return lib().oracle_energy_ell(s.size, idx.shape[1], _p(idx, c_ip), _p(J, c_fp), _p(s, c_bp), 1)


# ------------------------------------------------------------------ the compiled reference
def ref():
    """The reference's own Cython modules (oracle/_ref/piqmc_ref), or None if not built."""
    import importlib
    import sys
    from . import build_ref
    try:
        build_ref.build()
    except Exception:
        if not build_ref.built():
            return None
    d = os.path.join(HERE, "_ref")
    if d not in sys.path:
        sys.path.insert(0, d)
    return importlib.import_module("piqmc_ref")


def colour_counters(reset=False):
    """(attempts, attempts that needed a uniform) of the colour-semantics functions so far."""
    arr = (ctypes.c_uint64 * 2).in_dll(lib(), "oracle_colour_counters")
    out = (int(arr[0]), int(arr[1]))
    if reset:
        arr[0] = arr[1] = 0
    return out
