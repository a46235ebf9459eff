"""CPU tests of the C-ABI boundary: libpiqmc_b200.so loads, exports every symbol declared in
include/piqmc_b200.h, and its host-only entry points behave.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported_and_bound():
    from piqmc import _lib
    hdr = open(os.path.join(ROOT, "include", "piqmc_b200.h")).read()
    declared = set(re.findall(r"\b(piqmc_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"piqmc_ctx"}
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(_lib.lib, name), "libpiqmc_b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.lib.piqmc_version() == 100


def test_host_rand_matches_live_libc():
    from piqmc import _lib
    libc = ctypes.CDLL("libc.so.6")
    for seed in (1, 0, 77, 4000000000):
        s = _lib.RandState()
        _lib.lib.piqmc_rand_seed(ctypes.byref(s), seed)
        libc.srand(seed)
        assert [libc.rand() for _ in range(500)] == \
            [_lib.lib.piqmc_rand_next(ctypes.byref(s)) for _ in range(500)]


def test_capture_and_restore_libc_stream():
    from piqmc import device
    libc = ctypes.CDLL(None)
    libc.srand(99)
    for _ in range(17):
        libc.rand()
    st = device.capture_libc_rand()
    expect = [libc.rand() for _ in range(50)]           # capture must not disturb the stream
    from piqmc._lib import lib
    mine = [lib.piqmc_rand_next(ctypes.byref(st)) for _ in range(50)]
    assert mine == expect
    nxt = [libc.rand() for _ in range(5)]
    # rewind libc to 20 draws after the capture point
    libc.srand(99)
    for _ in range(17):
        libc.rand()
    st2 = device.capture_libc_rand()
    for _ in range(50):
        lib.piqmc_rand_next(ctypes.byref(st2))
    libc.srand(12345)                                   # scramble
    device.restore_libc_rand(st2)
    assert [libc.rand() for _ in range(5)] == nxt


def test_jperp_matches_oracle():
    from piqmc.qmc import JPerp
    for gamma in np.linspace(1.5, 1e-8, 23):
        for P, T in ((20, 0.01), (5, 0.01), (64, 0.01), (10, 0.7)):
            assert JPerp(gamma, P, T) == O.jperp(gamma, P, T)


def test_no_device_is_a_loud_error():
    from piqmc import _lib, device
    n = ctypes.c_int(-1)
    rc = _lib.lib.piqmc_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError):
        device.Device(0)
    assert "cuda" in _lib.last_error().lower()


def test_argument_validation_mirrors_reference_errors():
    import piqmc.qmc as qmc
    import piqmc.sa as sa
    nbs = np.zeros((4, 2, 2))
    with pytest.raises(ValueError):                     # Cython: Buffer dtype mismatch
        sa.Anneal(np.linspace(1, 0.1, 3).astype(np.float32), 1, np.ones(4), nbs, np.random.RandomState(0))
    with pytest.raises(ValueError):                     # wrong ndim
        qmc.QuantumAnneal(np.linspace(1, 0.1, 3), 1, 3, 0.1, 4, np.ones(4), nbs, np.random.RandomState(0))
    with pytest.raises(ValueError):
        sa.Anneal_multispin(np.linspace(1, 0.1, 3), 1, np.zeros((63, 4)), nbs, np.random.RandomState(0))


def test_dense_wrappers_validate_like_cython_memoryviews():
    """QuantumAnneal_dense / Anneal_dense take np.float_t[:, :] J (qmc.pyx:149, sa.pyx:133): float64
    only, two dimensions, and the shapes must agree -- checked before any device is touched."""
    import piqmc.qmc as qmc
    import piqmc.sa as sa
    rng = np.random.RandomState(0)
    sched = np.linspace(1.0, 0.1, 3)
    sv = np.ones(4)
    with pytest.raises(ValueError):
        sa.Anneal_dense(sched, 1, sv, np.zeros((4, 4), dtype=np.float32), rng)
    with pytest.raises(ValueError):
        sa.Anneal_dense(sched, 1, sv, np.zeros(16), rng)
    with pytest.raises(ValueError):
        sa.Anneal_dense(sched, 1, sv, np.zeros((5, 5)), rng)
    confs = np.ones((4, 3))
    with pytest.raises(ValueError):
        qmc.QuantumAnneal_dense(sched, 1, 3, 0.1, 4, confs, np.zeros((4, 5)), rng)
    with pytest.raises(ValueError):
        qmc.QuantumAnneal_dense(sched, 1, 3, 0.1, 4, confs.astype(np.float32), np.zeros((4, 4)), rng)
