"""ctypes binding of libpiqmc_b200.so (include/piqmc_b200.h).

The library is the product: there is NO CPU fallback.  If the shared object is missing or
cannot be loaded, importing this module raises ImportError; if no CUDA device is present the
first device call raises RuntimeError.
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("PIQMC_B200_LIB") or os.path.join(HERE, "libpiqmc_b200.so")
CSRC = os.path.join(os.path.dirname(HERE), "csrc")

OK, EINVAL, EZERODIV, ECUDA, ENOGRAPH, ENOSTATE, ENOMEM = 0, -1, -2, -3, -4, -5, -6


class RandState(ctypes.Structure):
    """piqmc_rand_state: glibc rand() TYPE_3 generator state."""
    _fields_ = [("r", ctypes.c_uint32 * 31), ("f", ctypes.c_int32), ("b", ctypes.c_int32)]


c_void = ctypes.c_void_p
c_int = ctypes.c_int
c_f = ctypes.c_float
c_d = ctypes.c_double
c_u32 = ctypes.c_uint32
c_u64 = ctypes.c_uint64
P = ctypes.POINTER

# name -> (restype, argtypes); mirrors include/piqmc_b200.h one to one
SIGNATURES = {
    "piqmc_version": (c_int, []),
    "piqmc_last_error": (ctypes.c_char_p, []),
    "piqmc_device_count": (c_int, [P(c_int)]),
    "piqmc_create": (c_int, [c_int, P(c_void)]),
    "piqmc_destroy": (c_int, [c_void]),
    "piqmc_synchronize": (c_int, [c_void]),
    "piqmc_stream": (c_void, [c_void]),
    "piqmc_launch_count": (c_u64, [c_void]),
    "piqmc_host_alloc": (c_int, [c_u64, P(c_void)]),
    "piqmc_host_free": (c_int, [c_void]),
    "piqmc_rand_seed": (None, [P(RandState), ctypes.c_uint]),
    "piqmc_rand_next": (ctypes.c_int32, [P(RandState)]),
    "piqmc_rand_capture_libc": (c_int, [P(RandState)]),
    "piqmc_rand_restore_libc": (c_int, [P(RandState)]),
    "piqmc_set_graph": (c_int, [c_void, c_int, c_int, c_void, c_void, c_int, c_void]),
    "piqmc_set_colouring": (c_int, [c_void, c_int, c_void]),
    "piqmc_order_levels": (c_int, [c_int, c_int, c_void, c_void, c_void, c_void]),
    "piqmc_qa_det": (c_int, [c_void, c_void, c_int, c_int, c_int, c_f, c_int, c_void, c_void, c_void,
                             c_void, c_u64, c_void]),
    "piqmc_sa_det": (c_int, [c_void, c_void, c_int, c_int, c_int, c_void, c_void, c_void, c_void,
                             c_u64, c_void]),
    "piqmc_qa_dense_det": (c_int, [c_void, c_int, c_void, c_void, c_int, c_int, c_int, c_f, c_int, c_void,
                                   c_void, c_void, c_void, c_u64, c_void]),
    "piqmc_sa_dense_det": (c_int, [c_void, c_int, c_void, c_void, c_int, c_int, c_int, c_void, c_void,
                                   c_void, c_void, c_u64, c_void]),
    "piqmc_sa_multispin_det": (c_int, [c_void, c_void, c_int, c_int, c_int, c_void, c_void, c_void]),
    "piqmc_jperp": (c_f, [c_d, c_int, c_f]),
    "piqmc_state_alloc": (c_int, [c_void, c_int, c_int]),
    "piqmc_state_alloc_packed": (c_int, [c_void, c_int, c_int, c_int]),
    "piqmc_state_replicas_to_slices": (c_int, [c_void, c_int, c_int]),
    "piqmc_state_replicas_to_slices_packed": (c_int, [c_void, c_int, c_int, c_int]),
    "piqmc_state_init_random": (c_int, [c_void, c_u64, c_u32, c_int]),
    "piqmc_state_upload_spins": (c_int, [c_void, c_void, c_int]),
    "piqmc_state_upload_words": (c_int, [c_void, c_void]),
    "piqmc_state_download_words": (c_int, [c_void, c_void]),
    "piqmc_state_devptr": (c_void, [c_void]),
    "piqmc_energy_devptr": (c_void, [c_void]),
    "piqmc_qa_colour": (c_int, [c_void, c_void, c_int, c_int, c_f, c_u64, c_u32, c_u32, c_int, c_void]),
    "piqmc_qa_carry": (c_int, [c_void, c_void, c_int, c_int, c_f, c_u64, c_u32, c_u32, c_void]),
    "piqmc_sa_colour": (c_int, [c_void, c_void, c_int, c_int, c_u64, c_u32, c_u32, c_void]),
    "piqmc_set_variant": (c_int, [c_void, c_int]),
    "piqmc_set_chain": (c_int, [c_void, c_int]),
    "piqmc_chain_info": (c_int, [c_void, P(c_int), P(c_int), P(c_d), P(c_int)]),
    "piqmc_chain_plan": (c_int, [c_int, c_int, c_void, c_void, c_int, c_void, P(c_int)]),
    "piqmc_set_global_moves": (c_int, [c_void, c_int]),
    "piqmc_energy": (c_int, [c_void, c_void]),
    "piqmc_results": (c_int, [c_void, c_void, c_void]),
    "piqmc_qa_colour_results": (c_int, [c_void, c_void, c_int, c_int, c_f, c_u64, c_u32, c_u32, c_int, c_void,
                                        c_void, c_void]),
    "piqmc_pipelined_runs": (c_u64, [c_void]),
    "piqmc_last_phase_seconds": (c_int, [c_void, c_void, c_void]),
    "piqmc_energy_histogram": (c_int, [c_void, c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       c_int, c_void, c_void]),
    "piqmc_energy_coo": (c_int, [c_void, c_int, c_int, c_void, c_void, c_void, c_int, c_void, c_void]),
}


def build(verbose=False):
    """Compile libpiqmc_b200.so in-tree with nvcc for sm_100a (csrc/Makefile)."""
    out = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libpiqmc_b200.so failed")
    return SO


def _load():
    if not os.path.exists(SO):
        raise ImportError(
            "libpiqmc_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C pathintegral-qmc_b200/csrc`. There is no CPU fallback." % SO)
    try:
        L = ctypes.CDLL(SO)
    except OSError as e:  # pragma: no cover
        raise ImportError("cannot load %s: %s" % (SO, e))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return L


lib = _load()


def last_error():
    return lib.piqmc_last_error().decode("utf-8", "replace")


def check(rc):
    """Map a status code to the exception the reference would raise."""
    if rc == OK:
        return
    msg = last_error()
    if rc == EZERODIV:
        raise ZeroDivisionError(msg)           # piqmc/qmc.c:2065-2074
    if rc in (EINVAL, ENOGRAPH, ENOSTATE):
        raise ValueError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
