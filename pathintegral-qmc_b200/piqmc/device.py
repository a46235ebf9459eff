"""Device handle: the host-side object behind piqmc.qmc / piqmc.sa / piqmc.tools.

One `Device` owns one GPU context handle of libpiqmc_b200 (stream, graph tables, packed replica
state).  The reference has no such object -- it is a set of free functions mutating NumPy
arrays -- so the drop-in functions in qmc.py / sa.py use a process-wide default Device and this
class is the explicit, batched interface (R replicas per call).
"""
import ctypes
import hashlib

import numpy as np

from . import _lib
from ._lib import RandState, check, lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def split_nbs(nbs):
    """nbs float64[N,maxnb,2] (tools.GenerateNeighbors) -> (idx int32[N,maxnb], J float64[N,maxnb]).
    idx uses the reference's int(...) truncation (piqmc/qmc.pyx:104)."""
    nbs = np.asarray(nbs)
    if nbs.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'float_t' but got '%s'" % nbs.dtype)
    if nbs.ndim != 3 or nbs.shape[2] != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % nbs.ndim)
    idx = np.ascontiguousarray(nbs[:, :, 0].astype(np.int32))
    J = np.ascontiguousarray(nbs[:, :, 1], dtype=np.float64)
    return idx, J


def rand_states(seeds):
    """ctypes array of glibc rand() states, one per seed (== srand(seed) in a fresh process)."""
    arr = (RandState * len(seeds))()
    for k, s in enumerate(seeds):
        lib.piqmc_rand_seed(ctypes.byref(arr[k]), int(s) & 0xFFFFFFFF)
    return arr


class DeviceArray:
    """A device buffer exposed through __cuda_array_interface__ (zero-copy into torch for the
    final NCCL gather).  The memory belongs to the Device; keep the Device alive."""

    def __init__(self, ptr, shape, typestr, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2, "strides": None}


class Device:
    def __init__(self, index=0):
        self._h = ctypes.c_void_p()
        check(lib.piqmc_create(int(index), ctypes.byref(self._h)))
        self.index = int(index)
        self.nspins = 0
        self.maxnb = 0
        self.ncolors = 0
        self.nrows = 0
        self.lanes = 0
        self.slices = 0          # slices per replica (== lanes unless several replicas share a word)
        self.per_word = 1        # replicas per word
        self.variant = 0
        self._graph_key = None

    def close(self):
        if self._h:
            lib.piqmc_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ misc
    @property
    def stream(self):
        """cudaStream_t (int) all work of this Device is issued on."""
        return lib.piqmc_stream(self._h)

    @property
    def launch_count(self):
        return int(lib.piqmc_launch_count(self._h))

    def synchronize(self):
        check(lib.piqmc_synchronize(self._h))

    def set_variant(self, variant):
        check(lib.piqmc_set_variant(self._h, int(variant)))
        self.variant = int(variant)

    def set_chain(self, chain_len=0):
        """Chain pipeline: 0 = let the library choose the chain length, > 0 = force it (and the kernel)."""
        check(lib.piqmc_set_chain(self._h, int(chain_len)))

    def chain_info(self):
        """(chain length, chains per ring, modelled steps per sweep, selected) of the current plan;
        (0, 0, 0.0, False) if there is none.  selected: a static-colouring QA/SA run takes the pipeline."""
        c, n, per, sel = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0.0), ctypes.c_int(0)
        check(lib.piqmc_chain_info(self._h, ctypes.byref(c), ctypes.byref(n), ctypes.byref(per), ctypes.byref(sel)))
        return c.value, n.value, per.value, bool(sel.value)

    def set_global_moves(self, enable):
        """World-line (all-slices) moves after the local moves of every spin (QA, maxnb <= 4)."""
        check(lib.piqmc_set_global_moves(self._h, 1 if enable else 0))

    # ------------------------------------------------------------------ graph
    def set_graph(self, nbs, color=None):
        """Upload the neighbour table (and, for the colour paths, a proper colouring).
        Cached on content: re-uploading the same table is free."""
        # the same table and colouring as last time (the usual case for repeated anneals of one instance): an exact
        # comparison with the private copies kept below, a fraction of a millisecond, instead of casts + a digest
        last = getattr(self, "_graph_last", None)
        if last is not None and self._graph_key is not None:
            nbs_l, col_l = last
            a = np.asarray(nbs)
            if (a.shape == nbs_l.shape and a.dtype == nbs_l.dtype and (color is None) == (col_l is None)
                    and np.array_equal(a, nbs_l)
                    and (color is None or (np.shape(color) == col_l.shape and np.array_equal(color, col_l)))):
                return
        idx, J = split_nbs(nbs)
        col = None if color is None else np.ascontiguousarray(color, dtype=np.int32)
        hsh = hashlib.blake2b(digest_size=16)
        hsh.update(idx.tobytes())
        hsh.update(J.tobytes())
        if col is not None:
            hsh.update(col.tobytes())
        key = (idx.shape, hsh.digest())
        keep = (np.array(nbs, copy=True), None if color is None else np.array(color, copy=True))
        if key == self._graph_key:
            self._graph_last = keep
            return
        self._graph_last = None
        ncol = 0 if col is None else int(col.max()) + 1
        if col is not None and col.shape != (idx.shape[0],):
            raise ValueError("color must have one entry per spin")
        check(lib.piqmc_set_graph(self._h, idx.shape[0], idx.shape[1], _ptr(idx), _ptr(J), ncol, _ptr(col)))
        if idx.shape[0] != self.nspins:             # the library keeps a state on the same spins
            self.nrows = self.lanes = 0
        self.nspins, self.maxnb, self.ncolors = idx.shape[0], idx.shape[1], ncol
        self._graph_key = key
        self._graph_last = keep

    # ------------------------------------------------------------------ deterministic paths
    def qa_det(self, sched, mcsteps, slices, temp, spins, perms, rstates=None, uniforms=None):
        """qmc.QuantumAnneal for R replicas, bit-exact.  spins int8[R,N,P] (in place),
        perms int32[R,nsweeps,N]; rstates = ctypes array of RandState (advanced in place) or
        uniforms float64[R,nuni].  Returns consumed uint64[R]."""
        sched = np.ascontiguousarray(sched, dtype=np.float64)
        R = spins.shape[0]
        self._check_det(spins, (R, self.nspins, slices), perms, (R, sched.size * mcsteps, self.nspins))
        consumed = np.zeros(R, dtype=np.uint64)
        uni, nuni = self._uniforms(uniforms, R)
        check(lib.piqmc_qa_det(self._h, _ptr(sched), sched.size, int(mcsteps), int(slices),
                               ctypes.c_float(temp), R, _ptr(spins), _ptr(perms),
                               None if rstates is None else ctypes.cast(rstates, ctypes.c_void_p),
                               _ptr(uni), nuni, _ptr(consumed)))
        return consumed

    def sa_det(self, sched, mcsteps, spins, perms, rstates=None, uniforms=None):
        """sa.Anneal for R replicas, bit-exact.  spins int8[R,N] in place."""
        sched = np.ascontiguousarray(sched, dtype=np.float64)
        R = spins.shape[0]
        self._check_det(spins, (R, self.nspins), perms, (R, sched.size * mcsteps, self.nspins))
        consumed = np.zeros(R, dtype=np.uint64)
        uni, nuni = self._uniforms(uniforms, R)
        check(lib.piqmc_sa_det(self._h, _ptr(sched), sched.size, int(mcsteps), R, _ptr(spins), _ptr(perms),
                               None if rstates is None else ctypes.cast(rstates, ctypes.c_void_p),
                               _ptr(uni), nuni, _ptr(consumed)))
        return consumed

    @staticmethod
    def _dense(J):
        J = np.asarray(J)
        if J.dtype != np.float64 or J.ndim != 2 or J.shape[0] != J.shape[1]:
            raise ValueError("J must be a square float64 matrix")
        return np.ascontiguousarray(J)

    def qa_dense_det(self, sched, mcsteps, slices, temp, spins, J, perms, rstates=None, uniforms=None):
        """qmc.QuantumAnneal_dense for R replicas, bit-exact.  J float64[N,N] (upper triangle +
        diagonal read); spins int8[R,N,P] in place; other arguments as qa_det.  Needs no graph."""
        sched = np.ascontiguousarray(sched, dtype=np.float64)
        J = self._dense(J)
        n, R = J.shape[0], spins.shape[0]
        self._check_det(spins, (R, n, slices), perms, (R, sched.size * mcsteps, n), nspins=n)
        consumed = np.zeros(R, dtype=np.uint64)
        uni, nuni = self._uniforms(uniforms, R)
        check(lib.piqmc_qa_dense_det(self._h, n, _ptr(J), _ptr(sched), sched.size, int(mcsteps), int(slices),
                                     ctypes.c_float(temp), R, _ptr(spins), _ptr(perms),
                                     None if rstates is None else ctypes.cast(rstates, ctypes.c_void_p),
                                     _ptr(uni), nuni, _ptr(consumed)))
        return consumed

    def sa_dense_det(self, sched, mcsteps, spins, J, perms, rstates=None, uniforms=None):
        """sa.Anneal_dense for R replicas, bit-exact.  spins int8[R,N] in place."""
        sched = np.ascontiguousarray(sched, dtype=np.float64)
        J = self._dense(J)
        n, R = J.shape[0], spins.shape[0]
        self._check_det(spins, (R, n), perms, (R, sched.size * mcsteps, n), nspins=n)
        consumed = np.zeros(R, dtype=np.uint64)
        uni, nuni = self._uniforms(uniforms, R)
        check(lib.piqmc_sa_dense_det(self._h, n, _ptr(J), _ptr(sched), sched.size, int(mcsteps), R, _ptr(spins),
                                     _ptr(perms),
                                     None if rstates is None else ctypes.cast(rstates, ctypes.c_void_p),
                                     _ptr(uni), nuni, _ptr(consumed)))
        return consumed

    def sa_multispin_det(self, sched, mcsteps, words, perms, rands):
        """sa.Anneal_multispin for G groups of 64 replicas.  words uint64[G,N] in place,
        perms int32[G,nsweeps,N], rands float64[G,nsweeps*N,64]."""
        sched = np.ascontiguousarray(sched, dtype=np.float64)
        G = words.shape[0]
        nsw = sched.size * mcsteps
        if words.dtype != np.uint64 or not words.flags.c_contiguous or words.shape != (G, self.nspins):
            raise ValueError("words must be C-contiguous uint64[G,N]")
        if perms.dtype != np.int32 or not perms.flags.c_contiguous or perms.shape != (G, nsw, self.nspins):
            raise ValueError("perms must be C-contiguous int32[G,nsweeps,N]")
        rands = np.ascontiguousarray(rands, dtype=np.float64)
        if rands.size != G * nsw * self.nspins * 64:
            raise ValueError("rands must hold 64 doubles per attempt")
        check(lib.piqmc_sa_multispin_det(self._h, _ptr(sched), sched.size, int(mcsteps), G, _ptr(words),
                                         _ptr(perms), _ptr(rands)))

    def _check_det(self, spins, sshape, perms, pshape, nspins=None):
        if nspins is None:
            nspins = self.nspins
        if nspins == 0:
            raise ValueError("set_graph has not been called")
        if spins.dtype != np.int8 or not spins.flags.c_contiguous or spins.shape != sshape:
            raise ValueError("spins must be C-contiguous int8 of shape %s" % (sshape,))
        if perms.dtype != np.int32 or not perms.flags.c_contiguous or perms.shape != pshape:
            raise ValueError("perms must be C-contiguous int32 of shape %s" % (pshape,))
        if perms.size and (perms.min() < 0 or perms.max() >= nspins):
            raise ValueError("perms entries out of range")

    @staticmethod
    def _uniforms(uniforms, R):
        if uniforms is None:
            return None, 0
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        if u.ndim != 2 or u.shape[0] != R:
            raise ValueError("uniforms must be float64[R,nuni]")
        return u, u.shape[1]

    # ------------------------------------------------------------------ packed state
    def state_alloc(self, nrows, lanes, per_word=1):
        """nrows words per spin of `lanes` lanes.  per_word > 1 (QA, lanes <= 32 slices, a multiple
        of 4): every word holds per_word replicas of `lanes` slices each, replica row*per_word + g in
        bits [g*lanes, (g+1)*lanes); the host-facing methods below then count in replicas."""
        check(lib.piqmc_state_alloc_packed(self._h, int(nrows), int(lanes), int(per_word)))
        self.nrows, self.slices, self.per_word = int(nrows), int(lanes), int(per_word)
        self.lanes = self.slices * self.per_word

    def state_replicas_to_slices(self, nreplicas, slices, per_word=1):
        """SA state (64 replicas per word) -> QA state (all slices of a replica = its spin), on the
        device."""
        check(lib.piqmc_state_replicas_to_slices_packed(self._h, int(nreplicas), int(slices), int(per_word)))
        self.per_word, self.slices = int(per_word), int(slices)
        self.nrows = (int(nreplicas) + self.per_word - 1) // self.per_word
        self.lanes = self.slices * self.per_word

    @property
    def nreplicas_held(self):
        """replicas a QA state holds (rows x replicas per word)"""
        return self.nrows * self.per_word

    def _unpack_replicas(self, raw):
        """packed words [nspins, nrows] -> [nrows*per_word, nspins] with bit k = slice k"""
        if self.per_word == 1:
            return raw.T
        ones = np.uint64((1 << self.slices) - 1)
        out = np.empty((self.nrows, self.per_word, self.nspins), dtype=np.uint64)
        for g in range(self.per_word):
            out[:, g, :] = ((raw >> np.uint64(g * self.slices)) & ones).T
        return out.reshape(self.nrows * self.per_word, self.nspins)

    def state_init_random(self, seed, row0=0, tile=True):
        check(lib.piqmc_state_init_random(self._h, int(seed), int(row0), 1 if tile else 0))

    def state_upload_spins(self, spins, tile=True):
        """tile: spins int8[R,N] copied to every slice; else int8[R,slices,N]; R = nrows*per_word."""
        if not (isinstance(spins, np.ndarray) and spins.dtype == np.int8 and spins.flags.c_contiguous):
            spins = np.ascontiguousarray(spins, dtype=np.int8)
        R = self.nrows * self.per_word
        want = (R, self.nspins) if tile else (R, self.slices, self.nspins)
        if spins.shape != want:
            raise ValueError("spins must have shape %s" % (want,))
        check(lib.piqmc_state_upload_spins(self._h, _ptr(spins), 1 if tile else 0))

    def state_upload_words(self, words):
        """words uint64[nrows, nspins] (logical shape; stored spin-major on the device)."""
        words = np.asarray(words, dtype=np.uint64)
        if words.shape != (self.nrows, self.nspins):
            raise ValueError("words must have shape (nrows, nspins)")
        dev = np.ascontiguousarray(words.T)
        check(lib.piqmc_state_upload_words(self._h, _ptr(dev)))

    def state_download_words(self, out=None):
        """uint64 words of logical shape (nrows, nspins).  The device layout is word[spin][row], so
        the result is the transposed VIEW of a C-contiguous (nspins, nrows) buffer (no host copy);
        pass `out` (C-contiguous uint64[nspins, nrows], e.g. pinned) to reuse a buffer."""
        if out is None:
            out = np.empty((self.nspins, self.nrows), dtype=np.uint64)
        elif out.dtype != np.uint64 or not out.flags.c_contiguous or out.shape != (self.nspins, self.nrows):
            raise ValueError("out must be C-contiguous uint64[nspins, nrows]")
        check(lib.piqmc_state_download_words(self._h, _ptr(out)))
        return self._unpack_replicas(out)

    def state_download_spins(self):
        """int8[nrows, lanes, N] of +-1 (host-side unpack of the packed words)."""
        w = self.state_download_words()
        lanes = np.arange(self.slices or self.lanes, dtype=np.uint64)
        bits = (w[:, None, :] >> lanes[None, :, None]) & np.uint64(1)
        return (1 - 2 * bits.astype(np.int8)).astype(np.int8)

    def state_device_array(self):
        """Device view of the packed state in its native (nspins, nrows) layout."""
        return DeviceArray(lib.piqmc_state_devptr(self._h), (self.nspins, self.nrows), "<u8", self)

    def energy_device_array(self):
        return DeviceArray(lib.piqmc_energy_devptr(self._h),
                           (self.nrows * self.per_word, self.lanes // self.per_word), "<f8", self)

    # ------------------------------------------------------------------ production sweeps
    def _orders(self, orders, nsweeps):
        if orders is None:
            return None
        o = np.ascontiguousarray(orders, dtype=np.int32)
        if o.shape != (nsweeps, self.nspins):
            raise ValueError("orders must have shape (nsweeps, nspins) = (%d, %d)" % (nsweeps, self.nspins))
        return o

    def qa_colour(self, sched, mcsteps, temp, seed, replica0=0, sweep0=0, trotter=0, orders=None):
        """QA sweeps over the resident state (returns when they have run: the library synchronises its stream,
        the per-sweep parameter arrays live only for the call).  orders=None: the graph's colouring.
        orders int32[nsweeps, N]: sweep s is the
        sequential sweep visiting spins in orders[s] (run through its level colouring)."""
        sched = np.ascontiguousarray(sched, dtype=np.float64)
        o = self._orders(orders, sched.size * int(mcsteps))
        check(lib.piqmc_qa_colour(self._h, _ptr(sched), sched.size, int(mcsteps), ctypes.c_float(temp),
                                  int(seed), int(replica0), int(sweep0), int(trotter), _ptr(o)))

    def qa_carry(self, sched, mcsteps, temp, seed, replica0=0, sweep0=0, orders=None):
        """QA sweeps with the AS-SHIPPED semantics of qmc.QuantumAnneal (energy difference carried over a whole
        slice sweep, piqmc/qmc.pyx:98-136) over the resident state; orders int32[nsweeps, N] = the visiting order
        of every sweep, shared by all replicas (None: 0..N-1)."""
        sched = np.ascontiguousarray(sched, dtype=np.float64)
        o = self._orders(orders, sched.size * int(mcsteps))
        check(lib.piqmc_qa_carry(self._h, _ptr(sched), sched.size, int(mcsteps), ctypes.c_float(temp), int(seed),
                                 int(replica0), int(sweep0), _ptr(o)))

    def sa_colour(self, sched, mcsteps, seed, row0=0, sweep0=0, orders=None):
        sched = np.ascontiguousarray(sched, dtype=np.float64)
        o = self._orders(orders, sched.size * int(mcsteps))
        check(lib.piqmc_sa_colour(self._h, _ptr(sched), sched.size, int(mcsteps), int(seed), int(row0),
                                  int(sweep0), _ptr(o)))

    def set_colouring(self, color):
        col = np.ascontiguousarray(color, dtype=np.int32)
        if col.shape != (self.nspins,):
            raise ValueError("color must have one entry per spin")
        check(lib.piqmc_set_colouring(self._h, int(col.max()) + 1, _ptr(col)))
        self.ncolors = int(col.max()) + 1
        self._graph_key = None

    # ------------------------------------------------------------------ energies
    def energy(self, download=True):
        """ClassicalIsingEnergy of every (row, lane): float64[nrows, lanes]."""
        if not download:
            check(lib.piqmc_energy(self._h, None))
            return None
        out = np.empty((self.nrows, self.lanes), dtype=np.float64)
        check(lib.piqmc_energy(self._h, _ptr(out)))
        return out.reshape(self.nrows * self.per_word, self.lanes // self.per_word)

    def results(self, words_out=None):
        """(energies float64[nrows, lanes], words uint64 view [nrows, nspins]) in one call; the
        state download overlaps the energy reduction."""
        en = np.empty((self.nrows, self.lanes), dtype=np.float64)
        if words_out is None:
            words_out = np.empty((self.nspins, self.nrows), dtype=np.uint64)
        elif (words_out.dtype != np.uint64 or not words_out.flags.c_contiguous
              or words_out.shape != (self.nspins, self.nrows)):
            raise ValueError("words_out must be C-contiguous uint64[nspins, nrows]")
        check(lib.piqmc_results(self._h, _ptr(en), _ptr(words_out)))
        return (en.reshape(self.nrows * self.per_word, self.lanes // self.per_word),
                self._unpack_replicas(words_out))

    def qa_colour_results(self, sched, mcsteps, temp, seed, replica0=0, sweep0=0, trotter=0, orders=None,
                          words_out=None):
        """qa_colour + results as one library call: with the dataflow kernel the row chunks of the state are
        staggered inside the launch and downloaded / reduced to energies as they finish, under the sweeps of
        the others (piqmc_qa_colour_results).  Same return value as results()."""
        sched = np.ascontiguousarray(sched, dtype=np.float64)
        o = self._orders(orders, sched.size * int(mcsteps))
        en = np.empty((self.nrows, self.lanes), dtype=np.float64)
        if words_out is None:
            words_out = np.empty((self.nspins, self.nrows), dtype=np.uint64)
        elif (words_out.dtype != np.uint64 or not words_out.flags.c_contiguous
              or words_out.shape != (self.nspins, self.nrows)):
            raise ValueError("words_out must be C-contiguous uint64[nspins, nrows]")
        check(lib.piqmc_qa_colour_results(self._h, _ptr(sched), sched.size, int(mcsteps), ctypes.c_float(temp),
                                          int(seed), int(replica0), int(sweep0), int(trotter), _ptr(o), _ptr(en),
                                          _ptr(words_out)))
        return (en.reshape(self.nrows * self.per_word, self.lanes // self.per_word),
                self._unpack_replicas(words_out))

    def last_phase_seconds(self):
        """(sweeps, energies + download) host seconds of the last qa_colour_results call"""
        a, b = ctypes.c_double(0.0), ctypes.c_double(0.0)
        check(lib.piqmc_last_phase_seconds(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    @property
    def pipelined_runs(self):
        """calls of qa_colour_results that took the overlapped (staggered chunks) path"""
        return int(lib.piqmc_pipelined_runs(self._h))

    def energy_histogram(self, e0=0.0, scale=1.0, lo=0.0, hi=1.0, nbins=64, reduce="mean"):
        """Histogram of (E - e0) * scale over the replicas of the last energy() / results() call, on the device
        (e.g. e0 = ground-state energy, scale = 1 / nspins: residual energy per spin, examples/santoro80.py:290-323
        of the reference).  reduce: "mean" over a replica's slices, "min" (its best slice) or "all" (every slice).
        Returns dict(counts uint64[nbins], below, above, mean, min, max, edges float64[nbins + 1])."""
        mode = {"mean": 0, "min": 1, "all": 2}[reduce]
        counts = np.zeros(int(nbins) + 2, dtype=np.uint64)
        stats = np.zeros(3, dtype=np.float64)
        check(lib.piqmc_energy_histogram(self._h, mode, float(e0), float(scale), float(lo), float(hi), int(nbins),
                                         _ptr(counts), _ptr(stats)))
        total = int(counts.sum())
        return {"counts": counts[:nbins], "below": int(counts[nbins]), "above": int(counts[nbins + 1]),
                "mean": stats[0] / max(total, 1), "min": stats[1], "max": stats[2],
                "edges": np.linspace(lo, hi, int(nbins) + 1)}

    def energy_coo(self, nspins, row, col, val, spins):
        row = np.ascontiguousarray(row, dtype=np.int32)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        spins = np.ascontiguousarray(spins, dtype=np.int8)
        if spins.ndim != 2 or spins.shape[1] != nspins:
            raise ValueError("spins must be int8[nconfs, nspins]")
        out = np.empty(spins.shape[0], dtype=np.float64)
        check(lib.piqmc_energy_coo(self._h, int(nspins), val.size, _ptr(row), _ptr(col), _ptr(val),
                                   spins.shape[0], _ptr(spins), _ptr(out)))
        return out


_default = {}


def default_device(index=0):
    """Process-wide Device used by the drop-in functions."""
    d = _default.get(index)
    if d is None:
        d = _default[index] = Device(index)
    return d


class _Pinned:
    def __init__(self, nbytes):
        self.ptr = ctypes.c_void_p()
        check(lib.piqmc_host_alloc(int(nbytes), ctypes.byref(self.ptr)))

    def __del__(self):
        try:
            lib.piqmc_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """NumPy array in page-locked host memory (cudaHostAlloc): host<->device copies of such
    arrays run at full PCIe speed.  The memory is released when the array is collected."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    owner = _Pinned(n)
    buf = (ctypes.c_char * max(n, 1)).from_address(owner.ptr.value)
    buf._piqmc_owner = owner                                  # the array's base keeps the allocation alive
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def order_levels(nbs, order=None):
    """Level colouring of a sequential visiting order (piqmc_order_levels): int32[N]."""
    idx, J = split_nbs(nbs)
    o = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
    level = np.empty(idx.shape[0], dtype=np.int32)
    rc = lib.piqmc_order_levels(idx.shape[0], idx.shape[1], _ptr(idx), _ptr(J), _ptr(o), _ptr(level))
    if rc < 0:
        check(rc)
    return level


def chain_plan(nbs, chain_len=0):
    """Host-side plan of the chain pipeline for the natural-order sweep of `nbs` (piqmc_chain_plan, no
    device needed): (chain length, kinds uint8[N,4], wrap).  kinds[i, k]: where sorted table column k of
    spin i gets its neighbour word from (0 none, 1 previous step of the chain, 2 next own-row word,
    3 preceding chain, 4 following chain); wrap: first and last chain are coupled."""
    idx, J = split_nbs(nbs)
    n = idx.shape[0]
    kinds = np.zeros((n, 4), dtype=np.uint8)
    wrap = ctypes.c_int(0)
    rc = lib.piqmc_chain_plan(n, idx.shape[1], _ptr(idx), _ptr(J), int(chain_len), _ptr(kinds), ctypes.byref(wrap))
    if rc < 0:
        check(rc)
    return rc, kinds, bool(wrap.value)


def device_count():
    n = ctypes.c_int(0)
    check(lib.piqmc_device_count(ctypes.byref(n)))
    return n.value


def capture_libc_rand():
    s = RandState()
    check(lib.piqmc_rand_capture_libc(ctypes.byref(s)))
    return s


def restore_libc_rand(state):
    check(lib.piqmc_rand_restore_libc(ctypes.byref(state)))
