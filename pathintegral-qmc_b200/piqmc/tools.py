"""piqmc.tools -- problem set-up for the annealing kernels (host side).

Mirror of the reference module piqmc/tools.pyx: same function names, argument names and
results, re-implemented in O(nnz) NumPy/Python instead of the reference's O(N*nnz) and O(N^2)
loops, plus what the B200 kernels additionally need (a graph colouring, instance loaders).
"""
import numpy as np
import scipy.sparse as sps

__all__ = ["bits2spins", "spins2bits", "GenerateNeighbors", "Generate2DIsingInstance",
           "ColourGraph", "OrderLevels", "LoadIsingInstance", "IsingFromTriples", "GaussianTorusNeighbors", "TorusNaturalLevels",
           "PackWords", "UnpackWords", "NeighborsToCSR", "CSRToNeighbors"]


def bits2spins(vec):
    """Convert a bitvector @vec to a spinvector (bit 0 -> +1, bit 1 -> -1).
    Reference: piqmc/tools.pyx:20-22."""
    return [-1 if k == 1 else 1 for k in vec]


def spins2bits(vec):
    """Convert a spinvector @vec to a bitvector (+1 -> 0, anything else -> 1).
    Reference: piqmc/tools.pyx:24-26."""
    return [0 if k == 1 else 1 for k in vec]


def GenerateNeighbors(nspins, J, maxnb, savepath=None, colouring=None, format="ell"):
    """Neighbour table of the Ising graph @J: float64[nspins, maxnb, 2] with
    [:, :, 0] = neighbour index and [:, :, 1] = coupling; a diagonal entry J[i,i] (local field)
    appears as a self-neighbour of i; unused rows stay [0, 0].

    Reference: piqmc/tools.pyx:28-96.  Row i lists, in DOK key-iteration order, every stored key
    (a, b) with a == i (-> neighbour b) or b == i (-> neighbour a).  The reference rescans all keys
    for every spin (O(N*nnz)); this walks the keys once and appends to both endpoint rows, which
    yields the same rows in the same order.  A spin with more than @maxnb entries raises
    IndexError, as the reference's bounds-checked buffer write does.

    @colouring (not in the reference): None returns the table alone, like the reference;
    "natural" / "checkerboard" / an int order array returns (table, colour classes) with the
    classes of ColourGraph(table, colouring) -- what the production kernels sweep by.
    @format (not in the reference): "ell" the table; "csr" the (indptr, indices, data) triple of
    NeighborsToCSR(table)."""
    nspins = int(nspins)
    maxnb = int(maxnb)
    if format not in ("ell", "csr"):
        raise ValueError("format must be 'ell' or 'csr'")
    nbs = np.zeros((nspins, maxnb, 2))
    fill = np.zeros(nspins, dtype=np.int64)
    Jd = J.todok()
    for (a, b), val in Jd.items():
        a = int(a)
        b = int(b)
        ends = (a,) if a == b else (a, b)
        for ispin in ends:
            if not 0 <= ispin < nspins:
                continue                      # the reference's scan never matches such keys
            k = fill[ispin]
            if k >= maxnb:
                raise IndexError("Out of bounds on buffer access (axis 1)")
            nbs[ispin, k, 0] = b if ispin == a else a
            nbs[ispin, k, 1] = val
            fill[ispin] = k + 1
    if savepath is not None:
        np.save(savepath, nbs)
    out = nbs if format == "ell" else NeighborsToCSR(nbs)
    if colouring is not None:
        return out, ColourGraph(nbs, colouring)
    return out


def NeighborsToCSR(nbs):
    """CSR form of a neighbour table (the north-star's "ELL/CSR"): (indptr int64[N+1], indices int32[nnz],
    data float64[nnz]); row i holds the entries of table row i with a non-zero coupling, in table order, a
    local field as the self entry (i, J_ii) -- every bond therefore appears in the rows of both its spins, as
    in the table of tools.pyx:74-96."""
    nbs = np.asarray(nbs, dtype=np.float64)
    live = nbs[:, :, 1] != 0.0
    indptr = np.concatenate([[0], np.cumsum(live.sum(axis=1))]).astype(np.int64)
    return indptr, nbs[:, :, 0][live].astype(np.int32), nbs[:, :, 1][live].copy()


def CSRToNeighbors(indptr, indices, data, maxnb=None):
    """The neighbour table float64[N, maxnb, 2] of a CSR graph in the convention of NeighborsToCSR (maxnb:
    the longest row unless given; a longer row raises IndexError like GenerateNeighbors)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n = indptr.size - 1
    deg = np.diff(indptr)
    width = int(deg.max()) if maxnb is None else int(maxnb)
    if n and deg.max() > width:
        raise IndexError("Out of bounds on buffer access (axis 1)")
    nbs = np.zeros((n, max(width, 1), 2))
    col = np.arange(indptr[-1]) - np.repeat(indptr[:-1], deg)
    row = np.repeat(np.arange(n), deg)
    nbs[row, col, 0] = np.asarray(indices)
    nbs[row, col, 1] = np.asarray(data, dtype=np.float64)
    return nbs


def Generate2DIsingInstance(nRows, rng):
    """2D square Ising model on a torus, couplings uniform in [-2, 2], upper-triangular DOK.

    Reference: piqmc/tools.pyx:98-130.  The reference draws inside an O(N^2) double loop; in
    particular the periodic-horizontal coupling of every row-start spin is redrawn once per
    remaining column (tools.pyx:120-123), so the generator state advances by O(N) per such row.
    This reproduces exactly the same draws (bulk `rng.uniform` calls consume the same stream),
    the same final values and the same key insertion order in O(N) Python steps."""
    nRows = int(nRows)
    nSpins = nRows ** 2
    J = sps.dok_matrix((nSpins, nSpins), dtype=np.float64)
    for row in range(nSpins):
        if row < nRows:
            J[row, row + (nRows * (nRows - 1))] = rng.uniform(low=-2, high=2)
        n = nSpins - row                           # columns row .. nSpins-1
        has_right = (n > 1) and (row % nRows != nRows - 1)
        has_bottom = nRows < n
        if row % nRows == 0:
            total = n + int(has_right) + int(has_bottom)
            u = rng.uniform(low=-2, high=2, size=total)
            writes = [(0, (row, row + nRows - 1), u[0])]          # first insertion at col == row
            last = (n - 1) + int(has_right and n - 1 > 1) + int(has_bottom and n - 1 > nRows)
            if has_right:
                writes.append((2, (row, row + 1), u[2]))
            if has_bottom:
                tb = nRows + 1 + int(has_right)
                writes.append((tb, (row, row + nRows), u[tb]))
            writes.append((last, (row, row + nRows - 1), u[last]))
            for _, key, val in sorted(writes, key=lambda w: w[0]):
                J[key] = val
        else:
            if has_right:
                J[row, row + 1] = rng.uniform(low=-2, high=2)
            if has_bottom:
                J[row, row + nRows] = rng.uniform(low=-2, high=2)
    return J


def Generate2DLattice(nrows, ncols, rng, periodic=0):
    """nrows x ncols square-lattice Ising model, couplings uniform in [-1e-8, 1e-8], upper-
    triangular DOK; @periodic=1 closes both directions into a torus.

    Reference: piqmc/tools.pyx:132-176.  One draw per stored bond in the reference's order (for
    every spin: wrap-around vertical, wrap-around horizontal, right neighbour, bottom neighbour),
    so the same @rng gives the same matrix; O(N) instead of the reference's O(N^2) column scan."""
    nrows, ncols = int(nrows), int(ncols)
    nspins = nrows * ncols
    J = sps.dok_matrix((nspins, nspins), dtype=np.float64)
    for jrow in range(nspins):
        if jrow < ncols and periodic:
            J[jrow, jrow + ncols * (nrows - 1)] = rng.uniform(low=-1e-8, high=1e-8)
        if jrow % ncols == 0 and periodic:
            J[jrow, jrow + ncols - 1] = rng.uniform(low=-1e-8, high=1e-8)
        # the reference scans columns upwards: right neighbour (jrow+1) before bottom (jrow+ncols)
        if jrow + 1 < nspins and jrow % ncols != ncols - 1:
            J[jrow, jrow + 1] = rng.uniform(low=-1e-8, high=1e-8)
        if jrow + ncols < nspins:
            J[jrow, jrow + ncols] = rng.uniform(low=-1e-8, high=1e-8)
    return J


def GenerateKblockLattice(nrows, ncols, rng, k=1):
    """Lattice in which every spin is coupled to the spins on the square rings of radius 1..@k
    around it (open boundaries), couplings uniform in [-1e-8, 1e-8], upper-triangular DOK.

    Reference: piqmc/tools.pyx:178-272.  Follows the reference ring by ring and side by side (top,
    bottom, left, right), including how it clips a ring at the lattice border and that a ring
    corner shared by two sides is drawn twice (the later draw wins), so the same @rng gives the
    same matrix."""
    nrows, ncols, k = int(nrows), int(ncols), int(k)
    nspins = nrows * ncols
    J = sps.dok_matrix((nspins, nspins), dtype=np.float64)
    for ispin in range(nspins):
        col = ispin % ncols
        room_right = ncols - 1 - col                       # spins to the right in this lattice row
        for ki in range(1, k + 1):
            kleft = min(ki, col)
            kright = min(ki, room_right)
            kup = ki
            while ispin - kup * ncols < 0:
                kup -= 1
            kdown = ki
            while ispin + kdown * ncols >= nspins:
                kdown -= 1
            corner = max(ispin - kup * ncols - kleft, 0)   # top-left spin of the (clipped) ring
            width, height = kleft + kright, kup + kdown
            ring = [corner + r for r in range(width + 1)]                              # top side
            ring += [corner + r + height * ncols for r in range(width + 1)]            # bottom side
            ring += [corner + r * ncols for r in range(1, height + 1)]                 # left side
            ring += [corner + width + r * ncols for r in range(1, height + 1)]         # right side
            for idx in ring:
                if ispin < idx:
                    J[ispin, idx] = rng.uniform(low=-1e-8, high=1e-8)
    return J


# ---------------------------------------------------------------------------------------------
# Additions for the B200 kernels (no counterpart in the reference)
# ---------------------------------------------------------------------------------------------
def _adjacency(nbs):
    nbs = np.asarray(nbs)
    n = nbs.shape[0]
    idx = nbs[:, :, 0].astype(np.int64)
    live = (nbs[:, :, 1] != 0.0) & (idx != np.arange(n)[:, None])
    return [idx[i][live[i]].tolist() for i in range(n)]


def OrderLevels(nbs, order=None):
    """Level colouring of a sequential visiting order: level[i] = 1 + max(level[j]) over coupled
    neighbours j visited before i (0 if there is none).  Updating level 0, then level 1, ... with
    all spins of a level at once is EXACTLY the sequential sweep in that order: every spin sees
    the new value of the neighbours visited before it and the old value of the others.  This is
    how the colour-class kernels reproduce the reference's natural-order sweep
    (piqmc/qmc.pyx:320, QuantumAnneal_parallel) and permutation-order sweeps (piqmc/sa.pyx:100).
    order: None = natural order 0..N-1, or an int array (order[t] = spin visited at step t)."""
    adj = _adjacency(nbs)
    n = len(adj)
    level = -np.ones(n, dtype=np.int32)
    for i in (range(n) if order is None else np.asarray(order).tolist()):
        m = -1
        for j in adj[i]:
            if level[j] > m:
                m = level[j]
        level[i] = m + 1
    return level


def ColourGraph(nbs, order=None):
    """Colour classes for the colour-parallel kernels: int32[N], classes are swept in ascending
    colour index and no two spins joined by a non-zero coupling share a class.  Self entries
    (local fields) and zero couplings (pad rows) are ignored.

    order=None or "checkerboard": fewest classes -- 2 by breadth-first search when the graph is
        bipartite (the even-L tori, bipartite8, boixo), else greedy largest-degree-first (K_n -> n).
    order="natural" or an int array: the dependency levels of that sequential visiting order
        (OrderLevels): the sweep then equals the sequential sweep in that order."""
    if order is not None and not (isinstance(order, str) and order == "checkerboard"):
        return OrderLevels(nbs, None if isinstance(order, str) and order == "natural" else order)
    adj = _adjacency(nbs)
    n = len(adj)
    color = -np.ones(n, dtype=np.int32)
    bipartite = True
    for s in range(n):
        if color[s] >= 0:
            continue
        color[s] = 0
        frontier = [s]
        while frontier and bipartite:
            nxt = []
            for i in frontier:
                ci = color[i]
                for j in adj[i]:
                    if color[j] < 0:
                        color[j] = 1 - ci
                        nxt.append(j)
                    elif color[j] == ci:
                        bipartite = False
                        break
                if not bipartite:
                    break
            frontier = nxt
        if not bipartite:
            break
    if bipartite:
        return color
    color[:] = -1
    for i in sorted(range(n), key=lambda i: -len(adj[i])):
        used = {color[j] for j in adj[i] if color[j] >= 0}
        c = 0
        while c in used:
            c += 1
        color[i] = c
    return color


def IsingFromTriples(ijv, nspins=None):
    """DOK matrix from rows (i, j, J) with 1-indexed spins, i == j rows being local fields --
    the reference's instance text format (examples/boixo.py:40-44)."""
    ijv = np.atleast_2d(np.asarray(ijv, dtype=np.float64))
    if nspins is None:
        nspins = int(ijv[:, :2].max())
    J = sps.dok_matrix((nspins, nspins))
    for i, j, val in ijv:
        J[int(i) - 1, int(j) - 1] = val
    return J


def LoadIsingInstance(path, nspins=None):
    """Read a 3-column `i j J` text instance (examples/ising_instances/*.txt)."""
    return IsingFromTriples(np.loadtxt(path), nspins)


def GaussianTorusNeighbors(L, seed=2024, dtype=np.float32):
    """Synthetic benchmark instance (BASELINE.json configs[4]): L x L periodic torus, bonds
    (i, right(i)) and (i, down(i)) with J ~ N(0,1) drawn from RandomState(seed) in row-major bond
    order and rounded to @dtype; no fields.  Returns (nbs float64[N,4,2], color int32[N]) built
    directly (rows in the order GenerateNeighbors would produce for keys inserted in bond order)."""
    L = int(L)
    n = L * L
    rng = np.random.RandomState(seed)
    Jb = rng.standard_normal(size=(n, 2)).astype(dtype).astype(np.float64)   # [:,0]=right, [:,1]=down
    i = np.arange(n)
    y, x = i // L, i % L
    right = y * L + (x + 1) % L
    down = ((y + 1) % L) * L + x
    left = y * L + (x - 1) % L
    up = ((y - 1) % L) * L + x
    nb = np.stack([right, down, left, up], axis=1)
    jv = np.stack([Jb[i, 0], Jb[i, 1], Jb[left, 0], Jb[up, 1]], axis=1)
    when = np.stack([2 * i, 2 * i + 1, 2 * left, 2 * up + 1], axis=1)          # key insertion time
    order = np.argsort(when, axis=1, kind="stable")
    nbs = np.zeros((n, 4, 2))
    nbs[:, :, 0] = np.take_along_axis(nb, order, axis=1)
    nbs[:, :, 1] = np.take_along_axis(jv, order, axis=1)
    if L % 2 == 0:
        color = ((x + y) % 2).astype(np.int32)
    else:
        color = ColourGraph(nbs)
    return nbs, color


def TorusNaturalLevels(L):
    """OrderLevels(nbs, None) of an L x L torus numbered row-major, in closed form: the
    anti-diagonals level(y, x) = x + y (2L-1 classes)."""
    i = np.arange(L * L)
    return ((i // L) + (i % L)).astype(np.int32)


def PackWords(spins):
    """int8/float +-1 array [..., lanes, N] -> uint64 [..., N]: bit `lane` set <-> spin -1."""
    s = np.asarray(spins)
    lanes = s.shape[-2]
    if lanes > 64:
        raise ValueError("at most 64 lanes per word")
    bits = (s < 0).astype(np.uint64)
    sh = np.arange(lanes, dtype=np.uint64).reshape((lanes, 1))
    return np.bitwise_or.reduce(bits << sh, axis=-2)


def UnpackWords(words, lanes):
    """uint64 [..., N] -> int8 +-1 [..., lanes, N]."""
    w = np.asarray(words, dtype=np.uint64)
    sh = np.arange(lanes, dtype=np.uint64).reshape((lanes, 1))
    bits = (w[..., None, :] >> sh) & np.uint64(1)
    return (1 - 2 * bits.astype(np.int8)).astype(np.int8)
