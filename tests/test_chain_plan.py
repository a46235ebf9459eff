"""CPU tests of the chain pipeline's plan (piqmc_chain_plan, csrc/api.cu) and of its hand-over protocol.

The CUDA kernel (csrc/chain_kernels.cu) cuts the natural-order sweep into chains and lets them run
concurrently; what keeps it equal to the sequential sweep is the per-slot dependency kinds the host
computes.  Here the protocol the kernel implements -- neighbour words requested one step ahead,
hand-over rings of 4 self-validating packets that the producer overwrites without waiting,
progress counters published every 8 steps, fall-back to the state word when a packet was lost -- is
executed by a Python model under adversarial schedules on random graphs and lattices, and must (a)
never deadlock and (b) end in exactly the state of the sequential sweep.
"""
import numpy as np
import pytest
import scipy.sparse as sps

import piqmc.tools as tools
from piqmc import device

K_ZERO, K_PREV, K_NEXT, K_LL_CUR, K_LL_OLD, K_MEM_SELF, K_MEM_CUR, K_MEM_OLD = range(8)
D, G = 4, 8


def _sorted_cols(nbs):
    """table column of every sorted slot, as build_chain_stat sorts them (|float32 J| descending, stable)"""
    n, maxnb = nbs.shape[:2]
    J = np.zeros((n, 4), dtype=np.float32)
    J[:, :maxnb] = nbs[:, :, 1].astype(np.float32)
    return np.argsort(-np.abs(J), axis=1, kind="stable")


def _sequential(nbs, state, nsweeps):
    n, maxnb = nbs.shape[:2]
    cols = _sorted_cols(nbs)
    st = state.copy()
    for s in range(nsweeps):
        for i in range(n):
            acc = int(st[i]) * 31 + s + 7
            for k in range(4):
                c = cols[i, k]
                if c < maxnb and nbs[i, c, 1] != 0.0 and int(nbs[i, c, 0]) != i:
                    acc += (k + 1) * int(st[int(nbs[i, c, 0])])
            st[i] = acc % 2147483647
    return st


def _pipeline(nbs, state, nsweeps, C, kinds, loc, rng, greedy):
    """The kernel's protocol, one Python object per chain, scheduled by `rng` (greedy: run the picked
    chain until it blocks -- maximal run-ahead, which is what exposes a missing dependency)."""
    n = nbs.shape[0]
    nch = (n + C - 1) // C
    st = state.copy()
    prog = np.zeros(nch, dtype=np.int64)
    ring = [[(0, 0)] * D for _ in range(nch)]              # (tag, value) per slot, written by the chain
    lens = [min(C, n - c * C) for c in range(nch)]

    class Chain:
        def __init__(self, c):
            self.c, self.t, self.T = c, 0, nsweeps * lens[c]
            self.pred = nch - 1 if c == 0 else c - 1
            self.phase = "fetch"                                # fetch(t) then run(t)
            self.wn, self.pending, self.result = [0] * 4, None, 0
            self.w1 = st[c * C]                                 # own word of step 0
            self.w = self.w2 = 0
            self.first = True

        def sp(self, t):
            return divmod(t, lens[self.c])

        def try_ll(self, need, pj):
            tag, val = ring[self.pred][pj % D]
            if tag == need:
                return "ok", val
            return ("lost", None) if tag > need else ("wait", None)

        def step(self):
            """advance one phase if possible; False if blocked"""
            c = self.c
            if self.t >= self.T:
                return False
            s, p = self.sp(self.t)
            i = c * C + p
            if self.phase == "fetch":
                # all waits first (the kernel blocks inside fetch), then the loads
                vals, pend = [0] * 4, None
                for k in range(4):
                    kind = int(kinds[i, k])
                    cj, pj = int(loc[i, k]) >> 16, int(loc[i, k]) & 0xFFFF
                    need = s * C + pj + 1 - (C if kind in (K_LL_OLD, K_MEM_OLD) else 0)
                    mem = kind >= K_MEM_SELF
                    if kind in (K_LL_CUR, K_LL_OLD):
                        if need <= 0:
                            mem = True
                        else:
                            r, v = self.try_ll(need, pj)
                            if r == "ok":
                                vals[k] = v
                            elif r == "lost":
                                mem = True
                            else:
                                pend = (k, need, cj, pj)
                    if mem:
                        if kind != K_MEM_SELF and need > 0 and prog[cj] < need:
                            return False
                        vals[k] = st[cj * C + pj]
                    elif kind == K_PREV:
                        vals[k] = self.result
                    elif kind == K_NEXT:
                        vals[k] = None                          # w2, read below
                p2 = 0 if p + 1 == lens[c] else p + 1
                self.w2 = st[c * C + p2] if self.t + 1 < self.T else 0
                for k in range(4):
                    if vals[k] is None:
                        vals[k] = self.w2
                self.wn, self.pending = vals, pend
                self.w, self.w1 = self.w1, self.w2              # rotation (kernel: after fetch)
                self.phase = "run"
                return True
            # run
            if self.pending is not None:
                k, need, cj, pj = self.pending
                r, v = self.try_ll(need, pj)
                if r == "wait":
                    return False
                if r == "lost":
                    if prog[cj] < need:
                        return False
                    v = st[cj * C + pj]
                self.wn[k] = v
                self.pending = None
            acc = int(self.w) * 31 + s + 7
            for k in range(4):
                if int(kinds[i, k]) != K_ZERO:
                    acc += (k + 1) * int(self.wn[k])
            self.result = acc % 2147483647
            assert st[i] == self.w, "own word changed under the chain"
            st[i] = self.result
            tag = s * C + p + 1
            ring[c][p % D] = (tag, self.result)
            last = p + 1 == lens[c]
            if last or (p + 1) % G == 0:
                prog[c] = (s + 1) * C if last else tag
            self.t += 1
            self.phase = "fetch"
            return True

    chains = [Chain(c) for c in range(nch)]
    while True:
        live = [ch for ch in chains if ch.t < ch.T]
        if not live:
            break
        order = rng.permutation(len(live))
        moved = False
        for q in order:
            ch = live[q]
            if ch.step():
                moved = True
                if greedy:
                    while ch.step():
                        pass
                if rng.randint(3) == 0:
                    break
        assert moved, "deadlock: no chain can advance"
    return st


def _random_graph(rng, n, torus=None):
    J = sps.dok_matrix((n, n))
    if torus:
        L = torus
        for y in range(L):
            for x in range(L):
                i = y * L + x
                for j in (y * L + (x + 1) % L, ((y + 1) % L) * L + x):
                    J[min(i, j), max(i, j)] = rng.uniform(-2, 2)
    else:
        deg = np.zeros(n, dtype=int)
        for _ in range(3 * n):
            a, b = rng.randint(n, size=2)
            if a != b and (min(a, b), max(a, b)) not in J and deg[a] < 3 and deg[b] < 3:
                J[min(a, b), max(a, b)] = rng.uniform(-2, 2)
                deg[a] += 1
                deg[b] += 1
        for i in rng.choice(n, n // 3, replace=False):
            J[i, i] = rng.uniform(-1, 1)
    return tools.GenerateNeighbors(n, J, 4)


@pytest.mark.parametrize("seed", range(8))
def test_chain_protocol_equals_sequential_sweep_on_random_graphs(seed):
    rng = np.random.RandomState(4200 + seed)
    n = 40 + 7 * seed
    nbs = _random_graph(rng, n)
    state = rng.randint(1, 1000, size=n).astype(np.int64)
    want = _sequential(nbs, state, 5)
    for C in (0, 4, 5, 9, 16, n):
        got_C, kinds, loc, per = device.chain_plan(nbs, C)
        if C and n % C == 1:
            assert got_C == 0                                   # a last chain of one spin is not planned
            continue
        assert got_C >= 4 and (C == 0 or got_C == min(C, n)) and per > 0
        for greedy in (False, True):
            got = _pipeline(nbs, state, 5, got_C, kinds, loc, rng, greedy)
            assert np.array_equal(got, want), (C, greedy)


@pytest.mark.parametrize("L,C", [(6, 0), (6, 4), (8, 0), (8, 16), (8, 5), (12, 0)])
def test_chain_protocol_on_torus(L, C):
    rng = np.random.RandomState(L * 100 + C)
    nbs = _random_graph(rng, L * L, torus=L)
    got_C, kinds, loc, per = device.chain_plan(nbs, C)
    if C == 0:
        assert got_C == L, "a lattice row is the natural chain"
        assert per < 2.5 * L, "rows pipeline: about one step per spin of a row per sweep"
        # the row pipeline uses registers and hand-over rings only, plus the old value of the row below
        i = 2 * L + 3
        ks = sorted(int(k) for k in kinds[i])
        assert ks == [K_PREV, K_NEXT, K_LL_CUR, K_MEM_OLD]
        assert sorted(int(k) for k in kinds[3]) == [K_PREV, K_NEXT, K_LL_OLD, K_MEM_OLD]      # row 0: up is the last row, old
    state = rng.randint(1, 1000, size=L * L).astype(np.int64)
    want = _sequential(nbs, state, 4)
    for greedy in (False, True):
        assert np.array_equal(_pipeline(nbs, state, 4, got_C, kinds, loc, rng, greedy), want)


def test_chain_plan_path_graph_and_rejects():
    """A path graph (the dataflow kernel's unit count explodes on chain-like graphs in natural
    order) is one chain of registers; maxnb > 4 has no plan."""
    n = 50
    J = sps.dok_matrix((n, n))
    for i in range(n - 1):
        J[i, i + 1] = 1.0 + i
    nbs = tools.GenerateNeighbors(n, J, 2)
    C, kinds, loc, per = device.chain_plan(nbs, 0)
    assert C == n and set(int(k) for k in kinds.ravel()) <= {K_ZERO, K_PREV, K_NEXT}
    with pytest.raises(ValueError):
        device.chain_plan(np.zeros((8, 5, 2)), 0)
