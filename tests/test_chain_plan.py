"""CPU tests of the chain pipeline's plan (piqmc_chain_plan, csrc/api.cu) and of its hand-over protocol.

The CUDA kernel (csrc/chain_kernels.cu) cuts the natural-order sweep of a lattice into chains (lattice
rows) and lets them run concurrently; what keeps it equal to the sequential sweep is (a) the per-slot
kinds the host computes and (b) the protocol: inside a block a result is handed to the following chain
as soon as that chain has read the previous one, between blocks tagged packets travel in rings of 16, own row and row below requested three steps
ahead (they may land any time between the request and two steps later), progress words published
every G steps by the first chain of a band, which guard packet-slot reuse and the row below.  Here that
protocol is executed by a Python model under adversarial schedules and must (1) never deadlock and
(2) end in exactly the state of the sequential sweep.
"""
import numpy as np
import pytest
import scipy.sparse as sps

import piqmc.tools as tools
from piqmc import device

K_ZERO, K_LEFT, K_RIGHT, K_UP, K_DOWN = range(5)
DR, DG, PRE = 4, 16, 3


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


def _pipeline(n, state, nsweeps, C, kinds, wrap, nbands, G, rng, greedy):
    """The kernel's protocol, one Python object per chain, scheduled by `rng` (greedy: run the picked
    chain until it blocks -- maximal run-ahead, which is what exposes a missing guard)."""
    nch = n // C
    st = state.copy()                                           # global state words
    T = nsweeps * C
    first = [(nch * b) // nbands for b in range(nbands + 1)]
    band_of = np.zeros(nch, dtype=int)
    for b in range(nbands):
        band_of[first[b]:first[b + 1]] = b
    prog = np.zeros(nbands, dtype=np.int64)                     # published by the first chain of a band
    gll = [[(0, 0)] * DG for _ in range(nbands)]                # (tag, value), written by the last chain of a band
    res = [[None] * DR for _ in range(nch)]                     # result rings
    res_full = [[-1] * DR for _ in range(nch)]                  # step whose result the slot holds
    res_taken = [[-1] * DR for _ in range(nch)]                 # last step whose result the consumer has read

    class Chain:
        def __init__(self, c):
            self.c, self.t = c, 0
            b = band_of[c]
            self.band = b
            self.w = c - first[b]
            self.nW = first[b + 1] - first[b]
            self.succ_in = self.w + 1 < self.nW
            self.ext_out = (not self.succ_in) and (b + 1 < nbands or wrap)
            self.ext_in = self.w == 0 and (b > 0 or wrap)
            self.has_down = c + 1 < nch or wrap
            self.nband = b + 1 if b + 1 < nbands else 0
            self.pband = b - 1 if b > 0 else nbands - 1
            self.in_off = C if b == 0 else 0
            self.out_off = C if self.nband == 0 else 0
            self.pv = 0
            self.phase = -1                                     # -1: prologue, 0: guards + request, 1: row above + compute + publish
            self.own, self.down = {}, {}                        # landed words by step
            self.flight = []                                    # (step requested at, step fetched for)
            self.fpos = 0
            self.res_prev = None

        def request(self, tf, at):
            if tf <= T:
                self.flight.append((at, tf, self.fpos))
            self.fpos = (self.fpos + 1) % C

        def land_one(self, item):
            at, tf, pos = item
            self.own[tf] = st[self.c * C + pos]
            if self.has_down:
                dc = self.c + 1 if self.c + 1 < nch else 0
                self.down[tf] = st[dc * C + pos]

        def land(self, upto, force):
            """requests made at steps <= upto must land now (force); the others may land (adversary)"""
            keep = []
            for item in self.flight:
                if (force and item[0] <= upto) or rng.randint(3) == 0:
                    self.land_one(item)
                else:
                    keep.append(item)
            self.flight = keep

        def step(self):
            c, t = self.c, self.t
            if t >= T:
                return False
            s, p = divmod(t, C)
            i = c * C + p
            tm = t % DR
            if self.phase == -1:
                # prologue.  The last chain of a torus reads chain 0's NEW values as its row below: chain 0
                # must have got that far before the first words are requested
                if self.ext_out and self.has_down and self.nband == 0:
                    need = min(PRE, T)
                    if prog[0] < need:
                        return False
                    self.pv = prog[0]
                self.res_prev = st[c * C + C - 1]               # the chain's last spin
                for tf in range(PRE):
                    self.request(tf, tf - PRE)                  # one group per step, as in the loop
                self.land(upto=-PRE - 1, force=False)           # ... any of which may land at once
                self.phase = 0
                return True
            if self.phase == 0:
                if self.ext_out:
                    need = t + self.out_off + 1 - DG
                    if self.has_down:
                        nd = t + PRE + 1 if self.nband == 0 else t + PRE + 1 - C
                        need = max(need, nd)
                    need = min(need, T)
                    if need > 0 and self.pv < need:
                        if prog[self.nband] < need:
                            return False
                        self.pv = prog[self.nband]
                self.request(t + PRE, t)
                self.land(upto=t - 2, force=True)               # wait_group 2
                self.phase = 1
                return True
            # row above
            up = None
            if self.w > 0:
                if res_full[c - 1][tm] != t:
                    return False
                up = res[c - 1][tm]
            elif self.ext_in:
                if t < self.in_off:
                    up = st[(nch - 1) * C + p]
                else:
                    tp = t - self.in_off
                    tag, val = gll[self.pband][tp % DG]
                    assert tag <= tp + 1, "a packet was overwritten before it was read"
                    if tag != tp + 1:
                        return False
                    up = val
            if self.succ_in and t >= 1 and res_taken[c][(t - 1) % DR] < t - 1:
                return False                                    # the consumer has not read the result of step t - 1 yet
            assert t in self.own and t + 1 in self.own, "own-row word not landed"
            src = {K_ZERO: 0, K_LEFT: self.res_prev, K_RIGHT: self.own[t + 1], K_UP: up,
                   K_DOWN: self.down.get(t)}
            acc = int(self.own[t]) * 31 + s + 7
            for k in range(4):
                kind = int(kinds[i, k])
                if kind != K_ZERO:
                    assert src[kind] is not None, (c, t, kind)
                    acc += (k + 1) * int(src[kind])
            result = acc % 2147483647
            assert st[i] == self.own[t], "own word changed under the chain"
            # publish
            res[c][tm] = result
            res_full[c][tm] = t
            if self.ext_out:
                gll[self.band][t % DG] = (t + 1, result)
            if self.w > 0:
                res_taken[c - 1][tm] = t
            st[i] = result
            self.res_prev = result
            if self.w == 0 and (t + 1 == T or (t + 1) % G == 0 or t + 1 == PRE):
                prog[self.band] = t + 1
            self.own.pop(t, None)
            self.down.pop(t, None)
            self.t += 1
            self.phase = 0
            return True

    chains = [Chain(c) for c in range(nch)]
    while True:
        live = [ch for ch in chains if ch.t < T]
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


def _lattice(rng, H, W, torus_x=True, torus_y=True, fields=False, holes=0.0):
    """H rows of W spins, row-major; optional periodic wraps, local fields and missing bonds"""
    n = H * W
    J = sps.dok_matrix((n, n))
    for y in range(H):
        for x in range(W):
            i = y * W + x
            nb = []
            if x + 1 < W or (torus_x and W > 2):
                nb.append(y * W + (x + 1) % W)
            if y + 1 < H or (torus_y and H > 2):
                nb.append(((y + 1) % H) * W + x)
            for j in nb:
                if rng.uniform() >= holes:
                    J[min(i, j), max(i, j)] = rng.uniform(-2, 2)
    if fields:
        for i in rng.choice(n, n // 3, replace=False):
            J[i, i] = rng.uniform(-1, 1)
    return tools.GenerateNeighbors(n, J, 4 if not fields else 5)


@pytest.mark.parametrize("H,W,tx,ty", [(8, 8, True, True), (5, 12, True, True), (12, 8, True, False),
                                       (9, 8, False, True), (16, 8, False, False), (3, 16, True, True)])
def test_chain_plan_kinds(H, W, tx, ty):
    rng = np.random.RandomState(H * 100 + W)
    nbs = _lattice(rng, H, W, tx, ty)
    C, kinds, wrap = device.chain_plan(nbs, 0)
    assert C == W, "a lattice row is the natural chain"
    assert wrap == (ty and H > 2)
    cols = _sorted_cols(nbs)
    n = H * W
    for i in range(n):
        y, x = divmod(i, W)
        for k in range(4):
            c = cols[i, k]
            kind = int(kinds[i, k])
            if c >= nbs.shape[1] or nbs[i, c, 1] == 0.0:
                assert kind == K_ZERO
                continue
            j = int(nbs[i, c, 0])
            yj, xj = divmod(j, W)
            if yj == y:
                left = xj == x - 1 or (x == 0 and xj == W - 1)
                assert kind == (K_LEFT if left else K_RIGHT)
            else:
                up = yj == y - 1 or (y == 0 and yj == H - 1 and H > 2)
                assert xj == x and kind == (K_UP if up else K_DOWN)


def test_chain_plan_rejects():
    """No plan: maxnb > 4 with a fifth live column, random graphs, chains shorter than 8, a chain length
    that does not divide the number of spins."""
    rng = np.random.RandomState(3)
    nbs = _lattice(rng, 8, 8)
    assert device.chain_plan(nbs, 4)[0] == 0
    assert device.chain_plan(nbs, 12)[0] == 0
    assert device.chain_plan(nbs, 16)[0] == 0                   # two lattice rows per chain: (y, x) -- (y+1, x) is inside a chain
    assert device.chain_plan(nbs, 8)[0] == 8
    n = 64
    J = sps.dok_matrix((n, n))
    for _ in range(100):
        a, b = rng.randint(n, size=2)
        if a != b:
            J[min(a, b), max(a, b)] = 1.0
    assert device.chain_plan(tools.GenerateNeighbors(n, J, 12), 0)[0] == 0
    small = _lattice(rng, 4, 4)
    assert device.chain_plan(small, 0)[0] == 0
    # local fields on a lattice: the self entry reads as "none"
    withf = _lattice(rng, 8, 8, torus_x=False, torus_y=False, fields=True)
    C, kinds, wrap = device.chain_plan(withf, 0)
    assert C == 8 and not wrap


@pytest.mark.parametrize("H,W,tx,ty,nbands,G", [
    (8, 8, True, True, 1, 2), (8, 8, True, True, 2, 2), (8, 8, True, True, 8, 1), (8, 8, True, True, 3, 2),
    (12, 8, True, True, 4, 2), (5, 16, True, True, 2, 4), (6, 16, True, False, 3, 4), (20, 8, False, True, 5, 2),
    (7, 8, False, False, 2, 2), (3, 24, True, True, 3, 4), (32, 32, True, True, 2, 8), (16, 16, True, True, 1, 4),
])
def test_chain_protocol_equals_sequential_sweep(H, W, tx, ty, nbands, G):
    rng = np.random.RandomState(H * 1000 + W * 10 + nbands)
    nbs = _lattice(rng, H, W, tx, ty, holes=0.1)
    n = H * W
    C, kinds, wrap = device.chain_plan(nbs, W)
    assert C == W
    state = rng.randint(1, 1000, size=n).astype(np.int64)
    nsweeps = 4
    want = _sequential(nbs, state, nsweeps)
    for greedy in (False, True):
        got = _pipeline(n, state, nsweeps, C, kinds, wrap, nbands, G, rng, greedy)
        assert np.array_equal(got, want), greedy
