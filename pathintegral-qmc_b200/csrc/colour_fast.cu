// colour_fast.cu -- the production Metropolis sweep kernel (maxnb <= 4; QA with the reference's
// Trotter neighbours, or SA), one launch for a whole run of sweeps.
//
// Work unit = block = (sweep s, spin i, chunk of rows).  Within a sweep the spins are ordered by
// colour class ("level"); a unit may run as soon as, for the same chunk of rows,
//     every coupled neighbour of a LOWER level has finished sweep s       (it reads their new value)
//     every coupled neighbour of a HIGHER level has finished sweep s-1    (it reads their old value
//                                                                          and they have read ours)
//     the spin itself has finished sweep s-1.
// Completion is published in done[spin][chunk] = tag(s) with release/acquire ordering, so the
// sweeps need no kernel boundary and no grid barrier: colour classes, and successive sweeps,
// overlap like a wavefront.  Units are handed out by an atomic ticket in (sweep, level) order,
// so a unit only ever waits for units with smaller tickets, which are already resident or done:
// the scheme cannot deadlock whatever the hardware's block dispatch order.
//
// Per unit: (1) the per-spin decision tables are built in shared memory while the unit waits;
// (2) each thread owns one 64-lane word per pass: the 2*NC boolean functions "accept by sign" /
// "needs a uniform" of the 4 neighbour-disagreement masks are evaluated for all 64 lanes at once
// in algebraic normal form; (3) the few words with lanes that need a uniform are resolved
// cooperatively by the warp (one Philox block per thread).  Semantics: oracle_qa_colour /
// oracle_sa_colour (oracle/piqmc_oracle.c part 3), bit for bit.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int FAST_THREADS = 128;
constexpr int FAST_WARPS = FAST_THREADS / 32;
constexpr int QCAP = 64;            // pooled draw requests per warp and pass
constexpr int COOP_MIN = 4;         // words with at least this many needy lanes are drawn by the whole warp

__device__ __forceinline__ float flip_sign(float negJ2, uint32_t bit)
{
    return __int_as_float(__float_as_int(negJ2) ^ (int)(bit << 31));
}

__device__ __forceinline__ uint32_t pick(const u32x4 &r, int q)
{
    return q == 0 ? r.x : (q == 1 ? r.y : (q == 2 ? r.z : r.w));
}

// the uniform of (row, lane, spin, sweep): Philox block (spin, lane>>2, sweep, row), word lane&3
__device__ __forceinline__ uint32_t lane_uniform(int lane, uint32_t spin, uint32_t sweep, uint32_t prow,
                                                 uint32_t k0, uint32_t k1)
{
    return pick(philox4x32_10(spin, (uint32_t)(lane >> 2) | (PIQMC_STREAM_SWEEP << 16), sweep, prow, k0, k1),
                lane & 3);
}

struct SpinTable {
    uint32_t cacc[3][16];    // ANF coefficient masks (0 / ~0) of "accept by sign", per Trotter class
    uint32_t cneed[3][16];   // ... of "needs a uniform"
    uint32_t thr[3][16];     // acceptance threshold of pattern p in class c
    uint32_t hacc[3], hneed[3];   // truth tables (bit p)
    float insum[16];         // in-slice energy difference of pattern p (no Trotter term): world-line moves
};

// Per-spin decision tables, built by warp c for Trotter class c without any block barrier
// (the caller synchronises once).  Trotter class of a lane = number of Trotter neighbours it
// disagrees with (0,1,2): tsum = -2*jp2, +0, +2*jp2.
template <bool QA>
__device__ __forceinline__ void build_table_warp(SpinTable &tab, int c, const float (&Jn)[4], float jp2,
                                                 float invT)
{
    const int lane = threadIdx.x & 31;
    const int p = lane & 15;                     // lanes 16..31 mirror lanes 0..15
    float e = 0.0f;
#pragma unroll
    for (int n = 0; n < 4; n++)                  // unused columns carry J = 0: adding +-0 changes nothing
        e = __fadd_rn(e, flip_sign(-2.0f * Jn[n], (uint32_t)(p >> n) & 1u));
    if (c == 0 && lane < 16) tab.insum[p] = e;
    if (QA) {
        const float tsum = (c == 0) ? -2.0f * jp2 : (c == 1 ? 0.0f : 2.0f * jp2);   // exact
        e = __fadd_rn(e, tsum);
    }
    e = __fadd_rn(e, 0.0f);
    const bool acc = QA ? (e > 0.0f) : (e >= 0.0f);
    const float x = __fmul_rn(e, invT);
    const bool need = !acc && (x >= PIQMC_XCUT);
    if (lane < 16) tab.thr[c][p] = need ? colour_thresh(x) : 0u;
    const uint32_t ha = __ballot_sync(0xffffffffu, acc) & 0xFFFFu;     // truth tables (bit p)
    const uint32_t hn = __ballot_sync(0xffffffffu, need) & 0xFFFFu;
    if (lane == 0) {
        tab.hacc[c] = ha;
        tab.hneed[c] = hn;
    }
    // Moebius transform: ANF coefficient of monomial S = parity of the truth table over subsets of S
    const int S = p;
    uint32_t sub = 1u;                           // bit T set <=> T is a subset of S
    if (S & 1) sub |= sub << 1;
    if (S & 2) sub |= sub << 2;
    if (S & 4) sub |= sub << 4;
    if (S & 8) sub |= sub << 8;
    const uint32_t coef = (__popc((lane < 16 ? ha : hn) & sub) & 1) ? 0xFFFFFFFFu : 0u;
    if (lane < 16) tab.cacc[c][S] = coef;
    else           tab.cneed[c][S] = coef;
}

__device__ __forceinline__ uint32_t pattern_at(const uint64_t (&x)[4], int k)
{
    return (uint32_t)((x[0] >> k) & 1) | (uint32_t)((x[1] >> k) & 1) << 1 |
           (uint32_t)((x[2] >> k) & 1) << 2 | (uint32_t)((x[3] >> k) & 1) << 3;
}

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src)
{
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
    const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

// Resolve the lanes in NEED (those whose Metropolis test needs a uniform); returns the accepted
// ones.  Must be called by all 32 threads of a warp.  Needy lanes cluster: the slices of one
// replica are strongly correlated, so a word has either none or most of its 64 lanes needy.
//   (a) a word with >= COOP_MIN needy lanes is broadcast to the warp; thread l draws slices
//       2l and 2l+1 (one Philox block), results come back by warp OR-reduction;
//   (b) the remaining scattered lanes are pooled in a per-warp queue and drawn 32 at a time.
// All loops have warp-uniform trip counts (a per-thread `while (mask)` would make the hardware
// run the threads' iterations one after another).
template <bool QA>
__device__ __forceinline__ uint64_t resolve_draws(uint64_t NEED, const uint64_t (&x)[4], uint64_t XL, uint64_t XR,
                                                  const SpinTable &tab, uint2 *queue, uint32_t spin,
                                                  uint32_t sweep, uint32_t prow_warp, uint32_t k0, uint32_t k1)
{
    const int lane = threadIdx.x & 31;
    uint64_t ACC = 0;
    uint32_t big = __ballot_sync(0xffffffffu, __popcll(NEED) >= COOP_MIN);
    while (big) {                                                   // warp-uniform
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint64_t needs = shfl64(NEED, src);
        uint64_t xs[4];
#pragma unroll
        for (int n = 0; n < 4; n++) xs[n] = shfl64(x[n], src);
        const uint64_t xls = QA ? shfl64(XL, src) : 0ull, xrs = QA ? shfl64(XR, src) : 0ull;
        const int ka = 2 * lane;
        const uint32_t need2 = (uint32_t)(needs >> ka) & 3u;
        uint32_t acc2 = 0;
        if (need2) {
            const u32x4 r = philox4x32_10(spin, (uint32_t)(ka >> 2) | (PIQMC_STREAM_SWEEP << 16), sweep,
                                          prow_warp + (uint32_t)src, k0, k1);
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const int k = ka + b;
                const uint32_t c = QA ? (uint32_t)((xls >> k) & 1) + (uint32_t)((xrs >> k) & 1) : 0u;
                const uint32_t u = (ka & 2) ? (b ? r.w : r.z) : (b ? r.y : r.x);
                if (((need2 >> b) & 1u) && u < tab.thr[c][pattern_at(xs, k)]) acc2 |= 1u << b;
            }
        }
        const uint32_t lo = __reduce_or_sync(0xffffffffu, lane < 16 ? acc2 << (2 * lane) : 0u);
        const uint32_t hi = __reduce_or_sync(0xffffffffu, lane >= 16 ? acc2 << (2 * lane - 32) : 0u);
        if (lane == src) {
            ACC |= ((uint64_t)hi << 32) | lo;
            NEED = 0;
        }
    }
    while (true) {
        const int n = __popcll(NEED);
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) break;
        const int excl = incl - n;
        uint64_t mask = NEED;
        int pos = excl;
#pragma unroll
        for (int j = 0; j < COOP_MIN - 1; j++) {                    // n < COOP_MIN here
            if (mask && pos < QCAP) {
                const int k = __ffsll((long long)mask) - 1;
                mask &= mask - 1;
                const uint32_t c = QA ? (uint32_t)((XL >> k) & 1) + (uint32_t)((XR >> k) & 1) : 0u;
                queue[pos++] = make_uint2(tab.thr[c][pattern_at(x, k)], (uint32_t)lane | ((uint32_t)k << 5));
            }
        }
        const uint64_t taken = NEED ^ mask;
        __syncwarp();
        const int T = total < QCAP ? total : QCAP;
        uint32_t res[QCAP / 32];
#pragma unroll
        for (int r = 0; r < QCAP / 32; r++) {
            res[r] = 0u;
            if (r * 32 < T) {                                       // warp-uniform
                bool a = false;
                const int it = r * 32 + lane;
                if (it < T) {
                    const uint2 q = queue[it];
                    a = lane_uniform((int)(q.y >> 5), spin, sweep, prow_warp + (q.y & 31u), k0, k1) < q.x;
                }
                res[r] = __ballot_sync(0xffffffffu, a);
            }
        }
        uint64_t tk = taken;
        pos = excl;
#pragma unroll
        for (int j = 0; j < COOP_MIN - 1; j++) {
            if (tk) {
                const int k = __ffsll((long long)tk) - 1;
                tk &= tk - 1;
                const uint32_t word = (pos < 32) ? res[0] : res[QCAP / 32 - 1];
                if ((word >> (pos & 31)) & 1u) ACC |= 1ull << k;
                pos++;
            }
        }
        NEED = mask;
        __syncwarp();
    }
    return ACC;
}

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct FastArgs {
    uint64_t *words;            // [N][nrows]
    const PiqmcUnitRec *recs;   // per sweep (or shared): one record per member, in ticket order
    int nsweeps;                // sweeps covered by this launch
    const float *jp2, *invT;    // per sweep
    uint32_t *done;             // [N][nchunks] tag of the last finished sweep
    unsigned int *ticket;
    int nspins, nrows, maxnb, lanes, nchunks, rows_per_block;
    int per_sweep_lists;        // members/level advance by N per sweep
    int global_moves;           // QA: attempt a world-line move after the local moves of every spin
    uint32_t k0, k1, row0, sweep0, tag0;
    unsigned int ticket_base;   // units handed out by earlier launches of the same run
    unsigned int poll_ns;       // back-off between polls of a completion flag (0 = spin)
};

// MINB = resident blocks per SM the register allocation is capped for (7 -> 72 registers, 8 -> 64):
// more resident blocks hide the per-unit latencies (ticket, record, flags) better, which matters
// most when a unit has little work (few rows).
template <bool QA, int TROT, int MINB>
__global__ void __launch_bounds__(FAST_THREADS, MINB) colour_sweep_fast(const FastArgs a)
{
    __shared__ SpinTable tab;
    __shared__ uint2 queues[FAST_WARPS][QCAP];
    __shared__ unsigned int s_ticket;
    constexpr int NC = QA ? 3 : 1;

    if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1u) - a.ticket_base;
    __syncthreads();
    const unsigned int t = s_ticket;
    // Tickets run over "periods" of N spins.  With a static colouring the member list is sorted by
    // (level mod D, level) and sweepoff[m] = level div D, D = 1 + the largest level gap across an
    // edge: period q then holds level rho of sweep q next to level rho+D of sweep q-1, ..., i.e.
    // consecutive sweeps overlap as far as the dependencies allow and ticket order is still a
    // topological order (a unit only waits for smaller tickets).
    const unsigned int per_sweep = (unsigned int)a.nspins * (unsigned int)a.nchunks;
    const int q = (int)(t / per_sweep);
    const unsigned int rem = t - (unsigned int)q * per_sweep;
    const int chunk = (int)(rem % (unsigned int)a.nchunks);
    const unsigned int m = rem / (unsigned int)a.nchunks;
    // everything the unit needs about its spin in ONE 48-byte record (one L2 round trip instead
    // of the member -> neighbour table -> level chain of dependent loads)
    const int4 *rp = reinterpret_cast<const int4 *>(a.recs + (a.per_sweep_lists ? (size_t)q * a.nspins : 0) + m);
    const int4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
    const int i = r0.x;
    const int s = q - r0.y;
    if (s < 0 || s >= a.nsweeps) return;                           // ramp-up / ramp-down periods
    const int nb[4] = {r0.z, r0.w, r1.x, r1.y};
    const float Jn[4] = {__int_as_float(r1.z), __int_as_float(r1.w), __int_as_float(r2.x), __int_as_float(r2.y)};
    const uint32_t deps = (uint32_t)r2.z;                           // byte n: 0 none, 1 wait s-1, 2 wait s
    const uint32_t tag = a.tag0 + (uint32_t)s + 1u;
    const uint32_t sweep = a.sweep0 + (uint32_t)s;
    const int nrows = a.nrows, maxnb = a.maxnb, lanes = a.lanes;

    // ---- warps 0..NC-1 build the decision tables while warp 3 waits until this unit's inputs
    //      are final (see the header comment); one barrier joins them
    const int warp = threadIdx.x >> 5;
    if (warp < NC) {
        build_table_warp<QA>(tab, warp, Jn, a.jp2[s], a.invT[s]);
    } else if (warp == FAST_WARPS - 1) {
        const int ql = threadIdx.x & 31;
        if (ql <= 4) {
            int j = i;
            uint32_t want = tag - 1u;                  // lane 4: the spin itself, previous sweep
            bool must = true;
            if (ql < 4) {
                const uint32_t d = (deps >> (8 * ql)) & 0xFFu;
                j = nb[ql];
                must = d != 0u;
                if (d == 2u) want = tag;
            }
            if (must) {
                const uint32_t *flag = a.done + (size_t)j * a.nchunks + chunk;
                while ((int32_t)(ld_acquire(flag) - want) < 0)
                    if (a.poll_ns) __nanosleep(a.poll_ns);
            }
        }
    }
    __syncthreads();

    const int rbeg = chunk * a.rows_per_block;
    const int rend = min(nrows, rbeg + a.rows_per_block);
    const uint64_t valid = (lanes >= 64) ? ~0ull : ((1ull << lanes) - 1ull);
    uint2 *queue = queues[threadIdx.x >> 5];
    uint64_t *words = a.words;

    // software pipeline: the 5 words of the next pass are requested before this pass computes.
    // L2-only loads (ld.global.cg): another unit may have rewritten these words during this very
    // launch and L1 is not coherent.
    uint64_t w_nx = 0, wn_nx[4] = {0, 0, 0, 0};
    {
        const int row = rbeg + threadIdx.x;
        if (row < rend) {
            w_nx = __ldcg(words + (size_t)i * nrows + row);
#pragma unroll
            for (int n = 0; n < 4; n++)
                if (nb[n] != i) wn_nx[n] = __ldcg(words + (size_t)nb[n] * nrows + row);
        }
    }
    for (int base = rbeg; base < rend; base += FAST_THREADS) {     // block-uniform trip count
        const int row = base + threadIdx.x;
        const bool live = row < rend;
        uint64_t *wrow = words + row;
        const uint64_t w = w_nx;
        uint64_t x[4];
#pragma unroll
        for (int n = 0; n < 4; n++)          // self entries (local fields) were not loaded: x = w ^ 0;
            x[n] = w ^ wn_nx[n];             // unused columns: any x does (the tables ignore that bit)
        {
            const int rown = row + FAST_THREADS;
            w_nx = 0;
            if (rown < rend) {
                w_nx = __ldcg(words + (size_t)i * nrows + rown);
#pragma unroll
                for (int n = 0; n < 4; n++)
                    if (nb[n] != i) wn_nx[n] = __ldcg(words + (size_t)nb[n] * nrows + rown);
            }
        }
        const uint32_t prow_warp = a.row0 + (uint32_t)(base + (threadIdx.x & ~31));

        // ---- monomials of the 4 neighbour-disagreement masks (shared by all boolean functions)
        uint32_t mlo[16], mhi[16];
        mlo[0] = mhi[0] = 0xFFFFFFFFu;
#pragma unroll
        for (int n = 0; n < 4; n++) {
            const uint32_t lo = (uint32_t)x[n], hi = (uint32_t)(x[n] >> 32);
#pragma unroll
            for (int S = 0; S < (1 << n); S++) {
                mlo[(1 << n) + S] = S ? (mlo[S] & lo) : lo;
                mhi[(1 << n) + S] = S ? (mhi[S] & hi) : hi;
            }
        }
        // "accept by sign" / "needs a uniform" of Trotter class c for all 64 lanes at once: a 4-input
        // boolean function of the disagreement masks in algebraic normal form.  No pattern of a class
        // needs a uniform for ~70% of (spin, class) pairs at T << J: block-uniform skip.
        auto anf_acc = [&](int c) -> uint64_t {
            uint32_t lo = tab.cacc[c][0], hi = lo;
#pragma unroll
            for (int S = 1; S < 16; S++) {
                const uint32_t ca = tab.cacc[c][S];
                lo ^= ca & mlo[S];
                hi ^= ca & mhi[S];
            }
            return ((uint64_t)hi << 32) | lo;
        };
        auto anf_need = [&](int c) -> uint64_t {
            if (tab.hneed[c] == 0u) return 0ull;
            uint32_t lo = tab.cneed[c][0], hi = lo;
#pragma unroll
            for (int S = 1; S < 16; S++) {
                const uint32_t cn = tab.cneed[c][S];
                lo ^= cn & mlo[S];
                hi ^= cn & mhi[S];
            }
            return ((uint64_t)hi << 32) | lo;
        };

        uint64_t result;
        if (!QA) {
            // ---- SA: no Trotter terms, one class
            const uint64_t todo = live ? valid : 0ull;
            uint64_t ACC = anf_acc(0) & todo;
            const uint64_t NEED = anf_need(0) & todo;
            if (__any_sync(0xffffffffu, NEED != 0))
                ACC |= resolve_draws<QA>(NEED, x, 0ull, 0ull, tab, queue, (uint32_t)i, sweep, prow_warp, a.k0, a.k1);
            result = w ^ ACC;
        } else if (TROT == 0) {
            // ---- reference Trotter neighbours: slices P-1 (old value for everyone; itself for lane
            //      P-1) and 1 (old for lane 0, itself for lane 1, new for lanes >= 2): lane 1 is decided
            //      first, straight from the truth tables.
            uint64_t flips = 0;
            const uint64_t bl = ((w >> (lanes - 1)) & 1ull) ? ~0ull : 0ull;
            const uint64_t br_old = ((w >> 1) & 1ull) ? ~0ull : 0ull;
            const uint64_t XL = (w ^ bl) & ~(1ull << (lanes - 1));
            const uint32_t c1 = (uint32_t)(XL >> 1) & 1u;               // right neighbour of lane 1 is itself
            const uint32_t p1 = pattern_at(x, 1);
            if (live) {
                if ((tab.hacc[c1] >> p1) & 1u) flips = 2ull;
                else if ((tab.hneed[c1] >> p1) & 1u)                     // ~1% of the words
                    if (lane_uniform(1, (uint32_t)i, sweep, a.row0 + (uint32_t)row, a.k0, a.k1) < tab.thr[c1][p1])
                        flips = 2ull;
            }
            const uint64_t br_new = br_old ^ (flips ? ~0ull : 0ull);
            const uint64_t XR = ((w ^ br_new) & ~1ull) | ((w ^ br_old) & 1ull);   // lane 0 sees the old bit 1
            const uint64_t todo = live ? (valid & ~2ull) : 0ull;
            uint64_t ACC = 0, NEED = 0;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const uint64_t Cc = (c == 0 ? ~(XL | XR) : (c == 1 ? (XL ^ XR) : (XL & XR))) & todo;
                ACC |= anf_acc(c) & Cc;
                NEED |= anf_need(c) & Cc;
            }
            if (__any_sync(0xffffffffu, NEED != 0))
                ACC |= resolve_draws<QA>(NEED, x, XL, XR, tab, queue, (uint32_t)i, sweep, prow_warp, a.k0, a.k1);
            result = w ^ flips ^ ACC;
        } else {
            // ---- periodic Trotter neighbours k-1, k+1: even slices first, then odd slices (the lanes
            //      of one pass do not see each other).  With an odd slice count lanes 0 and P-1 are
            //      both even and adjacent: lane P-1 gets a pass of its own after the even pass.
            uint64_t Fa[NC], Fn[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                Fa[c] = anf_acc(c);
                Fn[c] = anf_need(c);
            }
            const uint64_t evens = 0x5555555555555555ull & valid, odds = 0xAAAAAAAAAAAAAAAAull & valid;
            const uint64_t top = 1ull << (lanes - 1);
            const bool oddP = (lanes & 1) != 0;
            uint64_t cur = w;
#pragma unroll 1
            for (int pass = 0; pass < 3; pass++) {
                uint64_t sel = pass == 0 ? (oddP ? (evens & ~top) : evens) : (pass == 1 ? (oddP ? top : 0ull) : odds);
                if (!live) sel = 0ull;
                const uint64_t Lw = ((cur << 1) | (cur >> (lanes - 1))) & valid;     // bit k = slice k-1
                const uint64_t Rw = ((cur >> 1) | (cur << (lanes - 1))) & valid;     // bit k = slice k+1
                const uint64_t XL = cur ^ Lw, XR = cur ^ Rw;
                const uint64_t C0 = ~(XL | XR), C1 = XL ^ XR, C2 = XL & XR;
                uint64_t ACC = ((C0 & Fa[0]) | (C1 & Fa[1]) | (C2 & Fa[NC - 1])) & sel;
                const uint64_t NEED = ((C0 & Fn[0]) | (C1 & Fn[1]) | (C2 & Fn[NC - 1])) & sel;
                // lanes of this pass have not flipped yet, so their rows of x are still current
                if (__any_sync(0xffffffffu, NEED != 0))
                    ACC |= resolve_draws<QA>(NEED, x, XL, XR, tab, queue, (uint32_t)i, sweep, prow_warp, a.k0, a.k1);
                cur ^= ACC;
            }
            result = cur;
        }
        if (QA && a.global_moves) {
            // ---- world-line move: flip the spin in all slices at once.  The Trotter terms cancel;
            //      ediff = sum_p count[p] * insum[p] over the patterns of the UPDATED word (float64
            //      accumulation in pattern order, as oracle_qa_colour states it).
            const uint64_t d = w ^ result;                       // own flips toggle every disagreement bit
            uint64_t xa[2][2], xb[2][2];                         // [literal value][variable]
#pragma unroll
            for (int n = 0; n < 2; n++) {
                const uint64_t v0 = (n < maxnb) ? (x[n] ^ d) : 0ull, v1 = (n + 2 < maxnb) ? (x[n + 2] ^ d) : 0ull;
                xa[1][n] = v0; xa[0][n] = ~v0;
                xb[1][n] = v1; xb[0][n] = ~v1;
            }
            double accd = 0.0;
#pragma unroll
            for (int pat = 0; pat < 16; pat++) {
                const uint64_t mt = xa[pat & 1][0] & xa[(pat >> 1) & 1][1] & xb[(pat >> 2) & 1][0] &
                                    xb[(pat >> 3) & 1][1] & valid;
                accd += (double)__popcll(mt) * (double)tab.insum[pat];
            }
            const float g = __fadd_rn((float)accd, 0.0f);
            bool flip = g > 0.0f;
            if (!flip) {
                const float xg = __fmul_rn(g, a.invT[s]);
                if (xg >= PIQMC_XCUT)
                    flip = philox4x32_10((uint32_t)i, PIQMC_STREAM_GLOBAL << 16, sweep, a.row0 + (uint32_t)row,
                                         a.k0, a.k1).x < colour_thresh(xg);
            }
            if (flip) result ^= valid;
        }
        if (live) wrow[(size_t)i * nrows] = result;
    }

    // ---- publish: all stores of the block happen-before the flag
    __syncthreads();
    if (threadIdx.x == 0)                                           // release orders the block's stores
        st_release(a.done + (size_t)i * a.nchunks + chunk, tag);   // (cumulative through the barrier)
}

}  // namespace

// Runs `nsweeps` sweeps in as few launches as the grid-size limit allows (normally one).
// members/level: device arrays, level-major spin order and level per spin; either one list for
// all sweeps or one per sweep.  d_jp2/d_invT: per sweep.
int launch_fast_sweeps(piqmc_ctx *c, int qa, int trotter, int nsweeps, const PiqmcUnitRec *d_recs,
                       int nperiods_extra, int per_sweep_lists, const float *d_jp2, const float *d_invT,
                       uint64_t seed, uint32_t row0, uint32_t sweep0)
{
    if (nsweeps <= 0) return PIQMC_OK;
    // rows per block: the per-spin table is built once per block, so more rows per block is less
    // overhead, but fewer independent chunks; keep >= 4 chunks when the state allows
    // (measured on B200, 256x256 P=64: 4 chunks is the sweet spot from 512 to 4096 rows)
    int rpb = 512;
    while (rpb > FAST_THREADS && (c->nrows + rpb - 1) / rpb < 4) rpb >>= 1;
    if (const char *e = getenv("PIQMC_ROWS_PER_BLOCK")) {          // tuning knob
        const int v = atoi(e);
        if (v >= FAST_THREADS && v % FAST_THREADS == 0) rpb = v;
    }
    const int nchunks = (c->nrows + rpb - 1) / rpb;
    const size_t nflags = (size_t)c->nspins * nchunks;
    if (c->flow_nchunks != nchunks || c->d_done == nullptr) {
        PIQMC_CUDA(cudaStreamSynchronize(c->stream));
        if (c->d_done) PIQMC_CUDA(cudaFree(c->d_done));
        c->d_done = nullptr;
        PIQMC_CUDA(cudaMalloc(&c->d_done, nflags * sizeof(uint32_t)));
        PIQMC_CUDA(cudaMemsetAsync(c->d_done, 0, nflags * sizeof(uint32_t), c->stream));
        c->flow_nchunks = nchunks;
        c->flow_tag = 0;
    }
    if (!c->d_ticket) {
        PIQMC_CUDA(cudaMalloc(&c->d_ticket, sizeof(unsigned int)));
    }
    PIQMC_CUDA(cudaMemsetAsync(c->d_ticket, 0, sizeof(unsigned int), c->stream));

    FastArgs a;
    a.words = c->d_words;
    a.done = c->d_done;
    a.ticket = c->d_ticket;
    a.nspins = c->nspins;
    a.nrows = c->nrows;
    a.maxnb = c->maxnb;
    a.lanes = c->lanes;
    a.nchunks = nchunks;
    a.rows_per_block = rpb;
    a.per_sweep_lists = per_sweep_lists;
    a.k0 = (uint32_t)seed;
    a.k1 = (uint32_t)(seed >> 32);
    a.row0 = row0;
    a.global_moves = c->global_moves;
    a.poll_ns = 0;
    if (const char *e = getenv("PIQMC_POLL_NS")) a.poll_ns = (unsigned int)atoi(e);

    const size_t per_sweep = nflags;
    const int max_sweeps = (int)std::max<long long>(1, (long long)(((size_t)1 << 30) / per_sweep) - nperiods_extra);
    unsigned int ticket_base = 0;
    for (int s0 = 0; s0 < nsweeps; s0 += max_sweeps) {
        const int ns = std::min(max_sweeps, nsweeps - s0);
        a.recs = d_recs + (per_sweep_lists ? (size_t)s0 * c->nspins : 0);
        a.nsweeps = ns;
        a.jp2 = d_jp2 + s0;
        a.invT = d_invT + s0;
        a.sweep0 = sweep0 + (uint32_t)s0;
        a.tag0 = c->flow_tag + (uint32_t)s0;
        a.ticket_base = ticket_base;
        const size_t nunits = (size_t)(ns + nperiods_extra) * per_sweep;
        ticket_base += (unsigned int)nunits;
        dim3 block(FAST_THREADS), grid((unsigned)nunits);
        const bool wide = c->nrows >= 2048;       // measured on B200: 7 blocks/SM best at 4096 rows, 8 at 512
        if (!qa) {
            if (wide) colour_sweep_fast<false, 0, 7><<<grid, block, 0, c->stream>>>(a);
            else      colour_sweep_fast<false, 0, 8><<<grid, block, 0, c->stream>>>(a);
        } else if (trotter) {
            if (wide) colour_sweep_fast<true, 1, 7><<<grid, block, 0, c->stream>>>(a);
            else      colour_sweep_fast<true, 1, 8><<<grid, block, 0, c->stream>>>(a);
        } else {
            if (wide) colour_sweep_fast<true, 0, 7><<<grid, block, 0, c->stream>>>(a);
            else      colour_sweep_fast<true, 0, 8><<<grid, block, 0, c->stream>>>(a);
        }
        c->launches++;
        PIQMC_CUDA(cudaGetLastError());
    }
    c->flow_tag += (uint32_t)nsweeps;
    return PIQMC_OK;
}
