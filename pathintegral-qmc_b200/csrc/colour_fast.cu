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
// "needs a uniform" of the 4 neighbour-disagreement masks are evaluated for all 64 lanes at once,
// each recognised as one of the 27 regular functions and computed in <= 3 LOP3 per 32 lanes;
// (3) the few words with lanes that need a uniform are resolved cooperatively by the warp (one
// Philox block per thread).  Semantics: oracle_qa_colour /
// oracle_sa_colour (oracle/piqmc_oracle.c part 3), bit for bit.
#include <stdlib.h>

#include "colour_device.cuh"

namespace {

constexpr int FAST_MAX_LAG = 32;     // largest stagger of the last chunk, in ticket periods

struct FastArgs {
    uint64_t *words;            // [N][nrows]
    const PiqmcUnitRec *recs;   // per sweep (or shared): one record per member, in ticket order
    const float *pate;          // [members][16] in-slice energy difference of every z-pattern, same order
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
    unsigned int *err;          // watchdog word: a wait that outlasts watchdog_ns sets it and gives up
    unsigned long long watchdog_ns;
    int force_generic;          // testing: evaluate every decision function by the generic pattern loop
    uint64_t valid, top;        // lane masks: all `lanes` slices / the last slice (kernel constants: no per-pass arithmetic)
    // several replicas per word (SEG): seg_S segments of seg_P lanes; masks of the first, second and last lane of
    // every segment, and of one whole segment
    int seg_P, seg_S;
    uint64_t seg_low, seg_l1, seg_top, seg_ones;
    // staggered row chunks (piqmc_qa_colour_results): chunk c runs (c * lag16) / 16 ticket periods behind chunk 0,
    // so the chunks finish one after the other and the host downloads each while the others still sweep.
    // Row chunks never interact, so any stagger leaves the result and the topological ticket order intact.
    // Chunk c is active in the ticket periods [lag(c), lag(c) + nper), nper = sweeps + ramp periods of the colouring;
    // only active (period, chunk) pairs get tickets.  With lmax = lag(last chunk) <= nper the periods fall into a
    // ramp-up [0, lmax) (chunks 0 .. nact_up[q] - 1), a steady part [lmax, nper) (all chunks) and a ramp-down
    // [nper, nper + lmax) (chunks clo_dn[q - nper] .. nchunks - 1); pref_* = first ticket of each ramp period.
    int lag16, lmax, nper;
    unsigned int base_steady, base_down;
    unsigned int pref_up[FAST_MAX_LAG + 1], pref_dn[FAST_MAX_LAG + 1];
    unsigned char nact_up[FAST_MAX_LAG + 1], clo_dn[FAST_MAX_LAG + 1];
    unsigned int *chunk_count;  // [nchunks] units of the launch's last sweep that have finished (null: no signal)
    unsigned int *chunk_flag;   // [nchunks] mapped host memory: set when a chunk has finished all its sweeps
};

// MINB = resident blocks per SM the register allocation is capped for (7 -> 72 registers, 8 -> 64, 9 -> 56):
// more resident blocks hide the per-unit latencies (ticket, record, flags) better, which matters
// most when a unit has little work (few rows).
// PIPE: staggered row chunks with a completion signal per chunk (piqmc_qa_colour_results); a separate
// instantiation, so that the plain kernel pays nothing for it.
template <bool QA, int TROT, int MINB, bool SEG = false, bool PIPE = false>
__global__ void __launch_bounds__(FAST_THREADS, MINB) colour_sweep_fast(const FastArgs a)
{
    __shared__ SpinTable tab;
    __shared__ uint2 queues[FAST_WARPS][QCAP];
    __shared__ int s_unit[4];                    // period q, row chunk, member index m of this block's ticket
    constexpr int NC = QA ? 3 : 1;

    // Tickets run over "periods" of N spins.  With a static colouring the member list is sorted by
    // (level mod D, level) and sweepoff[m] = level div D, D = 1 + the largest level gap across an
    // edge: period q then holds level rho of sweep q next to level rho+D of sweep q-1, ..., i.e.
    // consecutive sweeps overlap as far as the dependencies allow and ticket order is still a
    // topological order (a unit only waits for smaller tickets).
    if (threadIdx.x == 0) {                      // one thread takes and decodes the ticket for the block
        const unsigned int t = atomicAdd(a.ticket, 1u) - a.ticket_base;
        const unsigned int per_sweep = (unsigned int)a.nspins * (unsigned int)a.nchunks;
        unsigned int q, rem, nact = (unsigned int)a.nchunks, clo = 0u;
        if (!PIPE || (t >= a.base_steady && t < a.base_down)) {   // all chunks active (every ticket when lag16 == 0)
            const unsigned int u = t - a.base_steady;
            q = u / per_sweep;
            rem = u - q * per_sweep;
            q += (unsigned int)a.lmax;
        } else if (t < a.base_steady) {                            // ramp-up: the first nact_up[q] chunks
            q = 0u;
            while (a.pref_up[q + 1] <= t) q++;
            rem = t - a.pref_up[q];
            nact = a.nact_up[q];
        } else {                                                   // ramp-down: chunks clo_dn[j] and up
            unsigned int j = 0u;
            while (a.pref_dn[j + 1] <= t) j++;
            rem = t - a.pref_dn[j];
            clo = a.clo_dn[j];
            nact -= clo;
            q = (unsigned int)a.nper + j;
        }
        const unsigned int m = rem / nact;
        s_unit[0] = (int)q;
        s_unit[1] = (int)(clo + rem - m * nact);
        s_unit[2] = (int)m;
    }
    __syncthreads();
    const int q = s_unit[0], chunk = s_unit[1];
    const unsigned int m = (unsigned int)s_unit[2];
    // everything the unit needs about its spin in ONE 48-byte record (one L2 round trip instead
    // of the member -> neighbour table -> level chain of dependent loads)
    const int4 *rp = reinterpret_cast<const int4 *>(a.recs + (a.per_sweep_lists ? (size_t)q * a.nspins : 0) + m);
    const int4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
    // the pattern energy of this lane (warps 0 .. NC-1 build the tables), requested together with the record
    const float e0 = __ldg(a.pate + ((a.per_sweep_lists ? (size_t)q * a.nspins : 0) + m) * 16 + (threadIdx.x & 15));
    const int i = r0.x;
    const int s = q - r0.y - (PIPE ? ((chunk * a.lag16) >> 4) : 0);
    if (s < 0 || s >= a.nsweeps) return;                           // ramp-up / ramp-down periods
    const int nb[4] = {r0.z, r0.w, r1.x, r1.y};
    const float Jn[4] = {__int_as_float(r1.z), __int_as_float(r1.w), __int_as_float(r2.x), __int_as_float(r2.y)};
    const uint32_t deps = (uint32_t)r2.z;                           // byte n: 0 none, 1 wait s-1, 2 wait s
    const uint32_t pad = (uint32_t)r2.w;                            // column permutation and signs
    const uint32_t tag = a.tag0 + (uint32_t)s + 1u;
    const uint32_t sweep = a.sweep0 + (uint32_t)s;
    const int nrows = a.nrows, maxnb = a.maxnb, lanes = a.lanes;

    // ---- warps 0..NC-1 build the decision tables while warp 3 waits until this unit's inputs
    //      are final (see the header comment); one barrier joins them
    const int warp = threadIdx.x >> 5;
    if (warp < NC) {
        build_table_warp_pre<QA>(tab, warp, e0, Jn, pad, a.jp2[s], a.invT[s], a.force_generic != 0, QA && a.global_moves);
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
                unsigned int polls = 0;
                unsigned long long t0 = 0ull;
                while ((int32_t)(ld_acquire(flag) - want) < 0) {
                    if (a.poll_ns) __nanosleep(a.poll_ns);
                    if ((++polls & 4095u) == 0u) {             // watchdog: an error, not a hang (the unit then
                        unsigned long long now;                // runs on stale inputs; the host discards the state)
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (t0 == 0ull) t0 = now;
                        if (*(volatile unsigned int *)a.err != 0u) break;
                        if (now - t0 > a.watchdog_ns) {
                            atomicCAS(a.err, 0u, 1u);
                            break;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();

    uint32_t sgn[4];                              // sign of the sorted couplings, as 32-lane masks
#pragma unroll
    for (int n = 0; n < 4; n++) sgn[n] = ((pad >> (8 + n)) & 1u) ? 0xFFFFFFFFu : 0u;
    // names of the decision functions: byte 2c = "accept by sign", byte 2c+1 = "... or needs a uniform"
    const uint64_t fnames = *reinterpret_cast<const uint64_t *>(tab.names);
    const uint64_t lane1tab = *reinterpret_cast<const uint64_t *>(tab.lane1);

    const auto thr_tab = [&](uint32_t c, uint32_t pat) -> uint32_t { return tab.thr[c][pat]; };
    const int rbeg = chunk * a.rows_per_block;
    const int rend = min(nrows, rbeg + a.rows_per_block);
    const uint64_t valid = a.valid;
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
            for (int n = 0; n < 4; n++)      // self entries and unused columns point at the zero row
                wn_nx[n] = __ldcg(words + (size_t)nb[n] * nrows + row);
        }
    }
    for (int base = rbeg; base < rend; base += FAST_THREADS) {     // block-uniform trip count
        const int row = base + threadIdx.x;
        const bool live = row < rend;
        uint64_t *wrow = words + row;
        const uint64_t w = w_nx;
        uint64_t z[4];
#pragma unroll
        for (int n = 0; n < 4; n++) {        // self entries (local fields) read the zero row: w ^ 0;
            const uint64_t sg = ((uint64_t)sgn[n] << 32) | sgn[n];
            z[n] = w ^ wn_nx[n] ^ sg;        // unused columns: any value does (the functions ignore it)
            asm volatile("" : "+l"(z[n]));   // keep z in registers: no recomputation inside the class loop
        }
        {
            const int rown = row + FAST_THREADS;
            w_nx = 0;
            if (rown < rend) {
                w_nx = __ldcg(words + (size_t)i * nrows + rown);
#pragma unroll
                for (int n = 0; n < 4; n++)
                    wn_nx[n] = __ldcg(words + (size_t)nb[n] * nrows + rown);
            }
        }
        const uint32_t prow_warp = a.row0 + (uint32_t)(base + (threadIdx.x & ~31)) * (SEG ? (uint32_t)a.seg_S : 1u);

        // "accept by sign" and "accept by sign or needs a uniform" of every Trotter class for all 64
        // lanes at once, by name (fnames byte 2c / 2c+1), through two evaluation sites in a rolled
        // loop over the classes (the hot code stays within the instruction cache).  No pattern of a class needs a uniform for ~70% of
        // (spin, class) pairs at T << J: such a function is named FID_NONE and skipped.
        uint64_t result;
        if (!QA || TROT == 0) {
            uint64_t C[3], todo, XL = 0ull, XR = 0ull;
            uint64_t flip1 = 0ull;
            if (!QA) {
                // ---- SA: no Trotter terms, one class
                todo = live ? valid : 0ull;
                C[0] = C[1] = C[2] = todo;
            } else if (SEG) {
                // ---- several replicas per word: the same rules segment by segment.  The last slice
                //      and slice 1 of every segment are broadcast over their segment by a multiplication
                //      (the products of the segments' base bits with 2^P - 1 do not overlap).
                const uint64_t ones = a.seg_ones;
                const uint64_t bl = ((w & a.seg_top) >> (a.seg_P - 1)) * ones;
                const uint64_t br_old = ((w & a.seg_l1) >> 1) * ones;
                XL = (w ^ bl) & ~a.seg_top;
                // slice 1 of every segment first (its right neighbour is itself: class = left bit),
                // straight from the decision functions of classes 0 and 1
                const uint32_t fa0 = (uint32_t)fnames & 0xFFu, fb0 = (uint32_t)(fnames >> 8) & 0xFFu;
                const uint32_t fa1 = (uint32_t)(fnames >> 16) & 0xFFu, fb1 = (uint32_t)(fnames >> 24) & 0xFFu;
                const uint64_t F0 = eval_fn(fa0, &tab.hacc[0], z);
                const uint64_t F1 = (fa1 == fa0 && fa0 != FID_GENERIC) ? F0 : eval_fn(fa1, &tab.hacc[1], z);
                uint64_t N0 = 0ull, N1 = 0ull;
                if (fb0 != FID_NONE) N0 = eval_fn(fb0, &tab.hall[0], z) & ~F0;
                if (fb1 != FID_NONE) N1 = eval_fn(fb1, &tab.hall[1], z) & ~F1;
                const uint64_t l1 = live ? a.seg_l1 : 0ull;
                flip1 = ((XL & F1) | (~XL & F0)) & l1;
                const uint64_t need1 = ((XL & N1) | (~XL & N0)) & l1;
                if (__any_sync(0xffffffffu, need1 != 0ull)) {            // ~1% of the words
                    for (int g = 0; g < a.seg_S; g++) {                  // warp-uniform trip count
                        const int k = g * a.seg_P + 1;
                        if ((need1 >> k) & 1ull) {
                            const uint32_t c1 = (uint32_t)(XL >> k) & 1u;
                            if (lane_uniform(1, (uint32_t)i, sweep, a.row0 + (uint32_t)(row * a.seg_S + g), a.k0, a.k1) <
                                tab.thr[c1][pattern_at(z, k)])
                                flip1 |= 1ull << k;
                        }
                    }
                }
                const uint64_t br_new = (((w ^ flip1) & a.seg_l1) >> 1) * ones;
                XR = ((w ^ br_new) & ~a.seg_low) | ((w ^ br_old) & a.seg_low);   // slice 0 sees the old slice 1
                todo = live ? (valid & ~a.seg_l1) : 0ull;
                C[0] = ~(XL | XR) & todo;
                C[1] = (XL ^ XR) & todo;
                C[2] = (XL & XR) & todo;
            } else {
                // ---- reference Trotter neighbours: slices P-1 (old value for everyone; itself for lane
                //      P-1) and 1 (old for lane 0, itself for lane 1, new for lanes >= 2): lane 1 is
                //      decided first, straight from the truth tables.
                const uint64_t bl = (w & a.top) ? ~0ull : 0ull;
                const uint64_t br_old = (w & 2ull) ? ~0ull : 0ull;
                XL = (w ^ bl) & ~a.top;
                const uint32_t c1 = (uint32_t)(XL >> 1) & 1u;               // right neighbour of lane 1 is itself
                const uint32_t p1 = pattern_at(z, 1);
                if (live) {
                    const uint32_t t1 = (uint32_t)(lane1tab >> (16u * c1 + p1));    // bit 0: accept, bit 32: accept or need
                    const uint32_t t2 = (uint32_t)(lane1tab >> (32u + 16u * c1 + p1));
                    if (t1 & 1u) flip1 = 2ull;
                    else if (t2 & 1u)                                        // ~1% of the words
                        if (lane_uniform(1, (uint32_t)i, sweep, a.row0 + (uint32_t)row, a.k0, a.k1) < tab.thr[c1][p1])
                            flip1 = 2ull;
                }
                const uint64_t br_new = br_old ^ (flip1 ? ~0ull : 0ull);
                XR = ((w ^ br_new) & ~1ull) | ((w ^ br_old) & 1ull);        // lane 0 sees the old bit 1
                todo = live ? (valid & ~2ull) : 0ull;
                C[0] = ~(XL | XR) & todo;
                C[1] = (XL ^ XR) & todo;
                C[2] = (XL & XR) & todo;
            }
            uint64_t ACC = 0ull, NEED = 0ull, V = 0ull;
            uint32_t last = FID_NONE;                       // name of the function V holds
            bool may_need = true;
            // Most word passes of a cold anneal have every lane of the warp in Trotter class 0 (the slices of a
            // replica agree): one vote, one evaluation, no loop (ncu: 1.4 populated classes per pass on average,
            // but three trips through the loop head).
            if (QA && !__any_sync(0xffffffffu, (C[1] | C[2]) != 0ull)) {
                const uint32_t fa = (uint32_t)fnames & 0xFFu, fb = ((uint32_t)fnames >> 8) & 0xFFu;
                V = eval_fn(fa, &tab.hacc[0], z);
                ACC = V & C[0];
                if (fb != FID_NONE) NEED = eval_fn(fb, &tab.hall[0], z) & ~V & C[0];
                else may_need = false;                       // (block-uniform) no pattern of class 0 draws: no vote below
            } else
#pragma unroll 1
            for (int c = 0; c < NC; c++) {
                const uint64_t Cc = (c == 0) ? C[0] : (c == 1 ? C[1] : C[2]);
                // a class no lane of the warp is in costs nothing (late in the anneal the slices of a
                // replica agree: only class 0 is populated)
                if (!__any_sync(0xffffffffu, Cc != 0ull)) continue;
                const uint32_t names = (uint32_t)(fnames >> (16 * c));
                const uint32_t fa = names & 0xFFu, fb = (names >> 8) & 0xFFu;
                // a function that was just evaluated is reused (early in the anneal J_perp is small
                // and the classes decide alike)
                if (fa != last || fa == FID_GENERIC) {
                    V = eval_fn(fa, &tab.hacc[c], z);
                    last = fa;
                }
                ACC |= V & Cc;
                if (fb != FID_NONE) NEED |= eval_fn(fb, &tab.hall[c], z) & ~V & Cc;
            }
            if (may_need && __any_sync(0xffffffffu, NEED != 0))
                ACC |= resolve_draws<QA>(NEED, z, XL, XR, thr_tab, queue, (uint32_t)i, sweep, prow_warp, a.k0, a.k1,
                                         SEG ? a.seg_P : 64, SEG ? a.seg_S : 1);
            result = w ^ flip1 ^ ACC;
        } else {
            // ---- periodic Trotter neighbours k-1, k+1: even slices first, then odd slices (the lanes
            //      of one pass do not see each other).  With an odd slice count lanes 0 and P-1 are
            //      both even and adjacent: lane P-1 gets a pass of its own after the even pass.
            uint64_t F[NC], N[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                F[c] = eval_fn((uint32_t)(fnames >> (16 * c)) & 0xFFu, &tab.hacc[c], z);
                const uint32_t fb = (uint32_t)(fnames >> (16 * c + 8)) & 0xFFu;
                N[c] = 0ull;
                if (fb != FID_NONE) N[c] = eval_fn(fb, &tab.hall[c], z) & ~F[c];
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
                uint64_t ACC = ((C0 & F[0]) | (C1 & F[1]) | (C2 & F[NC - 1])) & sel;
                const uint64_t NEED = ((C0 & N[0]) | (C1 & N[1]) | (C2 & N[NC - 1])) & sel;
                // lanes of this pass have not flipped yet, so their rows of z are still current
                if (__any_sync(0xffffffffu, NEED != 0))
                    ACC |= resolve_draws<QA>(NEED, z, XL, XR, thr_tab, queue, (uint32_t)i, sweep, prow_warp, a.k0, a.k1);
                cur ^= ACC;
            }
            result = cur;
        }
        if (QA && a.global_moves) {
            // ---- world-line move: flip the spin in all slices at once.  The Trotter terms cancel;
            //      ediff = sum_p count[p] * insum[p] over the patterns of the UPDATED word (float64
            //      accumulation in pattern order, as oracle_qa_colour states it).
            const uint64_t d = w ^ result;                       // own flips toggle every disagreement bit
            uint64_t x[4] = {0ull, 0ull, 0ull, 0ull};            // disagreement masks in TABLE order
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint64_t sg = ((uint64_t)sgn[k] << 32) | sgn[k];
#pragma unroll
                for (int n = 0; n < 4; n++)
                    if (((pad >> (2 * k)) & 3u) == (uint32_t)n) x[n] = z[k] ^ sg;
            }
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
    if (threadIdx.x == 0) {                                         // release orders the block's stores
        st_release(a.done + (size_t)i * a.nchunks + chunk, tag);   // (cumulative through the barrier)
        if (PIPE && a.chunk_count != nullptr && tag == a.tag0 + (uint32_t)a.nsweeps) {
            // last sweep of the launch: count the spins of this row chunk that are final; the unit that completes
            // the chunk tells the host (fence; atomic = release, atomic; fence = acquire, both system-wide: the
            // copy engine then reads final words)
            __threadfence_system();
            if (atomicAdd(a.chunk_count + chunk, 1u) == (unsigned int)a.nspins - 1u) {
                __threadfence_system();
                *(volatile unsigned int *)(a.chunk_flag + chunk) = 1u;
            }
        }
    }
}

}  // namespace

// rows per block: the per-spin table is built once per block, so more rows per block is less
// overhead, but fewer independent chunks; keep >= 4 chunks when the state allows
// (measured on B200, 256x256 P=64: 4 chunks is the sweet spot from 512 to 4096 rows)
static int fast_rows_per_block(const piqmc_ctx *c)
{
    int rpb = 512;
    while (rpb > FAST_THREADS && (c->nrows + rpb - 1) / rpb < 4) rpb >>= 1;
    if (const char *e = getenv("PIQMC_ROWS_PER_BLOCK")) {          // tuning knob
        const int v = atoi(e);
        if (v >= FAST_THREADS && v % FAST_THREADS == 0) rpb = v;
    }
    return rpb;
}

// One launch holds (sweeps + ramp periods) * N * chunks units; the ticket arithmetic is 32-bit and a
// grid has at most 2^31 - 1 blocks.  A colouring with many levels and a small level gap (a path graph
// in natural order) can exceed that even for one sweep: the caller then takes another kernel.
int fast_chunk_rows(const piqmc_ctx *c) { return fast_rows_per_block(c); }

bool launch_fast_fits(const piqmc_ctx *c, int nperiods_extra)
{
    const int rpb = fast_rows_per_block(c);
    const size_t per_sweep = (size_t)c->nspins * ((c->nrows + rpb - 1) / rpb);
    return (size_t)(1 + nperiods_extra) * per_sweep < ((size_t)1 << 31);
}

// Runs `nsweeps` sweeps in as few launches as the grid-size limit allows (normally one).
// members/level: device arrays, level-major spin order and level per spin; either one list for
// all sweeps or one per sweep.  d_jp2/d_invT: per sweep.
int launch_fast_sweeps(piqmc_ctx *c, int qa, int trotter, int nsweeps, const PiqmcUnitRec *d_recs, const float *d_pate,
                       int nperiods_extra, int per_sweep_lists, const float *d_jp2, const float *d_invT,
                       uint64_t seed, uint32_t row0, uint32_t sweep0)
{
    if (nsweeps <= 0) return PIQMC_OK;
    const int rpb = fast_rows_per_block(c);
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
    if (!c->d_err) {
        PIQMC_CUDA(cudaMalloc(&c->d_err, sizeof(unsigned int)));
        PIQMC_CUDA(cudaMemsetAsync(c->d_err, 0, sizeof(unsigned int), c->stream));
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
    a.valid = (c->lanes >= 64) ? ~0ull : ((1ull << c->lanes) - 1ull);
    a.top = 1ull << (c->lanes - 1);
    a.seg_P = c->seg_P;
    a.seg_S = c->seg_S;
    a.seg_low = a.seg_l1 = a.seg_top = 0ull;
    a.seg_ones = (c->seg_P >= 64) ? ~0ull : ((1ull << c->seg_P) - 1ull);
    for (int g = 0; g < c->seg_S; g++) {
        a.seg_low |= 1ull << (g * c->seg_P);
        a.seg_l1 |= 2ull << (g * c->seg_P);
        a.seg_top |= 1ull << (g * c->seg_P + c->seg_P - 1);
    }
    a.nchunks = nchunks;
    a.rows_per_block = rpb;
    a.per_sweep_lists = per_sweep_lists;
    a.k0 = (uint32_t)seed;
    a.k1 = (uint32_t)(seed >> 32);
    a.row0 = row0;
    a.global_moves = c->global_moves;
    a.poll_ns = 0;
    if (const char *e = getenv("PIQMC_POLL_NS")) a.poll_ns = (unsigned int)atoi(e);
    a.force_generic = 0;
    if (const char *e = getenv("PIQMC_FORCE_GENERIC_FN")) a.force_generic = atoi(e);
    a.err = c->d_err;
    a.watchdog_ns = 20000000000ull;
    if (const char *e = getenv("PIQMC_WATCHDOG_MS")) a.watchdog_ns = (unsigned long long)atoll(e) * 1000000ull;

    const size_t per_sweep = nflags;
    PIQMC_REQUIRE(launch_fast_fits(c, nperiods_extra), PIQMC_EINVAL,
                  "colouring with %d ramp periods x %zu units per sweep does not fit one launch of the dataflow kernel",
                  nperiods_extra, per_sweep);
    const int max_sweeps = (int)std::max<long long>(1, (long long)(((size_t)1 << 30) / per_sweep) - nperiods_extra);
    // staggered chunks with a completion signal per chunk: only when the run is one launch
    a.lag16 = 0;
    a.chunk_count = nullptr;
    a.chunk_flag = nullptr;
    c->pipe_armed = 0;
    if (c->pipe_request && qa && !trotter && c->seg_S == 1 && nchunks <= c->pipe_chunk_cap && nchunks <= 255 &&
        nsweeps <= max_sweeps) {              // (the one instantiation with the stagger code: QA, reference Trotter)
        // the stagger of the last chunk stays within the table size and within the periods of one chunk
        int lag16 = c->pipe_lag16;
        const int cap = std::min(FAST_MAX_LAG, nsweeps + nperiods_extra);
        if (nchunks > 1 && (((nchunks - 1) * lag16) >> 4) > cap) lag16 = (cap << 4) / (nchunks - 1);
        a.lag16 = c->pipe_lag16 = lag16;
        a.chunk_count = c->d_chunk_count;
        a.chunk_flag = c->d_chunk_flag;
        c->pipe_armed = 1;
    }
    unsigned int ticket_base = 0;
    for (int s0 = 0; s0 < nsweeps; s0 += max_sweeps) {
        const int ns = std::min(max_sweeps, nsweeps - s0);
        a.recs = d_recs + (per_sweep_lists ? (size_t)s0 * c->nspins : 0);
        a.pate = d_pate + (per_sweep_lists ? (size_t)s0 * c->nspins * 16 : 0);
        a.nsweeps = ns;
        a.jp2 = d_jp2 + s0;
        a.invT = d_invT + s0;
        a.sweep0 = sweep0 + (uint32_t)s0;
        a.tag0 = c->flow_tag + (uint32_t)s0;
        a.ticket_base = ticket_base;
        const size_t nunits = (size_t)(ns + nperiods_extra) * per_sweep;   // staggered or not: active pairs only
        // ticket -> (period, chunk) tables of the stagger (trivial when lag16 == 0)
        a.nper = ns + nperiods_extra;
        a.lmax = ((nchunks - 1) * a.lag16) >> 4;
        {
            unsigned int acc = 0;
            for (int q = 0; q <= a.lmax; q++) {                  // ramp-up: chunks with lag <= q
                int nact = 0;
                while (nact < nchunks && ((nact * a.lag16) >> 4) <= q) nact++;
                a.pref_up[q] = acc;
                a.nact_up[q] = (unsigned char)nact;
                if (q < a.lmax) acc += (unsigned int)c->nspins * (unsigned int)nact;
            }
            a.base_steady = acc;
            acc += (unsigned int)(a.nper - a.lmax) * (unsigned int)per_sweep;
            a.base_down = acc;
            for (int j = 0; j <= a.lmax; j++) {                  // ramp-down: chunks with lag + nper > nper + j
                int clo = 0;
                while (clo < nchunks && ((clo * a.lag16) >> 4) <= j) clo++;
                a.pref_dn[j] = acc;
                a.clo_dn[j] = (unsigned char)clo;
                if (j < a.lmax) acc += (unsigned int)c->nspins * (unsigned int)(nchunks - clo);
            }
            if (acc != (unsigned int)nunits) {
                piqmc_set_error("internal: stagger tables cover %u tickets, grid has %zu", acc, nunits);
                return PIQMC_EINVAL;
            }
        }
        ticket_base += (unsigned int)nunits;
        dim3 block(FAST_THREADS), grid((unsigned)nunits);
        int minb = (qa && !trotter) ? 9 : 8;      // measured on B200: 9 blocks/SM (56 registers) best from 512 to 4096 rows
        if (const char *e = getenv("PIQMC_MINB")) minb = atoi(e);      // tuning knob
        if (!qa) {
            if (minb == 7) colour_sweep_fast<false, 0, 7><<<grid, block, 0, c->stream>>>(a);
            else           colour_sweep_fast<false, 0, 8><<<grid, block, 0, c->stream>>>(a);
        } else if (trotter) {
            if (minb == 7) colour_sweep_fast<true, 1, 7><<<grid, block, 0, c->stream>>>(a);
            else           colour_sweep_fast<true, 1, 8><<<grid, block, 0, c->stream>>>(a);
        } else {
            if (c->pipe_armed)   colour_sweep_fast<true, 0, 9, false, true><<<grid, block, 0, c->stream>>>(a);
            else if (c->seg_S > 1) colour_sweep_fast<true, 0, 8, true><<<grid, block, 0, c->stream>>>(a);
            else if (minb == 7)  colour_sweep_fast<true, 0, 7><<<grid, block, 0, c->stream>>>(a);
            else if (minb == 9)  colour_sweep_fast<true, 0, 9><<<grid, block, 0, c->stream>>>(a);
            else if (minb == 10) colour_sweep_fast<true, 0, 10><<<grid, block, 0, c->stream>>>(a);
            else                 colour_sweep_fast<true, 0, 8><<<grid, block, 0, c->stream>>>(a);
        }
        c->launches++;
        PIQMC_CUDA(cudaGetLastError());
    }
    c->flow_tag += (uint32_t)nsweeps;
    return PIQMC_OK;
}
