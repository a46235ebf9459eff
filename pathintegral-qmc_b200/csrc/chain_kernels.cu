// chain_kernels.cu -- the natural-order Metropolis sweep as a pipeline of chains (maxnb <= 4; QA with the
// reference's Trotter neighbours, or SA), one launch for a whole run of sweeps.
//
// The sweep being reproduced is sequential: spins 0..N-1 in order, every spin seeing the new value of
// the neighbours visited before it and the old value of the others (piqmc/qmc.pyx:320-357, the order of
// the reference's per-spin-reset variant).  Its dependency graph on a lattice is a rigid wavefront: every
// (sweep, spin) is on a critical path, so the cost of ONE dependency hand-over is what bounds a sweep
// when there are few replicas per GPU.  colour_fast.cu pays an L2 flag round trip (~3.6 us) per
// wavefront step.  Here:
//
//   * the order is cut into contiguous CHAINS of C spins (a lattice row); one WARP owns one chain for
//     32 rows (replicas), one thread one row, for all sweeps of the launch.  Inside a chain the
//     dependency on the previous spin is a register; the old value of the next spin is the own word
//     of the next step, loaded once.
//   * the same position of the preceding chain ("up") arrives through a HAND-OVER RING: 16-byte
//     packets {lo, tag, hi, tag} per lane, in shared memory when both chains sit in one block
//     (8 chains per block), in L2 otherwise.  A packet validates itself (tag = sweep * C + pos + 1), so
//     the producer needs no fence and never waits; a consumer that finds a newer tag (the producer
//     ran ahead and reused the slot) falls back to the state word in global memory.
//   * every other dependency ("down": the old value of the following chain, and whatever an
//     arbitrary graph brings) is a state word guarded by a per-chain progress counter
//     (st.release / ld.acquire, published every 8 steps); these have a sweep of slack and are
//     prefetched one step ahead.
//   * the decision functions of a (schedule step, spin) -- by name, see colour_device.cuh -- are
//     computed ONCE for all replicas by chain_tables_kernel and staged into shared memory by cp.async,
//     16 steps ahead; acceptance thresholds are rebuilt by the warp only on the steps where a lane
//     needs a uniform.
//
// Blocks take (ring, chain group) tickets in ring-major order: a block only waits for blocks of its
// own ring, and all blocks of earlier rings are resident or done, so the scheme cannot deadlock
// whatever the dispatch order.  Every wait has a watchdog (err word + PIQMC_ECUDA instead of a hang).
// Semantics: oracle_qa_colour / oracle_sa_colour (oracle/piqmc_oracle.c part 3), bit for bit.
#include <stdlib.h>

#include "colour_device.cuh"

namespace {

constexpr int CH_W = 8;                  // chains (warps) per block
constexpr int CH_THREADS = 32 * CH_W;
constexpr int CH_D = 4;                  // hand-over ring depth (packets per lane)
constexpr int CH_B = 16;                 // steps per staged batch of records
constexpr unsigned FULL = 0xffffffffu;

struct ChainArgs {
    uint64_t *words;                     // [N + 1][nrows]
    const PiqmcChainStat *stat;          // [N]
    const PiqmcChainDyn *dyn;            // [schedule steps of this launch][N]
    const float *jp2, *invT;             // per schedule step of this launch
    uint32_t *prog;                      // [nrings][nchains] steps completed (sweep * C + pos + 1)
    uint4 *gll;                          // [nrings][bpr][CH_D][32] hand-over rings between blocks
    unsigned int *ticket, *err;
    int nspins, nrows, C, nchains, bpr, nsweeps, mcsteps;
    uint32_t gmask;                      // progress is published when ((pos + 1) & gmask) == 0 and at the chain end
    uint32_t k0, k1, row0, sweep0;
    unsigned long long watchdog_ns;
    uint64_t valid, top;
    int seg_P, seg_S;
    uint64_t seg_low, seg_l1, seg_top, seg_ones;
};

__device__ __forceinline__ uint4 ld_vol_v4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.volatile.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_vol_v4(uint4 *p, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    asm volatile("st.volatile.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}

// One poll of a wait loop went by.  Returns true when the wait must be given up: another warp has
// already reported a failure, or this wait has lasted longer than the watchdog allows.
__device__ __forceinline__ bool wait_expired(const ChainArgs &a, unsigned &polls, unsigned long long &t0, unsigned code)
{
    ++polls;
    if (polls > 16u) __nanosleep(polls > 256u ? 200u : 40u);
    if ((polls & 1023u) != 0u) return false;
    if (*(volatile unsigned int *)a.err != 0u) return true;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t0 == 0ull) t0 = now;
    else if (now - t0 > a.watchdog_ns) {
        atomicCAS(a.err, 0u, code);
        return true;
    }
    return false;
}

// The out-of-line (rare or slow) parts of a step.  They are kept out of the sweep loop so that its
// code stays within the instruction cache.

// wait until *flag >= need; the value seen, or 0xFFFFFFFF when the watchdog gave up
__device__ __noinline__ uint32_t chain_wait_progress(const ChainArgs &a, const uint32_t *flag, int32_t need)
{
    unsigned polls = 0;
    unsigned long long t0 = 0ull;
    uint32_t v;
    while ((int32_t)((v = ld_acquire(flag)) - (uint32_t)need) < 0)
        if (wait_expired(a, polls, t0, 2u)) return 0xFFFFFFFFu;
    return v;
}

// wait for the hand-over packet tagged `need`.  0: *val holds its word; 1: the slot holds a newer
// packet (take the state word instead); 2: the watchdog gave up
__device__ __noinline__ int chain_wait_packet(const ChainArgs &a, const uint4 *slot, int32_t need, uint64_t *val)
{
    unsigned polls = 0;
    unsigned long long t0 = 0ull;
    while (true) {
        const uint4 pk = ld_vol_v4(slot);
        const int32_t d1 = (int32_t)(pk.y - (uint32_t)need), d2 = (int32_t)(pk.w - (uint32_t)need);
        if (__all_sync(FULL, d1 == 0 && d2 == 0)) {
            *val = ((uint64_t)pk.z << 32) | pk.x;
            return 0;
        }
        if (__any_sync(FULL, d1 > 0 || d2 > 0)) return 1;
        if (wait_expired(a, polls, t0, 3u)) return 2;
    }
}

// acceptance thresholds of the 16 z-patterns in every Trotter class, by the warp: the sequence of
// float32 operations of build_table_warp (colour_device.cuh)
template <bool QA>
__device__ __noinline__ void chain_build_thr(uint32_t *thr, float J0, float J1, float J2, float J3, uint32_t pad,
                                             float jp2, float invT)
{
    const int lane = threadIdx.x & 31;
    const uint32_t p = lane & 15;
    const float Jz[4] = {J0, J1, J2, J3};
    float ex;
    const float e0 = pattern_energy(Jz, pad, p, ex);
    __syncwarp();                                   // readers of the previous table are done
#pragma unroll
    for (int r = 0; r < (QA ? 2 : 1); r++) {
        const int c = QA ? ((lane >> 4) + 2 * r) : 0;      // round 0: classes 0 | 1, round 1: class 2 | idle
        if (c < (QA ? 3 : 1) && (QA || lane < 16)) {
            float e = e0;
            if (QA) e = __fadd_rn(e, (c == 0) ? -2.0f * jp2 : (c == 1 ? 0.0f : 2.0f * jp2));
            e = __fadd_rn(e, 0.0f);
            const bool acc = QA ? (e > 0.0f) : (e >= 0.0f);
            const float x = __fmul_rn(e, invT);
            thr[c * 16 + p] = (!acc && x >= PIQMC_XCUT) ? colour_thresh(x) : 0u;
        }
    }
    __syncwarp();
}

// staging of the per-step records of one chain: CH_B steps per batch, by cp.async
struct ChainStage {
    uint32_t bs, bp, nissued;      // (sweep, position) of the first step of the next batch; batches issued
};
__device__ __noinline__ void chain_issue_batch(const ChainArgs &a, ChainStage &g, uint4 (*rec)[CH_B][4], int cbase, int len)
{
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
        const int q = lane + 32 * h, st = q >> 2, part = q & 3;
        uint32_t pl = g.bp + (uint32_t)st, sl = g.bs;
        while (pl >= (uint32_t)len) {
            pl -= (uint32_t)len;
            sl++;
        }
        if (sl < (uint32_t)a.nsweeps) {
            const uint32_t spin = (uint32_t)cbase + pl;
            const uint32_t f = sl / (uint32_t)a.mcsteps;
            const char *src = part < 2 ? (const char *)(a.stat + spin) + 16 * part
                                       : (const char *)(a.dyn + (size_t)f * a.nspins + spin) + 16 * (part - 2);
            cp_async16(&rec[g.nissued & 1u][st][part], src);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    g.nissued++;
    g.bp += CH_B;
    while (g.bp >= (uint32_t)len) {
        g.bp -= (uint32_t)len;
        g.bs++;
    }
}

// ---- decision functions of every (schedule step, spin), once for all replicas ----------------------
template <bool QA>
__global__ void __launch_bounds__(128) chain_tables_kernel(const PiqmcChainStat *stat, PiqmcChainDyn *dyn,
                                                           const float *jp2, const float *invT, int nspins,
                                                           int nsched, int force_generic)
{
    __shared__ SpinTable tabs[4];
    constexpr int NC = QA ? 3 : 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t g = (size_t)blockIdx.x * 4 + warp;
    if (g >= (size_t)nspins * nsched) return;
    const int f = (int)(g / nspins), i = (int)(g - (size_t)f * nspins);
    const PiqmcChainStat st = stat[i];
    const float Jz[4] = {st.J01[0], st.J01[1], st.J23[0], st.J23[1]};
    SpinTable &tab = tabs[warp];
    if (lane < 8) tab.names[lane] = 0;
    __syncwarp();
    for (int c = 0; c < NC; c++) build_table_warp<QA>(tab, c, Jz, st.pad, jp2[f], invT[f], force_generic != 0);
    __syncwarp();
    if (lane == 0) {
        PiqmcChainDyn d;
#pragma unroll
        for (int k = 0; k < 8; k++) d.names[k] = tab.names[k];
        d.lane1[0] = (uint16_t)tab.hacc[0];
        d.lane1[1] = (uint16_t)(QA ? tab.hacc[1] : 0u);
        d.lane1[2] = (uint16_t)tab.hall[0];
        d.lane1[3] = (uint16_t)(QA ? tab.hall[1] : 0u);
        d.hacc2 = (uint16_t)(QA ? tab.hacc[2] : 0u);
        d.hall2 = (uint16_t)(QA ? tab.hall[2] : 0u);
        d.J23[0] = st.J23[0];
        d.J23[1] = st.J23[1];
        d.spare = 0u;
        dyn[g] = d;
    }
}

// ---- the sweep kernel -------------------------------------------------------------------------------
template <bool QA, bool SEG, int MINB>
__global__ void __launch_bounds__(CH_THREADS, MINB) chain_sweep(const ChainArgs a)
{
    __shared__ uint4 s_ll[CH_W][CH_D][32];         // hand-over rings written by the chains of this block
    __shared__ uint4 s_rec[CH_W][2][CH_B][4];      // staged records: [0] loc, [1] kinds pad J0 J1, [2..3] PiqmcChainDyn
    __shared__ uint32_t s_thr[CH_W][48];
    __shared__ uint2 s_queue[CH_W][QCAP];
    __shared__ unsigned int s_ticket;
    constexpr int NC = QA ? 3 : 1;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1u);
    {
        uint4 *z = &s_ll[0][0][0];
        for (int k = threadIdx.x; k < CH_W * CH_D * 32; k += CH_THREADS) z[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();                               // the only block barrier: warps are independent from here on
    const unsigned int tk = s_ticket;
    const int ring = (int)(tk / (unsigned int)a.bpr);
    const int cg = (int)(tk - (unsigned int)ring * (unsigned int)a.bpr);
    const int chain = cg * CH_W + warp;
    if (chain >= a.nchains) return;

    const int C = a.C, nrows = a.nrows;
    const int cbase = chain * C;
    const int len = min(C, a.nspins - cbase);
    const uint32_t T = (uint32_t)a.nsweeps * (uint32_t)len;
    const int row = ring * 32 + lane;
    const bool live = row < nrows;
    uint64_t *wbase = a.words + (live ? row : 0);
    uint32_t *prog = a.prog + (size_t)ring * a.nchains;
    const uint32_t prow_warp = a.row0 + (uint32_t)(ring * 32) * (SEG ? (uint32_t)a.seg_S : 1u);
    uint2 *queue = s_queue[warp];
    uint32_t *thr = s_thr[warp];
    uint4(*rec)[CH_B][4] = s_rec[warp];

    // hand-over rings: in from the preceding chain of the ring, out to the following one
    const int pred = chain == 0 ? a.nchains - 1 : chain - 1;
    const int succ = chain + 1 == a.nchains ? 0 : chain + 1;
    const uint4 *ll_in = (pred / CH_W == cg) ? &s_ll[pred % CH_W][0][lane]
                                             : a.gll + ((size_t)(ring * a.bpr + pred / CH_W) * CH_D) * 32 + lane;
    uint4 *ll_out = (succ / CH_W == cg) ? &s_ll[warp][0][lane]
                                        : a.gll + ((size_t)(ring * a.bpr + cg) * CH_D) * 32 + lane;

    ChainStage stage = {0u, 0u, 0u};
    // progress of other chains: 2-entry cache of the last values seen
    uint32_t pc_a = 0xFFFFFFFFu, pc_b = 0xFFFFFFFFu, pv_a = 0u, pv_b = 0u;
    auto prog_wait = [&](uint32_t cj, int32_t need) -> bool {
        if (cj == pc_a && (int32_t)(pv_a - (uint32_t)need) >= 0) return true;
        if (cj == pc_b && (int32_t)(pv_b - (uint32_t)need) >= 0) return true;
        const uint32_t v = chain_wait_progress(a, prog + cj, need);
        if (v == 0xFFFFFFFFu) return false;
        if (cj == pc_b) pv_b = v;
        else {
            if (cj != pc_a) {
                pc_b = pc_a;
                pv_b = pv_a;
                pc_a = cj;
            }
            pv_a = v;
        }
        return true;
    };

    uint64_t w = 0ull, w1 = 0ull, w2 = 0ull, wn[4] = {0ull, 0ull, 0ull, 0ull};
    uint64_t result = 0ull;
    bool ll_pending = false;
    int ll_slot = 0;
    uint32_t ll_loc = 0u;
    int32_t ll_need = 0;

    if (T == 0u) return;
    chain_issue_batch(a, stage, rec, cbase, len);
    if (live) w1 = __ldcg(wbase + (size_t)cbase * nrows);          // own word of step 0

    const uint64_t valid = a.valid;
    uint32_t s = 0u, p = 0u, f = 0u, ms = 0u;                      // step t = (sweep s, position p); schedule step f
    // iteration t = -1 only requests the inputs of step 0
    for (int32_t t = -1; t < (int32_t)T; t++) {
        uint32_t sn = s, pn = p, msn = ms, fn = f;
        if (t >= 0) {
            const uint32_t i = (uint32_t)cbase + p;
            const uint32_t sweep = a.sweep0 + s;
            const uint4 *r = rec[((uint32_t)t / CH_B) & 1u][(uint32_t)t % CH_B];
            const uint4 r1 = r[1], r2 = r[2], r3 = r[3];
            const uint32_t pad = r1.y;
            const uint64_t fnames = ((uint64_t)r2.y << 32) | r2.x;
            const uint64_t lane1tab = ((uint64_t)r2.w << 32) | r2.z;

            // ---- the hand-over packet, if it had not arrived when it was first looked for
            if (ll_pending) {
                const uint32_t cj = ll_loc >> 16, pj = ll_loc & 0xFFFFu;
                uint64_t v = 0ull;
                const int rc = chain_wait_packet(a, ll_in + (pj & (CH_D - 1)) * 32, ll_need, &v);
                if (rc == 2) return;
                if (rc == 1) {
                    if (!prog_wait(cj, ll_need)) return;
                    v = 0ull;
                    if (live) v = __ldcg(wbase + (size_t)(cj * (uint32_t)C + pj) * nrows);
                }
#pragma unroll
                for (int n = 0; n < 4; n++)
                    if (n == ll_slot) wn[n] = v;
            }

            uint64_t z[4];
#pragma unroll
            for (int n = 0; n < 4; n++) {
                const uint64_t sg = ((pad >> (8 + n)) & 1u) ? ~0ull : 0ull;
                z[n] = w ^ wn[n] ^ sg;
                asm volatile("" : "+l"(z[n]));
            }

            // ---- thresholds on demand (only steps on which some lane needs a uniform)
            bool thr_ready = false;
            auto ensure_thr = [&]() {
                if (!thr_ready) {
                    chain_build_thr<QA>(thr, __uint_as_float(r1.z), __uint_as_float(r1.w), __uint_as_float(r3.y),
                                        __uint_as_float(r3.z), pad, a.jp2[f], a.invT[f]);
                    thr_ready = true;
                }
            };
            const auto thr_tab = [&](uint32_t c, uint32_t pat) -> uint32_t { return thr[c * 16u + pat]; };
            // truth tables, for the functions the list of 27 does not hold
            auto hacc_of = [&](int c) -> uint32_t { return c < 2 ? (uint32_t)(lane1tab >> (16 * c)) & 0xFFFFu : (r3.x & 0xFFFFu); };
            auto hall_of = [&](int c) -> uint32_t { return c < 2 ? (uint32_t)(lane1tab >> (32 + 16 * c)) & 0xFFFFu : (r3.x >> 16); };
            auto evalf = [&](uint32_t fid, uint32_t h) -> uint64_t { return eval_fn(fid, &h, z); };

            uint64_t C3[3], todo, XL = 0ull, XR = 0ull, flip1 = 0ull;
            if (!QA) {
                todo = live ? valid : 0ull;
                C3[0] = C3[1] = C3[2] = todo;
            } else if (SEG) {
                // several replicas per word: the rules of colour_fast.cu, segment by segment
                const uint64_t ones = a.seg_ones;
                const uint64_t bl = ((w & a.seg_top) >> (a.seg_P - 1)) * ones;
                const uint64_t br_old = ((w & a.seg_l1) >> 1) * ones;
                XL = (w ^ bl) & ~a.seg_top;
                const uint32_t fa0 = (uint32_t)fnames & 0xFFu, fb0 = (uint32_t)(fnames >> 8) & 0xFFu;
                const uint32_t fa1 = (uint32_t)(fnames >> 16) & 0xFFu, fb1 = (uint32_t)(fnames >> 24) & 0xFFu;
                const uint64_t F0 = evalf(fa0, hacc_of(0));
                const uint64_t F1 = (fa1 == fa0 && fa0 != FID_GENERIC) ? F0 : evalf(fa1, hacc_of(1));
                uint64_t N0 = 0ull, N1 = 0ull;
                if (fb0 != FID_NONE) N0 = evalf(fb0, hall_of(0)) & ~F0;
                if (fb1 != FID_NONE) N1 = evalf(fb1, hall_of(1)) & ~F1;
                const uint64_t l1 = live ? a.seg_l1 : 0ull;
                flip1 = ((XL & F1) | (~XL & F0)) & l1;
                const uint64_t need1 = ((XL & N1) | (~XL & N0)) & l1;
                if (__any_sync(FULL, need1 != 0ull)) {
                    ensure_thr();
                    for (int g = 0; g < a.seg_S; g++) {
                        const int k = g * a.seg_P + 1;
                        if ((need1 >> k) & 1ull) {
                            const uint32_t c1 = (uint32_t)(XL >> k) & 1u;
                            if (lane_uniform(1, i, sweep, a.row0 + (uint32_t)(row * a.seg_S + g), a.k0, a.k1) <
                                thr_tab(c1, pattern_at(z, k)))
                                flip1 |= 1ull << k;
                        }
                    }
                }
                const uint64_t br_new = (((w ^ flip1) & a.seg_l1) >> 1) * ones;
                XR = ((w ^ br_new) & ~a.seg_low) | ((w ^ br_old) & a.seg_low);
                todo = live ? (valid & ~a.seg_l1) : 0ull;
                C3[0] = ~(XL | XR) & todo;
                C3[1] = (XL ^ XR) & todo;
                C3[2] = (XL & XR) & todo;
            } else {
                // reference Trotter neighbours: slices P-1 and 1 for every slice; lane 1 is decided first
                const uint64_t bl = (w & a.top) ? ~0ull : 0ull;
                const uint64_t br_old = (w & 2ull) ? ~0ull : 0ull;
                XL = (w ^ bl) & ~a.top;
                const uint32_t c1 = (uint32_t)(XL >> 1) & 1u;
                const uint32_t p1 = pattern_at(z, 1);
                const uint32_t t1 = (uint32_t)(lane1tab >> (16u * c1 + p1));
                const uint32_t t2 = (uint32_t)(lane1tab >> (32u + 16u * c1 + p1));
                if (live && (t1 & 1u)) flip1 = 2ull;
                const bool need1 = live && !(t1 & 1u) && (t2 & 1u);
                if (__any_sync(FULL, need1)) {                       // ~1% of the words
                    ensure_thr();
                    if (need1 && lane_uniform(1, i, sweep, a.row0 + (uint32_t)row, a.k0, a.k1) < thr_tab(c1, p1))
                        flip1 = 2ull;
                }
                const uint64_t br_new = br_old ^ (flip1 ? ~0ull : 0ull);
                XR = ((w ^ br_new) & ~1ull) | ((w ^ br_old) & 1ull);
                todo = live ? (valid & ~2ull) : 0ull;
                C3[0] = ~(XL | XR) & todo;
                C3[1] = (XL ^ XR) & todo;
                C3[2] = (XL & XR) & todo;
            }
            uint64_t ACC = 0ull, NEED = 0ull, V = 0ull;
            uint32_t last = FID_NONE;
#pragma unroll 1
            for (int c = 0; c < NC; c++) {
                const uint64_t Cc = (c == 0) ? C3[0] : (c == 1 ? C3[1] : C3[2]);
                if (!__any_sync(FULL, Cc != 0ull)) continue;
                const uint32_t names = (uint32_t)(fnames >> (16 * c));
                const uint32_t fa = names & 0xFFu, fb = (names >> 8) & 0xFFu;
                if (fa != last || fa == FID_GENERIC) {
                    V = evalf(fa, hacc_of(c));
                    last = fa;
                }
                ACC |= V & Cc;
                if (fb != FID_NONE) NEED |= evalf(fb, hall_of(c)) & ~V & Cc;
            }
            if (__any_sync(FULL, NEED != 0ull)) {
                ensure_thr();
                ACC |= resolve_draws<QA>(NEED, z, XL, XR, thr_tab, queue, i, sweep, prow_warp, a.k0, a.k1,
                                         SEG ? a.seg_P : 64, SEG ? a.seg_S : 1);
            }
            result = w ^ flip1 ^ ACC;

            // ---- publish: the state word, the packet for the following chain, progress every few steps
            if (live) wbase[(size_t)i * nrows] = result;
            const uint32_t tag = s * (uint32_t)C + p + 1u;
            st_vol_v4(ll_out + (p & (CH_D - 1)) * 32, (uint32_t)result, tag, (uint32_t)(result >> 32), tag);
            const bool last_step = p + 1u == (uint32_t)len;
            if (last_step || ((p + 1u) & a.gmask) == 0u) {
                __syncwarp();                                   // the release below covers the stores of all lanes
                if (lane == 0) st_release(prog + chain, last_step ? (s + 1u) * (uint32_t)C : tag);
            }
            pn = p + 1u;
            if (pn == (uint32_t)len) {
                pn = 0u;
                sn++;
                if (++msn == (uint32_t)a.mcsteps) {
                    msn = 0u;
                    fn++;
                }
            }
        }

        // ---- request everything step tn = t + 1 = (sn, pn) reads, one step ahead: the own word of the
        //      step after it (w2) and its four neighbour words (wn) -- the word just written, w2, a
        //      hand-over packet (left pending if it is not there yet), or a state word
        const uint32_t tn = (uint32_t)(t + 1);
        if (tn < T) {
            if ((tn % CH_B) == 0u) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                chain_issue_batch(a, stage, rec, cbase, len);
            }
            w2 = 0ull;
            if (tn + 1u < T) {
                const uint32_t p2 = (pn + 1u == (uint32_t)len) ? 0u : pn + 1u;
                if (live) w2 = __ldcg(wbase + (size_t)((uint32_t)cbase + p2) * nrows);
            }
            const uint4 *r = rec[(tn / CH_B) & 1u][tn % CH_B];
            const uint4 locv = r[0];
            const uint32_t kinds = r[1].x;
            ll_pending = false;
#pragma unroll
            for (int n = 0; n < 4; n++) {
                const uint32_t kind = (kinds >> (8 * n)) & 0xFFu;
                uint64_t v = 0ull;
                if (kind == PIQMC_K_PREV) v = result;
                else if (kind == PIQMC_K_NEXT) v = w2;
                else if (kind >= PIQMC_K_LL_CUR) {
                    const uint32_t loc = n == 0 ? locv.x : (n == 1 ? locv.y : (n == 2 ? locv.z : locv.w));
                    const uint32_t cj = loc >> 16, pj = loc & 0xFFFFu;
                    const int32_t need = (int32_t)(sn * (uint32_t)C + pj + 1u) -
                                         ((kind == PIQMC_K_LL_OLD || kind == PIQMC_K_MEM_OLD) ? C : 0);
                    bool mem = kind >= PIQMC_K_MEM_SELF;
                    if (!mem) {
                        if (need <= 0) mem = true;             // old value in the first sweep: nothing was handed over
                        else {
                            const uint4 pk = ld_vol_v4(ll_in + (pj & (CH_D - 1)) * 32);
                            const int32_t d1 = (int32_t)(pk.y - (uint32_t)need), d2 = (int32_t)(pk.w - (uint32_t)need);
                            if (__all_sync(FULL, d1 == 0 && d2 == 0)) v = ((uint64_t)pk.z << 32) | pk.x;
                            else if (__any_sync(FULL, d1 > 0 || d2 > 0)) mem = true;     // slot reused: producer far ahead
                            else {
                                ll_pending = true;
                                ll_slot = n;
                                ll_loc = loc;
                                ll_need = need;
                            }
                        }
                    }
                    if (mem) {
                        if (kind != PIQMC_K_MEM_SELF && need > 0)
                            if (!prog_wait(cj, need)) return;
                        if (live) v = __ldcg(wbase + (size_t)(cj * (uint32_t)C + pj) * nrows);
                    }
                }
                wn[n] = v;
            }
        }
        w = w1;
        w1 = w2;
        s = sn;
        p = pn;
        ms = msn;
        f = fn;
    }
}

template <typename T>
int grow(T *&p, size_t &have, size_t want, cudaStream_t stream)
{
    if (want <= have && p) return PIQMC_OK;
    PIQMC_CUDA(cudaStreamSynchronize(stream));
    if (p) PIQMC_CUDA(cudaFree(p));
    p = nullptr;
    have = 0;
    PIQMC_CUDA(cudaMalloc(&p, want * sizeof(T)));
    have = want;
    return PIQMC_OK;
}

}  // namespace

int piqmc_check_watchdog(piqmc_ctx *c, const char *what)
{
    if (!c->d_err) return PIQMC_OK;
    unsigned int e = 0;
    PIQMC_CUDA(cudaMemcpyAsync(&e, c->d_err, sizeof(e), cudaMemcpyDeviceToHost, c->stream));
    PIQMC_CUDA(cudaStreamSynchronize(c->stream));
    if (e != 0u) {
        cudaMemsetAsync(c->d_err, 0, sizeof(unsigned int), c->stream);
        piqmc_set_error("%s: a dependency wait timed out on the device (code %u): the state is not valid", what, e);
        return PIQMC_ECUDA;
    }
    return PIQMC_OK;
}

int launch_chain_sweeps(piqmc_ctx *c, int qa, int nsched, int mcsteps, const float *h_jp2, const float *h_invT,
                        uint64_t seed, uint32_t row0, uint32_t sweep0)
{
    if (nsched <= 0 || mcsteps <= 0) return PIQMC_OK;
    PIQMC_REQUIRE(c->chain_C >= 4 && c->d_cstat, PIQMC_EINVAL, "no chain plan for this graph");
    const int N = c->nspins, C = c->chain_C, nchains = c->chain_n;
    const int nrings = (c->nrows + 31) / 32;
    const int bpr = (nchains + CH_W - 1) / CH_W;
    PIQMC_REQUIRE((size_t)nrings * bpr < ((size_t)1 << 31), PIQMC_EINVAL, "too many blocks for one launch");
    // schedule steps per launch: bounded by the memory of the decision tables (512 MB) and by the
    // 31-bit range of the step tags
    size_t max_f = std::max<size_t>(1, ((size_t)512 << 20) / ((size_t)N * sizeof(PiqmcChainDyn)));
    max_f = std::min<size_t>(max_f, std::max<size_t>(1, ((size_t)1 << 30) / ((size_t)C * mcsteps)));
    const size_t nf_max = std::min<size_t>(max_f, nsched);

    size_t err_have = c->d_err ? 1 : 0, tk_have = c->d_ticket ? 1 : 0;
    if (int rc = grow(c->d_err, err_have, 1, c->stream)) return rc;
    if (int rc = grow(c->d_ticket, tk_have, 1, c->stream)) return rc;
    if (int rc = grow(c->d_cdyn, c->cdyn_elems, nf_max * N, c->stream)) return rc;
    if (int rc = grow(c->d_cprog, c->cprog_elems, (size_t)nrings * nchains, c->stream)) return rc;
    {
        char *p = (char *)c->d_cll;
        const size_t want = (size_t)nrings * bpr * CH_D * 32 * sizeof(uint4);
        if (int rc = grow(p, c->cll_bytes, want, c->stream)) return rc;
        c->d_cll = p;
    }
    float *d_par = nullptr;                                   // jp2[nsched] then invT[nsched]
    PIQMC_CUDA(cudaMalloc(&d_par, 2 * (size_t)nsched * sizeof(float)));
    cudaError_t e = cudaMemcpyAsync(d_par, h_jp2, nsched * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_par + nsched, h_invT, nsched * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_err, 0, sizeof(unsigned int), c->stream);

    ChainArgs a;
    a.words = c->d_words;
    a.stat = c->d_cstat;
    a.dyn = c->d_cdyn;
    a.prog = c->d_cprog;
    a.gll = (uint4 *)c->d_cll;
    a.ticket = c->d_ticket;
    a.err = c->d_err;
    a.nspins = N;
    a.nrows = c->nrows;
    a.C = C;
    a.nchains = nchains;
    a.bpr = bpr;
    a.mcsteps = mcsteps;
    a.gmask = 7u;
    if (const char *s = getenv("PIQMC_CHAIN_G")) {            // tuning knob: power of two
        const int v = atoi(s);
        if (v >= 1 && (v & (v - 1)) == 0) a.gmask = (uint32_t)v - 1u;
    }
    a.k0 = (uint32_t)seed;
    a.k1 = (uint32_t)(seed >> 32);
    a.row0 = row0;
    a.watchdog_ns = 20000000000ull;
    if (const char *s = getenv("PIQMC_WATCHDOG_MS")) a.watchdog_ns = (unsigned long long)atoll(s) * 1000000ull;
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
    int force_generic = 0;
    if (const char *s = getenv("PIQMC_FORCE_GENERIC_FN")) force_generic = atoi(s);
    int minb = 4;
    if (const char *s = getenv("PIQMC_CHAIN_MINB")) minb = atoi(s);

    for (size_t f0 = 0; f0 < (size_t)nsched && e == cudaSuccess; f0 += nf_max) {
        const int nf = (int)std::min<size_t>(nf_max, nsched - f0);
        a.jp2 = d_par + f0;
        a.invT = d_par + nsched + f0;
        a.nsweeps = nf * mcsteps;
        a.sweep0 = sweep0 + (uint32_t)(f0 * mcsteps);
        const unsigned tgrid = (unsigned)(((size_t)nf * N + 3) / 4);
        if (qa) chain_tables_kernel<true><<<tgrid, 128, 0, c->stream>>>(c->d_cstat, c->d_cdyn, a.jp2, a.invT, N, nf, force_generic);
        else    chain_tables_kernel<false><<<tgrid, 128, 0, c->stream>>>(c->d_cstat, c->d_cdyn, a.jp2, a.invT, N, nf, force_generic);
        c->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemsetAsync(c->d_ticket, 0, sizeof(unsigned int), c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(c->d_cprog, 0, (size_t)nrings * nchains * sizeof(uint32_t), c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(c->d_cll, 0, (size_t)nrings * bpr * CH_D * 32 * sizeof(uint4), c->stream);
        if (e != cudaSuccess) break;
        const dim3 grid((unsigned)(nrings * bpr)), block(CH_THREADS);
        if (!qa) {
            if (minb == 3) chain_sweep<false, false, 3><<<grid, block, 0, c->stream>>>(a);
            else           chain_sweep<false, false, 4><<<grid, block, 0, c->stream>>>(a);
        } else if (c->seg_S > 1) {
            chain_sweep<true, true, 4><<<grid, block, 0, c->stream>>>(a);
        } else {
            if (minb == 3)      chain_sweep<true, false, 3><<<grid, block, 0, c->stream>>>(a);
            else if (minb == 5) chain_sweep<true, false, 5><<<grid, block, 0, c->stream>>>(a);
            else                chain_sweep<true, false, 4><<<grid, block, 0, c->stream>>>(a);
        }
        c->launches++;
        e = cudaGetLastError();
    }
    cudaError_t e2 = cudaStreamSynchronize(c->stream);        // d_par must outlive the launches
    cudaFree(d_par);
    if (e != cudaSuccess || e2 != cudaSuccess) {
        piqmc_set_error("chain sweep launch failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        return PIQMC_ECUDA;
    }
    return piqmc_check_watchdog(c, "chain sweeps");
}
