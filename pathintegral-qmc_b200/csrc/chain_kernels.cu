// chain_kernels.cu -- the natural-order Metropolis sweep of a 2-D lattice as a pipeline of chains
// (maxnb <= 4; QA with the reference's Trotter neighbours, or SA), one launch for a whole run of sweeps.
//
// The sweep being reproduced is sequential: spins 0..N-1 in order, every spin seeing the new value of
// the neighbours visited before it and the old value of the others (piqmc/qmc.pyx:320-357, the order of
// the reference's per-spin-reset variant).  On an L x L torus its dependency graph is a rigid ring:
// (y, x) needs the new (y-1, x) and (y, x-1), and row 0 of the next sweep needs row L-1 of this one, so
// every one of the L steps per sweep lies on the critical path and the cost of ONE hand-over bounds a
// sweep when there are few replicas per GPU.  colour_fast.cu pays an L2 flag round trip per step.  Here:
//
//   * the order is cut into CHAINS of C consecutive spins (a lattice row); one WARP walks one chain for
//     32 or 64 rows (replicas; RPT = 1 or 2 rows per thread) through all sweeps of the launch; a BLOCK
//     holds a band of consecutive chains, a RING is the set of blocks that cover all chains for one
//     group of rows.
//   * every word a step reads is in shared memory: the own row (old values, also the right-hand
//     neighbour) and the row below (old values) are streamed in three steps ahead (cp.async), the
//     left-hand neighbour is the warp's own previous result, the row above is written by the preceding
//     warp straight into a ring of this warp and announced through a hardware (named) barrier --
//     bar.arrive by the producer, bar.sync by the consumer: a waiting warp issues nothing, which
//     matters because waiting and working warps share schedulers (a polling consumer was measured to
//     starve its own producer); the way back ("this slot has been read") is an mbarrier that is
//     complete long before the producer looks at it.  The four neighbour words are then fetched in
//     the order of the spin's sorted couplings by four shared-memory loads whose ring is a byte of
//     the step's record -- no selects, no branches.
//   * between blocks (band boundaries, and from the last chain back to the first for the next sweep)
//     results travel as self-validating 16-byte packets {lo, tag, hi, tag} through L2: the producer
//     needs no fence; slot reuse and the row-below guard use a progress word the first chain of every
//     band publishes every few steps (st.release / ld.acquire).
//   * the decision functions of a (schedule step, spin) -- by name, see colour_device.cuh -- are
//     computed ONCE for all replicas by chain_tables_kernel (16-byte records, staged 32 steps at a
//     time); acceptance thresholds are rebuilt by the warp only on steps where a lane needs a uniform.
//
// Blocks take (ring, band) tickets in ring-major order; a ring needs all its bands resident (the host
// checks that one ring fits the device), earlier rings are resident or done, so the scheme cannot
// deadlock whatever the dispatch order.  Every wait has a watchdog (err word + PIQMC_ECUDA, not a hang).
// Semantics: oracle_qa_colour / oracle_sa_colour (oracle/piqmc_oracle.c part 3), bit for bit.
#include <stdio.h>
#include <stdlib.h>

#include "colour_device.cuh"

namespace {

constexpr int LT_MAXW = 16;              // chains (warps) per block, at most
constexpr int LT_DR = 4;                 // ring depth in slots (results, own row, row below)
constexpr int LT_DG = 16;                // packets per lane in a hand-over ring between blocks
constexpr int LT_REC = 32;               // steps per staged batch of records
constexpr int LT_PRE = 3;                // own row / row below are requested this many steps ahead
constexpr int LT_QCAP = QCAP;
constexpr unsigned FULL = 0xffffffffu;

struct LatArgs {
    uint64_t *words;                     // [N + 1][nrows]
    const PiqmcChainStat *stat;          // [N]
    const uint4 *hot;                    // [schedule steps of this launch][N]   PiqmcChainHot
    const uint4 *cold;                   // [schedule steps of this launch][N]   PiqmcChainCold
    const float *jp2, *invT;             // per schedule step of this launch
    const int *band_first;               // [nbands + 1] first chain of every band
    uint32_t *prog;                      // [nrings][nbands] steps completed by the first chain of the band
    uint4 *gll;                          // [nrings][nbands][LT_DG][32 * RPT] packets written by the last chain of the band
    unsigned int *ticket, *err;
    unsigned long long *dbg;             // profiling (PIQMC_CHAIN_PROF): cycles per segment of a step, [block][warp][16]
    int nspins, nrows, C, nchains, nbands, nsweeps, mcsteps;
    int wrap;                            // the last chain is coupled to the first one (torus)
    uint32_t gmask;                      // progress is published when ((t + 1) & gmask) == 0 and at the end
    uint32_t k0, k1, row0, sweep0;
    unsigned long long watchdog_ns;
    uint64_t valid;
    int seg_P, seg_S;                    // SEG: seg_S replicas of seg_P slices per word; else seg_P = lanes, seg_S = 1
    uint64_t seg_low, seg_l1, seg_top, seg_ones;
};

// Shared memory of a block: per warp a REGION of five rings of LT_DR slots (a slot = 32 lanes x RPT words),
// indexed by where a neighbour word comes from (PIQMC_K_*): [0] zeros, [1] the warp's own results,
// [2] own row, [3] results of the preceding chain (written by that warp, or by this warp from the packets
// of the preceding block), [4] row below; behind the regions one WarpMisc per warp.  The rings are
// phased so that step t reads slot t & 3 of every ring: a result of step t sits in slot (t + 1) & 3 of
// ring 1 and slot t & 3 of the consumer's ring 3, the own-row word of step t in slot (t + 3) & 3 of ring 2.
struct WarpMisc {
    uint4 rec[2 * LT_REC];               // staged hot records, two batches back to back (index t & 63)
    uint32_t thr[48];                    // acceptance thresholds (rare path)
    uint32_t thr_step, pad_[3];          // step + 1 the thresholds were built for (0: none)
    uint2 queue[LT_QCAP];                // pooled draw requests (rare path)
    uint64_t empty[2 * LT_DR];           // [0]: mbarrier "ring 3 has been read" (phase k: the row above of step k)
};
static_assert(sizeof(WarpMisc) % 16 == 0, "WarpMisc must keep 16-byte alignment");

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// one try: the hardware suspends the thread until the phase completes or its (short) time limit runs out
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void cp_async16(uint32_t smem, const void *gmem, uint32_t srcsize)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem), "l"(gmem), "r"(srcsize) : "memory");
}
__device__ __forceinline__ void cp_async16_ca(uint32_t smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint4 ld_vol_v4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_vol_v4(uint4 *p, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
template <int RPT>
__device__ __forceinline__ void lds_words(uint32_t addr, uint64_t (&v)[RPT])
{
    if (RPT == 2) asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v[0]), "=l"(v[RPT - 1]) : "r"(addr));
    else asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v[0]) : "r"(addr));
}
template <int RPT>
__device__ __forceinline__ void sts_words(uint32_t addr, const uint64_t (&v)[RPT])
{
    if (RPT == 2) asm volatile("st.shared.v2.u64 [%0], {%1, %2};" ::"r"(addr), "l"(v[0]), "l"(v[RPT - 1]) : "memory");
    else asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v[0]) : "memory");
}

// A wait has gone through `polls` unsuccessful tries.  Returns true when it must be given up: another
// warp has already reported a failure, or this wait has lasted longer than the watchdog allows.
__device__ __noinline__ bool wait_expired(const LatArgs &a, unsigned long long t0, unsigned code)
{
    if (*(volatile unsigned int *)a.err != 0u) return true;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (now - t0 > a.watchdog_ns) {
        atomicCAS(a.err, 0u, code);
        return true;
    }
    return false;
}
__device__ __forceinline__ unsigned long long timer_now()
{
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    return now;
}

// mbarrier wait with the watchdog; false: give up.  The polling loop is kept to a handful of
// instructions: waiting warps share their scheduler with working ones.
__device__ __forceinline__ bool mbar_wait(const LatArgs &a, uint32_t bar, uint32_t parity, unsigned code)
{
    if (mbar_try(bar, parity)) return true;
    unsigned long long t0 = 0ull;
    while (true) {
#pragma unroll 1
        for (int k = 0; k < 1024; k++)
            if (mbar_try(bar, parity)) return true;
        if (t0 == 0ull) t0 = timer_now();
        else if (wait_expired(a, t0, code)) return false;
    }
}

// wait until *flag >= need; the value seen, or 0xFFFFFFFF when the watchdog gave up
__device__ __noinline__ uint32_t wait_progress(const LatArgs &a, const uint32_t *flag, int32_t need)
{
    unsigned polls = 0;
    unsigned long long t0 = 0ull;
    uint32_t v;
    while ((int32_t)((v = ld_acquire(flag)) - (uint32_t)need) < 0) {
        if ((++polls & 63u) == 0u) {
            if (t0 == 0ull) t0 = timer_now();
            else if (wait_expired(a, t0, 2u)) return 0xFFFFFFFFu;
        }
    }
    return v;
}

__device__ __forceinline__ WarpMisc *warp_misc(int cw, int slot_bytes, int warp)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    return reinterpret_cast<WarpMisc *>(smem + (size_t)cw * 5 * LT_DR * slot_bytes) + warp;
}

// acceptance thresholds of the 16 z-patterns in every Trotter class, by the warp: the sequence of
// float32 operations of build_table_warp (colour_device.cuh)
template <bool QA>
__device__ __forceinline__ void chain_build_thr(WarpMisc *m, uint32_t step1, const PiqmcChainStat *st, float jp2, float invT)
{
    if (m->thr_step == step1) return;               // built earlier in this step (warp-uniform)
    uint32_t *thr = m->thr;
    const int lane = threadIdx.x & 31;
    const uint32_t p = lane & 15;
    const float Jz[4] = {st->J[0], st->J[1], st->J[2], st->J[3]};
    const uint32_t pad = st->pad;
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
    if (lane == 0) m->thr_step = step1;
    __syncwarp();
}

// ---- the rare paths of a step, out of line: the sweep loop stays small -------------------------------
// Resolve the lanes of NEED (one word per thread: row `prow0 + lane * lstride`, replica segments as in
// resolve_draws); the accepted ones come back.
template <bool QA>
__device__ __noinline__ uint64_t rare_draws(const LatArgs &a, uint64_t NEED, uint64_t z0, uint64_t z1, uint64_t z2, uint64_t z3,
                                            uint64_t XL, uint64_t XR, uint32_t i, uint32_t f, uint32_t sweep,
                                            uint32_t prow0, int lstride, uint32_t step1)
{
    WarpMisc *m = warp_misc((int)(blockDim.x >> 5), 256 * lstride, (int)(threadIdx.x >> 5));
    uint32_t *thr = m->thr;
    chain_build_thr<QA>(m, step1, a.stat + i, a.jp2[f], a.invT[f]);
    const auto thr_tab = [&](uint32_t c, uint32_t pat) -> uint32_t { return thr[c * 16u + pat]; };
    const uint64_t z[4] = {z0, z1, z2, z3};
    return resolve_draws<QA>(NEED, z, XL, XR, thr_tab, m->queue, i, sweep, prow0, a.k0, a.k1, QA ? a.seg_P : 64,
                             QA ? a.seg_S : 1, lstride);
}
// Slice 1 of every replica segment of one word per thread: the lanes of need1 draw their uniform.
__device__ __noinline__ uint64_t rare_slice1(const LatArgs &a, uint64_t need1, uint64_t z0, uint64_t z1, uint64_t z2, uint64_t z3,
                                             uint64_t XL, uint32_t i, uint32_t f, uint32_t sweep, uint32_t prow, int lstride,
                                             uint32_t step1)
{
    WarpMisc *m = warp_misc((int)(blockDim.x >> 5), 256 * lstride, (int)(threadIdx.x >> 5));
    uint32_t *thr = m->thr;
    chain_build_thr<true>(m, step1, a.stat + i, a.jp2[f], a.invT[f]);
    const uint64_t z[4] = {z0, z1, z2, z3};
    uint64_t flip = 0ull;
    for (int g = 0; g < a.seg_S; g++) {                       // warp-uniform trip count
        const int k = g * a.seg_P + 1;
        if ((need1 >> k) & 1ull) {
            const uint32_t c1 = (uint32_t)(XL >> k) & 1u;
            if (lane_uniform(1, i, sweep, prow + (uint32_t)g, a.k0, a.k1) < thr[c1 * 16u + pattern_at(z, k)])
                flip |= 1ull << k;
        }
    }
    return flip;
}
// a function outside the list of 27: from its truth table (class c, "accept by sign" or "... or needs a uniform")
__device__ __noinline__ uint64_t rare_generic(const LatArgs &a, uint32_t f, uint32_t i, int c, int need, uint64_t z0,
                                              uint64_t z1, uint64_t z2, uint64_t z3)
{
    const uint4 cold = __ldg(a.cold + (size_t)f * a.nspins + i);
    const uint32_t packed[3] = {cold.x, cold.y, cold.z};      // hacc0 hacc1 | hacc2 hall0 | hall1 hall2
    const int idx = need ? 3 + c : c;
    const uint32_t h = (packed[idx >> 1] >> (16 * (idx & 1))) & 0xFFFFu;
    return eval_generic(h, z0, z1, z2, z3);
}

// ---- decision functions of every (schedule step, spin), once for all replicas ----------------------
// hot record: x = names fa0 fb0 fa1 fb1 (fa: "accept by sign", fb: "... or needs a uniform" of a Trotter
// class), y = fa2 | fb2 << 8 | flags << 16 (bit 0: some fb is not FID_NONE, bit 1: some name is
// FID_GENERIC), z = byte k 0xFF when sorted coupling k is negative, w = byte k the kind of sorted slot k
template <bool QA>
__global__ void __launch_bounds__(128) chain_tables_kernel(const PiqmcChainStat *stat, uint4 *hot, uint4 *cold,
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
    const float Jz[4] = {st.J[0], st.J[1], st.J[2], st.J[3]};
    SpinTable &tab = tabs[warp];
    if (lane < 8) tab.names[lane] = (uint8_t)FID_NONE;
    if (lane < 3) tab.hacc[lane] = tab.hall[lane] = 0u;
    __syncwarp();
    for (int c = 0; c < NC; c++) build_table_warp<QA>(tab, c, Jz, st.pad, jp2[f], invT[f], force_generic != 0);
    __syncwarp();
    if (lane == 0) {
        uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
        bool anyneed = false, anygen = false;
        for (int k = 0; k < 4; k++) w0 |= (uint32_t)tab.names[k] << (8 * k);
        w1 = (uint32_t)tab.names[4] | ((uint32_t)tab.names[5] << 8);
        for (int c = 0; c < NC; c++) {
            anyneed = anyneed || tab.names[2 * c + 1] != FID_NONE;
            anygen = anygen || tab.names[2 * c] == FID_GENERIC || tab.names[2 * c + 1] == FID_GENERIC;
        }
        if (anyneed) w1 |= 1u << 16;
        if (anygen) w1 |= 2u << 16;
        for (int k = 0; k < 4; k++) {
            if ((st.pad >> (8 + k)) & 1u) w2 |= 0xFFu << (8 * k);
            w3 |= ((st.kinds >> (8 * k)) & 0xFFu) << (8 * k);
        }
        hot[g] = make_uint4(w0, w1, w2, w3);
        cold[g] = make_uint4(tab.hacc[0] | (tab.hacc[1] << 16), tab.hacc[2] | (tab.hall[0] << 16),
                             tab.hall[1] | (tab.hall[2] << 16), 0u);
    }
}

// ---- the sweep kernel -------------------------------------------------------------------------------
template <bool QA, bool SEG, int RPT>
__global__ void __launch_bounds__(LT_MAXW * 32, 1) chain_sweep(const LatArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned int s_ticket;
    constexpr uint32_t SLOT = 256 * RPT;           // bytes of one ring slot: 32 lanes x RPT words
    constexpr uint32_t RING = LT_DR * SLOT, REGION = 5 * RING;
    constexpr int RSH = RPT == 2 ? 11 : 10;        // log2(RING)
    constexpr int RW = 32 * RPT;                   // rows per ring of blocks

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int CW = (int)(blockDim.x >> 5);
    WarpMisc *misc_all = warp_misc(CW, SLOT, 0);
    WarpMisc &m = misc_all[warp];
    if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1u);
    {   // ring 0 of this warp's region: zeros
        uint4 *z0 = reinterpret_cast<uint4 *>(smem + (size_t)warp * REGION);
        for (int k = lane; k < (int)(RING / 16); k += 32) z0[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (lane == 0) mbar_init(&m.empty[0], 1u);
    if (lane == 0) m.thr_step = 0u;
    __syncthreads();                               // the only block barrier: warps are independent from here on
    const unsigned int tk = s_ticket;
    const int ring = (int)(tk / (unsigned int)a.nbands);
    const int band = (int)(tk - (unsigned int)ring * (unsigned int)a.nbands);
    const int cfirst = a.band_first[band];
    const int nW = a.band_first[band + 1] - cfirst;
    if (warp >= nW) return;
    const int chain = cfirst + warp;

    const int C = a.C, nrows = a.nrows;
    const size_t cstride = (size_t)C * nrows;      // words between the same position of consecutive chains
    const int cbase = chain * C;
    const uint32_t T = (uint32_t)a.nsweeps * (uint32_t)C;
    if (T == 0u) return;
    const int row = ring * RW + lane * RPT;        // first of this thread's RPT rows
    bool live[RPT];
#pragma unroll
    for (int r = 0; r < RPT; r++) live[r] = row + r < nrows;
    const bool succ_in = warp + 1 < nW;                                    // consumer of my results in this block
    const bool ext_out = !succ_in && (band + 1 < a.nbands || a.wrap);     // ... or in the next block of the ring
    const bool ext_in = warp == 0 && (band > 0 || a.wrap);                // my row above comes from another block
    const bool has_down = chain + 1 < a.nchains || a.wrap;
    const int nband = band + 1 < a.nbands ? band + 1 : 0;                 // band that consumes my packets
    const int pband = band > 0 ? band - 1 : a.nbands - 1;                 // band that produces my row above
    const uint32_t in_off = band == 0 ? (uint32_t)C : 0u;                 // chain 0 reads the last chain one sweep back
    const uint32_t out_off = nband == 0 ? (uint32_t)C : 0u;
    const uint32_t *prog_next = a.prog + (size_t)ring * a.nbands + nband;
    uint32_t *prog_mine = a.prog + (size_t)ring * a.nbands + band;
    const uint4 *gll_in = a.gll + ((size_t)(ring * a.nbands + pband) * LT_DG) * RW + lane * RPT;
    uint4 *gll_out = a.gll + ((size_t)(ring * a.nbands + band) * LT_DG) * RW + lane * RPT;

    // shared-memory addresses (32-bit, this lane's bytes within a slot included)
    const uint32_t region = smem_u32(smem) + (uint32_t)warp * REGION + (uint32_t)lane * 8u * RPT;
    const uint32_t ring_res = region + RING, ring_own = region + 2u * RING, ring_up = region + 3u * RING,
                   ring_down = region + 4u * RING;
    const uint32_t succ_up = ring_up + REGION;                             // ring 3 of the following warp
    const uint32_t recbase = smem_u32(&m.rec[0]);
    const uint32_t bar_empty = smem_u32(&m.empty[0]);
    const uint32_t succ_empty = bar_empty + (uint32_t)sizeof(WarpMisc);
    bool dead = false;                             // a wait timed out: no more waiting, the state is invalid (the
                                                   // warp keeps its place in the hardware-barrier hand-shakes)

    // global words of this thread's rows: the spin fetched next (LT_PRE steps ahead) and the spin stored next
    const char *fetch = reinterpret_cast<const char *>(a.words + (size_t)cbase * nrows + (live[0] ? row : 0));
    ptrdiff_t down_delta = 8 * (chain + 1 < a.nchains ? (ptrdiff_t)cstride : -(ptrdiff_t)((size_t)(a.nchains - 1) * cstride));
    char *store = reinterpret_cast<char *>(a.words + (size_t)cbase * nrows + (live[0] ? row : 0));
    ptrdiff_t step_bytes = 8 * (ptrdiff_t)nrows, wrap_bytes = 8 * (ptrdiff_t)cstride;
    asm volatile("" : "+l"(down_delta), "+l"(step_bytes), "+l"(wrap_bytes));   // keep them in registers (no re-derivation per step)
    uint32_t fpos = 0u;                                                    // position of `fetch` in the chain
    uint32_t pv_next = 0u;                                                 // last progress seen of the consuming band

    // RPT == 1: plain loads one step in flight (8-byte cp.async cannot bypass L1)
    uint64_t f_own = 0ull, f_down = 0ull;
    uint32_t f_tf = 0u;
    bool f_valid = false;

    auto issue_fetch = [&](uint32_t tf) {          // own row and row below of step tf
        if (RPT == 2) {
            const uint32_t sz = live[0] ? 16u : 0u;
            cp_async16(ring_own + ((tf + 3u) & 3u) * SLOT, fetch, sz);
            if (has_down) cp_async16(ring_down + (tf & 3u) * SLOT, fetch + down_delta, sz);
        } else {
            f_own = live[0] ? __ldcg(reinterpret_cast<const uint64_t *>(fetch)) : 0ull;
            f_down = (has_down && live[0]) ? __ldcg(reinterpret_cast<const uint64_t *>(fetch + down_delta)) : 0ull;
            f_tf = tf;
            f_valid = true;
        }
        fetch += step_bytes;
        if (++fpos == (uint32_t)C) {
            fpos = 0u;
            fetch -= wrap_bytes;
        }
    };
    auto land_fetch = [&]() {                      // RPT == 1: the words requested one step ago go to their slots
        if (RPT == 1 && f_valid) {
            asm volatile("st.shared.u64 [%0], %1;" ::"r"(ring_own + ((f_tf + 3u) & 3u) * SLOT), "l"(f_own) : "memory");
            asm volatile("st.shared.u64 [%0], %1;" ::"r"(ring_down + (f_tf & 3u) * SLOT), "l"(f_down) : "memory");
            f_valid = false;
        }
    };
    auto issue_records = [&](uint32_t tb) {        // hot records of steps tb .. tb + LT_REC - 1, one per lane
        const uint32_t tl = tb + (uint32_t)lane;
        if (tl < T) {
            const uint32_t sl = tl / (uint32_t)C, pl = tl - sl * (uint32_t)C;
            const uint32_t fl = sl / (uint32_t)a.mcsteps;
            cp_async16_ca(recbase + ((tl & (2u * LT_REC - 1u)) << 4), a.hot + (size_t)fl * a.nspins + (uint32_t)cbase + pl);
        }
    };

    // ---- prologue: the value the first step's left-hand neighbour may need (the chain's last spin, for
    //      a chain that closes on itself), records, and the first LT_PRE steps' words
    {
        const uint64_t *lastw = a.words + ((size_t)cbase + (size_t)(C - 1)) * nrows + (live[0] ? row : 0);
        uint64_t v[RPT];
#pragma unroll
        for (int r = 0; r < RPT; r++) v[r] = live[r] ? __ldcg(lastw + r) : 0ull;
        sts_words<RPT>(ring_res, v);               // "result of step -1": slot (-1 + 1) & 3
        if (ext_out && has_down && nband == 0) {   // last chain of a torus: its row below is chain 0, THIS sweep
            const int32_t need = (int32_t)(T < (uint32_t)LT_PRE ? T : (uint32_t)LT_PRE);
            pv_next = wait_progress(a, prog_next, need);
            if (pv_next == 0xFFFFFFFFu) dead = true;
        }
        issue_records(0u);
        cp_commit();
        for (uint32_t tf = 0; tf < (uint32_t)LT_PRE; tf++) {
            land_fetch();
            if (tf <= T) issue_fetch(tf);
            cp_commit();
        }
        if (LT_REC < T) issue_records(LT_REC);     // joins the group of step 0 below
    }

    uint4 pk[RPT];                                 // packets of the row above, requested one step ahead
#pragma unroll
    for (int r = 0; r < RPT; r++) pk[r] = make_uint4(0u, 0u, 0u, 0u);
    const uint64_t valid = a.valid;
    uint32_t s = 0u, p = 0u, f = 0u, ms = 0u;      // step t = (sweep s, position p); schedule step f
#ifdef LT_PROFILE                                  // build with -DLT_PROFILE, run with PIQMC_CHAIN_PROF=1: cycles per segment
    const bool prof = a.dbg != nullptr;
    long long pt = 0, pacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define LT_MARK(k)                                  \
    if (prof) {                                     \
        const long long now_ = clock64();           \
        pacc[k] += now_ - pt;                       \
        pt = now_;                                  \
    }
    if (prof) pt = clock64();
#else
#define LT_MARK(k)
#endif
    for (uint32_t t = 0; t < T; t++) {
        const uint32_t tm = t & 3u;
        const uint32_t i = (uint32_t)cbase + p;
        const uint32_t sweep = a.sweep0 + s;

        // ---- (A) last chain of a block: the packet slot this step's result goes to must have been read, and
        //      the consuming block must be far enough for the row below that is requested now.  (Inside a
        //      block both follow from the hand-shake of the previous step, see (F).)
        if (ext_out && !dead) {
            int32_t need = (int32_t)(t + out_off + 1u) - LT_DG;                       // packet slot t % LT_DG is free
            if (has_down) {
                const int32_t nd = nband == 0 ? (int32_t)(t + LT_PRE + 1u) : (int32_t)(t + LT_PRE + 1u) - C;
                if (nd > need) need = nd;
            }
            if (need > (int32_t)T) need = (int32_t)T;
            if (need > 0 && (int32_t)(pv_next - (uint32_t)need) < 0) {
                pv_next = wait_progress(a, prog_next, need);
                if (pv_next == 0xFFFFFFFFu) dead = true;
            }
        }
        LT_MARK(0)
        // ---- (B) request the words of step t + LT_PRE (and, every LT_REC steps, the records after the next batch)
        land_fetch();
        if (t + LT_PRE <= T) issue_fetch(t + LT_PRE);
        if ((t & (LT_REC - 1)) == 0u && t != 0u && t + LT_REC < T) issue_records(t + LT_REC);
        cp_commit();
        cp_wait<LT_PRE - 1>();                     // the groups of steps <= t + 1 have landed
        if ((t & (LT_REC - 1)) == 0u) __syncwarp();                                   // records are read across lanes

        LT_MARK(1)
        // ---- (C) the record: function names, signs, where the neighbour words are
        uint4 rc;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rc.x), "=r"(rc.y), "=r"(rc.z), "=r"(rc.w)
                     : "r"(recbase + ((t & (2u * LT_REC - 1u)) << 4)));
        const uint32_t slot_t = region + tm * SLOT;
        uint32_t src[4], sgn[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            src[k] = slot_t + (__byte_perm(rc.w, 0u, 0x4440u + k) << RSH);           // ring of kind k, slot t & 3
            sgn[k] = __byte_perm(rc.z, 0u, 0x1111u * k);                            // byte k, four times
        }

        LT_MARK(2)
        // ---- (D) the row above: handed over inside the block, or packets from the preceding block
        if (warp > 0) {
            asm volatile("bar.sync %0, 64;" ::"r"(warp) : "memory");                  // barrier `warp`: my producer and me
        } else if (ext_in) {
            uint64_t up[RPT];
            if (t < in_off) {                      // first sweep of chain 0: the last chain's initial words
                const uint64_t *q = a.words + ((size_t)(a.nchains - 1) * C + p) * nrows + (live[0] ? row : 0);
#pragma unroll
                for (int r = 0; r < RPT; r++) up[r] = live[r] ? __ldcg(q + r) : 0ull;
            } else {
                // the packets of this step were requested at the end of the previous one (the preceding
                // block runs ahead by up to LT_DG steps: normally they are there); poll only if not
                const uint32_t tp = t - in_off, need = tp + 1u;
                const uint4 *slot = gll_in + (size_t)(tp & (LT_DG - 1)) * RW;
                unsigned polls = 0;
                unsigned long long t0 = 0ull;
                while (true) {
                    bool ok = true;
#pragma unroll
                    for (int r = 0; r < RPT; r++) {
                        ok = ok && pk[r].y == need && pk[r].w == need;
                        up[r] = ((uint64_t)pk[r].z << 32) | pk[r].x;
                    }
                    if (__all_sync(FULL, ok) || dead) break;
#pragma unroll
                    for (int r = 0; r < RPT; r++) pk[r] = ld_vol_v4(slot + r);
                    if ((++polls & 63u) == 0u) {
                        if (t0 == 0ull) t0 = timer_now();
                        else if (wait_expired(a, t0, 3u)) dead = true;
                    }
                }
            }
            sts_words<RPT>(ring_up + tm * SLOT, up);
        }

        LT_MARK(3)
        // ---- (E) the five words of this step
        uint64_t w[RPT], z[RPT][4];
        lds_words<RPT>(ring_own + ((tm + 3u) & 3u) * SLOT, w);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint64_t v[RPT];
            lds_words<RPT>(src[k], v);
#pragma unroll
            for (int r = 0; r < RPT; r++) z[r][k] = v[r] ^ w[r] ^ (((uint64_t)sgn[k] << 32) | sgn[k]);
        }
        if (warp > 0) {                            // ring 3 may be overwritten: every lane has its words
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty);
        }

        LT_MARK(4)
        const uint32_t fa0 = rc.x & 0xFFu, fb0 = (rc.x >> 8) & 0xFFu;
        const bool anyneed = (rc.y >> 16) & 1u, anygen = (rc.y >> 17) & 1u;
        // function fid of class c (need: "... or needs a uniform"); the list of 27 inline, anything else out of line
        auto evalf = [&](uint32_t fid, int c, int need, uint64_t (&out)[RPT]) {
            if (!anygen || fid < PIQMC_NCANON) eval_canon<RPT>(fid, z, out);
            else {
#pragma unroll
                for (int r = 0; r < RPT; r++) out[r] = rare_generic(a, f, i, c, need, z[r][0], z[r][1], z[r][2], z[r][3]);
            }
        };

        uint64_t result[RPT];
        if (!QA) {
            // ---- SA: one class, all lanes
            uint64_t V[RPT];
            evalf(fa0, 0, 0, V);
            if (fb0 != FID_NONE) {
                uint64_t NEED[RPT];
                evalf(fb0, 0, 1, NEED);
                bool anyn = false;
#pragma unroll
                for (int r = 0; r < RPT; r++) {
                    NEED[r] &= ~V[r] & valid;
                    anyn = anyn || NEED[r] != 0ull;
                }
                if (__any_sync(FULL, anyn)) {
#pragma unroll
                    for (int r = 0; r < RPT; r++)
                        if (__any_sync(FULL, NEED[r] != 0ull))
                            V[r] |= rare_draws<false>(a, NEED[r], z[r][0], z[r][1], z[r][2], z[r][3], 0ull, 0ull, i, f, sweep,
                                                      a.row0 + (uint32_t)(ring * RW + r), RPT, t + 1u);
                }
            }
#pragma unroll
            for (int r = 0; r < RPT; r++) result[r] = w[r] ^ (V[r] & valid);
        } else {
            // ---- QA, the reference's Trotter neighbours: slices P-1 (old value for everyone; itself for
            //      slice P-1) and 1 (old for slice 0, itself for slice 1, new for slices >= 2), per segment
            //      when a word holds several replicas.  Slice 1 is decided first, from the functions of
            //      classes 0 and 1 (its right-hand neighbour is itself).  Class of a lane = number of
            //      Trotter neighbours it disagrees with: XL + XR.
            const uint32_t fa1 = (rc.x >> 16) & 0xFFu, fb1 = rc.x >> 24;
            const uint32_t fa2 = rc.y & 0xFFu, fb2 = (rc.y >> 8) & 0xFFu;
            uint64_t V0[RPT], V1[RPT], V2[RPT], N0[RPT], N1[RPT], N2[RPT], XL[RPT], XR[RPT], flip1[RPT], ACC[RPT];
            const uint64_t todo = valid & ~a.seg_l1;
            evalf(fa0, 0, 0, V0);
#pragma unroll
            for (int r = 0; r < RPT; r++) N0[r] = N1[r] = N2[r] = 0ull;
            if (SEG) {
                // ---- several replicas per word: all functions up front, 64-bit segment arithmetic
                if (fa1 == fa0 && fa0 != FID_GENERIC) {
#pragma unroll
                    for (int r = 0; r < RPT; r++) V1[r] = V0[r];
                } else evalf(fa1, 1, 0, V1);
                if (anyneed) {
                    if (fb0 != FID_NONE) evalf(fb0, 0, 1, N0);
                    if (fb1 != FID_NONE) evalf(fb1, 1, 1, N1);
#pragma unroll
                    for (int r = 0; r < RPT; r++) {
                        N0[r] &= ~V0[r];
                        N1[r] &= ~V1[r];
                    }
                }
                LT_MARK(7)
                uint64_t brold[RPT];
#pragma unroll
                for (int r = 0; r < RPT; r++) {
                    const uint64_t bl = ((w[r] & a.seg_top) >> (a.seg_P - 1)) * a.seg_ones;
                    brold[r] = ((w[r] & a.seg_l1) >> 1) * a.seg_ones;
                    XL[r] = (w[r] ^ bl) & ~a.seg_top;
                    flip1[r] = ((XL[r] & V1[r]) | (~XL[r] & V0[r])) & a.seg_l1;
                }
                if (anyneed) {
                    bool any1 = false;
                    uint64_t need1[RPT];
#pragma unroll
                    for (int r = 0; r < RPT; r++) {
                        need1[r] = ((XL[r] & N1[r]) | (~XL[r] & N0[r])) & a.seg_l1;
                        any1 = any1 || need1[r] != 0ull;
                    }
                    if (__any_sync(FULL, any1)) {                                 // ~1% of the words
#pragma unroll
                        for (int r = 0; r < RPT; r++)
                            flip1[r] |= rare_slice1(a, need1[r], z[r][0], z[r][1], z[r][2], z[r][3], XL[r], i, f, sweep,
                                                    a.row0 + (uint32_t)((row + r) * a.seg_S), RPT, t + 1u);
                    }
                }
#pragma unroll
                for (int r = 0; r < RPT; r++) {
                    const uint64_t brnew = (((w[r] ^ flip1[r]) & a.seg_l1) >> 1) * a.seg_ones;
                    XR[r] = ((w[r] ^ brnew) & ~a.seg_low) | ((w[r] ^ brold[r]) & a.seg_low);   // slice 0 sees the old slice 1
                }
                LT_MARK(8)
                // class 2 (both Trotter neighbours disagree): late in the anneal no lane of the warp is in it
                if (fa2 == fa1 && fa1 != FID_GENERIC) {
#pragma unroll
                    for (int r = 0; r < RPT; r++) V2[r] = V1[r];
                } else {
                    bool any2 = false;
#pragma unroll
                    for (int r = 0; r < RPT; r++) {
                        V2[r] = V1[r];
                        any2 = any2 || (XL[r] & XR[r] & todo) != 0ull;
                    }
                    if (__any_sync(FULL, any2)) evalf(fa2, 2, 0, V2);
                }
                if (anyneed && fb2 != FID_NONE) {
                    evalf(fb2, 2, 1, N2);
#pragma unroll
                    for (int r = 0; r < RPT; r++) N2[r] &= ~V2[r];
                }
            } else {
                // ---- one replica per word, on 32-bit halves (masks of "slice P-1 set" / "slice 1 set" by sign
                //      extension), and every function only when a lane of the warp is in its class: early in the
                //      anneal J_perp is small and the classes share one function, late in the anneal the slices
                //      of a replica agree and only class 0 is populated -- one evaluation per step in both regimes
                const int tsh = a.seg_P - 1;
                uint32_t brold[RPT];
                bool s1 = false;
#pragma unroll
                for (int r = 0; r < RPT; r++) {
                    const uint32_t wlo = (uint32_t)w[r];
                    const uint32_t bl = (uint32_t)((int32_t)((uint32_t)(w[r] >> tsh) << 31) >> 31);
                    brold[r] = (uint32_t)((int32_t)(wlo << 30) >> 31);
                    XL[r] = (w[r] ^ (((uint64_t)bl << 32) | bl)) & ~a.seg_top;
                    s1 = s1 || ((uint32_t)XL[r] & 2u) != 0u;
                }
                const bool same01 = fa1 == fa0 && fa0 != FID_GENERIC;
                bool haveV1 = same01, haveN1 = false;
#pragma unroll
                for (int r = 0; r < RPT; r++) V1[r] = V0[r];                      // stands in until class 1 is needed
                const bool s1c1 = __any_sync(FULL, s1);                           // slice 1 of some word is in class 1
                if (!haveV1 && s1c1) {
                    evalf(fa1, 1, 0, V1);
                    haveV1 = true;
                }
#pragma unroll
                for (int r = 0; r < RPT; r++)
                    flip1[r] = (((uint32_t)XL[r] & (uint32_t)V1[r]) | (~(uint32_t)XL[r] & (uint32_t)V0[r])) & 2u;
                LT_MARK(7)
                if (anyneed) {
                    if (fb0 != FID_NONE) {
                        evalf(fb0, 0, 1, N0);
#pragma unroll
                        for (int r = 0; r < RPT; r++) N0[r] &= ~V0[r];
                    }
                    if (s1c1) {
                        if (fb1 != FID_NONE) {
                            evalf(fb1, 1, 1, N1);
#pragma unroll
                            for (int r = 0; r < RPT; r++) N1[r] &= ~V1[r];
                        }
                        haveN1 = true;
                    }
                    bool any1 = false;
                    uint32_t need1[RPT];
#pragma unroll
                    for (int r = 0; r < RPT; r++) {
                        need1[r] = (((uint32_t)XL[r] & (uint32_t)N1[r]) | (~(uint32_t)XL[r] & (uint32_t)N0[r])) & 2u;
                        any1 = any1 || need1[r] != 0u;
                    }
                    if (__any_sync(FULL, any1)) {                                 // ~1% of the words
#pragma unroll
                        for (int r = 0; r < RPT; r++)
                            flip1[r] |= rare_slice1(a, (uint64_t)need1[r], z[r][0], z[r][1], z[r][2], z[r][3], XL[r], i, f, sweep,
                                                    a.row0 + (uint32_t)(row + r), RPT, t + 1u);
                    }
                }
                uint32_t present = 0u;                                            // bit 0: class 1, bit 1: class 2
#pragma unroll
                for (int r = 0; r < RPT; r++) {
                    const uint32_t fm = (uint32_t)((int32_t)((uint32_t)flip1[r] << 30) >> 31);   // slice 1 flipped: all ones
                    const uint32_t brnew = brold[r] ^ fm;
                    const uint32_t xrlo = (uint32_t)w[r] ^ brnew ^ (fm & 1u);                     // slice 0 sees the old slice 1
                    const uint32_t xrhi = (uint32_t)(w[r] >> 32) ^ brnew;
                    XR[r] = ((uint64_t)xrhi << 32) | xrlo;
                    if (((XL[r] ^ XR[r]) & todo) != 0ull) present |= 1u;
                    if ((XL[r] & XR[r] & todo) != 0ull) present |= 2u;
                }
                present = __reduce_or_sync(FULL, present);
                LT_MARK(8)
                if (present & 1u) {
                    if (!haveV1) {
                        evalf(fa1, 1, 0, V1);
                        haveV1 = true;
                    }
                    if (anyneed && !haveN1 && fb1 != FID_NONE) {
                        evalf(fb1, 1, 1, N1);
#pragma unroll
                        for (int r = 0; r < RPT; r++) N1[r] &= ~V1[r];
                    }
                }
#pragma unroll
                for (int r = 0; r < RPT; r++) V2[r] = V1[r];                      // stands in when class 2 is empty
                if (present & 2u) {
                    if (fa2 == fa0 && fa0 != FID_GENERIC) {
#pragma unroll
                        for (int r = 0; r < RPT; r++) V2[r] = V0[r];
                    } else if (!(haveV1 && fa2 == fa1 && fa1 != FID_GENERIC)) evalf(fa2, 2, 0, V2);
                    if (anyneed && fb2 != FID_NONE) {
                        evalf(fb2, 2, 1, N2);
#pragma unroll
                        for (int r = 0; r < RPT; r++) N2[r] &= ~V2[r];
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RPT; r++) {
                const uint64_t hi = (XR[r] & V2[r]) | (~XR[r] & V1[r]), lo = (XR[r] & V1[r]) | (~XR[r] & V0[r]);
                ACC[r] = ((XL[r] & hi) | (~XL[r] & lo)) & todo;
            }
            LT_MARK(9)
            if (anyneed) {
                uint64_t NEED[RPT];
                bool anyn = false;
#pragma unroll
                for (int r = 0; r < RPT; r++) {
                    const uint64_t hi = (XR[r] & N2[r]) | (~XR[r] & N1[r]), lo = (XR[r] & N1[r]) | (~XR[r] & N0[r]);
                    NEED[r] = ((XL[r] & hi) | (~XL[r] & lo)) & todo;
                    anyn = anyn || NEED[r] != 0ull;
                }
                if (__any_sync(FULL, anyn)) {
#pragma unroll
                    for (int r = 0; r < RPT; r++)
                        if (__any_sync(FULL, NEED[r] != 0ull))
                            ACC[r] |= rare_draws<true>(a, NEED[r], z[r][0], z[r][1], z[r][2], z[r][3], XL[r], XR[r], i, f, sweep,
                                                       a.row0 + (uint32_t)((ring * RW + r) * a.seg_S), RPT, t + 1u);
                }
            }
#pragma unroll
            for (int r = 0; r < RPT; r++) result[r] = w[r] ^ flip1[r] ^ ACC[r];
        }

        LT_MARK(5)
        // ---- (F) publish: the consumer's ring (or packets for the next block), my own result ring (next
        //      step's left-hand neighbour), the state word, progress every few steps
        if (succ_in) {
            // the consumer has read the result of step t - 1 (a whole step ago in steady state: never a wait);
            // it has then also stored its own step t - 2, which covers the row below requested at step t + 1
            if (!dead && !mbar_wait(a, succ_empty, (t & 1u) ^ 1u, 4u)) dead = true;
            sts_words<RPT>(succ_up + tm * SLOT, result);
            asm volatile("bar.arrive %0, 64;" ::"r"(warp + 1) : "memory");
        } else if (ext_out) {
            const uint32_t tag = t + 1u;
            uint4 *slot = gll_out + (size_t)(t & (LT_DG - 1)) * RW;
#pragma unroll
            for (int r = 0; r < RPT; r++)
                st_vol_v4(slot + r, (uint32_t)result[r], tag, (uint32_t)(result[r] >> 32), tag);
        }
        sts_words<RPT>(ring_res + ((tm + 1u) & 3u) * SLOT, result);
        if (RPT == 2) {
            if (live[0]) *reinterpret_cast<ulonglong2 *>(store) = make_ulonglong2(result[0], result[RPT - 1]);
        } else if (live[0]) *reinterpret_cast<uint64_t *>(store) = result[0];
        store += step_bytes;
        const bool last_step = t + 1u == T;
        if (ext_in && t + 1u >= in_off && !last_step) {
            const uint4 *slot = gll_in + (size_t)((t + 1u - in_off) & (LT_DG - 1)) * RW;
#pragma unroll
            for (int r = 0; r < RPT; r++) pk[r] = ld_vol_v4(slot + r);
        }
        if (warp == 0 && (last_step || ((t + 1u) & a.gmask) == 0u || t + 1u == (uint32_t)LT_PRE)) {
            __syncwarp();                          // the release below covers the stores of all lanes
            if (lane == 0) st_release(prog_mine, t + 1u);
        }
        LT_MARK(6)
        if (++p == (uint32_t)C) {
            p = 0u;
            s++;
            store -= wrap_bytes;
            if (++ms == (uint32_t)a.mcsteps) {
                ms = 0u;
                f++;
            }
        }
    }
#ifdef LT_PROFILE
    if (prof && lane == 0) {
        unsigned long long *d = a.dbg + ((size_t)tk * LT_MAXW + warp) * 16;
        for (int k = 0; k < 10; k++) d[k] = (unsigned long long)pacc[k];
        d[15] = T;
    }
#endif
#undef LT_MARK
}

template <typename T>
int grow(T *&p, size_t &have, size_t want, cudaStream_t stream) { return piqmc_grow(p, have, want, stream); }

size_t chain_smem_bytes(int cw, int rpt)
{
    return (size_t)cw * 5 * LT_DR * 256 * rpt + (size_t)cw * sizeof(WarpMisc);
}

typedef void (*chain_kernel_t)(const LatArgs);
chain_kernel_t chain_kernel(int qa, int seg, int rpt)
{
    if (!qa) return rpt == 2 ? chain_sweep<false, false, 2> : chain_sweep<false, false, 1>;
    if (seg) return rpt == 2 ? chain_sweep<true, true, 2> : chain_sweep<true, true, 1>;
    return rpt == 2 ? chain_sweep<true, false, 2> : chain_sweep<true, false, 1>;
}

int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return s ? atoi(s) : dflt;
}

}  // namespace

int launch_decision_tables(piqmc_ctx *c, int qa, const PiqmcChainStat *d_stat, int nf, const float *d_jp2,
                           const float *d_invT, int force_generic)
{
    const unsigned tgrid = (unsigned)(((size_t)nf * c->nspins + 3) / 4);
    if (qa) chain_tables_kernel<true><<<tgrid, 128, 0, c->stream>>>(d_stat, (uint4 *)c->d_chot, (uint4 *)c->d_ccold, d_jp2, d_invT, c->nspins, nf, force_generic);
    else    chain_tables_kernel<false><<<tgrid, 128, 0, c->stream>>>(d_stat, (uint4 *)c->d_chot, (uint4 *)c->d_ccold, d_jp2, d_invT, c->nspins, nf, force_generic);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}

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

// The launch geometry of the chain pipeline for the current state: rows per thread, chains per block,
// bands per ring.  A ring (all bands of one group of rows) must be resident at once.  Returns false when
// the pipeline cannot run this state (the caller takes the dataflow kernel).
bool chain_geometry(const piqmc_ctx *c, int qa, ChainGeom *g)
{
    if (c->chain_C < 8 || !c->d_cstat || c->nrows <= 0) return false;
    int rpt = (c->nrows % 2 == 0 && c->nrows >= 64) ? 2 : 1;
    const int e_rpt = env_int("PIQMC_CHAIN_RPT", 0);
    if (e_rpt == 1 || (e_rpt == 2 && c->nrows % 2 == 0)) rpt = e_rpt;
    int cwmax = env_int("PIQMC_CHAIN_CW", LT_MAXW);
    cwmax = std::max(1, std::min(cwmax, LT_MAXW));
    const int nchains = c->chain_n;
    const int rw = 32 * rpt;
    const int nrings = (c->nrows + rw - 1) / rw;
    // resident blocks per SM for blocks of cw warps: the occupancy of the kernel that would run
    const chain_kernel_t kern = chain_kernel(qa, qa && c->seg_S > 1, rpt);
    auto blocks_per_sm = [&](int cw) {
        const size_t sm = chain_smem_bytes(cw, rpt);
        if (sm > 227 * 1024) return 0;
        if (cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *)kern, cw * 32, sm) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        return nb;
    };
    int nbands = (nchains + cwmax - 1) / cwmax;
    int cw = (nchains + nbands - 1) / nbands;
    // everything resident at once (few rows): spread the chains over more, smaller blocks so that every
    // SM gets its share -- but not below 8 chains per block (every band boundary is a trip through L2)
    const int e_nb = env_int("PIQMC_CHAIN_BANDS", 0);
    if (e_nb > 0) nbands = std::max(nbands, std::min(e_nb, nchains));
    else {
        const long cap = (long)c->sm_count * blocks_per_sm(cw);
        if ((long)nrings * nbands <= cap) {
            const int want = (int)std::min<long>(c->sm_count / std::max(1, nrings), nchains / 8);
            if (want > nbands) nbands = want;
        }
    }
    cw = (nchains + nbands - 1) / nbands;
    const int bps = blocks_per_sm(cw);
    if (bps < 1 || (long)nbands > (long)c->sm_count * bps) return false;      // a ring must fit the device
    if ((size_t)nrings * nbands >= ((size_t)1 << 31)) return false;
    g->rpt = rpt;
    g->cw = cw;
    g->nbands = nbands;
    g->nrings = nrings;
    g->smem = chain_smem_bytes(cw, rpt);
    return true;
}

int launch_chain_sweeps(piqmc_ctx *c, int qa, int nsched, int mcsteps, const float *h_jp2, const float *h_invT,
                        uint64_t seed, uint32_t row0, uint32_t sweep0)
{
    if (nsched <= 0 || mcsteps <= 0) return PIQMC_OK;
    ChainGeom g;
    PIQMC_REQUIRE(chain_geometry(c, qa, &g), PIQMC_EINVAL, "no chain plan for this graph and state");
    const int N = c->nspins, C = c->chain_C, nchains = c->chain_n;
    // schedule steps per launch: bounded by the memory of the decision tables (512 MB) and by the
    // 31-bit range of the step tags
    size_t max_f = std::max<size_t>(1, ((size_t)512 << 20) / ((size_t)N * 32));
    max_f = std::min<size_t>(max_f, std::max<size_t>(1, ((size_t)1 << 30) / ((size_t)C * mcsteps)));
    const size_t nf_max = std::min<size_t>(max_f, nsched);

    size_t err_have = c->d_err ? 1 : 0, tk_have = c->d_ticket ? 1 : 0;
    if (int rc = grow(c->d_err, err_have, 1, c->stream)) return rc;
    if (int rc = grow(c->d_ticket, tk_have, 1, c->stream)) return rc;
    if (int rc = grow(c->d_chot, c->chot_elems, nf_max * N, c->stream)) return rc;
    if (int rc = grow(c->d_ccold, c->ccold_elems, nf_max * N, c->stream)) return rc;
    const size_t nprog = (size_t)g.nrings * g.nbands;
    if (int rc = grow(c->d_cprog, c->cprog_elems, nprog, c->stream)) return rc;
    const size_t gll_bytes = nprog * LT_DG * 32 * g.rpt * sizeof(uint4);
    {
        char *p = (char *)c->d_cll;
        if (int rc = grow(p, c->cll_bytes, gll_bytes, c->stream)) return rc;
        c->d_cll = p;
    }
    // bands: nchains split as evenly as possible
    std::vector<int> first(g.nbands + 1);
    for (int b = 0; b <= g.nbands; b++) first[b] = (int)(((long)nchains * b) / g.nbands);
    size_t bf_have = c->d_cband ? c->cband_elems : 0;
    if (int rc = grow(c->d_cband, bf_have, (size_t)g.nbands + 1, c->stream)) return rc;
    c->cband_elems = bf_have;
    float *d_par = nullptr;                                   // jp2[nsched] then invT[nsched]
    PIQMC_CUDA(cudaMalloc(&d_par, 2 * (size_t)nsched * sizeof(float)));
    cudaError_t e = cudaMemcpyAsync(d_par, h_jp2, nsched * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_par + nsched, h_invT, nsched * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(c->d_cband, first.data(), first.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_err, 0, sizeof(unsigned int), c->stream);

    LatArgs a;
    a.words = c->d_words;
    a.stat = c->d_cstat;
    a.hot = (const uint4 *)c->d_chot;
    a.cold = (const uint4 *)c->d_ccold;
    a.band_first = c->d_cband;
    a.prog = c->d_cprog;
    a.gll = (uint4 *)c->d_cll;
    a.ticket = c->d_ticket;
    a.err = c->d_err;
    a.nspins = N;
    a.nrows = c->nrows;
    a.C = C;
    a.nchains = nchains;
    a.nbands = g.nbands;
    a.mcsteps = mcsteps;
    a.wrap = c->chain_wrap;
    {
        int G = std::max(1, std::min(8, C / 4));
        const int v = env_int("PIQMC_CHAIN_G", 0);
        if (v >= 1) G = v;
        while (G & (G - 1)) G &= G - 1;                       // power of two
        a.gmask = (uint32_t)G - 1u;
    }
    a.k0 = (uint32_t)seed;
    a.k1 = (uint32_t)(seed >> 32);
    a.row0 = row0;
    a.watchdog_ns = 20000000000ull;
    if (const char *s = getenv("PIQMC_WATCHDOG_MS")) a.watchdog_ns = (unsigned long long)atoll(s) * 1000000ull;
    a.valid = (c->lanes >= 64) ? ~0ull : ((1ull << c->lanes) - 1ull);
    a.seg_P = qa ? c->seg_P : 64;
    a.seg_S = qa ? c->seg_S : 1;
    a.seg_low = a.seg_l1 = a.seg_top = 0ull;
    a.seg_ones = (a.seg_P >= 64) ? ~0ull : ((1ull << a.seg_P) - 1ull);
    for (int k = 0; k < a.seg_S; k++) {
        a.seg_low |= 1ull << (k * a.seg_P);
        a.seg_l1 |= 2ull << (k * a.seg_P);
        a.seg_top |= 1ull << (k * a.seg_P + a.seg_P - 1);
    }
    unsigned long long *d_dbg = nullptr;
    const size_t ndbg = nprog * LT_MAXW * 16;
    if (env_int("PIQMC_CHAIN_PROF", 0)) {
        PIQMC_CUDA(cudaMalloc(&d_dbg, ndbg * sizeof(unsigned long long)));
        PIQMC_CUDA(cudaMemsetAsync(d_dbg, 0, ndbg * sizeof(unsigned long long), c->stream));
    }
    a.dbg = d_dbg;
    const int force_generic = env_int("PIQMC_FORCE_GENERIC_FN", 0);
    const chain_kernel_t kern = chain_kernel(qa, qa && c->seg_S > 1, g.rpt);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);

    for (size_t f0 = 0; f0 < (size_t)nsched && e == cudaSuccess; f0 += nf_max) {
        const int nf = (int)std::min<size_t>(nf_max, nsched - f0);
        a.jp2 = d_par + f0;
        a.invT = d_par + nsched + f0;
        a.nsweeps = nf * mcsteps;
        a.sweep0 = sweep0 + (uint32_t)(f0 * mcsteps);
        const unsigned tgrid = (unsigned)(((size_t)nf * N + 3) / 4);
        if (qa) chain_tables_kernel<true><<<tgrid, 128, 0, c->stream>>>(c->d_cstat, (uint4 *)c->d_chot, (uint4 *)c->d_ccold, a.jp2, a.invT, N, nf, force_generic);
        else    chain_tables_kernel<false><<<tgrid, 128, 0, c->stream>>>(c->d_cstat, (uint4 *)c->d_chot, (uint4 *)c->d_ccold, a.jp2, a.invT, N, nf, force_generic);
        c->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemsetAsync(c->d_ticket, 0, sizeof(unsigned int), c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(c->d_cprog, 0, nprog * sizeof(uint32_t), c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(c->d_cll, 0, gll_bytes, c->stream);
        if (e != cudaSuccess) break;
        kern<<<dim3((unsigned)nprog), dim3((unsigned)(g.cw * 32)), g.smem, c->stream>>>(a);
        c->launches++;
        e = cudaGetLastError();
    }
    cudaError_t e2 = cudaStreamSynchronize(c->stream);        // d_par must outlive the launches
    cudaFree(d_par);
    if (d_dbg) {
        std::vector<unsigned long long> h(ndbg);
        cudaMemcpy(h.data(), d_dbg, ndbg * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        cudaFree(d_dbg);
        static const char *names[10] = {"A guards", "B fetch+wait", "C record", "D row above", "E loads", "draws+result", "F publish", "V0 V1", "N0 N1 slice1 XR", "V2 ACC"};
        double tot[3][10] = {{0}}, steps[3] = {0, 0, 0};      // all warps | first warp of a block | last warp
        for (size_t b = 0; b < nprog; b++) {
            const int band = (int)(b % g.nbands), nW = first[band + 1] - first[band];
            for (int w = 0; w < nW; w++) {
                const unsigned long long *d = &h[(b * LT_MAXW + w) * 16];
                const int cls[3] = {1, w == 0, w == nW - 1};
                for (int q = 0; q < 3; q++)
                    if (cls[q]) {
                        for (int k = 0; k < 10; k++) tot[q][k] += (double)d[k];
                        steps[q] += (double)d[15];
                    }
            }
        }
        fprintf(stderr, "[chain prof] cycles per step: %-14s %8s %8s %8s\n", "", "all", "first", "last");
        for (int k = 0; k < 10; k++)
            fprintf(stderr, "[chain prof]                  %-14s %8.0f %8.0f %8.0f\n", names[k], tot[0][k] / steps[0],
                    tot[1][k] / std::max(1.0, steps[1]), tot[2][k] / std::max(1.0, steps[2]));
    }
    if (e != cudaSuccess || e2 != cudaSuccess) {
        piqmc_set_error("chain sweep launch failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        return PIQMC_ECUDA;
    }
    return piqmc_check_watchdog(c, "chain sweeps");
}
