// level_kernels.cu -- level-synchronous Metropolis sweeps (maxnb <= 4; QA with the reference's Trotter
// neighbours, or SA), one launch for a whole run of sweeps, NO dependency traffic between thread blocks
// of different row groups.
//
// The sweep being reproduced is sequential (piqmc/qmc.pyx:320-357: spins in visiting order, every spin
// sees the new value of the neighbours visited before it); its dependency levels are the colour classes
// (DESIGN.md section 2).  Replicas never interact, so the state is cut the other way round than in
// colour_fast.cu: a GROUP of 32 replica rows (one warp wide: coalesced 256-byte accesses, and the decision
// functions of a spin are warp-uniform) belongs to ONE thread-block cluster, which walks the levels of
// all sweeps of the launch on its own:
//
//     for every step (period q, level class rho):            -- D steps per sweep on a torus of side D
//         warp (cta, w) takes members  first + cta*W + w, + K*W, ...  of the class: one word per lane
//         cluster barrier (K > 1) or block barrier (K == 1)
//
// With the period-major member order of apply_colouring (api.cu) step (q, rho) holds level rho of sweep q
// next to level rho + D of sweep q - 1, ...: every step is as wide as the widest level and consecutive
// sweeps overlap.  The cluster size K is chosen by the host so that the groups fill the device: 128 groups
// of config 5 at 4096 rows run on one block each (a step is 256 words per block), the 16 groups of its
// 8-GPU shard on clusters of 8 (a step is one word per warp).  A step costs one barrier and one trip to
// L2 for the words written in the step before -- not a flag round trip per unit and no ticket, no table
// build, no polling: the decision functions of a (schedule step, spin) are computed once for all replicas
// by the decision-table kernel (chain_kernels.cu) and read as 16-byte records.
// Semantics: oracle_qa_colour / oracle_sa_colour (oracle/piqmc_oracle.c part 3), bit for bit.
#include <stdio.h>
#include <stdlib.h>

#include "colour_device.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int LV_MAXW = 32;              // warps per block, at most

struct LevelArgs {
    uint64_t *words;                     // [N + 1][nrows]
    const PiqmcUnitRec *recs;            // member records in step order
    const PiqmcLevelRec *xrecs;          // staged variant: the same members with their word sources
    const PiqmcLevelRec *stream;         // streamed variant: [K * W][stream_len] records of one period, per warp slot
    int stream_len;
    const int *step_off;                 // static: [period_len + 1]; per-sweep lists: [nsteps + 1] (absolute)
    const int *step_sweep;               // per-sweep lists: sweep of every step
    const PiqmcChainStat *stat;          // [N] sorted couplings and pad (acceptance thresholds, rare path)
    const uint4 *hot, *cold;             // [schedule steps of this launch][N] decision-function records
    const float *jp2, *invT;             // per schedule step of this launch
    int nspins, nrows, nsweeps, nperiods_extra, mcsteps, f_off, period_len, nsteps, per_sweep_lists, K;   // sweep s belongs to schedule step (s + f_off) / mcsteps
    uint32_t k0, k1, row0, sweep0;
    uint64_t valid;
    int seg_P, seg_S;                    // SEG: seg_S replicas of seg_P slices per word; else seg_P = lanes, seg_S = 1
    uint64_t seg_low, seg_l1, seg_top, seg_ones;
    unsigned long long *dbg;             // -DLV_PROFILE: cycles per segment of a step, [block][warp][16]
    int dry;                             // diagnostics (PIQMC_LEVEL_DRY): barriers only
    int shared_draws;                    // serve the uniforms of a step by the whole block (few members per warp and step)
};

struct LvScratch {                       // per warp, rare paths only
    uint32_t thr[48];                    // acceptance thresholds of the 16 patterns x 3 Trotter classes
    uint32_t thr_key;                    // item the thresholds were built for (0: none)
    uint32_t spin, sweep, prow0;         // Philox counter words of the warp's current member (requests served by others)
    uint2 queue[QCAP];                   // pooled draw requests
};

// acceptance thresholds by the warp: the float32 sequence of build_table_warp (colour_device.cuh)
template <bool QA>
__device__ __forceinline__ void lv_build_thr(LvScratch *m, uint32_t key, const PiqmcChainStat *st, float jp2, float invT)
{
    if (m->thr_key == key) return;                  // built earlier for this item (warp-uniform)
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
    if (lane == 0) m->thr_key = key;
    __syncwarp();
}

// ---- the rare paths of an item, out of line ---------------------------------------------------------
template <bool QA>
__device__ __noinline__ uint64_t lv_rare_draws(const LevelArgs &a, LvScratch *m, uint64_t NEED, uint64_t z0, uint64_t z1,
                                               uint64_t z2, uint64_t z3, uint64_t XL, uint64_t XR, uint32_t i, uint32_t f,
                                               uint32_t sweep, uint32_t prow0, uint32_t key)
{
    uint32_t *thr = m->thr;
    lv_build_thr<QA>(m, key, a.stat + i, a.jp2[f], a.invT[f]);
    const auto thr_tab = [&](uint32_t c, uint32_t pat) -> uint32_t { return thr[c * 16u + pat]; };
    const uint64_t z[4] = {z0, z1, z2, z3};
    return resolve_draws<QA>(NEED, z, XL, XR, thr_tab, m->queue, i, sweep, prow0, a.k0, a.k1, QA ? a.seg_P : 64,
                             QA ? a.seg_S : 1, 1);
}
// slice 1 of every replica segment of one word per thread: the lanes of need1 draw their uniform
__device__ __noinline__ uint64_t lv_rare_slice1(const LevelArgs &a, LvScratch *m, const PiqmcChainStat *st, uint64_t need1,
                                                uint64_t z0, uint64_t z1, uint64_t z2, uint64_t z3, uint64_t XL, uint32_t i,
                                                uint32_t f, uint32_t sweep, uint32_t prow, uint32_t key)
{
    uint32_t *thr = m->thr;
    lv_build_thr<true>(m, key, st, __ldg(a.jp2 + f), __ldg(a.invT + f));
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
// a function outside the list of 27: from its truth table
__device__ __noinline__ uint64_t lv_rare_generic(const LevelArgs &a, uint32_t f, uint32_t i, int c, int need, uint64_t z0,
                                                 uint64_t z1, uint64_t z2, uint64_t z3)
{
    const uint4 cold = __ldg(a.cold + (size_t)f * a.nspins + i);
    const uint32_t packed[3] = {cold.x, cold.y, cold.z};      // hacc0 hacc1 | hacc2 hall0 | hall1 hall2
    const int idx = need ? 3 + c : c;
    const uint32_t h = (packed[idx >> 1] >> (16 * (idx & 1))) & 0xFFFFu;
    return eval_generic(h, z0, z1, z2, z3);
}

// ---- one word: the Metropolis decisions of all its lanes ---------------------------------------------
// w: the spin's word; z: disagreement masks with the 4 sorted neighbours, sign-normalised; rc: the hot
// record of (schedule step, spin).  live: the thread's row exists.  Returns the new word.
struct LvWord {
    uint64_t base;            // the word after the decisions that need no uniform (slice 1 included)
    uint64_t NEED, XL, XR;    // lanes whose Metropolis test needs a uniform; their Trotter disagreements
};
template <bool QA, bool SEG>
__device__ __forceinline__ LvWord lv_decide(const LevelArgs &a, LvScratch *m, const PiqmcChainStat *st, uint64_t w0,
                                            const uint64_t (&z)[1][4], const uint4 rc, uint32_t i, uint32_t f,
                                            uint32_t sweep, uint32_t row, bool live, uint32_t key)
{
    const uint64_t w[1] = {w0};
    const uint64_t valid = live ? a.valid : 0ull;
    const uint32_t fa0 = rc.x & 0xFFu, fb0 = (rc.x >> 8) & 0xFFu;
    const bool anyneed = (rc.y >> 16) & 1u, anygen = (rc.y >> 17) & 1u;
    auto evalf = [&](uint32_t fid, int c, int need, uint64_t (&out)[1]) {
        if (!anygen || fid < PIQMC_NCANON) eval_canon<1>(fid, z, out);
        else out[0] = lv_rare_generic(a, f, i, c, need, z[0][0], z[0][1], z[0][2], z[0][3]);
    };
    if (!QA) {
        uint64_t V[1];
        evalf(fa0, 0, 0, V);
        if (fb0 != FID_NONE) {
            uint64_t NEED[1];
            evalf(fb0, 0, 1, NEED);
            NEED[0] &= ~V[0] & valid;
            return LvWord{w[0] ^ (V[0] & valid), NEED[0], 0ull, 0ull};
        }
        return LvWord{w[0] ^ (V[0] & valid), 0ull, 0ull, 0ull};
    }
    // ---- QA, the reference's Trotter neighbours: slices P-1 (old value for everyone; itself for slice
    //      P-1) and 1 (old for slice 0, itself for slice 1, new for slices >= 2), per segment when a word
    //      holds several replicas.  Slice 1 is decided first, from the functions of classes 0 and 1 (its
    //      right-hand neighbour is itself).  Class of a lane = number of Trotter neighbours it disagrees
    //      with: XL + XR.
    const uint32_t fa1 = (rc.x >> 16) & 0xFFu, fb1 = rc.x >> 24;
    const uint32_t fa2 = rc.y & 0xFFu, fb2 = (rc.y >> 8) & 0xFFu;
    uint64_t V0[1], V1[1], V2[1], N0[1] = {0ull}, N1[1] = {0ull}, N2[1] = {0ull}, XL, XR, flip1, ACC;
    const uint64_t todo = valid & ~a.seg_l1;
    evalf(fa0, 0, 0, V0);
    if (SEG) {
        // ---- several replicas per word: all functions up front, 64-bit segment arithmetic
        if (fa1 == fa0 && fa0 != FID_GENERIC) V1[0] = V0[0];
        else evalf(fa1, 1, 0, V1);
        if (anyneed) {
            if (fb0 != FID_NONE) evalf(fb0, 0, 1, N0);
            if (fb1 != FID_NONE) evalf(fb1, 1, 1, N1);
            N0[0] &= ~V0[0];
            N1[0] &= ~V1[0];
        }
        const uint64_t l1 = live ? a.seg_l1 : 0ull;
        const uint64_t bl = ((w[0] & a.seg_top) >> (a.seg_P - 1)) * a.seg_ones;
        const uint64_t brold = ((w[0] & a.seg_l1) >> 1) * a.seg_ones;
        XL = (w[0] ^ bl) & ~a.seg_top;
        flip1 = ((XL & V1[0]) | (~XL & V0[0])) & l1;
        if (anyneed) {
            const uint64_t need1 = ((XL & N1[0]) | (~XL & N0[0])) & l1;
            if (__any_sync(FULL, need1 != 0ull))                               // ~1% of the words
                flip1 |= lv_rare_slice1(a, m, st, need1, z[0][0], z[0][1], z[0][2], z[0][3], XL, i, f, sweep,
                                        a.row0 + row * (uint32_t)a.seg_S, key);
        }
        const uint64_t brnew = (((w[0] ^ flip1) & a.seg_l1) >> 1) * a.seg_ones;
        XR = ((w[0] ^ brnew) & ~a.seg_low) | ((w[0] ^ brold) & a.seg_low);     // slice 0 sees the old slice 1
        if (fa2 == fa1 && fa1 != FID_GENERIC) V2[0] = V1[0];
        else {
            V2[0] = V1[0];
            if (__any_sync(FULL, (XL & XR & todo) != 0ull)) evalf(fa2, 2, 0, V2);
        }
        if (anyneed && fb2 != FID_NONE) {
            evalf(fb2, 2, 1, N2);
            N2[0] &= ~V2[0];
        }
    } else {
        // ---- one replica per word, on 32-bit halves, and every function only when a lane of the warp is in
        //      its class: early in the anneal J_perp is small and the classes share one function, late in the
        //      anneal the slices of a replica agree and only class 0 is populated
        const int tsh = a.seg_P - 1;
        const uint32_t l1lo = live ? 2u : 0u;
        const uint32_t wlo = (uint32_t)w[0];
        const uint32_t bl = (uint32_t)((int32_t)((uint32_t)(w[0] >> tsh) << 31) >> 31);
        const uint32_t brold = (uint32_t)((int32_t)(wlo << 30) >> 31);
        XL = (w[0] ^ (((uint64_t)bl << 32) | bl)) & ~a.seg_top;
        const bool same01 = fa1 == fa0 && fa0 != FID_GENERIC;
        bool haveV1 = same01, haveN1 = false;
        V1[0] = V0[0];                                                         // stands in until class 1 is needed
        const bool s1c1 = __any_sync(FULL, ((uint32_t)XL & l1lo) != 0u);       // slice 1 of some word is in class 1
        if (!haveV1 && s1c1) {
            evalf(fa1, 1, 0, V1);
            haveV1 = true;
        }
        uint32_t f1 = (((uint32_t)XL & (uint32_t)V1[0]) | (~(uint32_t)XL & (uint32_t)V0[0])) & l1lo;
        if (anyneed) {
            if (fb0 != FID_NONE) {
                evalf(fb0, 0, 1, N0);
                N0[0] &= ~V0[0];
            }
            if (s1c1) {
                if (fb1 != FID_NONE) {
                    evalf(fb1, 1, 1, N1);
                    N1[0] &= ~V1[0];
                }
                haveN1 = true;
            }
            const uint32_t need1 = (((uint32_t)XL & (uint32_t)N1[0]) | (~(uint32_t)XL & (uint32_t)N0[0])) & l1lo;
            if (__any_sync(FULL, need1 != 0u))                                 // ~1% of the words
                f1 |= (uint32_t)lv_rare_slice1(a, m, st, (uint64_t)need1, z[0][0], z[0][1], z[0][2], z[0][3], XL, i, f, sweep,
                                               a.row0 + row, key);
        }
        flip1 = (uint64_t)f1;
        const uint32_t fm = (uint32_t)((int32_t)(f1 << 30) >> 31);             // slice 1 flipped: all ones
        const uint32_t brnew = brold ^ fm;
        const uint32_t xrlo = wlo ^ brnew ^ (fm & 1u);                         // slice 0 sees the old slice 1
        const uint32_t xrhi = (uint32_t)(w[0] >> 32) ^ brnew;
        XR = ((uint64_t)xrhi << 32) | xrlo;
        uint32_t present = 0u;                                                 // bit 0: class 1, bit 1: class 2
        if (((XL ^ XR) & todo) != 0ull) present |= 1u;
        if ((XL & XR & todo) != 0ull) present |= 2u;
        present = __reduce_or_sync(FULL, present);
        if (present & 1u) {
            if (!haveV1) {
                evalf(fa1, 1, 0, V1);
                haveV1 = true;
            }
            if (anyneed && !haveN1 && fb1 != FID_NONE) {
                evalf(fb1, 1, 1, N1);
                N1[0] &= ~V1[0];
            }
        }
        V2[0] = V1[0];                                                         // stands in when class 2 is empty
        if (present & 2u) {
            if (fa2 == fa0 && fa0 != FID_GENERIC) V2[0] = V0[0];
            else if (!(haveV1 && fa2 == fa1 && fa1 != FID_GENERIC)) evalf(fa2, 2, 0, V2);
            if (anyneed && fb2 != FID_NONE) {
                evalf(fb2, 2, 1, N2);
                N2[0] &= ~V2[0];
            }
        }
    }
    {
        const uint64_t hi = (XR & V2[0]) | (~XR & V1[0]), lo = (XR & V1[0]) | (~XR & V0[0]);
        ACC = ((XL & hi) | (~XL & lo)) & todo;
    }
    uint64_t NEED = 0ull;
    if (anyneed) {
        const uint64_t hi = (XR & N2[0]) | (~XR & N1[0]), lo = (XR & N1[0]) | (~XR & N0[0]);
        NEED = ((XL & hi) | (~XL & lo)) & todo;
    }
    return LvWord{w[0] ^ flip1 ^ ACC, NEED, XL, XR};
}

// ---- draws shared by the whole block (latency regime: one word per thread and step) --------------------
// A thread whose word has lanes that need a uniform leaves a request; after a block barrier the requests
// are served two per warp and round -- thread (h, b) draws the Philox block of slices 4b..4b+3 of request h
// -- and the accepted lanes are OR-ed into the owner's slot.  Balanced over the block whatever the
// distribution of the requests (they cluster by spin, i.e. by warp).
struct LvReq {
    uint64_t NEED, z[4], XL, XR;
    uint32_t owner;           // warp << 5 | lane of the requesting thread
    uint32_t pad_;
};
static_assert(sizeof(LvReq) == 64, "LvReq must be 64 bytes");

// XOUT: the accepted lanes are flipped in the owner's word of `out` (exchange buffer); else OR-ed into out[owner]
template <bool QA, bool XOUT>
__device__ __forceinline__ void lv_serve_requests(const LevelArgs &a, const LvReq *reqs, int n, const LvScratch *scratch,
                                                  unsigned long long *out, int W)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = lane >> 4, b = lane & 15;
    const int segP = QA ? a.seg_P : 64, segS = QA ? a.seg_S : 1;
    for (int e0 = 2 * warp; e0 < n; e0 += 2 * W) {            // warp-uniform
        const int e = e0 + h;
        if (e < n) {
            const LvReq &r = reqs[e];
            const uint32_t need4 = (uint32_t)(r.NEED >> (4 * b)) & 0xFu;
            if (need4) {
                const LvScratch &o = scratch[r.owner >> 5];
                const int seg = (segS > 1) ? (4 * b) / segP : 0;
                const u32x4 u = philox4x32_10(o.spin, (uint32_t)(b - seg * (segP >> 2)) | ((QA ? PIQMC_STREAM_SWEEP : PIQMC_STREAM_SA) << 16), o.sweep,
                                              o.prow0 + (r.owner & 31u) * (uint32_t)segS + (uint32_t)seg, a.k0, a.k1);
                const uint32_t *thr = o.thr;
                const uint64_t zz[4] = {r.z[0], r.z[1], r.z[2], r.z[3]};
                uint32_t acc4 = 0u;
#pragma unroll
                for (int qd = 0; qd < 4; qd++) {
                    const int k = 4 * b + qd;
                    const uint32_t c = QA ? (uint32_t)((r.XL >> k) & 1) + (uint32_t)((r.XR >> k) & 1) : 0u;
                    if (((need4 >> qd) & 1u) && pick(u, qd) < thr[c * 16u + pattern_at(zz, k)]) acc4 |= 1u << qd;
                }
                if (acc4) {
                    if (XOUT) atomicXor(out + (r.owner & 1023u), (unsigned long long)acc4 << (4 * b));
                    else atomicOr(out + (r.owner & 1023u), (unsigned long long)acc4 << (4 * b));
                }
            }
        }
    }
}

// ---- the sweep kernel ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lv_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// CLUSTER: the row group is shared by the K blocks of a cluster, else one block per group.
// Step barrier of a cluster: a block barrier, then ONE thread per block arrives (release, cluster scope --
// cumulative over the block's stores through the block barrier) on the mbarrier of every block of the
// cluster through distributed shared memory; everybody waits on the local one.  The wait is an acquire at
// CTA scope only: a cluster-scope acquire makes every warp invalidate the L1 (CCTL.IVALL), which was
// measured to stall the loads that follow by ~4 us per step; the words other blocks have written are
// read past the L1 (ld.global.cg) anyway, and everything else the kernel reads is constant.
template <bool QA, bool SEG, bool CLUSTER>
__global__ void __launch_bounds__(LV_MAXW * 32, 1) level_sweep(const __grid_constant__ LevelArgs a)
{
    extern __shared__ __align__(128) unsigned char lv_dyn[];         // shared draws: requests + result slots
    __shared__ LvScratch scratch[LV_MAXW];
    __shared__ __align__(8) uint64_t cl_bar;
    __shared__ int q_count[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = (int)(blockDim.x >> 5);
    const int K = CLUSTER ? a.K : 1;
    const int group = (int)blockIdx.x / K, rank = (int)blockIdx.x - group * K;
    LvScratch *m = &scratch[warp];
    LvReq *reqs = reinterpret_cast<LvReq *>(lv_dyn);
    unsigned long long *acc_out = reinterpret_cast<unsigned long long *>(lv_dyn + (size_t)W * 32 * sizeof(LvReq));
    const bool shared_draws = a.shared_draws != 0;
    if (lane == 0) m->thr_key = 0u;
    if (threadIdx.x == 0) {
        q_count[0] = q_count[1] = 0;
        if (CLUSTER) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lv_smem_u32(&cl_bar)), "r"(K) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if (CLUSTER) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
    uint32_t cl_phase = 0u;

    const int row = group * 32 + lane;
    const bool live = row < a.nrows;
    uint64_t *wrow = a.words + (live ? row : 0);
    const size_t nrows = (size_t)a.nrows;
    const uint32_t segS = (uint32_t)(QA ? a.seg_S : 1);
    const uint32_t prow_warp = a.row0 + (uint32_t)(group * 32) * segS;
    const int slot = rank * W + warp, nslots = K * W;
    uint32_t key = 0u;
#ifdef LV_PROFILE
    long long pt = clock64(), pacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define LV_MARK(k)                          \
    {                                       \
        const long long now_ = clock64();   \
        pacc[k] += now_ - pt;               \
        pt = now_;                          \
    }
#else
#define LV_MARK(k)
#endif

    for (int t = 0; t < a.nsteps; t++) {
        int q, beg, end;
        if (a.per_sweep_lists) {
            q = a.step_sweep[t];
            beg = a.step_off[t];
            end = a.step_off[t + 1];
        } else {
            q = t / a.period_len;
            const int rho = t - q * a.period_len;
            beg = a.step_off[rho];
            end = a.step_off[rho + 1];
        }
        if (a.dry == 1) end = beg;
        LV_MARK(0)
        // shared draws: one round of members per block barrier pair (every warp at most one member)
        const int nrounds = shared_draws ? (end - beg + nslots - 1) / nslots : 1;
        for (int rd = 0; rd < nrounds; rd++) {
            int *qc = &q_count[(t + rd) & 1];
            uint64_t pending = 0ull, pbase = 0ull;
            uint32_t pspin = 0u;
            bool have = false;
            const int mfirst = beg + slot + (shared_draws ? rd * nslots : 0);
            const int mend = shared_draws ? min(end, mfirst + 1) : end;
            for (int mi = mfirst; mi < mend; mi += nslots) {                    // warp-uniform
                const int4 *rp = reinterpret_cast<const int4 *>(a.recs + mi);
                const int4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                const int s = q - r0.y;
                LV_MARK(1)
                if (s < 0 || s >= a.nsweeps) continue;                          // ramp-up / ramp-down periods
                const uint32_t i = (uint32_t)r0.x;
                const uint32_t f = (uint32_t)(s + a.f_off) / (uint32_t)a.mcsteps;
                const uint32_t sweep = a.sweep0 + (uint32_t)s;
                const uint4 rc = __ldg(a.hot + (size_t)f * a.nspins + i);
                const int nb[4] = {r0.z, r0.w, r1.x, r1.y};
#ifdef LV_PROFILE
                if (rc.x == 0x12345678u) key += 7;
                LV_MARK(2)
#endif
                uint64_t w = 0ull, wn[4] = {0ull, 0ull, 0ull, 0ull};
                if (live && a.dry != 2) {
                    // L2 loads: the words may have been written by another block of the cluster one step ago
                    w = __ldcg(wrow + (size_t)i * nrows);
#pragma unroll
                    for (int k = 0; k < 4; k++) wn[k] = __ldcg(wrow + (size_t)nb[k] * nrows);
                }
                uint64_t z[1][4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t sg = __byte_perm(rc.z, 0u, 0x1111u * k);     // byte k, four times
                    z[0][k] = w ^ wn[k] ^ (((uint64_t)sg << 32) | sg);
                }
                key++;
#ifdef LV_PROFILE
                if (z[0][0] == 0x12345678u && z[0][1] == 7 && z[0][2] == 1 && z[0][3] == 2) key += 7;
                LV_MARK(3)
#endif
                if (a.dry >= 2) {                                               // diagnostics: memory traffic only
                    if (a.dry == 2 && rc.x == 0x12345678u && live) wrow[(size_t)i * nrows] = 0ull;
                    if (a.dry == 3 && live) wrow[(size_t)i * nrows] = w ^ ((z[0][0] & z[0][1] & z[0][2] & z[0][3]) & 1ull);
                    continue;
                }
                const LvWord d = lv_decide<QA, SEG>(a, m, a.stat + i, w, z, rc, i, f, sweep, (uint32_t)row, live, key);
                uint64_t result = d.base;
#ifdef LV_PROFILE
                if (result == 0x12345678u) key += 7;
                LV_MARK(4)
#endif
                const uint32_t needy = __ballot_sync(FULL, d.NEED != 0ull);
                if (needy) {
                    if (!shared_draws) {
                        result ^= QA ? lv_rare_draws<true>(a, m, d.NEED, z[0][0], z[0][1], z[0][2], z[0][3], d.XL, d.XR, i, f,
                                                           sweep, prow_warp, key)
                                     : lv_rare_draws<false>(a, m, d.NEED, z[0][0], z[0][1], z[0][2], z[0][3], 0ull, 0ull, i, f,
                                                            sweep, prow_warp, key);
                    } else {
                        // thresholds of this spin for whoever serves the requests; one request per needy word
                        lv_build_thr<QA>(m, key, a.stat + i, a.jp2[f], a.invT[f]);
                        if (lane == 0) {
                            m->spin = i;
                            m->sweep = sweep;
                            m->prow0 = prow_warp;
                        }
                        int base = 0;
                        if (lane == 0) base = atomicAdd(qc, __popc(needy));
                        base = __shfl_sync(FULL, base, 0);
                        if (d.NEED != 0ull) {
                            LvReq &r = reqs[base + __popc(needy & ((1u << lane) - 1u))];
                            r.NEED = d.NEED;
#pragma unroll
                            for (int k = 0; k < 4; k++) r.z[k] = z[0][k];
                            r.XL = d.XL;
                            r.XR = d.XR;
                            r.owner = (uint32_t)threadIdx.x;
                            acc_out[threadIdx.x] = 0ull;
                            pending = d.NEED;
                        }
                        pbase = result;
                        pspin = i;
                        have = true;
                        continue;
                    }
                }
                if (live) wrow[(size_t)i * nrows] = result;
            }
            LV_MARK(5)
            if (shared_draws) {
                __syncthreads();                                                // requests are complete
                const int n = *qc;
                if (n > 0) {                                                    // block-uniform
                    lv_serve_requests<QA, false>(a, reqs, n, scratch, acc_out, W);
                    __syncthreads();
                    if (threadIdx.x == 0) *qc = 0;                              // next use: two rounds from now
                }
                if (have) {
                    if (pending != 0ull) pbase ^= acc_out[threadIdx.x];
                    if (live) wrow[(size_t)pspin * nrows] = pbase;
                }
            }
            LV_MARK(6)
        }
        if (CLUSTER) {
            __syncthreads();
            LV_MARK(7)
            if (threadIdx.x < (unsigned)K) {
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(lv_smem_u32(&cl_bar)), "r"(threadIdx.x));
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
            }
            uint32_t ok = 0u;
            while (!ok) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(ok)
                    : "r"(lv_smem_u32(&cl_bar)), "r"(cl_phase)
                    : "memory");
            }
            cl_phase ^= 1u;
        } else {
            __syncthreads();
        }
        LV_MARK(8)
    }
#ifdef LV_PROFILE
    if (a.dbg && lane == 0) {
        unsigned long long *d = a.dbg + ((size_t)blockIdx.x * LV_MAXW + warp) * 16;
        for (int k = 0; k < 10; k++) d[k] = (unsigned long long)pacc[k];
        d[15] = (unsigned long long)a.nsteps;
    }
#endif
#undef LV_MARK
}


// ---- the staged variant: nothing but shared memory on a step's critical path ---------------------------
// Applies when every warp of the cluster has at most ONE member per step (K * W >= widest step: the
// latency-bound regime) and the colouring is static with a period of at least 3 steps.  Measured on B200, a
// dependent trip to L2 / HBM costs 0.4-0.8 us here, and the plain kernel above makes three per step.  So:
//   * the words written in the last two steps are read from EXCHANGE BUFFERS in the shared memory of the
//     block that wrote them (distributed shared memory inside the cluster): a result goes to
//     xbuf[step % 3][warp][lane] first and to global memory one step later (so that the release of the
//     step barrier never waits for a store in flight);
//   * everything else is requested a step or two ahead with cp.async into per-warp staging slots -- no
//     registers: the member record of step t + 2, and, from the record of step t + 1, its hot record, its
//     own word and the neighbour words that are three or more steps old;
//   * the step offsets of the period sit in shared memory; lanes that need a uniform are served by the
//     whole block (lv_serve_requests), which flips the accepted lanes in the exchange buffer itself.
__device__ __forceinline__ void lv_cp16(uint32_t smem, const void *gmem, uint32_t srcsize)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem), "l"(gmem), "r"(srcsize) : "memory");
}

template <bool QA, bool SEG, bool CLUSTER>
__global__ void __launch_bounds__(LV_MAXW * 32, 1) level_sweep_x(const __grid_constant__ LevelArgs a)
{
    extern __shared__ __align__(128) unsigned char lv_dyn[];
    __shared__ LvScratch scratch[LV_MAXW];
    __shared__ __align__(8) uint64_t cl_bar;
    __shared__ int q_count[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = (int)(blockDim.x >> 5);
    const int K = CLUSTER ? a.K : 1;
    const int group = (int)blockIdx.x / K, rank = (int)blockIdx.x - group * K;
    const int P = a.period_len;
    LvScratch *m = &scratch[warp];
    // carve the dynamic shared memory
    unsigned char *dp = lv_dyn;
    unsigned long long *xbuf = reinterpret_cast<unsigned long long *>(dp);         // [3][W][32]
    dp += (size_t)3 * W * 256;
    unsigned long long *wbuf = reinterpret_cast<unsigned long long *>(dp) + (size_t)warp * 2 * 5 * 32;   // [W][2][5][32]
    dp += (size_t)W * 2560;
    PiqmcLevelRec *recb = reinterpret_cast<PiqmcLevelRec *>(dp) + warp * 3;          // [W][3]
    dp += (size_t)W * 96;
    uint4 *hotb = reinterpret_cast<uint4 *>(dp) + warp * 2;                          // [W][2]
    dp += (size_t)W * 32;
    PiqmcChainStat *statb = reinterpret_cast<PiqmcChainStat *>(dp) + warp * 2;        // [W][2]
    dp += (size_t)W * 64;
    LvReq *reqs = reinterpret_cast<LvReq *>(dp);                                      // [W * 32]
    dp += (size_t)W * 32 * sizeof(LvReq);
    int *soff = reinterpret_cast<int *>(dp);                                          // [P + 1]

    for (int k = (int)threadIdx.x; k <= P; k += (int)blockDim.x) soff[k] = a.step_off[k];
    if (lane == 0) m->thr_key = 0u;
    if (threadIdx.x == 0) {
        q_count[0] = q_count[1] = 0;
        if (CLUSTER) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lv_smem_u32(&cl_bar)), "r"(K) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if (CLUSTER) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
    uint32_t cl_phase = 0u;

    const int row = group * 32 + lane;
    const bool live = row < a.nrows;                         // nrows is even: lanes 2l and 2l + 1 live together
    const size_t nrows = (size_t)a.nrows;
    uint64_t *wrow = a.words + (live ? row : 0);
    const uint32_t cpsz = live ? 16u : 0u;
    const uint32_t segS = (uint32_t)(QA ? a.seg_S : 1);
    const uint32_t prow_warp = a.row0 + (uint32_t)(group * 32) * segS;
    const int slot = rank * W + warp;
    const int nsteps = a.nsteps;
    const uint32_t xbuf_u32 = lv_smem_u32(xbuf), wbuf_u32 = lv_smem_u32(wbuf), recb_u32 = lv_smem_u32(recb),
                   hotb_u32 = lv_smem_u32(hotb), statb_u32 = lv_smem_u32(statb);
    uint32_t key = 0u;

    // request the member record of step u (period q, class rho) into slot u % 3
    auto issue_rec = [&](int u, int rho) {
        const int mi = soff[rho] + slot;
        const bool has = u < nsteps && mi < soff[rho + 1];
        if (has) {
            if (lane < 2) lv_cp16(recb_u32 + (uint32_t)(u % 3) * 32u + (uint32_t)lane * 16u,
                                  reinterpret_cast<const char *>(a.xrecs + mi) + lane * 16, 16u);
        } else if (lane == 0) recb[u % 3].spin = -1;
    };
    // a neighbour word comes from global memory: written three or more steps ago, or never in this launch
    auto from_global = [&](uint32_t src, int s) { return ((src >> 10) & 3u) == 0u || (s == 0 && ((src >> 12) & 1u)); };
    // request everything step u needs from global memory (its record has landed): slots u % 2
    auto issue_data = [&](int u, int q) {
        const PiqmcLevelRec &r = recb[u % 3];
        const int i = r.spin;
        const int s = q - r.sweepoff;
        if (i < 0 || s < 0 || s >= a.nsweeps) return;                          // warp-uniform
        const uint32_t f = (uint32_t)(s + a.f_off) / (uint32_t)a.mcsteps;
        if (lane == 1) lv_cp16(hotb_u32 + (uint32_t)(u & 1) * 16u, a.hot + (size_t)f * a.nspins + i, 16u);
        if (lane == 3 || lane == 5)                                            // thresholds are built from it (rare path)
            lv_cp16(statb_u32 + (uint32_t)(u & 1) * 32u + (uint32_t)(lane - 3) * 8u,
                    reinterpret_cast<const char *>(a.stat + i) + (lane - 3) * 8, 16u);
        if ((lane & 1) == 0) {
            const uint32_t dst = wbuf_u32 + (uint32_t)(u & 1) * 1280u + (uint32_t)lane * 8u;
            lv_cp16(dst, wrow + (size_t)i * nrows, cpsz);
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (from_global(r.src[k], s)) lv_cp16(dst + 256u * (uint32_t)(k + 1), wrow + (size_t)r.nb[k] * nrows, cpsz);
        }
    };

    // (period, class) of steps t, t + 1, t + 2
    int q0 = 0, r0 = 0, q1 = 0, r1 = 1, q2, r2;
    if (r1 == P) { r1 = 0; q1 = 1; }
    q2 = q1; r2 = r1 + 1;
    if (r2 == P) { r2 = 0; q2++; }
    // prologue: records of steps 0 and 1, then the data of step 0
    issue_rec(0, r0);
    issue_rec(1, r1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    issue_data(0, q0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    int prev_spin = -1;                                    // member of the step before (its word goes to global memory now)
#ifdef LV_PROFILE
    long long pt = clock64(), pacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define LV_MARK(k)                          \
    {                                       \
        const long long now_ = clock64();   \
        pacc[k] += now_ - pt;               \
        pt = now_;                          \
    }
#else
#define LV_MARK(k)
#endif

    for (int t = 0; t < nsteps; t++) {
        // (1) what was requested a step ago has landed: record of t + 1, data of t
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        LV_MARK(0)
        // (2) the word of step t - 1, from the exchange buffer to global memory
        if (prev_spin >= 0 && live)
            wrow[(size_t)prev_spin * nrows] = xbuf[((size_t)((t + 2) % 3) * W + warp) * 32 + lane];
        // (3) requests for the next steps
        issue_data(t + 1, q1);
        issue_rec(t + 2, r2);
        asm volatile("cp.async.commit_group;" ::: "memory");
        LV_MARK(1)

        // (4) this step's member
        int *qc = &q_count[t & 1];
        const PiqmcLevelRec &rec = recb[t % 3];
        const int ispin = rec.spin;
        const int s = q0 - rec.sweepoff;
        const bool valid = ispin >= 0 && s >= 0 && s < a.nsweeps;               // warp-uniform
        prev_spin = valid ? ispin : -1;
        if (valid) {
            const uint32_t i = (uint32_t)ispin;
            const uint32_t f = (uint32_t)(s + a.f_off) / (uint32_t)a.mcsteps;
            const uint32_t sweep = a.sweep0 + (uint32_t)s;
            const uint4 rc = hotb[t & 1];
            const unsigned long long *wb = wbuf + (size_t)(t & 1) * 160 + lane;
            const uint64_t w = wb[0];
            uint64_t z[1][4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t src = rec.src[k];
                uint64_t wn;
                if (from_global(src, s)) wn = wb[32 * (k + 1)];
                else {
                    const uint32_t age = (src >> 10) & 3u;                      // 1 or 2 steps ago
                    const uint32_t buf = (uint32_t)(t + 3 - (int)age) % 3u;
                    const uint32_t local = xbuf_u32 + ((buf * (uint32_t)W + (src & 31u)) * 32u + (uint32_t)lane) * 8u;
                    if (CLUSTER) {
                        uint32_t remote;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"((src >> 5) & 15u));
                        asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(wn) : "r"(remote) : "memory");
                    } else {
                        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(wn) : "r"(local) : "memory");
                    }
                }
                const uint32_t sg = __byte_perm(rc.z, 0u, 0x1111u * k);         // byte k, four times
                z[0][k] = w ^ wn ^ (((uint64_t)sg << 32) | sg);
            }
            key++;
#ifdef LV_PROFILE
            if (z[0][0] == 0x12345678u && z[0][1] == 7 && z[0][2] == 1 && z[0][3] == 2) key += 7;
            LV_MARK(2)
#endif
            const LvWord d = lv_decide<QA, SEG>(a, m, statb + (t & 1), w, z, rc, i, f, sweep, (uint32_t)row, live, key);
            xbuf[((size_t)(t % 3) * W + warp) * 32 + lane] = d.base;
            const uint32_t needy = __ballot_sync(FULL, d.NEED != 0ull);
            LV_MARK(3)
            if (needy) {
                // thresholds of this spin for whoever serves the requests; one request per needy word
                lv_build_thr<QA>(m, key, statb + (t & 1), __ldg(a.jp2 + f), __ldg(a.invT + f));
                int base = 0;
                if (lane == 0) {
                    m->spin = i;
                    m->sweep = sweep;
                    m->prow0 = prow_warp;
                    base = atomicAdd(qc, __popc(needy));
                }
                base = __shfl_sync(FULL, base, 0);
                if (d.NEED != 0ull) {
                    LvReq &r = reqs[base + __popc(needy & ((1u << lane) - 1u))];
                    r.NEED = d.NEED;
#pragma unroll
                    for (int k = 0; k < 4; k++) r.z[k] = z[0][k];
                    r.XL = d.XL;
                    r.XR = d.XR;
                    r.owner = (uint32_t)threadIdx.x;
                }
            }
        }
        // (5) the block's uniforms, then the step barrier
        LV_MARK(4)
        __syncthreads();                                                        // words and requests are complete
        const int n = *qc;
        LV_MARK(5)
        if (n > 0) {                                                            // block-uniform
            lv_serve_requests<QA, true>(a, reqs, n, scratch, xbuf + (size_t)(t % 3) * W * 32, W);
            __syncthreads();
            if (threadIdx.x == 0) *qc = 0;                                      // next use: two steps from now
        }
        LV_MARK(6)
        if (CLUSTER) {
            if (threadIdx.x < (unsigned)K) {
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(lv_smem_u32(&cl_bar)), "r"(threadIdx.x));
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
            }
            uint32_t ok = 0u;
            while (!ok) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(ok)
                    : "r"(lv_smem_u32(&cl_bar)), "r"(cl_phase)
                    : "memory");
            }
            cl_phase ^= 1u;
        }
        LV_MARK(7)
        q0 = q1; r0 = r1; q1 = q2; r1 = r2;
        if (++r2 == P) { r2 = 0; q2++; }
    }
#ifdef LV_PROFILE
    if (a.dbg && lane == 0) {
        unsigned long long *d = a.dbg + ((size_t)blockIdx.x * LV_MAXW + warp) * 16;
        for (int k = 0; k < 10; k++) d[k] = (unsigned long long)pacc[k];
        d[15] = (unsigned long long)nsteps;
    }
#endif
#undef LV_MARK
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (prev_spin >= 0 && live)
        wrow[(size_t)prev_spin * nrows] = xbuf[((size_t)((nsteps + 2) % 3) * W + warp) * 32 + lane];
    if (CLUSTER) {
        // nobody may leave while a block of the cluster can still read its exchange buffers
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
}


// ---- the streamed variant: many members per warp and step (throughput regime) --------------------------
// Every warp slot of the cluster owns a STREAM of 32-byte member records, laid out by the host in the order
// the warp will process them: ceil(width / slots) positions per step (dummy records where a step has fewer
// members), the last position of a step flagged.  All slots have the same length and step boundaries, the
// same stream serves every period.  A warp runs a three-stage pipeline over its stream:
//   * records arrive 16 positions ahead (cp.async, a ring of 32 in shared memory);
//   * the data of position e + 2 -- own row, four neighbour rows (256 bytes each: 32 replica rows of one
//     spin) and the hot record -- is requested by ONE lane as bulk copies (TMA, cp.async.bulk) that complete
//     on the stage's mbarrier, while position e computes.  A position of a later step is requested only
//     after the barrier that opens its step (its neighbours' words are written until then);
//   * position e: wait for its stage, five shared-memory reads, decide, store.
// No registers are spent on the pipeline, one instruction requests 256 bytes, and the loop has no ticket,
// flag, table build or per-unit block barrier: ~230 warp-instructions per 32 words against 383 of the
// dataflow kernel.
constexpr int LT_STAGES = 3;
constexpr int LT_STAGE_BYTES = 1408;       // 5 rows of 256 bytes, the hot record, padding to 128
constexpr int LT_RING = 32;                // member records per warp in shared memory
constexpr uint32_t LT_LAST = 1u << 16;     // sweepoff field: the last position of a step

template <bool QA, bool SEG, bool CLUSTER>
__global__ void __launch_bounds__(LV_MAXW * 32, 1) level_sweep_t(const __grid_constant__ LevelArgs a)
{
    extern __shared__ __align__(128) unsigned char lv_dyn[];
    __shared__ LvScratch scratch[LV_MAXW];
    __shared__ __align__(8) uint64_t cl_bar;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = (int)(blockDim.x >> 5);
    const int K = CLUSTER ? a.K : 1;
    const int group = (int)blockIdx.x / K, rank = (int)blockIdx.x - group * K;
    LvScratch *m = &scratch[warp];
    unsigned char *stage0 = lv_dyn + (size_t)warp * LT_STAGES * LT_STAGE_BYTES;
    PiqmcLevelRec *ring = reinterpret_cast<PiqmcLevelRec *>(lv_dyn + (size_t)W * LT_STAGES * LT_STAGE_BYTES) + warp * LT_RING;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(lv_dyn + (size_t)W * LT_STAGES * LT_STAGE_BYTES +
                                                  (size_t)W * LT_RING * sizeof(PiqmcLevelRec)) + warp * LT_STAGES;
    if (lane == 0) {
        m->thr_key = 0u;
        for (int k = 0; k < LT_STAGES; k++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lv_smem_u32(&mbar[k])) : "memory");
    }
    if (threadIdx.x == 0 && CLUSTER) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lv_smem_u32(&cl_bar)), "r"(K) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (CLUSTER) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
    uint32_t cl_phase = 0u;

    const int row = group * 32 + lane;
    const bool live = row < a.nrows;                         // nrows is even
    const size_t nrows = (size_t)a.nrows;
    uint64_t *wrow = a.words + (live ? row : 0);
    const uint64_t *grow = a.words + (size_t)group * 32;     // the group's first row
    const uint32_t rowbytes = (uint32_t)min(32, a.nrows - group * 32) * 8u;
    const uint32_t txbytes = 5u * rowbytes + 16u;
    const uint32_t segS = (uint32_t)(QA ? a.seg_S : 1);
    const uint32_t prow_warp = a.row0 + (uint32_t)(group * 32) * segS;
    const PiqmcLevelRec *stream = a.stream + (size_t)(rank * W + warp) * a.stream_len;
    const int L = a.stream_len;
    const int total = L * (a.nsweeps + a.nperiods_extra);    // positions of this launch (the host keeps it below 2^31)
    const uint32_t ring_u32 = lv_smem_u32(ring), stage_u32 = lv_smem_u32(stage0), mbar_u32 = lv_smem_u32(mbar);
    uint32_t key = 0u;

    int q0 = 0, pos0 = 0;                                    // period and stream index of the current position
    // records of positions [e0, e0 + 16) into their ring slots (position mod 32); idx0: stream index of e0
    auto issue_recs = [&](int e0, int idx0) {
        const int e = e0 + (lane >> 1);
        if (e < total) {
            int idx = idx0 + (lane >> 1);
            while (idx >= L) idx -= L;
            lv_cp16(ring_u32 + (uint32_t)(e & (LT_RING - 1)) * 32u + (uint32_t)(lane & 1) * 16u,
                    reinterpret_cast<const char *>(stream + idx) + (lane & 1) * 16, 16u);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // the member of position e = current + dx (its record is in the ring): spin or -1, sweep, last-of-step flag
    auto member = [&](int e, int dx, int &s, bool &last) {
        const PiqmcLevelRec &r = ring[e & (LT_RING - 1)];
        int q = q0, p = pos0 + dx;
        while (p >= L) {
            p -= L;
            q++;
        }
        s = q - (int)(r.sweepoff & 0xFFFF);
        last = (r.sweepoff & LT_LAST) != 0u;
        return (r.spin >= 0 && s >= 0 && s < a.nsweeps) ? r.spin : -1;
    };
    uint32_t ph_issue = 0u, ph_wait = 0u;                    // mbarrier phases per stage (bit k)
    // request the data of position e: one lane, six bulk copies on the stage's mbarrier
    auto issue_data = [&](int e, int dx) {
        int s;
        bool last;
        const int i = member(e, dx, s, last);
        if (i < 0) return;                                                     // warp-uniform
        const uint32_t st = (uint32_t)(e % LT_STAGES);
        ph_issue ^= 1u << st;
        if (lane == 0) {
            const PiqmcLevelRec &r = ring[e & (LT_RING - 1)];
            const uint32_t f = (uint32_t)(s + a.f_off) / (uint32_t)a.mcsteps;
            const uint32_t dst = stage_u32 + st * LT_STAGE_BYTES, bar = mbar_u32 + st * 8u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(txbytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(grow + (size_t)i * nrows), "r"(rowbytes), "r"(bar) : "memory");
#pragma unroll
            for (int k = 0; k < 4; k++)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst + 256u * (uint32_t)(k + 1)), "l"(grow + (size_t)r.nb[k] * nrows), "r"(rowbytes), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                         ::"r"(dst + 1280u), "l"(a.hot + (size_t)f * a.nspins + i), "r"(bar) : "memory");
        }
    };

    // prologue: the first 32 records, then the data of the first positions of step 0
    issue_recs(0, 0);
    issue_recs(16, 16);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    int issued = 0;                                          // positions whose data has been requested

    for (int e = 0; e < total; e++) {
        // records: at the start of a block of 16 positions the block after the next is requested (its ring slots
        // are those of the block just finished); it has landed 8 positions later and is first needed after 14
        if ((e & 15) == 0 && e > 0) issue_recs(e + 16, pos0 + 16);
        if ((e & 15) == 8) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
        }
        // data: position e itself if the barrier before it kept it back, then up to two positions ahead
        // within the same step
        if (issued < e) issued = e;
        while (issued <= e + 2 && issued < total) {
            if (issued > e) {
                int s_;
                bool last_;
                member(issued - 1, issued - 1 - e, s_, last_);
                if (last_) break;                                               // the next position opens a new step
            }
            issue_data(issued, issued - e);
            issued++;
        }
        int s;
        bool last;
        const int ispin = member(e, 0, s, last);
        if (ispin >= 0) {                                                       // warp-uniform
            const uint32_t st = (uint32_t)(e % LT_STAGES);
            const uint32_t bar = mbar_u32 + st * 8u, par = (ph_wait >> st) & 1u;
            ph_wait ^= 1u << st;
            uint32_t ok = 0u;
            while (!ok) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(ok)
                    : "r"(bar), "r"(par)
                    : "memory");
            }
            const unsigned char *sp = stage0 + st * LT_STAGE_BYTES;
            const uint32_t i = (uint32_t)ispin;
            const uint32_t f = (uint32_t)(s + a.f_off) / (uint32_t)a.mcsteps;
            const uint32_t sweep = a.sweep0 + (uint32_t)s;
            const uint4 rc = *reinterpret_cast<const uint4 *>(sp + 1280);
            const uint64_t w = reinterpret_cast<const uint64_t *>(sp)[lane];
            uint64_t z[1][4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint64_t wn = reinterpret_cast<const uint64_t *>(sp + 256 * (k + 1))[lane];
                const uint32_t sg = __byte_perm(rc.z, 0u, 0x1111u * k);         // byte k, four times
                z[0][k] = w ^ wn ^ (((uint64_t)sg << 32) | sg);
            }
            __syncwarp();                                                       // the stage may be refilled
            key++;
            const LvWord d = lv_decide<QA, SEG>(a, m, a.stat + i, w, z, rc, i, f, sweep, (uint32_t)row, live, key);
            uint64_t result = d.base;
            if (__any_sync(FULL, d.NEED != 0ull))
                result ^= QA ? lv_rare_draws<true>(a, m, d.NEED, z[0][0], z[0][1], z[0][2], z[0][3], d.XL, d.XR, i, f, sweep,
                                                   prow_warp, key)
                             : lv_rare_draws<false>(a, m, d.NEED, z[0][0], z[0][1], z[0][2], z[0][3], 0ull, 0ull, i, f, sweep,
                                                    prow_warp, key);
            if (live) wrow[(size_t)i * nrows] = result;
        }
        if (last) {
            // the words of this step are read by bulk copies (async proxy) after the barrier
            asm volatile("fence.proxy.async.global;" ::: "memory");
            __syncthreads();
            if (CLUSTER) {
                if (threadIdx.x < (unsigned)K) {
                    uint32_t remote;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(lv_smem_u32(&cl_bar)), "r"(threadIdx.x));
                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
                }
                uint32_t ok = 0u;
                while (!ok) {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                        "selp.u32 %0, 1, 0, p;\n\t}"
                        : "=r"(ok)
                        : "r"(lv_smem_u32(&cl_bar)), "r"(cl_phase)
                        : "memory");
                }
                cl_phase ^= 1u;
            }
        }
        if (++pos0 == L) {
            pos0 = 0;
            q0++;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

typedef void (*level_kernel_t)(const LevelArgs);
level_kernel_t level_kernel(int qa, int seg, int cluster)
{
    if (!qa) return cluster ? level_sweep<false, false, true> : level_sweep<false, false, false>;
    if (seg) return cluster ? level_sweep<true, true, true> : level_sweep<true, true, false>;
    return cluster ? level_sweep<true, false, true> : level_sweep<true, false, false>;
}

level_kernel_t level_kernel_t_(int qa, int seg, int cluster)
{
    if (!qa) return cluster ? level_sweep_t<false, false, true> : level_sweep_t<false, false, false>;
    if (seg) return cluster ? level_sweep_t<true, true, true> : level_sweep_t<true, true, false>;
    return cluster ? level_sweep_t<true, false, true> : level_sweep_t<true, false, false>;
}

level_kernel_t level_kernel_x(int qa, int seg, int cluster)
{
    if (!qa) return cluster ? level_sweep_x<false, false, true> : level_sweep_x<false, false, false>;
    if (seg) return cluster ? level_sweep_x<true, true, true> : level_sweep_x<true, true, false>;
    return cluster ? level_sweep_x<true, false, true> : level_sweep_x<true, false, false>;
}

int lv_env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return s ? atoi(s) : dflt;
}

}  // namespace

// Launch geometry: warps per block and blocks per cluster for `width` members in the widest step.
// Every group of 32 rows gets K blocks; K grows while the groups do not fill the device and a step still
// has a word for every warp.
void level_geometry(const piqmc_ctx *c, int width, int *warps, int *K)
{
    const int groups = (c->nrows + 31) / 32;
    int w = LV_MAXW;
    int k = 1;
    while (k < 8 && groups * k * 2 <= c->sm_count && width >= 2 * k * w) k *= 2;
    if (width < w) w = std::max(1, width);
    const int ew = lv_env_int("PIQMC_LEVEL_WARPS", 0), ek = lv_env_int("PIQMC_LEVEL_K", 0);
    if (ew >= 1 && ew <= LV_MAXW) w = ew;
    if (ek == 1 || ek == 2 || ek == 4 || ek == 8 || ek == 16) k = ek;
    *warps = w;
    *K = k;
}

// Staged variant: one member per warp and step.  The fewest blocks per cluster whose 32-warp blocks cover
// the widest step, then the warps spread evenly.  false: the step is too wide (throughput regime).
bool level_staged_geometry(int width, int *warps, int *K)
{
    int k = 1;
    while (k < 16 && k * LV_MAXW < width) k *= 2;
    if (k * LV_MAXW < width || width < 1) return false;
    *K = k;
    *warps = (width + k - 1) / k;
    return true;
}

// Streamed variant: the member records of one period, laid out per warp slot in processing order (see
// level_sweep_t).  Cached in the context for the last geometry.
static int level_build_stream(piqmc_ctx *c, int nslots)
{
    if (c->d_stream && c->stream_slots == nslots) return PIQMC_OK;
    const int P = c->lv_period;
    const std::vector<int> &off = c->h_lvoff;
    int L = 0;
    for (int r = 0; r < P; r++) L += std::max(1, (off[r + 1] - off[r] + nslots - 1) / nslots);
    std::vector<PiqmcLevelRec> st((size_t)nslots * L);
    int base = 0;
    for (int r = 0; r < P; r++) {
        const int cnt = std::max(1, (off[r + 1] - off[r] + nslots - 1) / nslots);
        for (int sl = 0; sl < nslots; sl++)
            for (int j = 0; j < cnt; j++) {
                PiqmcLevelRec &x = st[(size_t)sl * L + base + j];
                const int mi = off[r] + sl + j * nslots;
                x.spin = -1;
                x.sweepoff = 0;
                for (int z = 0; z < 4; z++) {
                    x.nb[z] = c->nspins;
                    x.src[z] = 0;
                }
                if (mi < off[r + 1]) {
                    const PiqmcUnitRec &u = c->h_recs[mi];
                    x.spin = u.spin;
                    x.sweepoff = u.sweepoff;
                    for (int z = 0; z < 4; z++) x.nb[z] = u.nb[z];
                }
                if (j == cnt - 1) x.sweepoff |= (int32_t)LT_LAST;
            }
        base += cnt;
    }
    PIQMC_CUDA(cudaStreamSynchronize(c->stream));
    if (c->d_stream) PIQMC_CUDA(cudaFree(c->d_stream));
    c->d_stream = nullptr;
    PIQMC_CUDA(cudaMalloc(&c->d_stream, st.size() * sizeof(PiqmcLevelRec)));
    PIQMC_CUDA(cudaMemcpy(c->d_stream, st.data(), st.size() * sizeof(PiqmcLevelRec), cudaMemcpyHostToDevice));
    c->stream_slots = nslots;
    c->stream_len = L;
    return PIQMC_OK;
}

size_t level_streamed_smem(int W)
{
    return (size_t)W * LT_STAGES * LT_STAGE_BYTES + (size_t)W * LT_RING * sizeof(PiqmcLevelRec) + (size_t)W * LT_STAGES * 8;
}

size_t level_staged_smem(int W, int P)
{
    return (size_t)3 * W * 256 + (size_t)W * 2560 + (size_t)W * 96 + (size_t)W * 32 + (size_t)W * 64 +
           (size_t)W * 32 * sizeof(LvReq) + (size_t)(P + 1) * sizeof(int);
}

// nsweeps sweeps; sweep s belongs to schedule step (s + f_off) / mcsteps of the nf steps in h_jp2/h_invT.
// recs/step_off/step_sweep: device arrays (see LevelArgs); static colouring: period_len steps per period,
// nperiods_extra ramp periods; per-sweep lists (d_step_sweep != null): nsteps_lists steps.
int launch_level_sweeps(piqmc_ctx *c, int qa, int nsweeps, int mcsteps, int f_off, const float *h_jp2,
                        const float *h_invT, int nf_all, uint64_t seed, uint32_t row0, uint32_t sweep0,
                        const PiqmcUnitRec *d_recs, const int *d_step_off, const int *d_step_sweep, int period_len,
                        int nperiods_extra, int nsteps_lists, int width)
{
    if (nsweeps <= 0 || mcsteps <= 0) return PIQMC_OK;
    const int N = c->nspins;
    const int per_sweep_lists = d_step_sweep != nullptr;
    // schedule steps per launch: bounded by the memory of the decision tables (512 MB)
    const size_t nf_cap = std::max<size_t>(1, ((size_t)512 << 20) / ((size_t)N * 32));
    PIQMC_REQUIRE(!per_sweep_lists || (size_t)nf_all <= nf_cap, PIQMC_EINVAL,
                  "too many schedule steps in one chunk of per-sweep orders");
    const size_t nf_max = std::min<size_t>(nf_cap, nf_all);
    if (int rc = piqmc_grow(c->d_chot, c->chot_elems, nf_max * N, c->stream)) return rc;
    if (int rc = piqmc_grow(c->d_ccold, c->ccold_elems, nf_max * N, c->stream)) return rc;
    float *d_par = nullptr;                                   // jp2[nf_all] then invT[nf_all]
    PIQMC_CUDA(cudaMalloc(&d_par, 2 * (size_t)nf_all * sizeof(float)));
    cudaError_t e = cudaMemcpyAsync(d_par, h_jp2, nf_all * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_par + nf_all, h_invT, nf_all * sizeof(float), cudaMemcpyHostToDevice, c->stream);

    LevelArgs a;
    a.words = c->d_words;
    a.recs = d_recs;
    a.step_off = d_step_off;
    a.step_sweep = d_step_sweep;
    a.stat = c->d_lstat;
    a.hot = (const uint4 *)c->d_chot;
    a.cold = (const uint4 *)c->d_ccold;
    a.nspins = N;
    a.nrows = c->nrows;
    a.mcsteps = mcsteps;
    a.period_len = period_len;
    a.per_sweep_lists = per_sweep_lists;
    a.k0 = (uint32_t)seed;
    a.k1 = (uint32_t)(seed >> 32);
    a.row0 = row0;
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
    int warps = LV_MAXW, K = 1;
    level_geometry(c, width, &warps, &K);
    // the staged variant whenever it applies: static colouring, a period of 3 or more steps, one member per
    // warp and step, an even number of rows (16-byte copies of row pairs)
    const bool staged = !per_sweep_lists && c->d_xrecs && c->lvx_K > 0 && period_len >= 3 && c->nrows % 2 == 0 &&
                        d_recs == c->d_recs && lv_env_int("PIQMC_LEVEL_STAGED", 1) != 0 &&
                        // ... and all clusters resident at once (else the streamed variant keeps the device busier)
                        (((c->nrows + 31) / 32) * c->lvx_K <= c->sm_count || lv_env_int("PIQMC_LEVEL_STAGED", 1) == 2) &&
                        lv_env_int("PIQMC_LEVEL_DRY", 0) == 0 && level_staged_smem(c->lvx_W, period_len) <= 200 * 1024;
    if (staged) {
        warps = c->lvx_W;
        K = c->lvx_K;
    }
    // the streamed variant for the rest of the static colourings (several members per warp and step)
    const bool streamed = !staged && !per_sweep_lists && d_recs == c->d_recs && !c->h_recs.empty() && c->nrows % 2 == 0 &&
                          lv_env_int("PIQMC_LEVEL_STREAM", 1) != 0 && lv_env_int("PIQMC_LEVEL_DRY", 0) == 0 &&
                          c->lv_period > 0 && c->lv_period < 32768;
    a.stream = nullptr;
    a.stream_len = 0;
    if (streamed) {
        warps = LV_MAXW;
        if (int rc = level_build_stream(c, K * warps)) {
            cudaStreamSynchronize(c->stream);
            cudaFree(d_par);
            return rc;
        }
        a.stream = c->d_stream;
        a.stream_len = c->stream_len;
        if ((long long)c->stream_len * ((long long)nsweeps + nperiods_extra) >= (1ll << 31)) {
            cudaStreamSynchronize(c->stream);
            cudaFree(d_par);
            piqmc_set_error("too many sweeps for one call of the streamed level kernel");
            return PIQMC_EINVAL;
        }
    }
    a.nperiods_extra = nperiods_extra;
    a.xrecs = c->d_xrecs;
    a.K = K;
    a.dry = lv_env_int("PIQMC_LEVEL_DRY", 0);
    a.dbg = nullptr;
#ifdef LV_PROFILE
    const size_t ndbg = (size_t)((c->nrows + 31) / 32) * K * LV_MAXW * 16;
    if (lv_env_int("PIQMC_LEVEL_PROF", 0)) {
        PIQMC_CUDA(cudaMalloc(&a.dbg, ndbg * sizeof(unsigned long long)));
        PIQMC_CUDA(cudaMemsetAsync(a.dbg, 0, ndbg * sizeof(unsigned long long), c->stream));
    }
#endif
    // few members per warp and step: the step is latency-bound and ends with its slowest warp
    a.shared_draws = lv_env_int("PIQMC_LEVEL_SHARED", (width + K * warps - 1) / (K * warps) <= 2 ? 1 : 0);
    const size_t dyn = streamed ? level_streamed_smem(warps) : staged ? level_staged_smem(warps, period_len)
                              : (a.shared_draws ? (size_t)warps * 32 * (sizeof(LvReq) + 8) : 0);
    const int groups = (c->nrows + 31) / 32;
    const int force_generic = lv_env_int("PIQMC_FORCE_GENERIC_FN", 0);
    const level_kernel_t kern = streamed ? level_kernel_t_(qa, qa && c->seg_S > 1, K > 1)
                                : staged ? level_kernel_x(qa, qa && c->seg_S > 1, K > 1) : level_kernel(qa, qa && c->seg_S > 1, K > 1);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(dyn, 1024));
    if (e == cudaSuccess && K > 8)
        e = cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);

    // sub-launches of at most nf_max schedule steps (static colourings only; per-sweep lists fit by the check above)
    int s0 = 0;
    while (s0 < nsweeps && e == cudaSuccess) {
        const int fa = (s0 + f_off) / mcsteps;                               // first schedule step of this launch
        const int nf = (int)std::min<size_t>(nf_max, (size_t)(nf_all - fa));
        const int send = std::min<long long>(nsweeps, (long long)(fa + nf) * mcsteps - f_off);
        a.jp2 = d_par + fa;
        a.invT = d_par + nf_all + fa;
        a.nsweeps = send - s0;
        a.f_off = (s0 + f_off) - fa * mcsteps;
        a.sweep0 = sweep0 + (uint32_t)s0;
        a.nsteps = per_sweep_lists ? nsteps_lists : (a.nsweeps + nperiods_extra) * period_len;
        if (int rc = launch_decision_tables(c, qa, c->d_lstat, nf, a.jp2, a.invT, force_generic)) {
            cudaStreamSynchronize(c->stream);
            cudaFree(d_par);
            return rc;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(groups * K));
        cfg.blockDim = dim3((unsigned)(warps * 32));
        cfg.dynamicSmemBytes = dyn;
        cfg.stream = c->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)K;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = K > 1 ? 1 : 0;
        e = cudaLaunchKernelEx(&cfg, kern, a);
        c->launches++;
        s0 = send;
    }
    cudaError_t e2 = cudaStreamSynchronize(c->stream);        // d_par must outlive the launches
    cudaFree(d_par);
#ifdef LV_PROFILE
    if (a.dbg) {
        std::vector<unsigned long long> h(ndbg);
        cudaMemcpy(h.data(), a.dbg, ndbg * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        cudaFree(a.dbg);
        static const char *names_p[9] = {"step offsets", "member record", "hot record", "state words", "decide", "requests", "serve+store", "block barrier", "cluster wait"};
        static const char *names_x[9] = {"cp.async wait", "store + issue", "operands", "decide", "requests", "barrier A", "serve", "step barrier", "-"};
        const char **names = staged ? names_x : names_p;
        double tot[9] = {0}, mx[9] = {0}, steps = 0;
        for (size_t b = 0; b < ndbg / 16; b++) {
            if (h[b * 16 + 15] == 0) continue;
            for (int k = 0; k < 9; k++) {
                tot[k] += (double)h[b * 16 + k];
                mx[k] = std::max(mx[k], (double)h[b * 16 + k] / (double)h[b * 16 + 15]);
            }
            steps += (double)h[b * 16 + 15];
        }
        for (int k = 0; k < 9; k++)
            fprintf(stderr, "[level prof] cycles per step: %-14s mean %8.0f  max over warps %8.0f\n", names[k], tot[k] / steps, mx[k]);
    }
#endif
    if (e != cudaSuccess || e2 != cudaSuccess) {
        piqmc_set_error("level sweep launch failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        return PIQMC_ECUDA;
    }
    return PIQMC_OK;
}
