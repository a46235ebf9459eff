// colour_kernels.cu -- production colour-class Metropolis sweeps on bit-packed state.
//
// State: uint64 word[spin][row] (row fastest: consecutive threads = consecutive rows, so every
// access is coalesced whatever the shape of the colour class); bit `lane` = Trotter slice (QA)
// or replica-in-group (SA).  One thread owns one (spin, row) word of the class being updated;
// all neighbour words belong to other classes and are constant during the launch, so the launch
// is race-free for any proper colouring.  Classes are processed in ascending order; when they
// are the dependency levels of a sequential visiting order (piqmc_order_levels) the sweep
// equals the sequential sweep in that order.
//
// Semantics: oracle/piqmc_oracle.c part 3 (oracle_qa_colour / oracle_sa_colour), reproduced
// bit-exactly by both variants below:
//   generic  any maxnb / lane count / Trotter mode; float sums per lane in registers;
//   fast     (colour_fast.cu) maxnb <= 4, reference Trotter mode or SA: bitwise evaluation of
//            all 64 lanes at once, one dataflow launch for a whole run of sweeps.
//
// Bandwidth: per launch each word of the class is read and written once and its neighbour
// words are read once (mostly L2 hits): ~0.25 B of HBM traffic per attempt at 64 lanes.  The
// kernels are bound by instruction issue, not by memory.
#include "common.cuh"

namespace {

// +-2J selected by a bit: bit==0 (spins agree) -> -2J, bit==1 -> +2J.  negJ2 = -2J.
__device__ __forceinline__ float flip_sign(float negJ2, uint32_t bit)
{
    return __int_as_float(__float_as_int(negJ2) ^ (int)(bit << 31));
}

// the uniform of (row, lane, spin, sweep): Philox block (spin, lane>>2, sweep, row), word lane&3
__device__ __forceinline__ uint32_t lane_uniform(int lane, uint32_t spin, uint32_t sweep, uint32_t prow,
                                                 uint32_t k0, uint32_t k1, uint32_t stream)
{
    const u32x4 r = philox4x32_10(spin, (uint32_t)(lane >> 2) | (stream << 16), sweep, prow, k0, k1);
    const int q = lane & 3;
    return q == 0 ? r.x : (q == 1 ? r.y : (q == 2 ? r.z : r.w));
}

// ------------------------------------------------------------------------------------------
// Generic variant.  Phase 1 accumulates the in-slice sums of all lanes in registers (neighbour
// loop outside, lanes unrolled inside: per-lane order is table order, as the specification
// requires).  Phase 2 walks the lanes in order, adds the Trotter terms from the *current* word
// (so slice k >= 2 sees the new bit 1, as in the sequential slice-major order) and applies
// Metropolis; the uniform of a lane is only computed when the lane needs it.
// ------------------------------------------------------------------------------------------
template <int NL, bool QA, int TROTTER>
__global__ void __launch_bounds__(128) colour_sweep_generic(
    uint64_t *__restrict__ words, int nspins, int nrows, int maxnb, const int32_t *__restrict__ idx_t,
    const float *__restrict__ J_t, const int32_t *__restrict__ members, int nmembers, int lanes,
    float jp2, float invT, uint32_t k0, uint32_t k1, uint32_t row0, uint32_t sweep)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)nmembers * nrows) return;
    const int row = (int)(tid % nrows);
    const int i = members[tid / nrows];
    uint64_t *wrow = words + row;                     // word of spin s at wrow[s*nrows]
    uint64_t w = wrow[(size_t)i * nrows];

    float e[NL];
#pragma unroll
    for (int k = 0; k < NL; k++) e[k] = 0.0f;
    for (int n = 0; n < maxnb; n++) {
        const int j = idx_t[(size_t)n * nspins + i];
        const float negJ2 = -2.0f * J_t[(size_t)n * nspins + i];
        const uint64_t x = (j == i) ? w : (w ^ wrow[(size_t)j * nrows]);
#pragma unroll
        for (int k = 0; k < NL; k++)
            e[k] = __fadd_rn(e[k], flip_sign(negJ2, (uint32_t)(x >> k) & 1u));
    }

    const float njp2 = -jp2;
    const uint32_t prow = row0 + (uint32_t)row;
#pragma unroll
    for (int pass = 0; pass < ((QA && TROTTER == 1) ? 2 : 1); pass++) {
#pragma unroll
        for (int k = 0; k < NL; k++) {
            if (QA && TROTTER == 1 && (k & 1) != pass) continue;
            if (k < lanes) {
                float ee = e[k];
                if (QA) {
                    const uint32_t own = (uint32_t)(w >> k) & 1u;
                    int kl, kr;
                    if (TROTTER == 1) {
                        kl = (k == 0) ? lanes - 1 : k - 1;
                        kr = (k == lanes - 1) ? 0 : k + 1;
                    } else {
                        kl = lanes - 1;
                        kr = 1;
                    }
                    const uint32_t bl = (uint32_t)(w >> kl) & 1u;
                    const uint32_t br = (uint32_t)(w >> kr) & 1u;
                    // exact sum of the two Trotter terms (0 or +-2*jp2), one rounding
                    const float tsum = __fadd_rn(flip_sign(njp2, own ^ bl), flip_sign(njp2, own ^ br));
                    ee = __fadd_rn(ee, tsum);
                }
                ee = __fadd_rn(ee, 0.0f);          // canonical zero (-0 -> +0)
                bool acc = QA ? (ee > 0.0f) : (ee >= 0.0f);
                if (!acc) {
                    const float x = __fmul_rn(ee, invT);
                    if (x >= PIQMC_XCUT)
                        acc = lane_uniform(k, (uint32_t)i, sweep, prow, k0, k1, QA ? PIQMC_STREAM_SWEEP : PIQMC_STREAM_SA) < colour_thresh(x);
                }
                if (acc) w ^= (1ull << k);
            }
        }
    }
    wrow[(size_t)i * nrows] = w;
}


// ------------------------------------------------------------------------------------------
// Resident variant: small graphs.  When the words of all spins of a block's rows fit in shared memory
// (N * 8 bytes per row), the whole run of sweeps happens there: a row is private to the thread(s) that own
// it, so the spins are visited one after the other in the sequential order the colouring stands for (the
// classes in ascending order; or the per-sweep visiting orders themselves) without any barrier, flag or
// launch in between, and global memory is touched twice -- state in, state out.  Any maxnb.  Per-lane
// arithmetic of the generic variant (same float sequence, same uniforms).  SPLIT = threads per word: SA
// words hold 64 independent replicas, 8 threads take 8 lanes each (few rows would otherwise leave the device
// empty: 65536 SA replicas are 1024 words per spin); QA lanes are coupled and stay with one thread.
// ------------------------------------------------------------------------------------------
template <int NL, bool QA, int TROTTER, int SPLIT>
__global__ void __launch_bounds__(128) resident_sweeps(
    uint64_t *__restrict__ words, int nspins, int nrows, int maxnb, const int32_t *__restrict__ idx,
    const float *__restrict__ J32, const int32_t *__restrict__ order, int per_sweep_orders, int nsweeps,
    const float *__restrict__ jp2s, const float *__restrict__ invTs, int lanes, uint32_t k0, uint32_t k1,
    uint32_t row0, uint32_t sweep0)
{
    extern __shared__ __align__(16) unsigned char rs_smem[];
    const int B = (int)blockDim.x / SPLIT;                         // rows of this block
    uint64_t *st = reinterpret_cast<uint64_t *>(rs_smem);          // [nspins][B]
    int32_t *sidx = reinterpret_cast<int32_t *>(st + (size_t)nspins * B);   // [nspins][maxnb]
    float *snj2 = reinterpret_cast<float *>(sidx + (size_t)nspins * maxnb); // -2 J
    const int r = (int)threadIdx.x / SPLIT, part = (int)threadIdx.x % SPLIT;
    const int row = (int)blockIdx.x * B + r;
    const bool live = row < nrows;
    for (int i = part; i < nspins; i += SPLIT) st[(size_t)i * B + r] = live ? words[(size_t)i * nrows + row] : 0ull;
    for (int e = (int)threadIdx.x; e < nspins * maxnb; e += (int)blockDim.x) {
        sidx[e] = idx[e];
        snj2[e] = -2.0f * J32[e];
    }
    __syncthreads();
    const unsigned amask = __ballot_sync(0xffffffffu, live);       // the threads of a word stay or leave together
    if (!live) return;
    const int l0 = part * NL;                                     // first lane of this thread
    const uint32_t prow = row0 + (uint32_t)row;
    constexpr uint32_t STREAM = QA ? PIQMC_STREAM_SWEEP : PIQMC_STREAM_SA;

    for (int s = 0; s < nsweeps; s++) {
        const float invT = invTs[s];
        const float njp2 = QA ? -jp2s[s] : 0.0f;
        const uint32_t sweep = sweep0 + (uint32_t)s;
        const int32_t *ord = order + (per_sweep_orders ? (size_t)s * nspins : 0);
        for (int t = 0; t < nspins; t++) {
            const int i = __ldg(ord + t);
            uint64_t w = st[(size_t)i * B + r];
            float e[NL];
#pragma unroll
            for (int k = 0; k < NL; k++) e[k] = 0.0f;
            for (int n = 0; n < maxnb; n++) {
                const int j = sidx[i * maxnb + n];
                const float negJ2 = snj2[i * maxnb + n];
                const uint64_t x = ((j == i) ? w : (w ^ st[(size_t)j * B + r])) >> l0;
#pragma unroll
                for (int k = 0; k < NL; k++)
                    e[k] = __fadd_rn(e[k], flip_sign(negJ2, (uint32_t)(x >> k) & 1u));
            }
            u32x4 blk[(NL + 3) / 4];                               // Philox blocks of this thread's lanes, on demand
            uint32_t have = 0u;
            uint64_t flips = 0ull;                                 // SPLIT > 1: this thread's accepted lanes
#pragma unroll
            for (int pass = 0; pass < ((QA && TROTTER == 1) ? 2 : 1); pass++) {
#pragma unroll
                for (int k = 0; k < NL; k++) {
                    if (QA && TROTTER == 1 && (k & 1) != pass) continue;
                    const int lane = l0 + k;
                    if (lane < lanes) {
                        float ee = e[k];
                        if (QA) {
                            const uint32_t own = (uint32_t)(w >> lane) & 1u;
                            int kl, kr;
                            if (TROTTER == 1) {
                                kl = (lane == 0) ? lanes - 1 : lane - 1;
                                kr = (lane == lanes - 1) ? 0 : lane + 1;
                            } else {
                                kl = lanes - 1;
                                kr = 1;
                            }
                            const uint32_t bl = (uint32_t)(w >> kl) & 1u;
                            const uint32_t br = (uint32_t)(w >> kr) & 1u;
                            const float tsum = __fadd_rn(flip_sign(njp2, own ^ bl), flip_sign(njp2, own ^ br));
                            ee = __fadd_rn(ee, tsum);
                        }
                        ee = __fadd_rn(ee, 0.0f);          // canonical zero (-0 -> +0)
                        bool acc = QA ? (ee > 0.0f) : (ee >= 0.0f);
                        if (!acc) {
                            const float x = __fmul_rn(ee, invT);
                            if (x >= PIQMC_XCUT) {
                                const int q = k >> 2;                      // lanes are 4-aligned: block (lane >> 2)
                                if (!((have >> q) & 1u)) {
                                    blk[q] = philox4x32_10((uint32_t)i, (uint32_t)(lane >> 2) | (STREAM << 16), sweep, prow, k0, k1);
                                    have |= 1u << q;
                                }
                                const uint32_t u = (k & 3) == 0 ? blk[q].x : ((k & 3) == 1 ? blk[q].y : ((k & 3) == 2 ? blk[q].z : blk[q].w));
                                acc = u < colour_thresh(x);
                            }
                        }
                        if (acc) {
                            if (SPLIT == 1) w ^= (1ull << lane);
                            else flips |= 1ull << lane;
                        }
                    }
                }
            }
            if (SPLIT > 1) {                               // lanes are independent (SA): merge the threads' slices
#pragma unroll
                for (int d = 1; d < SPLIT; d <<= 1) {
                    const uint32_t lo = __shfl_xor_sync(amask, (uint32_t)flips, d);
                    const uint32_t hi = __shfl_xor_sync(amask, (uint32_t)(flips >> 32), d);
                    flips |= ((uint64_t)hi << 32) | lo;
                }
                w ^= flips;                                // (the shuffles: every thread of the word has read its inputs)
                if (part == 0) st[(size_t)i * B + r] = w;
                __syncwarp(amask);
            } else {
                st[(size_t)i * B + r] = w;
            }
        }
    }
    for (int i = part; i < nspins; i += SPLIT) words[(size_t)i * nrows + row] = st[(size_t)i * B + r];
}

// ------------------------------------------------------------------------------------------
// Resident variant for INTEGER couplings (the multispin-coded +-J kernel of BASELINE.json configs[3];
// piqmc/sa.pyx:339-382 XORs bit-packed replicas and then walks the 64 lanes one by one): when every
// coupling and field of the graph is a small integer multiple of a power of two q, the in-slice energy
// difference of a lane is 2q (2n - W) with n = sum of the weights of the columns the lane disagrees with
// (after flipping the columns with negative coupling) -- every partial sum of the specification's float32
// sequence is exact, so the order does not matter.  n is accumulated for all 64 lanes at once in five bit
// planes (ripple-carry adders on 64-bit words); the decisions "accept by sign" / "needs a uniform" are
// monotone in n, so per Trotter class they are two comparisons of the bit-sliced n with thresholds that
// the warp works out once per (sweep, spin) -- lane n of the warp evaluates the specification's float32
// expressions for n -- and only the lanes that need a uniform are visited one by one.
// QA with the reference's Trotter neighbours, or SA; one replica per word (QA) / 64 replicas per word (SA).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t ge_const(const uint64_t (&b)[5], uint32_t theta)
{
    // [n >= theta] for the bit-sliced n (theta <= 32: 32 and more is never reached)
    if (theta >= 32u) return 0ull;
    uint64_t ge = ~0ull;
#pragma unroll
    for (int p = 0; p < 5; p++) ge = ((theta >> p) & 1u) ? (b[p] & ge) : (b[p] | ge);   // theta is warp-uniform
    return ge;
}

template <bool QA>
__global__ void __launch_bounds__(128) resident_sweeps_int(
    uint64_t *__restrict__ words, int nspins, int nrows, int maxnb, const int32_t *__restrict__ idx,
    const int8_t *__restrict__ iw, float unit, const int32_t *__restrict__ order, int per_sweep_orders, int nsweeps,
    const float *__restrict__ jp2s, const float *__restrict__ invTs, int lanes, uint32_t k0, uint32_t k1,
    uint32_t row0, uint32_t sweep0)
{
    extern __shared__ __align__(16) unsigned char rs_smem[];
    const int B = (int)blockDim.x;
    uint64_t *st = reinterpret_cast<uint64_t *>(rs_smem);          // [nspins][B]
    int32_t *sidx = reinterpret_cast<int32_t *>(st + (size_t)nspins * B);   // [nspins][maxnb]
    int8_t *sw = reinterpret_cast<int8_t *>(sidx + (size_t)nspins * maxnb); // signed weights
    __shared__ uint32_t thrtab[4][3][32];                          // per warp: threshold of (class, n)
    const int r = (int)threadIdx.x, lane = r & 31, warp = r >> 5;
    const int row = (int)blockIdx.x * B + r;
    const bool live = row < nrows;                                 // dead rows keep going: the warp needs its lanes
    for (int i = 0; i < nspins; i++) st[(size_t)i * B + r] = live ? words[(size_t)i * nrows + row] : 0ull;
    for (int e = r; e < nspins * maxnb; e += B) {
        sidx[e] = idx[e];
        sw[e] = iw[e];
    }
    __syncthreads();
    const uint32_t prow = row0 + (uint32_t)row;
    constexpr uint32_t STREAM = QA ? PIQMC_STREAM_SWEEP : PIQMC_STREAM_SA;
    constexpr int NC = QA ? 3 : 1;
    const uint64_t valid = (lanes >= 64) ? ~0ull : ((1ull << lanes) - 1ull);
    const uint64_t top = 1ull << (lanes - 1);
    uint32_t(*thr)[32] = thrtab[warp];

    for (int s = 0; s < nsweeps; s++) {
        const float invT = invTs[s];
        const float jp2 = QA ? jp2s[s] : 0.0f;
        const uint32_t sweep = sweep0 + (uint32_t)s;
        const int32_t *ord = order + (per_sweep_orders ? (size_t)s * nspins : 0);
        for (int t = 0; t < nspins; t++) {
            const int i = __ldg(ord + t);
            const uint64_t w = st[(size_t)i * B + r];
            // ---- n in bit planes
            uint64_t b[5] = {0ull, 0ull, 0ull, 0ull, 0ull};
            int W = 0;
            for (int n = 0; n < maxnb; n++) {
                const int wt = sw[i * maxnb + n];                 // warp-uniform
                if (wt == 0) continue;
                const int j = sidx[i * maxnb + n];
                uint64_t z = (j == i) ? w : (w ^ st[(size_t)j * B + r]);
                if (wt < 0) z = ~z;
                const int aw = wt < 0 ? -wt : wt;
                W += aw;
#pragma unroll
                for (int jb = 0; jb < 3; jb++)
                    if ((aw >> jb) & 1) {
                        uint64_t carry = z;
#pragma unroll
                        for (int p = jb; p < 5; p++) {
                            const uint64_t tt = b[p] & carry;
                            b[p] ^= carry;
                            carry = tt;
                        }
                    }
            }
            // ---- thresholds of n, once per warp: lane n evaluates the specification for n
            uint32_t tha[NC], thn[NC];
            {
                const float e0 = __fmul_rn(2.0f * unit, (float)(2 * lane - W));   // exact
                __syncwarp();
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    float ee = e0;
                    if (QA) ee = __fadd_rn(ee, (c == 0) ? -2.0f * jp2 : (c == 1 ? 0.0f : 2.0f * jp2));
                    ee = __fadd_rn(ee, 0.0f);
                    const bool acc = QA ? (ee > 0.0f) : (ee >= 0.0f);
                    const float x = __fmul_rn(ee, invT);
                    const bool need = !acc && x >= PIQMC_XCUT;
                    thr[c][lane] = need ? colour_thresh(x) : 0u;
                    const uint32_t ma = __ballot_sync(0xffffffffu, acc && lane <= W);
                    const uint32_t mn = __ballot_sync(0xffffffffu, (acc || need) && lane <= W);
                    tha[c] = ma ? (uint32_t)(__ffs(ma) - 1) : 32u;            // monotone in n: {n >= tha}
                    thn[c] = mn ? (uint32_t)(__ffs(mn) - 1) : 32u;
                }
                __syncwarp();
            }
            uint64_t V[3], Nd[3];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                V[c] = ge_const(b, tha[c]);
                Nd[c] = (thn[c] < tha[c]) ? (ge_const(b, thn[c]) & ~V[c]) : 0ull;
            }
            // one needy lane: its uniform against the threshold of (class, n)
            u32x4 blk;
            int blk_q = -1;
            auto draw = [&](int k, uint32_t c) -> bool {
                const uint32_t n = (uint32_t)((b[0] >> k) & 1) | (uint32_t)((b[1] >> k) & 1) << 1 | (uint32_t)((b[2] >> k) & 1) << 2 |
                                   (uint32_t)((b[3] >> k) & 1) << 3 | (uint32_t)((b[4] >> k) & 1) << 4;
                if (blk_q != (k >> 2)) {
                    blk = philox4x32_10((uint32_t)i, (uint32_t)(k >> 2) | (STREAM << 16), sweep, prow, k0, k1);
                    blk_q = k >> 2;
                }
                const uint32_t u = (k & 3) == 0 ? blk.x : ((k & 3) == 1 ? blk.y : ((k & 3) == 2 ? blk.z : blk.w));
                return u < thr[c][n];
            };
            uint64_t result;
            if (!QA) {
                uint64_t acc = V[0] & valid, need = Nd[0] & valid;
                while (need) {
                    const int k = __ffsll((long long)need) - 1;
                    need &= need - 1;
                    if (draw(k, 0u)) acc |= 1ull << k;
                }
                result = w ^ acc;
            } else {
                // the reference's Trotter neighbours: slices P-1 (old value; itself for slice P-1) and 1 (old for
                // slice 0, itself for slice 1, new for slices >= 2): slice 1 is decided first
                const uint64_t bl = (w & top) ? ~0ull : 0ull;
                const uint64_t br_old = (w & 2ull) ? ~0ull : 0ull;
                const uint64_t XL = (w ^ bl) & ~top;
                const uint32_t c1 = (uint32_t)(XL >> 1) & 1u;                // right neighbour of slice 1 is itself
                uint64_t flip1 = 0ull;
                if ((c1 ? V[1] : V[0]) & 2ull) flip1 = 2ull;
                else if (((c1 ? Nd[1] : Nd[0]) & 2ull) && draw(1, c1)) flip1 = 2ull;
                const uint64_t br_new = br_old ^ (flip1 ? ~0ull : 0ull);
                const uint64_t XR = ((w ^ br_new) & ~1ull) | ((w ^ br_old) & 1ull);   // slice 0 sees the old slice 1
                const uint64_t todo = valid & ~2ull;
                const uint64_t C0 = ~(XL | XR), C1 = XL ^ XR, C2 = XL & XR;
                uint64_t acc = ((C0 & V[0]) | (C1 & V[1]) | (C2 & V[2])) & todo;
                uint64_t need = ((C0 & Nd[0]) | (C1 & Nd[1]) | (C2 & Nd[2])) & todo;
                while (need) {
                    const int k = __ffsll((long long)need) - 1;
                    need &= need - 1;
                    const uint32_t c = (uint32_t)((XL >> k) & 1) + (uint32_t)((XR >> k) & 1);
                    if (draw(k, c)) acc |= 1ull << k;
                }
                result = w ^ flip1 ^ acc;
            }
            st[(size_t)i * B + r] = result;
        }
    }
    if (live)
        for (int i = 0; i < nspins; i++) words[(size_t)i * nrows + row] = st[(size_t)i * B + r];
}

// ------------------------------------------------------------------------------------------
// Carry mode: the AS-SHIPPED semantics of qmc.QuantumAnneal (piqmc/qmc.pyx:98-136) at production speed.
// The reference resets its running energy difference once per slice, not per spin: within a slice sweep
// ediff is a float32 carry over all spins visited so far, and the carry decides the moves.  A slice is
// therefore one sequential chain per sweep -- but the slices of a replica only meet in the Trotter terms,
// which read slices P-1 and 1 of the spin being visited, and every spin is visited once per sweep: the P
// chains of a replica can advance in lockstep over the visiting order if, at every step, slice 1 decides
// first (slice 0 still sees its old value, slices >= 2 its new one) and everybody reads slice P-1 before it
// moves.  One warp = one replica (row); lane l carries the chains of slices l and l + 32; the word of the
// visited spin and its neighbours are warp-uniform loads; the flips of a step are collected by ballot.
// Specification: oracle_qa_carry (oracle/piqmc_oracle.c part 4), bit for bit.
// ------------------------------------------------------------------------------------------
template <bool RESIDENT>
__global__ void __launch_bounds__(256) qa_carry_kernel(
    uint64_t *__restrict__ words, int nspins, int nrows, int maxnb, const int32_t *__restrict__ idx,
    const float *__restrict__ J32, const int32_t *__restrict__ order, int per_sweep_orders, int nsweeps,
    const float *__restrict__ jp2s, const float *__restrict__ invTs, int lanes, uint32_t k0, uint32_t k1,
    uint32_t row0, uint32_t sweep0)
{
    extern __shared__ __align__(16) unsigned char rs_smem[];
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= nrows) return;                                  // whole warps
    // RESIDENT: the replica's words live in shared memory for the whole run (8 N bytes per warp)
    uint64_t *mine = reinterpret_cast<uint64_t *>(rs_smem) + (size_t)(threadIdx.x >> 5) * nspins;
    uint64_t *wrow = words + row;
    if (RESIDENT) {
        for (int i = lane; i < nspins; i += 32) mine[i] = wrow[(size_t)i * nrows];
        __syncwarp();
    }
    const auto load = [&](int i) -> uint64_t { return RESIDENT ? mine[i] : __ldcg(wrow + (size_t)i * nrows); };
    const uint32_t prow = row0 + (uint32_t)row;
    const int top = lanes - 1;
    for (int s = 0; s < nsweeps; s++) {
        const float invT = invTs[s];
        const float njp2 = -jp2s[s];
        const uint32_t sweep = sweep0 + (uint32_t)s;
        const int32_t *ord = per_sweep_orders ? order + (size_t)s * nspins : order;
        float ed[2] = {0.0f, 0.0f};                            // the carries of slices lane, lane + 32
        for (int t = 0; t < nspins; t++) {
            const int i = ord ? __ldg(ord + t) : t;
            const uint64_t w = load(i);
            // in-slice terms, table order
#pragma unroll 1
            for (int n = 0; n < maxnb; n++) {
                const int j = __ldg(idx + (size_t)i * maxnb + n);
                const float negJ2 = -2.0f * __ldg(J32 + (size_t)i * maxnb + n);
                const uint64_t x = (j == i) ? w : (w ^ load(j));
                ed[0] = __fadd_rn(ed[0], flip_sign(negJ2, (uint32_t)(x >> lane) & 1u));
                if (lanes > 32) ed[1] = __fadd_rn(ed[1], flip_sign(negJ2, (uint32_t)(x >> (lane + 32)) & 1u));
            }
            const uint32_t old_top = (uint32_t)(w >> top) & 1u, old1 = (uint32_t)(w >> 1) & 1u;
            // the decision of slice k with the right-hand Trotter neighbour's bit `rb`: carries on, no reset
            auto decide = [&](int k, int h, uint32_t rb) -> bool {
                const uint32_t own = (uint32_t)(w >> k) & 1u;
                float e = __fadd_rn(ed[h], flip_sign(njp2, own ^ (k == top ? own : old_top)));
                e = __fadd_rn(e, flip_sign(njp2, own ^ rb));
                ed[h] = e;
                if (e > 0.0f) return true;
                const float x = __fmul_rn(e, invT);
                if (!(x >= PIQMC_XCUT)) return false;
                const u32x4 r = philox4x32_10((uint32_t)i, (uint32_t)(k >> 2) | (PIQMC_STREAM_CARRY << 16), sweep, prow, k0, k1);
                const uint32_t u = (k & 3) == 0 ? r.x : ((k & 3) == 1 ? r.y : ((k & 3) == 2 ? r.z : r.w));
                return u < colour_thresh(x);
            };
            // slice 1 first (its right-hand neighbour is itself)
            bool f0 = false, f1 = false;
            if (lane == 1) f0 = decide(1, 0, old1);
            const uint32_t new1 = old1 ^ (uint32_t)__shfl_sync(0xffffffffu, (int)f0, 1);
            if (lane != 1 && lane < lanes) f0 = decide(lane, 0, lane == 0 ? old1 : new1);
            if (lane + 32 < lanes) f1 = decide(lane + 32, 1, new1);
            const uint32_t lo = __ballot_sync(0xffffffffu, f0), hi = lanes > 32 ? __ballot_sync(0xffffffffu, f1) : 0u;
            if (lane == 0) {
                if (RESIDENT) mine[i] = w ^ (((uint64_t)hi << 32) | lo);
                else wrow[(size_t)i * nrows] = w ^ (((uint64_t)hi << 32) | lo);
            }
            __syncwarp();
        }
    }
    if (RESIDENT)
        for (int i = lane; i < nspins; i += 32) wrow[(size_t)i * nrows] = mine[i];
}

// ------------------------------------------------------------------------------------------
// state initialisation / packing
// ------------------------------------------------------------------------------------------
// all lanes of a segment of P lanes
__device__ __forceinline__ uint64_t seg_ones(int P) { return (P >= 64) ? ~0ull : ((1ull << P) - 1ull); }

__global__ void state_init_kernel(uint64_t *words, int nspins, int nrows, int lanes, int segP, int segS,
                                  uint32_t k0, uint32_t k1, uint32_t row0, int tile)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)nrows * nspins) return;
    const uint32_t row = (uint32_t)(tid % nrows), i = (uint32_t)(tid / nrows);
    uint64_t w = 0;
    if (tile) {                                      // one bit per (replica, spin), copied to all its slices
        for (int g = 0; g < segS; g++) {
            const u32x4 r = philox4x32_10(i, PIQMC_STREAM_INIT << 16, 0u, row0 + row * (uint32_t)segS + g, k0, k1);
            if (r.x >> 31) w |= seg_ones(segP) << (g * segP);
        }
    } else {
        for (int l = 0; l < lanes; l++) {
            const u32x4 r = philox4x32_10(i, PIQMC_STREAM_INIT << 16, 0u, (row0 + row) * 64u + l, k0, k1);
            w |= (uint64_t)(r.x >> 31) << l;
        }
    }
    words[tid] = w;
}

// spins: [nrows*segS][nspins] (tile) or [nrows*segS][segP][nspins]; replica row*segS + g -> segment g
__global__ void pack_spins_kernel(uint64_t *words, const int8_t *spins, int nspins, int nrows,
                                  int segP, int segS, int tile)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)nrows * nspins) return;
    const size_t row = tid % nrows, i = tid / nrows;
    uint64_t w = 0;
    for (int g = 0; g < segS; g++) {
        const size_t rep = row * segS + g;
        if (tile) {
            if (spins[rep * nspins + i] < 0) w |= seg_ones(segP) << (g * segP);
        } else {
            for (int l = 0; l < segP; l++)
                if (spins[(rep * segP + l) * nspins + i] < 0) w |= 1ull << (g * segP + l);
        }
    }
    words[tid] = w;
}

// SA state (64 replicas per word) -> QA state (one row per replica, every slice = the replica's
// spin): the reference's np.tile(spinVector, (P,1)).T start (examples/spinglass32.py:94-96) without
// a trip through the host.
__global__ void replicas_to_slices_kernel(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst,
                                          int nspins, int src_rows, int dst_rows, int nreplicas, int segP, int segS)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)dst_rows * nspins) return;
    const size_t row = tid % dst_rows, i = tid / dst_rows;
    uint64_t w = 0;
    for (int g = 0; g < segS; g++) {
        const size_t r = row * segS + g;                  // replica; those beyond nreplicas stay +1
        if (r < (size_t)nreplicas && ((src[i * src_rows + (r >> 6)] >> (r & 63)) & 1ull)) w |= seg_ones(segP) << (g * segP);
    }
    dst[tid] = w;
}

}  // namespace

int launch_replicas_to_slices(piqmc_ctx *c, const uint64_t *d_src, int src_rows, uint64_t *d_dst, int dst_rows,
                              int nreplicas, int segP, int segS)
{
    const size_t n = (size_t)dst_rows * c->nspins;
    replicas_to_slices_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_src, d_dst, c->nspins, src_rows,
                                                                                dst_rows, nreplicas, segP, segS);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}

int launch_state_init(piqmc_ctx *c, uint64_t seed, uint32_t row0, int tile)
{
    const size_t n = (size_t)c->nrows * c->nspins;
    state_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        c->d_words, c->nspins, c->nrows, c->lanes, c->seg_P, c->seg_S, (uint32_t)seed, (uint32_t)(seed >> 32), row0,
        tile);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}

int launch_pack_spins(piqmc_ctx *c, const int8_t *d_spins, int tile)
{
    const size_t n = (size_t)c->nrows * c->nspins;
    pack_spins_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_words, d_spins, c->nspins,
                                                                         c->nrows, c->seg_P, c->seg_S, tile);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}

#define SWEEP_ARGS c->d_words, c->nspins, c->nrows, c->maxnb, c->d_idx_t, c->d_J32_t, members, nmem, \
                   c->lanes, jp2, invT, k0, k1, row0, sweep

template <int NL>
static int launch_generic(piqmc_ctx *c, int qa, int trotter, const int32_t *members, int nmem,
                          float jp2, float invT, uint64_t seed, uint32_t row0, uint32_t sweep)
{
    const size_t total = (size_t)nmem * c->nrows;
    dim3 block(128), grid((unsigned)((total + 127) / 128));
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    if (!qa)
        colour_sweep_generic<NL, false, 0><<<grid, block, 0, c->stream>>>(SWEEP_ARGS);
    else if (trotter == 1)
        colour_sweep_generic<NL, true, 1><<<grid, block, 0, c->stream>>>(SWEEP_ARGS);
    else
        colour_sweep_generic<NL, true, 0><<<grid, block, 0, c->stream>>>(SWEEP_ARGS);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}

// rows per block of the resident variant for this graph and state (0: the state does not fit)
int resident_rows_per_block(const piqmc_ctx *c, int qa)
{
    const int split = qa ? 1 : 8;
    for (int threads = 128; threads >= 32; threads >>= 1) {
        const int B = threads / split;
        const size_t smem = (size_t)c->nspins * B * 8 + (size_t)c->nspins * c->maxnb * 8;
        if (smem <= 160 * 1024) return B;
    }
    return 0;
}

template <int NL, bool QA, int TROTTER, int SPLIT>
static int launch_resident_t(piqmc_ctx *c, const int32_t *d_order, int per_sweep_orders, int nsweeps,
                             const float *d_jp2, const float *d_invT, uint64_t seed, uint32_t row0, uint32_t sweep0)
{
    const int B = resident_rows_per_block(c, QA);
    const size_t smem = (size_t)c->nspins * B * 8 + (size_t)c->nspins * c->maxnb * 8;
    auto kern = resident_sweeps<NL, QA, TROTTER, SPLIT>;
    PIQMC_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)((c->nrows + B - 1) / B), B * SPLIT, smem, c->stream>>>(
        c->d_words, c->nspins, c->nrows, c->maxnb, c->d_idx, c->d_J32, d_order, per_sweep_orders, nsweeps, d_jp2, d_invT,
        c->lanes, (uint32_t)seed, (uint32_t)(seed >> 32), row0, sweep0);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}

// the bit-sliced kernel can run this graph and mode (integer couplings; reference Trotter neighbours or SA)
bool resident_int_ok(const piqmc_ctx *c, int qa, int trotter)
{
    return c->int_unit > 0.0f && c->d_iw && !(qa && trotter) && !getenv("PIQMC_NO_INT_KERNEL");
}

// integer couplings: the bit-sliced kernel
static int launch_resident_int(piqmc_ctx *c, int qa, const int32_t *d_order, int per_sweep_orders, int nsweeps,
                               const float *d_jp2, const float *d_invT, uint64_t seed, uint32_t row0, uint32_t sweep0)
{
    int B = 128;
    while (B > 32 && (size_t)c->nspins * B * 8 + (size_t)c->nspins * c->maxnb * 5 > 160 * 1024) B >>= 1;
    const size_t smem = (size_t)c->nspins * B * 8 + (size_t)c->nspins * c->maxnb * 5;
    PIQMC_REQUIRE(smem <= 160 * 1024, PIQMC_EINVAL, "the state of one row does not fit in shared memory");
    if (qa) {
        PIQMC_CUDA(cudaFuncSetAttribute((const void *)resident_sweeps_int<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        resident_sweeps_int<true><<<(unsigned)((c->nrows + B - 1) / B), B, smem, c->stream>>>(
            c->d_words, c->nspins, c->nrows, c->maxnb, c->d_idx, c->d_iw, c->int_unit, d_order, per_sweep_orders, nsweeps,
            d_jp2, d_invT, c->lanes, (uint32_t)seed, (uint32_t)(seed >> 32), row0, sweep0);
    } else {
        PIQMC_CUDA(cudaFuncSetAttribute((const void *)resident_sweeps_int<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        resident_sweeps_int<false><<<(unsigned)((c->nrows + B - 1) / B), B, smem, c->stream>>>(
            c->d_words, c->nspins, c->nrows, c->maxnb, c->d_idx, c->d_iw, c->int_unit, d_order, per_sweep_orders, nsweeps,
            d_jp2, d_invT, c->lanes, (uint32_t)seed, (uint32_t)(seed >> 32), row0, sweep0);
    }
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}

// nsweeps sweeps, each the sequential sweep in d_order (one list, or one per sweep), state resident in shared memory
int launch_resident_sweeps(piqmc_ctx *c, int qa, int trotter, const int32_t *d_order, int per_sweep_orders, int nsweeps,
                           const float *d_jp2, const float *d_invT, uint64_t seed, uint32_t row0, uint32_t sweep0,
                           int bit_sliced)
{
    if (nsweeps <= 0) return PIQMC_OK;
    PIQMC_REQUIRE(resident_rows_per_block(c, qa) > 0, PIQMC_EINVAL, "the state of one row does not fit in shared memory");
    if (bit_sliced && resident_int_ok(c, qa, trotter))
        return launch_resident_int(c, qa, d_order, per_sweep_orders, nsweeps, d_jp2, d_invT, seed, row0, sweep0);
#define RS_ARGS c, d_order, per_sweep_orders, nsweeps, d_jp2, d_invT, seed, row0, sweep0
    if (!qa) return launch_resident_t<8, false, 0, 8>(RS_ARGS);
    if (trotter == 1) {
        if (c->lanes <= 8) return launch_resident_t<8, true, 1, 1>(RS_ARGS);
        if (c->lanes <= 16) return launch_resident_t<16, true, 1, 1>(RS_ARGS);
        if (c->lanes <= 32) return launch_resident_t<32, true, 1, 1>(RS_ARGS);
        return launch_resident_t<64, true, 1, 1>(RS_ARGS);
    }
    if (c->lanes <= 8) return launch_resident_t<8, true, 0, 1>(RS_ARGS);
    if (c->lanes <= 16) return launch_resident_t<16, true, 0, 1>(RS_ARGS);
    if (c->lanes <= 32) return launch_resident_t<32, true, 0, 1>(RS_ARGS);
    return launch_resident_t<64, true, 0, 1>(RS_ARGS);
#undef RS_ARGS
}

// as-shipped QuantumAnneal semantics (per-slice energy carry): d_order null = natural order
int launch_qa_carry(piqmc_ctx *c, const int32_t *d_order, int per_sweep_orders, int nsweeps, const float *d_jp2,
                    const float *d_invT, uint64_t seed, uint32_t row0, uint32_t sweep0)
{
    if (nsweeps <= 0) return PIQMC_OK;
    // the words of a replica in shared memory when at least one warp's worth fits (as many warps per block as fit, 8 at most)
    // (measured: with fewer than 8 warps per block the kernel is latency-bound and slower than reading through L2)
    int wpb = (int)std::min<size_t>(8, (size_t)(160 * 1024) / ((size_t)c->nspins * 8));
    if (getenv("PIQMC_CARRY_GLOBAL")) wpb = 0;
    if (wpb >= 8 || (wpb >= 1 && getenv("PIQMC_CARRY_RESIDENT"))) {
        const size_t smem = (size_t)wpb * c->nspins * 8;
        PIQMC_CUDA(cudaFuncSetAttribute((const void *)qa_carry_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        qa_carry_kernel<true><<<(unsigned)((c->nrows + wpb - 1) / wpb), wpb * 32, smem, c->stream>>>(
            c->d_words, c->nspins, c->nrows, c->maxnb, c->d_idx, c->d_J32, d_order, per_sweep_orders, nsweeps, d_jp2, d_invT,
            c->lanes, (uint32_t)seed, (uint32_t)(seed >> 32), row0, sweep0);
    } else {
        wpb = 8;
        qa_carry_kernel<false><<<(unsigned)((c->nrows + wpb - 1) / wpb), wpb * 32, 0, c->stream>>>(
            c->d_words, c->nspins, c->nrows, c->maxnb, c->d_idx, c->d_J32, d_order, per_sweep_orders, nsweeps, d_jp2, d_invT,
            c->lanes, (uint32_t)seed, (uint32_t)(seed >> 32), row0, sweep0);
    }
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}

int launch_colour_sweep(piqmc_ctx *c, int qa, int trotter, const int32_t *members, int nmem, float jp2,
                        float invT, uint64_t seed, uint32_t row0, uint32_t sweep)
{
    if (nmem == 0) return PIQMC_OK;
    if (c->lanes <= 8) return launch_generic<8>(c, qa, trotter, members, nmem, jp2, invT, seed, row0, sweep);
    if (c->lanes <= 16) return launch_generic<16>(c, qa, trotter, members, nmem, jp2, invT, seed, row0, sweep);
    if (c->lanes <= 32) return launch_generic<32>(c, qa, trotter, members, nmem, jp2, invT, seed, row0, sweep);
    return launch_generic<64>(c, qa, trotter, members, nmem, jp2, invT, seed, row0, sweep);
}

