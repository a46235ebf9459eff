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
//   fast     maxnb <= 4, reference Trotter mode: every lane's energy difference is one of
//            3 x 16 values per spin (Trotter class x neighbour pattern), so "accept by sign"
//            and "needs a uniform" are 4-input boolean functions of the disagreement masks.
//            They are evaluated for all 64 lanes at once in algebraic normal form with
//            per-spin coefficient masks from shared memory (~4 instructions per attempt); only
//            the ~1% of lanes that need a uniform take the scalar Philox path.
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

// u < thr for lane `lane` of word (row, spin, sweep), most significant bit first.  Bit planes:
// Philox block j serves planes 2j (words x,y = lanes 0..31, 32..63) and 2j+1 (words z,w).
__device__ __noinline__ bool draw_below(uint32_t thr, int lane, uint32_t spin, uint32_t sweep,
                                        uint32_t prow, uint32_t k0, uint32_t k1)
{
    const int sh = lane & 31;
    const bool hi = lane >= 32;
#pragma unroll 1
    for (uint32_t j = 0; j < 16; j++) {
        const u32x4 r = philox4x32_10(spin, j | (PIQMC_STREAM_SWEEP << 16), sweep, prow, k0, k1);
        uint32_t ub = ((hi ? r.y : r.x) >> sh) & 1u;
        uint32_t tb = (thr >> (31 - 2 * j)) & 1u;
        if (ub != tb) return tb != 0;
        ub = ((hi ? r.w : r.z) >> sh) & 1u;
        tb = (thr >> (30 - 2 * j)) & 1u;
        if (ub != tb) return tb != 0;
    }
    return false;   // u == thr
}

// ------------------------------------------------------------------------------------------
// Generic variant.  Phase 1 accumulates the in-slice sums of all lanes in registers (neighbour
// loop outside, lanes unrolled inside: per-lane order is table order, as the specification
// requires).  Phase 2 walks the lanes in order, adds the Trotter terms from the *current* word
// (so slice k >= 2 sees the new bit 1, as in the sequential slice-major order) and applies
// Metropolis.
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
                    if (x >= PIQMC_XCUT) acc = draw_below(colour_thresh(x), k, (uint32_t)i, sweep, prow, k0, k1);
                }
                if (acc) w ^= (1ull << k);
            }
        }
    }
    wrow[(size_t)i * nrows] = w;
}

// ------------------------------------------------------------------------------------------
// Fast variant (maxnb <= 4, reference Trotter mode or SA).
// ------------------------------------------------------------------------------------------
constexpr int FAST_THREADS = 128;
constexpr int FAST_ROWS = 512;      // rows of one spin handled by one block (table built once)

struct SpinTable {
    uint32_t cacc[3][16];    // ANF coefficient masks (0 / ~0) of "accept by sign", per Trotter class
    uint32_t cneed[3][16];   // ... of "needs a uniform"
    uint32_t thr[3][16];     // acceptance threshold of pattern p in class c (lazy path)
    uint32_t hacc[3], hneed[3];   // truth tables (bit p)
};

// Per-spin decision tables for this launch.  Trotter class of a lane = number of Trotter
// neighbours it disagrees with (0,1,2): tsum = -2*jp2, +0, +2*jp2.
template <bool QA>
__device__ __forceinline__ void build_table(SpinTable &tab, int i, int nspins, int maxnb,
                                            const float *__restrict__ J_t, float jp2, float invT)
{
    constexpr int NC = QA ? 3 : 1;
    if (threadIdx.x < 6) (&tab.hacc[0])[threadIdx.x] = 0u;
    __syncthreads();
    if (threadIdx.x < NC * 16) {
        const int c = threadIdx.x / 16, p = threadIdx.x % 16;
        float e = 0.0f;
        for (int n = 0; n < maxnb; n++)
            e = __fadd_rn(e, flip_sign(-2.0f * J_t[(size_t)n * nspins + i], (uint32_t)(p >> n) & 1u));
        if (QA) {
            const float tsum = (c == 0) ? -2.0f * jp2 : (c == 1 ? 0.0f : 2.0f * jp2);   // exact
            e = __fadd_rn(e, tsum);
        }
        e = __fadd_rn(e, 0.0f);
        const bool acc = QA ? (e > 0.0f) : (e >= 0.0f);
        const float x = __fmul_rn(e, invT);
        const bool need = !acc && (x >= PIQMC_XCUT);
        tab.thr[c][p] = need ? colour_thresh(x) : 0u;
        if (acc) atomicOr(&tab.hacc[c], 1u << p);
        if (need) atomicOr(&tab.hneed[c], 1u << p);
    }
    __syncthreads();
    // Moebius transform: ANF coefficient of monomial S = parity of the truth table over subsets of S
    if (threadIdx.x < NC * 32) {
        const int c = threadIdx.x / 32, f = (threadIdx.x / 16) & 1, S = threadIdx.x % 16;
        uint32_t sub = 1u;                       // bit T set <=> T is a subset of S
        if (S & 1) sub |= sub << 1;
        if (S & 2) sub |= sub << 2;
        if (S & 4) sub |= sub << 4;
        if (S & 8) sub |= sub << 8;
        const uint32_t h = f ? tab.hneed[c] : tab.hacc[c];
        const uint32_t coef = (__popc(h & sub) & 1) ? 0xFFFFFFFFu : 0u;
        if (f) tab.cneed[c][S] = coef;
        else   tab.cacc[c][S] = coef;
    }
    __syncthreads();
}

// F(x0..x3) in algebraic normal form for one 32-bit half: XOR_S coef[S] & monomial[S]
__device__ __forceinline__ uint32_t anf16(const uint32_t *__restrict__ coef, const uint32_t (&m)[16])
{
    uint32_t f = coef[0];
#pragma unroll
    for (int S = 1; S < 16; S++) f ^= coef[S] & m[S];
    return f;
}

__device__ __forceinline__ void monomials(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t (&m)[16])
{
    m[0] = 0xFFFFFFFFu;
    m[1] = x0; m[2] = x1; m[3] = x0 & x1;
    m[4] = x2; m[5] = x0 & x2; m[6] = x1 & x2; m[7] = m[3] & x2;
    m[8] = x3;
#pragma unroll
    for (int S = 1; S < 8; S++) m[8 + S] = m[S] & x3;
}

// grid: one block per (class member, chunk of FAST_ROWS rows); each thread walks its rows.
template <bool QA>
__global__ void __launch_bounds__(FAST_THREADS) colour_sweep_fast(
    uint64_t *__restrict__ words, int nspins, int nrows, int maxnb, const int32_t *__restrict__ idx_t,
    const float *__restrict__ J_t, const int32_t *__restrict__ members, int nchunks, int lanes,
    float jp2, float invT, uint32_t k0, uint32_t k1, uint32_t row0, uint32_t sweep)
{
    __shared__ SpinTable tab;
    const int i = members[blockIdx.x / nchunks];
    const int rbeg = (blockIdx.x % nchunks) * FAST_ROWS;
    const int rend = min(nrows, rbeg + FAST_ROWS);
    build_table<QA>(tab, i, nspins, maxnb, J_t, jp2, invT);

    int nb[4];
#pragma unroll
    for (int n = 0; n < 4; n++) nb[n] = (n < maxnb) ? idx_t[(size_t)n * nspins + i] : i;
    const uint64_t valid = (lanes >= 64) ? ~0ull : ((1ull << lanes) - 1ull);

    for (int row = rbeg + threadIdx.x; row < rend; row += FAST_THREADS) {
        uint64_t *wrow = words + row;
        const uint64_t w = wrow[(size_t)i * nrows];
        const uint32_t prow = row0 + (uint32_t)row;
        uint64_t x[4];
#pragma unroll
        for (int n = 0; n < 4; n++)          // self entries (local fields) and unused columns: x = w / 0
            x[n] = (n < maxnb) ? ((nb[n] == i) ? w : (w ^ wrow[(size_t)nb[n] * nrows])) : 0ull;

        uint64_t todo = valid, flips = 0, XL = 0, XR = 0;
        if (QA) {
            // reference Trotter neighbours: slices P-1 (old value for everyone; itself for lane P-1)
            // and 1 (old for lane 0, itself for lane 1, new for lanes >= 2): decide lane 1 first.
            const uint32_t bl = (uint32_t)(w >> (lanes - 1)) & 1u;
            const uint32_t br_old = (uint32_t)(w >> 1) & 1u;
            const uint32_t c1 = (lanes - 1 == 1) ? 0u : (br_old ^ bl);   // lane 1: right neighbour is itself
            const uint32_t p1 = (uint32_t)((x[0] >> 1) & 1) | (uint32_t)((x[1] >> 1) & 1) << 1 |
                                (uint32_t)((x[2] >> 1) & 1) << 2 | (uint32_t)((x[3] >> 1) & 1) << 3;
            bool f1 = (tab.hacc[c1] >> p1) & 1u;
            if (!f1 && ((tab.hneed[c1] >> p1) & 1u))
                f1 = draw_below(tab.thr[c1][p1], 1, (uint32_t)i, sweep, prow, k0, k1);
            flips = f1 ? 2ull : 0ull;
            const uint32_t br_new = br_old ^ (f1 ? 1u : 0u);
            XL = (w ^ (bl ? ~0ull : 0ull)) & ~(1ull << (lanes - 1));
            XR = (w ^ (br_new ? ~0ull : 0ull));
            XR = (XR & ~1ull) | (uint64_t)(((uint32_t)w & 1u) ^ br_old);   // lane 0 sees the old bit 1
            todo = valid & ~2ull;
        }

        uint64_t ACC = 0, NEED = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t m16[16];
            monomials((uint32_t)(x[0] >> (32 * h)), (uint32_t)(x[1] >> (32 * h)),
                      (uint32_t)(x[2] >> (32 * h)), (uint32_t)(x[3] >> (32 * h)), m16);
            uint32_t acc, need;
            if (QA) {
                const uint32_t xl = (uint32_t)(XL >> (32 * h)), xr = (uint32_t)(XR >> (32 * h));
                const uint32_t C0 = ~(xl | xr), C1 = xl ^ xr, C2 = xl & xr;
                acc = (C0 & anf16(tab.cacc[0], m16)) | (C1 & anf16(tab.cacc[1], m16)) |
                      (C2 & anf16(tab.cacc[2], m16));
                need = (C0 & anf16(tab.cneed[0], m16)) | (C1 & anf16(tab.cneed[1], m16)) |
                       (C2 & anf16(tab.cneed[2], m16));
            } else {
                acc = anf16(tab.cacc[0], m16);
                need = anf16(tab.cneed[0], m16);
            }
            ACC |= (uint64_t)acc << (32 * h);
            NEED |= (uint64_t)need << (32 * h);
        }
        ACC &= todo;
        NEED &= todo;

        // lanes that need a uniform (~1% at T << J): scalar path, exact threshold from the table
        while (NEED) {
            const int k = __ffsll((long long)NEED) - 1;
            NEED &= NEED - 1;
            const uint32_t p = (uint32_t)((x[0] >> k) & 1) | (uint32_t)((x[1] >> k) & 1) << 1 |
                               (uint32_t)((x[2] >> k) & 1) << 2 | (uint32_t)((x[3] >> k) & 1) << 3;
            const uint32_t c = QA ? (uint32_t)((XL >> k) & 1) + (uint32_t)((XR >> k) & 1) : 0u;
            if (draw_below(tab.thr[c][p], k, (uint32_t)i, sweep, prow, k0, k1)) ACC |= 1ull << k;
        }
        wrow[(size_t)i * nrows] = w ^ flips ^ ACC;
    }
}

// ------------------------------------------------------------------------------------------
// state initialisation / packing
// ------------------------------------------------------------------------------------------
__global__ void state_init_kernel(uint64_t *words, int nspins, int nrows, int lanes, uint32_t k0,
                                  uint32_t k1, uint32_t row0, int tile)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)nrows * nspins) return;
    const uint32_t row = (uint32_t)(tid % nrows), i = (uint32_t)(tid / nrows);
    uint64_t w = 0;
    if (tile) {
        const u32x4 r = philox4x32_10(i, PIQMC_STREAM_INIT << 16, 0u, row0 + row, k0, k1);
        if (r.x >> 31) w = (lanes == 64) ? ~0ull : ((1ull << lanes) - 1ull);
    } else {
        for (int l = 0; l < lanes; l++) {
            const u32x4 r = philox4x32_10(i, PIQMC_STREAM_INIT << 16, 0u, (row0 + row) * 64u + l, k0, k1);
            w |= (uint64_t)(r.x >> 31) << l;
        }
    }
    words[tid] = w;
}

__global__ void pack_spins_kernel(uint64_t *words, const int8_t *spins, int nspins, int nrows,
                                  int lanes, int tile)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)nrows * nspins) return;
    const size_t row = tid % nrows, i = tid / nrows;
    uint64_t w = 0;
    if (tile) {
        if (spins[row * nspins + i] < 0) w = (lanes == 64) ? ~0ull : ((1ull << lanes) - 1ull);
    } else {
        for (int l = 0; l < lanes; l++)
            if (spins[(row * lanes + l) * nspins + i] < 0) w |= 1ull << l;
    }
    words[tid] = w;
}

}  // namespace

int launch_state_init(piqmc_ctx *c, uint64_t seed, uint32_t row0, int tile)
{
    const size_t n = (size_t)c->nrows * c->nspins;
    state_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        c->d_words, c->nspins, c->nrows, c->lanes, (uint32_t)seed, (uint32_t)(seed >> 32), row0, tile);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}

int launch_pack_spins(piqmc_ctx *c, const int8_t *d_spins, int tile)
{
    const size_t n = (size_t)c->nrows * c->nspins;
    pack_spins_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_words, d_spins, c->nspins,
                                                                         c->nrows, c->lanes, tile);
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

// variant: 0 auto (fast when the graph and state qualify), 1 generic, 2 fast-if-possible
static bool fast_ok(const piqmc_ctx *c, int qa, int trotter)
{
    return c->variant != 1 && c->maxnb <= 4 && c->nrows >= 32 && (!qa || trotter == 0);
}

int launch_colour_sweep(piqmc_ctx *c, int qa, int trotter, const int32_t *members, int nmem, float jp2,
                        float invT, uint64_t seed, uint32_t row0, uint32_t sweep)
{
    if (nmem == 0) return PIQMC_OK;
    if (fast_ok(c, qa, trotter)) {
        const int nchunks = (c->nrows + FAST_ROWS - 1) / FAST_ROWS;
        dim3 block(FAST_THREADS), grid((unsigned)nmem * (unsigned)nchunks);
        const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#define FAST_ARGS c->d_words, c->nspins, c->nrows, c->maxnb, c->d_idx_t, c->d_J32_t, members, nchunks, \
                  c->lanes, jp2, invT, k0, k1, row0, sweep
        if (qa) colour_sweep_fast<true><<<grid, block, 0, c->stream>>>(FAST_ARGS);
        else    colour_sweep_fast<false><<<grid, block, 0, c->stream>>>(FAST_ARGS);
#undef FAST_ARGS
        c->launches++;
        PIQMC_CUDA(cudaGetLastError());
        return PIQMC_OK;
    }
    if (c->lanes <= 8) return launch_generic<8>(c, qa, trotter, members, nmem, jp2, invT, seed, row0, sweep);
    if (c->lanes <= 16) return launch_generic<16>(c, qa, trotter, members, nmem, jp2, invT, seed, row0, sweep);
    if (c->lanes <= 32) return launch_generic<32>(c, qa, trotter, members, nmem, jp2, invT, seed, row0, sweep);
    return launch_generic<64>(c, qa, trotter, members, nmem, jp2, invT, seed, row0, sweep);
}

int build_lut(piqmc_ctx *c)
{
    (void)c;
    return PIQMC_OK;
}
