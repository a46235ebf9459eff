// common.cuh -- shared declarations of libpiqmc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/piqmc_b200.h"

// ------------------------------------------------------------------------------------------
// error plumbing (never throw across the C ABI)
// ------------------------------------------------------------------------------------------
void piqmc_set_error(const char *fmt, ...);

#define PIQMC_CUDA(call)                                                                   \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            piqmc_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),       \
                            __FILE__, __LINE__);                                           \
            return PIQMC_ECUDA;                                                            \
        }                                                                                  \
    } while (0)

#define PIQMC_REQUIRE(cond, code, ...)                                                     \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            piqmc_set_error(__VA_ARGS__);                                                  \
            return (code);                                                                 \
        }                                                                                  \
    } while (0)

// One work-unit record of the dataflow kernel: everything a unit needs to know about its spin.
struct alignas(16) PiqmcUnitRec {
    int32_t spin;
    int32_t sweepoff;     // sweep lag of this member in the period-major order (0 for per-sweep lists)
    int32_t nb[4];        // neighbour spins (== nspins, the all-zero row, for self entries and unused
                          // columns), table columns sorted by |J| descending
    float J[4];           // couplings (0 for unused columns), same order
    uint8_t dep[4];       // 0: nothing to wait for; 1: neighbour of a higher level (its previous sweep);
                          // 2: neighbour of a lower level (its current sweep)
    int32_t pad;          // bits 2z..2z+1: table column of sorted entry z; bit 8+z: J[z] < 0
};
static_assert(sizeof(PiqmcUnitRec) == 48, "PiqmcUnitRec must be 48 bytes");

// ---- level kernel, staged variant (level_kernels.cu) -------------------------------------------------
// One member of a step: where its words come from.  src[k] of sorted neighbour k: bits 0..4 warp and bits
// 5..8 block (within the cluster) of the thread that wrote the neighbour's word, bits 10..11 how long ago
// (0: three or more steps, or never in this launch -- read from global memory; 1: in the step before;
// 2: two steps before -- both still in the blocks' exchange buffers), bit 12: that write belongs to the
// previous sweep (in the first sweep of a launch the word is the initial state: global memory).
struct alignas(16) PiqmcLevelRec {
    int32_t spin;
    int32_t sweepoff;     // as PiqmcUnitRec
    int32_t nb[4];        // as PiqmcUnitRec (sorted by |J|, nspins = the all-zero row)
    uint16_t src[4];
};
static_assert(sizeof(PiqmcLevelRec) == 32, "PiqmcLevelRec must be 32 bytes");

// ---- chain kernel (chain_kernels.cu) --------------------------------------------------------
// The sequential (natural-order) sweep of a 2-D lattice is cut into chains of C consecutive spins (a
// lattice row); one warp walks one chain.  Per spin (static, graph only): where each of the 4 sorted
// table columns gets its neighbour word from.
enum : uint32_t {
    PIQMC_K_ZERO = 0,      // self entry / unused column: reads as 0
    PIQMC_K_LEFT = 1,      // the spin visited one step earlier by this chain: spin i-1, or -- for the first
                           // spin of a chain that closes on itself -- the chain's last spin (its old value)
    PIQMC_K_RIGHT = 2,     // the spin visited one step later: spin i+1 (old value), or -- for the last spin
                           // of a closed chain -- the chain's first spin (its new value)
    PIQMC_K_UP = 3,        // same position in the preceding chain, this sweep; for chain 0 of a torus: in
                           // the last chain, previous sweep
    PIQMC_K_DOWN = 4,      // same position in the following chain, previous sweep; for the last chain of a
                           // torus: in chain 0, this sweep
};
struct alignas(16) PiqmcChainStat {
    float J[4];           // couplings of the table columns sorted by |J| descending (stable), 0 for unused
    uint32_t kinds;       // byte k: PIQMC_K_* of sorted slot k
    uint32_t pad;         // as PiqmcUnitRec::pad
    uint32_t spare[2];
};
static_assert(sizeof(PiqmcChainStat) == 32, "PiqmcChainStat must be 32 bytes");
// Per (schedule step, spin), written by chain_tables_kernel: a 16-byte record the sweep reads every
// step {names of "accept by sign" / "... or needs a uniform" of Trotter classes 0, 1 | class 2, flags |
// sign bytes of the sorted couplings | kinds * 16} and a 16-byte record of truth tables for the rare
// function outside the list of 27.
struct ChainGeom {
    int rpt, cw, nbands, nrings;
    size_t smem;
};

// ------------------------------------------------------------------------------------------
// device context
// ------------------------------------------------------------------------------------------
struct piqmc_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // second stream: result download overlapping the energy reduction
    cudaEvent_t copy_event = nullptr;
    uint64_t launches = 0;

    // graph (set by piqmc_set_graph)
    int nspins = 0, maxnb = 0, ncolors = 0;
    int32_t *d_idx = nullptr;       // [N][maxnb]  row-major, table order (deterministic kernels)
    float *d_J32 = nullptr;         // [N][maxnb]
    double *d_J64 = nullptr;        // [N][maxnb]  exact reference values
    // energy reduction: per spin the table entries that count (neighbour index >= own index, coupling != 0), in
    // table order: offsets [N + 1], neighbour (own index for a field), coupling
    int32_t *d_fb_off = nullptr, *d_fb_j = nullptr;
    double *d_fb_J = nullptr;
    int32_t *d_idx_t = nullptr;     // [maxnb][N]  transposed, coalesced over spins (colour kernels)
    float *d_J32_t = nullptr;       // [maxnb][N]
    int32_t *d_members = nullptr;   // spins sorted by colour (static colouring)
    std::vector<int> color_off;     // ncolors+1 offsets into d_members
    std::vector<int32_t> h_idx;     // host copies for level colourings of per-sweep orders
    std::vector<uint8_t> h_live;    // J != 0 && idx != self
    std::vector<float> h_J32;       // [N][maxnb]
    int32_t *d_level = nullptr;     // colour (level) of every spin (static colouring)
    // integer couplings (resident bit-sliced kernel): every J = iw * int_unit, |iw| <= 7, row sums <= 31
    float int_unit = 0.0f;          // 0: the graph does not qualify
    int8_t *d_iw = nullptr;         // [N][maxnb] signed weights
    PiqmcUnitRec *d_recs = nullptr; // unit records, spins sorted by (level mod D, level): period-major order
    float *d_pate = nullptr;        // [N][16] in-slice energy difference of every z-pattern, same order (build_pattern_energies)
    int flow_extra = 0;             // ceil(ncolors / D) - 1 ramp periods
    // dataflow sweep kernel (colour_fast.cu)
    uint32_t *d_done = nullptr;     // [N][flow_nchunks] tag of the last finished sweep
    unsigned int *d_ticket = nullptr;
    int flow_nchunks = 0;
    uint32_t flow_tag = 0;
    unsigned int *d_err = nullptr;  // watchdog word of the dataflow / chain kernels (0 = fine)
    // chain kernel (chain_kernels.cu): natural-order sweep cut into chains of chain_C spins
    int chain_ok = 0;               // the colouring is a level colouring of the natural order, maxnb <= 4
    int chain_force_C = 0;          // piqmc_set_chain: 0 = choose, > 0 = this chain length
    int chain_C = 0, chain_n = 0;   // chain length, chains per ring (0: no plan)
    int chain_wrap = 0;             // the last chain is coupled to chain 0 (torus)
    double chain_period = 0.0;      // pipeline steps per sweep: max(chain length, chains) on a torus
    PiqmcChainStat *d_cstat = nullptr;
    uint4 *d_chot = nullptr, *d_ccold = nullptr;   // per (schedule step, spin) records of the current launch
    size_t chot_elems = 0, ccold_elems = 0;
    int *d_cband = nullptr;         // first chain of every band
    size_t cband_elems = 0;
    uint32_t *d_cprog = nullptr;
    size_t cprog_elems = 0;
    void *d_cll = nullptr;
    size_t cll_bytes = 0;
    // level-synchronous kernel (level_kernels.cu): steps of the period-major member order
    PiqmcChainStat *d_lstat = nullptr;   // [N] sorted couplings + pad of every spin (any graph with maxnb <= 4)
    int *d_lvoff = nullptr;              // [lv_period + 1] offsets of the steps of one period into d_recs
    int lv_period = 0, lv_width = 0;     // steps per period, members of the widest step
    PiqmcLevelRec *d_xrecs = nullptr;    // staged variant: member records with word sources (null: not applicable)
    int lvx_K = 0, lvx_W = 0;            // its geometry: blocks per cluster, warps per block (one member per warp and step)
    std::vector<PiqmcUnitRec> h_recs;    // host copy of d_recs and the step offsets (streams are laid out per launch geometry)
    std::vector<int> h_lvoff;
    PiqmcLevelRec *d_stream = nullptr;   // streamed variant: [stream_slots][stream_len] records of one period
    int stream_slots = 0, stream_len = 0;

    // packed state.  QA states with at most 32 slices may hold several replicas per word: seg_S
    // segments of seg_P lanes each (lanes = seg_P * seg_S); replica of (row, segment g) = row*seg_S + g
    int nrows = 0, lanes = 0;
    int seg_P = 0, seg_S = 1;
    uint64_t *d_words = nullptr;    // [N + 1][nrows]  (row fastest); row N stays all-zero
    double *d_energy = nullptr;     // [nrows][lanes]
    void *d_stage = nullptr;        // grow-only staging buffer for host spins
    size_t stage_bytes = 0;
    double *d_epart = nullptr;      // scratch for the energy reduction
    size_t epart_elems = 0;

    int variant = 0;
    int global_moves = 0;           // QA world-line moves (fast kernel only)

    // anneal + results in one call (piqmc_qa_colour_results): the row chunks of the dataflow launch are
    // staggered and tell the host when they are final; it downloads them while the others still sweep
    int pipe_request = 0;           // set around the launch by piqmc_qa_colour_results
    int pipe_lag16 = 0;             // stagger between consecutive chunks, 1/16 ticket periods
    int pipe_armed = 0;             // launch_fast_sweeps took the request (one launch, few enough chunks)
    int pipe_chunk_cap = 0;         // entries of the two arrays below
    unsigned int *d_chunk_count = nullptr;
    unsigned int *h_chunk_flag = nullptr;   // cudaHostAlloc (mapped)
    unsigned int *d_chunk_flag = nullptr;   // its device alias
    cudaStream_t aux_stream = nullptr;      // high priority: per-chunk energy reductions next to the running sweeps
    cudaEvent_t aux_event = nullptr;
    uint64_t pipe_runs = 0;         // calls that went through the pipelined path (reported to tests / bench)
    double phase_s[2] = {0.0, 0.0};         // last piqmc_qa_colour_results: seconds in the sweeps / in energies + download
    double *pipe_energies = nullptr;        // host destinations of the call in flight
    uint64_t *pipe_words = nullptr;
};

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), as in oracle/piqmc_oracle.c part 3
// ------------------------------------------------------------------------------------------
#define PIQMC_STREAM_SWEEP 0u
#define PIQMC_STREAM_INIT 1u
#define PIQMC_STREAM_GLOBAL 2u
#define PIQMC_STREAM_CARRY 4u   // the carry (as-shipped) QA sweeps
#define PIQMC_STREAM_SA 3u      // SA sweeps: a pre-anneal and the anneal that follows may share a seed
#define PIQMC_XCUT (-22.0f)

struct u32x4 {
    uint32_t x, y, z, w;
};

__device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return u32x4{c0, c1, c2, c3};
}

// Acceptance threshold floor(exp_det(x) * 2^32), saturated; every step one IEEE float32 op.
// Must stay in lock-step with oracle_colour_thresh().
__device__ __forceinline__ uint32_t colour_thresh(float x)
{
    float t = __fmul_rn(x, 1.44269504088896341f);
    float n = rintf(t);
    float f = __fsub_rn(t, n);
    float p = 1.5403530393381609e-4f;
    p = __fmaf_rn(p, f, 1.3333558146428443e-3f);
    p = __fmaf_rn(p, f, 9.6181291076284772e-3f);
    p = __fmaf_rn(p, f, 5.5504108664821580e-2f);
    p = __fmaf_rn(p, f, 2.4022650695910071e-1f);
    p = __fmaf_rn(p, f, 6.9314718055994531e-1f);
    p = __fmaf_rn(p, f, 1.0f);
    int bits = __float_as_int(p) + (((int)n) << 23);
    float s = __fmul_rn(__int_as_float(bits), 4294967296.0f);
    return __float2uint_rz(s);   // cvt.rzi.u32.f32 saturates at 2^32-1
}

// ------------------------------------------------------------------------------------------
// kernel launchers (defined in the *_kernels.cu files)
// ------------------------------------------------------------------------------------------
int launch_qa_det(piqmc_ctx *c, const float *d_jperp, int nsched, int mcsteps, int slices, float temp,
                  int nreplicas, int8_t *d_spins, const int32_t *d_perms, piqmc_rand_state *d_rstate,
                  const double *d_uniforms, uint64_t nuniforms, unsigned long long *d_consumed,
                  const double *d_dense = nullptr, int dense_n = 0);
int launch_sa_det(piqmc_ctx *c, const float *d_temps, int nsched, int mcsteps, int nreplicas,
                  int8_t *d_spins, const int32_t *d_perms, piqmc_rand_state *d_rstate,
                  const double *d_uniforms, uint64_t nuniforms, unsigned long long *d_consumed,
                  const double *d_dense = nullptr, int dense_n = 0);
int launch_sa_multispin_det(piqmc_ctx *c, const float *d_temps, int nsched, int mcsteps, int ngroups,
                            uint64_t *d_words, const int32_t *d_perms, const double *d_rands);

int launch_state_init(piqmc_ctx *c, uint64_t seed, uint32_t row0, int tile);   // row0: first replica id
int launch_pack_spins(piqmc_ctx *c, const int8_t *d_spins, int tile);
int launch_replicas_to_slices(piqmc_ctx *c, const uint64_t *d_src, int src_rows, uint64_t *d_dst, int dst_rows,
                              int nreplicas, int segP, int segS);
// one colour class (device list `members`) of one sweep; qa != 0: QA rules (jp2 = 2*jperp)
int launch_colour_sweep(piqmc_ctx *c, int qa, int trotter, const int32_t *members, int nmem, float jp2,
                        float invT, uint64_t seed, uint32_t row0, uint32_t sweep);
// small graphs: the whole run of sweeps with the state resident in shared memory (colour_kernels.cu)
int launch_resident_sweeps(piqmc_ctx *c, int qa, int trotter, const int32_t *d_order, int per_sweep_orders, int nsweeps,
                           const float *d_jp2, const float *d_invT, uint64_t seed, uint32_t row0, uint32_t sweep0,
                           int bit_sliced);
bool resident_int_ok(const piqmc_ctx *c, int qa, int trotter);
int resident_rows_per_block(const piqmc_ctx *c, int qa);
int launch_qa_carry(piqmc_ctx *c, const int32_t *d_order, int per_sweep_orders, int nsweeps, const float *d_jp2,
                    const float *d_invT, uint64_t seed, uint32_t row0, uint32_t sweep0);
int launch_energy(piqmc_ctx *c);
// rows [row_lo, row_lo + count) only, on `stream`; the scratch must have been sized by energy_reserve
int energy_reserve(piqmc_ctx *c);
int launch_energy_rows(piqmc_ctx *c, int row_lo, int count, cudaStream_t stream);
int fast_chunk_rows(const piqmc_ctx *c);   // rows per unit of the dataflow kernel for the current state
int launch_energy_histogram(piqmc_ctx *c, int reduce, double e0, double scale, double lo, double hi, int nbins,
                            unsigned long long *d_counts, double *d_stats);
int launch_energy_coo(piqmc_ctx *c, int nspins, int nnz, const int32_t *d_row, const int32_t *d_col,
                      const double *d_val, int nconfs, const int8_t *d_spins, double *d_out);
// the production kernel: nsweeps sweeps in one dataflow launch (colour_fast.cu)
int launch_fast_sweeps(piqmc_ctx *c, int qa, int trotter, int nsweeps, const PiqmcUnitRec *d_recs, const float *d_pate,
                       int nperiods_extra, int per_sweep_lists, const float *d_jp2, const float *d_invT,
                       uint64_t seed, uint32_t row0, uint32_t sweep0);
bool launch_fast_fits(const piqmc_ctx *c, int nperiods_extra);   // the grid of one launch stays below 2^31 units
// chain kernel: nsweeps natural-order sweeps, one launch (chain_kernels.cu).  jp2/invT: host arrays per
// schedule step; sweep s of the run belongs to schedule step s / mcsteps.
int launch_chain_sweeps(piqmc_ctx *c, int qa, int nsched, int mcsteps, const float *h_jp2, const float *h_invT,
                        uint64_t seed, uint32_t row0, uint32_t sweep0);
int piqmc_check_watchdog(piqmc_ctx *c, const char *what);
// decision functions of every (schedule step, spin), once for all replicas: c->d_chot / c->d_ccold
// (chain_kernels.cu; d_jp2 / d_invT: device arrays of nf schedule steps)
int launch_decision_tables(piqmc_ctx *c, int qa, const PiqmcChainStat *d_stat, int nf, const float *d_jp2,
                           const float *d_invT, int force_generic);
// level-synchronous sweeps (level_kernels.cu)
int launch_level_sweeps(piqmc_ctx *c, int qa, int nsweeps, int mcsteps, int f_off, const float *h_jp2,
                        const float *h_invT, int nf_all, uint64_t seed, uint32_t row0, uint32_t sweep0,
                        const PiqmcUnitRec *d_recs, const int *d_step_off, const int *d_step_sweep, int period_len,
                        int nperiods_extra, int nsteps_lists, int width);
void level_geometry(const piqmc_ctx *c, int width, int *warps, int *K);
bool level_staged_geometry(int width, int *warps, int *K);   // one member per warp and step: K * warps >= width

// grow-only device buffer
template <typename T>
int piqmc_grow(T *&p, size_t &have, size_t want, cudaStream_t stream)
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
// launch geometry of the chain pipeline for the current plan and state; false: it cannot run them
bool chain_geometry(const piqmc_ctx *c, int qa, ChainGeom *g);
// variant: 0 auto, 1 generic, 2 dataflow kernel whenever the graph qualifies, 3 chain pipeline whenever there
// is a plan, else as 2, 4 level-synchronous kernel whenever the graph qualifies, else as 2, 5 resident kernel
// whenever the state of a row fits in shared memory, else as 0 (2 ... 5: used by the parity tests)
static inline bool piqmc_fast_ok(const piqmc_ctx *c, int qa, int trotter)
{
    (void)qa;
    (void)trotter;
    if (c->variant == 1 || c->maxnb > 4) return false;
    return c->variant >= 2 || c->nrows >= 32;
}
