// energy_kernels.cu -- sa.ClassicalIsingEnergy (piqmc/sa.pyx:25-44) as device reductions.
//
//   E = - sum_{bonds} J_ij s_i s_j - sum_i J_ii s_i          (float64)
//
// Packed path: the ELL table of tools.GenerateNeighbors lists every stored key (a, b) in both
// endpoint rows (tools.pyx:84-92), so each key is counted once by keeping only the entries whose
// neighbour index is larger than the row index; diagonal keys appear once as self-neighbours.
// Summation order is fixed (no atomics): results are reproducible.  Padding entries (0, 0.0) of a
// row i > 0 point at spin 0 < i and are skipped; those of row 0 add +-0.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int EN_THREADS = 256;

// Thread = (row, group of LPT lanes), LPT = 32 for states with many rows (a warp covers 16 consecutive rows: every
// word load of a warp is 128 contiguous bytes) and 8 otherwise (4 rows x 8 lane groups per warp: more threads for
// few rows).  grid = (row tiles, spin splits); block (tile, sp) sums spins sp, sp + nsplit, ... for its rows; LPT
// float64 accumulators per thread, one per lane.  Fields and bonds go into the same sum: E = -(sum).
// The entries come from the per-spin list of piqmc_set_graph (neighbour index >= own index, J != 0, table order);
// two spins are in flight per iteration (their words and the first two neighbour words of each are requested
// before the first add).  A lane accumulates D = the sum of J over its DISAGREEING entries -- one bit test and one
// predicated DADD per (entry, lane) -- next to the lane-independent total T of all J; sum_b J_b s_i s_j = T - 2 D.
// Measured on B200, 256x256 torus, 4096 rows x 64 lanes (3.4e10 adds): 13.9 / 13.2 / 12.0 ms with 8 / 16 / 32 lanes
// per thread (8.9 warp-instructions per add in the round-1 version, ~4.5 here); 512 rows: 2.6 -> 1.8 ms (fewer
// redundant loads).  A variant with 64-bit fixed-point accumulators (exact, order-independent sums) took 13.7 ms:
// ncu there showed the integer / logic pipe 65 % busy at 49 % issue and 36 % of the stalls on loads with 4 warps
// per scheduler (91 registers) -- no pipe is saturated, the loop is a mix of ALU rate and exposed load latency,
// and it did not pay to change the arithmetic.  The same float64 values for every LPT.
template <int LPT>
__device__ __forceinline__ void energy_add(double (&acc)[LPT], double &tot, double J, uint64_t x, int lg)
{
    const uint32_t bits = (LPT == 32) ? (uint32_t)(x >> (32 * lg)) : ((uint32_t)(x >> (LPT * lg)) & ((1u << LPT) - 1u));
    tot += J;
#pragma unroll
    for (int b = 0; b < LPT; b++)        // (ptxas makes this an unconditional DADD into a temporary + two selects,
        if ((bits >> b) & 1u) acc[b] += J;   //  also when the add is written as a predicated instruction)
}

template <int LPT>
__global__ void __launch_bounds__(EN_THREADS) energy_partial_kernel(
    const uint64_t *__restrict__ words, int nspins, int nrows, const int32_t *__restrict__ fb_off,
    const int32_t *__restrict__ fb_j, const double *__restrict__ fb_J, double *__restrict__ part, int row_lo, int row_hi)
{
    constexpr int GROUPS = 64 / LPT;                    // threads per row
    const int lg = threadIdx.x % GROUPS;
    const int row = row_lo + blockIdx.x * (EN_THREADS / GROUPS) + threadIdx.x / GROUPS;
    const bool live = row < row_hi;
    const uint64_t *wrow = words + (live ? row : 0);   // word of spin s at wrow[s*nrows]
    double acc[LPT], tot = 0.0;
#pragma unroll
    for (int b = 0; b < LPT; b++) acc[b] = 0.0;
    const int stride = (int)gridDim.y;
    for (int i0 = blockIdx.y; i0 < nspins; i0 += 2 * stride) {
        int sp[2], e0[2], e1[2];
        uint64_t w[2], x[2][2];
#pragma unroll
        for (int k = 0; k < 2; k++) {                   // all loads of both spins first
            sp[k] = i0 + k * stride;
            const bool on = sp[k] < nspins;
            e0[k] = on ? fb_off[sp[k]] : 0;
            e1[k] = on ? fb_off[sp[k] + 1] : 0;
            w[k] = on ? wrow[(size_t)sp[k] * nrows] : 0ull;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                x[k][t] = 0ull;
                if (e0[k] + t < e1[k]) {
                    const int j = fb_j[e0[k] + t];
                    x[k][t] = (j == sp[k]) ? 0ull : wrow[(size_t)j * nrows];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {                   // then the adds, spin by spin, entry by entry
#pragma unroll
            for (int t = 0; t < 2; t++)
                if (e0[k] + t < e1[k]) energy_add<LPT>(acc, tot, fb_J[e0[k] + t], w[k] ^ x[k][t], lg);
            for (int e = e0[k] + 2; e < e1[k]; e++) {   // spins with more than two entries
                const int j = fb_j[e];
                energy_add<LPT>(acc, tot, fb_J[e], (j == sp[k]) ? w[k] : (w[k] ^ wrow[(size_t)j * nrows]), lg);
            }
        }
    }
    if (live) {
        double *p = part + ((size_t)blockIdx.y * nrows + row) * 64 + LPT * lg;
#pragma unroll
        for (int b = 0; b < LPT; b++) p[b] = tot - 2.0 * acc[b];
    }
}

__global__ void energy_final_kernel(const double *__restrict__ part, int nsplit, int nrows, int lanes,
                                    double *__restrict__ out, int row_lo, int row_hi)
{
    const int tid = row_lo * lanes + blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= row_hi * lanes) return;
    const int row = tid / lanes, lane = tid % lanes;
    double q = 0.0;
    for (int b = 0; b < nsplit; b++) q += part[((size_t)b * nrows + row) * 64 + lane];   // fixed order
    out[tid] = -q;
}

// COO path for host configurations (drop-in ClassicalIsingEnergy): one block per configuration.
__global__ void __launch_bounds__(256) energy_coo_kernel(
    int nspins, int nnz, const int32_t *__restrict__ row, const int32_t *__restrict__ col,
    const double *__restrict__ val, const int8_t *__restrict__ spins, double *__restrict__ out)
{
    const int8_t *s = spins + (size_t)blockIdx.x * nspins;
    double acc = 0.0;
    for (int e = threadIdx.x; e < nnz; e += blockDim.x) {
        const int i = row[e], j = col[e];
        const double v = val[e];
        acc += (i == j) ? v * (double)s[i] : v * (double)(s[i] * s[j]);
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) sh[threadIdx.x] += sh[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = -sh[0];
}

// one thread per replica (or per (replica, slice) when reduce == 2): value -> bin; block-level sums first
__global__ void __launch_bounds__(256) energy_histogram_kernel(
    const double *__restrict__ en, int nreplicas, int slices, int reduce, double e0, double scale, double lo,
    double hi, int nbins, unsigned long long *__restrict__ counts, double *__restrict__ stats)
{
    const int n = reduce == 2 ? nreplicas * slices : nreplicas;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    const bool on = tid < n;
    if (on) {
        if (reduce == 2) v = en[tid];
        else {
            const double *e = en + (size_t)tid * slices;
            double acc = e[0];
            for (int k = 1; k < slices; k++) acc = reduce == 0 ? acc + e[k] : fmin(acc, e[k]);
            v = reduce == 0 ? acc / slices : acc;
        }
        v = (v - e0) * scale;
        int b;
        if (v < lo) b = nbins;
        else if (!(v < hi)) b = nbins + 1;
        else b = min(nbins - 1, (int)((v - lo) / (hi - lo) * nbins));
        atomicAdd(counts + b, 1ull);
    }
    __shared__ double ssum[256], smin[256], smax[256];
    ssum[threadIdx.x] = on ? v : 0.0;
    smin[threadIdx.x] = on ? v : 1e300;
    smax[threadIdx.x] = on ? v : -1e300;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) {
            ssum[threadIdx.x] += ssum[threadIdx.x + st];
            smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + st]);
            smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + st]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicAdd(stats, ssum[0]);
        // minimum / maximum through the order-preserving integer image of a double
        auto key = [](double x) {
            long long k = __double_as_longlong(x);
            return k >= 0 ? k : (long long)(0x8000000000000000ull - (unsigned long long)k);
        };
        atomicMin(reinterpret_cast<long long *>(stats) + 1, key(smin[0]));
        atomicMax(reinterpret_cast<long long *>(stats) + 2, key(smax[0]));
    }
}

}  // namespace

int launch_energy_histogram(piqmc_ctx *c, int reduce, double e0, double scale, double lo, double hi, int nbins,
                            unsigned long long *d_counts, double *d_stats)
{
    const int slices = c->seg_P, nrep = c->nrows * c->seg_S;
    const int n = reduce == 2 ? nrep * slices : nrep;
    energy_histogram_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_energy, nrep, slices, reduce, e0, scale, lo, hi,
                                                                   nbins, d_counts, d_stats);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}

// the split of the spin sum depends on the graph only, never on the number of rows (nor on the row range of
// a launch): a replica's energy is the same float64 whichever shard of the replicas it is computed in
static int energy_nsplit(const piqmc_ctx *c) { return std::max(1, std::min(64, (c->nspins + 63) / 64)); }

int energy_reserve(piqmc_ctx *c)
{
    const size_t need = (size_t)energy_nsplit(c) * c->nrows * 64;
    if (need > c->epart_elems) {
        if (c->d_epart) PIQMC_CUDA(cudaFree(c->d_epart));
        c->d_epart = nullptr;
        c->epart_elems = 0;
        PIQMC_CUDA(cudaMalloc(&c->d_epart, need * sizeof(double)));
        c->epart_elems = need;
    }
    return PIQMC_OK;
}

int launch_energy_rows(piqmc_ctx *c, int row_lo, int count, cudaStream_t stream)
{
    if (count <= 0) return PIQMC_OK;
    const int nsplit = energy_nsplit(c);
    // the lane grouping depends on the size of the state, never on the row range of a launch: a replica's energy
    // is the same float64 whichever rows are reduced together with it
    int lpt = c->nrows >= 256 ? 32 : 8;
    if (const char *e = getenv("PIQMC_ENERGY_LPT")) lpt = atoi(e);          // tuning knob: 8, 16 or 32 lanes per thread
    const int rows_per_block = EN_THREADS / (64 / lpt);
    dim3 grid((count + rows_per_block - 1) / rows_per_block, nsplit);
    if (lpt == 32)
        energy_partial_kernel<32><<<grid, EN_THREADS, 0, stream>>>(c->d_words, c->nspins, c->nrows, c->d_fb_off, c->d_fb_j,
                                                                  c->d_fb_J, c->d_epart, row_lo, row_lo + count);
    else if (lpt == 16)
        energy_partial_kernel<16><<<grid, EN_THREADS, 0, stream>>>(c->d_words, c->nspins, c->nrows, c->d_fb_off, c->d_fb_j,
                                                                  c->d_fb_J, c->d_epart, row_lo, row_lo + count);
    else
        energy_partial_kernel<8><<<grid, EN_THREADS, 0, stream>>>(c->d_words, c->nspins, c->nrows, c->d_fb_off, c->d_fb_j,
                                                                 c->d_fb_J, c->d_epart, row_lo, row_lo + count);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    const int n = count * c->lanes;
    energy_final_kernel<<<(n + 255) / 256, 256, 0, stream>>>(c->d_epart, nsplit, c->nrows, c->lanes, c->d_energy,
                                                            row_lo, row_lo + count);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}

int launch_energy(piqmc_ctx *c)
{
    int rc = energy_reserve(c);
    if (rc != PIQMC_OK) return rc;
    return launch_energy_rows(c, 0, c->nrows, c->stream);
}

int launch_energy_coo(piqmc_ctx *c, int nspins, int nnz, const int32_t *d_row, const int32_t *d_col,
                      const double *d_val, int nconfs, const int8_t *d_spins, double *d_out)
{
    energy_coo_kernel<<<nconfs, 256, 0, c->stream>>>(nspins, nnz, d_row, d_col, d_val, d_spins, d_out);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}
