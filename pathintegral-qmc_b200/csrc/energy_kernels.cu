// energy_kernels.cu -- sa.ClassicalIsingEnergy (piqmc/sa.pyx:25-44) as device reductions.
//
//   E = - sum_{bonds} J_ij s_i s_j - sum_i J_ii s_i          (float64)
//
// Packed path: the ELL table of tools.GenerateNeighbors lists every stored key (a, b) in both
// endpoint rows (tools.pyx:84-92), so each key is counted once by keeping only the entries whose
// neighbour index is larger than the row index; diagonal keys appear once as self-neighbours.
// Summation order is fixed (no atomics): results are reproducible.  Padding entries (0, 0.0) of a
// row i > 0 point at spin 0 < i and are skipped; those of row 0 add +-0.
#include "common.cuh"

namespace {

constexpr int EN_ROWS = 32;     // rows per block: 8 warps x 4 rows
constexpr int EN_THREADS = 256;

// Thread = (row, group of 8 lanes): a warp covers 4 consecutive rows x 8 lane groups, so every
// word load of a warp is one 32-byte sector (4 rows x 8 bytes, each broadcast to the 8 threads of
// its row) and the table entries (idx, J) are warp-uniform.  grid = (row tiles, spin splits);
// block (tile, sp) sums spins sp, sp + nsplit, ... for its 32 rows; 8 float64 accumulators per
// thread, one per lane.  Fields and bonds go into the same sum: E = -(sum).
__global__ void __launch_bounds__(EN_THREADS) energy_partial_kernel(
    const uint64_t *__restrict__ words, int nspins, int nrows, int maxnb,
    const int32_t *__restrict__ idx, const double *__restrict__ J, double *__restrict__ part, int row_lo, int row_hi)
{
    const int lg = threadIdx.x & 7;
    const int row = row_lo + blockIdx.x * EN_ROWS + (threadIdx.x >> 3);
    const bool live = row < row_hi;
    const uint64_t *wrow = words + (live ? row : 0);   // word of spin s at wrow[s*nrows]
    double acc[8];
#pragma unroll
    for (int b = 0; b < 8; b++) acc[b] = 0.0;
    for (int i = blockIdx.y; i < nspins; i += gridDim.y) {
        const uint64_t w = wrow[(size_t)i * nrows];
        for (int n = 0; n < maxnb; n++) {
            const int j = idx[(size_t)i * maxnb + n];
            if (j < i) continue;                              // the key is counted in row j
            const long long jb = __double_as_longlong(J[(size_t)i * maxnb + n]);
            const uint64_t x = (j == i) ? w : (w ^ wrow[(size_t)j * nrows]);
            const uint32_t bits = (uint32_t)(x >> (8 * lg)) & 0xFFu;
#pragma unroll
            for (int b = 0; b < 8; b++)      // +J where the lane agrees (s_i s_j = +1), -J where it does not
                acc[b] += __longlong_as_double(jb ^ ((long long)((bits >> b) & 1u) << 63));
        }
    }
    if (live) {
        double *p = part + ((size_t)blockIdx.y * nrows + row) * 64 + 8 * lg;
#pragma unroll
        for (int b = 0; b < 8; b++) p[b] = acc[b];
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
    const int tiles = (count + EN_ROWS - 1) / EN_ROWS;
    const int nsplit = energy_nsplit(c);
    dim3 grid(tiles, nsplit);
    energy_partial_kernel<<<grid, EN_THREADS, 0, stream>>>(c->d_words, c->nspins, c->nrows, c->maxnb, c->d_idx,
                                                          c->d_J64, c->d_epart, row_lo, row_lo + count);
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
