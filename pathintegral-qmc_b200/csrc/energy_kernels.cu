// energy_kernels.cu -- sa.ClassicalIsingEnergy (piqmc/sa.pyx:25-44) as device reductions.
//
//   E = - sum_{bonds} J_ij s_i s_j - sum_i J_ii s_i          (float64)
//
// Packed path: the ELL table of tools.GenerateNeighbors lists every stored key (a, b) in both
// endpoint rows (tools.pyx:84-92), so each key is counted once by keeping only the entries whose
// neighbour index is larger than the row index; diagonal keys appear once as self-neighbours.
// Summation order is fixed (no atomics): results are reproducible.
#include "common.cuh"

namespace {

constexpr int EN_CHUNKS = 4;   // threadIdx.y sub-chunks per block

// grid = (nblocks_x, nrows); block = (64 lanes, EN_CHUNKS).  Thread (lane, y) of block bx sums
// spins i = (bx*EN_CHUNKS + y), stepping by gridDim.x*EN_CHUNKS.  All 64 lanes of a warp pair
// read the same words (broadcast), each extracting its own bit.
__global__ void __launch_bounds__(64 * EN_CHUNKS) energy_partial_kernel(
    const uint64_t *__restrict__ words, int nspins, int nrows, int maxnb,
    const int32_t *__restrict__ idx, const double *__restrict__ J, int lanes, double *__restrict__ part)
{
    const int lane = threadIdx.x, y = threadIdx.y, row = blockIdx.y;
    const uint64_t *wrow = words + row;               // word of spin s at wrow[s*nrows]
    double eq = 0.0, el = 0.0;
    for (int i = blockIdx.x * EN_CHUNKS + y; i < nspins; i += gridDim.x * EN_CHUNKS) {
        const uint64_t w = wrow[(size_t)i * nrows];
        for (int n = 0; n < maxnb; n++) {
            const int j = idx[(size_t)i * maxnb + n];
            if (j < i) continue;                              // the key is counted in row j
            const double jv = J[(size_t)i * maxnb + n];
            if (j == i) {
                el += ((w >> lane) & 1) ? -jv : jv;
            } else {
                const uint64_t x = w ^ wrow[(size_t)j * nrows];
                eq += ((x >> lane) & 1) ? -jv : jv;
            }
        }
    }
    __shared__ double sq[EN_CHUNKS][64], sl[EN_CHUNKS][64];
    sq[y][lane] = eq;
    sl[y][lane] = el;
    __syncthreads();
    if (y == 0 && lane < lanes) {
        double q = 0.0, l = 0.0;
        for (int c = 0; c < EN_CHUNKS; c++) {
            q += sq[c][lane];
            l += sl[c][lane];
        }
        // part[(row*gridDim.x + bx)*2*64 + {0,1}*64 + lane]
        double *p = part + ((size_t)row * gridDim.x + blockIdx.x) * 128;
        p[lane] = q;
        p[64 + lane] = l;
    }
}

__global__ void energy_final_kernel(const double *__restrict__ part, int nblocks, int nrows, int lanes,
                                    double *__restrict__ out)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= nrows * lanes) return;
    const int row = tid / lanes, lane = tid % lanes;
    double q = 0.0, l = 0.0;
    for (int b = 0; b < nblocks; b++) {
        const double *p = part + ((size_t)row * nblocks + b) * 128;
        q += p[lane];
        l += p[64 + lane];
    }
    out[tid] = -q - l;
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

}  // namespace

int launch_energy(piqmc_ctx *c)
{
    int nbx = (c->nspins + EN_CHUNKS * 64 - 1) / (EN_CHUNKS * 64);   // >= 64 spins per thread-column
    if (nbx < 1) nbx = 1;
    if (nbx > 64) nbx = 64;
    const size_t need = (size_t)c->nrows * nbx * 128;
    if (need > c->epart_elems) {
        if (c->d_epart) PIQMC_CUDA(cudaFree(c->d_epart));
        c->d_epart = nullptr;
        PIQMC_CUDA(cudaMalloc(&c->d_epart, need * sizeof(double)));
        c->epart_elems = need;
    }
    dim3 block(64, EN_CHUNKS), grid(nbx, c->nrows);
    energy_partial_kernel<<<grid, block, 0, c->stream>>>(c->d_words, c->nspins, c->nrows, c->maxnb,
                                                        c->d_idx, c->d_J64, c->lanes, c->d_epart);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    const int n = c->nrows * c->lanes;
    energy_final_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_epart, nbx, c->nrows, c->lanes,
                                                               c->d_energy);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}

int launch_energy_coo(piqmc_ctx *c, int nspins, int nnz, const int32_t *d_row, const int32_t *d_col,
                      const double *d_val, int nconfs, const int8_t *d_spins, double *d_out)
{
    energy_coo_kernel<<<nconfs, 256, 0, c->stream>>>(nspins, nnz, d_row, d_col, d_val, d_spins, d_out);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}
