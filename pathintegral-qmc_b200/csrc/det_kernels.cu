// det_kernels.cu -- deterministic, reference-stream kernels (bit-exact replay).
//
// The reference's sequential algorithms cannot be parallelised inside one replica without
// changing their output: rand() is one process-global stream consumed lazily (only when
// ediff <= 0, piqmc/qmc.pyx:130-133), so which uniform attempt (slice k, step t) sees depends
// on every accept decision before it; and QuantumAnneal's ediff is a float32 carry chain over
// the whole slice sweep (piqmc/qmc.pyx:134-135).  So these kernels run ONE REPLICA PER THREAD
// and take their parallelism from the replica dimension only.  They are the parity path; the
// throughput path is colour_kernels.cu.
#include "common.cuh"

namespace {

struct GlibcRand {
    uint32_t r[31];
    int f, b;
    __device__ __forceinline__ void load(const piqmc_rand_state *s)
    {
#pragma unroll 1
        for (int i = 0; i < 31; i++) r[i] = s->r[i];
        f = s->f;
        b = s->b;
    }
    __device__ __forceinline__ void store(piqmc_rand_state *s) const
    {
#pragma unroll 1
        for (int i = 0; i < 31; i++) s->r[i] = r[i];
        s->f = f;
        s->b = b;
    }
    // glibc random_r(), TYPE_3: r[f] += r[b]; result = r[f] >> 1
    __device__ __forceinline__ int32_t next()
    {
        uint32_t v = (r[f] += r[b]);
        f = (f + 1 == 31) ? 0 : f + 1;
        b = (b + 1 == 31) ? 0 : b + 1;
        return (int32_t)(v >> 1);
    }
};

struct Uniforms {
    GlibcRand g;
    const double *table;
    uint64_t ntable;
    unsigned long long consumed;
    __device__ __forceinline__ double next()
    {
        double u;
        if (table != nullptr)
            u = (consumed < ntable) ? table[consumed] : 2.0;   // exhausted: never accept
        else
            u = (double)g.next() / 2147483647.0;                // rand()/(double)RAND_MAX
        consumed++;
        return u;
    }
};

// exp(x) > u without the exponential where it cannot matter (x <= 0 here, or any x for the multispin
// variant): below x = -22 the exponential is smaller than 2.79e-10, so every u >= 2.8e-10 rejects -- at the
// reference's T = 0.01 that is nearly every rejected attempt.  (rand()/RAND_MAX is 0 or >= 4.66e-10.)
// Exact: the decision is the one the full comparison gives.
__device__ __forceinline__ bool exp_exceeds(double x, double u)
{
    if (x < -22.0 && u >= 2.8e-10) return false;
    return exp(x) > u;
}

// The lazy Metropolis test of the sequential variants: exp((double)(ediff / temp)) > rand() / RAND_MAX with
// ediff <= 0 (piqmc/qmc.pyx:130-133, piqmc/sa.pyx:114-117), consuming one uniform.  With the libc generator
// the uniform is k / (2^31 - 1) for an integer k: far below the cut (x < -22, nearly every rejected attempt at
// the reference's T = 0.01) only k == 0 can accept, so neither the double division nor the exponential is
// evaluated -- the decision is still exactly the one the full expression gives.
__device__ __forceinline__ bool lazy_accept(Uniforms &us, float ediff, float temp)
{
    if (us.table == nullptr) {
        const int32_t k = us.g.next();
        us.consumed++;
        if (temp > 0.0f && ediff < -22.01f * temp) return k == 0 && exp((double)__fdiv_rn(ediff, temp)) > 0.0;
        return exp((double)__fdiv_rn(ediff, temp)) > (double)k / 2147483647.0;
    }
    return exp_exceeds((double)__fdiv_rn(ediff, temp), us.next());
}

// -(2*jv) with the sign of s_i*s_j applied: the exact value of (-2.0*s_i)*(jv*s_j) for
// s = +-1 (piqmc/qmc.pyx:111-113).  neg != 0 means s_i*s_j == -1.
__device__ __forceinline__ float signed_term(float jv, int neg)
{
    float t = -2.0f * jv;   // exact
    return neg ? -t : t;
}

// one dense term added to the float running sum the way the generated C does it (double add, then
// narrowing): ediff = (float)(ediff + (-2.0*s_i) * (J*s_j)); the products are exact for s = +-1.
__device__ __forceinline__ float dense_add(float ediff, double jv, int prod)
{
    const double t = __dmul_rn(prod < 0 ? 2.0 : -2.0, jv);
    return __double2float_rn(__dadd_rn((double)ediff, t));
}

// qmc.QuantumAnneal, piqmc/qmc.pyx:76-136; DENSE: qmc.QuantumAnneal_dense, piqmc/qmc.pyx:141-242
// (coupling matrix Jd[N][N] float64, upper triangle + diagonal read, couplings not narrowed).
template <bool DENSE>
__global__ void __launch_bounds__(32) qa_det_kernel(
    const float *__restrict__ jperp_tab, int nsched, int mcsteps, int slices, float temp,
    int nspins, int maxnb, const int32_t *__restrict__ idx, const float *__restrict__ J,
    const double *__restrict__ Jd,
    int nreplicas, int8_t *__restrict__ spins, const int32_t *__restrict__ perms,
    piqmc_rand_state *rstate, const double *__restrict__ uniforms, uint64_t nuniforms,
    unsigned long long *consumed)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreplicas) return;
    int8_t *s = spins + (size_t)r * nspins * slices;
    const int32_t *perm_r = perms + (size_t)r * nsched * mcsteps * nspins;
    Uniforms us;
    us.table = uniforms ? uniforms + (size_t)r * nuniforms : nullptr;
    us.ntable = nuniforms;
    us.consumed = 0;
    if (!uniforms) us.g.load(rstate + r);

    float ediff = 0.0f;
    const int tleft = slices - 1, tright = 1;      // tidx is never assigned (qmc.pyx:83,115-117)
    for (int ifield = 0; ifield < nsched; ifield++) {
        const float jperp = jperp_tab[ifield];
        for (int step = 0; step < mcsteps; step++) {
            const int32_t *perm = perm_r + (size_t)(ifield * mcsteps + step) * nspins;
            for (int k = 0; k < slices; k++) {
                for (int t = 0; t < nspins; t++) {
                    const int sidx = perm[t];
                    int8_t *row = s + (size_t)sidx * slices;
                    const int own = row[k];
                    if (DENSE) {
                        for (int si = 0; si < nspins; si++) {           // qmc.pyx:198-211
                            const int other = (si == sidx) ? 1 : (int)s[(size_t)si * slices + k];
                            const double jv = (sidx <= si) ? Jd[(size_t)sidx * nspins + si]
                                                           : Jd[(size_t)si * nspins + sidx];
                            ediff = dense_add(ediff, jv, own * other);
                        }
                    } else {
                        for (int n = 0; n < maxnb; n++) {
                            const int j = idx[(size_t)sidx * maxnb + n];
                            const float jv = J[(size_t)sidx * maxnb + n];
                            const int other = (j == sidx) ? 1 : (int)s[(size_t)j * slices + k];
                            ediff = __fadd_rn(ediff, signed_term(jv, own * other < 0));
                        }
                    }
                    ediff = __fadd_rn(ediff, signed_term(jperp, own * (int)row[tleft] < 0));
                    ediff = __fadd_rn(ediff, signed_term(jperp, own * (int)row[tright] < 0));
                    bool flip;
                    if (ediff > 0.0f) {
                        flip = true;
                    } else {
                        flip = lazy_accept(us, ediff, temp);
                    }
                    if (flip) row[k] = (int8_t)-own;
                }
                ediff = 0.0f;                       // once per slice (qmc.pyx:134-135)
            }
        }
    }
    if (!uniforms) us.g.store(rstate + r);
    if (consumed) consumed[r] = us.consumed;
}

// sa.Anneal, piqmc/sa.pyx:80-120; DENSE: sa.Anneal_dense, piqmc/sa.pyx:126-187 (accepts on
// ediff > 0, strictly, where the sparse variant accepts on >=).
template <bool DENSE>
__global__ void __launch_bounds__(32) sa_det_kernel(
    const float *__restrict__ temps, int nsched, int mcsteps, int nspins, int maxnb,
    const int32_t *__restrict__ idx, const float *__restrict__ J, const double *__restrict__ Jd, int nreplicas,
    int8_t *__restrict__ spins, const int32_t *__restrict__ perms, piqmc_rand_state *rstate,
    const double *__restrict__ uniforms, uint64_t nuniforms, unsigned long long *consumed)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreplicas) return;
    int8_t *s = spins + (size_t)r * nspins;
    const int32_t *perm_r = perms + (size_t)r * nsched * mcsteps * nspins;
    Uniforms us;
    us.table = uniforms ? uniforms + (size_t)r * nuniforms : nullptr;
    us.ntable = nuniforms;
    us.consumed = 0;
    if (!uniforms) us.g.load(rstate + r);

    for (int itemp = 0; itemp < nsched; itemp++) {
        const float temp = temps[itemp];
        for (int step = 0; step < mcsteps; step++) {
            const int32_t *perm = perm_r + (size_t)(itemp * mcsteps + step) * nspins;
            for (int t = 0; t < nspins; t++) {
                const int sidx = perm[t];
                const int own = s[sidx];
                float ediff = 0.0f;                 // per spin (sa.pyx:119)
                if (DENSE) {
                    for (int si = 0; si < nspins; si++) {               // sa.pyx:166-177
                        const int other = (si == sidx) ? 1 : (int)s[si];
                        const double jv = (sidx <= si) ? Jd[(size_t)sidx * nspins + si]
                                                       : Jd[(size_t)si * nspins + sidx];
                        ediff = dense_add(ediff, jv, own * other);
                    }
                } else {
                    for (int n = 0; n < maxnb; n++) {
                        const int j = idx[(size_t)sidx * maxnb + n];
                        const float jv = J[(size_t)sidx * maxnb + n];
                        const int other = (j == sidx) ? 1 : (int)s[j];
                        ediff = __fadd_rn(ediff, signed_term(jv, own * other < 0));
                    }
                }
                bool flip;
                if (DENSE ? (ediff > 0.0f) : (ediff >= 0.0f)) {   // >= (sa.pyx:114); dense: > (sa.pyx:180)
                    flip = true;
                } else {
                    flip = lazy_accept(us, ediff, temp);
                }
                if (flip) s[sidx] = (int8_t)-own;
            }
        }
    }
    if (!uniforms) us.g.store(rstate + r);
    if (consumed) consumed[r] = us.consumed;
}

// sa.Anneal_multispin, piqmc/sa.pyx:318-405.  One block of 64 threads per group of 64
// replicas; thread k owns replica k = bit 63-k.  Sequential over attempts.
__global__ void __launch_bounds__(64) sa_multispin_det_kernel(
    const float *__restrict__ temps, int nsched, int mcsteps, int nspins, int maxnb,
    const int32_t *__restrict__ idx, const float *__restrict__ J, uint64_t *__restrict__ words,
    const int32_t *__restrict__ perms, const double *__restrict__ rands)
{
    const int g = blockIdx.x;
    const int k = threadIdx.x;
    const int bitpos = 63 - k;
    uint64_t *sv = words + (size_t)g * nspins;
    const size_t nsweeps = (size_t)nsched * mcsteps;
    const int32_t *perm_g = perms + (size_t)g * nsweeps * nspins;
    const double *rands_g = rands + (size_t)g * nsweeps * nspins * 64;
    __shared__ uint32_t ballots[2];

    for (int itemp = 0; itemp < nsched; itemp++) {
        const double temp = (double)temps[itemp];
        for (int step = 0; step < mcsteps; step++) {
            const size_t sweep = (size_t)itemp * mcsteps + step;
            const int32_t *perm = perm_g + sweep * nspins;
            for (int t = 0; t < nspins; t++) {
                const int sidx = perm[t];
                const uint64_t w = sv[sidx];
                double ediff = 0.0;
                for (int n = 0; n < maxnb; n++) {
                    const int j = idx[(size_t)sidx * maxnb + n];
                    const double jv = (double)J[(size_t)sidx * maxnb + n];
                    const uint64_t m = (j == sidx) ? w : (w ^ sv[j]);
                    if ((m >> bitpos) & 1) ediff += 2.0 * jv;   // sa.pyx:365-368,376-379
                    else                   ediff -= 2.0 * jv;
                }
                const double u = rands_g[(sweep * nspins + t) * 64 + k];
                const bool flip = exp_exceeds(ediff / temp, u);     // no ediff>0 shortcut (sa.pyx:382)
                const uint32_t b = __ballot_sync(0xffffffffu, flip);
                if ((k & 31) == 0) ballots[k >> 5] = b;
                __syncthreads();
                if (k == 0) {
                    // thread k -> bit 63-k: warp 0 (k=0..31) fills bits 63..32, lane l at 63-l
                    uint64_t hi = __brev(ballots[0]);
                    uint64_t lo = __brev(ballots[1]);
                    sv[sidx] = w ^ ((hi << 32) | lo);
                }
                __syncthreads();
            }
        }
    }
}

}  // namespace

int launch_qa_det(piqmc_ctx *c, const float *d_jperp, int nsched, int mcsteps, int slices, float temp,
                  int nreplicas, int8_t *d_spins, const int32_t *d_perms, piqmc_rand_state *d_rstate,
                  const double *d_uniforms, uint64_t nuniforms, unsigned long long *d_consumed,
                  const double *d_dense, int dense_n)
{
    dim3 block(32), grid((nreplicas + 31) / 32);
    if (d_dense)
        qa_det_kernel<true><<<grid, block, 0, c->stream>>>(d_jperp, nsched, mcsteps, slices, temp, dense_n, 0,
                                                           nullptr, nullptr, d_dense, nreplicas, d_spins, d_perms,
                                                           d_rstate, d_uniforms, nuniforms, d_consumed);
    else
        qa_det_kernel<false><<<grid, block, 0, c->stream>>>(d_jperp, nsched, mcsteps, slices, temp, c->nspins,
                                                            c->maxnb, c->d_idx, c->d_J32, nullptr, nreplicas,
                                                            d_spins, d_perms, d_rstate, d_uniforms, nuniforms,
                                                            d_consumed);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}

int launch_sa_det(piqmc_ctx *c, const float *d_temps, int nsched, int mcsteps, int nreplicas,
                  int8_t *d_spins, const int32_t *d_perms, piqmc_rand_state *d_rstate,
                  const double *d_uniforms, uint64_t nuniforms, unsigned long long *d_consumed,
                  const double *d_dense, int dense_n)
{
    dim3 block(32), grid((nreplicas + 31) / 32);
    if (d_dense)
        sa_det_kernel<true><<<grid, block, 0, c->stream>>>(d_temps, nsched, mcsteps, dense_n, 0, nullptr, nullptr,
                                                           d_dense, nreplicas, d_spins, d_perms, d_rstate,
                                                           d_uniforms, nuniforms, d_consumed);
    else
        sa_det_kernel<false><<<grid, block, 0, c->stream>>>(d_temps, nsched, mcsteps, c->nspins, c->maxnb,
                                                            c->d_idx, c->d_J32, nullptr, nreplicas, d_spins,
                                                            d_perms, d_rstate, d_uniforms, nuniforms, d_consumed);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}

int launch_sa_multispin_det(piqmc_ctx *c, const float *d_temps, int nsched, int mcsteps, int ngroups,
                            uint64_t *d_words, const int32_t *d_perms, const double *d_rands)
{
    sa_multispin_det_kernel<<<ngroups, 64, 0, c->stream>>>(d_temps, nsched, mcsteps, c->nspins,
                                                          c->maxnb, c->d_idx, c->d_J32, d_words,
                                                          d_perms, d_rands);
    c->launches++;
    PIQMC_CUDA(cudaGetLastError());
    return PIQMC_OK;
}
