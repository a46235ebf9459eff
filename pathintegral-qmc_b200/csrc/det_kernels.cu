// det_kernels.cu -- deterministic, reference-stream kernels (bit-exact replay).
//
// The reference's sequential algorithms cannot be parallelised inside one replica without
// changing their output: rand() is one process-global stream consumed lazily (only when
// ediff <= 0, piqmc/qmc.pyx:130-133), so which uniform attempt (slice k, step t) sees depends
// on every accept decision before it; and QuantumAnneal's ediff is a float32 carry chain over
// the whole slice sweep (piqmc/qmc.pyx:134-135).  So these kernels run ONE REPLICA PER THREAD
// and take their parallelism from the replica dimension only.  They are the parity path; the
// throughput path is colour_kernels.cu.
#include <stdlib.h>
#include <string.h>

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
template <typename U>
__device__ __forceinline__ bool lazy_accept(U &us, float ediff, float temp)
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

// ---- the same two replays with the replica on chip ------------------------------------------------------
// One warp per replica.  Spins (int8 [N][slices]), the ELL table, the sweep's visiting order and the libc
// generator state live in shared memory; lane 0 walks the sequential chain (it cannot be parallelised, see the
// header), all lanes move data in and out.  An attempt is then ~a hundred cycles of shared-memory latency and
// dependent float adds instead of ~800 cycles of dependent global loads; the table row of the NEXT attempt is
// fetched while the current one is decided (rows and orders never change, spins are read in program order).
struct GlibcRandShared {
    uint32_t *r;        // 31 words in shared memory
    int f, b;
    __device__ __forceinline__ int32_t next()
    {
        uint32_t v = (r[f] += r[b]);
        f = (f + 1 == 31) ? 0 : f + 1;
        b = (b + 1 == 31) ? 0 : b + 1;
        return (int32_t)(v >> 1);
    }
};

struct UniformsShared {
    GlibcRandShared g;
    const double *table;
    uint64_t ntable;
    unsigned long long consumed;
    __device__ __forceinline__ double next()
    {
        double u;
        if (table != nullptr)
            u = (consumed < ntable) ? table[consumed] : 2.0;
        else
            u = (double)g.next() / 2147483647.0;
        consumed++;
        return u;
    }
};

constexpr int DET_MAXNB = 8;          // table columns the on-chip kernels hold in registers (4 or 8)

__host__ __device__ inline size_t det_smem_bytes(int nspins, int slices, int mnb, bool tables)
{
    // [offsets | -2J] | order | generator (36 words) | sign bytes of nspins + 1 rows (padded to 16)
    return (tables ? (size_t)nspins * mnb * 8 : 0) + (size_t)nspins * 4 + 36 * 4 +
           (((size_t)(nspins + 1) * slices + 15) & ~(size_t)15);
}

// QA: qmc.QuantumAnneal (piqmc/qmc.pyx:76-136): sched_tab = J_perp per schedule step, temp fixed.
// !QA: sa.Anneal (piqmc/sa.pyx:80-120): sched_tab = temperature per schedule step, slices == 1.
// The ELL table arrives transformed and padded to MNB columns (det_table, below): byte offset of the
// neighbour's row of sign bytes (self entries and padding: a spare +1 row behind the last spin, so no compare
// per attempt) and -2 J (exact; padding +0, which leaves the running sum as it is).  TSM: the table is copied
// to shared memory too; otherwise it is read through L1 (larger lattices).
// LIBC: the uniforms are the libc stream (the drop-in case): the generator runs one value ahead, the Metropolis
// test of the two common outcomes (accept by sign; reject far below the cut with a non-zero integer uniform) is
// branch-free, and only the rare attempt near the cut takes the exponential.  !LIBC: uniforms from a table.
template <bool QA, bool TSM, int MNB, bool LIBC>
__global__ void __launch_bounds__(32) det_onchip_kernel(
    const float *__restrict__ sched_tab, int nsched, int mcsteps, int slices, float temp_qa, int nspins,
    const int32_t *__restrict__ toff, const uint32_t *__restrict__ tm2j, int8_t *__restrict__ spins,
    const int32_t *__restrict__ perms, piqmc_rand_state *rstate, const double *__restrict__ uniforms,
    uint64_t nuniforms, unsigned long long *consumed)
{
    // Lane 0 walks the chain and one warp's issue rate bounds it, so every instruction per attempt counts: a
    // spin is a sign byte (0x80 = -1) and the reference's term (-2 s_i)(J s_j) is one shift, one XOR on the
    // sign bit of -2 J and the float add, in table order.
    extern __shared__ __align__(16) unsigned char det_smem[];
    int32_t *off_sm = reinterpret_cast<int32_t *>(det_smem);
    uint32_t *m2J_sm = reinterpret_cast<uint32_t *>(off_sm + (TSM ? (size_t)nspins * MNB : 0));
    int32_t *ord_s = reinterpret_cast<int32_t *>(m2J_sm + (TSM ? (size_t)nspins * MNB : 0));
    uint32_t *gen_s = reinterpret_cast<uint32_t *>(ord_s + nspins);
    uint8_t *sp = reinterpret_cast<uint8_t *>(gen_s + 36);         // [nspins + 1][slices]

    const int r = blockIdx.x, lane = threadIdx.x;
    const size_t nstate = (size_t)nspins * slices;
    int8_t *s_glob = spins + (size_t)r * nstate;
    const int32_t *perm_r = perms + (size_t)r * nsched * mcsteps * nspins;
    if (TSM)
        for (int e = lane; e < nspins * MNB; e += 32) {
            off_sm[e] = toff[e];
            m2J_sm[e] = tm2j[e];
        }
    for (size_t e = lane; e < nstate; e += 32) sp[e] = s_glob[e] < 0 ? (uint8_t)0x80 : (uint8_t)0;
    for (int e = lane; e < slices; e += 32) sp[nstate + e] = 0;
    if (!uniforms && lane < 31) gen_s[lane] = rstate[r].r[lane];
    __syncwarp();

    UniformsShared us;
    us.g.r = gen_s;
    us.g.f = uniforms ? 0 : rstate[r].f;
    us.g.b = uniforms ? 0 : rstate[r].b;
    us.table = uniforms ? uniforms + (size_t)r * nuniforms : nullptr;
    us.ntable = nuniforms;
    us.consumed = 0;
    // LIBC: glibc random_r() TYPE_3 (r[f] += r[b]; result r[f] >> 1), one value ahead: knext is the next
    // rand(); gf / gb are where the value after it will come from, pf / pb where knext came from (the step is
    // undone at the end if knext was never consumed, so the state goes back exactly as the reference leaves it)
    int gf = us.g.f, gb = us.g.b, pf = 0, pb = 0;
    uint32_t knext = 0u;
    unsigned long long nconsumed = 0ull;
    if (LIBC && lane == 0) {
        pf = gf;
        pb = gb;
        const uint32_t v = (gen_s[gf] += gen_s[gb]);
        knext = v >> 1;
        gf = (gf + 1 == 31) ? 0 : gf + 1;
        gb = (gb + 1 == 31) ? 0 : gb + 1;
    }

    const int tleft = slices - 1, tright = 1;      // tidx is never assigned (qmc.pyx:83,115-117)
    for (int ifield = 0; ifield < nsched; ifield++) {
        const float par = sched_tab[ifield];        // J_perp (QA) or the temperature (SA) of this step
        const uint32_t m2jp = __float_as_uint(-2.0f * par);
        // lazy_accept's cut: below it only the integer uniform 0 can accept
        const float temp_now = QA ? temp_qa : par;
        const float cut = (temp_now > 0.0f) ? -22.01f * temp_now : -__int_as_float(0x7f800000);
        for (int step = 0; step < mcsteps; step++) {
            const int32_t *perm = perm_r + (size_t)(ifield * mcsteps + step) * nspins;
            for (int e = lane; e < nspins; e += 32) ord_s[e] = perm[e];
            __syncwarp();
            if (lane == 0) {
                for (int k = 0; k < slices; k++) {
                    float ediff = 0.0f;                          // QA: carried over the slice sweep (qmc.pyx:134-135)
                    int nsidx;
                    int4 noff[MNB / 4];
                    uint4 nm2j[MNB / 4];
                    const auto fetch = [&](int t) {              // the table row of attempt t
                        nsidx = ord_s[t];
#pragma unroll
                        for (int v = 0; v < MNB / 4; v++) {
                            if (TSM) {
                                noff[v] = reinterpret_cast<const int4 *>(off_sm)[nsidx * (MNB / 4) + v];
                                nm2j[v] = reinterpret_cast<const uint4 *>(m2J_sm)[nsidx * (MNB / 4) + v];
                            } else {
                                noff[v] = __ldg(reinterpret_cast<const int4 *>(toff) + nsidx * (MNB / 4) + v);
                                nm2j[v] = __ldg(reinterpret_cast<const uint4 *>(tm2j) + nsidx * (MNB / 4) + v);
                            }
                        }
                    };
                    fetch(0);
                    for (int t = 0; t < nspins; t++) {
                        const int sidx = nsidx;
                        int off[MNB];
                        uint32_t m2j[MNB];
#pragma unroll
                        for (int v = 0; v < MNB / 4; v++) {
                            off[4 * v] = noff[v].x; off[4 * v + 1] = noff[v].y; off[4 * v + 2] = noff[v].z; off[4 * v + 3] = noff[v].w;
                            m2j[4 * v] = nm2j[v].x; m2j[4 * v + 1] = nm2j[v].y; m2j[4 * v + 2] = nm2j[v].z; m2j[4 * v + 3] = nm2j[v].w;
                        }
                        fetch((t + 1 < nspins) ? t + 1 : t);     // the next attempt's row, while this one is decided
                        uint8_t *row = sp + sidx * slices;
                        const uint32_t own = (uint32_t)row[k] << 24;
                        if (!QA) ediff = 0.0f;                   // SA: per spin (sa.pyx:119)
                        uint32_t oth[MNB];
#pragma unroll
                        for (int n = 0; n < MNB; n++)            // all neighbour reads in flight together
                            oth[n] = (uint32_t)sp[off[n] + k] << 24;
#pragma unroll
                        for (int n = 0; n < MNB; n++)            // summed in table order (the reference's rounding)
                            ediff = __fadd_rn(ediff, __uint_as_float(m2j[n] ^ own ^ oth[n]));
                        if (QA) {
                            const uint32_t tl = (uint32_t)row[tleft] << 24, tr = (uint32_t)row[tright] << 24;
                            ediff = __fadd_rn(ediff, __uint_as_float(m2jp ^ own ^ tl));
                            ediff = __fadd_rn(ediff, __uint_as_float(m2jp ^ own ^ tr));
                        }
                        const bool bysign = QA ? (ediff > 0.0f) : (ediff >= 0.0f);           // >= (sa.pyx:114)
                        bool flip = bysign;
                        if (LIBC) {
                            // the value after knext, speculatively (committed only if knext is consumed now)
                            const uint32_t v2 = gen_s[gf] + gen_s[gb];
                            const int gf1 = (gf + 1 == 31) ? 0 : gf + 1, gb1 = (gb + 1 == 31) ? 0 : gb + 1;
                            const uint32_t kk = knext;
                            const bool consume = !bysign;
                            if (consume) gen_s[gf] = v2;
                            knext = consume ? (v2 >> 1) : knext;
                            pf = consume ? gf : pf;
                            pb = consume ? gb : pb;
                            gf = consume ? gf1 : gf;
                            gb = consume ? gb1 : gb;
                            nconsumed += consume ? 1ull : 0ull;
                            if (consume && !(ediff < cut && kk != 0u)) {                     // rare: near the cut, or k == 0
                                const double ex = exp((double)__fdiv_rn(ediff, temp_now));
                                flip = (ediff < cut) ? (ex > 0.0) : (ex > (double)(int32_t)kk / 2147483647.0);
                            }
                        } else if (!bysign) {
                            flip = lazy_accept(us, ediff, temp_now);
                        }
                        if (flip) row[k] = (uint8_t)((own >> 24) ^ 0x80u);
                    }
                }
            }
            __syncwarp();
        }
    }
    for (size_t e = lane; e < nstate; e += 32) s_glob[e] = (sp[e] & 0x80) ? (int8_t)-1 : (int8_t)1;
    if (LIBC) {
        if (lane == 0) {                      // knext was produced but not consumed: take the step back
            gen_s[pf] -= gen_s[pb];
            rstate[r].f = pf;
            rstate[r].b = pb;
        }
        __syncwarp();
        if (lane < 31) rstate[r].r[lane] = gen_s[lane];
        if (consumed && lane == 0) consumed[r] = nconsumed;
    } else {
        if (consumed && lane == 0) consumed[r] = us.consumed;
    }
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

// the replica (spins, one visiting order, generator) fits in the shared memory of one block;
// PIQMC_DET_ONCHIP=0 keeps the one-thread-per-replica kernels (testing)
static int det_mnb(const piqmc_ctx *c) { return c->maxnb <= 4 ? 4 : 8; }

static bool det_onchip_ok(const piqmc_ctx *c, int slices)
{
    if (const char *e = getenv("PIQMC_DET_ONCHIP"))
        if (atoi(e) == 0) return false;
    return c->maxnb >= 1 && c->maxnb <= DET_MAXNB && !c->h_idx.empty() &&
           det_smem_bytes(c->nspins, slices, det_mnb(c), false) <= (size_t)200 * 1024;
}

// ... and the table with it, in a third of an SM's shared memory (three replicas per SM stay resident)
static bool det_tables_fit(const piqmc_ctx *c, int slices)
{
    return det_smem_bytes(c->nspins, slices, det_mnb(c), true) <= (size_t)72 * 1024;
}

// One launch of the on-chip replay: the transformed table (see det_onchip_kernel) is built on the host from
// the graph's host copy, lives in stream-ordered memory for the launch.
template <bool QA>
static int launch_det_onchip(piqmc_ctx *c, const float *d_sched, int nsched, int mcsteps, int slices, float temp,
                             int nreplicas, int8_t *d_spins, const int32_t *d_perms, piqmc_rand_state *d_rstate,
                             const double *d_uniforms, uint64_t nuniforms, unsigned long long *d_consumed)
{
    const int N = c->nspins, mnb = det_mnb(c);
    std::vector<int32_t> off((size_t)N * mnb, N * slices);
    std::vector<uint32_t> m2j((size_t)N * mnb, 0u);
    for (int i = 0; i < N; i++)
        for (int n = 0; n < c->maxnb; n++) {
            const int j = c->h_idx[(size_t)i * c->maxnb + n];
            const float v = -2.0f * c->h_J32[(size_t)i * c->maxnb + n];
            off[(size_t)i * mnb + n] = (j == i) ? N * slices : j * slices;
            memcpy(&m2j[(size_t)i * mnb + n], &v, 4);
        }
    int32_t *d_off = nullptr;
    uint32_t *d_m2j = nullptr;
    PIQMC_CUDA(cudaMallocAsync((void **)&d_off, off.size() * 4, c->stream));
    if (cudaMallocAsync((void **)&d_m2j, m2j.size() * 4, c->stream) != cudaSuccess) {
        cudaFreeAsync(d_off, c->stream);
        piqmc_set_error("cudaMallocAsync of the replay table failed: %s", cudaGetErrorString(cudaGetLastError()));
        return PIQMC_ECUDA;
    }
    if (cudaMemcpyAsync(d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        cudaMemcpyAsync(d_m2j, m2j.data(), m2j.size() * 4, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
        cudaFreeAsync(d_off, c->stream);
        cudaFreeAsync(d_m2j, c->stream);
        piqmc_set_error("upload of the replay table failed: %s", cudaGetErrorString(cudaGetLastError()));
        return PIQMC_ECUDA;
    }
    const bool tsm = det_tables_fit(c, slices);
    const size_t smem = det_smem_bytes(N, slices, mnb, tsm);
    auto kern = d_uniforms
        ? (mnb == 4 ? (tsm ? det_onchip_kernel<QA, true, 4, false> : det_onchip_kernel<QA, false, 4, false>)
                    : (tsm ? det_onchip_kernel<QA, true, 8, false> : det_onchip_kernel<QA, false, 8, false>))
        : (mnb == 4 ? (tsm ? det_onchip_kernel<QA, true, 4, true> : det_onchip_kernel<QA, false, 4, true>)
                    : (tsm ? det_onchip_kernel<QA, true, 8, true> : det_onchip_kernel<QA, false, 8, true>));
    PIQMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nreplicas, 32, smem, c->stream>>>(d_sched, nsched, mcsteps, slices, temp, N, d_off, d_m2j, d_spins, d_perms,
                                            d_rstate, d_uniforms, nuniforms, d_consumed);
    c->launches++;
    const cudaError_t launched = cudaGetLastError();
    cudaFreeAsync(d_off, c->stream);                 // stream-ordered: behind the kernel, also when the launch failed
    cudaFreeAsync(d_m2j, c->stream);
    PIQMC_CUDA(launched);
    return PIQMC_OK;
}

int launch_qa_det(piqmc_ctx *c, const float *d_jperp, int nsched, int mcsteps, int slices, float temp,
                  int nreplicas, int8_t *d_spins, const int32_t *d_perms, piqmc_rand_state *d_rstate,
                  const double *d_uniforms, uint64_t nuniforms, unsigned long long *d_consumed,
                  const double *d_dense, int dense_n)
{
    if (!d_dense && det_onchip_ok(c, slices))
        return launch_det_onchip<true>(c, d_jperp, nsched, mcsteps, slices, temp, nreplicas, d_spins, d_perms, d_rstate,
                                       d_uniforms, nuniforms, d_consumed);
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
    if (!d_dense && det_onchip_ok(c, 1))
        return launch_det_onchip<false>(c, d_temps, nsched, mcsteps, 1, 0.0f, nreplicas, d_spins, d_perms, d_rstate,
                                        d_uniforms, nuniforms, d_consumed);
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
