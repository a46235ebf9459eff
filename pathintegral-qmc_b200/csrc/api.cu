// api.cu -- the extern "C" surface of libpiqmc_b200.so (include/piqmc_b200.h).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <new>

#include "common.cuh"

static thread_local char g_err[512] = "";

void piqmc_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

// RAII device scratch buffer (freed on every exit path of an API call)
template <typename T>
struct DevBuf {
    T *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
};

template <typename T>
void free_dev(T *&p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

void free_graph(piqmc_ctx *c)
{
    free_dev(c->d_idx);
    free_dev(c->d_J32);
    free_dev(c->d_J64);
    free_dev(c->d_fb_off);
    free_dev(c->d_fb_j);
    free_dev(c->d_fb_J);
    free_dev(c->d_idx_t);
    free_dev(c->d_J32_t);
    free_dev(c->d_members);
    free_dev(c->d_level);
    free_dev(c->d_recs);
    free_dev(c->d_pate);
    free_dev(c->d_done);
    free_dev(c->d_cstat);
    free_dev(c->d_lstat);
    free_dev(c->d_iw);
    c->int_unit = 0.0f;
    free_dev(c->d_lvoff);
    free_dev(c->d_xrecs);
    free_dev(c->d_stream);
    c->stream_slots = c->stream_len = 0;
    c->h_recs.clear();
    c->h_lvoff.clear();
    c->lv_period = c->lv_width = 0;
    c->lvx_K = c->lvx_W = 0;
    c->chain_ok = c->chain_C = c->chain_n = 0;
    c->flow_nchunks = 0;
    c->color_off.clear();
    c->nspins = c->maxnb = c->ncolors = 0;
}

void free_state(piqmc_ctx *c)
{
    free_dev(c->d_words);
    free_dev(c->d_energy);
    free_dev(c->d_done);            // completion flags belong to a state
    c->flow_nchunks = 0;
    c->nrows = c->lanes = 0;
}

int use(piqmc_ctx *c)
{
    PIQMC_REQUIRE(c != nullptr, PIQMC_EINVAL, "null handle");
    PIQMC_CUDA(cudaSetDevice(c->device));
    return PIQMC_OK;
}

#define USE(h)                                  \
    do {                                        \
        int rc__ = use(h);                      \
        if (rc__ != PIQMC_OK) return rc__;      \
    } while (0)

#define TRY(expr)                               \
    do {                                        \
        int rc__ = (expr);                      \
        if (rc__ != PIQMC_OK) return rc__;      \
    } while (0)


// proper colouring: coupled spins (non-zero J, not self) never share a colour
static int check_colouring(int nspins, int maxnb, const int32_t *idx, const double *J, int ncolors,
                           const int32_t *color)
{
    PIQMC_REQUIRE(ncolors > 0, PIQMC_EINVAL, "ncolors must be positive");
    for (int i = 0; i < nspins; i++) {
        PIQMC_REQUIRE(color[i] >= 0 && color[i] < ncolors, PIQMC_EINVAL, "color[%d] out of range", i);
        for (int n = 0; n < maxnb; n++) {
            const int j = idx[(size_t)i * maxnb + n];
            if (j != i && J[(size_t)i * maxnb + n] != 0.0)
                PIQMC_REQUIRE(color[j] != color[i], PIQMC_EINVAL,
                              "improper colouring: spins %d and %d are coupled and share colour %d", i, j,
                              color[i]);
        }
    }
    return PIQMC_OK;
}

// bucket spins by colour (ascending spin index inside a class); offsets -> off[ncolors+1]
static void bucket_members(int nspins, int ncolors, const int32_t *color, std::vector<int> &off,
                           int32_t *members)
{
    off.assign(ncolors + 1, 0);
    for (int i = 0; i < nspins; i++) off[color[i] + 1]++;
    for (int c = 0; c < ncolors; c++) off[c + 1] += off[c];
    std::vector<int> fill(off.begin(), off.end() - 1);
    for (int i = 0; i < nspins; i++) members[fill[color[i]]++] = i;
}

// unit records (dataflow kernel) for the members listed in `order`, given the level of every spin.
// The (up to 4) table columns of a spin are stored sorted by |J| descending (stable); `pad` keeps
// the original column of every sorted entry (2 bits each) and the sign bits (J < 0) in bits 8..11:
// in the sorted, sign-normalised variables the Metropolis decision of every Trotter class is a
// *regular* monotone function -- one of only 27 -- which the kernel evaluates by name
// (colour_fast.cu).  The energies themselves are still summed in table order.
static void build_unit_recs(const piqmc_ctx *h, const int32_t *level, const int32_t *order,
                            const int32_t *sweepoff, PiqmcUnitRec *out)
{
    const int mb = h->maxnb < 4 ? h->maxnb : 4;
    for (int k = 0; k < h->nspins; k++) {
        const int i = order[k];
        PiqmcUnitRec &r = out[k];
        r.spin = i;
        r.sweepoff = sweepoff ? sweepoff[k] : 0;
        int32_t nb[4];
        float J[4];
        uint8_t dep[4];
        for (int n = 0; n < 4; n++) {
            // self entries (local fields) and unused columns read the all-zero row behind the last
            // spin (piqmc_state_alloc), so the kernel loads every column unconditionally
            nb[n] = h->nspins;
            J[n] = 0.0f;
            dep[n] = 0;
            if (n < mb) {
                const size_t e = (size_t)i * h->maxnb + n;
                if (h->h_idx[e] != i) nb[n] = h->h_idx[e];
                J[n] = h->h_J32[e];
                if (h->h_live[e]) dep[n] = level[h->h_idx[e]] < level[i] ? 2 : 1;
            }
        }
        int col[4] = {0, 1, 2, 3};
        std::stable_sort(col, col + 4, [&](int a, int b) { return fabsf(J[a]) > fabsf(J[b]); });
        int32_t pad = 0;
        for (int z = 0; z < 4; z++) {
            const int n = col[z];
            r.nb[z] = nb[n];
            r.J[z] = J[n];
            r.dep[z] = dep[n];
            pad |= n << (2 * z);
            if (J[n] < 0.0f) pad |= 1 << (8 + z);
        }
        r.pad = pad;
    }
}


// The in-slice energy difference of every z-pattern of a unit's spin, summed in TABLE order with the float32
// operations of pattern_energy (colour_device.cuh) -- it depends on the graph only, so the dataflow kernel
// reads it next to the unit record instead of recomputing it in every unit (16 selects per table column).
static void build_pattern_energies(const PiqmcUnitRec *recs, size_t n, float *out)
{
    for (size_t k = 0; k < n; k++) {
        const PiqmcUnitRec &r = recs[k];
        for (int p = 0; p < 16; p++) {
            volatile float e = 0.0f;                 // every step rounded to float32, as __fadd_rn does
            for (int col = 0; col < 4; col++)
                for (int z = 0; z < 4; z++)
                    if (((r.pad >> (2 * z)) & 3) == col) {
                        float term = -2.0f * r.J[z];
                        if (((p >> z) ^ (r.pad >> (8 + z))) & 1) term = -term;
                        e = e + term;
                    }
            out[k * 16 + p] = e;
        }
    }
}

// ---- chain plan (chain_kernels.cu) -------------------------------------------------------------
// The natural-order sweep of a 2-D lattice cut into chains of C consecutive spins.  Slot kinds of spin i
// (sorted columns, as in build_unit_recs): see PIQMC_K_* in common.cuh.  A graph qualifies when every
// coupled pair is one of: consecutive spins of a chain, the two ends of a chain (a chain that closes
// on itself), the same position of consecutive chains, or the same position of the first and the last
// chain (a torus).
struct GraphView {            // host view of an ELL table (float32 couplings, live = coupled to another spin)
    int nspins, maxnb;
    const int32_t *h_idx;
    const float *h_J32;
    const uint8_t *h_live;
};

// false: some coupled pair does not fit the pattern; *wrap: the last chain is coupled to chain 0
static bool build_chain_stat(const GraphView *h, int C, std::vector<PiqmcChainStat> &out, int *wrap)
{
    const int N = h->nspins, mb = h->maxnb < 4 ? h->maxnb : 4;
    if (C < 8 || N % C != 0) return false;
    const int nchains = N / C;
    if (nchains > 65535) return false;
    out.assign(N, PiqmcChainStat());
    *wrap = 0;
    for (int i = 0; i < N; i++) {
        int32_t nb[4];
        float J[4];
        uint8_t livef[4];
        for (int n = 0; n < 4; n++) {
            nb[n] = -1;
            J[n] = 0.0f;
            livef[n] = 0;
            if (n < mb) {
                const size_t e = (size_t)i * h->maxnb + n;
                J[n] = h->h_J32[e];
                if (h->h_live[e]) {
                    nb[n] = h->h_idx[e];
                    livef[n] = 1;
                }
            }
        }
        for (int n = 4; n < h->maxnb; n++)
            if (h->h_live[(size_t)i * h->maxnb + n]) return false;
        int col[4] = {0, 1, 2, 3};
        std::stable_sort(col, col + 4, [&](int a, int b) { return fabsf(J[a]) > fabsf(J[b]); });
        PiqmcChainStat &r = out[i];
        const int ci = i / C, pi = i % C;
        uint32_t kinds = 0, pad = 0;
        for (int z = 0; z < 4; z++) {
            const int n = col[z];
            r.J[z] = J[n];
            pad |= (uint32_t)n << (2 * z);
            if (J[n] < 0.0f) pad |= 1u << (8 + z);
            uint32_t kind = PIQMC_K_ZERO;
            if (livef[n]) {
                const int j = nb[n];
                if (j == i - 1 && pi > 0) kind = PIQMC_K_LEFT;
                else if (j == i + 1 && pi < C - 1) kind = PIQMC_K_RIGHT;
                else if (j == i + C - 1 && pi == 0) kind = PIQMC_K_LEFT;          // closed chain, first spin
                else if (j == i - (C - 1) && pi == C - 1) kind = PIQMC_K_RIGHT;   // closed chain, last spin
                else if (j == i - C) kind = PIQMC_K_UP;
                else if (j == i + C) kind = PIQMC_K_DOWN;
                else if (j == i + (N - C) && ci == 0) {
                    kind = PIQMC_K_UP;
                    *wrap = 1;
                } else if (j == i - (N - C) && ci == nchains - 1) {
                    kind = PIQMC_K_DOWN;
                    *wrap = 1;
                } else return false;
            }
            kinds |= kind << (8 * z);
        }
        r.kinds = kinds;
        r.pad = pad;
        r.spare[0] = r.spare[1] = 0;
    }
    return true;
}

// Choose the chain length (or take the forced one) and build the static records.  Candidates: the
// most frequent index distances of coupled pairs (a lattice row).
static int choose_chain(const GraphView *h, int force_C, std::vector<PiqmcChainStat> &best, double &best_period,
                        int *wrap)
{
    const int N = h->nspins;
    std::vector<int> cand;
    if (force_C > 0) cand.push_back(std::min(force_C, N));
    else {
        std::vector<std::pair<int, int>> hist;                 // (distance, count)
        std::vector<int> d;
        for (int i = 0; i < N; i++)
            for (int n = 0; n < h->maxnb; n++) {
                const size_t e = (size_t)i * h->maxnb + n;
                if (h->h_live[e] && h->h_idx[e] > i + 1) d.push_back(h->h_idx[e] - i);
            }
        std::sort(d.begin(), d.end());
        for (size_t k = 0; k < d.size();) {
            size_t m = k;
            while (m < d.size() && d[m] == d[k]) m++;
            hist.push_back({d[k], (int)(m - k)});
            k = m;
        }
        std::sort(hist.begin(), hist.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) {
            return a.second != b.second ? a.second > b.second : a.first < b.first;
        });
        for (size_t k = 0; k < hist.size() && cand.size() < 3; k++) cand.push_back(hist[k].first);
        cand.push_back(N);
    }
    best_period = 0.0;
    for (int C : cand) {
        std::vector<PiqmcChainStat> st;
        int w = 0;
        if (!build_chain_stat(h, C, st, &w)) continue;
        best.swap(st);
        *wrap = w;
        best_period = (double)std::max(C, w ? N / C : 1);      // steps per sweep of the rigid ring
        return C;
    }
    return 0;
}

static int build_chain_plan(piqmc_ctx *h)
{
    h->chain_C = h->chain_n = h->chain_wrap = 0;
    if (!h->chain_ok) return PIQMC_OK;
    const int N = h->nspins;
    const GraphView gv = {N, h->maxnb, h->h_idx.data(), h->h_J32.data(), h->h_live.data()};
    std::vector<PiqmcChainStat> best;
    double best_period = 0.0;
    int wrap = 0;
    const int best_C = choose_chain(&gv, h->chain_force_C, best, best_period, &wrap);
    if (best_C == 0) return PIQMC_OK;
    if (!h->d_cstat) PIQMC_CUDA(cudaMalloc(&h->d_cstat, (size_t)N * sizeof(PiqmcChainStat)));
    PIQMC_CUDA(cudaMemcpy(h->d_cstat, best.data(), (size_t)N * sizeof(PiqmcChainStat), cudaMemcpyHostToDevice));
    h->chain_C = best_C;
    h->chain_n = N / best_C;
    h->chain_wrap = wrap;
    h->chain_period = best_period;
    return PIQMC_OK;
}

// Which kernel runs a static colouring.  The chain pipeline is opt-in (variant 3, piqmc_set_chain, or
// PIQMC_CHAIN=1): measured on B200 it equals the dataflow kernel when few rows make the sweep
// latency-bound and loses when many rows make it throughput-bound.
static bool chain_selected(const piqmc_ctx *h, int qa, int trotter)
{
    if (h->chain_C < 8 || h->variant == 1 || h->variant == 2 || h->global_moves || (qa && trotter)) return false;
    if (h->variant != 3 && h->chain_force_C == 0) {
        const char *e = getenv("PIQMC_CHAIN");
        if (!e || atoi(e) == 0) return false;
    }
    if (h->d_words == nullptr) return true;                    // no state yet: the plan exists
    ChainGeom g;
    return chain_geometry(h, qa, &g);
}

// the sorted couplings and pad of every spin (level-synchronous kernel: decision tables, thresholds)
static int upload_level_stat(piqmc_ctx *h)
{
    free_dev(h->d_lstat);
    if (h->maxnb > 4) return PIQMC_OK;
    std::vector<int32_t> zero(h->nspins, 0), ident(h->nspins);
    for (int i = 0; i < h->nspins; i++) ident[i] = i;
    std::vector<PiqmcUnitRec> recs(h->nspins);
    build_unit_recs(h, zero.data(), ident.data(), nullptr, recs.data());
    std::vector<PiqmcChainStat> st(h->nspins);
    for (int i = 0; i < h->nspins; i++) {
        for (int z = 0; z < 4; z++) st[i].J[z] = recs[i].J[z];
        st[i].kinds = 0;
        st[i].pad = (uint32_t)recs[i].pad;
        st[i].spare[0] = st[i].spare[1] = 0;
    }
    PIQMC_CUDA(cudaMalloc(&h->d_lstat, st.size() * sizeof(PiqmcChainStat)));
    PIQMC_CUDA(cudaMemcpy(h->d_lstat, st.data(), st.size() * sizeof(PiqmcChainStat), cudaMemcpyHostToDevice));
    return PIQMC_OK;
}

// The level-synchronous kernel runs any level colouring of a graph with maxnb <= 4 (reference Trotter
// neighbours, no world-line moves).  Opt-in for now: variant 4 or PIQMC_LEVEL=1.
static bool level_selected(const piqmc_ctx *h, int qa, int trotter)
{
    if (h->maxnb > 4 || !h->d_lstat || h->global_moves || (qa && trotter)) return false;
    if (h->variant == 4) return true;
    if (h->variant != 0) return false;
    const char *e = getenv("PIQMC_LEVEL");
    return e && atoi(e) != 0;
}

// The resident kernel runs any graph whose per-row state fits in shared memory (one replica per word, no
// world-line moves).  Chosen for graphs the table kernels cannot run (maxnb > 4) and for tiny graphs, whose
// colour classes are too small for anything that synchronises between classes.
static bool resident_selected(const piqmc_ctx *h, int qa)
{
    if (h->global_moves || h->seg_S != 1 || !h->d_words || resident_rows_per_block(h, qa) == 0) return false;
    if (h->variant == 5) return true;
    if (h->variant != 0) return false;
    if (const char *e = getenv("PIQMC_RESIDENT")) return atoi(e) != 0;
    return h->maxnb > 4 || h->nspins <= 64;
}

static int apply_colouring(piqmc_ctx *h, int ncolors, const int32_t *color, bool validate)
{
    if (validate) {
        PIQMC_REQUIRE(ncolors > 0 && color, PIQMC_EINVAL, "bad colouring");
        for (int i = 0; i < h->nspins; i++) {
            PIQMC_REQUIRE(color[i] >= 0 && color[i] < ncolors, PIQMC_EINVAL, "color[%d] out of range", i);
            for (int n = 0; n < h->maxnb; n++) {
                const size_t e = (size_t)i * h->maxnb + n;
                if (h->h_live[e])
                    PIQMC_REQUIRE(color[h->h_idx[e]] != color[i], PIQMC_EINVAL,
                                  "improper colouring: spins %d and %d are coupled and share colour %d", i,
                                  h->h_idx[e], color[i]);
            }
        }
    }
    std::vector<int32_t> members(h->nspins);
    bucket_members(h->nspins, ncolors, color, h->color_off, members.data());
    PIQMC_CUDA(cudaStreamSynchronize(h->stream));     // the old lists may still be in use
    PIQMC_CUDA(cudaMemcpy(h->d_members, members.data(), (size_t)h->nspins * sizeof(int32_t),
                          cudaMemcpyHostToDevice));
    PIQMC_CUDA(cudaMemcpy(h->d_level, color, (size_t)h->nspins * sizeof(int32_t), cudaMemcpyHostToDevice));
    // period-major order for the dataflow kernel: D = 1 + largest level gap across a coupled pair
    int D = 1;
    for (int i = 0; i < h->nspins; i++)
        for (int n = 0; n < h->maxnb; n++) {
            const size_t e = (size_t)i * h->maxnb + n;
            if (h->h_live[e]) D = std::max(D, std::abs(color[i] - color[h->h_idx[e]]) + 1);
        }
    std::vector<int32_t> order(h->nspins), pm(h->nspins), po(h->nspins);
    for (int i = 0; i < h->nspins; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        const int rx = color[x] % D, ry = color[y] % D;
        return rx != ry ? rx < ry : color[x] < color[y];
    });
    for (int k = 0; k < h->nspins; k++) {
        pm[k] = order[k];
        po[k] = color[order[k]] / D;
    }
    if (h->maxnb <= 4) {
        std::vector<PiqmcUnitRec> recs(h->nspins);
        build_unit_recs(h, color, pm.data(), po.data(), recs.data());
        PIQMC_CUDA(cudaMemcpy(h->d_recs, recs.data(), recs.size() * sizeof(PiqmcUnitRec), cudaMemcpyHostToDevice));
        std::vector<float> pate(recs.size() * 16);
        build_pattern_energies(recs.data(), recs.size(), pate.data());
        PIQMC_CUDA(cudaMemcpy(h->d_pate, pate.data(), pate.size() * sizeof(float), cudaMemcpyHostToDevice));
        // level-synchronous kernel: where the steps of a period (members with the same level mod D) begin
        // in the period-major list
        const int Dp = std::min(D, ncolors);
        std::vector<int> off(Dp + 1, 0);
        for (int i = 0; i < h->nspins; i++) off[color[i] % D + 1]++;
        int width = 0;
        for (int r = 0; r < Dp; r++) {
            width = std::max(width, off[r + 1]);
            off[r + 1] += off[r];
        }
        free_dev(h->d_lvoff);
        PIQMC_CUDA(cudaMalloc(&h->d_lvoff, off.size() * sizeof(int)));
        PIQMC_CUDA(cudaMemcpy(h->d_lvoff, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
        h->lv_period = Dp;
        h->lv_width = width;
        h->h_recs = recs;
        h->h_lvoff = off;
        free_dev(h->d_stream);                                   // laid out at the next launch
        h->stream_slots = h->stream_len = 0;
        // staged variant: one member per warp and step; where every neighbour word was written
        free_dev(h->d_xrecs);
        h->lvx_K = h->lvx_W = 0;
        int xw = 0, xk = 0;
        if (Dp >= 3 && level_staged_geometry(width, &xw, &xk)) {
            std::vector<int> pos(h->nspins);                      // position of a spin within its step
            for (int k = 0; k < h->nspins; k++) pos[pm[k]] = k - off[color[pm[k]] % D];
            std::vector<PiqmcLevelRec> xr(h->nspins);
            for (int k = 0; k < h->nspins; k++) {
                const PiqmcUnitRec &r = recs[k];
                PiqmcLevelRec &x = xr[k];
                x.spin = r.spin;
                x.sweepoff = r.sweepoff;
                for (int z = 0; z < 4; z++) {
                    x.nb[z] = r.nb[z];
                    uint16_t src = 0;
                    if (r.dep[z] != 0) {
                        const int j = r.nb[z];
                        const int gap = r.dep[z] == 2 ? color[r.spin] - color[j] : Dp - (color[j] - color[r.spin]);
                        if (gap == 1 || gap == 2)
                            src = (uint16_t)((pos[j] % xw) | ((pos[j] / xw) << 5) | (gap << 10) | (r.dep[z] == 1 ? 1 << 12 : 0));
                    }
                    x.src[z] = src;
                }
            }
            PIQMC_CUDA(cudaMalloc(&h->d_xrecs, xr.size() * sizeof(PiqmcLevelRec)));
            PIQMC_CUDA(cudaMemcpy(h->d_xrecs, xr.data(), xr.size() * sizeof(PiqmcLevelRec), cudaMemcpyHostToDevice));
            h->lvx_K = xk;
            h->lvx_W = xw;
        }
    }
    h->flow_extra = (ncolors + D - 1) / D - 1;
    h->ncolors = ncolors;
    // the chain kernel runs the natural order: usable when that order has this level colouring
    h->chain_ok = h->maxnb <= 4 ? 1 : 0;
    for (int i = 0; i < h->nspins && h->chain_ok; i++)
        for (int n = 0; n < h->maxnb; n++) {
            const size_t e = (size_t)i * h->maxnb + n;
            if (h->h_live[e] && h->h_idx[e] > i && color[h->h_idx[e]] <= color[i]) {
                h->chain_ok = 0;
                break;
            }
        }
    return build_chain_plan(h);
}

// level[i] = 1 + max(level of coupled neighbours visited earlier); returns the number of levels
static int order_levels(int nspins, int maxnb, const int32_t *idx, const uint8_t *live,
                        const int32_t *order, int32_t *level)
{
    for (int i = 0; i < nspins; i++) level[i] = -1;
    int nlev = 0;
    for (int t = 0; t < nspins; t++) {
        const int i = order ? order[t] : t;
        int m = -1;
        for (int n = 0; n < maxnb; n++) {
            const size_t e = (size_t)i * maxnb + n;
            if (live[e]) {
                const int l = level[idx[e]];
                if (l > m) m = l;
            }
        }
        level[i] = m + 1;
        if (m + 2 > nlev) nlev = m + 2;
    }
    return nlev;
}

static int check_orders(int nspins, size_t nsweeps, const int32_t *orders)
{
    std::vector<uint8_t> seen(nspins);
    for (size_t s = 0; s < nsweeps; s++) {
        std::fill(seen.begin(), seen.end(), 0);
        for (int t = 0; t < nspins; t++) {
            const int i = orders[s * nspins + t];
            PIQMC_REQUIRE(i >= 0 && i < nspins && !seen[i], PIQMC_EINVAL,
                          "orders[%zu] is not a permutation of 0..%d", s, nspins - 1);
            seen[i] = 1;
        }
    }
    return PIQMC_OK;
}

// Sweeps driver shared by QA and SA.  value[f] is jp2 (QA) or unused; invT[f] per schedule step.
// ---- anneal + results in one call: the host side of the staggered launch ------------------------------
// Streams, the per-chunk counters and the mapped flags; everything is allocated BEFORE the sweep kernel is
// launched (an allocation would wait for it).
static int pipe_prepare(piqmc_ctx *h, int nchunks)
{
    if (!h->copy_stream) PIQMC_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->copy_event) PIQMC_CUDA(cudaEventCreateWithFlags(&h->copy_event, cudaEventDisableTiming));
    if (!h->aux_stream) {
        int lo = 0, hi = 0;
        PIQMC_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // hi = numerically lowest = greatest priority
        PIQMC_CUDA(cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, hi));
    }
    if (!h->aux_event) PIQMC_CUDA(cudaEventCreateWithFlags(&h->aux_event, cudaEventDisableTiming));
    if (nchunks > h->pipe_chunk_cap) {
        PIQMC_CUDA(cudaStreamSynchronize(h->stream));
        free_dev(h->d_chunk_count);
        if (h->h_chunk_flag) cudaFreeHost(h->h_chunk_flag);
        h->h_chunk_flag = h->d_chunk_flag = nullptr;
        h->pipe_chunk_cap = 0;
        const int cap = std::max(64, nchunks);
        PIQMC_CUDA(cudaMalloc(&h->d_chunk_count, cap * sizeof(unsigned int)));
        PIQMC_CUDA(cudaHostAlloc((void **)&h->h_chunk_flag, cap * sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable));
        PIQMC_CUDA(cudaHostGetDevicePointer((void **)&h->d_chunk_flag, h->h_chunk_flag, 0));
        h->pipe_chunk_cap = cap;
    }
    PIQMC_CUDA(cudaMemsetAsync(h->d_chunk_count, 0, nchunks * sizeof(unsigned int), h->stream));
    for (int c = 0; c < nchunks; c++) h->h_chunk_flag[c] = 0u;
    __sync_synchronize();
    return energy_reserve(h);
}

// Stagger between consecutive row chunks: they should finish at the rate the host link takes them (one chunk's
// download per stagger), but the ramps at both ends of the run stay a fraction of it.  Measured on B200
// (profiles/r2_e2e_overlap.md): a ticket period with few active rows is bound by its dependency chain and costs
// as much as the download it hides (256x256 torus: 1.8 ms per period with one 512-row chunk active, 3.6 ms with
// all 4096 rows); with 16384 rows the ramps are cheap (sweeps 280 -> 299 ms at lag16 = 5) but the chunk-wise 2-D
// copies next to the running kernel reach 31 GB/s against 57 GB/s for the one contiguous copy, and the call ends
// later than the plain sequence (504 ms against 458 ms).  So the stagger is opt-in: PIQMC_PIPE_LAG16 (tests,
// experiments); without it this returns 0 and the call runs the plain sequence.  The model below is what a
// layout with contiguous chunks would use.
static int pipe_choose_lag16(const piqmc_ctx *h, size_t nsweeps, int rpb, int nchunks)
{
    if (const char *e = getenv("PIQMC_PIPE_LAG16")) return std::max(0, atoi(e));
    if (nchunks < 2 || !getenv("PIQMC_PIPE_AUTO")) return 0;
    const double t_copy = (double)rpb * h->nspins * 8.0 / 43e9;                         // one chunk over the host link, GPU busy
    const double levels = (double)std::max(1, h->ncolors) / (double)(h->flow_extra + 1);
    const double t_period = std::max((double)h->nrows * h->nspins * 64.0 / 4.5e12,     // issue-bound sweep ...
                                     levels * 3.9e-6);                                 // ... or its dependency chain
    double lag = t_copy / t_period;
    lag = std::min(lag, 0.5 * (double)nsweeps / (double)(nchunks - 1));
    return (int)(lag * 16.0 + 0.5);
}

// Called between the launch of the staggered sweeps and the stream synchronisation: takes the chunks in the
// order they finish, downloads each (copy engine) and reduces its energies (high-priority stream) while the
// rest of the launch is still running.
static int pipe_drain(piqmc_ctx *h)
{
    const int rpb = fast_chunk_rows(h);
    const int nchunks = (h->nrows + rpb - 1) / rpb;
    volatile unsigned int *flag = h->h_chunk_flag;
    bool kernel_over = false;
    // PIQMC_PIPE_TRACE=1: when each chunk was seen final, and when its download and its energy reduction ran
    // (milliseconds since the drain began), on stderr.  The energies are reduced once, behind the sweeps and under
    // the tail of the downloads (PIQMC_PIPE_ENERGY=chunk: per chunk on a high-priority stream instead -- measured:
    // those launches only start when the queued downloads have drained, profiles/r2_e2e_overlap.md)
    const bool trace = getenv("PIQMC_PIPE_TRACE") != nullptr && nchunks <= 64;
    const char *een = getenv("PIQMC_PIPE_ENERGY");
    const bool energy_at_end = !(een != nullptr && een[0] == 'c');
    struct Events {                                  // released on every exit path
        std::vector<cudaEvent_t> v;
        ~Events() { for (auto &e : v) if (e) cudaEventDestroy(e); }
    } evs;
    std::vector<cudaEvent_t> &ev = evs.v;
    std::vector<double> seen(nchunks, 0.0);
    struct timespec ts0;
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    auto now_ms = [&]() {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return 1e3 * (double)(ts.tv_sec - ts0.tv_sec) + 1e-6 * (double)(ts.tv_nsec - ts0.tv_nsec);
    };
    if (trace) {
        ev.assign(1 + 4 * (size_t)nchunks, nullptr);
        for (auto &e : ev) PIQMC_CUDA(cudaEventCreate(&e));
        PIQMC_CUDA(cudaEventRecord(ev[0], h->copy_stream));
    }
    for (int c = 0; c < nchunks; c++) {
        unsigned int spins = 0;
        while (!kernel_over && flag[c] == 0u) {
            if ((++spins & 0x3FFu) == 0u) {
                const cudaError_t q = cudaStreamQuery(h->stream);
                if (q == cudaSuccess) kernel_over = true;          // every flag has been written (or the watchdog fired)
                else if (q != cudaErrorNotReady) {
                    piqmc_set_error("dataflow sweeps failed: %s", cudaGetErrorString(q));
                    return PIQMC_ECUDA;
                }
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        __sync_synchronize();
        if (trace) seen[c] = now_ms();
        const int row_lo = c * rpb, cnt = std::min(rpb, h->nrows - row_lo);
        if (trace) PIQMC_CUDA(cudaEventRecord(ev[1 + 4 * c], h->copy_stream));
        PIQMC_CUDA(cudaMemcpy2DAsync(h->pipe_words + row_lo, (size_t)h->nrows * sizeof(uint64_t), h->d_words + row_lo,
                                     (size_t)h->nrows * sizeof(uint64_t), (size_t)cnt * sizeof(uint64_t), h->nspins,
                                     cudaMemcpyDeviceToHost, h->copy_stream));
        if (trace) PIQMC_CUDA(cudaEventRecord(ev[2 + 4 * c], h->copy_stream));
        if (energy_at_end) continue;
        if (trace) PIQMC_CUDA(cudaEventRecord(ev[3 + 4 * c], h->aux_stream));
        TRY(launch_energy_rows(h, row_lo, cnt, h->aux_stream));
        if (trace) PIQMC_CUDA(cudaEventRecord(ev[4 + 4 * c], h->aux_stream));
    }
    PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    const double t_sweeps = now_ms();
    cudaStream_t es = energy_at_end ? h->stream : h->aux_stream;
    if (energy_at_end) TRY(launch_energy_rows(h, 0, h->nrows, es));
    // one download of all energies (small) behind the last reduction
    PIQMC_CUDA(cudaMemcpyAsync(h->pipe_energies, h->d_energy, (size_t)h->nrows * h->lanes * sizeof(double),
                               cudaMemcpyDeviceToHost, es));
    PIQMC_CUDA(cudaStreamSynchronize(es));
    const double t_energy = now_ms();
    PIQMC_CUDA(cudaStreamSynchronize(h->copy_stream));
    if (trace) {
        fprintf(stderr, "[piqmc pipe] %d chunks of %d rows, lag16 %d: sweeps over at %.2f ms, energies at %.2f, downloads at %.2f\n",
                nchunks, rpb, h->pipe_lag16, t_sweeps, t_energy, now_ms());
        for (int c = 0; c < nchunks; c++) {
            float a = 0, b = 0, e0 = 0, e1 = 0;
            cudaEventElapsedTime(&a, ev[0], ev[1 + 4 * c]);
            cudaEventElapsedTime(&b, ev[0], ev[2 + 4 * c]);
            if (!energy_at_end) {
                cudaEventElapsedTime(&e0, ev[0], ev[3 + 4 * c]);
                cudaEventElapsedTime(&e1, ev[0], ev[4 + 4 * c]);
            }
            fprintf(stderr, "[piqmc pipe]   chunk %d final at %.2f ms; download %.2f -> %.2f; energy %.2f -> %.2f\n", c,
                    seen[c], a, b, e0, e1);
        }
    }
    h->pipe_runs++;
    return PIQMC_OK;
}

static int run_colour_sweeps(piqmc_ctx *h, int qa, int trotter, int nsched, int mcsteps,
                             const std::vector<float> &jp2, const std::vector<float> &invT, uint64_t seed,
                             uint32_t row0, uint32_t sweep0, const int32_t *orders)
{
    const int N = h->nspins;
    const size_t nsweeps = (size_t)nsched * mcsteps;
    if (nsweeps == 0) return PIQMC_OK;
    // static colourings with many levels and a small level gap (a path graph in natural order) would need
    // more units than one grid holds: those run class by class (or through the chain pipeline below)
    const bool resident = resident_selected(h, qa);
    const bool use_level = !resident && level_selected(h, qa, trotter);
    const bool fast = !resident && !use_level && piqmc_fast_ok(h, qa, trotter) && (orders || launch_fast_fits(h, h->flow_extra));
    if (!orders) PIQMC_REQUIRE(h->ncolors > 0, PIQMC_ENOGRAPH, "graph has no colouring");
    else TRY(check_orders(N, nsweeps, orders));
    if (!orders && use_level)
        return launch_level_sweeps(h, qa, (int)nsweeps, mcsteps, 0, jp2.data(), invT.data(), nsched, seed, row0, sweep0,
                                   h->d_recs, h->d_lvoff, nullptr, h->lv_period, h->flow_extra, 0, h->lv_width);

    // per-sweep parameters (fast path)
    DevBuf<float> d_jp2, d_invT;
    if (fast || resident) {
        std::vector<float> a(nsweeps), b(nsweeps);
        for (size_t s = 0; s < nsweeps; s++) {
            a[s] = jp2[s / mcsteps];
            b[s] = invT[s / mcsteps];
        }
        PIQMC_CUDA(d_jp2.alloc(nsweeps));
        PIQMC_CUDA(d_invT.alloc(nsweeps));
        PIQMC_CUDA(cudaMemcpyAsync(d_jp2.p, a.data(), nsweeps * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        PIQMC_CUDA(cudaMemcpyAsync(d_invT.p, b.data(), nsweeps * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        PIQMC_CUDA(cudaStreamSynchronize(h->stream));     // a, b are about to go out of scope
    }

    if (resident) {
        // sequential sweeps in shared memory: the classes in ascending order, or the visiting orders themselves.
        // Integer couplings: the bit-sliced kernel takes the sweeps that are cold enough for it (few lanes need
        // a uniform: T <= unit / 2; it visits those lanes one by one) -- with a falling schedule the hot head of
        // the run goes through the per-lane kernel; SA words are not split over threads there, so it also needs
        // enough rows to fill the device.
        size_t cold_from = nsweeps;
        if (resident_int_ok(h, qa, trotter) && (qa || h->nrows >= 4096 || getenv("PIQMC_FORCE_INT_KERNEL"))) {
            cold_from = 0;
            if (!getenv("PIQMC_FORCE_INT_KERNEL"))
                for (size_t s = 0; s < nsweeps; s++)
                    if (1.0f / invT[s / mcsteps] > 0.5f * h->int_unit) cold_from = s + 1;
        }
        const size_t chunk = !orders ? nsweeps
                                     : std::max<size_t>(1, std::min<size_t>(nsweeps, (size_t)(64u << 20) / ((size_t)N * 4)));
        DevBuf<int32_t> d_ord;
        if (orders) PIQMC_CUDA(d_ord.alloc(chunk * N));
        for (size_t base = 0; base < nsweeps;) {
            size_t m = std::min(chunk, nsweeps - base);
            if (base < cold_from) m = std::min(m, cold_from - base);          // a segment is hot or cold, not both
            if (orders)
                PIQMC_CUDA(cudaMemcpyAsync(d_ord.p, orders + base * N, m * N * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
            TRY(launch_resident_sweeps(h, qa, trotter, orders ? d_ord.p : h->d_members, orders ? 1 : 0, (int)m, d_jp2.p + base,
                                       d_invT.p + base, seed, row0, sweep0 + (uint32_t)base, base >= cold_from ? 1 : 0));
            if (orders) PIQMC_CUDA(cudaStreamSynchronize(h->stream));         // the orders are reused or freed next
            base += m;
        }
        PIQMC_CUDA(cudaStreamSynchronize(h->stream));        // the per-sweep parameter arrays must outlive the launches
        return PIQMC_OK;
    }
    if (!orders && !use_level && chain_selected(h, qa, trotter))
        return launch_chain_sweeps(h, qa, nsched, mcsteps, jp2.data(), invT.data(), seed, row0, sweep0);
    if (!orders) {
        if (fast) {
            if (h->pipe_request) {
                const int rpb = fast_chunk_rows(h);
                const int nchunks = (h->nrows + rpb - 1) / rpb;
                h->pipe_lag16 = pipe_choose_lag16(h, nsweeps, rpb, nchunks);
                // no stagger, no overlap: the plain sequence (one contiguous download under the energy reduction) is faster
                if (nchunks <= 255 && (h->pipe_lag16 > 0 || getenv("PIQMC_PIPE_LAG16"))) TRY(pipe_prepare(h, nchunks));
                else h->pipe_request = 0;
            }
            TRY(launch_fast_sweeps(h, qa, trotter, (int)nsweeps, h->d_recs, h->d_pate, h->flow_extra, 0, d_jp2.p, d_invT.p, seed,
                                   row0, sweep0));
            if (h->pipe_armed) TRY(pipe_drain(h));
            // the per-sweep parameter arrays must outlive the launch
            PIQMC_CUDA(cudaStreamSynchronize(h->stream));
            return piqmc_check_watchdog(h, "dataflow sweeps");
        }
        uint32_t sweep = sweep0;
        for (int f = 0; f < nsched; f++)
            for (int s = 0; s < mcsteps; s++, sweep++)
                for (int c = 0; c < h->ncolors; c++)
                    TRY(launch_colour_sweep(h, qa, trotter, h->d_members + h->color_off[c],
                                            h->color_off[c + 1] - h->color_off[c], jp2[f], invT[f], seed, row0,
                                            sweep));
        return PIQMC_OK;
    }

    // per-sweep visiting orders: level-colour each sweep on the host, ship member lists (and
    // levels) in chunks of sweeps (bounded device memory)
    const bool wantrec = fast || use_level;
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(nsweeps, (size_t)(64u << 20) /
                                                                        ((size_t)N * (wantrec ? sizeof(PiqmcUnitRec) : 4))));
    DevBuf<int32_t> d_mem;
    DevBuf<PiqmcUnitRec> d_rec;
    DevBuf<int> d_soff, d_ssweep;
    DevBuf<float> d_pate;
    if (wantrec) PIQMC_CUDA(d_rec.alloc(chunk * N));
    if (fast) PIQMC_CUDA(d_pate.alloc(chunk * N * 16));
    else         PIQMC_CUDA(d_mem.alloc(chunk * N));
    std::vector<int32_t> members(chunk * N), levels(chunk * N);
    std::vector<PiqmcUnitRec> recs(wantrec ? chunk * N : 0);
    std::vector<std::vector<int>> offs(chunk);
    for (size_t base = 0; base < nsweeps; base += chunk) {
        const size_t m = std::min(chunk, nsweeps - base);
        for (size_t s = 0; s < m; s++) {
            int32_t *lev = levels.data() + s * N;
            const int nlev = order_levels(N, h->maxnb, h->h_idx.data(), h->h_live.data(),
                                          orders + (base + s) * N, lev);
            bucket_members(N, nlev, lev, offs[s], members.data() + s * N);
            if (wantrec) build_unit_recs(h, lev, members.data() + s * N, nullptr, recs.data() + s * N);
        }
        PIQMC_CUDA(cudaStreamSynchronize(h->stream));     // previous chunk's lists no longer in use
        if (use_level) {
            // steps = the levels of every sweep of the chunk, in order
            std::vector<int> soff, ssweep;
            int width = 1;
            for (size_t s = 0; s < m; s++)
                for (size_t c = 0; c + 1 < offs[s].size(); c++) {
                    soff.push_back((int)(s * N) + offs[s][c]);
                    ssweep.push_back((int)s);
                    width = std::max(width, offs[s][c + 1] - offs[s][c]);
                }
            soff.push_back((int)(m * N));
            if (d_soff.p) cudaFree(d_soff.p), d_soff.p = nullptr;
            if (d_ssweep.p) cudaFree(d_ssweep.p), d_ssweep.p = nullptr;
            PIQMC_CUDA(d_soff.alloc(soff.size()));
            PIQMC_CUDA(d_ssweep.alloc(ssweep.size()));
            PIQMC_CUDA(cudaMemcpyAsync(d_rec.p, recs.data(), m * N * sizeof(PiqmcUnitRec), cudaMemcpyHostToDevice,
                                       h->stream));
            PIQMC_CUDA(cudaMemcpyAsync(d_soff.p, soff.data(), soff.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
            PIQMC_CUDA(cudaMemcpyAsync(d_ssweep.p, ssweep.data(), ssweep.size() * sizeof(int), cudaMemcpyHostToDevice,
                                       h->stream));
            const int f0 = (int)(base / mcsteps), foff = (int)(base % mcsteps);
            const int nfc = (int)((foff + m + mcsteps - 1) / mcsteps);
            TRY(launch_level_sweeps(h, qa, (int)m, mcsteps, foff, jp2.data() + f0, invT.data() + f0, nfc, seed, row0,
                                    sweep0 + (uint32_t)base, d_rec.p, d_soff.p, d_ssweep.p, 0, 0, (int)ssweep.size(),
                                    width));               // synchronises the stream before it returns
            continue;
        }
        if (fast) {
            PIQMC_CUDA(cudaMemcpyAsync(d_rec.p, recs.data(), m * N * sizeof(PiqmcUnitRec), cudaMemcpyHostToDevice,
                                       h->stream));
            std::vector<float> pate(m * N * 16);
            build_pattern_energies(recs.data(), m * N, pate.data());
            PIQMC_CUDA(cudaMemcpyAsync(d_pate.p, pate.data(), pate.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
            TRY(launch_fast_sweeps(h, qa, trotter, (int)m, d_rec.p, d_pate.p, 0, 1, d_jp2.p + base, d_invT.p + base, seed, row0,
                                   sweep0 + (uint32_t)base));
        } else {
            PIQMC_CUDA(cudaMemcpyAsync(d_mem.p, members.data(), m * N * sizeof(int32_t), cudaMemcpyHostToDevice,
                                       h->stream));
            for (size_t s = 0; s < m; s++) {
                const int f = (int)((base + s) / mcsteps);
                const std::vector<int> &off = offs[s];
                for (size_t c = 0; c + 1 < off.size(); c++)
                    TRY(launch_colour_sweep(h, qa, trotter, d_mem.p + s * N + off[c], off[c + 1] - off[c], jp2[f],
                                            invT[f], seed, row0, sweep0 + (uint32_t)(base + s)));
            }
        }
        PIQMC_CUDA(cudaStreamSynchronize(h->stream));     // lists are reused or freed next
        if (fast) TRY(piqmc_check_watchdog(h, "dataflow sweeps"));
    }
    return PIQMC_OK;
}

}  // namespace

extern "C" {

int piqmc_version(void) { return 100; }

const char *piqmc_last_error(void) { return g_err; }

int piqmc_device_count(int *count)
{
    PIQMC_REQUIRE(count != nullptr, PIQMC_EINVAL, "null count");
    PIQMC_CUDA(cudaGetDeviceCount(count));
    return PIQMC_OK;
}

int piqmc_create(int device, piqmc_handle *out)
{
    PIQMC_REQUIRE(out != nullptr, PIQMC_EINVAL, "null out");
    int n = 0;
    PIQMC_CUDA(cudaGetDeviceCount(&n));
    PIQMC_REQUIRE(device >= 0 && device < n, PIQMC_EINVAL, "device %d out of range (have %d)", device, n);
    PIQMC_CUDA(cudaSetDevice(device));
    piqmc_ctx *c = new (std::nothrow) piqmc_ctx();
    PIQMC_REQUIRE(c != nullptr, PIQMC_ENOMEM, "out of host memory");
    c->device = device;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        piqmc_set_error("device init failed: %s", cudaGetErrorString(e));
        delete c;
        return PIQMC_ECUDA;
    }
    if (prop.major != 10) {                         // the kernels exist for sm_100a only: say so now, not at the first launch
        piqmc_set_error("device %d is sm_%d%d; libpiqmc_b200 is built for sm_100a (B200) only", device, prop.major,
                        prop.minor);
        cudaStreamDestroy(c->stream);
        delete c;
        return PIQMC_ECUDA;
    }
    c->sm_count = prop.multiProcessorCount;
    *out = c;
    return PIQMC_OK;
}

int piqmc_destroy(piqmc_handle h)
{
    if (!h) return PIQMC_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_graph(h);
    free_state(h);
    free_dev(h->d_epart);
    free_dev(h->d_ticket);
    free_dev(h->d_stage);
    free_dev(h->d_err);
    free_dev(h->d_chot);
    free_dev(h->d_ccold);
    free_dev(h->d_cband);
    free_dev(h->d_cprog);
    free_dev(h->d_cll);
    free_dev(h->d_chunk_count);
    if (h->h_chunk_flag) cudaFreeHost(h->h_chunk_flag);
    if (h->aux_event) cudaEventDestroy(h->aux_event);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if (h->copy_event) cudaEventDestroy(h->copy_event);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    cudaStreamDestroy(h->stream);
    delete h;
    return PIQMC_OK;
}

int piqmc_synchronize(piqmc_handle h)
{
    USE(h);
    PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    return PIQMC_OK;
}

void *piqmc_stream(piqmc_handle h) { return h ? (void *)h->stream : nullptr; }
uint64_t piqmc_launch_count(piqmc_handle h) { return h ? h->launches : 0; }

int piqmc_host_alloc(uint64_t bytes, void **out)
{
    PIQMC_REQUIRE(out != nullptr, PIQMC_EINVAL, "null out");
    PIQMC_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return PIQMC_OK;
}

int piqmc_host_free(void *p)
{
    if (p) PIQMC_CUDA(cudaFreeHost(p));
    return PIQMC_OK;
}

// ---- glibc rand() on the host ---------------------------------------------------------------
void piqmc_rand_seed(piqmc_rand_state *s, unsigned int seed)
{
    // glibc srandom_r for TYPE_3: r[0] = seed (0 -> 1); r[i] = 16807*r[i-1] mod (2^31-1) by
    // Schrage's method; front = 3, rear = 0; discard 310 outputs.
    if (seed == 0) seed = 1;
    s->r[0] = seed;
    int32_t word = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
        long hi = word / 127773, lo = word % 127773;
        word = (int32_t)(16807 * lo - 2836 * hi);
        if (word < 0) word += 2147483647;
        s->r[i] = (uint32_t)word;
    }
    s->f = 3;
    s->b = 0;
    for (int i = 0; i < 310; i++) (void)piqmc_rand_next(s);
}

int32_t piqmc_rand_next(piqmc_rand_state *s)
{
    uint32_t v = (s->r[s->f] += s->r[s->b]);
    s->f = (s->f + 1 == 31) ? 0 : s->f + 1;
    s->b = (s->b + 1 == 31) ? 0 : s->b + 1;
    return (int32_t)(v >> 1);
}

#if defined(__GLIBC__)
// glibc's initstate()/setstate() hand back a pointer to the previous state array whose word 0 is
// (rear_index * 5 + type) and whose words 1..31 are r[0..30] (stdlib/random_r.c).
int piqmc_rand_capture_libc(piqmc_rand_state *s)
{
    PIQMC_REQUIRE(s != nullptr, PIQMC_EINVAL, "null state");
    static char scratch[128];
    int32_t *old = (int32_t *)initstate(1u, scratch, sizeof(scratch));
    PIQMC_REQUIRE(old != nullptr, PIQMC_EINVAL, "initstate failed");
    const int type = old[0] % 5, rear = old[0] / 5;
    int rc = PIQMC_OK;
    if (type != 3 || rear < 0 || rear >= 31) {
        piqmc_set_error("libc rand() is not in its default TYPE_3 state (type %d)", type);
        rc = PIQMC_EINVAL;
    } else {
        for (int i = 0; i < 31; i++) s->r[i] = (uint32_t)old[1 + i];
        s->b = rear;
        s->f = (rear + 3) % 31;
    }
    setstate((char *)old);
    return rc;
}

int piqmc_rand_restore_libc(const piqmc_rand_state *s)
{
    PIQMC_REQUIRE(s != nullptr && s->b >= 0 && s->b < 31, PIQMC_EINVAL, "bad state");
    static char scratch[128];
    int32_t *old = (int32_t *)initstate(1u, scratch, sizeof(scratch));
    PIQMC_REQUIRE(old != nullptr, PIQMC_EINVAL, "initstate failed");
    for (int i = 0; i < 31; i++) old[1 + i] = (int32_t)s->r[i];
    old[0] = s->b * 5 + 3;
    setstate((char *)old);
    return PIQMC_OK;
}
#else
int piqmc_rand_capture_libc(piqmc_rand_state *) { piqmc_set_error("glibc only"); return PIQMC_EINVAL; }
int piqmc_rand_restore_libc(const piqmc_rand_state *) { piqmc_set_error("glibc only"); return PIQMC_EINVAL; }
#endif

float piqmc_jperp(double gamma, int slices, float temp)
{
    // piqmc/qmc.pyx:95 with the casts of piqmc/qmc.c:2063-2075
    const float t12 = (float)slices * temp;
    return (float)(((-0.5 * slices) * (double)temp) * log(tanh(gamma / (double)t12)));
}

// ---- graph --------------------------------------------------------------------------------
int piqmc_set_graph(piqmc_handle h, int nspins, int maxnb, const int32_t *idx, const double *J,
                    int ncolors, const int32_t *color)
{
    USE(h);
    PIQMC_REQUIRE(nspins > 0 && maxnb > 0 && idx && J, PIQMC_EINVAL, "bad graph arguments");
    const size_t ne = (size_t)nspins * maxnb;
    for (size_t e = 0; e < ne; e++)
        PIQMC_REQUIRE(idx[e] >= 0 && idx[e] < nspins, PIQMC_EINVAL,
                      "neighbour index %d out of range at entry %zu", idx[e], e);
    if (color) TRY(check_colouring(nspins, maxnb, idx, J, ncolors, color));
    // every bond must appear in the rows of both its spins with the same coupling, as tools.GenerateNeighbors
    // lists it (piqmc/tools.pyx:74-96): the sweeps read row i for spin i, the energy reduction counts a bond
    // in the row of its smaller index
    {
        std::vector<int> pending;
        for (int i = 0; i < nspins; i++)
            for (int n = 0; n < maxnb; n++) {
                const size_t e = (size_t)i * maxnb + n;
                const int j = idx[e];
                if (j == i || J[e] == 0.0) continue;
                int listed = 0, listed_back = 0;
                for (int m = 0; m < maxnb; m++) {
                    if (idx[(size_t)i * maxnb + m] == j && J[(size_t)i * maxnb + m] == J[e]) listed++;
                    if (idx[(size_t)j * maxnb + m] == i && J[(size_t)j * maxnb + m] == J[e]) listed_back++;
                }
                PIQMC_REQUIRE(listed == listed_back, PIQMC_EINVAL,
                              "neighbour table is not symmetric: the bond (%d, %d) with J = %g is listed %d time(s) in "
                              "row %d and %d time(s) in row %d", i, j, J[e], listed, i, listed_back, j);
            }
    }
    PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    const int old_nspins = h->nspins;
    free_graph(h);

    std::vector<float> j32(ne), j32t(ne);
    std::vector<int32_t> idxt(ne);
    for (int i = 0; i < nspins; i++)
        for (int n = 0; n < maxnb; n++) {
            const size_t e = (size_t)i * maxnb + n;
            j32[e] = (float)J[e];                       // C-float narrowing, qmc.pyx:106
            j32t[(size_t)n * nspins + i] = j32[e];
            idxt[(size_t)n * nspins + i] = idx[e];
        }
    PIQMC_CUDA(cudaMalloc(&h->d_idx, ne * sizeof(int32_t)));
    PIQMC_CUDA(cudaMalloc(&h->d_J32, ne * sizeof(float)));
    PIQMC_CUDA(cudaMalloc(&h->d_J64, ne * sizeof(double)));
    PIQMC_CUDA(cudaMalloc(&h->d_idx_t, ne * sizeof(int32_t)));
    PIQMC_CUDA(cudaMalloc(&h->d_J32_t, ne * sizeof(float)));
    PIQMC_CUDA(cudaMemcpy(h->d_idx, idx, ne * sizeof(int32_t), cudaMemcpyHostToDevice));
    PIQMC_CUDA(cudaMemcpy(h->d_J32, j32.data(), ne * sizeof(float), cudaMemcpyHostToDevice));
    PIQMC_CUDA(cudaMemcpy(h->d_J64, J, ne * sizeof(double), cudaMemcpyHostToDevice));
    PIQMC_CUDA(cudaMemcpy(h->d_idx_t, idxt.data(), ne * sizeof(int32_t), cudaMemcpyHostToDevice));
    PIQMC_CUDA(cudaMemcpy(h->d_J32_t, j32t.data(), ne * sizeof(float), cudaMemcpyHostToDevice));
    {
        // the entries the energy reduction adds: every stored key once (the row of its smaller index), fields as
        // self entries; explicit zeros add +-0 and are left out
        std::vector<int32_t> off(nspins + 1, 0), fj;
        std::vector<double> fJ;
        for (int i = 0; i < nspins; i++) {
            for (int n = 0; n < maxnb; n++) {
                const size_t e = (size_t)i * maxnb + n;
                if (idx[e] >= i && J[e] != 0.0) {
                    fj.push_back(idx[e]);
                    fJ.push_back(J[e]);
                }
            }
            off[i + 1] = (int32_t)fj.size();
        }
        PIQMC_CUDA(cudaMalloc(&h->d_fb_off, off.size() * sizeof(int32_t)));
        PIQMC_CUDA(cudaMalloc(&h->d_fb_j, std::max<size_t>(fj.size(), 1) * sizeof(int32_t)));
        PIQMC_CUDA(cudaMalloc(&h->d_fb_J, std::max<size_t>(fJ.size(), 1) * sizeof(double)));
        PIQMC_CUDA(cudaMemcpy(h->d_fb_off, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        if (!fj.empty()) {
            PIQMC_CUDA(cudaMemcpy(h->d_fb_j, fj.data(), fj.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            PIQMC_CUDA(cudaMemcpy(h->d_fb_J, fJ.data(), fJ.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
    }
    h->nspins = nspins;
    h->maxnb = maxnb;
    h->h_idx.assign(idx, idx + ne);
    h->h_J32 = j32;
    h->h_live.resize(ne);
    for (int i = 0; i < nspins; i++)
        for (int n = 0; n < maxnb; n++) {
            const size_t e = (size_t)i * maxnb + n;
            h->h_live[e] = (idx[e] != i && J[e] != 0.0) ? 1 : 0;
        }
    PIQMC_CUDA(cudaMalloc(&h->d_members, (size_t)nspins * sizeof(int32_t)));
    PIQMC_CUDA(cudaMalloc(&h->d_level, (size_t)nspins * sizeof(int32_t)));
    PIQMC_CUDA(cudaMalloc(&h->d_recs, (size_t)nspins * sizeof(PiqmcUnitRec)));
    PIQMC_CUDA(cudaMalloc(&h->d_pate, (size_t)nspins * 16 * sizeof(float)));
    TRY(upload_level_stat(h));
    // integer couplings: the largest power of two q <= the smallest |J| such that every J32 is an integer
    // multiple of q with |multiple| <= 7 and row sums of at most 31 (bit-sliced resident kernel)
    {
        float amin = 0.0f;
        for (size_t e = 0; e < ne; e++)
            if (j32[e] != 0.0f && (amin == 0.0f || fabsf(j32[e]) < amin)) amin = fabsf(j32[e]);
        float q = 0.0f;
        if (amin > 0.0f && maxnb <= 8) {
            int ex = 0;
            frexpf(amin, &ex);
            q = ldexpf(1.0f, ex - 1);                              // 2^floor(log2(amin))
            for (int tries = 0; tries < 3 && q > 0.0f; tries++, q *= 0.5f) {
                bool ok = true;
                for (int i = 0; i < nspins && ok; i++) {
                    int rowsum = 0;
                    for (int n = 0; n < maxnb && ok; n++) {
                        const float m = j32[(size_t)i * maxnb + n] / q;
                        if (m != rintf(m) || fabsf(m) > 7.0f) ok = false;
                        rowsum += (int)fabsf(m);
                    }
                    if (rowsum > 31) ok = false;
                }
                if (ok) break;
                if (tries == 2) q = 0.0f;
            }
        }
        if (q > 0.0f) {
            bool ok = true;
            std::vector<int8_t> iw(ne);
            for (size_t e = 0; e < ne && ok; e++) {
                const float m = j32[e] / q;
                ok = m == rintf(m) && fabsf(m) <= 7.0f;
                iw[e] = (int8_t)m;
            }
            if (ok) {
                PIQMC_CUDA(cudaMalloc(&h->d_iw, ne));
                PIQMC_CUDA(cudaMemcpy(h->d_iw, iw.data(), ne, cudaMemcpyHostToDevice));
                h->int_unit = q;
            }
        }
    }
    if (color) TRY(apply_colouring(h, ncolors, color, false));
    // a resident state stays valid for a new graph on the same spins (words are [spin][row],
    // whatever the couplings): the SA pre-anneal -> PIQMC hand-over may change the colouring
    if (old_nspins != nspins) free_state(h);
    return PIQMC_OK;
}

int piqmc_set_colouring(piqmc_handle h, int ncolors, const int32_t *color)
{
    USE(h);
    PIQMC_REQUIRE(h->nspins > 0, PIQMC_ENOGRAPH, "piqmc_set_graph has not been called");
    return apply_colouring(h, ncolors, color, true);
}

int piqmc_order_levels(int nspins, int maxnb, const int32_t *idx, const double *J, const int32_t *order,
                       int32_t *level)
{
    PIQMC_REQUIRE(nspins > 0 && maxnb > 0 && idx && J && level, PIQMC_EINVAL, "bad arguments");
    const size_t ne = (size_t)nspins * maxnb;
    std::vector<uint8_t> live(ne);
    for (int i = 0; i < nspins; i++)
        for (int n = 0; n < maxnb; n++) {
            const size_t e = (size_t)i * maxnb + n;
            PIQMC_REQUIRE(idx[e] >= 0 && idx[e] < nspins, PIQMC_EINVAL, "neighbour index out of range");
            live[e] = (idx[e] != i && J[e] != 0.0) ? 1 : 0;
        }
    if (order) TRY(check_orders(nspins, 1, order));
    return order_levels(nspins, maxnb, idx, live.data(), order, level);
}

// ---- deterministic paths ---------------------------------------------------------------------
static int det_common_checks(piqmc_handle h, const double *sched, int nsched, int mcsteps, int nreplicas,
                             const void *spins, const int32_t *perms, const piqmc_rand_state *rstate,
                             const double *uniforms, const double *dense = nullptr, int dense_n = 0)
{
    if (dense) PIQMC_REQUIRE(dense_n > 0, PIQMC_EINVAL, "nspins must be positive");
    else PIQMC_REQUIRE(h->nspins > 0, PIQMC_ENOGRAPH, "piqmc_set_graph has not been called");
    PIQMC_REQUIRE(sched && nsched >= 0 && mcsteps >= 0 && nreplicas > 0 && spins && perms, PIQMC_EINVAL,
                  "bad arguments");
    PIQMC_REQUIRE(rstate || uniforms, PIQMC_EINVAL, "need either rstate or uniforms");
    return PIQMC_OK;
}

static int qa_det_impl(piqmc_handle h, const double *sched, int nsched, int mcsteps, int slices, float temp,
                       int nreplicas, int8_t *spins, const int32_t *perms, piqmc_rand_state *rstate,
                       const double *uniforms, uint64_t nuniforms, uint64_t *consumed, const double *dense,
                       int dense_n)
{
    USE(h);
    TRY(det_common_checks(h, sched, nsched, mcsteps, nreplicas, spins, perms, rstate, uniforms, dense, dense_n));
    PIQMC_REQUIRE(slices >= 2, PIQMC_EINVAL, "slices must be >= 2 (the reference reads slice 1)");
    // ZeroDivisionError conditions of piqmc/qmc.c:2065-2074 and :2407-2416
    PIQMC_REQUIRE((float)slices * temp != 0.0f && temp != 0.0f, PIQMC_EZERODIV, "float division");
    const int N = dense ? dense_n : h->nspins;
    DevBuf<double> d_dense;
    if (dense) {
        PIQMC_CUDA(d_dense.alloc((size_t)N * N));
        PIQMC_CUDA(cudaMemcpyAsync(d_dense.p, dense, (size_t)N * N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    const size_t nsweeps = (size_t)nsched * mcsteps;
    std::vector<float> jperp(std::max(nsched, 1));
    for (int f = 0; f < nsched; f++) jperp[f] = piqmc_jperp(sched[f], slices, temp);

    DevBuf<float> d_jperp;
    DevBuf<int8_t> d_spins;
    DevBuf<int32_t> d_perms;
    DevBuf<piqmc_rand_state> d_rs;
    DevBuf<double> d_uni;
    DevBuf<unsigned long long> d_cons;
    const size_t nsp = (size_t)nreplicas * N * slices, npm = (size_t)nreplicas * nsweeps * N;
    PIQMC_CUDA(d_jperp.alloc(nsched));
    PIQMC_CUDA(d_spins.alloc(nsp));
    PIQMC_CUDA(d_perms.alloc(npm));
    PIQMC_CUDA(d_cons.alloc(nreplicas));
    PIQMC_CUDA(cudaMemcpyAsync(d_jperp.p, jperp.data(), nsched * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    PIQMC_CUDA(cudaMemcpyAsync(d_spins.p, spins, nsp, cudaMemcpyHostToDevice, h->stream));
    PIQMC_CUDA(cudaMemcpyAsync(d_perms.p, perms, npm * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    if (uniforms) {
        PIQMC_CUDA(d_uni.alloc((size_t)nreplicas * nuniforms));
        PIQMC_CUDA(cudaMemcpyAsync(d_uni.p, uniforms, (size_t)nreplicas * nuniforms * sizeof(double),
                                   cudaMemcpyHostToDevice, h->stream));
    } else {
        PIQMC_CUDA(d_rs.alloc(nreplicas));
        PIQMC_CUDA(cudaMemcpyAsync(d_rs.p, rstate, (size_t)nreplicas * sizeof(piqmc_rand_state),
                                   cudaMemcpyHostToDevice, h->stream));
    }
    TRY(launch_qa_det(h, d_jperp.p, nsched, mcsteps, slices, temp, nreplicas, d_spins.p, d_perms.p,
                      uniforms ? nullptr : d_rs.p, uniforms ? d_uni.p : nullptr, nuniforms, d_cons.p,
                      dense ? d_dense.p : nullptr, N));
    PIQMC_CUDA(cudaMemcpyAsync(spins, d_spins.p, nsp, cudaMemcpyDeviceToHost, h->stream));
    if (!uniforms)
        PIQMC_CUDA(cudaMemcpyAsync(rstate, d_rs.p, (size_t)nreplicas * sizeof(piqmc_rand_state),
                                   cudaMemcpyDeviceToHost, h->stream));
    if (consumed)
        PIQMC_CUDA(cudaMemcpyAsync(consumed, d_cons.p, (size_t)nreplicas * sizeof(uint64_t),
                                   cudaMemcpyDeviceToHost, h->stream));
    PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    return PIQMC_OK;
}

static int sa_det_impl(piqmc_handle h, const double *sched, int nsched, int mcsteps, int nreplicas,
                       int8_t *spins, const int32_t *perms, piqmc_rand_state *rstate, const double *uniforms,
                       uint64_t nuniforms, uint64_t *consumed, const double *dense, int dense_n)
{
    USE(h);
    TRY(det_common_checks(h, sched, nsched, mcsteps, nreplicas, spins, perms, rstate, uniforms, dense, dense_n));
    const int N = dense ? dense_n : h->nspins;
    DevBuf<double> d_dense;
    if (dense) {
        PIQMC_CUDA(d_dense.alloc((size_t)N * N));
        PIQMC_CUDA(cudaMemcpyAsync(d_dense.p, dense, (size_t)N * N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    const size_t nsweeps = (size_t)nsched * mcsteps;
    std::vector<float> temps(std::max(nsched, 1));
    for (int t = 0; t < nsched; t++) temps[t] = (float)sched[t];      // sa.pyx:96

    DevBuf<float> d_temps;
    DevBuf<int8_t> d_spins;
    DevBuf<int32_t> d_perms;
    DevBuf<piqmc_rand_state> d_rs;
    DevBuf<double> d_uni;
    DevBuf<unsigned long long> d_cons;
    const size_t nsp = (size_t)nreplicas * N, npm = (size_t)nreplicas * nsweeps * N;
    PIQMC_CUDA(d_temps.alloc(nsched));
    PIQMC_CUDA(d_spins.alloc(nsp));
    PIQMC_CUDA(d_perms.alloc(npm));
    PIQMC_CUDA(d_cons.alloc(nreplicas));
    PIQMC_CUDA(cudaMemcpyAsync(d_temps.p, temps.data(), nsched * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    PIQMC_CUDA(cudaMemcpyAsync(d_spins.p, spins, nsp, cudaMemcpyHostToDevice, h->stream));
    PIQMC_CUDA(cudaMemcpyAsync(d_perms.p, perms, npm * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    if (uniforms) {
        PIQMC_CUDA(d_uni.alloc((size_t)nreplicas * nuniforms));
        PIQMC_CUDA(cudaMemcpyAsync(d_uni.p, uniforms, (size_t)nreplicas * nuniforms * sizeof(double),
                                   cudaMemcpyHostToDevice, h->stream));
    } else {
        PIQMC_CUDA(d_rs.alloc(nreplicas));
        PIQMC_CUDA(cudaMemcpyAsync(d_rs.p, rstate, (size_t)nreplicas * sizeof(piqmc_rand_state),
                                   cudaMemcpyHostToDevice, h->stream));
    }
    TRY(launch_sa_det(h, d_temps.p, nsched, mcsteps, nreplicas, d_spins.p, d_perms.p,
                      uniforms ? nullptr : d_rs.p, uniforms ? d_uni.p : nullptr, nuniforms, d_cons.p,
                      dense ? d_dense.p : nullptr, N));
    PIQMC_CUDA(cudaMemcpyAsync(spins, d_spins.p, nsp, cudaMemcpyDeviceToHost, h->stream));
    if (!uniforms)
        PIQMC_CUDA(cudaMemcpyAsync(rstate, d_rs.p, (size_t)nreplicas * sizeof(piqmc_rand_state),
                                   cudaMemcpyDeviceToHost, h->stream));
    if (consumed)
        PIQMC_CUDA(cudaMemcpyAsync(consumed, d_cons.p, (size_t)nreplicas * sizeof(uint64_t),
                                   cudaMemcpyDeviceToHost, h->stream));
    PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    return PIQMC_OK;
}

int piqmc_qa_det(piqmc_handle h, const double *sched, int nsched, int mcsteps, int slices, float temp,
                 int nreplicas, int8_t *spins, const int32_t *perms, piqmc_rand_state *rstate,
                 const double *uniforms, uint64_t nuniforms, uint64_t *consumed)
{
    return qa_det_impl(h, sched, nsched, mcsteps, slices, temp, nreplicas, spins, perms, rstate, uniforms,
                       nuniforms, consumed, nullptr, 0);
}

int piqmc_sa_det(piqmc_handle h, const double *sched, int nsched, int mcsteps, int nreplicas,
                 int8_t *spins, const int32_t *perms, piqmc_rand_state *rstate, const double *uniforms,
                 uint64_t nuniforms, uint64_t *consumed)
{
    return sa_det_impl(h, sched, nsched, mcsteps, nreplicas, spins, perms, rstate, uniforms, nuniforms, consumed,
                       nullptr, 0);
}

int piqmc_qa_dense_det(piqmc_handle h, int nspins, const double *J, const double *sched, int nsched, int mcsteps,
                       int slices, float temp, int nreplicas, int8_t *spins, const int32_t *perms,
                       piqmc_rand_state *rstate, const double *uniforms, uint64_t nuniforms, uint64_t *consumed)
{
    PIQMC_REQUIRE(J != nullptr, PIQMC_EINVAL, "null coupling matrix");
    return qa_det_impl(h, sched, nsched, mcsteps, slices, temp, nreplicas, spins, perms, rstate, uniforms,
                       nuniforms, consumed, J, nspins);
}

int piqmc_sa_dense_det(piqmc_handle h, int nspins, const double *J, const double *sched, int nsched, int mcsteps,
                       int nreplicas, int8_t *spins, const int32_t *perms, piqmc_rand_state *rstate,
                       const double *uniforms, uint64_t nuniforms, uint64_t *consumed)
{
    PIQMC_REQUIRE(J != nullptr, PIQMC_EINVAL, "null coupling matrix");
    return sa_det_impl(h, sched, nsched, mcsteps, nreplicas, spins, perms, rstate, uniforms, nuniforms, consumed,
                       J, nspins);
}

int piqmc_sa_multispin_det(piqmc_handle h, const double *sched, int nsched, int mcsteps, int ngroups,
                           uint64_t *words, const int32_t *perms, const double *rands)
{
    USE(h);
    PIQMC_REQUIRE(h->nspins > 0, PIQMC_ENOGRAPH, "piqmc_set_graph has not been called");
    PIQMC_REQUIRE(sched && nsched >= 0 && mcsteps >= 0 && ngroups > 0 && words && perms && rands,
                  PIQMC_EINVAL, "bad arguments");
    const int N = h->nspins;
    const size_t nsweeps = (size_t)nsched * mcsteps;
    std::vector<float> temps(std::max(nsched, 1));
    for (int t = 0; t < nsched; t++) temps[t] = (float)sched[t];      // sa.pyx:349
    DevBuf<float> d_temps;
    DevBuf<uint64_t> d_words;
    DevBuf<int32_t> d_perms;
    DevBuf<double> d_rands;
    const size_t nw = (size_t)ngroups * N, npm = (size_t)ngroups * nsweeps * N, nr = npm * 64;
    PIQMC_CUDA(d_temps.alloc(nsched));
    PIQMC_CUDA(d_words.alloc(nw));
    PIQMC_CUDA(d_perms.alloc(npm));
    PIQMC_CUDA(d_rands.alloc(nr));
    PIQMC_CUDA(cudaMemcpyAsync(d_temps.p, temps.data(), nsched * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    PIQMC_CUDA(cudaMemcpyAsync(d_words.p, words, nw * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    PIQMC_CUDA(cudaMemcpyAsync(d_perms.p, perms, npm * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    PIQMC_CUDA(cudaMemcpyAsync(d_rands.p, rands, nr * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    TRY(launch_sa_multispin_det(h, d_temps.p, nsched, mcsteps, ngroups, d_words.p, d_perms.p, d_rands.p));
    PIQMC_CUDA(cudaMemcpyAsync(words, d_words.p, nw * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    return PIQMC_OK;
}

// ---- packed state ----------------------------------------------------------------------------
static int check_packing(int slices, int per_word)
{
    PIQMC_REQUIRE(per_word >= 1 && slices >= 1 && slices * per_word <= 64, PIQMC_EINVAL,
                  "slices * per_word must be at most 64");
    PIQMC_REQUIRE(per_word == 1 || (slices % 4 == 0 && slices >= 4), PIQMC_EINVAL,
                  "several replicas per word need a slice count that is a multiple of 4 (a Philox block "
                  "serves 4 consecutive slices of one replica)");
    return PIQMC_OK;
}

int piqmc_state_alloc(piqmc_handle h, int nrows, int lanes) { return piqmc_state_alloc_packed(h, nrows, lanes, 1); }

int piqmc_state_alloc_packed(piqmc_handle h, int nrows, int slices, int per_word)
{
    USE(h);
    PIQMC_REQUIRE(h->nspins > 0, PIQMC_ENOGRAPH, "piqmc_set_graph has not been called");
    PIQMC_REQUIRE(nrows > 0 && nrows <= (1 << 24), PIQMC_EINVAL, "nrows must be in [1, 2^24]");
    PIQMC_REQUIRE(slices >= 1 && slices <= 64, PIQMC_EINVAL,
                  "lanes must be in [1, 64] (the packed path holds all slices of a spin in one 64-bit word)");
    TRY(check_packing(slices, per_word));
    const int lanes = slices * per_word;
    if (h->d_words && h->nrows == nrows && h->lanes == lanes && h->seg_P == slices && h->seg_S == per_word) {
        // same shape: keep the buffers (and the dataflow flags, which are consistent between runs)
        PIQMC_CUDA(cudaMemsetAsync(h->d_words, 0, (size_t)nrows * h->nspins * sizeof(uint64_t), h->stream));
        return PIQMC_OK;
    }
    PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    free_state(h);
    // one extra, all-zero row behind the last spin: the "neighbour" of self entries and unused columns
    PIQMC_CUDA(cudaMalloc(&h->d_words, (size_t)nrows * (h->nspins + 1) * sizeof(uint64_t)));
    PIQMC_CUDA(cudaMalloc(&h->d_energy, (size_t)nrows * lanes * sizeof(double)));
    PIQMC_CUDA(cudaMemsetAsync(h->d_words, 0, (size_t)nrows * (h->nspins + 1) * sizeof(uint64_t), h->stream));
    h->nrows = nrows;
    h->lanes = lanes;
    h->seg_P = slices;
    h->seg_S = per_word;
    return PIQMC_OK;
}

int piqmc_state_replicas_to_slices(piqmc_handle h, int nreplicas, int slices)
{
    return piqmc_state_replicas_to_slices_packed(h, nreplicas, slices, 1);
}

int piqmc_state_replicas_to_slices_packed(piqmc_handle h, int nreplicas, int slices, int per_word)
{
    USE(h);
    PIQMC_REQUIRE(h->d_words && h->lanes == 64 && h->seg_S == 1, PIQMC_ENOSTATE,
                  "no resident SA state (64 replicas per word)");
    PIQMC_REQUIRE(nreplicas > 0 && nreplicas <= h->nrows * 64, PIQMC_EINVAL, "nreplicas must be in [1, 64*rows]");
    PIQMC_REQUIRE(slices >= 2 && slices <= 64, PIQMC_EINVAL, "slices must be in [2, 64]");
    TRY(check_packing(slices, per_word));
    const int dst_rows = (nreplicas + per_word - 1) / per_word;
    PIQMC_REQUIRE(dst_rows <= (1 << 24), PIQMC_EINVAL, "too many rows");
    uint64_t *src = h->d_words;
    const int src_rows = h->nrows;
    uint64_t *dst = nullptr;
    PIQMC_CUDA(cudaMalloc(&dst, (size_t)dst_rows * (h->nspins + 1) * sizeof(uint64_t)));   // + the zero row
    if (cudaMemsetAsync(dst + (size_t)dst_rows * h->nspins, 0, (size_t)dst_rows * sizeof(uint64_t), h->stream) !=
        cudaSuccess) {
        cudaFree(dst);
        piqmc_set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
        return PIQMC_ECUDA;
    }
    int rc = launch_replicas_to_slices(h, src, src_rows, dst, dst_rows, nreplicas, slices, per_word);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (rc != PIQMC_OK || e != cudaSuccess) {
        cudaFree(dst);
        if (rc == PIQMC_OK) piqmc_set_error("replicas_to_slices failed: %s", cudaGetErrorString(e));
        return rc != PIQMC_OK ? rc : PIQMC_ECUDA;
    }
    free_state(h);
    h->d_words = dst;
    PIQMC_CUDA(cudaMalloc(&h->d_energy, (size_t)dst_rows * slices * per_word * sizeof(double)));
    h->nrows = dst_rows;
    h->lanes = slices * per_word;
    h->seg_P = slices;
    h->seg_S = per_word;
    return PIQMC_OK;
}

int piqmc_state_init_random(piqmc_handle h, uint64_t seed, uint32_t row0, int tile)
{
    USE(h);
    PIQMC_REQUIRE(h->d_words, PIQMC_ENOSTATE, "no packed state (call piqmc_state_alloc)");
    return launch_state_init(h, seed, row0, tile);
}

int piqmc_state_upload_spins(piqmc_handle h, const int8_t *spins, int tile)
{
    USE(h);
    PIQMC_REQUIRE(h->d_words, PIQMC_ENOSTATE, "no packed state (call piqmc_state_alloc)");
    PIQMC_REQUIRE(spins, PIQMC_EINVAL, "null spins");
    const size_t n = (size_t)h->nrows * h->seg_S * h->nspins * (tile ? 1 : h->seg_P);   // nrows*seg_S replicas
    if (n > h->stage_bytes) {                       // grow-only staging buffer: no malloc/free per call
        PIQMC_CUDA(cudaStreamSynchronize(h->stream));
        free_dev(h->d_stage);
        h->stage_bytes = 0;
        PIQMC_CUDA(cudaMalloc(&h->d_stage, n));
        h->stage_bytes = n;
    }
    PIQMC_CUDA(cudaMemcpyAsync(h->d_stage, spins, n, cudaMemcpyHostToDevice, h->stream));
    TRY(launch_pack_spins(h, (const int8_t *)h->d_stage, tile));
    PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    return PIQMC_OK;
}

int piqmc_state_upload_words(piqmc_handle h, const uint64_t *words)
{
    USE(h);
    PIQMC_REQUIRE(h->d_words, PIQMC_ENOSTATE, "no packed state (call piqmc_state_alloc)");
    PIQMC_REQUIRE(words, PIQMC_EINVAL, "null words");
    PIQMC_CUDA(cudaMemcpyAsync(h->d_words, words, (size_t)h->nrows * h->nspins * sizeof(uint64_t),
                               cudaMemcpyHostToDevice, h->stream));
    PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    return PIQMC_OK;
}

int piqmc_state_download_words(piqmc_handle h, uint64_t *words)
{
    USE(h);
    PIQMC_REQUIRE(h->d_words, PIQMC_ENOSTATE, "no packed state (call piqmc_state_alloc)");
    PIQMC_REQUIRE(words, PIQMC_EINVAL, "null words");
    PIQMC_CUDA(cudaMemcpyAsync(words, h->d_words, (size_t)h->nrows * h->nspins * sizeof(uint64_t),
                               cudaMemcpyDeviceToHost, h->stream));
    PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    return PIQMC_OK;
}

void *piqmc_state_devptr(piqmc_handle h) { return h ? (void *)h->d_words : nullptr; }
void *piqmc_energy_devptr(piqmc_handle h) { return h ? (void *)h->d_energy : nullptr; }

int piqmc_set_global_moves(piqmc_handle h, int enable)
{
    PIQMC_REQUIRE(h != nullptr, PIQMC_EINVAL, "null handle");
    h->global_moves = enable ? 1 : 0;
    return PIQMC_OK;
}

int piqmc_set_variant(piqmc_handle h, int variant)
{
    PIQMC_REQUIRE(h != nullptr && variant >= 0 && variant <= 5, PIQMC_EINVAL, "variant must be 0 ... 5");
    h->variant = variant;
    return PIQMC_OK;
}

int piqmc_set_chain(piqmc_handle h, int chain_len)
{
    USE(h);
    PIQMC_REQUIRE(chain_len >= 0, PIQMC_EINVAL, "chain_len must be >= 0 (0 = choose)");
    h->chain_force_C = chain_len;
    if (h->nspins > 0 && h->ncolors > 0) {
        PIQMC_CUDA(cudaStreamSynchronize(h->stream));
        TRY(build_chain_plan(h));
        PIQMC_REQUIRE(chain_len == 0 || h->chain_C > 0, PIQMC_EINVAL,
                      "no chain plan: the colouring is not the natural order's, or the graph is not a lattice of rows of "
                      "chain_len spins (chain_len >= 8, dividing the number of spins)");
    }
    return PIQMC_OK;
}

int piqmc_chain_plan(int nspins, int maxnb, const int32_t *idx, const double *J, int chain_len, uint8_t *kinds,
                     int *wrap)
{
    PIQMC_REQUIRE(nspins > 0 && maxnb > 0 && idx && J && chain_len >= 0, PIQMC_EINVAL, "bad arguments");
    const size_t ne = (size_t)nspins * maxnb;
    std::vector<float> j32(ne);
    std::vector<uint8_t> live(ne);
    for (size_t e = 0; e < ne; e++) {
        PIQMC_REQUIRE(idx[e] >= 0 && idx[e] < nspins, PIQMC_EINVAL, "neighbour index out of range");
        j32[e] = (float)J[e];
        live[e] = (idx[e] != (int32_t)(e / maxnb) && J[e] != 0.0) ? 1 : 0;
    }
    const GraphView gv = {nspins, maxnb, idx, j32.data(), live.data()};
    std::vector<PiqmcChainStat> st;
    double per = 0.0;
    int w = 0;
    const int C = choose_chain(&gv, chain_len, st, per, &w);
    if (C > 0 && kinds)
        for (int i = 0; i < nspins; i++)
            for (int z = 0; z < 4; z++) kinds[(size_t)i * 4 + z] = (uint8_t)((st[i].kinds >> (8 * z)) & 0xFFu);
    if (wrap) *wrap = w;
    return C;
}

int piqmc_chain_info(piqmc_handle h, int *chain_len, int *nchains, double *period, int *selected)
{
    PIQMC_REQUIRE(h != nullptr, PIQMC_EINVAL, "null handle");
    if (chain_len) *chain_len = h->chain_C;
    if (nchains) *nchains = h->chain_n;
    if (period) *period = h->chain_period;
    if (selected) *selected = chain_selected(h, 1, 0) ? 1 : 0;
    return PIQMC_OK;
}

// ---- production sweeps -----------------------------------------------------------------------
int piqmc_qa_colour(piqmc_handle h, const double *sched, int nsched, int mcsteps, float temp,
                    uint64_t seed, uint32_t replica0, uint32_t sweep0, int trotter, const int32_t *orders)
{
    USE(h);
    PIQMC_REQUIRE(h->d_words, PIQMC_ENOSTATE, "no packed state (call piqmc_state_alloc)");
    PIQMC_REQUIRE(sched && nsched >= 0 && mcsteps >= 0, PIQMC_EINVAL, "bad schedule");
    PIQMC_REQUIRE(trotter == 0 || trotter == 1, PIQMC_EINVAL, "trotter must be 0 or 1");
    const int slices = h->seg_P;                    // per replica (several replicas may share a word)
    PIQMC_REQUIRE(slices >= 2, PIQMC_EINVAL, "slices must be >= 2");
    PIQMC_REQUIRE((float)slices * temp != 0.0f && temp != 0.0f, PIQMC_EZERODIV, "float division");
    PIQMC_REQUIRE(!h->global_moves || piqmc_fast_ok(h, 1, trotter), PIQMC_EINVAL,
                  "world-line moves need the table kernel (maxnb <= 4, variant != 1, >= 32 rows)");
    PIQMC_REQUIRE(h->seg_S == 1 || ((piqmc_fast_ok(h, 1, trotter) || (!orders && chain_selected(h, 1, trotter))) &&
                                    trotter == 0 && !h->global_moves), PIQMC_EINVAL,
                  "several replicas per word need the table kernel (maxnb <= 4, variant != 1, >= 32 rows), the "
                  "reference Trotter neighbours and no world-line moves");
    std::vector<float> jp2(std::max(nsched, 1)), invT(std::max(nsched, 1), 1.0f / temp);
    for (int f = 0; f < nsched; f++) jp2[f] = 2.0f * piqmc_jperp(sched[f], slices, temp);
    return run_colour_sweeps(h, 1, trotter, nsched, mcsteps, jp2, invT, seed, replica0, sweep0, orders);
}

// Anneal, energies and download in one call.  With the dataflow kernel and a static colouring the row chunks
// of the launch are staggered and downloaded as they finish (pipe_drain); every other configuration runs the
// three steps one after the other.  Same results either way.
int piqmc_qa_colour_results(piqmc_handle h, const double *sched, int nsched, int mcsteps, float temp, uint64_t seed,
                            uint32_t replica0, uint32_t sweep0, int trotter, const int32_t *orders, double *energies,
                            uint64_t *words)
{
    USE(h);
    PIQMC_REQUIRE(energies && words, PIQMC_EINVAL, "null output");
    h->pipe_armed = 0;
    h->pipe_request = 1;
    if (const char *e = getenv("PIQMC_PIPE")) h->pipe_request = atoi(e) != 0;
    h->pipe_energies = energies;
    h->pipe_words = words;
    struct timespec t0, t1, t2;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    const int rc = piqmc_qa_colour(h, sched, nsched, mcsteps, temp, seed, replica0, sweep0, trotter, orders);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const bool done = h->pipe_armed != 0;
    h->pipe_request = h->pipe_armed = 0;
    h->pipe_energies = nullptr;
    h->pipe_words = nullptr;
    if (rc != PIQMC_OK) return rc;
    const int rc2 = done ? PIQMC_OK : piqmc_results(h, energies, words);
    clock_gettime(CLOCK_MONOTONIC, &t2);
    h->phase_s[0] = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);   // overlapped path: all of it
    h->phase_s[1] = (double)(t2.tv_sec - t1.tv_sec) + 1e-9 * (double)(t2.tv_nsec - t1.tv_nsec);
    return rc2;
}

uint64_t piqmc_pipelined_runs(piqmc_handle h) { return h ? h->pipe_runs : 0; }

int piqmc_last_phase_seconds(piqmc_handle h, double *sweeps, double *results)
{
    PIQMC_REQUIRE(h != nullptr, PIQMC_EINVAL, "null handle");
    if (sweeps) *sweeps = h->phase_s[0];
    if (results) *results = h->phase_s[1];
    return PIQMC_OK;
}

int piqmc_qa_carry(piqmc_handle h, const double *sched, int nsched, int mcsteps, float temp, uint64_t seed,
                   uint32_t replica0, uint32_t sweep0, const int32_t *orders)
{
    USE(h);
    PIQMC_REQUIRE(h->d_words, PIQMC_ENOSTATE, "no packed state (call piqmc_state_alloc)");
    PIQMC_REQUIRE(sched && nsched >= 0 && mcsteps >= 0, PIQMC_EINVAL, "bad schedule");
    PIQMC_REQUIRE(h->seg_S == 1, PIQMC_EINVAL, "the carry sweeps work on states with one replica per word");
    const int slices = h->seg_P, N = h->nspins;
    PIQMC_REQUIRE(slices >= 2, PIQMC_EINVAL, "slices must be >= 2");
    PIQMC_REQUIRE((float)slices * temp != 0.0f && temp != 0.0f, PIQMC_EZERODIV, "float division");
    const size_t nsweeps = (size_t)nsched * mcsteps;
    if (nsweeps == 0) return PIQMC_OK;
    if (orders) TRY(check_orders(N, nsweeps, orders));
    std::vector<float> a(nsweeps), b(nsweeps, 1.0f / temp);
    for (size_t s = 0; s < nsweeps; s++) a[s] = 2.0f * piqmc_jperp(sched[s / mcsteps], slices, temp);
    DevBuf<float> d_jp2, d_invT;
    PIQMC_CUDA(d_jp2.alloc(nsweeps));
    PIQMC_CUDA(d_invT.alloc(nsweeps));
    PIQMC_CUDA(cudaMemcpyAsync(d_jp2.p, a.data(), nsweeps * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    PIQMC_CUDA(cudaMemcpyAsync(d_invT.p, b.data(), nsweeps * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    if (!orders) {
        TRY(launch_qa_carry(h, nullptr, 0, (int)nsweeps, d_jp2.p, d_invT.p, seed, replica0, sweep0));
        PIQMC_CUDA(cudaStreamSynchronize(h->stream));
        return PIQMC_OK;
    }
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(nsweeps, (size_t)(64u << 20) / ((size_t)N * 4)));
    DevBuf<int32_t> d_ord;
    PIQMC_CUDA(d_ord.alloc(chunk * N));
    for (size_t base = 0; base < nsweeps; base += chunk) {
        const size_t m = std::min(chunk, nsweeps - base);
        PIQMC_CUDA(cudaMemcpyAsync(d_ord.p, orders + base * N, m * N * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        TRY(launch_qa_carry(h, d_ord.p, 1, (int)m, d_jp2.p + base, d_invT.p + base, seed, replica0, sweep0 + (uint32_t)base));
        PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    }
    return PIQMC_OK;
}

int piqmc_sa_colour(piqmc_handle h, const double *sched, int nsched, int mcsteps, uint64_t seed,
                    uint32_t row0, uint32_t sweep0, const int32_t *orders)
{
    USE(h);
    PIQMC_REQUIRE(h->d_words, PIQMC_ENOSTATE, "no packed state (call piqmc_state_alloc)");
    PIQMC_REQUIRE(sched && nsched >= 0 && mcsteps >= 0, PIQMC_EINVAL, "bad schedule");
    PIQMC_REQUIRE(h->seg_S == 1, PIQMC_EINVAL, "the SA sweeps work on states with one group of 64 replicas per word");
    std::vector<float> jp2(std::max(nsched, 1), 0.0f), invT(std::max(nsched, 1));
    for (int t = 0; t < nsched; t++) invT[t] = 1.0f / (float)sched[t];
    return run_colour_sweeps(h, 0, 0, nsched, mcsteps, jp2, invT, seed, row0, sweep0, orders);
}

// ---- energies --------------------------------------------------------------------------------
int piqmc_energy(piqmc_handle h, double *energies)
{
    USE(h);
    PIQMC_REQUIRE(h->d_words, PIQMC_ENOSTATE, "no packed state (call piqmc_state_alloc)");
    TRY(launch_energy(h));
    if (energies) {
        PIQMC_CUDA(cudaMemcpyAsync(energies, h->d_energy, (size_t)h->nrows * h->lanes * sizeof(double),
                                   cudaMemcpyDeviceToHost, h->stream));
        PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    }
    return PIQMC_OK;
}

int piqmc_energy_histogram(piqmc_handle h, int reduce, double e0, double scale, double lo, double hi, int nbins,
                           uint64_t *counts, double *stats)
{
    USE(h);
    PIQMC_REQUIRE(h->d_words && h->d_energy, PIQMC_ENOSTATE, "no packed state");
    PIQMC_REQUIRE(reduce >= 0 && reduce <= 2 && nbins >= 1 && nbins <= (1 << 20) && hi > lo && counts, PIQMC_EINVAL,
                  "bad histogram arguments");
    DevBuf<unsigned long long> d_counts;
    DevBuf<double> d_stats;
    PIQMC_CUDA(d_counts.alloc(nbins + 2));
    PIQMC_CUDA(d_stats.alloc(3));
    // order-preserving integer images of +inf / -inf as the starting minimum / maximum
    const long long kmax = 0x7FF0000000000000ll, kmin = (long long)(0x8000000000000000ull - 0xFFF0000000000000ull);
    double init[3] = {0.0, 0.0, 0.0};
    memcpy(&init[1], &kmax, 8);
    memcpy(&init[2], &kmin, 8);
    PIQMC_CUDA(cudaMemsetAsync(d_counts.p, 0, (size_t)(nbins + 2) * sizeof(unsigned long long), h->stream));
    PIQMC_CUDA(cudaMemcpyAsync(d_stats.p, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    TRY(launch_energy_histogram(h, reduce, e0, scale, lo, hi, nbins, d_counts.p, d_stats.p));
    double out[3];
    PIQMC_CUDA(cudaMemcpyAsync(counts, d_counts.p, (size_t)(nbins + 2) * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    PIQMC_CUDA(cudaMemcpyAsync(out, d_stats.p, sizeof(out), cudaMemcpyDeviceToHost, h->stream));
    PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    if (stats) {
        auto unkey = [](double x) {
            long long k;
            memcpy(&k, &x, 8);
            if (k < 0) k = (long long)(0x8000000000000000ull - (unsigned long long)k);
            double v;
            memcpy(&v, &k, 8);
            return v;
        };
        stats[0] = out[0];
        stats[1] = unkey(out[1]);
        stats[2] = unkey(out[2]);
    }
    return PIQMC_OK;
}

int piqmc_results(piqmc_handle h, double *energies, uint64_t *words)
{
    USE(h);
    PIQMC_REQUIRE(h->d_words, PIQMC_ENOSTATE, "no packed state (call piqmc_state_alloc)");
    PIQMC_REQUIRE(energies && words, PIQMC_EINVAL, "null output");
    if (!h->copy_stream) PIQMC_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->copy_event) PIQMC_CUDA(cudaEventCreateWithFlags(&h->copy_event, cudaEventDisableTiming));
    // the download of the packed state (second stream) overlaps the energy reduction
    PIQMC_CUDA(cudaEventRecord(h->copy_event, h->stream));
    PIQMC_CUDA(cudaStreamWaitEvent(h->copy_stream, h->copy_event, 0));
    PIQMC_CUDA(cudaMemcpyAsync(words, h->d_words, (size_t)h->nrows * h->nspins * sizeof(uint64_t),
                               cudaMemcpyDeviceToHost, h->copy_stream));
    TRY(launch_energy(h));
    PIQMC_CUDA(cudaMemcpyAsync(energies, h->d_energy, (size_t)h->nrows * h->lanes * sizeof(double),
                               cudaMemcpyDeviceToHost, h->stream));
    PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    PIQMC_CUDA(cudaStreamSynchronize(h->copy_stream));
    return PIQMC_OK;
}

int piqmc_energy_coo(piqmc_handle h, int nspins, int nnz, const int32_t *row, const int32_t *col,
                     const double *val, int nconfs, const int8_t *spins, double *energies)
{
    USE(h);
    PIQMC_REQUIRE(nspins > 0 && nnz >= 0 && nconfs > 0 && spins && energies, PIQMC_EINVAL, "bad arguments");
    PIQMC_REQUIRE(nnz == 0 || (row && col && val), PIQMC_EINVAL, "null COO arrays");
    for (int e = 0; e < nnz; e++)
        PIQMC_REQUIRE(row[e] >= 0 && row[e] < nspins && col[e] >= 0 && col[e] < nspins, PIQMC_EINVAL,
                      "COO entry %d out of range", e);
    DevBuf<int32_t> d_row, d_col;
    DevBuf<double> d_val, d_out;
    DevBuf<int8_t> d_sp;
    PIQMC_CUDA(d_row.alloc(nnz));
    PIQMC_CUDA(d_col.alloc(nnz));
    PIQMC_CUDA(d_val.alloc(nnz));
    PIQMC_CUDA(d_out.alloc(nconfs));
    PIQMC_CUDA(d_sp.alloc((size_t)nconfs * nspins));
    PIQMC_CUDA(cudaMemcpyAsync(d_row.p, row, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    PIQMC_CUDA(cudaMemcpyAsync(d_col.p, col, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    PIQMC_CUDA(cudaMemcpyAsync(d_val.p, val, nnz * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    PIQMC_CUDA(cudaMemcpyAsync(d_sp.p, spins, (size_t)nconfs * nspins, cudaMemcpyHostToDevice, h->stream));
    TRY(launch_energy_coo(h, nspins, nnz, d_row.p, d_col.p, d_val.p, nconfs, d_sp.p, d_out.p));
    PIQMC_CUDA(cudaMemcpyAsync(energies, d_out.p, nconfs * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    PIQMC_CUDA(cudaStreamSynchronize(h->stream));
    return PIQMC_OK;
}

}  // extern "C"
