// colour_sweep_fast2 -- EXPERIMENT, not compiled into libpiqmc_b200 (profiles/r2_e2e_overlap.md section 4).
//
// The pass loop of the dataflow unit (csrc/colour_fast.cu) with two ADJACENT rows per thread: every word access
// is one 128-bit ld.global.cg.v2.u64 / st.v2 (nrows even), the class loop's control, the decoding of the function
// names, the votes and the indexed jumps are paid once per two words.  Bit-exact against the one-row kernel and
// the oracle on the GPU parity suite; measured 4.55e12 attempts/s (7 blocks / SM) against 4.70e12 for the one-row
// kernel at 4096 rows, so it was not kept.  The unit prologue (ticket, record, table build, dependency wait) and
// epilogue (barrier, release) are those of colour_sweep_fast<true, 0, MINB>; this is what replaced its pass loop
// (to rebuild: paste into a copy of that kernel, launch with rows_per_block a multiple of 2 * FAST_THREADS).

__device__ __forceinline__ ulonglong2 ldcg2(const uint64_t *p)
{
    return __ldcg(reinterpret_cast<const ulonglong2 *>(p));
}

#if 0   // inside the kernel, after the barrier that joins table build and dependency wait
    for (int base = rbeg; base < rend; base += 2 * FAST_THREADS) {     // block-uniform trip count
        const int row = base + 2 * threadIdx.x;                        // even; row + 1 < rend whenever row < rend
        const bool live = row < rend;
        uint64_t w[2] = {0ull, 0ull};
        uint64_t z[2][4];
        {
            ulonglong2 wn[4];
#pragma unroll
            for (int n = 0; n < 4; n++) wn[n] = make_ulonglong2(0ull, 0ull);
            if (live) {
                const ulonglong2 ww = ldcg2(words + (size_t)i * nrows + row);
                w[0] = ww.x;
                w[1] = ww.y;
#pragma unroll
                for (int n = 0; n < 4; n++) wn[n] = ldcg2(words + (size_t)nb[n] * nrows + row);
            }
#pragma unroll
            for (int n = 0; n < 4; n++) {
                const uint64_t sg = ((uint64_t)sgn[n] << 32) | sgn[n];
                z[0][n] = w[0] ^ wn[n].x ^ sg;
                z[1][n] = w[1] ^ wn[n].y ^ sg;
                asm volatile("" : "+l"(z[0][n]), "+l"(z[1][n]));
            }
        }
        const uint32_t prow_warp = a.row0 + (uint32_t)(base + 2 * (threadIdx.x & ~31));

        uint64_t XL[2], XR[2], flip1[2], C[2][3];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const uint64_t bl = (w[r] & a.top) ? ~0ull : 0ull;
            const uint64_t br_old = (w[r] & 2ull) ? ~0ull : 0ull;
            XL[r] = (w[r] ^ bl) & ~a.top;
            const uint32_t c1 = (uint32_t)(XL[r] >> 1) & 1u;
            const uint32_t p1 = pattern_at(z[r], 1);
            flip1[r] = 0ull;
            if (live) {
                const uint32_t t1 = (uint32_t)(lane1tab >> (16u * c1 + p1));
                const uint32_t t2 = (uint32_t)(lane1tab >> (32u + 16u * c1 + p1));
                if (t1 & 1u) flip1[r] = 2ull;
                else if (t2 & 1u)
                    if (lane_uniform(1, (uint32_t)i, sweep, a.row0 + (uint32_t)(row + r), a.k0, a.k1) < tab.thr[c1][p1])
                        flip1[r] = 2ull;
            }
            const uint64_t br_new = br_old ^ (flip1[r] ? ~0ull : 0ull);
            XR[r] = ((w[r] ^ br_new) & ~1ull) | ((w[r] ^ br_old) & 1ull);
            const uint64_t todo = live ? (valid & ~2ull) : 0ull;
            C[r][0] = ~(XL[r] | XR[r]) & todo;
            C[r][1] = (XL[r] ^ XR[r]) & todo;
            C[r][2] = (XL[r] & XR[r]) & todo;
            asm volatile("" : "+l"(C[r][0]), "+l"(C[r][1]), "+l"(C[r][2]));
        }
        uint64_t ACC[2] = {0ull, 0ull}, NEED[2] = {0ull, 0ull}, V[2] = {0ull, 0ull};
        uint32_t last = FID_NONE;
#pragma unroll 1
        for (int c = 0; c < NC; c++) {
            const uint64_t Ca = (c == 0) ? C[0][0] : (c == 1 ? C[0][1] : C[0][2]);
            const uint64_t Cb = (c == 0) ? C[1][0] : (c == 1 ? C[1][1] : C[1][2]);
            if (!__any_sync(0xffffffffu, (Ca | Cb) != 0ull)) continue;
            const uint32_t names = (uint32_t)(fnames >> (16 * c));
            const uint32_t fa = names & 0xFFu, fb = (names >> 8) & 0xFFu;
            if (fa != last || fa == FID_GENERIC) {
                if (fa < PIQMC_NCANON) eval_canon<2>(fa, z, V);          // one indexed jump for four 32-lane halves
                else {
                    V[0] = eval_generic(tab.hacc[c], z[0][0], z[0][1], z[0][2], z[0][3]);
                    V[1] = eval_generic(tab.hacc[c], z[1][0], z[1][1], z[1][2], z[1][3]);
                }
                last = fa;
            }
            ACC[0] |= V[0] & Ca;
            ACC[1] |= V[1] & Cb;
            if (fb != FID_NONE) {
                uint64_t U[2];
                if (fb < PIQMC_NCANON) eval_canon<2>(fb, z, U);
                else {
                    U[0] = eval_generic(tab.hall[c], z[0][0], z[0][1], z[0][2], z[0][3]);
                    U[1] = eval_generic(tab.hall[c], z[1][0], z[1][1], z[1][2], z[1][3]);
                }
                NEED[0] |= U[0] & ~V[0] & Ca;
                NEED[1] |= U[1] & ~V[1] & Cb;
            }
        }
#pragma unroll
        for (int r = 0; r < 2; r++)          // thread l of the warp holds rows 2l and 2l + 1: lstride = 2
            if (__any_sync(0xffffffffu, NEED[r] != 0))
                ACC[r] |= resolve_draws<true>(NEED[r], z[r], XL[r], XR[r], thr_tab, queue, (uint32_t)i, sweep,
                                              prow_warp + (uint32_t)r, a.k0, a.k1, 64, 1, 2);
        if (live) {
            ulonglong2 out;
            out.x = w[0] ^ flip1[0] ^ ACC[0];
            out.y = w[1] ^ flip1[1] ^ ACC[1];
            *reinterpret_cast<ulonglong2 *>(words + (size_t)i * nrows + row) = out;
        }
    }
#endif
