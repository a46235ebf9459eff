/*
 * piqmc_b200.h -- C ABI of libpiqmc_b200.so, the B200 (sm_100a) implementation of the
 * Metropolis-sweep hot path of hadsed/pathintegral-qmc.
 *
 * The reference has no FFI of its own: its "plugin API" is the Python call signature of the
 * Cython functions in piqmc/qmc.pyx, piqmc/sa.pyx and piqmc/tools.pyx.  The entry points below
 * are what a thin Python (ctypes) replacement of those modules binds; the reference function
 * each one stands in for is cited as file:line.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns 0 on success, a negative PIQMC_E* code otherwise;
 *     piqmc_last_error() returns a thread-local message for the last failure;
 *   - plain pointers and sizes only; all pointers are HOST pointers unless the name says dev;
 *   - the caller owns every buffer it passes; the library keeps no host pointer after return;
 *   - a handle owns one CUDA device, one stream and all device buffers; calls on one handle
 *     must not overlap in time (use one handle per thread / per GPU).
 *
 * Spin / bit convention (piqmc/tools.pyx:20-26): bit 0 <-> spin +1, bit 1 <-> spin -1.
 *
 * Packed state ("words"): uint64 word[spin][row] (row fastest); bit `lane` of a word is
 *   - quantum annealing: Trotter slice `lane` of replica `row`        (lanes = P <= 64)
 *   - simulated annealing: replica `row*64 + lane`                    (lanes = 64)
 */
#ifndef PIQMC_B200_H
#define PIQMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIQMC_OK            0
#define PIQMC_EINVAL       -1   /* bad argument (-> ValueError)                         */
#define PIQMC_EZERODIV     -2   /* slices*temp == 0 or temp == 0 (-> ZeroDivisionError,  */
                                /*   as piqmc/qmc.c:2065-2074,2407-2416)                 */
#define PIQMC_ECUDA        -3   /* CUDA runtime failure (-> RuntimeError)                */
#define PIQMC_ENOGRAPH     -4   /* piqmc_set_graph has not been called                   */
#define PIQMC_ENOSTATE     -5   /* no packed state resident                              */
#define PIQMC_ENOMEM       -6

typedef struct piqmc_ctx *piqmc_handle;

/* State of glibc's rand() (TYPE_3 additive feedback generator): what the reference consumes
 * through libc rand() at piqmc/qmc.pyx:132 and piqmc/sa.pyx:116.  f/b are the front and rear
 * indices into r[]. */
typedef struct {
    uint32_t r[31];
    int32_t f, b;
} piqmc_rand_state;

/* ---- library / device ---------------------------------------------------------------- */
int         piqmc_version(void);
const char *piqmc_last_error(void);
int         piqmc_device_count(int *count);
int         piqmc_create(int device, piqmc_handle *out);
int         piqmc_destroy(piqmc_handle h);
int         piqmc_synchronize(piqmc_handle h);
/* the handle's cudaStream_t, for callers that want to record their own events on it */
void       *piqmc_stream(piqmc_handle h);
/* number of kernel launches this handle has issued so far */
uint64_t    piqmc_launch_count(piqmc_handle h);

/* page-locked host buffers (full-speed PCIe transfers for the batched entry points) */
int         piqmc_host_alloc(uint64_t bytes, void **out);
int         piqmc_host_free(void *p);

/* ---- glibc rand() restated on the host (seeding helper for the deterministic paths) ---- */
void    piqmc_rand_seed(piqmc_rand_state *s, unsigned int seed);     /* == srand(seed)       */
int32_t piqmc_rand_next(piqmc_rand_state *s);                        /* == rand()            */
/* copy the process-global libc generator into *s / load *s back into it (glibc only) */
int     piqmc_rand_capture_libc(piqmc_rand_state *s);
int     piqmc_rand_restore_libc(const piqmc_rand_state *s);

/* ---- graph ---------------------------------------------------------------------------
 * The neighbour table of tools.GenerateNeighbors (piqmc/tools.pyx:74-96), split into its two
 * planes: idx[i*maxnb+n] = (int)nbs[i,n,0], J[i*maxnb+n] = nbs[i,n,1] (float64; the kernels
 * narrow it to float32 exactly as piqmc/qmc.pyx:106 does, the energy reduction keeps float64).
 * color[i] in [0,ncolors) is a proper colouring of the graph (ignoring self entries and
 * zero couplings); it may be NULL when only the deterministic paths are used. */
int piqmc_set_graph(piqmc_handle h, int nspins, int maxnb, const int32_t *idx, const double *J,
                    int ncolors, const int32_t *color);
/* replace the colouring of the current graph (cheap: only the class member lists change) */
int piqmc_set_colouring(piqmc_handle h, int ncolors, const int32_t *color);
/* Level colouring of a sequential visiting order (host only, no device work):
 * level[i] = 1 + max(level[j]) over coupled neighbours j visited before i (0 if none).  Updating
 * the levels in ascending order, all spins of a level at once, gives exactly the result of the
 * sequential sweep in that order -- this is how the colour kernels reproduce the reference's
 * natural-order (piqmc/qmc.pyx:320) and permutation-order (piqmc/sa.pyx:100) sweeps.
 * order[t] = spin visited at step t (NULL = natural order 0..N-1).  Returns the level count. */
int piqmc_order_levels(int nspins, int maxnb, const int32_t *idx, const double *J,
                       const int32_t *order, int32_t *level);

/* ---- deterministic, reference-stream paths (bit-exact) --------------------------------
 * One replica per GPU thread, each replaying the reference's sequential algorithm with the
 * reference's random streams:
 *   perms[(r*nsweeps + sweep)*nspins + t]  the spin order of sweep `sweep` of replica r, i.e. the
 *        successive results of rng.permutation (piqmc/qmc.pyx:90-91,136), nsweeps = nsched*mcsteps;
 *   rstate[r]   glibc rand() state of replica r, advanced in place (lazy consumption,
 *        piqmc/qmc.pyx:130-133); or, if uniforms != NULL, uniforms[r*nuniforms + c] is the c-th
 *        value of rand()/RAND_MAX for replica r and rstate may be NULL;
 *   consumed[r] receives the number of uniforms replica r drew (may be NULL).
 * spins are +-1 int8, modified in place.
 */

/* qmc.QuantumAnneal (piqmc/qmc.pyx:30-136), as shipped, including: float32 running ediff reset
 * once per slice, Trotter neighbours slices-1 and 1 for every slice, lazy rand().
 * spins[(r*nspins + i)*slices + k].   jperp is computed on the host by piqmc_jperp. */
int piqmc_qa_det(piqmc_handle h, const double *sched, int nsched, int mcsteps, int slices,
                 float temp, int nreplicas, int8_t *spins, const int32_t *perms,
                 piqmc_rand_state *rstate, const double *uniforms, uint64_t nuniforms,
                 uint64_t *consumed);

/* sa.Anneal (piqmc/sa.pyx:50-120).  spins[r*nspins + i]. */
int piqmc_sa_det(piqmc_handle h, const double *sched, int nsched, int mcsteps, int nreplicas,
                 int8_t *spins, const int32_t *perms, piqmc_rand_state *rstate,
                 const double *uniforms, uint64_t nuniforms, uint64_t *consumed);

/* qmc.QuantumAnneal_dense (piqmc/qmc.pyx:141-242) and sa.Anneal_dense (piqmc/sa.pyx:126-187): the
 * same sweeps with a dense coupling matrix J[nspins*nspins] (float64, row-major; only the upper
 * triangle and the diagonal -- the local fields -- are read, as in the reference).  No graph has
 * to be set.  Couplings stay float64 (the sparse variants narrow them to float); Anneal_dense
 * accepts on ediff > 0 where sa.Anneal accepts on >= 0.  Other arguments as above. */
int piqmc_qa_dense_det(piqmc_handle h, int nspins, const double *J, const double *sched, int nsched,
                       int mcsteps, int slices, float temp, int nreplicas, int8_t *spins,
                       const int32_t *perms, piqmc_rand_state *rstate, const double *uniforms,
                       uint64_t nuniforms, uint64_t *consumed);
int piqmc_sa_dense_det(piqmc_handle h, int nspins, const double *J, const double *sched, int nsched,
                       int mcsteps, int nreplicas, int8_t *spins, const int32_t *perms,
                       piqmc_rand_state *rstate, const double *uniforms, uint64_t nuniforms,
                       uint64_t *consumed);

/* sa.Anneal_multispin (piqmc/sa.pyx:282-405): groups of 64 replicas, float64 ediffs, no
 * ediff>0 shortcut.  words[g*nspins + i]: replica k of group g in bit 63-k (sa.pyx:339-345).
 * rands[((g*nsweeps + sweep)*nspins + t)*64 + k]: the rng.rand(64) block in force at attempt t. */
int piqmc_sa_multispin_det(piqmc_handle h, const double *sched, int nsched, int mcsteps,
                           int ngroups, uint64_t *words, const int32_t *perms,
                           const double *rands);

/* J_perp(Gamma) with the reference's cast order (piqmc/qmc.pyx:95, piqmc/qmc.c:2063-2075). */
float piqmc_jperp(double gamma, int slices, float temp);

/* ---- production colour-parallel paths -------------------------------------------------
 * Packed state resident on the device; Philox4x32-10 keyed by (seed; spin, lane, sweep, row).
 * Semantics are stated on the CPU in oracle/piqmc_oracle.c part 3 and reproduced bit-exactly.
 */
int piqmc_state_alloc(piqmc_handle h, int nrows, int lanes);
/* QA states with at most 32 slices: `per_word` replicas share a word, replica row*per_word + g in
 * bits [g*slices, (g+1)*slices).  Results are those of the one-replica-per-word layout, replica by
 * replica (same Philox keys); only the fast kernel with the reference Trotter neighbours runs on
 * such a state.  slices must be a multiple of 4 when per_word > 1, slices*per_word <= 64.
 * piqmc_state_upload_spins then takes nrows*per_word replicas; piqmc_energy returns
 * energies[(row*per_word + g)*slices + k]; row0/replica0 arguments are replica ids. */
int piqmc_state_alloc_packed(piqmc_handle h, int nrows, int slices, int per_word);
/* Turn the resident SA state (64 replicas per word) into a QA state on the device: one row per
 * replica, every one of the `slices` lanes equal to the replica's spin -- the reference's
 * np.tile(spinVector, (P,1)).T hand-over from the SA pre-anneal to PIQMC
 * (examples/spinglass32.py:94-96,124-127). */
int piqmc_state_replicas_to_slices(piqmc_handle h, int nreplicas, int slices);
int piqmc_state_replicas_to_slices_packed(piqmc_handle h, int nreplicas, int slices, int per_word);
/* random initial state from Philox: tile != 0 -> one bit per (row, spin) copied to every lane
 * (the reference's np.tile(spinVector,(P,1)).T start, examples/spinglass32.py:94-96);
 * tile == 0 -> independent bit per (row*64+lane, spin). */
int piqmc_state_init_random(piqmc_handle h, uint64_t seed, uint32_t row0, int tile);
/* spins[(row*nspins + i)] +-1, copied to every lane (tile != 0), or
 * spins[(row*lanes + lane)*nspins + i] (tile == 0) */
int piqmc_state_upload_spins(piqmc_handle h, const int8_t *spins, int tile);
/* words[i*nrows + row], the device layout */
int piqmc_state_upload_words(piqmc_handle h, const uint64_t *words);
int piqmc_state_download_words(piqmc_handle h, uint64_t *words);
/* device pointer of the packed state / of the last energy result (for NCCL gathers) */
void *piqmc_state_devptr(piqmc_handle h);
void *piqmc_energy_devptr(piqmc_handle h);

/* Path-integral QA sweeps over the resident state (lanes = slices).  Colour-class order,
 * per-spin ediff reset, J_perp recomputed per schedule step.  trotter = 0: neighbours slices-1
 * and 1 (reference, qmc.pyx:115-117); trotter = 1: periodic k-1,k+1 (requires even slices).
 * replica0: global id of row 0 (so results do not depend on how replicas are sharded).
 * orders: NULL -> every sweep uses the graph's colouring; else orders[s*nspins + t] is the spin
 * visited at step t of sweep s (s < nsched*mcsteps): each sweep is run through the level
 * colouring of its own order and so equals the sequential sweep in that order.
 * Returns when the sweeps have finished. */
int piqmc_qa_colour(piqmc_handle h, const double *sched, int nsched, int mcsteps, float temp,
                    uint64_t seed, uint32_t replica0, uint32_t sweep0, int trotter,
                    const int32_t *orders);
/* The AS-SHIPPED semantics of qmc.QuantumAnneal (piqmc/qmc.pyx:98-136) on the resident state: the running
 * energy difference is reset once per slice sweep, not per spin, so within a slice it is a float32 carry over
 * all spins visited so far (the reference's default function behaves like this; its `_parallel` variant and
 * piqmc_qa_colour reset per spin).  Trotter neighbours slices-1 and 1 for every slice, `> 0` shortcut,
 * orders int32[nsched*mcsteps][nspins] = the visiting order of every sweep, shared by all replicas (NULL:
 * 0..N-1), this library's Philox uniforms.  One warp per replica; any maxnb; one replica per word.
 * Specification: oracle_qa_carry (oracle/piqmc_oracle.c part 4). */
int piqmc_qa_carry(piqmc_handle h, const double *sched, int nsched, int mcsteps, float temp,
                   uint64_t seed, uint32_t replica0, uint32_t sweep0, const int32_t *orders);
/* Classical SA sweeps over the resident state (lanes = 64 replicas per word, sa.Anneal rules). */
int piqmc_sa_colour(piqmc_handle h, const double *sched, int nsched, int mcsteps,
                    uint64_t seed, uint32_t row0, uint32_t sweep0, const int32_t *orders);
/* World-line (global) moves for piqmc_qa_colour: after the local moves of a spin, attempt to flip
 * it in ALL slices at once (the Trotter terms cancel; ediff = slice-summed in-slice term, same
 * Metropolis rule).  A capability the reference does not have (BASELINE.json north_star); off by
 * default; needs maxnb <= 4.  Specification: oracle/piqmc_oracle.c, oracle_qa_colour. */
int piqmc_set_global_moves(piqmc_handle h, int enable);
/* kernel variant selection for piqmc_qa_colour / piqmc_sa_colour: 0 = auto, 1 = generic,
 * 2 = dataflow kernel (one unit per (sweep, spin, row chunk); falls back to generic when the graph
 * does not qualify), 3 = chain pipeline (natural-order colourings with maxnb <= 4; falls back to 2) */
int piqmc_set_variant(piqmc_handle h, int variant);
/* Chain pipeline (the natural-order sweep of qmc.pyx:320-357 on a 2-D lattice, cut into chains of
 * chain_len consecutive spins -- a lattice row --, one warp per chain and 32 or 64 rows): chain_len = 0
 * lets the library choose; > 0 forces that length and the pipeline (testing, tuning).
 * piqmc_chain_info reports the plan of the current graph + colouring: chain_len = 0 when there is none
 * (the colouring is not the natural order's, maxnb > 4, or some coupled pair is neither consecutive in
 * a chain, nor the two ends of a chain, nor the same position of adjacent chains / of the first and
 * last chain); period = pipeline steps per sweep; selected = 1 when a static-colouring run (reference
 * Trotter neighbours, no world-line moves) would take the pipeline under the current variant and state. */
int piqmc_set_chain(piqmc_handle h, int chain_len);
int piqmc_chain_info(piqmc_handle h, int *chain_len, int *nchains, double *period, int *selected);
/* The plan itself, on the host (no device needed): for the natural-order sweep of the ELL table
 * (idx, J) cut into chains of chain_len spins (0 = choose), kinds[i*4 + k] is where sorted table
 * column k of spin i (columns sorted by |J| descending, stable) gets its neighbour word from:
 * 0 none, 1 the word this chain wrote one step earlier, 2 the own-row word one step later, 3 the
 * same position of the preceding chain (chain 0 of a torus: the last chain, previous sweep), 4 the
 * same position of the following chain (last chain of a torus: chain 0, this sweep).  *wrap = 1
 * when the first and last chain are coupled.  Returns the chain length (0: no plan, < 0: error). */
int piqmc_chain_plan(int nspins, int maxnb, const int32_t *idx, const double *J, int chain_len,
                     uint8_t *kinds, int *wrap);

/* sa.ClassicalIsingEnergy (piqmc/sa.pyx:25-44) of every (row, lane) of the resident state,
 * float64: energies[row*lanes + lane].  energies may be NULL (result stays on the device). */
int piqmc_energy(piqmc_handle h, double *energies);

/* energies (as piqmc_energy) and packed words (as piqmc_state_download_words) in one call; the
 * download of the words runs on a second stream and overlaps the energy reduction. */
int piqmc_results(piqmc_handle h, double *energies, uint64_t *words);

/* piqmc_qa_colour followed by piqmc_results, as one call -- what a caller of the reference's
 * qmc.QuantumAnneal_parallel gets back (the annealed configurations, piqmc/qmc.pyx:247-357) plus the
 * per-slice energies its drivers compute next (examples/spinglass32.py:150-160).  Same results as the two
 * calls.  Opt-in (environment PIQMC_PIPE_LAG16 = stagger between consecutive row chunks in 1/16 sweeps): with the
 * dataflow kernel and a static colouring the row chunks of the state (512 replicas each) are staggered inside the
 * one launch, report to the host when they are final (mapped flags) and are downloaded while the remaining
 * chunks still sweep (words should be page-locked, piqmc_host_alloc).  Same results bit for bit; measured on B200
 * it does not end earlier than the plain sequence (profiles/r2_e2e_overlap.md section 1), which is therefore what
 * the call runs by default.  piqmc_pipelined_runs: calls that took the overlapped path so far. */
int piqmc_qa_colour_results(piqmc_handle h, const double *sched, int nsched, int mcsteps, float temp,
                            uint64_t seed, uint32_t replica0, uint32_t sweep0, int trotter,
                            const int32_t *orders, double *energies, uint64_t *words);
uint64_t piqmc_pipelined_runs(piqmc_handle h);
/* Host seconds the last piqmc_qa_colour_results spent in the sweeps and in energies + download (the second is ~0
 * when the overlapped path ran: the downloads then happen inside the first). */
int piqmc_last_phase_seconds(piqmc_handle h, double *sweeps, double *results);

/* Histogram of the energies of the last piqmc_energy / piqmc_results on the device (the residual-energy
 * statistics of examples/santoro80.py:290-323 without moving R x slices doubles to the host): one value per
 * replica -- reduce 0: mean over its slices, 1: its lowest slice, 2: every slice counts on its own --, binned
 * as floor(((E - e0) * scale - lo) / (hi - lo) * nbins); counts[nbins] in range, counts[nbins] below lo,
 * counts[nbins + 1] at or above hi; stats[0..2] = sum, minimum and maximum of (E - e0) * scale. */
int piqmc_energy_histogram(piqmc_handle h, int reduce, double e0, double scale, double lo, double hi, int nbins,
                           uint64_t *counts, double *stats);

/* ClassicalIsingEnergy for host configurations: J given as nnz COO triples holding each bond
 * once (diagonal = local fields); spins[c*nspins + i] +-1; energies[c]. */
int piqmc_energy_coo(piqmc_handle h, int nspins, int nnz, const int32_t *row, const int32_t *col,
                     const double *val, int nconfs, const int8_t *spins, double *energies);

#ifdef __cplusplus
}
#endif
#endif /* PIQMC_B200_H */
