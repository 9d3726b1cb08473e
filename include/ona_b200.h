/* ona_b200.h -- C ABI of the B200-native OptNetAlign scoring engine.
 *
 * Drop-in boundary for the reference's data-parallel hot path.  The reference
 * (crclark/optnetaligncpp) has no FFI layer: its "interface" for this path is the
 * C++ header pair Alignment.h + localSearch.h resolved at link time.  Every entry
 * point below names the reference interface it replaces (file:line relative to the
 * reference tree); INTEGRATION.md shows the replacement Alignment.cpp /
 * localSearch.cpp translation units (optnetaligncpp_b200/host/) that implement the
 * unchanged headers on top of this ABI, so steadystate.cpp / nsgaiimain.cpp /
 * hillclimber.cpp link unchanged.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every buffer is caller-owned HOST memory unless
 *     the parameter name ends in _dev (then it is a CUDA device pointer on the engine's
 *     device); ints are 32-bit node ids (Network.h:11 "typedef int node"), scores fp64.
 *   - every function returns ONA_OK or a negative error code; ona_last_error() returns a
 *     thread-local message (the reference throws C++ exceptions -- LineReadException,
 *     ArgError, std::out_of_range -- and nothing may throw across a C boundary; the C++
 *     host shim turns non-zero codes back into std::runtime_error).
 *   - there is NO CPU fallback: ona_engine_create fails if no CUDA device is usable.
 *   - an engine is bound to one GPU and one CUDA stream (ona_set_stream); calls on one
 *     engine are serialised by an internal mutex, so the reference's many host threads
 *     (steadystate.cpp:132-303) may share it.
 *   - population slots: the engine stores P individuals ("slots"); host Alignment objects
 *     stay the source of truth between batched calls and are keyed by slot index.
 */
#ifndef ONA_B200_H
#define ONA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ONA_OK 0
#define ONA_ERR_INVALID (-1) /* bad argument / malformed graph or permutation */
#define ONA_ERR_CUDA (-2)    /* CUDA runtime failure (message has the CUDA error string) */
#define ONA_ERR_STATE (-3)   /* call order problem (e.g. population not uploaded) */
#define ONA_ERR_NOMEM (-4)

/* objective ids == the reference's enum fitnessName (Network.h:31-32), same values */
enum ona_objective {
    ONA_ICS = 0, ONA_EC = 1, ONA_BITSCORE_SUM = 2, ONA_EVALS_SUM = 3, ONA_SIZE = 4,
    ONA_GOC = 5, ONA_S3 = 6, ONA_S3_DENOM = 7, ONA_ICS_TIMES_EC = 8, ONA_S3_VARIANT = 9
};
#define ONA_MAX_OBJECTIVES 10

/* BLASTDict / GOCDict (blastinfo.h:10, goc.h:12) in one of three forms:
 *   dense      the row-major n1 x n2 matrix the reference holds (pass the multi_array's data());
 *   COO        its non-zeros as triples (later duplicates win, like the reference's assignment
 *              loop blastinfo.cpp:42);
 *   term sets  (GOC only) the annotation sets loadAnnotations returns (goc.cpp:55-89) as CSR:
 *              ann_rowptr1/ann_terms1 for net1, ann_rowptr2/ann_terms2 for net2, every row sorted
 *              ascending without duplicates (std::set<int> order).  The engine then builds the
 *              dense Jaccard matrix double(|A n B|) / double(|A u B|) of loadGOC (goc.cpp:16-52)
 *              on the device, replacing the reference's O(n1 n2 terms) host loop and the n1 x n2
 *              host matrix.  A node is "annotated" when it has a row in the reference's map:
 *              ann_present1/2 (bytes, may be NULL = "row non-empty") carry that; pairs with an
 *              unannotated node stay 0, two annotated-but-empty nodes give 0/0 = NaN as in goc.cpp:47.
 * All-NULL = table absent (the reference's nullptr dict). */
typedef struct ona_score_table {
    const double* dense;
    const int* rows;
    const int* cols;
    const double* vals;
    long long nnz;
    const int* ann_rowptr1;
    const int* ann_terms1;
    const int* ann_rowptr2;
    const int* ann_terms2;
    const unsigned char* ann_present1;
    const unsigned char* ann_present2;
} ona_score_table;

/* Replaces: the immutable inputs an Alignment borrows -- const Network* net1/net2,
 * const BLASTDict* bitscores, const GOCDict* gocs (Alignment.h:73-98) -- plus the
 * fitnessNames vector every computeFitness/correctHillClimb call receives
 * (argParsing.h:212-271 fixes its order).
 *   rowptr/col : CSR of Network::adjList in adjList order (Network.cpp:89-92: a self-loop
 *                is listed once).  Duplicate edge lines ("a b" and "b a", repeated lines) are
 *                accepted and keep their copies, as adjList does: the incremental counters
 *                (conservedCount / inducedCount, :791-826) and fastICSDenominator (:667-690) weigh
 *                such an edge by its multiplicity while edges.size(), adjMatrix and the set-based
 *                ics() (:604-661) do not -- mirrored exactly.
 *   num_edges1 : net1->edges.size(), the EC / S3 denominator term (Alignment.cpp:536,541). */
typedef struct ona_problem {
    int n1, n2; /* |V(net1)| <= |V(net2)| (argParsing.h:207-210) */
    long long num_edges1;
    const int* rowptr1;
    const int* col1;
    const int* rowptr2;
    const int* col2;
    ona_score_table bitscores;
    ona_score_table gocs;
    const int* objectives; /* ona_objective values, in fitnessNames order */
    int num_objectives;
} ona_problem;

typedef struct ona_engine ona_engine;

/* ------------------------------------------------------------------ lifecycle */
int ona_device_count(void);
int ona_engine_create(int device, const ona_problem* problem, ona_engine** out);
void ona_engine_destroy(ona_engine* e);
/* A second command queue on the same problem: the fork shares the parent's immutable device data
 * (graph layouts, score tables -- nothing is copied or rebuilt) and has its own population slots,
 * scratch, stream and lock.  Replaces nothing in the reference; it is what lets the short
 * per-alignment calls of one host thread (a child's crossover / mutation / evaluation,
 * steadystate.cpp:211-232) run beside a long batched hill-climb launch of other threads instead of
 * queueing behind it.  Destroy parent and forks in any order. */
int ona_engine_fork(ona_engine* parent, ona_engine** out);
const char* ona_last_error(void);
/* run on an existing stream (e.g. torch.cuda.current_stream().cuda_stream); NULL = own stream */
int ona_set_stream(ona_engine* e, void* cuda_stream);
int ona_synchronize(ona_engine* e);
/* pinned host memory for the host-buffer entry points (optional; any host memory works) */
void* ona_host_alloc(long long bytes);
void ona_host_free(void* p);
/* number of engine kernels launched so far (bench.py's gpu_launches) */
long long ona_kernel_launches(const ona_engine* e);
/* Replaces: passing a different fitnessNames vector to computeFitness / correctHillClimb
 * (Alignment.cpp:559, localSearch.cpp:18): switch the objective list of an existing engine.
 * Cached objective rows of all slots are recomputed from their counters. */
int ona_set_objectives(ona_engine* e, const int* objectives, int num_objectives);
/* global index of slot 0 (island offset); only keys the on-device proposal RNG */
int ona_set_island_offset(ona_engine* e, long long first_global_individual);

/* ----------------------------------------------------------------- population */
/* Replaces: Alignment::aln / alnInv / alnMask state (Alignment.h:44-46).  aln is
 * [count][n2] (a permutation of V2 per row; positions >= n1 are the dummy nodes of
 * Alignment.cpp:25-34), mask is [count][n2] bytes or NULL for all-true (--total).
 * alnInv is rebuilt on the device.  Counters are NOT touched: follow with
 * ona_full_fitness(seed_mode != 0) or ona_pop_set_counters.
 * ona_pop_resize keeps the state of slots [0, min(old, new)) (permutation, mask, counters,
 * objective rows); slots added by a grow are zero-filled and must be uploaded before any
 * scoring call touches them.  On allocation failure the old population stays valid. */
int ona_pop_resize(ona_engine* e, int num_slots);
int ona_pop_upload(ona_engine* e, int first, int count, const int* aln, const unsigned char* mask);
int ona_pop_download(ona_engine* e, int first, int count, int* aln, int* aln_inv,
                     unsigned char* mask);
/* Replaces: direct access to currConservedCount / currInducedCount / currBitscore /
 * currGOC / fitness (Alignment.h:61,73,80,101,109).  Any output may be NULL. */
int ona_pop_set_counters(ona_engine* e, int first, int count, const int* conserved,
                         const int* induced, const double* bitscore, const double* goc);
int ona_pop_get_counters(ona_engine* e, int first, int count, int* conserved, int* induced,
                         double* bitscore, double* goc, double* size, double* fitness);

/* --------------------------------------------------------------- full fitness */
#define ONA_SEED_NONE 0    /* only report, leave the slots' cached counters alone */
#define ONA_SEED_TRUE 1    /* cache the exact conserved count */
#define ONA_SEED_REFCOMPAT 2 /* cache (int)(ics()*denom), the reference's truncated seed (H1) */
#define ONA_SUM_SEQUENTIAL 16 /* OR into seed_mode: fp64 sums in the reference's index order
                                 (bit-identical to sumBLAST/sumGOC) instead of a tree sum */
/* Replaces, population-batched: the from-scratch scoring idiom at the tail of
 * Alignment::shuf and the constructors (Alignment.cpp:437-446, 43-46, 129-137): sumBLAST
 * (:701-710), sumGOC (:712-722), fastICSDenominator (:667-690), ics() (:604-661), then
 * computeFitness (:559-602).  Outputs are [count] ([count][M] for fitness), any may be NULL:
 * conserved_true = the exact value of the counter doSwap maintains (|mapped E1 /\ induced E2|;
 * with duplicate edge lines each edge weighted by its adjList multiplicity), induced =
 * fastICSDenominator, ics = ics() (ratio of the two edge SETS), conserved_refseed =
 * (int)(ics * induced), the reference's seed value. */
int ona_full_fitness(ona_engine* e, int first, int count, int seed_mode, int* conserved_true,
                     int* conserved_refseed, int* induced, double* bitscore, double* goc,
                     double* size, double* fitness, double* ics);
/* same, nothing leaves the device (results stay in the slots / the device fitness matrix) */
int ona_full_fitness_dev(ona_engine* e, int first, int count, int seed_mode);

/* ---------------------------------------------------------------- swap deltas */
/* Replaces, batched and state-preserving: Alignment::doSwap + computeFitness + the accept
 * test of correctHillClimb (Alignment.cpp:491-520, 559-602; localSearch.cpp:32-47) and
 * swapNormalizedDelta (localSearch.cpp:88-120) for K candidate swaps per individual, all
 * evaluated against the slot's CURRENT state (which is not modified).
 *   xy       [count][K][2] proposals (x,y): positions in V1 u dummies, x != y not required
 *   obj      -1: accept iff no objective decreased (localSearch.cpp:36-41);
 *            j>=0: accept iff objective j strictly increased (:43-47)
 * outputs (any may be NULL), all [count][K] ([count][K][M] for fitness): the values the
 * reference's Alignment holds right after doSwap+computeFitness: currConservedCount,
 * currInducedCount, currBitscore, currGOC, fitness[]; improved = the accept decision;
 * snd = swapNormalizedDelta's return value. */
int ona_swap_deltas(ona_engine* e, int first, int count, int K, const int* xy, int obj,
                    int* conserved, int* induced, double* bitscore, double* goc,
                    double* fitness, unsigned char* improved, double* snd);
/* Proposals drawn on the device from a counter-based RNG keyed (seed, global individual,
 * candidate index), supports x in U[0,n1), y in U[0,n2) \ {x} as localSearch.cpp:21-28.
 * Results stay on the device; summary outputs (host, [count], may be NULL):
 * num_improved = how many of the K candidates would be accepted, best_snd / best_xy =
 * arg-max swapNormalizedDelta and its (x,y) (the steepestAscentHillClimb reduction,
 * localSearch.cpp:136-147, over a random candidate subset). */
int ona_swap_deltas_rng(ona_engine* e, int first, int count, int K, unsigned long long seed,
                        int obj, int* num_improved, double* best_snd, int* best_xy);
/* Device-resident form: proposals from xy_dev ([count][K][2] on the device) or, when xy_dev
 * is NULL, from the device RNG; every output pointer is a DEVICE buffer (or NULL) laid out as
 * in ona_swap_deltas; num_improved_dev is [count] and is accumulated into (zero it first).
 * Asynchronous on the engine's stream; nothing is copied or synchronised. */
int ona_swap_deltas_dev(ona_engine* e, int first, int count, int K, const int* xy_dev,
                        unsigned long long seed, int obj, int* conserved_dev, int* induced_dev,
                        double* bitscore_dev, double* goc_dev, double* fitness_dev,
                        unsigned char* improved_dev, double* snd_dev, int* num_improved_dev);
/* Replaces: one scan of steepestAscentHillClimb (localSearch.cpp:130-188): swapNormalizedDelta
 * of EVERY (x,y) in [0,n1) x [0,n2) against slot `slot` (all objectives), arg-max with the
 * reference's tie rule (first maximum in x-major order; only values > -1 can win, :134,141).
 * No state is modified; best_x = best_y = -1 and best_delta = -1 when nothing wins. */
int ona_steepest_step(ona_engine* e, int slot, double* best_delta, int* best_x, int* best_y);
/* the proposals the device RNG generates, for host-side checking */
int ona_rng_proposals(ona_engine* e, int first, int count, int K, unsigned long long seed,
                      long long start_index, int* xy);

/* ----------------------------------------------------------------- hill climb */
/* Replaces: correctHillClimb (localSearch.cpp:18-55; decl localSearch.h:9-11), for `count`
 * individuals at once, with the reference's exact sequential semantics per individual:
 * proposals are consumed in order, each is applied, scored, and undone unless accepted;
 * integer counters, accept decisions and the fp64 accumulators (including the not-quite-
 * inverse undo) equal what the reference produces for the same proposals.
 *   xy        [count][iters][2] proposals, or NULL to draw them on the device (seed)
 *   objs      [count] objective index per individual (-1 = all, as above), or NULL = all -1
 *   accepted  [count][iters] bytes out, or NULL
 *   num_accepted [count] out, or NULL
 * Precondition (as the reference, steadystate.cpp:232): the slots' counters are valid. */
int ona_hill_climb(ona_engine* e, int first, int count, long long iters, const int* xy,
                   unsigned long long seed, const int* objs, unsigned char* accepted,
                   long long* num_accepted);

/* Same, the proposals of individual i read from its own host array xy_lists[i] ([iters][2]): the
 * form in which the reference's callers have them, one correctHillClimb call = one PRNG = one
 * list per alignment (localSearch.cpp:18-28), without gathering them into one array first. */
int ona_hill_climb_lists(ona_engine* e, int first, int count, long long iters, const int* const* xy_lists,
                         const int* objs, unsigned char* accepted, long long* num_accepted);

/* Same, the proposals generated ON THE DEVICE as the reference's own stream (localSearch.cpp:21-28):
 *     x = uniform_int_distribution<int>(0, n1 - 1)(prng);  y = uniform_int_distribution<int>(0, n2 - 2)(prng);
 *     if (y >= x) y++;
 * for prng a std::mt19937 (RandGenT, Network.h:10) and GNU libstdc++'s distributions (GCC >= 11:
 * Lemire's multiply-shift with rejection, bits/uniform_int_dist.h) -- so neither 2*iters host draws nor
 * an 8-byte-per-iteration upload happen per call.
 *   mt_states  [count][625] in/out: per individual the generator's 624 state words followed by its
 *              position (0..624), exactly the numbers operator<<(ostream&, mt19937) prints; on return
 *              the state the caller's generator is in after the 2*iters (+ redrawn) draws, to be read
 *              back with operator>>.
 * Trajectory, counters and accumulators are those of ona_hill_climb fed with the host-drawn list. */
int ona_hill_climb_mt(ona_engine* e, int first, int count, long long iters, unsigned int* mt_states,
                      const int* objs, unsigned char* accepted, long long* num_accepted);
/* the proposals alone ([count][iters][2] out) for explicit ranges n1 >= 1, n2 >= 2, states advanced as
 * above -- for checking the stream against the host's libstdc++ */
int ona_mt_proposals(ona_engine* e, int n1, int n2, int count, long long iters, unsigned int* mt_states,
                     int* xy);

/* Throughput form of the hill climb (north_star item 3: "thousands of candidates per individual
 * at once; accepted swaps are applied in-kernel").  Per individual, rounds of 1024 proposals
 * are scored against the round's base state; each is accepted or rejected against the base
 * objectives exactly like one correctHillClimb iteration would (localSearch.cpp:32-47); a
 * conflict-free subset of the accepted ones (deterministic rule: include/..., csrc kernel
 * header) is applied and the next round starts from the new state.  Integer counters stay exact
 * (equal to a from-scratch count); fp64 accumulators agree with sequential application to
 * <= 1e-9 relative.  NOT trajectory-identical to the reference: proposals that conflict inside
 * a round are dropped.  Arguments as ona_hill_climb; num_applied [count] out (may be NULL). */
int ona_hill_climb_batched(ona_engine* e, int first, int count, long long iters, const int* xy,
                           unsigned long long seed, const int* objs, long long* num_applied);

/* Replaces: a run of unconditional Alignment::doSwap calls -- the swap part of
 * Alignment::mutate (Alignment.cpp:456-465) and of UPMX crossover (:197-252) given a
 * host-generated ordered swap list.  xy is [count][nswaps][2]; entries with x < 0 are
 * skipped (ragged lists).  Counters/accumulators are updated incrementally like doSwap. */
int ona_apply_swaps(ona_engine* e, int first, int count, int nswaps, const int* xy);

/* Replaces: the variation operators' edits of PARTIAL alignments (no --total), which are not pure
 * doSwap runs -- Alignment::mutate's mask flip (Alignment.cpp:468-480), Alignment::onBit
 * (:522-533) and the UPMX step with its mask AND/OR (:197-252) -- replayed in order with the
 * reference's own conservedCount / inducedCount / updateBitscore / updateGOC call pattern, so
 * counters and fp64 accumulators match the reference bit for bit.
 * ops is [count][nops][4] = {kind, a, b, c}; kind < 0: skipped (ragged lists).
 *   ONA_EDIT_SWAP  (a = x, b = y)          doSwap(x, y), :491-520
 *   ONA_EDIT_FLIP  (a = i)                 alnMask[i] = !alnMask[i] with mutate's updates
 *   ONA_EDIT_ONBIT (a = x)                 onBit(x)
 *   ONA_EDIT_CROSS (a = i, b = par2.aln[i], c = par2.alnMask[i] | coinFlip << 1 | total << 2)
 *                                          position i takes parent 2's image, the displaced image
 *                                          moves to where that image was (:197-252) */
#define ONA_EDIT_SWAP 0
#define ONA_EDIT_FLIP 1
#define ONA_EDIT_ONBIT 2
#define ONA_EDIT_CROSS 3
int ona_apply_edits(ona_engine* e, int first, int count, int nops, const int* ops);

/* ---------------------------------------------------------------- GOC matrix
 * Replaces: the pair loop of loadGOC (goc.cpp:23-50; 31.8 s on one host thread for the bundled example pair):
 * the dense n1 x n2 matrix of double(|A n B|) / double(|A u B|) over the annotation sets loadAnnotations
 * returns (goc.cpp:55-89), computed on the GPU and written to the caller's row-major host matrix `out`
 * (e.g. the data() of the GOCDict a replacement loadGOC returns).  Term sets as CSR, rows sorted ascending
 * without duplicates (std::set<int> order); present1/2[u] != 0 iff node u has an entry in the reference's
 * map (NULL: "row non-empty"); pairs with an absent node stay 0.0, two present-but-empty sets give 0/0 = NaN
 * as goc.cpp:47 does.  Needs no engine; values are bit-identical to the reference's. */
int ona_goc_matrix(int device, int n1, int n2, const int* ann_rowptr1, const int* ann_terms1,
                   const unsigned char* ann_present1, const int* ann_rowptr2, const int* ann_terms2,
                   const unsigned char* ann_present2, double* out);

/* ------------------------------------------------------------ measurement aid
 * Replaces nothing in the reference.  Measures, on the engine's GPU and stream, how many scattered
 * 4-byte reads per second (one 32-byte sector each, 8 independent loads in flight per thread, all SMs)
 * the memory system serves from a table of table_bytes (rounded down to a power of two; sized like the
 * engine's L2-resident structures it stays in L2).  This is the request pattern of the engine's gathers
 * -- adjacency words, node words, signatures, hash slots -- and so the ceiling bench.py sets the
 * swap-delta kernel's L2 sector rate against.  Best of 4 timed launches after 2 warm-up ones. */
int ona_probe_gather_rate(ona_engine* e, long long table_bytes, double* sectors_per_s, double* ms);

/* -------------------------------------------------------------- device access */
/* copy the [count][M] objective rows of slots first..first+count into a device buffer
 * (e.g. a torch tensor's data_ptr) on the engine's stream -- the send buffer of the
 * per-generation all-gather that feeds the host NSGA-II sort (nsga-ii.cpp:14-118). */
int ona_export_fitness_dev(ona_engine* e, int first, int count, double* dst_dev);

/* -------------------------------------------------------------- NSGA-II ranking
 * Replaces, for one [P][M] matrix of objective rows (larger is better), the sequence
 * nsgaiimain.cpp:157-175 runs every generation:
 *     nonDominatedSort(all)          nsga-ii.cpp:14-66   (dominates, :121-134)
 *     normalizeFitnesses(all)        nsga-ii.cpp:68-84
 *     setCrowdingDists(front)        nsga-ii.cpp:89-118  on every front
 * P and M are independent of the engine's problem and population (P <= 2^20, M <= 16).
 *   fitness        [P][M]   in
 *   has_norm       [P]      in, nullable: individual i already carries normalised values
 *                           (norm_io row i) from an earlier generation -- normalizeFitnesses
 *                           only appends (:79), so setCrowdingDists keeps using those
 *   norm_io        [P][M]   in/out, nullable unless has_norm: fitnessNormalized[0..M)
 *   dom_rank       [P]      out: Alignment::domRank
 *   front_order    [P]      out: the fronts nonDominatedSort returns, flattened, members in the
 *                           reference's order;  front f = [front_offsets[f], front_offsets[f+1])
 *   front_offsets  [P+1]    out (n_fronts+1 entries are written);  n_fronts: out, required
 *   order_after    [P]      out: the same fronts after setCrowdingDists' in-place sorts (ties are
 *                           placed as libstdc++'s std::sort places them)
 *   crowd_dist     [P]      out: Alignment::crowdDist (boundary members: DBL_MAX, :104-105)
 * Every out pointer except n_fronts may be NULL.  Fitness values must not be NaN (the
 * reference's std::sort is undefined on them).  Synchronous. */
int ona_nsga2_rank(ona_engine* e, int P, int M, const double* fitness, const unsigned char* has_norm,
                   double* norm_io, int* dom_rank, int* front_order, int* front_offsets, int* n_fronts,
                   int* order_after, double* crowd_dist);
/* Same with every array in device memory (e.g. the all-gathered objective rows of all islands);
 * n_fronts stays a host pointer.  Runs on the engine's stream and returns after the last kernel
 * has been launched; the peel loop synchronises with the host to learn the number of fronts. */
int ona_nsga2_rank_dev(ona_engine* e, int P, int M, const double* fitness_dev,
                       const unsigned char* has_norm_dev, double* norm_dev, int* dom_rank_dev,
                       int* front_order_dev, int* front_offsets_dev, int* n_fronts,
                       int* order_after_dev, double* crowd_dist_dev);

#ifdef __cplusplus
}
#endif
#endif /* ONA_B200_H */
