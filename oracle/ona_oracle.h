/* TEST INFRASTRUCTURE ONLY -- the CPU oracle for the OptNetAlign hot path.
 *
 * Plain-C restatement of the reference's scoring core (Alignment.cpp) and hill
 * climber (localSearch.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker or as the
 * reported CPU baseline -- never on the product path (optnetaligncpp_b200/ does not
 * import, link or execute anything under oracle/).
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function
 * here against the unmodified reference compiled into oracle/_ref/libona_ref.so
 * (when present) and tests/test_oracle_golden.py checks it against the golden
 * vectors under tests/golden/ that tests/golden/make_golden.py generated from that
 * same compiled reference.
 */
#ifndef ONA_ORACLE_H
#define ONA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* fitnessName enum, Network.h:31-32 (same numeric values) */
enum {
    ORC_ICS = 0, ORC_EC = 1, ORC_BITSCORE_SUM = 2, ORC_EVALS_SUM = 3, ORC_SIZE = 4,
    ORC_GOC = 5, ORC_S3 = 6, ORC_S3_DENOM = 7, ORC_ICS_TIMES_EC = 8, ORC_S3_VARIANT = 9
};
#define ORC_MAX_OBJ 10

typedef struct OrcProblem OrcProblem; /* two networks + two dense score matrices */
typedef struct OrcAln OrcAln;         /* one Alignment's state (Alignment.h:44-111) */

/* rp/col: CSR of Network::adjList in adjList order (self-loop listed once,
 * Network.cpp:89-92).  E1 = net1->edges.size().  S,G: dense row-major n1 x n2
 * doubles or NULL (BLASTDict/GOCDict, blastinfo.h:10, goc.h:12); borrowed. */
OrcProblem* orc_problem_create(int n1, int n2, long E1, const int* rp1, const int* col1,
                               const int* rp2, const int* col2, const double* S, const double* G);
void orc_problem_free(OrcProblem*);

/* aln[n2] permutation, mask[n2] (NULL = all true).  Seeds the counters like
 * Alignment.cpp:437-446, including the truncated conserved seed (hazard H1). */
OrcAln* orc_aln_create(const OrcProblem*, const int* aln, const uint8_t* mask);
OrcAln* orc_aln_clone(const OrcAln*);
void orc_aln_free(OrcAln*);
void orc_aln_get_state(const OrcAln*, int* aln, int* alnInv, uint8_t* mask);
void orc_aln_get_counters(const OrcAln*, int* cons, int* ind, double* bits, double* goc);
void orc_aln_set_counters(OrcAln*, int cons, int ind, double bits, double goc);

/* from-scratch pieces */
long orc_conserved_true(const OrcAln*);        /* |mapped E1 /\ induced E2| (numerator of ics() :604-661), every
                                                  adjList copy of a duplicate edge line counted: the exact
                                                  value of the counter doSwap maintains */
long orc_induced_true(const OrcAln*);          /* fastICSDenominator :667-690 */
double orc_ics(const OrcAln*);                 /* :604-661 */
int orc_refseed_conserved(long C, long I);     /* (int)(ics()*denom) :445 */
int orc_refseed_from_ics(double ics, long denom);
double orc_sum_blast(const OrcAln*);           /* :701-710 */
double orc_sum_goc(const OrcAln*);             /* :712-722 */
double orc_aln_size(const OrcAln*);            /* :724-732 */

/* incremental pieces */
int orc_conserved_count(const OrcAln*, int n1, int n2, int mask, int ignore); /* :791-808 */
int orc_induced_count(const OrcAln*, int n2, int ignore);                     /* :812-826 */
void orc_do_swap(OrcAln*, int x, int y);                                      /* :491-520 */
void orc_compute_fitness(OrcAln*, const int* names, int M, double* out);      /* :559-602 */
/* ordered {kind,a,b,c} edits: doSwap / mutate's mask flip :468-480 / onBit :522-533 / UPMX step
 * :197-252 (see ona_oracle.c) */
void orc_apply_edits(OrcAln*, const int* ops, int nops);

/* localSearch.cpp */
void orc_hill_climb_replay(OrcAln*, const int* xy, long iters, const int* names, int M, int obj,
                           uint8_t* accepted);                                /* :23-54 */
double orc_swap_normalized_delta(OrcAln*, const int* names, int M, int x, int y); /* :88-120 */
void orc_swap_delta_records(OrcAln*, const int* names, int M, int obj, const int* xy, long n,
                            int* cons, int* ind, double* bits, double* goc, double* fit,
                            uint8_t* improved, double* snd);

/* CPU baselines ("port" kind): same worker patterns as oracle/ref_shim.cpp */
double orc_time_swap_deltas(OrcAln** alns, int nalns, const int* names, int M, const int* xy,
                            long n, int nthreads, double* sink);
double orc_time_hill_climb(OrcAln** alns, int nalns, const int* xy, long iters, const int* names,
                           int M, int obj, int nthreads);
double orc_time_full_fitness(OrcAln** alns, int nalns, const int* names, int M, long reps,
                             int nthreads, double* sink);

/* the proposal stream of localSearch.cpp:21-28 (std::mt19937 + libstdc++'s uniform_int_distribution);
 * state = 624 words + position, advanced in place */
void orc_mt_seed(uint32_t seed, uint32_t* state);
void orc_mt_xy_stream(uint32_t* state, int n1, int n2, long iters, int* xy);

/* NSGA-II ranking of the objective rows (ona_oracle_nsga2.c; nsga-ii.cpp:14-134).
 * fitness [P][M]; has_norm [P] or NULL; norm_io [P][M] in/out; dom_rank, front_order,
 * order_after, crowd_dist [P]; front_offsets [P+1].  See oracle/ref_shim.cpp:ref_nsga2. */
int orc_dominates(const double* f1, const double* f2, int M);                 /* :121-134 */
void orc_gnu_sort_by_key(int n, const double* keys, int* idx);                /* std::sort, libstdc++ */
void orc_gnu_heap_sort_by_key(int n, const double* keys, int* idx);           /* its depth-limit fallback */
int orc_nsga2(int P, int M, const double* fitness, const unsigned char* has_norm, double* norm_io,
              int* dom_rank, int* front_order, int* front_offsets, int* n_fronts,
              int* order_after, double* crowd_dist);

#ifdef __cplusplus
}
#endif
#endif
