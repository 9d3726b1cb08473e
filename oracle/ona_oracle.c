/* TEST INFRASTRUCTURE ONLY -- see ona_oracle.h.  Plain-C restatement of the
 * reference hot path; every function cites the reference lines it follows
 * (paths relative to /root/reference).  Compiled with -ffp-contract=off so that
 * fp64 expressions keep the reference's (unfused) operation order. */
#define _POSIX_C_SOURCE 200809L
#include "ona_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

struct OrcProblem {
    int n1, n2;
    long E1;
    const int *rp1, *col1, *rp2, *col2; /* owned copies */
    uint32_t* adj2;                     /* n2 x words bit matrix == Network::adjMatrix of net2 */
    int words;
    uint8_t *first1, *first2; /* per CSR entry: first copy of that column in its row.  Duplicate edge
                                 lines stay in adjList (Network.cpp:89-92) but not in the `edges`
                                 set (:87) that ics() iterates */
    const double *S, *G; /* borrowed dense n1 x n2 */
};

struct OrcAln {
    const OrcProblem* p;
    int *aln, *alnInv;
    uint8_t* mask;
    int cons, ind;    /* currConservedCount, currInducedCount (Alignment.h:101,109) */
    double bits, goc; /* currBitscore, currGOC (Alignment.h:73,80) */
    double fitness[ORC_MAX_OBJ];
    int M;
};

static int* dup_ints(const int* a, long n) {
    int* r = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    memcpy(r, a, sizeof(int) * (size_t)n);
    return r;
}

static uint8_t* first_copies(int n, const int* rp, const int* col) {
    uint8_t* f = (uint8_t*)malloc((size_t)(rp[n] > 0 ? rp[n] : 1));
    for (int u = 0; u < n; ++u)
        for (int k = rp[u]; k < rp[u + 1]; ++k) {
            f[k] = 1;
            for (int j = rp[u]; j < k; ++j)
                if (col[j] == col[k]) { f[k] = 0; break; }
        }
    return f;
}

/* Network.cpp:99-109 -- adjMatrix[i][j] = true for every adjList entry */
OrcProblem* orc_problem_create(int n1, int n2, long E1, const int* rp1, const int* col1,
                               const int* rp2, const int* col2, const double* S, const double* G) {
    OrcProblem* p = (OrcProblem*)calloc(1, sizeof(OrcProblem));
    p->n1 = n1; p->n2 = n2; p->E1 = E1;
    p->rp1 = dup_ints(rp1, n1 + 1); p->col1 = dup_ints(col1, rp1[n1]);
    p->rp2 = dup_ints(rp2, n2 + 1); p->col2 = dup_ints(col2, rp2[n2]);
    p->words = (n2 + 31) / 32;
    p->adj2 = (uint32_t*)calloc((size_t)n2 * (size_t)p->words, sizeof(uint32_t));
    for (int i = 0; i < n2; ++i)
        for (int k = rp2[i]; k < rp2[i + 1]; ++k) {
            int j = col2[k];
            p->adj2[(size_t)i * p->words + (j >> 5)] |= 1u << (j & 31);
        }
    p->S = S; p->G = G;
    p->first1 = first_copies(n1, rp1, col1);
    p->first2 = first_copies(n2, rp2, col2);
    return p;
}
void orc_problem_free(OrcProblem* p) {
    if (!p) return;
    free((void*)p->rp1); free((void*)p->col1); free((void*)p->rp2); free((void*)p->col2);
    free(p->adj2); free(p->first1); free(p->first2); free(p);
}
static inline int adj2_bit(const OrcProblem* p, int u, int v) {
    return (p->adj2[(size_t)u * p->words + (v >> 5)] >> (v & 31)) & 1u;
}

/* ------------------------------------------------------------ from scratch */

/* Alignment.cpp:667-690 fastICSDenominator */
long orc_induced_true(const OrcAln* a) {
    const OrcProblem* p = a->p;
    uint8_t* mapped = (uint8_t*)calloc((size_t)p->n2, 1);
    for (int i = 0; i < p->n1; ++i) mapped[a->aln[i]] = a->mask[i];
    long count = 0;
    for (int i = 0; i < p->n2; ++i)
        if (mapped[i])
            for (int k = p->rp2[i]; k < p->rp2[i + 1]; ++k) {
                int y = p->col2[k];
                if (mapped[y]) count += (i == y) ? 2 : 1;
            }
    free(mapped);
    return count / 2;
}

/* numerator / denominator of the set-based ics(), Alignment.cpp:604-661.
 * v2Unaligned = images of masked-off positions and of dummies (:609-619);
 * inducedES = net2 edges with both ends aligned (:622-627); mappedNet1ES =
 * images of net1 edges with both ends masked on (:629-635); aln is a
 * permutation, so distinct edges have distinct images and set sizes are counts of
 * the DISTINCT edges (first copies; duplicate edge lines are not in `edges`).
 * *mult (nullable): the same numerator with every copy counted -- the value the adjList-based
 * incremental counter (doSwap, :491-520) holds for this alignment. */
static void ics_parts(const OrcAln* a, long* num, long* den, long* mult) {
    const OrcProblem* p = a->p;
    uint8_t* v2un = (uint8_t*)calloc((size_t)p->n2, 1);
    for (int i = 0; i < p->n2; ++i) {
        if (!a->mask[i]) v2un[a->aln[i]] = 1;
        if (i >= p->n1) v2un[a->aln[i]] = 1;
    }
    long d = 0, c = 0, cm = 0;
    for (int u = 0; u < p->n2; ++u)
        for (int k = p->rp2[u]; k < p->rp2[u + 1]; ++k) {
            int v = p->col2[k];
            if (v >= u && p->first2[k] && !v2un[u] && !v2un[v]) ++d; /* each undirected edge once */
        }
    for (int u = 0; u < p->n1; ++u)
        for (int k = p->rp1[u]; k < p->rp1[u + 1]; ++k) {
            int v = p->col1[k];
            if (v < u) continue;
            if (!a->mask[u] || !a->mask[v]) continue; /* v1Unaligned (:610-611,631) */
            int iu = a->aln[u], iv = a->aln[v];
            /* member of inducedES: a net2 edge whose ends are both aligned */
            if (adj2_bit(p, iu, iv) && !v2un[iu] && !v2un[iv]) {
                ++cm;
                if (p->first1[k]) ++c;
            }
        }
    free(v2un);
    *num = c; *den = d;
    if (mult) *mult = cm;
}
long orc_conserved_true(const OrcAln* a) { long n, d, m; ics_parts(a, &n, &d, &m); return m; }
double orc_ics(const OrcAln* a) {
    long n, d; ics_parts(a, &n, &d, NULL);
    if ((double)d == 0.0) return 0.0; /* :639-641 */
    return (double)n / (double)d;     /* :658-659 */
}
/* Alignment.cpp:444-446: currConservedCount = ics()*denom  (double -> int truncation, H1) */
int orc_refseed_conserved(long C, long I) {
    double denom = (double)I;
    double ics = (denom == 0.0) ? 0.0 : (double)C / denom;
    volatile double prod = ics * denom;
    return (int)prod;
}
/* the same with the two factors as the reference has them: ics() is the ratio of the two edge
 * SETS, denom = fastICSDenominator() (they only differ with duplicate edge lines) */
int orc_refseed_from_ics(double ics, long denom) {
    volatile double prod = ics * (double)denom;
    return (int)prod;
}
/* Alignment.cpp:701-710 */
double orc_sum_blast(const OrcAln* a) {
    const OrcProblem* p = a->p;
    double s = 0.0;
    if (!p->S) return 0.0;
    for (int i = 0; i < p->n1; ++i)
        if (a->mask[i]) s += p->S[(size_t)i * p->n2 + a->aln[i]];
    return s;
}
/* Alignment.cpp:712-722 */
double orc_sum_goc(const OrcAln* a) {
    const OrcProblem* p = a->p;
    double s = 0.0;
    if (!p->G) return 0.0;
    for (int i = 0; i < p->n1; ++i)
        if (a->mask[i]) s += p->G[(size_t)i * p->n2 + a->aln[i]];
    return s;
}
/* Alignment.cpp:724-732 */
double orc_aln_size(const OrcAln* a) {
    double s = 0.0;
    for (int i = 0; i < a->p->n1; ++i)
        if (a->mask[i]) s += 1.0;
    return s;
}

OrcAln* orc_aln_create(const OrcProblem* p, const int* aln, const uint8_t* mask) {
    OrcAln* a = (OrcAln*)calloc(1, sizeof(OrcAln));
    a->p = p;
    a->aln = dup_ints(aln, p->n2);
    a->alnInv = (int*)malloc(sizeof(int) * (size_t)p->n2);
    a->mask = (uint8_t*)malloc((size_t)p->n2);
    for (int i = 0; i < p->n2; ++i) {
        a->alnInv[aln[i]] = i;
        a->mask[i] = mask ? (mask[i] != 0) : 1;
    }
    /* Alignment.cpp:437-446 */
    a->bits = orc_sum_blast(a);
    a->goc = orc_sum_goc(a);
    long denom = orc_induced_true(a);
    a->cons = orc_refseed_from_ics(orc_ics(a), denom);
    a->ind = (int)denom;
    return a;
}
OrcAln* orc_aln_clone(const OrcAln* s) {
    OrcAln* a = (OrcAln*)malloc(sizeof(OrcAln));
    *a = *s;
    a->aln = dup_ints(s->aln, s->p->n2);
    a->alnInv = dup_ints(s->alnInv, s->p->n2);
    a->mask = (uint8_t*)malloc((size_t)s->p->n2);
    memcpy(a->mask, s->mask, (size_t)s->p->n2);
    return a;
}
void orc_aln_free(OrcAln* a) {
    if (!a) return;
    free(a->aln); free(a->alnInv); free(a->mask); free(a);
}
void orc_aln_get_state(const OrcAln* a, int* aln, int* alnInv, uint8_t* mask) {
    size_t n = (size_t)a->p->n2;
    if (aln) memcpy(aln, a->aln, n * sizeof(int));
    if (alnInv) memcpy(alnInv, a->alnInv, n * sizeof(int));
    if (mask) memcpy(mask, a->mask, n);
}
void orc_aln_get_counters(const OrcAln* a, int* cons, int* ind, double* bits, double* goc) {
    *cons = a->cons; *ind = a->ind; *bits = a->bits; *goc = a->goc;
}
void orc_aln_set_counters(OrcAln* a, int cons, int ind, double bits, double goc) {
    a->cons = cons; a->ind = ind; a->bits = bits; a->goc = goc;
}

/* ------------------------------------------------------------- incremental */

/* Alignment.cpp:791-808 */
int orc_conserved_count(const OrcAln* a, int n1, int n2, int mask, int ignore) {
    const OrcProblem* p = a->p;
    if (!mask || n1 >= p->n1 || n1 == ignore) return 0;
    int r = 0;
    for (int k = p->rp1[n1]; k < p->rp1[n1 + 1]; ++k) {
        int i = p->col1[k];
        if (i != ignore && a->mask[i] && adj2_bit(p, n2, a->aln[i])) ++r;
    }
    return r;
}
/* Alignment.cpp:812-826 */
int orc_induced_count(const OrcAln* a, int n2, int ignore) {
    const OrcProblem* p = a->p;
    if (a->alnInv[n2] >= p->n1 || !a->mask[a->alnInv[n2]]) return 0;
    int r = 0;
    for (int k = p->rp2[n2]; k < p->rp2[n2 + 1]; ++k) {
        int x = p->col2[k];
        if (a->alnInv[x] < p->n1 && a->mask[a->alnInv[x]] && x != ignore) ++r;
    }
    return r;
}
/* Alignment.cpp:745-766 / :768-785 (same body on either matrix) */
static void update_score(const OrcProblem* p, const double* D, double* acc, int n1, int n2old,
                         int n2new, int oldMask, int newMask) {
    if (!D) return;
    if (n1 >= p->n1) return;
    double oldScore = 0.0, newScore = 0.0;
    if (oldMask) oldScore = D[(size_t)n1 * p->n2 + n2old];
    if (newMask) newScore = D[(size_t)n1 * p->n2 + n2new];
    double delta = newScore - oldScore;
    *acc += delta;
}
/* Alignment.cpp:491-520 */
void orc_do_swap(OrcAln* a, int x, int y) {
    const OrcProblem* p = a->p;
    int oldConservedX = orc_conserved_count(a, x, a->aln[x], a->mask[x], -1);
    int oldConservedY = orc_conserved_count(a, y, a->aln[y], a->mask[y], x);
    int oldInducedX = orc_induced_count(a, a->aln[x], -1);
    int oldInducedY = orc_induced_count(a, a->aln[y], a->aln[x]);
    int temp = a->aln[x];
    a->aln[x] = a->aln[y];
    a->alnInv[a->aln[x]] = x;
    a->aln[y] = temp;
    a->alnInv[a->aln[y]] = y;
    uint8_t tb = a->mask[x];
    a->mask[x] = a->mask[y];
    a->mask[y] = tb;
    int newConservedX = orc_conserved_count(a, x, a->aln[x], a->mask[x], -1);
    int newConservedY = orc_conserved_count(a, y, a->aln[y], a->mask[y], x);
    int newInducedX = orc_induced_count(a, a->aln[y], -1);
    int newInducedY = orc_induced_count(a, a->aln[x], a->aln[y]);
    update_score(p, p->S, &a->bits, x, a->aln[y], a->aln[x], a->mask[y], a->mask[x]);
    update_score(p, p->G, &a->goc, x, a->aln[y], a->aln[x], a->mask[y], a->mask[x]);
    update_score(p, p->S, &a->bits, y, a->aln[x], a->aln[y], a->mask[x], a->mask[y]);
    update_score(p, p->G, &a->goc, y, a->aln[x], a->aln[y], a->mask[x], a->mask[y]);
    a->cons += ((newConservedX - oldConservedX) + (newConservedY - oldConservedY));
    a->ind += ((newInducedX - oldInducedX) + (newInducedY - oldInducedY));
}

/* Ordered edit lists of the variation operators (ops [nops][4] = kind, a, b, c; kind < 0 skipped):
 *   0 doSwap(a, b)                              Alignment.cpp:491-520
 *   1 mutate's mask flip at a                   :468-480
 *   2 onBit(a)                                  :522-533
 *   3 one UPMX step: position a takes image b (= par2->aln[a]); c bit0 = par2->alnMask[a],
 *     bit1 = the coin flip, bit2 = --total      :197-252
 * Same statement order as the reference, so the fp64 accumulators follow the same sequence. */
void orc_apply_edits(OrcAln* a, const int* ops, int nops) {
    const OrcProblem* p = a->p;
    for (int o = 0; o < nops; ++o) {
        const int kind = ops[4 * o], u = ops[4 * o + 1], v = ops[4 * o + 2], w = ops[4 * o + 3];
        if (kind < 0) continue;
        if (kind == 0) {
            orc_do_swap(a, u, v);
        } else if (kind == 1 || kind == 2) {
            const int i = u;
            int oldConserved = orc_conserved_count(a, i, a->aln[i], a->mask[i], -1);
            int oldInduced = orc_induced_count(a, a->aln[i], -1);
            const uint8_t old = a->mask[i];
            a->mask[i] = kind == 2 ? 1 : !old;
            int newConserved = orc_conserved_count(a, i, a->aln[i], a->mask[i], -1);
            int newInduced = orc_induced_count(a, a->aln[i], -1);
            update_score(p, p->S, &a->bits, i, a->aln[i], a->aln[i], old, a->mask[i]);
            update_score(p, p->G, &a->goc, i, a->aln[i], a->aln[i], old, a->mask[i]);
            a->cons += (newConserved - oldConserved);
            a->ind += (newInduced - oldInduced);
        } else {
            const int i = u, temp1 = a->aln[i], temp2 = v, j = a->alnInv[temp2];
            int oldConservedI = orc_conserved_count(a, i, a->aln[i], a->mask[i], -1);
            int oldConservedJ = orc_conserved_count(a, j, a->aln[j], a->mask[j], i);
            int oldInducedI = orc_induced_count(a, a->aln[i], -1);
            int oldInducedJ = orc_induced_count(a, a->aln[j], a->aln[i]);
            a->aln[i] = temp2;
            a->alnInv[temp2] = i;
            a->aln[j] = temp1;
            a->alnInv[temp1] = j;
            const uint8_t temp1Mask = a->mask[i], par1bool = a->mask[j], par2bool = (uint8_t)(w & 1);
            if (!(w & 4)) {
                a->mask[j] = temp1Mask;
                a->mask[i] = (w & 2) ? (par1bool && par2bool) : (par1bool || par2bool);
            }
            int newConservedI = orc_conserved_count(a, i, a->aln[i], a->mask[i], -1);
            int newConservedJ = orc_conserved_count(a, j, a->aln[j], a->mask[j], i);
            int newInducedI = orc_induced_count(a, a->aln[j], -1);
            int newInducedJ = orc_induced_count(a, a->aln[i], a->aln[j]);
            update_score(p, p->S, &a->bits, i, temp1, temp2, temp1Mask, a->mask[i]);
            update_score(p, p->G, &a->goc, i, temp1, temp2, temp1Mask, a->mask[i]);
            update_score(p, p->S, &a->bits, j, temp2, temp1, par1bool, a->mask[j]);
            update_score(p, p->G, &a->goc, j, temp2, temp1, par1bool, a->mask[j]);
            a->cons += ((newConservedI - oldConservedI) + (newConservedJ - oldConservedJ));
            a->ind += ((newInducedI - oldInducedI) + (newInducedJ - oldInducedJ));
        }
    }
}

/* Alignment.cpp:693-699 */
static double fast_ics(const OrcAln* a) {
    double denom = (double)a->ind;
    if (denom == 0.0) return 0.0;
    return ((double)a->cons) / denom;
}
/* Alignment.cpp:535-537 */
static double fast_ec(const OrcAln* a) { return (double)a->cons / (double)(size_t)a->p->E1; }

/* Alignment.cpp:559-602 (+ :539-556) */
void orc_compute_fitness(OrcAln* a, const int* names, int M, double* out) {
    a->M = M;
    for (int i = 0; i < M; ++i) {
        double f = 0.0;
        switch (names[i]) {
            case ORC_ICS: f = fast_ics(a); break;
            case ORC_EC: f = fast_ec(a); break;
            case ORC_BITSCORE_SUM: f = a->bits; break;
            case ORC_EVALS_SUM: f = -1.0 * a->bits; break;
            case ORC_SIZE: f = orc_aln_size(a); break;
            case ORC_GOC: f = a->goc; break;
            case ORC_S3: { /* :539-544: size_t + int, then double */
                double num = (double)a->cons;
                double denom = (double)((size_t)a->p->E1 + (size_t)(long)a->ind) - num;
                f = num / denom;
            } break;
            case ORC_S3_DENOM: /* :589-592 */
                f = -((double)((size_t)a->p->E1 + (size_t)(long)a->ind) - a->cons);
                break;
            case ORC_ICS_TIMES_EC: f = fast_ics(a) * fast_ec(a); break; /* :546-548 */
            case ORC_S3_VARIANT: { /* :550-556 */
                double x = (double)a->cons, b = (double)(size_t)a->p->E1, c = (double)a->ind;
                double denom = (b < c ? c : b) - x;
                f = x / denom;
            } break;
            default: f = 0.0;
        }
        a->fitness[i] = f;
        if (out) out[i] = f;
    }
}

/* -------------------------------------------------------------- localSearch */

/* localSearch.cpp:29-53 over an explicit proposal list (the proposals themselves
 * come from libstdc++'s mt19937 + uniform_int_distribution, :21-28, and are
 * generated by the compiled reference shim, never re-implemented). */
void orc_hill_climb_replay(OrcAln* a, const int* xy, long iters, const int* names, int M, int obj,
                           uint8_t* accepted) {
    double currFit[ORC_MAX_OBJ];
    for (long i = 0; i < iters; ++i) {
        int x = xy[2 * i], y = xy[2 * i + 1];
        for (int j = 0; j < M; ++j) currFit[j] = a->fitness[j];
        orc_do_swap(a, x, y);
        orc_compute_fitness(a, names, M, 0);
        int improved = 1;
        if (obj == -1) {
            for (int j = 0; j < M; ++j)
                if (a->fitness[j] < currFit[j]) improved = 0;
        } else if (a->fitness[obj] <= currFit[obj]) {
            improved = 0;
        }
        if (!improved) {
            orc_do_swap(a, x, y);
            orc_compute_fitness(a, names, M, 0);
        }
        if (accepted) accepted[i] = (uint8_t)improved;
    }
}

/* The proposal stream of localSearch.cpp:21-28 restated in plain C:
 *     x = uniform_int_distribution<int>(0, n1 - 1)(prng);  y = uniform_int_distribution<int>(0, n2 - 2)(prng);
 *     if (y >= x) y++;
 * prng = std::mt19937 (Network.h:10): MT19937 of Matsumoto & Nishimura (n = 624, m = 397, r = 31,
 * a = 0x9908b0df, tempering (11, 0xffffffff), (7, 0x9d2c5680), (15, 0xefc60000), 18), state regenerated
 * when all 624 words have been used.  The distributions are GNU libstdc++'s (GCC >= 11,
 * bits/uniform_int_dist.h:_S_nd): for a 32-bit generator and a range of R values, Lemire's
 * multiply-shift  v = (u * R) >> 32  with the draw repeated while  (uint32)(u * R) < 2^32 mod R.
 * state = the 625 numbers operator<< of std::mt19937 prints (624 words, then the position 0..624);
 * it is advanced in place.  Pinned against the compiled reference shim (libstdc++ itself) in
 * tests/test_oracle_vs_reference.py. */
static uint32_t orc_mt_next(uint32_t* st) {
    uint32_t* x = st;
    if (st[624] >= 624u) {
        for (int k = 0; k < 624; ++k) {
            uint32_t y = (x[k] & 0x80000000u) | (x[(k + 1) % 624] & 0x7fffffffu);
            x[k] = x[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        st[624] = 0;
    }
    uint32_t z = x[st[624]++];
    z ^= z >> 11;
    z ^= (z << 7) & 0x9d2c5680u;
    z ^= (z << 15) & 0xefc60000u;
    z ^= z >> 18;
    return z;
}
static int orc_mt_uniform(uint32_t* st, uint32_t range) {
    uint64_t prod = (uint64_t)orc_mt_next(st) * (uint64_t)range;
    uint32_t low = (uint32_t)prod;
    if (low < range) {
        uint32_t threshold = (0u - range) % range;
        while (low < threshold) {
            prod = (uint64_t)orc_mt_next(st) * (uint64_t)range;
            low = (uint32_t)prod;
        }
    }
    return (int)(prod >> 32);
}
void orc_mt_seed(uint32_t seed, uint32_t* state) {   /* mersenne_twister_engine::seed(value) */
    state[0] = seed;
    for (uint32_t i = 1; i < 624; ++i) state[i] = 1812433253u * (state[i - 1] ^ (state[i - 1] >> 30)) + i;
    state[624] = 624;
}
void orc_mt_xy_stream(uint32_t* state, int n1, int n2, long iters, int* xy) {
    for (long i = 0; i < iters; ++i) {
        int x = orc_mt_uniform(state, (uint32_t)n1);
        int y = orc_mt_uniform(state, (uint32_t)(n2 - 1));
        if (y >= x) y++;
        xy[2 * i] = x;
        xy[2 * i + 1] = y;
    }
}

/* localSearch.cpp:88-120 */
double orc_swap_normalized_delta(OrcAln* a, const int* names, int M, int x, int y) {
    double currFit[ORC_MAX_OBJ], newFit[ORC_MAX_OBJ], pct[ORC_MAX_OBJ];
    for (int j = 0; j < M; ++j) currFit[j] = a->fitness[j];
    orc_do_swap(a, x, y);
    orc_compute_fitness(a, names, M, newFit);
    orc_do_swap(a, x, y);
    orc_compute_fitness(a, names, M, 0);
    int oneNeg = 0;
    for (int i = 0; i < M; ++i) {
        pct[i] = (newFit[i] - currFit[i]) / currFit[i];
        if (pct[i] < 0) oneNeg = 1;
    }
    if (oneNeg) return -1.0;
    double sumSq = 0.0;
    for (int i = 0; i < M; ++i) sumSq += pct[i] * pct[i];
    return sqrt(sumSq);
}

/* same contract as ref_swap_delta_records in ref_shim.cpp */
void orc_swap_delta_records(OrcAln* a, const int* names, int M, int obj, const int* xy, long n,
                            int* cons, int* ind, double* bits, double* goc, double* fit,
                            uint8_t* improved, double* snd) {
    double base[ORC_MAX_OBJ];
    orc_compute_fitness(a, names, M, base);
    const double b0 = a->bits, g0 = a->goc;
    for (long i = 0; i < n; ++i) {
        int x = xy[2 * i], y = xy[2 * i + 1];
        if (snd) {
            snd[i] = orc_swap_normalized_delta(a, names, M, x, y);
            a->bits = b0; a->goc = g0;
            memcpy(a->fitness, base, sizeof(double) * (size_t)M);
        }
        orc_do_swap(a, x, y);
        orc_compute_fitness(a, names, M, 0);
        if (cons) cons[i] = a->cons;
        if (ind) ind[i] = a->ind;
        if (bits) bits[i] = a->bits;
        if (goc) goc[i] = a->goc;
        if (fit) memcpy(fit + i * M, a->fitness, sizeof(double) * (size_t)M);
        if (improved) {
            int imp = 1;
            if (obj == -1) {
                for (int j = 0; j < M; ++j)
                    if (a->fitness[j] < base[j]) imp = 0;
            } else if (a->fitness[obj] <= base[obj]) {
                imp = 0;
            }
            improved[i] = (uint8_t)imp;
        }
        orc_do_swap(a, x, y);
        a->bits = b0; a->goc = g0;
        memcpy(a->fitness, base, sizeof(double) * (size_t)M);
    }
}

/* ------------------------------------------------------------ CPU baselines */
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    OrcAln* a; const int* names; int M; const int* xy; long lo, hi; double sum;
    int obj; long iters; OrcAln** alns; int nalns; int* next; pthread_mutex_t* mu;
    long reps_lo, reps_hi;
} Job;

static void* job_swap_deltas(void* v) {
    Job* j = (Job*)v;
    double s = 0.0;
    for (long i = j->lo; i < j->hi; ++i) {
        double d = orc_swap_normalized_delta(j->a, j->names, j->M, j->xy[2 * i], j->xy[2 * i + 1]);
        if (d > 0.0 && d < 1e300) s += d;
    }
    j->sum = s;
    return 0;
}
/* worker pattern of localSearch.cpp:130-147: private copy per thread, a slab of candidates each */
double orc_time_swap_deltas(OrcAln** alns, int nalns, const int* names, int M, const int* xy,
                            long n, int nthreads, double* sink) {
    Job* jobs = (Job*)calloc((size_t)nthreads, sizeof(Job));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; ++t) {
        jobs[t].a = orc_aln_clone(alns[t % nalns]);
        orc_compute_fitness(jobs[t].a, names, M, 0);
        jobs[t].names = names; jobs[t].M = M; jobs[t].xy = xy;
        jobs[t].lo = n * t / nthreads; jobs[t].hi = n * (t + 1) / nthreads;
    }
    double t0 = now_s();
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], 0, job_swap_deltas, &jobs[t]);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
    double t1 = now_s();
    double s = 0.0;
    for (int t = 0; t < nthreads; ++t) { s += jobs[t].sum; orc_aln_free(jobs[t].a); }
    if (sink) *sink = s;
    free(jobs); free(th);
    return t1 - t0;
}

static void* job_hill_climb(void* v) {
    Job* j = (Job*)v;
    for (;;) {
        pthread_mutex_lock(j->mu);
        int i = (*j->next)++;
        pthread_mutex_unlock(j->mu);
        if (i >= j->nalns) break;
        orc_hill_climb_replay(j->alns[i], j->xy + 2 * j->iters * (long)i, j->iters, j->names, j->M,
                              j->obj, 0);
    }
    return 0;
}
/* worker pattern of steadystate.cpp:132-253; xy holds nalns*iters proposals */
double orc_time_hill_climb(OrcAln** alns, int nalns, const int* xy, long iters, const int* names,
                           int M, int obj, int nthreads) {
    for (int i = 0; i < nalns; ++i) orc_compute_fitness(alns[i], names, M, 0);
    Job* jobs = (Job*)calloc((size_t)nthreads, sizeof(Job));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    int next = 0;
    for (int t = 0; t < nthreads; ++t) {
        jobs[t].alns = alns; jobs[t].nalns = nalns; jobs[t].xy = xy; jobs[t].iters = iters;
        jobs[t].names = names; jobs[t].M = M; jobs[t].obj = obj; jobs[t].next = &next; jobs[t].mu = &mu;
    }
    double t0 = now_s();
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], 0, job_hill_climb, &jobs[t]);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
    double t1 = now_s();
    free(jobs); free(th);
    return t1 - t0;
}

static void* job_full_fitness(void* v) {
    Job* j = (Job*)v;
    double s = 0.0;
    for (long k = j->reps_lo; k < j->reps_hi; ++k) {
        OrcAln* a = j->alns[k % j->nalns];
        a->bits = orc_sum_blast(a);
        a->goc = orc_sum_goc(a);
        long denom = orc_induced_true(a);
        a->cons = orc_refseed_from_ics(orc_ics(a), denom);
        a->ind = (int)denom;
        orc_compute_fitness(a, j->names, j->M, 0);
        s += a->fitness[0];
    }
    j->sum = s;
    return 0;
}
/* tail of shuf (Alignment.cpp:437-446) + computeFitness, array based (no hash sets) */
double orc_time_full_fitness(OrcAln** alns, int nalns, const int* names, int M, long reps,
                             int nthreads, double* sink) {
    Job* jobs = (Job*)calloc((size_t)nthreads, sizeof(Job));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    OrcAln** priv = (OrcAln**)calloc((size_t)nthreads * (size_t)nalns, sizeof(OrcAln*));
    for (int t = 0; t < nthreads; ++t) {
        for (int i = 0; i < nalns; ++i) priv[(size_t)t * nalns + i] = orc_aln_clone(alns[i]);
        jobs[t].alns = priv + (size_t)t * nalns; jobs[t].nalns = nalns;
        jobs[t].names = names; jobs[t].M = M;
        jobs[t].reps_lo = reps * t / nthreads; jobs[t].reps_hi = reps * (t + 1) / nthreads;
    }
    double t0 = now_s();
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], 0, job_full_fitness, &jobs[t]);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
    double t1 = now_s();
    double s = 0.0;
    for (int t = 0; t < nthreads; ++t) s += jobs[t].sum;
    for (size_t i = 0; i < (size_t)nthreads * (size_t)nalns; ++i) orc_aln_free(priv[i]);
    if (sink) *sink = s;
    free(jobs); free(th); free(priv);
    return t1 - t0;
}
