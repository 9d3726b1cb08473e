// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or shipped with the
// product (optnetaligncpp_b200/).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library this
// file is compiled into.
//
// extern "C" shim around the UNMODIFIED reference translation units
//   /root/reference/{Network,blastinfo,goc,Alignment,localSearch,nsga-ii,Archive}.cpp
// which oracle/Makefile compiles *where they lie* together with this file into
// oracle/_ref/libona_ref.so.  Everything in here only CALLS the reference's own
// public C++ API (Alignment.h:13-113, localSearch.h:9-24, Network.h:98-108,
// blastinfo.h:12, goc.h:14) and moves data across a C boundary so that the
// Python tests and bench.py can drive the real reference through ctypes.
//
// The only pieces of logic that are restated here rather than called are
//   * ref_gen_xy_stream      -- the two uniform_int_distribution draws of
//                               localSearch.cpp:21-28, so that the very same
//                               (x,y) proposal stream can be fed to the GPU;
//   * ref_hill_climb_replay  -- the accept/undo loop of localSearch.cpp:29-53
//                               driven by an explicit (x,y) list instead of a
//                               PRNG (tests pin it against the real
//                               correctHillClimb with the same seed);
//   * ref_aln_set_state      -- the "seed the incremental counters" idiom of
//                               Alignment.cpp:437-446 after assigning a
//                               caller-provided permutation/mask.
// All three operate through the reference's doSwap/computeFitness/ics/... so the
// arithmetic is still the reference's.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <exception>
#include <random>
#include <string>
#include <thread>
#include <sstream>
#include <vector>

#include "Alignment.h"
#include "Network.h"
#include "blastinfo.h"
#include "goc.h"
#include "localSearch.h"
#include "nsga-ii.h"

static thread_local std::string g_err;

#define REF_TRY try {
#define REF_CATCH(ret)                   \
    }                                    \
    catch (const std::exception& e) {    \
        g_err = e.what();                \
        return ret;                      \
    }                                    \
    catch (...) {                        \
        g_err = "unknown C++ exception"; \
        return ret;                      \
    }

static std::vector<fitnessName> to_names(const int* names, int M) {
    std::vector<fitnessName> v;
    for (int i = 0; i < M; ++i) v.push_back(static_cast<fitnessName>(names[i]));
    return v;
}

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// ----------------------------------------------------------------- Network
void* ref_net_load(const char* path) {
    REF_TRY
    return new Network(std::string(path));
    REF_CATCH(nullptr)
}
void ref_net_free(void* n) { delete static_cast<Network*>(n); }
int ref_net_n(void* n) { return (int)static_cast<Network*>(n)->nodeToNodeName.size(); }
long ref_net_num_edges(void* n) { return (long)static_cast<Network*>(n)->edges.size(); }
long ref_net_adj_total(void* n) {
    long t = 0;
    for (auto& l : static_cast<Network*>(n)->adjList) t += (long)l.size();
    return t;
}
// rowptr has n+1 entries, col has ref_net_adj_total entries; adjList order kept.
void ref_net_export_csr(void* n, int* rowptr, int* col) {
    Network* net = static_cast<Network*>(n);
    long k = 0;
    for (size_t i = 0; i < net->adjList.size(); ++i) {
        rowptr[i] = (int)k;
        for (node v : net->adjList[i]) col[k++] = v;
    }
    rowptr[net->adjList.size()] = (int)k;
}
int ref_net_adjmatrix_bit(void* n, int u, int v) {
    return static_cast<Network*>(n)->adjMatrix[u][v] ? 1 : 0;
}
// writes the node name into buf (NUL terminated, truncated to cap-1)
void ref_net_node_name(void* n, int u, char* buf, int cap) {
    const std::string& s = static_cast<Network*>(n)->nodeToNodeName.at(u);
    std::strncpy(buf, s.c_str(), cap - 1);
    buf[cap - 1] = 0;
}

// ------------------------------------------------------- score dictionaries
void* ref_blast_load(void* n1, void* n2, const char* path) {
    REF_TRY
    return loadBLASTInfo(static_cast<Network*>(n1), static_cast<Network*>(n2), std::string(path));
    REF_CATCH(nullptr)
}
void* ref_goc_load(void* n1, void* n2, const char* f1, const char* f2) {
    REF_TRY
    return loadGOC(static_cast<Network*>(n1), static_cast<Network*>(n2), std::string(f1),
                   std::string(f2));
    REF_CATCH(nullptr)
}
// dense row-major rows x cols doubles -> a BLASTDict/GOCDict (same type)
void* ref_dict_from_dense(int rows, int cols, const double* data) {
    BLASTDict* d = new BLASTDict(boost::extents[rows][cols]);
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) (*d)[i][j] = data[(size_t)i * cols + j];
    return d;
}
void ref_dict_free(void* d) { delete static_cast<BLASTDict*>(d); }
double ref_dict_get(void* d, int i, int j) { return (*static_cast<BLASTDict*>(d))[i][j]; }
// copies row i (cols doubles) out
void ref_dict_row(void* d, int i, int cols, double* out) {
    const double* r = (*static_cast<const BLASTDict*>(d))[i];
    std::memcpy(out, r, sizeof(double) * (size_t)cols);
}

// --------------------------------------------------------------- Alignment
void* ref_aln_new(void* n1, void* n2, void* bit, void* goc) {
    REF_TRY
    Alignment* a = new Alignment(static_cast<Network*>(n1), static_cast<Network*>(n2),
                                 static_cast<BLASTDict*>(bit), static_cast<GOCDict*>(goc));
    // constructor 1 leaves the two fp64 accumulators uninitialised
    // (Alignment.cpp:23-47); give them the from-scratch values.
    if (a->bitscores) a->currBitscore = a->sumBLAST(); else a->currBitscore = 0.0;
    if (a->gocs) a->currGOC = a->sumGOC(); else a->currGOC = 0.0;
    return a;
    REF_CATCH(nullptr)
}
// the alignment-file constructor (Alignment.cpp:49-138)
void* ref_aln_load(void* n1, void* n2, const char* path, void* bit, void* goc) {
    REF_TRY
    return new Alignment(static_cast<Network*>(n1), static_cast<Network*>(n2), std::string(path),
                         static_cast<BLASTDict*>(bit), static_cast<GOCDict*>(goc));
    REF_CATCH(nullptr)
}
void* ref_aln_clone(void* a) { return new Alignment(*static_cast<Alignment*>(a)); }
void ref_aln_free(void* a) { delete static_cast<Alignment*>(a); }

void ref_aln_shuf(void* a, uint32_t seed, int uniformsize, int smallstart, int total) {
    RandGenT g(seed);
    static_cast<Alignment*>(a)->shuf(g, uniformsize != 0, smallstart != 0, total != 0);
}

// assign permutation + mask, then seed counters exactly like Alignment.cpp:437-446
void ref_aln_set_state(void* av, const int* aln, const uint8_t* mask) {
    Alignment* a = static_cast<Alignment*>(av);
    size_t n = a->aln.size();
    for (size_t i = 0; i < n; ++i) {
        a->aln[i] = aln[i];
        a->alnInv[aln[i]] = (int)i;
        a->alnMask[i] = mask ? (mask[i] != 0) : true;
    }
    if (a->bitscores) a->currBitscore = a->sumBLAST();
    if (a->gocs) a->currGOC = a->sumGOC();
    double denom = a->fastICSDenominator();
    a->currConservedCount = a->ics() * denom;
    a->currInducedCount = (int)denom;
    a->fitnessValid = false;
}
int ref_aln_size_n2(void* a) { return (int)static_cast<Alignment*>(a)->aln.size(); }
void ref_aln_get_state(void* av, int* aln, int* alnInv, uint8_t* mask) {
    Alignment* a = static_cast<Alignment*>(av);
    size_t n = a->aln.size();
    for (size_t i = 0; i < n; ++i) {
        if (aln) aln[i] = a->aln[i];
        if (alnInv) alnInv[i] = a->alnInv[i];
        if (mask) mask[i] = a->alnMask[i] ? 1 : 0;
    }
}
void ref_aln_get_counters(void* av, int* cons, int* ind, double* bits, double* goc) {
    Alignment* a = static_cast<Alignment*>(av);
    *cons = a->currConservedCount;
    *ind = a->currInducedCount;
    *bits = a->bitscores ? a->currBitscore : 0.0;
    *goc = a->gocs ? a->currGOC : 0.0;
}
void ref_aln_set_counters(void* av, int cons, int ind, double bits, double goc) {
    Alignment* a = static_cast<Alignment*>(av);
    a->currConservedCount = cons;
    a->currInducedCount = ind;
    a->currBitscore = bits;
    a->currGOC = goc;
}
void ref_aln_do_swap(void* a, int x, int y) { static_cast<Alignment*>(a)->doSwap(x, y); }
void ref_aln_compute_fitness(void* av, const int* names, int M, double* out) {
    Alignment* a = static_cast<Alignment*>(av);
    std::vector<fitnessName> fn = to_names(names, M);
    // computeFitness sizes the vector on first use only (Alignment.cpp:560-562): keep a caller
    // that switches objective lists from writing past it
    if ((int)a->fitness.size() != M) a->fitness.assign(M, 0.0);
    a->computeFitness(fn);
    for (int i = 0; i < M; ++i) out[i] = a->fitness[i];
}
double ref_aln_ics(void* a) { return static_cast<Alignment*>(a)->ics(); }
double ref_aln_fast_ics_denominator(void* a) {
    return static_cast<Alignment*>(a)->fastICSDenominator();
}
double ref_aln_sum_blast(void* a) { return static_cast<Alignment*>(a)->sumBLAST(); }
double ref_aln_sum_goc(void* a) { return static_cast<Alignment*>(a)->sumGOC(); }
double ref_aln_aln_size(void* a) { return static_cast<Alignment*>(a)->alnSize(); }
int ref_aln_conserved_count(void* a, int n1, int n2, int mask, int ignore) {
    return static_cast<Alignment*>(a)->conservedCount(n1, n2, mask != 0, ignore);
}
int ref_aln_induced_count(void* a, int n2, int ignore) {
    return static_cast<Alignment*>(a)->inducedCount(n2, ignore);
}
void ref_aln_mutate(void* a, uint32_t seed, float mutswappb, int total) {
    RandGenT g(seed);
    static_cast<Alignment*>(a)->mutate(g, mutswappb, total != 0);
}
void ref_aln_greedy_match(void* a, int bit) { static_cast<Alignment*>(a)->greedyMatch(bit != 0); }
void ref_aln_on_bit(void* a, int x) { static_cast<Alignment*>(a)->onBit(x); }
void* ref_aln_crossover(uint32_t seed, float cxswappb, void* p1, void* p2, int total) {
    RandGenT g(seed);
    return new Alignment(g, cxswappb, *static_cast<Alignment*>(p1), *static_cast<Alignment*>(p2),
                         total != 0);
}

// The PRNG draws of Alignment::mutate (Alignment.cpp:451-466) and of the UPMX constructor
// (:159-166, :197-198, :223) restated, like ref_gen_xy_stream, so that the very same edit list can
// be fed to the GPU: they depend on the generator only, never on the alignment.  Tests pin them
// by comparing the replayed list against the real mutate / constructor.
//   ops: [cap][4] = {kind, a, b, c} as include/ona_b200.h (ONA_EDIT_*); returns the count.
int ref_gen_mutate_ops(uint32_t seed, int size, float mutswappb, int total, int* ops, int cap) {
    RandGenT prng(seed);
    std::uniform_real_distribution<float> fltgen(0.0, 1.0);
    std::uniform_int_distribution<int> intgen(0, size - 2);
    int n = 0;
    for (int i = 0; i < size; i++) {
        if (fltgen(prng) < mutswappb) {
            int swapIndex = intgen(prng);
            if (swapIndex >= i) swapIndex++;
            if (n < cap) { ops[4 * n] = 0; ops[4 * n + 1] = i; ops[4 * n + 2] = swapIndex; ops[4 * n + 3] = 0; }
            n++;
        }
        if (!total && fltgen(prng) < mutswappb) {
            if (n < cap) { ops[4 * n] = 1; ops[4 * n + 1] = i; ops[4 * n + 2] = 0; ops[4 * n + 3] = 0; }
            n++;
        }
    }
    return n;
}
// first_is_p1: whether the constructor's first coin made p1 the primary parent;  sel: [cap][2] =
// {i, coin (0 under --total)} for every index the crossover touches.
int ref_gen_cross_draws(uint32_t seed, int size, float cxswappb, int total, int* first_is_p1, int* sel,
                        int cap) {
    RandGenT prng(seed);
    std::uniform_int_distribution<int> coinFlip(0, 1);
    std::uniform_real_distribution<float> fltgen(0.0, 1.0);
    *first_is_p1 = coinFlip(prng) ? 1 : 0;
    int n = 0;
    for (int i = 0; i < size; i++) {
        if (fltgen(prng) < cxswappb) {
            const int coin = total ? 0 : coinFlip(prng);
            if (n < cap) { sel[2 * n] = i; sel[2 * n + 1] = coin; }
            n++;
        }
    }
    return n;
}

// ------------------------------------------------------------ local search
// the real thing: correctHillClimb with mt19937(seed)  (localSearch.cpp:18-55)
void ref_correct_hill_climb(void* av, uint32_t seed, int iters, const int* names, int M, int obj) {
    Alignment* a = static_cast<Alignment*>(av);
    std::vector<fitnessName> fn = to_names(names, M);
    RandGenT g(seed);
    correctHillClimb(g, a, true, iters, fn, obj);
}

// proportionalSearch (localSearch.cpp:60-84) with mt19937(seed)
void ref_proportional_search(void* av, uint32_t seed, int iters, const int* names, int M, int obj,
                             double proportion) {
    Alignment* a = static_cast<Alignment*>(av);
    std::vector<fitnessName> fn = to_names(names, M);
    if ((int)a->fitness.size() != M) a->fitness.assign(M, 0.0);
    a->computeFitness(fn);
    RandGenT g(seed);
    proportionalSearch(g, a, true, iters, fn, obj, proportion);
}

// steepestAscentHillClimb (localSearch.cpp:122-203) until no swap improves
void ref_steepest_ascent(void* av, const int* names, int M, int nthreads) {
    Alignment* a = static_cast<Alignment*>(av);
    std::vector<fitnessName> fn = to_names(names, M);
    if ((int)a->fitness.size() != M) a->fitness.assign(M, 0.0);
    a->computeFitness(fn);
    steepestAscentHillClimb(a, fn, nthreads, false);
}

// the (x,y) proposals correctHillClimb would draw from mt19937(seed):
// same two distribution objects, same draw order as localSearch.cpp:21-28.
void ref_gen_xy_stream(uint32_t seed, int n1, int n2, long iters, int* xy) {
    RandGenT prng(seed);
    auto randNonDummyIndex = std::uniform_int_distribution<int>(0, n1 - 1);
    auto randIndex = std::uniform_int_distribution<int>(0, n2 - 2);
    for (long i = 0; i < iters; ++i) {
        node x = randNonDummyIndex(prng);
        node y = randIndex(prng);
        if (y >= x) y++;
        xy[2 * i] = x;
        xy[2 * i + 1] = y;
    }
}

// same, starting from an explicit generator state and returning the state the draws leave behind:
// state = the 625 numbers operator<< / operator>> of std::mt19937 exchange (624 words, position)
void ref_gen_xy_stream_state(uint32_t* state, int n1, int n2, long iters, int* xy) {
    std::stringstream ss;
    for (int i = 0; i < 625; ++i) ss << state[i] << ' ';
    RandGenT prng;
    ss >> prng;
    auto randNonDummyIndex = std::uniform_int_distribution<int>(0, n1 - 1);
    auto randIndex = std::uniform_int_distribution<int>(0, n2 - 2);
    for (long i = 0; i < iters; ++i) {
        node x = randNonDummyIndex(prng);
        node y = randIndex(prng);
        if (y >= x) y++;
        xy[2 * i] = x;
        xy[2 * i + 1] = y;
    }
    std::stringstream out;
    out << prng;
    for (int i = 0; i < 625; ++i) {
        unsigned long v;
        out >> v;
        state[i] = (uint32_t)v;
    }
}
// the state of mt19937(seed) in the same form
void ref_mt_state(uint32_t seed, uint32_t* state) {
    RandGenT prng(seed);
    std::stringstream out;
    out << prng;
    for (int i = 0; i < 625; ++i) {
        unsigned long v;
        out >> v;
        state[i] = (uint32_t)v;
    }
}

// accept/undo loop of localSearch.cpp:29-53 over an explicit proposal list.
// accepted (optional) receives 1/0 per proposal.  aln->fitness must be valid.
void ref_hill_climb_replay(void* av, const int* xy, long iters, const int* names, int M, int obj,
                           uint8_t* accepted) {
    Alignment* aln = static_cast<Alignment*>(av);
    std::vector<fitnessName> fn = to_names(names, M);
    std::vector<double> currFit(M);
    for (long i = 0; i < iters; ++i) {
        node x = xy[2 * i], y = xy[2 * i + 1];
        std::copy(aln->fitness.begin(), aln->fitness.end(), currFit.begin());
        aln->doSwap(x, y);
        aln->computeFitness(fn);
        bool improved = true;
        if (obj == -1) {
            for (int j = 0; j < M; ++j)
                if (aln->fitness[j] < currFit[j]) improved = false;
        } else {
            if (aln->fitness[obj] <= currFit[obj]) improved = false;
        }
        if (!improved) {
            aln->doSwap(x, y);
            aln->computeFitness(fn);
        }
        if (accepted) accepted[i] = improved ? 1 : 0;
    }
}

double ref_swap_normalized_delta(void* av, const int* names, int M, int x, int y) {
    Alignment* a = static_cast<Alignment*>(av);
    std::vector<fitnessName> fn = to_names(names, M);
    return swapNormalizedDelta(*a, fn, x, y);
}

// Per-candidate records against ONE base state: for each (x,y) do the swap with
// the reference's doSwap+computeFitness, record the new counters / accumulators /
// objectives and the accept decision for `obj`, undo, and restore the two fp64
// accumulators + fitness to the base values (the undo is not bit-exact, H3).
// Outputs (any may be null): cons[n], ind[n], bits[n], goc[n], fit[n*M], improved[n],
// snd[n] = swapNormalizedDelta value.
void ref_swap_delta_records(void* av, const int* names, int M, int obj, const int* xy, long n,
                            int* cons, int* ind, double* bits, double* goc, double* fit,
                            uint8_t* improved, double* snd) {
    Alignment* a = static_cast<Alignment*>(av);
    std::vector<fitnessName> fn = to_names(names, M);
    a->computeFitness(fn);
    const std::vector<double> base = a->fitness;
    const double b0 = a->currBitscore, g0 = a->currGOC;
    for (long i = 0; i < n; ++i) {
        node x = xy[2 * i], y = xy[2 * i + 1];
        if (snd) {
            snd[i] = swapNormalizedDelta(*a, fn, x, y);
            a->currBitscore = b0;
            a->currGOC = g0;
            a->fitness = base;
        }
        a->doSwap(x, y);
        a->computeFitness(fn);
        if (cons) cons[i] = a->currConservedCount;
        if (ind) ind[i] = a->currInducedCount;
        if (bits) bits[i] = a->bitscores ? a->currBitscore : 0.0;
        if (goc) goc[i] = a->gocs ? a->currGOC : 0.0;
        if (fit)
            for (int j = 0; j < M; ++j) fit[i * M + j] = a->fitness[j];
        if (improved) {
            bool imp = true;
            if (obj == -1) {
                for (int j = 0; j < M; ++j)
                    if (a->fitness[j] < base[j]) imp = false;
            } else if (a->fitness[obj] <= base[obj]) {
                imp = false;
            }
            improved[i] = imp ? 1 : 0;
        }
        a->doSwap(x, y);
        a->currBitscore = b0;
        a->currGOC = g0;
        a->fitness = base;
    }
}

// ------------------------------------------------------------ CPU baselines
static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch())
        .count();
}

// Threaded swapNormalizedDelta over a fixed candidate list: the worker pattern of
// localSearch.cpp:130-147 (private Alignment copy per thread, contiguous slab of
// the candidates each).  alns[] holds `nalns` base individuals; thread t works
// on individual t % nalns.  Returns wall seconds; *sink gets a checksum.
double ref_time_swap_deltas(void** alns, int nalns, const int* names, int M, const int* xy,
                            long n, int nthreads, double* sink) {
    std::vector<fitnessName> fn = to_names(names, M);
    std::vector<double> sums(nthreads, 0.0);
    std::vector<Alignment> locals;
    locals.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t) {
        locals.push_back(*static_cast<Alignment*>(alns[t % nalns]));
        locals.back().computeFitness(fn);
    }
    std::vector<std::thread> ts;
    double t0 = now_s();
    for (int t = 0; t < nthreads; ++t) {
        ts.push_back(std::thread([&, t]() {
            long lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
            Alignment& local = locals[t];
            double s = 0.0;
            for (long i = lo; i < hi; ++i) {
                double d = swapNormalizedDelta(local, fn, xy[2 * i], xy[2 * i + 1]);
                if (d > 0.0 && d < 1e300) s += d;
            }
            sums[t] = s;
        }));
    }
    for (auto& th : ts) th.join();
    double t1 = now_s();
    double s = 0.0;
    for (double v : sums) s += v;
    if (sink) *sink = s;
    return t1 - t0;
}

// Threaded correctHillClimb on independent individuals (the worker pattern of
// steadystate.cpp:132-253): individual i is climbed `iters` iterations with
// mt19937(seed0+i); individuals are handed out dynamically to nthreads threads.
// The alignments ARE modified.  Returns wall seconds.
double ref_time_hill_climb(void** alns, int nalns, uint32_t seed0, int iters, const int* names,
                           int M, int obj, int nthreads) {
    std::vector<fitnessName> fn = to_names(names, M);
    for (int i = 0; i < nalns; ++i) static_cast<Alignment*>(alns[i])->computeFitness(fn);
    std::atomic<int> next(0);
    std::vector<std::thread> ts;
    double t0 = now_s();
    for (int t = 0; t < nthreads; ++t) {
        ts.push_back(std::thread([&]() {
            for (;;) {
                int i = next.fetch_add(1);
                if (i >= nalns) break;
                RandGenT g(seed0 + (uint32_t)i);
                correctHillClimb(g, static_cast<Alignment*>(alns[i]), true, iters, fn, -1 == obj ? -1 : obj);
            }
        }));
    }
    for (auto& th : ts) th.join();
    return now_s() - t0;
}

// Threaded from-scratch scoring (the tail of shuf, Alignment.cpp:437-446, plus
// computeFitness): `reps` evaluations spread over nthreads threads, each on
// individual (k % nalns).  with_ics=0 skips the hash-set ics() and recounts the
// conserved edges through conservedCount() instead (adjacency based), so the
// comparison is not inflated by the reference's unordered_set implementation.
double ref_time_full_fitness(void** alns, int nalns, const int* names, int M, long reps,
                             int nthreads, int with_ics, double* sink) {
    std::vector<fitnessName> fn = to_names(names, M);
    std::vector<std::vector<Alignment>> locals(nthreads);
    for (int t = 0; t < nthreads; ++t)
        for (int i = 0; i < nalns; ++i) locals[t].push_back(*static_cast<Alignment*>(alns[i]));
    std::vector<double> sums(nthreads, 0.0);
    std::vector<std::thread> ts;
    double t0 = now_s();
    for (int t = 0; t < nthreads; ++t) {
        ts.push_back(std::thread([&, t]() {
            long lo = reps * t / nthreads, hi = reps * (t + 1) / nthreads;
            double s = 0.0;
            for (long k = lo; k < hi; ++k) {
                Alignment& a = locals[t][k % nalns];
                if (a.bitscores) a.currBitscore = a.sumBLAST();
                if (a.gocs) a.currGOC = a.sumGOC();
                double denom = a.fastICSDenominator();
                if (with_ics) {
                    a.currConservedCount = a.ics() * denom;
                } else {
                    int twice = 0;
                    for (int u = 0; u < a.actualSize; ++u) {
                        int c = a.conservedCount(u, a.aln[u], a.alnMask[u], -1);
                        // a self-loop appears once in adjList, all other edges twice
                        twice += c;
                        if (a.alnMask[u] && a.net1->adjMatrix[u][u] &&
                            a.net2->adjMatrix[a.aln[u]][a.aln[u]])
                            twice += 1;
                    }
                    a.currConservedCount = twice / 2;
                }
                a.currInducedCount = (int)denom;
                a.computeFitness(fn);
                s += a.fitness[0];
            }
            sums[t] = s;
        }));
    }
    for (auto& th : ts) th.join();
    double t1 = now_s();
    double s = 0.0;
    for (double v : sums) s += v;
    if (sink) *sink = s;
    return t1 - t0;
}

// nsga-ii.cpp:14-118 through real Alignment objects, in the order nsgaiimain.cpp:157-175 calls them:
// nonDominatedSort(all) -> normalizeFitnesses(all) -> setCrowdingDists(front) for every front.
//   has_norm[i] != 0: individual i already carries M normalised values (norm_io row i) from an
//   earlier generation; normalizeFitnesses only APPENDS (:79), so setCrowdingDists keeps reading
//   those stale leading entries.  norm_io returns the leading M entries of every individual.
//   front_order: the fronts as nonDominatedSort returned them; order_after: the same fronts after
//   setCrowdingDists' in-place sorts.  Returns the time spent inside the three reference calls.
double ref_nsga2(int P, int M, const double* fitness, const unsigned char* has_norm, double* norm_io,
                 int* dom_rank, int* front_order, int* front_offsets, int* n_fronts, int* order_after,
                 double* crowd_dist) {
    REF_TRY
    std::vector<Alignment> objs(P);
    std::vector<Alignment*> all(P);
    for (int i = 0; i < P; ++i) {
        objs[i].fitness.assign(fitness + (size_t)i * M, fitness + (size_t)(i + 1) * M);
        objs[i].domRank = -1;
        objs[i].crowdDist = -1.0;
        if (has_norm && has_norm[i])
            objs[i].fitnessNormalized.assign(norm_io + (size_t)i * M, norm_io + (size_t)(i + 1) * M);
        all[i] = &objs[i];
    }
    double t0 = now_s();
    std::vector<std::vector<Alignment*> > fronts = nonDominatedSort(all);
    normalizeFitnesses(all);
    double t1 = now_s();
    int off = 0;
    *n_fronts = (int)fronts.size();
    for (size_t f = 0; f < fronts.size(); ++f) {
        front_offsets[f] = off;
        for (Alignment* a : fronts[f]) front_order[off++] = (int)(a - &objs[0]);
    }
    front_offsets[fronts.size()] = off;
    double t2 = now_s();
    for (auto& fr : fronts)
        if (!fr.empty()) setCrowdingDists(fr);
    double t3 = now_s();
    off = 0;
    for (auto& fr : fronts)
        for (Alignment* a : fr) order_after[off++] = (int)(a - &objs[0]);
    for (int i = 0; i < P; ++i) {
        dom_rank[i] = objs[i].domRank;
        crowd_dist[i] = objs[i].crowdDist;
        for (int m = 0; m < M; ++m) norm_io[(size_t)i * M + m] = objs[i].fitnessNormalized[m];
    }
    return (t1 - t0) + (t3 - t2);
    REF_CATCH(-1.0)
}

// dominates (nsga-ii.cpp:121-134)
int ref_dominates(const double* f1, const double* f2, int M) {
    return dominates(std::vector<double>(f1, f1 + M), std::vector<double>(f2, f2 + M)) ? 1 : 0;
}

// std::sort of this toolchain's libstdc++ on an index array keyed by `keys` (the comparator shape of
// nsga-ii.cpp:98-101), used to pin the oracle's restatement of GNU introsort on tie-heavy inputs.
void ref_std_sort_by_key(int n, const double* keys, int* idx) {
    std::sort(idx, idx + n, [keys](int a, int b) { return keys[a] < keys[b]; });
}
void ref_std_partial_sort_by_key(int n, const double* keys, int* idx) {
    std::partial_sort(idx, idx + n, idx + n, [keys](int a, int b) { return keys[a] < keys[b]; });
}

int ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
