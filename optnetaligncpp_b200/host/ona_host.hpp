// C++ host mirror of the reference's Alignment / localSearch interface on top of the C ABI
// (include/ona_b200.h).  Shared plumbing of the two drop-in translation units
//   host/Alignment.cpp   -- replaces the reference's Alignment.cpp
//   host/localSearch.cpp -- replaces the reference's localSearch.cpp
// which implement the reference's UNCHANGED headers (Alignment.h, localSearch.h), so that
// steadystate.cpp / nsgaiimain.cpp / hillclimber.cpp / nsga-ii.cpp / Archive.cpp compile and
// link against them as they are.
//
// Design: a host Alignment object stays the source of truth (the reference copies, assigns
// and deletes Alignments freely: steadystate.cpp:204-219, Archive.cpp:30-66).  For every
// batched operation the object's permutation, mask and cached counters are uploaded into a
// scratch slot of the engine, the GPU does the graph work, and the results are copied back.
// One engine exists per (net1, net2, bitscores, gocs) tuple and per GPU; its objective list
// follows the fitnessNames vector of the call.
#pragma once
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "Alignment.h"
#include "../../include/ona_b200.h"

namespace ona_host {

// ONA_HOST_STATS=1: wall time and call counts of the mirror's GPU round trips, printed at exit
// (where a drop-in driver's time goes: engine creation, evaluations, swap / edit lists, climbs)
struct StatTimer {
    StatTimer(int slot, long long units = 1);
    ~StatTimer();
    int slot;
    long long units;
    double t0;
};
enum { kStatCreate = 0, kStatScratch, kStatSwaps, kStatEdits, kStatClimb, kStatClimbWait, kStatSteepest, kStatRank, kStatSmallLockWait, kStatSlots };

// non-zero C-ABI codes become exceptions again (the reference reports errors by throwing)
inline void check(int rc) {
    if (rc != ONA_OK) throw std::runtime_error(std::string("ona_b200: ") + ona_last_error());
}

// One command queue of the engine: scratch population slots, a lock, host staging buffers.
class Queue {
public:
    ona_engine* engine() { return e_; }
    std::mutex& lock() { return mu_; }
    // scratch slots [0, count) of the engine, grown on demand; caller holds lock()
    void ensure_slots(int count);
    // upload `alns` into slots 0..n-1 (permutation, mask, cached counters); caller holds lock()
    void upload(const std::vector<Alignment*>& alns);
    // copy slots 0..n-1 back into the objects (permutation, inverse, mask, counters)
    void download(const std::vector<Alignment*>& alns);

private:
    friend class Context;
    ona_engine* e_ = nullptr;
    int n2_ = 0, slots_ = 0;
    std::mutex mu_;
    std::vector<int> h_aln_, h_inv_;
    std::vector<unsigned char> h_mask_;
};

// The engine of one (net1, net2, bitscores, gocs) tuple with TWO queues on it: the batched hill
// climbs of the drivers' worker threads run on the main queue, the short per-alignment calls
// (seeding a new alignment's counters, a child's crossover / mutation swap list) on a few forks of
// it (ona_engine_fork: same device data, own slots / stream / lock), so that a worker preparing
// its next child neither queues behind a climb that runs for a second nor behind every other
// worker doing the same.
class Context {
public:
    Context(const Network* n1, const Network* n2, const BLASTDict* bit, const GOCDict* goc);
    ~Context();
    ona_engine* engine() { return main_.e_; }
    std::mutex& lock() { return main_.mu_; }
    Queue& main() { return main_; }
    // one of the side queues, locked: a free one if there is any, else the caller's home queue
    std::unique_lock<std::mutex> acquire_small(Queue*& q);
    int n1() const { return n1_; }
    int n2() const { return n2_; }
    // make `names` the main queue's objective list (no-op when unchanged); caller holds lock()
    void use_objectives(const std::vector<fitnessName>& names);
    void ensure_slots(int count) { main_.ensure_slots(count); }
    void upload(const std::vector<Alignment*>& alns) { main_.upload(alns); }
    void download(const std::vector<Alignment*>& alns) { main_.download(alns); }

private:
    Queue main_;
    std::vector<std::unique_ptr<Queue> > small_;   // kSmallQueues forks of the main engine
    int n1_ = 0, n2_ = 0;
    std::vector<int> names_;
};

// the engine of the networks / score tables an Alignment borrows (created on first use)
Context& context_for(const Alignment& a);
Context& context_for(const Network* n1, const Network* n2, const BLASTDict* bit, const GOCDict* goc);

// from-scratch counters of one alignment (Alignment.cpp:437-446 idiom) on the GPU
struct Scratch {
    int cons_true, cons_refseed, induced;
    double bits, goc, size, ics;
};
Scratch score_from_scratch(const Alignment& a);
// seed a's cached counters from scratch exactly like the reference's constructors / shuf
void seed_counters(Alignment& a);
// apply an ordered list of doSwap(x,y) to `a` on the GPU (incremental counter updates)
void apply_swaps(Alignment& a, const std::vector<int>& xy);
// apply an ordered list of {kind, a, b, c} edits (ONA_EDIT_*: doSwap, mask flip, onBit, UPMX step
// with mask AND/OR) to `a` on the GPU -- the partial-alignment operators, bit for bit
void apply_edits(Alignment& a, const std::vector<int>& ops);


// nsgaiimain.cpp:157-175 as ONE engine call (ona_nsga2_rank): nonDominatedSort(in)
// (nsga-ii.cpp:14-66), normalizeFitnesses(in) (:68-84) and setCrowdingDists on every front
// (:89-118).  Sets domRank, numThatDominate (0, as the peeling leaves it), crowdDist and appends
// the normalised objectives exactly like the reference; returns the fronts in the order
// setCrowdingDists leaves them (pass crowding = false for nonDominatedSort's own order and
// nothing but domRank set).  The `dominated` lists -- scratch of the reference's sort, read by
// nobody else -- are left empty.
std::vector<std::vector<Alignment*> > rankGeneration(const std::vector<Alignment*>& in, bool crowding = true);


// The result files optnetalign writes for its archive (steadystate.cpp:339-395):
//   <outprefix>.info       TSV, one row per alignment: file name, EC, ICS, S3, GOC, BitscoreSum,
//                          EvalsSum, Size, ICS*EC, S3Variant; a '*' after the names of the
//                          optimised objectives; '?' where the score data was not given
//   <outprefix>_<i>.aln    Alignment::save
void writeResults(const std::string& outprefix, const std::vector<Alignment*>& alns,
                  const std::vector<fitnessName>& fitnessNames);

}  // namespace ona_host
