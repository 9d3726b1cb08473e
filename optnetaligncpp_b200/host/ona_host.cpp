// see ona_host.hpp
#include "ona_host.hpp"

#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <tuple>

#include <atomic>
#include <chrono>
#include <cstdio>

namespace ona_host {

namespace {
const char* kStatNames[kStatSlots] = {"engine creation", "from-scratch evaluations", "swap lists", "edit lists",
                                      "hill-climb launches (individuals)", "correctHillClimb calls (caller wait)",
                                      "steepest scans", "NSGA-II rankings", "  (of the short calls: waiting for their queue)"};
std::atomic<long long> g_stat_ns[kStatSlots], g_stat_calls[kStatSlots], g_stat_units[kStatSlots];
bool stats_on() {
    static const bool on = [] {
        const bool v = std::getenv("ONA_HOST_STATS") != nullptr;
        if (v)
            std::atexit([] {
                std::fprintf(stderr, "ona_host stats (wall time inside the mirror's GPU round trips)\n");
                for (int i = 0; i < kStatSlots; ++i)
                    if (g_stat_calls[i].load())
                        std::fprintf(stderr, "  %-40s calls %8lld  units %10lld  %9.3f s\n", kStatNames[i],
                                     g_stat_calls[i].load(), g_stat_units[i].load(), g_stat_ns[i].load() * 1e-9);
            });
        return v;
    }();
    return on;
}
double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
std::mutex g_mu;
std::map<std::tuple<const Network*, const Network*, const BLASTDict*, const GOCDict*>,
         std::unique_ptr<Context>> g_ctx;

// Network::adjList -> CSR in adjList order (Network.cpp:89-92)
void to_csr(const Network* n, std::vector<int>& rp, std::vector<int>& col) {
    rp.assign(1, 0);
    col.clear();
    for (const auto& row : n->adjList) {
        col.insert(col.end(), row.begin(), row.end());
        rp.push_back((int)col.size());
    }
}
}  // namespace

}  // namespace ona_host
ona_host::StatTimer::StatTimer(int s, long long u) : slot(s), units(u), t0(stats_on() ? now_s() : 0.0) {}
ona_host::StatTimer::~StatTimer() {
    if (!stats_on()) return;
    g_stat_ns[slot] += (long long)((now_s() - t0) * 1e9);
    g_stat_calls[slot] += 1;
    g_stat_units[slot] += units;
}
namespace ona_host {

Context::Context(const Network* net1, const Network* net2, const BLASTDict* bit, const GOCDict* goc) {
    StatTimer st(kStatCreate);
    n1_ = (int)net1->nodeToNodeName.size();
    n2_ = (int)net2->nodeToNodeName.size();
    std::vector<int> rp1, c1, rp2, c2;
    to_csr(net1, rp1, c1);
    to_csr(net2, rp2, c2);
    ona_problem pr{};
    pr.n1 = n1_;
    pr.n2 = n2_;
    pr.num_edges1 = (long long)net1->edges.size();  // EC / S3 denominator term (Alignment.cpp:536,541)
    pr.rowptr1 = rp1.data(); pr.col1 = c1.data();
    pr.rowptr2 = rp2.data(); pr.col2 = c2.data();
    // BLASTDict / GOCDict are dense row-major n1 x n2 doubles (blastinfo.cpp:20, goc.cpp:23)
    if (bit) pr.bitscores.dense = &(*bit)[0][0];
    if (goc) pr.gocs.dense = &(*goc)[0][0];
    names_ = {ONA_ICS};
    pr.objectives = names_.data();
    pr.num_objectives = 1;
    int device = 0;
    if (const char* d = std::getenv("ONA_DEVICE")) device = std::atoi(d);
    // one stream per queue: ask for enough hardware connections that none of them shares a work
    // queue with the main queue's long launches (read by the driver when CUDA initialises, i.e.
    // here in a drop-in binary; the default is 8)
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    check(ona_engine_create(device, &pr, &main_.e_));
    main_.n2_ = n2_;
    int nq = 8;
    if (const char* q = std::getenv("ONA_HOST_QUEUES")) nq = std::max(1, std::min(64, std::atoi(q)));
    for (int i = 0; i < nq; ++i) {
        small_.emplace_back(new Queue());
        check(ona_engine_fork(main_.e_, &small_.back()->e_));
        small_.back()->n2_ = n2_;
    }
}

Context::~Context() {
    for (auto& q : small_) ona_engine_destroy(q->e_);
    ona_engine_destroy(main_.e_);
}

std::unique_lock<std::mutex> Context::acquire_small(Queue*& q) {
    static std::atomic<unsigned> next{0};
    thread_local unsigned home = next++;
    const size_t n = small_.size();
    for (size_t k = 0; k < n; ++k) {
        Queue* cand = small_[(home + k) % n].get();
        std::unique_lock<std::mutex> lk(cand->mu_, std::try_to_lock);
        if (lk.owns_lock()) {
            q = cand;
            return lk;
        }
    }
    q = small_[home % n].get();
    return std::unique_lock<std::mutex>(q->mu_);
}

void Context::use_objectives(const std::vector<fitnessName>& names) {
    std::vector<int> v(names.begin(), names.end());
    if (v == names_ || v.empty()) return;
    check(ona_set_objectives(main_.e_, v.data(), (int)v.size()));
    names_ = v;
}

void Queue::ensure_slots(int count) {
    if (count <= slots_) return;
    int want = slots_ ? slots_ : 16;
    while (want < count) want *= 2;
    check(ona_pop_resize(e_, want));
    slots_ = want;
}

void Queue::upload(const std::vector<Alignment*>& alns) {
    const int n = (int)alns.size();
    ensure_slots(n);
    h_aln_.resize((size_t)n * n2_);
    h_mask_.resize((size_t)n * n2_);
    std::vector<int> cons(n), ind(n);
    std::vector<double> bits(n), goc(n);
    bool any_off = false;
    for (int p = 0; p < n; ++p) {
        const Alignment& a = *alns[p];
        for (int i = 0; i < n2_; ++i) {
            h_aln_[(size_t)p * n2_ + i] = a.aln[i];
            const unsigned char m = a.alnMask[i] ? 1 : 0;
            h_mask_[(size_t)p * n2_ + i] = m;
            any_off |= !m;
        }
        cons[p] = a.currConservedCount;
        ind[p] = a.currInducedCount;
        bits[p] = a.bitscores ? a.currBitscore : 0.0;
        goc[p] = a.gocs ? a.currGOC : 0.0;
    }
    check(ona_pop_upload(e_, 0, n, h_aln_.data(), any_off ? h_mask_.data() : nullptr));
    check(ona_pop_set_counters(e_, 0, n, cons.data(), ind.data(), bits.data(), goc.data()));
}

void Queue::download(const std::vector<Alignment*>& alns) {
    const int n = (int)alns.size();
    h_aln_.resize((size_t)n * n2_);
    h_inv_.resize((size_t)n * n2_);
    h_mask_.resize((size_t)n * n2_);
    std::vector<int> cons(n), ind(n);
    std::vector<double> bits(n), goc(n);
    check(ona_pop_download(e_, 0, n, h_aln_.data(), h_inv_.data(), h_mask_.data()));
    check(ona_pop_get_counters(e_, 0, n, cons.data(), ind.data(), bits.data(), goc.data(), nullptr, nullptr));
    for (int p = 0; p < n; ++p) {
        Alignment& a = *alns[p];
        for (int i = 0; i < n2_; ++i) {
            a.aln[i] = h_aln_[(size_t)p * n2_ + i];
            a.alnInv[i] = h_inv_[(size_t)p * n2_ + i];
            a.alnMask[i] = h_mask_[(size_t)p * n2_ + i] != 0;
        }
        a.currConservedCount = cons[p];
        a.currInducedCount = ind[p];
        if (a.bitscores) a.currBitscore = bits[p];
        if (a.gocs) a.currGOC = goc[p];
    }
}

Context& context_for(const Network* n1, const Network* n2, const BLASTDict* bit, const GOCDict* goc) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(n1, n2, bit, goc);
    auto it = g_ctx.find(key);
    if (it == g_ctx.end()) it = g_ctx.emplace(key, std::unique_ptr<Context>(new Context(n1, n2, bit, goc))).first;
    return *it->second;
}
Context& context_for(const Alignment& a) { return context_for(a.net1, a.net2, a.bitscores, a.gocs); }

Scratch score_from_scratch(const Alignment& a) {
    StatTimer st(kStatScratch);
    StatTimer* wait = new StatTimer(kStatSmallLockWait);
    Queue* qp = nullptr;
    std::unique_lock<std::mutex> lk = context_for(a).acquire_small(qp);
    Queue& c = *qp;
    delete wait;
    std::vector<Alignment*> one{const_cast<Alignment*>(&a)};
    c.ensure_slots(1);
    // only the permutation and mask matter here; the cached counters are left alone
    std::vector<int> aln(a.aln.begin(), a.aln.end());
    std::vector<unsigned char> mask(a.alnMask.size());
    bool any_off = false;
    for (size_t i = 0; i < mask.size(); ++i) {
        mask[i] = a.alnMask[i] ? 1 : 0;
        any_off |= !mask[i];
    }
    check(ona_pop_upload(c.engine(), 0, 1, aln.data(), any_off ? mask.data() : nullptr));
    Scratch s{};
    check(ona_full_fitness(c.engine(), 0, 1, ONA_SEED_NONE | ONA_SUM_SEQUENTIAL, &s.cons_true, &s.cons_refseed,
                           &s.induced, &s.bits, &s.goc, &s.size, nullptr, &s.ics));
    return s;
}

void seed_counters(Alignment& a) {
    // Alignment.cpp:437-446: currBitscore = sumBLAST(); currGOC = sumGOC();
    // currConservedCount = ics()*denom (truncated, H1); currInducedCount = (int)denom
    const Scratch s = score_from_scratch(a);
    if (a.bitscores) a.currBitscore = s.bits;
    if (a.gocs) a.currGOC = s.goc;
    a.currConservedCount = s.cons_refseed;
    a.currInducedCount = s.induced;
}

void apply_swaps(Alignment& a, const std::vector<int>& xy) {
    if (xy.empty()) return;
    StatTimer st(kStatSwaps, (long long)xy.size() / 2);
    StatTimer* wait = new StatTimer(kStatSmallLockWait);
    Queue* qp = nullptr;
    std::unique_lock<std::mutex> lk = context_for(a).acquire_small(qp);
    Queue& c = *qp;
    delete wait;
    std::vector<Alignment*> one{&a};
    c.upload(one);
    check(ona_apply_swaps(c.engine(), 0, 1, (int)(xy.size() / 2), xy.data()));
    c.download(one);
}

void apply_edits(Alignment& a, const std::vector<int>& ops) {
    if (ops.empty()) return;
    StatTimer st(kStatEdits, (long long)ops.size() / 4);
    StatTimer* wait = new StatTimer(kStatSmallLockWait);
    Queue* qp = nullptr;
    std::unique_lock<std::mutex> lk = context_for(a).acquire_small(qp);
    Queue& c = *qp;
    delete wait;
    std::vector<Alignment*> one{&a};
    c.upload(one);
    check(ona_apply_edits(c.engine(), 0, 1, (int)(ops.size() / 4), ops.data()));
    c.download(one);
}

std::vector<std::vector<Alignment*> > rankGeneration(const std::vector<Alignment*>& in, bool crowding) {
    std::vector<std::vector<Alignment*> > fronts(1);  // nsga-ii.cpp:16: there is always a front 0
    if (in.empty()) return fronts;
    const int P = (int)in.size(), M = (int)in[0]->fitness.size();
    std::vector<double> fit((size_t)P * M), norm((size_t)P * M, 0.0), crowd(P);
    std::vector<unsigned char> has(P, 0);
    for (int i = 0; i < P; i++) {
        std::copy(in[i]->fitness.begin(), in[i]->fitness.end(), fit.begin() + (size_t)i * M);
        // normalizeFitnesses only appends (:79): whoever was normalised in an earlier generation
        // keeps being read through its first M entries
        if ((int)in[i]->fitnessNormalized.size() >= M) {
            has[i] = 1;
            std::copy(in[i]->fitnessNormalized.begin(), in[i]->fitnessNormalized.begin() + M,
                      norm.begin() + (size_t)i * M);
        }
    }
    std::vector<int> rank(P), order(P), after(P), offs(P + 1);
    int nf = 0;
    Context& c = context_for(*in[0]);
    {
        std::lock_guard<std::mutex> lk(c.lock());
        check(ona_nsga2_rank(c.engine(), P, M, fit.data(), has.data(), norm.data(), rank.data(), order.data(),
                             offs.data(), &nf, after.data(), crowd.data()));
    }
    for (int i = 0; i < P; i++) {
        in[i]->domRank = rank[i];
        in[i]->numThatDominate = 0;
        in[i]->dominated.clear();
    }
    if (crowding) {
        for (int m = 0; m < M; m++) {  // :68-84, the append itself
            double mx = 0.0;
            for (int i = 0; i < P; i++)
                if (fit[(size_t)i * M + m] > mx) mx = fit[(size_t)i * M + m];
            for (int i = 0; i < P; i++) in[i]->fitnessNormalized.push_back(fit[(size_t)i * M + m] / mx);
        }
        for (int i = 0; i < P; i++) in[i]->crowdDist = crowd[i];
    }
    fronts.assign(nf, std::vector<Alignment*>());
    const std::vector<int>& src = crowding ? after : order;
    for (int f = 0; f < nf; f++)
        for (int k = offs[f]; k < offs[f + 1]; k++) fronts[f].push_back(in[src[k]]);
    return fronts;
}

void writeResults(const std::string& outprefix, const std::vector<Alignment*>& alns,
                  const std::vector<fitnessName>& fitnessNames) {
    std::ofstream info(outprefix + ".info");
    if (!info) throw std::runtime_error("cannot write " + outprefix + ".info");
    const fitnessName columns[] = {ECFit, ICSFit, S3Fit, GOCFit, BitscoreSumFit, EvalsSumFit, SizeFit, ICSTimesEC, S3Variant};
    info << "filename";
    for (fitnessName col : columns) {
        info << '\t' << fitnessNameToStr(col);
        if (std::find(fitnessNames.begin(), fitnessNames.end(), col) != fitnessNames.end()) info << "*";
    }
    info << std::endl;
    for (size_t i = 0; i < alns.size(); i++) {
        const Alignment* a = alns[i];
        const std::string file = outprefix + "_" + std::to_string(i) + ".aln";
        a->save(file);
        info << file << '\t' << a->fastEC() << '\t' << a->fastICS() << '\t' << a->fastS3() << '\t';
        if (a->gocs) info << a->currGOC << '\t';
        else info << '?' << '\t';
        // one accumulator serves both BLAST columns in the reference as well (:375-380)
        if (a->bitscores) info << a->currBitscore << '\t' << a->currBitscore << '\t';
        else info << '?' << '\t' << '?' << '\t';
        info << a->alnSize() << '\t' << a->icsTimesEC() << '\t' << a->s3Variant() << std::endl;
    }
}

}  // namespace ona_host
