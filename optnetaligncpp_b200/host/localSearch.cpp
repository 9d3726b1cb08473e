// Drop-in replacement for the reference's localSearch.cpp: implements the reference's
// UNCHANGED header localSearch.h on top of the B200 engine's C ABI.  Each function cites the
// reference lines whose behaviour it reproduces (/root/reference/localSearch.cpp).
#include "localSearch.h"

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <sstream>

#include "nsga-ii.h"
#include "ona_host.hpp"

using namespace std;

namespace {

// correctHillClimb is called concurrently by the reference's worker threads, each on its own
// Alignment (steadystate.cpp:132-303, nsgaiimain.cpp:194-238).  Calls that arrive together and
// agree on (engine, objective list, iteration count) are combined into ONE batched GPU climb,
// one CTA per alignment: the first caller becomes the leader, waits a moment for company,
// runs the batch and wakes the others.
// Per-thread pinned buffer for the proposal list of the thread's current correctHillClimb call
// (pinned: the batched launch copies every caller's list with a true asynchronous DMA instead of a
// staged pageable copy, 55 GB/s instead of ~5).  Grown on demand, released with the thread.
struct PinnedList {
    int* p = nullptr;
    size_t cap = 0;
    int* need(size_t n) {
        if (n > cap) {
            if (p) ona_host_free(p);
            cap = max<size_t>(n, 1 << 16);
            p = static_cast<int*>(ona_host_alloc((long long)(cap * sizeof(int))));
            if (!p) {
                cap = 0;
                throw runtime_error("ona_b200: pinned host allocation failed");
            }
        }
        return p;
    }
    ~PinnedList() {
        if (p) ona_host_free(p);
    }
};

constexpr size_t kMtWords = 625;   // what operator<< of std::mt19937 prints: 624 state words, position

// From this many iterations on the proposals are drawn on the device (ona_hill_climb_mt): the
// generator's state travels instead of the list.  Below it (proportionalSearch's 500-iteration
// climbs) printing and parsing 625 numbers costs more than the draws.  ONA_HOST_MT_MIN overrides.
long long mt_threshold() {
    static const long long v = [] {
        const char* s = getenv("ONA_HOST_MT_MIN");
        return s ? atoll(s) : 8192ll;
    }();
    return v;
}

struct Request {
    Alignment* aln;
    const vector<fitnessName>* names;
    int* xy = nullptr;   // [iters][2], in the calling thread's pinned buffer; or
    unsigned int* mt = nullptr;   // the caller's mt19937 state (625 numbers): proposals drawn on the device
    int obj;
    long long iters;
    bool done = false;
    string error;
};

class Batcher {
public:
    void submit(ona_host::Context& ctx, Request& r) {
        unique_lock<mutex> lk(mu_);
        queue_.push_back(&r);
        while (!r.done) {
            if (leader_) {  // somebody else is driving the GPU: wait to be served (or to take over)
                cv_.wait(lk);
                continue;
            }
            leader_ = true;
            // Give concurrent callers time to join.  A launch takes as long as its slowest climb
            // whatever the batch size (one CTA per alignment, up to ~300 of them run at once), so
            // the leader waits until as many callers are queued as were around at the previous
            // launch (those it climbed plus those left waiting), for at most ~10-20 % of the launch's
            // expected duration (200 us for a 500-iteration climb, 100 ms for a 10^6-iteration one).
            // Before the first launch nothing is known about the callers: the leader waits the whole cap,
            // so that a driver's worker threads, which all reach their first climb within a few
            // milliseconds of each other, share a launch instead of queueing behind a launch of one.
            {
                const long long cap_us = min<long long>(100000, max<long long>(200, r.iters / 10));
                const auto t_end = chrono::steady_clock::now() + chrono::microseconds(cap_us);
                while (chrono::steady_clock::now() < t_end && (callers_ == 0 || queue_.size() < callers_))
                    cv_.wait_for(lk, chrono::microseconds(100));
            }
            vector<Request*> batch;
            for (Request* q : queue_)
                if (q->iters == queue_.front()->iters && *q->names == *queue_.front()->names &&
                    !q->mt == !queue_.front()->mt)
                    batch.push_back(q);
            for (Request* q : batch) queue_.erase(find(queue_.begin(), queue_.end(), q));
            callers_ = batch.size() + queue_.size();
            lk.unlock();
            run(ctx, batch);
            lk.lock();
            for (Request* q : batch) q->done = true;
            leader_ = false;
            cv_.notify_all();
        }
        lk.unlock();
        if (!r.error.empty()) throw runtime_error(r.error);
    }

private:
    static void run(ona_host::Context& ctx, vector<Request*>& batch) {
        ona_host::StatTimer st(ona_host::kStatClimb, (long long)batch.size());
        try {
            lock_guard<mutex> lk(ctx.lock());
            ctx.use_objectives(*batch[0]->names);
            vector<Alignment*> alns;
            vector<int> objs;
            vector<const int*> lists;       // every caller's own proposal list, uploaded where it lies
            for (Request* q : batch) {
                alns.push_back(q->aln);
                objs.push_back(q->obj);
                lists.push_back(q->xy);
            }
            ctx.upload(alns);
            if (batch[0]->mt) {   // generator states in, advanced states out
                vector<unsigned int> states(batch.size() * kMtWords);
                for (size_t i = 0; i < batch.size(); ++i)
                    copy(batch[i]->mt, batch[i]->mt + kMtWords, states.begin() + i * kMtWords);
                ona_host::check(ona_hill_climb_mt(ctx.engine(), 0, (int)alns.size(), batch[0]->iters, states.data(),
                                                  objs.data(), nullptr, nullptr));
                for (size_t i = 0; i < batch.size(); ++i)
                    copy(states.begin() + i * kMtWords, states.begin() + (i + 1) * kMtWords, batch[i]->mt);
            } else
                ona_host::check(ona_hill_climb_lists(ctx.engine(), 0, (int)alns.size(), batch[0]->iters,
                                                     lists.data(), objs.data(), nullptr, nullptr));
            ctx.download(alns);
            for (Request* q : batch) q->aln->computeFitness(*q->names);
        } catch (const exception& ex) {
            for (Request* q : batch) q->error = ex.what();
        }
    }
    mutex mu_;
    condition_variable cv_;
    vector<Request*> queue_;
    bool leader_ = false;
    size_t callers_ = 0;   // callers around at the previous launch (0: no launch yet)
};

Batcher& batcher_for(ona_host::Context& ctx) {
    static mutex mu;
    static map<ona_host::Context*, Batcher*> all;
    lock_guard<mutex> lk(mu);
    Batcher*& b = all[&ctx];
    if (!b) b = new Batcher();
    return *b;
}

}  // namespace

// localSearch.cpp:18-55.  The proposals are drawn from the caller's PRNG with the reference's
// own two distributions in the reference's order (:21-28), so the generator ends in the same
// state and the climb follows the same trajectory; the accept/undo loop runs on the GPU with
// the reference's exact sequential semantics (ona_hill_climb).
void correctHillClimb(RandGenT& prng, Alignment* aln, bool total, int maxIters,
                      const vector<fitnessName>& fitnessNames, int obj) {
    (void)total;  // unused by the reference as well
    if (maxIters <= 0) return;
    auto randNonDummyIndex = uniform_int_distribution<int>(0, aln->actualSize - 1);
    auto randIndex = uniform_int_distribution<int>(0, (int)aln->aln.size() - 2);
    Request r;
    r.aln = aln;
    r.names = &fitnessNames;
    r.obj = obj;
    r.iters = maxIters;
    ona_host::Context& ctx = ona_host::context_for(*aln);
    if (maxIters >= mt_threshold() && aln->actualSize >= 1 && aln->aln.size() >= 2) {
        // the stream of :21-28 is generated on the device from the generator's state (operator<<
        // gives the 624 words and the position, operator>> puts the advanced state back), so the
        // caller's prng ends exactly where the reference's loop leaves it
        unsigned int state[kMtWords];
        {
            stringstream ss;
            ss << prng;
            for (size_t i = 0; i < kMtWords; ++i) {
                unsigned long v;
                ss >> v;
                state[i] = (unsigned int)v;
            }
        }
        r.mt = state;
        {
            ona_host::StatTimer st(ona_host::kStatClimbWait, maxIters);
            batcher_for(ctx).submit(ctx, r);
        }
        stringstream ss;
        for (size_t i = 0; i < kMtWords; ++i) ss << state[i] << ' ';
        ss >> prng;
        return;
    }
    thread_local PinnedList pinned;
    r.xy = pinned.need(2 * (size_t)maxIters);
    for (int i = 0; i < maxIters; i++) {
        node x = randNonDummyIndex(prng);
        node y = randIndex(prng);
        if (y >= x) y++;
        r.xy[2 * i] = x;
        r.xy[2 * i + 1] = y;
    }
    ona_host::StatTimer st(ona_host::kStatClimbWait, maxIters);
    batcher_for(ctx).submit(ctx, r);
}

// localSearch.cpp:60-84
void proportionalSearch(RandGenT& prng, Alignment* aln, bool total, int iters,
                        const vector<fitnessName>& fitnessNames, int obj, double proportion) {
    auto prob = uniform_real_distribution<double>(0, 1);
    auto randObj = uniform_int_distribution<int>(0, (int)fitnessNames.size() - 1);
    for (int i = 0; i < iters; i++) {
        if (prob(prng) < proportion) {
            correctHillClimb(prng, aln, total, 500, fitnessNames, obj);
        } else {
            int robj = obj;
            while (robj == obj) robj = randObj(prng);
            correctHillClimb(prng, aln, total, 500, fitnessNames, robj);
        }
    }
}

// localSearch.cpp:88-120.  One candidate against the alignment's current state; the state is
// left exactly as it was (the reference's swap/un-swap can leave an ulp-level residue in the
// fp64 accumulators, SURVEY H3 -- not reproduced here).
double swapNormalizedDelta(Alignment& aln, const vector<fitnessName>& fitnessNames, node x, node y) {
    ona_host::Context& ctx = ona_host::context_for(aln);
    lock_guard<mutex> lk(ctx.lock());
    ctx.use_objectives(fitnessNames);
    vector<Alignment*> one{&aln};
    ctx.upload(one);
    const int xy[2] = {x, y};
    double snd = 0.0;
    ona_host::check(ona_swap_deltas(ctx.engine(), 0, 1, 1, xy, -1, nullptr, nullptr, nullptr, nullptr, nullptr,
                                    nullptr, &snd));
    return snd;
}

// localSearch.cpp:122-203.  The reference splits the n1 x n2 grid over nthreads CPU threads;
// here every scan is one pass of the swap-delta kernel over the whole grid.
void steepestAscentHillClimb(Alignment* aln, vector<fitnessName>& fitnessNames, int nthreads, bool verbose) {
    (void)nthreads;
    ona_host::Context& ctx = ona_host::context_for(*aln);
    int numiters = 0;
    for (;;) {
        double best = -1.0;
        int bx = -1, by = -1;
        {
            lock_guard<mutex> lk(ctx.lock());
            ctx.use_objectives(fitnessNames);
            vector<Alignment*> one{aln};
            ctx.upload(one);
            ona_host::check(ona_steepest_step(ctx.engine(), 0, &best, &bx, &by));
        }
        if (best <= 0.0) break;  // :190
        aln->doSwap(bx, by);
        aln->computeFitness(fitnessNames);
        if (verbose) {
            reportStats({aln}, fitnessNames, true, false);
            cout << (++numiters) << " swaps performed." << endl;
        }
    }
}

// localSearch.cpp:298-330 -- ring buffer of the last 500 delta vectors
VelocityTracker::VelocityTracker() : nextSpot(0), size(0), recentDeltas(500) {}
void VelocityTracker::reportDelta(const vector<double>& in) {
    recentDeltas[nextSpot] = in;
    nextSpot = (nextSpot + 1) % 500;
    if (size < 500) size++;
}
vector<double> VelocityTracker::getRecentVel() const {
    vector<double> mean(recentDeltas[0].size(), 0.0);
    for (int i = 0; i < size; i++)
        for (size_t j = 0; j < recentDeltas[i].size(); j++) mean[j] += recentDeltas[i][j];
    const double n = double(size);
    for (double& v : mean) v /= n;
    return mean;
}
