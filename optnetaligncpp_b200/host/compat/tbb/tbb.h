// Stand-in for <tbb/tbb.h> (TBB is absent in this image): nsgaiimain.cpp uses
// tbb::parallel_for over a tbb::blocked_range<size_t> with a lambda body (:77-121, :194-238).
// The range is cut into one contiguous chunk per hardware thread, each run on its own
// std::thread -- the same "each chunk has its own PRNG" structure the reference relies on.
#pragma once
#include <algorithm>
#include <cstddef>
#include <thread>
#include <vector>

#include "spin_mutex.h"

namespace tbb {
template <typename T>
class blocked_range {
public:
    blocked_range(T b, T e, std::size_t grain = 1) : b_(b), e_(e), grain_(grain) {}
    T begin() const { return b_; }
    T end() const { return e_; }
    std::size_t size() const { return (std::size_t)(e_ - b_); }
    bool empty() const { return !(b_ < e_); }
    std::size_t grainsize() const { return grain_; }

private:
    T b_, e_;
    std::size_t grain_;
};

template <typename Range, typename Body>
void parallel_for(const Range& range, const Body& body) {
    if (range.empty()) return;
    const std::size_t n = range.size();
    std::size_t workers = std::max<std::size_t>(1, std::thread::hardware_concurrency());
    workers = std::min(workers, n);
    std::vector<std::thread> ts;
    for (std::size_t w = 0; w < workers; ++w) {
        const auto lo = range.begin() + (decltype(range.begin()))(n * w / workers);
        const auto hi = range.begin() + (decltype(range.begin()))(n * (w + 1) / workers);
        if (lo == hi) continue;
        ts.emplace_back([lo, hi, &body]() { body(Range(lo, hi)); });
    }
    for (auto& t : ts) t.join();
}
}  // namespace tbb
