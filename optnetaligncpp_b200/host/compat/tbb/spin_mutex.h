// Stand-in for <tbb/spin_mutex.h> (TBB is absent in this image): the reference only uses
// tbb::spin_mutex with its scoped_lock around the archive (steadystate.cpp:27,194,257,278).
#pragma once
#include <atomic>
#include <thread>

namespace tbb {
class spin_mutex {
public:
    spin_mutex() { flag_.clear(); }
    class scoped_lock {
    public:
        explicit scoped_lock(spin_mutex& m) : m_(m) {
            // spin briefly, then yield (TBB backs off the same way): with more worker threads than
            // cores a pure spin would starve the thread that holds the lock
            for (int spins = 0; m_.flag_.test_and_set(std::memory_order_acquire); ++spins)
                if (spins > 64) std::this_thread::yield();
        }
        ~scoped_lock() { m_.flag_.clear(std::memory_order_release); }
        scoped_lock(const scoped_lock&) = delete;
        scoped_lock& operator=(const scoped_lock&) = delete;

    private:
        spin_mutex& m_;
    };

private:
    std::atomic_flag flag_;
};
}  // namespace tbb
