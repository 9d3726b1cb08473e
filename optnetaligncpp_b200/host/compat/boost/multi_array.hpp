// Build-compat header for hosts without Boost (the reference needs only this one class).
//
// Minimal stand-in for <boost/multi_array.hpp> so that the UNMODIFIED reference
// sources under /root/reference compile in an image that has no Boost.  The
// reference uses exactly one instantiation, boost::multi_array<double,2>, only
// as storage: constructed from boost::extents[r][c] (blastinfo.cpp:20,
// goc.cpp:23) and indexed as (*d)[i][j] (Alignment.cpp:706,717,756,762,777,782;
// blastinfo.cpp:23,42; goc.cpp:48).  No arithmetic lives here, so results of the
// oracle cannot depend on this file beyond "zero-initialised row-major doubles",
// which is also what the real boost::multi_array provides.
#pragma once
#include <cstddef>
#include <vector>

namespace boost {

namespace detail_standin {
template <std::size_t N> struct extent_gen {
    std::size_t dims[N ? N : 1];
    extent_gen<N + 1> operator[](std::size_t d) const {
        extent_gen<N + 1> g;
        for (std::size_t i = 0; i < N; ++i) g.dims[i] = dims[i];
        g.dims[N] = d;
        return g;
    }
};
}  // namespace detail_standin

static const detail_standin::extent_gen<0> extents = detail_standin::extent_gen<0>();

template <typename T, std::size_t D> class multi_array;

template <typename T> class multi_array<T, 2> {
public:
    multi_array() : rows_(0), cols_(0) {}
    explicit multi_array(const detail_standin::extent_gen<2>& e)
        : rows_(e.dims[0]), cols_(e.dims[1]), data_(e.dims[0] * e.dims[1], T()) {}
    T* operator[](std::size_t i) { return data_.data() + i * cols_; }
    const T* operator[](std::size_t i) const { return data_.data() + i * cols_; }
    const std::size_t* shape() const { return &rows_; }  // rows_, cols_ are adjacent
    T* data() { return data_.data(); }
    const T* data() const { return data_.data(); }
    std::size_t num_elements() const { return data_.size(); }

private:
    std::size_t rows_;
    std::size_t cols_;
    std::vector<T> data_;
};

}  // namespace boost
