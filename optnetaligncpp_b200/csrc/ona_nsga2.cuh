// NSGA-II ranking of the population's objective rows on the device:
//   nonDominatedSort     nsga-ii.cpp:14-66      dominates          nsga-ii.cpp:121-134
//   normalizeFitnesses   nsga-ii.cpp:68-84      setCrowdingDists   nsga-ii.cpp:89-118
// called in that order by nsgaiimain.cpp:157-175 on the [P, M] matrix the islands all-gather.
//
// The O(P^2 M) part -- who dominates whom -- is one pass that leaves a P x P bit matrix in HBM
// (dominators of q = row q), followed by one (order, peel) kernel pair per front:
//   peel(t)   every unranked q subtracts the dominators it has in front t (row AND front mask,
//             popcount); q whose count reaches zero joins front t+1 with the key
//             (highest position among its dominators in front t, q)
//   order(t)  ranks those keys.  That IS the reference's order: it walks front t in order and each
//             member's dominated list in index order, so q is appended when its LAST dominator of
//             front t is visited, and among the q's released by one member in index order.
// Crowding distances: one CTA per front, one objective after the other as the reference does.  The
// reference sorts with std::sort, whose placement of TIED keys is whatever GNU introsort does
// (unspecified by the standard, and it decides which of two tied individuals becomes the boundary
// point).  A tie-free front has exactly one sorted order, so the CTA sorts in parallel (bitonic
// network on order-preserving integer images of the keys) and checks neighbours for ties; only a
// front that has some falls back to one thread replaying libstdc++'s introsort step by step
// (median-of-3 to first, unguarded partition, 16-element threshold, depth limit 2*floor(lg n) with
// heap-sort fallback, final insertion sort) so that the permutation of the ties is the reference's.
#pragma once
#include <cfloat>
#include <cstdint>

namespace ona {

constexpr int kNsgaMaxM = 16;      // objectives per row (fitnessName has 10 values, Network.h:31)
constexpr int kDomCand = 256;      // candidate dominators per CTA (8 words)
constexpr int kDomRows = 64;       // rows q per CTA
constexpr int kCrowdThreads = 1024;
constexpr int kCrowdSmemItems = 14336;  // 16 B per front member; larger fronts use HBM scratch

// ------------------------------------------------------------------ dominance bit matrix
// dom[q*W + w] bit b  <=>  individual 32w+b dominates individual q.  W is the row stride in words,
// a multiple of 8 (pad words stay zero).  Lane = candidate dominator, so one ballot is one word;
// a CTA's 64 x 8 words are staged in shared memory and leave as full 32-byte sectors.
// MT > 0: the number of objectives is a compile-time constant and the lane's own row lives in registers.
template <int MT>
__global__ void __launch_bounds__(kDomCand)
k_nsga_dominators(const double* __restrict__ fit, int P, int Mrt, int W, uint32_t* __restrict__ dom) {
    const int M = MT > 0 ? MT : Mrt;
    extern __shared__ double nsga_sm[];
    double* sa = nsga_sm;                  // [M][kDomCand] candidates, objective-major
    double* sq = nsga_sm + M * kDomCand;   // [kDomRows][M]
    __shared__ uint4 tile4[kDomRows * (kDomCand / 32) / 4];
    uint32_t* tile = reinterpret_cast<uint32_t*>(tile4);
    const int tid = threadIdx.x, a0 = blockIdx.x * kDomCand, q0 = blockIdx.y * kDomRows;
    for (int i = tid; i < kDomCand * M; i += kDomCand) {
        const int a = i / M, m = i - a * M;
        sa[m * kDomCand + a] = (a0 + a < P) ? fit[(size_t)(a0 + a) * M + m] : 0.0;
    }
    for (int i = tid; i < kDomRows * M; i += kDomCand) {
        const int q = q0 + i / M;
        sq[i] = q < P ? fit[(size_t)q * M + (i % M)] : 0.0;
    }
    __syncthreads();
    const int a = a0 + tid;
    const int rows = min(kDomRows, P - q0);
    double xr[MT > 0 ? MT : 1];
#pragma unroll
    for (int m = 0; m < MT; ++m) xr[m] = sa[m * kDomCand + tid];
    for (int r = 0; r < rows; ++r) {
        bool ge = a < P, gt = false;       // dominates(), :121-134, without the early exit
        if (MT > 0) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const double y = sq[r * MT + m];
                ge &= !(xr[m] < y);
                gt |= xr[m] > y;
            }
        } else {
            for (int m = 0; m < M; ++m) {
                const double x = sa[m * kDomCand + tid], y = sq[r * M + m];
                ge &= !(x < y);
                gt |= x > y;
            }
        }
        const unsigned b = __ballot_sync(0xffffffffu, ge & gt);
        if ((tid & 31) == 0) tile[r * (kDomCand / 32) + (tid >> 5)] = b;
    }
    __syncthreads();
    if (tid < 2 * rows)   // two 16-byte halves per row
        *reinterpret_cast<uint4*>(dom + (size_t)(q0 + (tid >> 1)) * W + blockIdx.x * (kDomCand / 32) +
                                  (tid & 1) * 4) = tile4[tid];
}

// normalizeFitnesses :68-84: max starts at 0.0 and only a strictly greater value replaces it
__global__ void k_nsga_colmax(const double* __restrict__ fit, int P, int M, double* __restrict__ colmax) {
    __shared__ double red[32];
    const int m = blockIdx.x;
    double mx = 0.0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const double v = fit[(size_t)i * M + m];
        if (v > mx) mx = v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double v = __shfl_xor_sync(0xffffffffu, mx, o);
        if (v > mx) mx = v;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (red[w] > mx) mx = red[w];
        colmax[m] = mx;
    }
}
// an individual that already carries normalised values keeps them (the reference only appends, :79)
__global__ void k_nsga_normalize(const double* __restrict__ fit, const unsigned char* __restrict__ has_norm,
                                 const double* __restrict__ colmax, int P, int M, double* __restrict__ norm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * M) return;
    const int p = i / M;
    if (has_norm && has_norm[p]) return;
    norm[i] = fit[i] / colmax[i - p * M];
}

// ------------------------------------------------------------------ front peeling
struct NsgaPeel {
    const uint32_t* dom;   // [P][W]
    int* count;            // [P] dominators not yet ranked
    int* rank;             // [P] domRank, -1 = not ranked yet
    uint32_t* front;       // [2][W] membership mask of front t (t&1) / scratch for t+1
    int* pos;              // [P] position inside its front
    unsigned long long* keys;  // [P] (last dominator's position << 32 | q) of the front being built
    int* cnt;              // [2] size of front t (t&1) / of the front being built
    int* offsets;          // [P+2]
    int* order;            // [P] fronts, flattened
    int P, W;
};

// :17-35  dominator counts (row popcounts); front 0 = nobody dominates, in index order
__global__ void k_nsga_seed(NsgaPeel s) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= s.P) return;
    int c = 0;
    for (int w = lane; w < s.W; w += 32) c += __popc(__ldg(s.dom + (size_t)q * s.W + w));
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane) return;
    s.count[q] = c;
    s.rank[q] = c == 0 ? 0 : -1;
    if (c == 0) s.keys[atomicAdd(s.cnt, 1)] = (unsigned long long)q;
}

// front t: rank the keys gathered by seed / peel(t-1)
__global__ void k_nsga_order(NsgaPeel s, int t) {
    const int n = s.cnt[t & 1], base = s.offsets[t];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        s.offsets[t + 1] = base + n;
        s.cnt[(t + 1) & 1] = 0;
    }
    if (i >= n) return;
    const unsigned long long key = s.keys[i];
    int p = 0;
    for (int j = 0; j < n; ++j) p += __ldg(s.keys + j) < key ? 1 : 0;
    const int q = (int)(key & 0xffffffffull);
    s.order[base + p] = q;
    s.pos[q] = p;
    atomicOr(s.front + (size_t)(t & 1) * s.W + (q >> 5), 1u << (q & 31));
}

// :41-49  One warp per unranked row q: the row is ANDed with front t's membership mask 256 words at
// a time (the loads of a chunk are issued together), popcount = dominators that just left.
constexpr int kPeelRegs = 8;
__global__ void __launch_bounds__(256) k_nsga_peel(NsgaPeel s, int t) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t* next = s.front + (size_t)((t + 1) & 1) * s.W;
    for (int w = gt; w < s.W; w += gridDim.x * blockDim.x) next[w] = 0u;
    if (s.cnt[t & 1] == 0) return;
    const int q = gt >> 5, lane = threadIdx.x & 31;
    if (q >= s.P || s.rank[q] != -1) return;
    const uint32_t* fr = s.front + (size_t)(t & 1) * s.W;
    const uint32_t* row = s.dom + (size_t)q * s.W;
    int hit = 0, mp = -1;
    for (int c0 = lane; c0 < s.W; c0 += 32 * kPeelRegs) {
        uint32_t d[kPeelRegs];
#pragma unroll
        for (int k = 0; k < kPeelRegs; ++k) d[k] = c0 + 32 * k < s.W ? __ldg(fr + c0 + 32 * k) : 0u;
#pragma unroll
        for (int k = 0; k < kPeelRegs; ++k)
            if (d[k]) d[k] &= __ldg(row + c0 + 32 * k);
#pragma unroll
        for (int k = 0; k < kPeelRegs; ++k) {
            uint32_t v = d[k];
            hit += __popc(v);
            while (v) {
                const int b = __ffs(v) - 1;
                v &= v - 1;
                mp = max(mp, __ldg(s.pos + 32 * (c0 + 32 * k) + b));
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        hit += __shfl_xor_sync(0xffffffffu, hit, o);
        mp = max(mp, __shfl_xor_sync(0xffffffffu, mp, o));
    }
    if (lane == 0 && hit) {
        const int c = s.count[q] - hit;
        s.count[q] = c;
        if (c == 0) {
            s.rank[q] = t + 1;
            s.keys[atomicAdd(s.cnt + ((t + 1) & 1), 1)] = ((unsigned long long)(unsigned)mp << 32) | (unsigned)q;
        }
    }
}

// ------------------------------------------------------------------ small populations, one CTA
// P <= 1024 rows (the reference's usual 2 x popsize): dominator matrix, counts and the whole peel
// loop stay in one CTA's shared memory, so ranking a generation is two launches (this + the
// crowding kernel) and no host round trip.  Same keys, same order as the multi-kernel path.
constexpr int kSmallP = 1024;
__host__ __device__ inline size_t nsga_small_smem(int P) {
    const int W = (P + 31) / 32;
    return (size_t)P * 8 + ((size_t)P * W + 2 * (size_t)P + W) * 4;
}
__global__ void __launch_bounds__(1024)
k_nsga_rank_small(const double* __restrict__ fit, const unsigned char* __restrict__ has_norm,
                  double* __restrict__ norm, int P, int M, int* __restrict__ rank,
                  int* __restrict__ order, int* __restrict__ offsets, int* __restrict__ n_fronts) {
    extern __shared__ double nsga_sm[];
    const int W = (P + 31) / 32;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(nsga_sm);  // [P]
    uint32_t* dom = reinterpret_cast<uint32_t*>(keys + P);                      // [P][W]
    int* count = reinterpret_cast<int*>(dom + (size_t)P * W);                   // [P]; -1 = ranked
    int* pos = count + P;                                                       // [P]
    uint32_t* front = reinterpret_cast<uint32_t*>(pos + P);                     // [W]
    __shared__ int s_n;
    __shared__ double s_max[kNsgaMaxM];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp < M) {  // :71-76
        double mx = 0.0;
        for (int i = lane; i < P; i += 32) {
            const double v = fit[(size_t)i * M + warp];
            if (v > mx) mx = v;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double v = __shfl_xor_sync(0xffffffffu, mx, o);
            if (v > mx) mx = v;
        }
        if (lane == 0) s_max[warp] = mx;
    }
    for (int q = warp; q < P; q += 32) {  // :17-30, one warp per row, one ballot per word
        int cnt = 0;
        for (int w = 0; w < W; ++w) {
            const int a = 32 * w + lane;
            bool ge = a < P, gt = false;
            if (ge)
                for (int m = 0; m < M; ++m) {
                    const double x = fit[(size_t)a * M + m], y = fit[(size_t)q * M + m];
                    ge &= !(x < y);
                    gt |= x > y;
                }
            const unsigned b = __ballot_sync(0xffffffffu, ge & gt);
            if (lane == 0) dom[(size_t)q * W + w] = b;
            cnt += __popc(b);
        }
        if (lane == 0) count[q] = cnt;
    }
    if (tid == 0) {
        s_n = 0;
        offsets[0] = 0;
    }
    __syncthreads();
    for (int i = tid; i < P * M; i += 1024)  // :78-81
        if (!(has_norm && has_norm[i / M])) norm[i] = fit[i] / s_max[i % M];
    for (int q = tid; q < P; q += 1024) {    // :31-35
        const bool first = count[q] == 0;
        rank[q] = first ? 0 : -1;
        if (first) {
            count[q] = -1;
            keys[atomicAdd(&s_n, 1)] = (unsigned long long)q;
        }
    }
    __syncthreads();
    int t = 0, base = 0;
    for (;;) {
        const int n = s_n;
        if (n == 0) break;
        __syncthreads();
        if (tid < W) front[tid] = 0u;
        if (tid == 0) {
            s_n = 0;
            offsets[t + 1] = base + n;
        }
        unsigned long long key = 0ull;
        int p = 0;
        if (tid < n) {  // n <= P <= 1024
            key = keys[tid];
            for (int j = 0; j < n; ++j) p += keys[j] < key ? 1 : 0;
        }
        __syncthreads();
        if (tid < n) {
            const int q = (int)(key & 0xffffffffull);
            order[base + p] = q;
            pos[q] = p;
            atomicOr(front + (q >> 5), 1u << (q & 31));
        }
        __syncthreads();
        for (int q = warp; q < P; q += 32) {  // :41-49
            if (count[q] <= 0) continue;
            uint32_t d = lane < W ? (dom[(size_t)q * W + lane] & front[lane]) : 0u;
            if (!__any_sync(0xffffffffu, d != 0u)) continue;   // no dominator of q in this front
            int hit = __popc(d), mp = -1;
            while (d) {
                const int b = __ffs(d) - 1;
                d &= d - 1;
                mp = max(mp, pos[32 * lane + b]);
            }
            for (int o = 16; o > 0; o >>= 1) {
                hit += __shfl_xor_sync(0xffffffffu, hit, o);
                mp = max(mp, __shfl_xor_sync(0xffffffffu, mp, o));
            }
            if (lane == 0 && hit) {
                const int c = count[q] - hit;
                count[q] = c == 0 ? -1 : c;
                if (c == 0) {
                    rank[q] = t + 1;
                    keys[atomicAdd(&s_n, 1)] = ((unsigned long long)(unsigned)mp << 32) | (unsigned)q;
                }
            }
        }
        __syncthreads();
        base += n;
        ++t;
    }
    // fronts(1): there is always a front 0; the unused tail of `offsets` describes empty fronts
    if (tid == 0) *n_fronts = max(t, 1);
    for (int f = t + 1 + tid; f <= P; f += 1024) offsets[f] = base;
}

// ------------------------------------------------------------------ std::sort of libstdc++, on pairs
// (key[i], slot[i]) is one element; comp(i, j) = key[i] < key[j] (nsga-ii.cpp:98-101).
// The range guards [lo, hi) never trigger on NaN-free keys (the library's loops are unguarded
// there); they only keep NaN input from walking out of the arrays.
struct PairRef {
    double* k;
    int* s;
    __device__ __forceinline__ void swap(int i, int j) const {
        const double tk = k[i]; k[i] = k[j]; k[j] = tk;
        const int ts = s[i]; s[i] = s[j]; s[j] = ts;
    }
    __device__ __forceinline__ void move(int dst, int src) const { k[dst] = k[src]; s[dst] = s[src]; }
    __device__ __forceinline__ void put(int dst, double vk, int vs) const { k[dst] = vk; s[dst] = vs; }
};
__device__ inline void gnu_adjust_heap(const PairRef& a, int first, int hole, int len, double vk, int vs) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (a.k[first + child] < a.k[first + child - 1]) child--;
        a.move(first + hole, first + child);
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a.move(first + hole, first + child - 1);
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;  // __push_heap
    while (hole > top && a.k[first + parent] < vk) {
        a.move(first + hole, first + parent);
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a.put(first + hole, vk, vs);
}
__device__ inline void gnu_heap_sort(const PairRef& a, int first, int last) {  // partial_sort(f, l, l)
    const int len = last - first;
    if (len >= 2)
        for (int parent = (len - 2) / 2;; --parent) {
            gnu_adjust_heap(a, first, parent, len, a.k[first + parent], a.s[first + parent]);
            if (parent == 0) break;
        }
    while (last - first > 1) {
        --last;
        const double vk = a.k[last];
        const int vs = a.s[last];
        a.move(last, first);
        gnu_adjust_heap(a, first, 0, last - first, vk, vs);
    }
}
__device__ inline void gnu_unguarded_linear_insert(const PairRef& a, int last, int lo) {
    const double vk = a.k[last];
    const int vs = a.s[last];
    int next = last - 1;
    while (next >= lo && vk < a.k[next]) {
        a.move(last, next);
        last = next;
        --next;
    }
    a.put(last, vk, vs);
}
__device__ inline void gnu_insertion_sort(const PairRef& a, int first, int last) {
    for (int i = first + 1; i < last; ++i) {
        if (a.k[i] < a.k[first]) {
            const double vk = a.k[i];
            const int vs = a.s[i];
            for (int j = i; j > first; --j) a.move(j, j - 1);  // move_backward
            a.put(first, vk, vs);
        } else
            gnu_unguarded_linear_insert(a, i, first);
    }
}
__device__ inline void gnu_sort_pairs(double* key, int* slot, int n) {
    if (n <= 0) return;
    const PairRef a{key, slot};
    int lg = 0;
    for (int v = n; v > 1; v >>= 1) ++lg;
    // __introsort_loop; the recursion on [cut, last) becomes an explicit stack (the two halves are
    // disjoint, so the order in which they are finished does not matter)
    int sf[72], sl[72], sd[72], sp = 0;
    sf[0] = 0; sl[0] = n; sd[0] = 2 * lg; sp = 1;
    while (sp > 0) {
        --sp;
        const int first = sf[sp];
        int last = sl[sp], depth = sd[sp];
        while (last - first > 16) {
            if (depth == 0) {
                gnu_heap_sort(a, first, last);
                break;
            }
            --depth;
            const int mid = first + (last - first) / 2, x = first + 1, z = last - 1;
            // __move_median_to_first(first, first+1, mid, last-1)
            if (key[x] < key[mid]) {
                if (key[mid] < key[z]) a.swap(first, mid);
                else if (key[x] < key[z]) a.swap(first, z);
                else a.swap(first, x);
            } else if (key[x] < key[z]) a.swap(first, x);
            else if (key[mid] < key[z]) a.swap(first, z);
            else a.swap(first, mid);
            // __unguarded_partition(first+1, last, pivot = first)
            int lo = first + 1, hi = last;
            for (;;) {
                while (lo < last && key[lo] < key[first]) ++lo;
                --hi;
                while (hi > first && key[first] < key[hi]) --hi;
                if (!(lo < hi)) break;
                a.swap(lo, hi);
                ++lo;
            }
            if (sp < 72) { sf[sp] = lo; sl[sp] = last; sd[sp] = depth; ++sp; }
            last = lo;
        }
    }
    // __final_insertion_sort
    if (n > 16) {
        gnu_insertion_sort(a, 0, 16);
        for (int i = 16; i < n; ++i) gnu_unguarded_linear_insert(a, i, 0);
    } else
        gnu_insertion_sort(a, 0, n);
}

// order-preserving integer image of a double (-0.0 lands right below +0.0, NaNs at the ends;
// the tie test afterwards compares the doubles themselves)
__device__ __forceinline__ unsigned long long key_image(double v) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long u) {
    return __longlong_as_double((long long)((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u));
}

// ------------------------------------------------------------------ crowding distances, :89-118
// One CTA per front.  order_after starts as a copy of the fronts and is sorted in place, objective
// after objective, exactly as setCrowdingDists leaves its argument.
struct NsgaCrowd {
    const double* norm;   // [P][M]
    const int* offsets;   // [n_fronts + 1]
    int* order_after;     // [P]
    double* crowd;        // [P]
    double* big_key;      // [P] HBM scratch for fronts above kCrowdSmemItems
    int* big_slot;        // [P]
    int* big_old;         // [P]
    int M;
    int cap;              // front members that fit the CTA's shared memory (16 B each)
};
__global__ void __launch_bounds__(kCrowdThreads) k_nsga_crowd(NsgaCrowd c) {
    extern __shared__ double nsga_sm[];
    __shared__ int s_tie;
    const int base = c.offsets[blockIdx.x], n = c.offsets[blockIdx.x + 1] - base;
    if (n <= 0) return;
    const int tid = threadIdx.x;
    double* A;
    int *B, *C;
    if (n <= c.cap) {
        A = nsga_sm;
        B = reinterpret_cast<int*>(nsga_sm + c.cap);
        C = B + c.cap;
    } else {
        A = c.big_key + base; B = c.big_slot + base; C = c.big_old + base;
    }
    unsigned long long* U = reinterpret_cast<unsigned long long*>(A);
    int* in = c.order_after + base;
    for (int k = tid; k < n; k += kCrowdThreads) c.crowd[in[k]] = 0.0;  // :92-94
    for (int m = 0; m < c.M; ++m) {
        if (tid == 0) s_tie = 0;
        for (int k = tid; k < n; k += kCrowdThreads) {
            const int id = in[k];
            C[k] = id;
            B[k] = k;
            U[k] = key_image(c.norm[(size_t)id * c.M + m]);
        }
        __syncthreads();
        // bitonic network with every comparator ascending, so the virtual +inf padding above n
        // never moves and those pairs are simply skipped
        for (int k2 = 2; (k2 >> 1) < n; k2 <<= 1) {
            for (int i = tid; i < n; i += kCrowdThreads) {
                const int j = i ^ (k2 - 1);
                if (j > i && j < n && U[j] < U[i]) {
                    const unsigned long long tu = U[i]; U[i] = U[j]; U[j] = tu;
                    const int tb = B[i]; B[i] = B[j]; B[j] = tb;
                }
            }
            __syncthreads();
            for (int s2 = k2 >> 2; s2 > 0; s2 >>= 1) {
                for (int i = tid; i < n; i += kCrowdThreads) {
                    const int j = i | s2;
                    if (!(i & s2) && j < n && U[j] < U[i]) {
                        const unsigned long long tu = U[i]; U[i] = U[j]; U[j] = tu;
                        const int tb = B[i]; B[i] = B[j]; B[j] = tb;
                    }
                }
                __syncthreads();
            }
        }
        bool tie = false;
        for (int k = tid; k + 1 < n; k += kCrowdThreads) tie |= !(key_value(U[k]) < key_value(U[k + 1]));
        if (tie) s_tie = 1;
        __syncthreads();
        if (s_tie) {  // replay std::sort from the order the previous objective left
            for (int k = tid; k < n; k += kCrowdThreads) {
                B[k] = k;
                A[k] = c.norm[(size_t)C[k] * c.M + m];
            }
            __syncthreads();
            if (tid == 0) gnu_sort_pairs(A, B, n);
        } else {
            for (int k = tid; k < n; k += kCrowdThreads) A[k] = key_value(U[k]);
        }
        __syncthreads();
        const double denom = A[n - 1] - A[0];  // :109
        for (int k = tid; k < n; k += kCrowdThreads) {
            const int id = C[B[k]];
            if (k == 0 || k == n - 1) c.crowd[id] = DBL_MAX;          // :104-105
            else c.crowd[id] += (A[k + 1] - A[k - 1]) / denom;        // :110-113
            in[k] = id;
        }
        __syncthreads();
    }
}

}  // namespace ona
