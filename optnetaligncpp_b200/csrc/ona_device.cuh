// Device-side building blocks of the OptNetAlign scoring engine (sm_100a).
//
// Everything here restates, in a state-preserving closed form, what the reference
// computes by mutate -> recount -> mutate-back:
//   Alignment::conservedCount  Alignment.cpp:791-808
//   Alignment::inducedCount    Alignment.cpp:812-826
//   Alignment::updateBitscore / updateGOC  Alignment.cpp:745-785
//   Alignment::doSwap          Alignment.cpp:491-520
//   Alignment::computeFitness  Alignment.cpp:559-602 (+ fastEC/fastS3/... :535-556, :693-699)
//   accept test of correctHillClimb  localSearch.cpp:35-47
//   swapNormalizedDelta        localSearch.cpp:98-119
//
// Compiled with -fmad=false: every fp64 expression keeps the reference's separate
// multiply / add / divide roundings (the reference is built for plain x86-64, no FMA).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ona {

constexpr int kMaxObj = 10;

// objective ids == enum fitnessName, Network.h:31-32
enum : int {
    kICS = 0, kEC = 1, kBitscoreSum = 2, kEvalsSum = 3, kSize = 4,
    kGOC = 5, kS3 = 6, kS3Denom = 7, kICSTimesEC = 8, kS3Variant = 9
};

// ---------------------------------------------------------------- layouts in HBM
// Packed per-individual arrays.  aln[i] = image of V1 position i, top bit set when
// alnMask[i] is false; inv[v] = pre-image of V2 node v, top bit set when the pre-image's
// mask bit is false.  With the flag folded in, "pre-image is a real, unmasked node of
// net1" is the single unsigned compare inv[v] < n1, and a masked-off neighbour costs no
// extra load.  In shared memory the same arrays are held as T = uint16_t when
// n2 < 32768 (flag = bit 15) and as uint32_t otherwise.
template <typename T> struct Pk;
template <> struct Pk<uint16_t> {
    static constexpr uint32_t FLAG = 0x8000u, MASK = 0x7fffu;
};
template <> struct Pk<uint32_t> {
    static constexpr uint32_t FLAG = 0x80000000u, MASK = 0x7fffffffu;
};
constexpr uint32_t kGFlag = 0x80000000u, kGMask = 0x7fffffffu;  // global (HBM) form

struct DevGraph {
    const int* rp1;   // net1 CSR row pointers [n1+1]       (Network::adjList, Network.cpp:89-92)
    const int* col1;  // net1 CSR columns
    const int* rp2;   // net2 CSR
    const int* col2;
    const uint32_t* adj2;  // net2 adjacency bit matrix [n2][W] (Network::adjMatrix, :99-109), rows 128 B aligned
    int W;                 // words per row
    int n1, n2;
    const int* eu1;  // unique undirected edges of net1 (u <= v), [ne1]
    const int* ev1;
    const int* eu2;  // unique undirected edges of net2
    const int* ev2;
    int ne1, ne2;
    unsigned long long E1;  // net1->edges.size()
};

// BLASTDict / GOCDict on the device: dense fp64 matrix, or an open-addressing hash of the
// non-zeros (key = row * n2 + col, 16-byte slots {key, value bits}, linear probing).
struct DevTable {
    int kind;  // 0 absent, 1 dense, 2 hash
    const double* dense;
    const ulonglong2* slots;
    unsigned long long cap_mask;
    int shift;  // 64 - log2(capacity)
    int n2;
};
constexpr unsigned long long kEmptyKey = ~0ull;

struct DevObjectives {
    int M;
    int names[kMaxObj];
};

struct DevPop {
    uint32_t* aln;  // [P][n2p] packed
    uint32_t* inv;  // [P][n2p] packed
    int n2p;        // row stride (n2 rounded up to a multiple of 4)
    int* cons;      // currConservedCount
    int* ind;       // currInducedCount
    int* size;      // alnSize() as an integer
    double* bits;   // currBitscore
    double* goc;    // currGOC
    double* fit;    // [P][M]
};

// per-candidate outputs of the swap-delta kernel; null = not wanted
struct DevDeltaOut {
    int* cons;
    int* ind;
    double* bits;
    double* goc;
    double* fit;
    unsigned char* improved;
    double* snd;
};

// ------------------------------------------------------------------ small helpers
__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// Counter-based proposal stream: candidate i of global individual gp under `seed`.
// Supports as localSearch.cpp:21-28: x in [0,n1), y in [0,n2) \ {x}.
__host__ __device__ __forceinline__ void rng_proposal(unsigned long long ind_key,
                                                      unsigned long long i, int n1, int n2,
                                                      int& x, int& y) {
    unsigned long long r = mix64(ind_key + i);
    unsigned int lo = (unsigned int)r, hi = (unsigned int)(r >> 32);
    x = (int)(((unsigned long long)lo * (unsigned long long)n1) >> 32);
    y = (int)(((unsigned long long)hi * (unsigned long long)(n2 - 1)) >> 32);
    if (y >= x) y++;
}
__host__ __device__ __forceinline__ unsigned long long rng_individual_key(
    unsigned long long seed, unsigned long long global_individual) {
    return mix64(mix64(seed) + global_individual);
}

__device__ __forceinline__ int probe(const DevGraph& g, int row, int col) {
    return (int)((__ldg(g.adj2 + (size_t)row * (size_t)g.W + (col >> 5)) >> (col & 31)) & 1u);
}

__device__ __forceinline__ double table_lookup(const DevTable& t, int r, int c) {
    if (t.kind == 1) return __ldg(t.dense + (size_t)r * (size_t)t.n2 + c);
    unsigned long long key = (unsigned long long)r * (unsigned long long)t.n2 + (unsigned long long)c;
    unsigned long long h = (key * 0x9E3779B97F4A7C15ull) >> t.shift;
    for (;;) {
        ulonglong2 s = __ldg(t.slots + h);
        if (s.x == key) return __longlong_as_double((long long)s.y);
        if (s.x == kEmptyKey) return 0.0;
        h = (h + 1) & t.cap_mask;
    }
}

// ---------------------------------------------------------------- objectives (S5)
struct Scalars {
    int C, I, size;  // currConservedCount, currInducedCount, alnSize()
    double bits, goc;
};

// Alignment::computeFitness, Alignment.cpp:559-602
__device__ __forceinline__ double objective(int name, const Scalars& s, unsigned long long E1) {
    switch (name) {
        case kICS: {  // fastICS :693-699
            double denom = (double)s.I;
            return denom == 0.0 ? 0.0 : ((double)s.C) / denom;
        }
        case kEC:  // fastEC :535-537
            return (double)s.C / (double)E1;
        case kBitscoreSum: return s.bits;
        case kEvalsSum: return -1.0 * s.bits;
        case kSize: return (double)s.size;
        case kGOC: return s.goc;
        case kS3: {  // fastS3 :539-544   size_t + int, then double
            double num = (double)s.C;
            double denom = (double)(E1 + (unsigned long long)(long long)s.I) - num;
            return num / denom;
        }
        case kS3Denom:  // :589-592
            return -((double)(E1 + (unsigned long long)(long long)s.I) - (double)s.C);
        case kICSTimesEC: {  // :546-548
            double denom = (double)s.I;
            double ics = denom == 0.0 ? 0.0 : ((double)s.C) / denom;
            return ics * ((double)s.C / (double)E1);
        }
        case kS3Variant: {  // :550-556
            double a = (double)s.C, b = (double)E1, c = (double)s.I;
            double denom = (b < c ? c : b) - a;
            return a / denom;
        }
    }
    return 0.0;
}

// ------------------------------------------------------ topology delta (S1,S2,S4)
// One candidate swap (x,y) evaluated by a group of G lanes (gl = lane in group) against
// the packed state in shared memory.  Returns this lane's PARTIAL (dcons, dind); the
// caller sums over the group.  No state is modified.
//
//   dcons = (cc'(x,-1) + cc'(y,x)) - (cc(x,-1) + cc(y,x))      Alignment.cpp:492-493,506-507,516
//   dind  = (ic'(a_x,-1) + ic'(a_y,a_x)) - (ic(a_x,-1) + ic(a_y,a_x))   :494-495,509-510,518
// with primes = after exchanging the images AND mask bits of x and y (:496-504).
template <typename T, int G>
__device__ __forceinline__ void topo_delta(const DevGraph& g, const T* __restrict__ s_aln,
                                           const T* __restrict__ s_inv, int x, int y, int gl,
                                           int& dc, int& di) {
    constexpr uint32_t FLAG = Pk<T>::FLAG, MASK = Pk<T>::MASK;
    dc = 0;
    di = 0;
    if (x == y) return;  // doSwap(x,x): every term cancels (conservedCount n1==ignore, :795)
    const uint32_t rx = s_aln[x], ry = s_aln[y];
    const int ax = (int)(rx & MASK), ay = (int)(ry & MASK);
    const int mx = (rx & FLAG) ? 0 : 1, my = (ry & FLAG) ? 0 : 1;
    const bool xr = x < g.n1, yr = y < g.n1;
    const uint32_t* __restrict__ rowx = g.adj2 + (size_t)ax * (size_t)g.W;
    const uint32_t* __restrict__ rowy = g.adj2 + (size_t)ay * (size_t)g.W;

    if (xr) {  // conservedCount(x, ., ., -1) before and after
        const int b = __ldg(g.rp1 + x), e = __ldg(g.rp1 + x + 1);
        for (int k = b + gl; k < e; k += G) {
            const int t = __ldg(g.col1 + k);
            const uint32_t rt = s_aln[t];
            int it = (int)(rt & MASK), mt = (rt & FLAG) ? 0 : 1;
            int it2 = it, mt2 = mt;
            if (t == x) { it2 = ay; mt2 = my; }
            else if (t == y) { it2 = ax; mt2 = mx; }
            const int o = mx & mt & (int)((__ldg(rowx + (it >> 5)) >> (it & 31)) & 1u);
            const int n = my & mt2 & (int)((__ldg(rowy + (it2 >> 5)) >> (it2 & 31)) & 1u);
            dc += n - o;
        }
    }
    if (yr) {  // conservedCount(y, ., ., ignore = x)
        const int b = __ldg(g.rp1 + y), e = __ldg(g.rp1 + y + 1);
        for (int k = b + gl; k < e; k += G) {
            const int t = __ldg(g.col1 + k);
            if (t == x) continue;
            const uint32_t rt = s_aln[t];
            int it = (int)(rt & MASK), mt = (rt & FLAG) ? 0 : 1;
            int it2 = it, mt2 = mt;
            if (t == y) { it2 = ax; mt2 = mx; }
            const int o = my & mt & (int)((__ldg(rowy + (it >> 5)) >> (it & 31)) & 1u);
            const int n = mx & mt2 & (int)((__ldg(rowx + (it2 >> 5)) >> (it2 & 31)) & 1u);
            dc += n - o;
        }
    }
    // "alive" = pre-image is a real, unmasked node of net1 (inducedCount guard, :813,:820)
    const int Lx_old = (xr ? 1 : 0) & mx, Lx_new = (yr ? 1 : 0) & mx;
    const int Ly_old = (yr ? 1 : 0) & my, Ly_new = (xr ? 1 : 0) & my;
    if (Lx_old != Lx_new || Ly_old != Ly_new) {  // otherwise the induced subgraph is unchanged
        {
            const int b = __ldg(g.rp2 + ax), e = __ldg(g.rp2 + ax + 1);
            for (int k = b + gl; k < e; k += G) {
                const int v = __ldg(g.col2 + k);
                const int ao = ((uint32_t)s_inv[v] < (uint32_t)g.n1) ? 1 : 0;
                const int an = (v == ax) ? Lx_new : (v == ay) ? Ly_new : ao;
                di += (Lx_new & an) - (Lx_old & ao);
            }
        }
        {
            const int b = __ldg(g.rp2 + ay), e = __ldg(g.rp2 + ay + 1);
            for (int k = b + gl; k < e; k += G) {
                const int v = __ldg(g.col2 + k);
                if (v == ax) continue;  // ignore = a_x (:495,:510)
                const int ao = ((uint32_t)s_inv[v] < (uint32_t)g.n1) ? 1 : 0;
                const int an = (v == ay) ? Ly_new : ao;
                di += (Ly_new & an) - (Ly_old & ao);
            }
        }
    }
}

template <int G>
__device__ __forceinline__ int group_sum(int v, unsigned gmask) {
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(gmask, v, off, G);
    return v;
}

// ------------------------------------------------------- score deltas (S3) + size
struct ScoreDelta {
    double dbx, dby, dgx, dgy;  // updateBitscore(x), updateBitscore(y), updateGOC(x), updateGOC(y)
    int dsize;
};

// updateBitscore / updateGOC argument pattern of doSwap, Alignment.cpp:512-515:
//   x: old image a_x (mask m_x)  -> new image a_y (mask m_y)
//   y: old image a_y (mask m_y)  -> new image a_x (mask m_x);  skipped for dummies (:749,:771)
template <typename T>
__device__ __forceinline__ ScoreDelta score_delta(const DevGraph& g, const DevTable& S,
                                                  const DevTable& Gt, const T* s_aln, int x,
                                                  int y) {
    constexpr uint32_t FLAG = Pk<T>::FLAG, MASK = Pk<T>::MASK;
    ScoreDelta d;
    d.dbx = d.dby = d.dgx = d.dgy = 0.0;
    d.dsize = 0;
    if (x == y) return d;
    const uint32_t rx = s_aln[x], ry = s_aln[y];
    const int ax = (int)(rx & MASK), ay = (int)(ry & MASK);
    const bool mx = !(rx & FLAG), my = !(ry & FLAG);
    const bool xr = x < g.n1, yr = y < g.n1;
    if (xr) d.dsize += (int)my - (int)mx;
    if (yr) d.dsize += (int)mx - (int)my;
    if (S.kind) {
        if (xr) {
            double o = mx ? table_lookup(S, x, ax) : 0.0;
            double n = my ? table_lookup(S, x, ay) : 0.0;
            d.dbx = n - o;
        }
        if (yr) {
            double o = my ? table_lookup(S, y, ay) : 0.0;
            double n = mx ? table_lookup(S, y, ax) : 0.0;
            d.dby = n - o;
        }
    }
    if (Gt.kind) {
        if (xr) {
            double o = mx ? table_lookup(Gt, x, ax) : 0.0;
            double n = my ? table_lookup(Gt, x, ay) : 0.0;
            d.dgx = n - o;
        }
        if (yr) {
            double o = my ? table_lookup(Gt, y, ay) : 0.0;
            double n = mx ? table_lookup(Gt, y, ax) : 0.0;
            d.dgy = n - o;
        }
    }
    return d;
}

// state after doSwap: counters += delta (:516-519), accumulators (acc + dX) + dY (:512-515, H2)
__device__ __forceinline__ Scalars apply_delta(const Scalars& s, int dc, int di,
                                               const ScoreDelta& d) {
    Scalars n;
    n.C = s.C + dc;
    n.I = s.I + di;
    n.size = s.size + d.dsize;
    n.bits = (s.bits + d.dbx) + d.dby;
    n.goc = (s.goc + d.dgx) + d.dgy;
    return n;
}
// accumulators after the undoing doSwap (H3): the second swap's deltas are exactly the
// negated ones, added in the same x-then-y order
__device__ __forceinline__ void undo_accumulators(const Scalars& after, const ScoreDelta& d,
                                                  double& bits, double& goc) {
    bits = (after.bits + (-d.dbx)) + (-d.dby);
    goc = (after.goc + (-d.dgx)) + (-d.dgy);
}

// accept test, localSearch.cpp:35-47.  obj == -2: unconditional (ona_apply_swaps)
__device__ __forceinline__ bool accept_move(const DevObjectives& ob, int obj, const double* cur,
                                            const double* nw) {
    if (obj == -2) return true;
    bool improved = true;
    if (obj == -1) {
        for (int j = 0; j < ob.M; ++j)
            if (nw[j] < cur[j]) improved = false;
    } else {
        if (nw[obj] <= cur[obj]) improved = false;
    }
    return improved;
}

// swapNormalizedDelta's return value, localSearch.cpp:98-119
__device__ __forceinline__ double normalized_delta(const DevObjectives& ob, const double* cur,
                                                   const double* nw) {
    bool oneNeg = false;
    double sumSq = 0.0;
    for (int j = 0; j < ob.M; ++j) {
        double pct = (nw[j] - cur[j]) / cur[j];
        if (pct < 0) oneNeg = true;
        sumSq += pct * pct;
    }
    return oneNeg ? -1.0 : sqrt(sumSq);
}

// ------------------------------------------------------------- smem staging
template <typename T>
__device__ __forceinline__ void stage_state(const DevPop& pop, int p, int n2, T* s_aln,
                                            T* s_inv) {
    const uint32_t* ga = pop.aln + (size_t)p * pop.n2p;
    const uint32_t* gi = pop.inv + (size_t)p * pop.n2p;
    const int n4 = pop.n2p >> 2;
    const uint4* ga4 = reinterpret_cast<const uint4*>(ga);
    const uint4* gi4 = reinterpret_cast<const uint4*>(gi);
    constexpr uint32_t FLAG = Pk<T>::FLAG;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        uint4 a = ga4[i], b = gi4[i];
        const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s_aln[4 * i + j] = (T)((av[j] & kGMask) | ((av[j] & kGFlag) ? FLAG : 0u));
            s_inv[4 * i + j] = (T)((bv[j] & kGMask) | ((bv[j] & kGFlag) ? FLAG : 0u));
        }
    }
    (void)n2;
}
template <typename T>
__device__ __forceinline__ void unstage_state(const DevPop& pop, int p, const T* s_aln,
                                              const T* s_inv) {
    uint32_t* ga = pop.aln + (size_t)p * pop.n2p;
    uint32_t* gi = pop.inv + (size_t)p * pop.n2p;
    constexpr uint32_t FLAG = Pk<T>::FLAG, MASK = Pk<T>::MASK;
    for (int i = threadIdx.x; i < pop.n2p; i += blockDim.x) {
        uint32_t a = s_aln[i], b = s_inv[i];
        ga[i] = (a & MASK) | ((a & FLAG) ? kGFlag : 0u);
        gi[i] = (b & MASK) | ((b & FLAG) ? kGFlag : 0u);
    }
}

}  // namespace ona
