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
    // quad-padded CSR: every row padded with the dummy id n2 to a multiple of 4 columns so that
    // one 128-bit load fetches four neighbours; rq[u]..rq[u+1] index int4 quads
    const int* rq1;
    const int4* cq1;
    const int* rq2;
    const int4* cq2;   // = cq1 + (number of net1 quads): both networks' rows live in one array
    // neighbourhood signature of every net2 node: three 32-bit Bloom words -- for every
    // neighbour j bit (j & 31) of .x, bit ((j >> 5) & 31) of .y and bit ((j >> 10) & 31) of .z
    // are set.  A clear bit proves "not adjacent" without touching the bit matrix, and one
    // 16-byte signature serves all the tests against that row.
    const uint4* sig2;
    const uint4* sig1;  // same for net1 (hill-climb conflict tests)
    const uint32_t* adj1;  // net1 adjacency bit matrix [n1][W1] (hill-climb conflict tests)
    int W1;
    // Duplicate edge lines ("a b" + "b a", repeated lines): the reference keeps every copy in
    // adjList (Network.cpp:89-92), so its incremental counters weigh such an edge by its
    // multiplicity while `edges`, adjMatrix and the set-based ics() do not.  The CSR / quad CSR
    // above keep the copies; the unique edge lists carry the multiplicities (null = all 1).
    const int* ew1;
    const int* ew2;
    int dup;  // any duplicate in either network
    // fast swap-delta walk (k_swap_deltas_sorted<T, true>): one word per node = quad base << 8 |
    // has-self-loop << 7 | min(quads, 127) (127 = look the count up in rq)
    const uint32_t* pk1;
    const uint32_t* pk2;
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
    // hash form only: per-row Bloom signature (3 x 32 bits) of the columns that hold a non-zero.
    // Score tables are very sparse (0.1 %), so almost every lookup is answered "0.0" by the row
    // signature -- one 16-byte load per row serves both lookups of that row -- without probing.
    const uint4* rowsig;
};
constexpr unsigned long long kEmptyKey = ~0ull;

struct DevObjectives {
    int M;
    int names[kMaxObj];
};

struct DevPop {
    uint32_t* aln;  // [P][n2p] packed
    uint32_t* inv;  // [P][n2p] packed
    int n2p;        // row stride: n2 + 1 (slot n2 = the "masked off" dummy that pads CSR quads)
                    // rounded up to a multiple of 4
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

// Neighbourhood signatures (DevGraph::sig1 / sig2): bit (id & 31) of .x, bit ((id >> 5) & 31) of
// .y, bit ((id >> 10) & 31) of .z and -- ONA_BLOOM4 -- bit (id * 0x9E3779B1 >> 27) of .w.  For
// graphs of a few thousand nodes the third slice has only a handful of distinct values, so the
// fourth, multiplicative word is what brings the false-positive rate of a degree-10 row from ~2 %
// to ~0.5 %: every false positive is a predicated L2 probe that the whole warp waits for.
#ifndef ONA_BLOOM4
#define ONA_BLOOM4 1
#endif
__host__ __device__ __forceinline__ unsigned sig_slot4(unsigned id) { return (id * 0x9E3779B1u) >> 27; }
__device__ __forceinline__ unsigned nbr_sig_bit(const uint4& sg, unsigned id) {
    unsigned r = (sg.x >> (id & 31u)) & (sg.y >> ((id >> 5) & 31u)) & (sg.z >> ((id >> 10) & 31u));
#if ONA_BLOOM4
    r &= sg.w >> sig_slot4(id);
#endif
    return r & 1u;
}

__device__ __forceinline__ bool sig_has(const uint4& sg, int id) { return nbr_sig_bit(sg, (unsigned)id) != 0u; }
// the two lookups S[r][c0], S[r][c1] of one row, each gated by a flag
__device__ __forceinline__ void row_lookup2(const DevTable& t, int r, bool want0, int c0, bool want1, int c1,
                                            double& v0, double& v1) {
    v0 = 0.0;
    v1 = 0.0;
    if (t.kind == 2) {
        if (!(want0 | want1)) return;
        const uint4 sg = __ldg(t.rowsig + r);
        want0 = want0 && sig_has(sg, c0);
        want1 = want1 && sig_has(sg, c1);
    }
    if (want0) v0 = table_lookup(t, r, c0);
    if (want1) v1 = table_lookup(t, r, c1);
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
// ---- quad-CSR walkers (the sorted swap-delta kernel) -------------------------------------
// Per-candidate context: images, masks and "alive" flags before/after the exchange.
struct SwapCtx {
    int x, y, ax, ay;
    int mx, my;                          // alnMask[x], alnMask[y] before the exchange
    int Lx_old, Lx_new, Ly_old, Ly_new;  // is a_x / a_y the image of a real, unmasked node
    bool xr, yr;
};
template <typename T>
__device__ __forceinline__ SwapCtx make_ctx(const DevGraph& g, const T* s_aln, int x, int y) {
    constexpr uint32_t FLAG = Pk<T>::FLAG, MASK = Pk<T>::MASK;
    SwapCtx c;
    const uint32_t rx = s_aln[x], ry = s_aln[y];
    c.x = x; c.y = y;
    c.ax = (int)(rx & MASK); c.ay = (int)(ry & MASK);
    c.mx = (rx & FLAG) ? 0 : 1; c.my = (ry & FLAG) ? 0 : 1;
    c.xr = x < g.n1; c.yr = y < g.n1;
    c.Lx_old = (c.xr ? 1 : 0) & c.mx; c.Lx_new = (c.yr ? 1 : 0) & c.mx;
    c.Ly_old = (c.yr ? 1 : 0) & c.my; c.Ly_new = (c.xr ? 1 : 0) & c.my;
    return c;
}
__device__ __forceinline__ bool induced_changes(const SwapCtx& c) {
    return c.Lx_old != c.Lx_new || c.Ly_old != c.Ly_new;
}

// Scoring the four neighbours of one quad, whichever list it is from.
//
// Conserved lists (N1(x), N1(y)), Alignment.cpp:799-803 before and after the exchange: for a
// neighbour t with image it = aln[t] (and mask bit mt) the x-list term is
//     mt * ( my*A2[a_y][it] - mx*A2[a_x][it] )  =:  mt * D(it)
// and the y-list term is exactly -mt * D(it) (old and new rows trade places), so one item
// routine serves both lists and only a per-quad sign differs.  A2[row][it] is answered by the
// row's neighbourhood signature first (three 32-bit Bloom words; signature AND mask are folded
// into sgx/sgy once per candidate): only when all three bits are set is the exact word of
// the L2-resident bit matrix probed, with a predicated load -- the common "not an edge" case
// costs no memory traffic and no branch.
// Induced lists (N2(a_x), N2(a_y)), Alignment.cpp:818-822: a neighbour contributes
// coef = L_new - L_old of the list's owner when its pre-image is alive (inv[v] < n1).
// Neighbours that are themselves exchanged (x or y; a_x or a_y) are skipped here and only
// recorded; their closed-form terms are added once per candidate (special_terms).
struct QuadCtx {
    const void* table;   // s_aln (conserved lists) or s_inv (induced lists)
    bool cons;
    int sp0, sp1;        // the two exchanged ids of this list's id space
};
struct RowSig {
    uint4 sx, sy;        // masked signatures of rows a_x, a_y
    unsigned rax, ray;   // word offsets of the two bit-matrix rows
};

template <typename T>
__device__ __forceinline__ void quad_item(const DevGraph& g, const QuadCtx& qc, const RowSig& rs, int t,
                                          int& cv, int& iv, bool& is0, bool& is1) {
    constexpr uint32_t FLAG = Pk<T>::FLAG, MASK = Pk<T>::MASK;
    // row padding is the dummy id n2, whose packed entries are "masked off" in both tables, so
    // it needs no test here: it is never live and never alive
    const uint32_t r = static_cast<const T*>(qc.table)[t];
    is0 |= t == qc.sp0;
    is1 |= t == qc.sp1;
    const bool plain = (t != qc.sp0) & (t != qc.sp1);
    const unsigned it = r & MASK, h1 = it >> 5;
    const bool live = plain & qc.cons & !(r & FLAG);
    const bool pX = live & (nbr_sig_bit(rs.sx, it) != 0u);
    const bool pY = live & (nbr_sig_bit(rs.sy, it) != 0u);
    unsigned wX = 0u, wY = 0u;
    if (pX) wX = __ldg(g.adj2 + (unsigned)(rs.rax + h1));
    if (pY) wY = __ldg(g.adj2 + (unsigned)(rs.ray + h1));
    cv += (int)((wY >> (it & 31u)) & 1u) - (int)((wX >> (it & 31u)) & 1u);
    iv += (plain & !qc.cons & (r < (uint32_t)g.n1)) ? 1 : 0;
}

// `seen` bits
constexpr int kSeenXinX = 1, kSeenYinY = 2, kSeenAXinAX = 4, kSeenAYinAX = 8, kSeenAYinAY = 16;

// Closed-form terms of the exchanged nodes themselves (rare: self-loops and the a_x - a_y edge):
//   self-loop at x:   new my*A2[ay][ay] - old mx*A2[ax][ax]      (x in N1(x), Alignment.cpp:801)
//   self-loop at y:   new mx*A2[ax][ax] - old my*A2[ay][ay]
//   edge x-y:         cancels (the image pair is only mirrored); y's copy is the `ignore` (:493)
//   self-loop at a_x: Lx_new - Lx_old;  self-loop at a_y: Ly_new - Ly_old        (:820)
//   edge a_x-a_y:     (Lx_new & Ly_new) - (Lx_old & Ly_old); a_y's copy is the `ignore` (:495)
// With duplicate edge lines (g.dup) every term is weighted by how often the list holds the id.
__device__ __forceinline__ int list_multiplicity(const int* rp, const int* col, int row, int id) {
    int m = 0;
    for (int k = __ldg(rp + row); k < __ldg(rp + row + 1); ++k) m += __ldg(col + k) == id ? 1 : 0;
    return m;
}
__device__ __forceinline__ void special_terms(const DevGraph& g, const SwapCtx& c, int seen, int& dc,
                                              int& di) {
    if (seen & (kSeenXinX | kSeenYinY)) {
        const int sx = probe(g, c.ax, c.ax), sy = probe(g, c.ay, c.ay);
        const int wx = g.dup && (seen & kSeenXinX) ? list_multiplicity(g.rp1, g.col1, c.x, c.x) : 1;
        const int wy = g.dup && (seen & kSeenYinY) ? list_multiplicity(g.rp1, g.col1, c.y, c.y) : 1;
        if (seen & kSeenXinX) dc += wx * ((c.my & sy) - (c.mx & sx));
        if (seen & kSeenYinY) dc += wy * ((c.mx & sx) - (c.my & sy));
    }
    if (!g.dup) {
        if (seen & kSeenAXinAX) di += c.Lx_new - c.Lx_old;
        if (seen & kSeenAYinAY) di += c.Ly_new - c.Ly_old;
        if (seen & kSeenAYinAX) di += (c.Lx_new & c.Ly_new) - (c.Lx_old & c.Ly_old);
    } else {
        if (seen & kSeenAXinAX) di += list_multiplicity(g.rp2, g.col2, c.ax, c.ax) * (c.Lx_new - c.Lx_old);
        if (seen & kSeenAYinAY) di += list_multiplicity(g.rp2, g.col2, c.ay, c.ay) * (c.Ly_new - c.Ly_old);
        if (seen & kSeenAYinAX)
            di += list_multiplicity(g.rp2, g.col2, c.ax, c.ay) * ((c.Lx_new & c.Ly_new) - (c.Lx_old & c.Ly_old));
    }
}

// The (up to) four quad lists of one candidate as one index space [0, e3):
// [0,e0) N1(x)  [e0,e1) N1(y)  [e1,e2) N2(a_x)  [e2,e3) N2(a_y); quad q of list l is b[l] + q.
struct QuadPlan {
    int e0, e1, e2, e3;
    int b0, b1, b2, b3;
};
__device__ __forceinline__ QuadPlan make_plan(const DevGraph& g, const SwapCtx& c) {
    QuadPlan p;
    p.e0 = p.e1 = p.e2 = p.e3 = 0;
    p.b0 = p.b1 = p.b2 = p.b3 = 0;
    if (c.x == c.y) return p;  // doSwap(x,x): every term cancels
    int q0 = 0, q1 = 0, q2 = 0, q3 = 0;
    if (c.xr) { p.b0 = __ldg(g.rq1 + c.x); q0 = __ldg(g.rq1 + c.x + 1) - p.b0; }
    if (c.yr) { p.b1 = __ldg(g.rq1 + c.y); q1 = __ldg(g.rq1 + c.y + 1) - p.b1; }
    if (induced_changes(c)) {  // otherwise the induced subgraph is unchanged
        p.b2 = __ldg(g.rq2 + c.ax); q2 = __ldg(g.rq2 + c.ax + 1) - p.b2;
        p.b3 = __ldg(g.rq2 + c.ay); q3 = __ldg(g.rq2 + c.ay + 1) - p.b3;
    }
    p.e0 = q0; p.e1 = q0 + q1; p.e2 = p.e1 + q2; p.e3 = p.e2 + q3;
    p.b1 -= p.e0; p.b2 -= p.e1; p.b3 -= p.e2;
    return p;
}

// (dcons, dind) of one candidate: quads [first, first+stride, ...) of its plan.  One loop and
// one branch-free item routine for all four lists, so that the lanes of a warp (each on its
// own candidate, or all on one hub) stay converged whatever the individual list lengths are.
// `seen` must be OR-reduced over the lanes that share a candidate before special_terms.
// Prefetch: request the quad of the next iteration before the current one is scored (pays when a
// group of lanes shares one candidate, as in the hill climb; thread-per-candidate walks already
// overlap their round trips across the lanes' candidates).
template <typename T, bool Prefetch = false>
__device__ __forceinline__ void walk_plan(const DevGraph& g, const T* s_aln, const T* s_inv,
                                          const SwapCtx& c, const QuadPlan& p, int first, int stride,
                                          int& dc, int& di, int& seen) {
    dc = 0;
    di = 0;
    seen = 0;
    RowSig rs;
    rs.rax = (unsigned)c.ax * (unsigned)g.W;
    rs.ray = (unsigned)c.ay * (unsigned)g.W;
    rs.sx = make_uint4(0u, 0u, 0u, 0u);
    rs.sy = rs.sx;
    if (p.e1 > 0) {  // some conserved list will be walked; a masked-off row never matches
        if (c.mx) rs.sx = __ldg(g.sig2 + c.ax);
        if (c.my) rs.sy = __ldg(g.sig2 + c.ay);
    }
    const int coefx = c.Lx_new - c.Lx_old, coefy = c.Ly_new - c.Ly_old;
    int l_next = 0;
    int4 t_next = make_int4(0, 0, 0, 0);
    if (Prefetch && first < p.e3) {
        l_next = (first >= p.e0 ? 1 : 0) + (first >= p.e1 ? 1 : 0) + (first >= p.e2 ? 1 : 0);
        t_next = __ldg((l_next < 2 ? g.cq1 : g.cq2) +
                       (l_next == 0 ? p.b0 : l_next == 1 ? p.b1 : l_next == 2 ? p.b2 : p.b3) + first);
    }
    for (int q = first; q < p.e3; q += stride) {
        int l;
        int4 t;
        if (!Prefetch) {
            l = (q >= p.e0 ? 1 : 0) + (q >= p.e1 ? 1 : 0) + (q >= p.e2 ? 1 : 0);
            const int idx = (l == 0 ? p.b0 : l == 1 ? p.b1 : l == 2 ? p.b2 : p.b3) + q;
            t = __ldg((l < 2 ? g.cq1 : g.cq2) + idx);
        } else {
            l = l_next;
            t = t_next;
            const int qn = q + stride;
            if (qn < p.e3) {
                l_next = (qn >= p.e0 ? 1 : 0) + (qn >= p.e1 ? 1 : 0) + (qn >= p.e2 ? 1 : 0);
                t_next = __ldg((l_next < 2 ? g.cq1 : g.cq2) +
                               (l_next == 0 ? p.b0 : l_next == 1 ? p.b1 : l_next == 2 ? p.b2 : p.b3) + qn);
            }
        }
        QuadCtx qc;
        qc.cons = l < 2;
        qc.table = qc.cons ? static_cast<const void*>(s_aln) : static_cast<const void*>(s_inv);
        qc.sp0 = qc.cons ? c.x : c.ax;
        qc.sp1 = qc.cons ? c.y : c.ay;
        int cv = 0, iv = 0;
        bool is0 = false, is1 = false;
        quad_item<T>(g, qc, rs, t.x, cv, iv, is0, is1);
        quad_item<T>(g, qc, rs, t.y, cv, iv, is0, is1);
        quad_item<T>(g, qc, rs, t.z, cv, iv, is0, is1);
        quad_item<T>(g, qc, rs, t.w, cv, iv, is0, is1);
        const bool odd = (l & 1) != 0;  // lists owned by y / a_y
        dc += odd ? -cv : cv;
        di += iv * (odd ? coefy : coefx);
        // what meeting sp0 / sp1 in this list means (0 = contributes nothing)
        if (is0) seen |= l == 0 ? kSeenXinX : l == 2 ? kSeenAXinAX : 0;
        if (is1) seen |= l == 1 ? kSeenYinY : l == 2 ? kSeenAYinAX : l == 3 ? kSeenAYinAY : 0;
    }
}

// ------------------------------------------------ fast walk on precomputed node counts
// For K candidates against ONE state most of the delta is a function of single nodes:
//   cc(u) = sum over adjList1[u] entries t with mask bit on of A2[aln[u]][aln[t]]   (u < n1)
//   ic(v) = number of adjList2[v] entries whose pre-image is a real, unmasked node  (v < n2)
// (k_node_counts, once per individual per call).  With them the "old" halves of doSwap's
// recounts (Alignment.cpp:492-495) and BOTH induced recounts are O(1), and what is left to walk is
//   I(x, a_y) = sum over t in adjList1[x] \ {x,y} of m_t A2[a_y][aln[t]]      (if x real and m_y)
//   I(y, a_x) = sum over t in adjList1[y] \ {x,y} of m_t A2[a_x][aln[t]]      (if y real and m_x)
// each the size of an intersection of a net1 list with the pull-back of a net2 list, so it is
// counted from whichever list is SHORTER: over adjList1[u] (gather aln, test row a of A2) or over
// adjList2[a] (gather alnInv, test row u of A1).  PPI degrees are heavy-tailed, the minimum of two
// independent degrees is not: 7.4 items per candidate instead of 21 on the yeast-fly pair, one
// adjacency test per item instead of two, and hardly any hubs.  With duplicate edge lines (g.dup)
// the two lists weigh an edge differently, so the net1 list is always the one walked.
struct FastPlan {
    int e0, e1;   // quads of walk 0 = [0,e0), of walk 1 = [e0,e1)
    int b0, b1;   // cq index of quad q is b0 + q (walk 0) / b1 + q (walk 1; b1 is pre-shifted by -e0)
    int side0, side1;  // 0: the net1 list is walked, 1: the net2 list
};
// the packed node words of the four nodes a candidate touches (0 for a dummy position)
struct NodeWords {
    uint32_t wx, wy, wax, way;
};
__device__ __forceinline__ NodeWords load_node_words(const DevGraph& g, const SwapCtx& c) {
    NodeWords w;
    w.wx = c.xr ? __ldg(g.pk1 + c.x) : 0u;
    w.wy = c.yr ? __ldg(g.pk1 + c.y) : 0u;
    w.wax = __ldg(g.pk2 + c.ax);
    w.way = __ldg(g.pk2 + c.ay);
    return w;
}
__device__ __forceinline__ void quad_range(uint32_t w, const int* rq, int node, int& base, int& cnt) {
    base = (int)(w >> 8);     // index into the combined quad array (net1's rows, then net2's)
    cnt = (int)(w & 127u);
    if (cnt == 127) cnt = __ldg(rq + node + 1) - __ldg(rq + node);
}
__device__ __forceinline__ int word_self(uint32_t w) { return (int)((w >> 7) & 1u); }
__device__ __forceinline__ FastPlan make_fast_plan(const DevGraph& g, const SwapCtx& c, const NodeWords& w) {
    FastPlan p;
    p.e0 = p.e1 = p.b0 = p.b1 = p.side0 = p.side1 = 0;
    if (c.x == c.y) return p;
    if (c.xr & (c.my != 0)) {   // I(x, a_y)
        int b1l, q1l, b2l, q2l;
        quad_range(w.wx, g.rq1, c.x, b1l, q1l);
        quad_range(w.way, g.rq2, c.ay, b2l, q2l);
        p.side0 = (g.dup == 0 && q2l < q1l) ? 1 : 0;
        p.b0 = p.side0 ? b2l : b1l;
        p.e0 = p.side0 ? q2l : q1l;
    }
    int len1 = 0;
    if (c.yr & (c.mx != 0)) {   // I(y, a_x)
        int b1l, q1l, b2l, q2l;
        quad_range(w.wy, g.rq1, c.y, b1l, q1l);
        quad_range(w.wax, g.rq2, c.ax, b2l, q2l);
        p.side1 = (g.dup == 0 && q2l < q1l) ? 1 : 0;
        p.b1 = p.side1 ? b2l : b1l;
        len1 = p.side1 ? q2l : q1l;
    }
    p.e1 = p.e0 + len1;
    p.b1 -= p.e0;
    return p;
}

// one walk in 32 bits: quad base (24) | quads (7) | side (1); quads beyond 127 only occur for
// hub candidates (> kHubQuads quads), whose plan is rebuilt by the warp that walks them
__device__ __forceinline__ uint2 pack_fast_plan(const FastPlan& p) {
    const int len1 = p.e1 - p.e0, b1 = p.b1 + p.e0;
    return make_uint2(((uint32_t)p.b0 << 8) | ((uint32_t)min(p.e0, 127) << 1) | (uint32_t)p.side0,
                      ((uint32_t)b1 << 8) | ((uint32_t)min(len1, 127) << 1) | (uint32_t)p.side1);
}
__device__ __forceinline__ FastPlan unpack_fast_plan(const uint2& w) {
    FastPlan p;
    p.b0 = (int)(w.x >> 8);
    p.e0 = (int)((w.x >> 1) & 127u);
    p.side0 = (int)(w.x & 1u);
    const int len1 = (int)((w.y >> 1) & 127u);
    p.side1 = (int)(w.y & 1u);
    p.e1 = p.e0 + len1;
    p.b1 = (int)(w.y >> 8) - p.e0;
    return p;
}

template <typename T>
__device__ __forceinline__ int fast_item(const T* table, uint32_t lim, const uint4& sg, const uint32_t* row,
                                         int sp0, int sp1, int t) {
    constexpr uint32_t MASK = Pk<T>::MASK;
    // side 0: r = packed aln[t], live = mask bit on (r < FLAG), id = image;  side 1: r = packed
    // alnInv[t], live = pre-image real and unmasked (r < n1), id = pre-image.  Row padding (id n2)
    // is masked off in both tables.
    const uint32_t r = table[t];
    const bool live = (t != sp0) & (t != sp1) & (r < lim);
    const unsigned it = r & MASK, s0 = it & 31u, h1 = it >> 5;
    const bool hit = live & (nbr_sig_bit(sg, it) != 0u);
    unsigned w = 0u;
    if (hit) w = __ldg(row + h1);
    return (int)((w >> s0) & 1u);
}

// I(x, a_y) and I(y, a_x) of one candidate: quads [first, first + stride, ...) of its plan
template <typename T>
__device__ __forceinline__ void walk_fast(const DevGraph& g, const T* s_aln, const T* s_inv, const SwapCtx& c,
                                          const FastPlan& p, int first, int stride, int& ix, int& iy) {
    constexpr uint32_t FLAG = Pk<T>::FLAG;
    ix = 0;
    iy = 0;
    if (first >= p.e1) return;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    uint4 sg0 = zero4, sg1 = zero4;
    const uint32_t *row0 = g.adj2, *row1 = g.adj2;
    if (p.e0 > 0) {
        sg0 = p.side0 ? __ldg(g.sig1 + c.x) : __ldg(g.sig2 + c.ay);
        row0 = p.side0 ? g.adj1 + (size_t)c.x * (size_t)g.W1 : g.adj2 + (size_t)c.ay * (size_t)g.W;
    }
    if (p.e1 > p.e0) {
        sg1 = p.side1 ? __ldg(g.sig1 + c.y) : __ldg(g.sig2 + c.ax);
        row1 = p.side1 ? g.adj1 + (size_t)c.y * (size_t)g.W1 : g.adj2 + (size_t)c.ax * (size_t)g.W;
    }
// Two quads per trip, both requested before either is scored, halve the dependent L2 round trips of a
// walk at the price of scoring a padding quad whenever a walk has an odd number of them.  It paid while
// the kernel had 60 KB of L1 (0.4106 vs 0.4075 ms); with 92 KB one quad per trip is faster (0.3154 vs
// 0.3230 ms), so it is off.
#ifndef ONA_FAST_UNROLL2
#define ONA_FAST_UNROLL2 0
#endif
    auto quad = [&](const int4& t, bool w1) {
        const int side = w1 ? p.side1 : p.side0;
        const T* table = side ? s_inv : s_aln;
        const uint32_t lim = side ? (uint32_t)g.n1 : FLAG;
        const int sp0 = side ? c.ax : c.x, sp1 = side ? c.ay : c.y;
        const uint4 sg = w1 ? sg1 : sg0;
        const uint32_t* row = w1 ? row1 : row0;
        const int n = fast_item<T>(table, lim, sg, row, sp0, sp1, t.x) + fast_item<T>(table, lim, sg, row, sp0, sp1, t.y) +
                      fast_item<T>(table, lim, sg, row, sp0, sp1, t.z) + fast_item<T>(table, lim, sg, row, sp0, sp1, t.w);
        if (w1) iy += n;
        else ix += n;
    };
#if ONA_FAST_UNROLL2
    // two quads per trip, both requested before either is scored: the lists are short (1-3 quads),
    // so this halves the number of dependent L2 round trips of a candidate.  A missing second quad
    // is the padding quad (id n2: masked off in both tables, contributes nothing).
    for (int q = first; q < p.e1; q += 2 * stride) {
        const int q2 = q + stride;
        const bool wa = q >= p.e0, wb = q2 >= p.e0;
        const int4 ta = __ldg(g.cq1 + (wa ? p.b1 : p.b0) + q);
        int4 tb = make_int4(g.n2, g.n2, g.n2, g.n2);
        if (q2 < p.e1) tb = __ldg(g.cq1 + (wb ? p.b1 : p.b0) + q2);
        quad(ta, wa);
        quad(tb, wb);
    }
#else
    for (int q = first; q < p.e1; q += stride) {
        const bool w1 = q >= p.e0;
        quad(__ldg(g.cq1 + (w1 ? p.b1 : p.b0) + q), w1);
    }
#endif
}

// The closed forms around the two walked sums (see the derivation in DESIGN.md, section 4.1):
// conservedCount before / after the exchange for x and y (Alignment.cpp:492-493, 506-507) and
// inducedCount before / after for a_x and a_y (:494-495, 508-509), with the node counts standing
// in for every sum that does not involve the partner node.
//   nc: this individual's node counts, cc at [0, n1), ic at [ic_off, ic_off + n2)
__device__ __forceinline__ void fast_delta(const DevGraph& g, const SwapCtx& c, const NodeWords& nw,
                                           const unsigned short* nc, int ic_off, int ix, int iy, int& dc, int& di) {
    dc = 0;
    di = 0;
    if (c.x == c.y) return;
    // multiplicities of the entries that involve the exchanged nodes themselves
    int sxx = 0, syy = 0, mxy = 0;
    if (c.xr) sxx = word_self(nw.wx);
    if (c.yr) syy = word_self(nw.wy);
    if (c.xr & c.yr) mxy = (int)((__ldg(g.adj1 + (size_t)c.x * (size_t)g.W1 + (c.y >> 5)) >> (c.y & 31)) & 1u);
    if (g.dup) {
        if (sxx) sxx = list_multiplicity(g.rp1, g.col1, c.x, c.x);
        if (syy) syy = list_multiplicity(g.rp1, g.col1, c.y, c.y);
        if (mxy) mxy = list_multiplicity(g.rp1, g.col1, c.x, c.y);
    }
    const bool ich = induced_changes(c);
    int a_xx = 0, a_yy = 0, a_xy = 0;
    if (sxx | syy) { a_xx = probe(g, c.ax, c.ax); a_yy = probe(g, c.ay, c.ay); }
    if (mxy != 0 || ich) a_xy = probe(g, c.ax, c.ay);
    const int ccx = c.xr ? (int)__ldg(nc + c.x) : 0, ccy = c.yr ? (int)__ldg(nc + c.y) : 0;
    const int xr = c.xr ? 1 : 0, yr = c.yr ? 1 : 0;
    const int cX = c.mx * xr * ccx;
    const int cY = c.my * yr * (ccy - mxy * c.mx * a_xy);
    const int cXn = c.my * xr * (ix + sxx * c.my * a_yy + mxy * c.mx * a_xy);
    const int cYn = c.mx * yr * (iy + syy * c.mx * a_xx);
    dc = cXn + cYn - cX - cY;
    if (ich) {
        int s2x = word_self(nw.wax), s2y = word_self(nw.way), m2 = a_xy;
        if (g.dup) {
            if (s2x) s2x = list_multiplicity(g.rp2, g.col2, c.ax, c.ax);
            if (s2y) s2y = list_multiplicity(g.rp2, g.col2, c.ay, c.ay);
            if (m2) m2 = list_multiplicity(g.rp2, g.col2, c.ax, c.ay);
        }
        const int icx = (int)__ldg(nc + ic_off + c.ax), icy = (int)__ldg(nc + ic_off + c.ay);
        const int iX = c.Lx_old * icx;
        const int iY = c.Ly_old * (icy - m2 * c.Lx_old);
        const int iXn = c.Lx_new * (icx - s2x * c.Lx_old - m2 * c.Ly_old + s2x * c.Lx_new + m2 * c.Ly_new);
        const int iYn = c.Ly_new * (icy - s2y * c.Ly_old - m2 * c.Lx_old + s2y * c.Ly_new);
        di = iXn + iYn - iX - iY;
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
        double o, n;
        if (xr) { row_lookup2(S, x, mx, ax, my, ay, o, n); d.dbx = n - o; }
        if (yr) { row_lookup2(S, y, my, ay, mx, ax, o, n); d.dby = n - o; }
    }
    if (Gt.kind) {
        double o, n;
        if (xr) { row_lookup2(Gt, x, mx, ax, my, ay, o, n); d.dgx = n - o; }
        if (yr) { row_lookup2(Gt, y, my, ay, mx, ax, o, n); d.dgy = n - o; }
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

// computeFitness on the swapped state + the accept test of correctHillClimb
// (localSearch.cpp:35-47) + swapNormalizedDelta's return value (:98-119), one objective at a
// time so that no objective vector has to live in registers or local memory.
//   obj == -1: accept iff no objective decreased; obj = j: iff objective j strictly increased;
//   obj == -2: unconditional (ona_apply_swaps).
//   curfit: the current state's objectives (any address space), or nullptr to derive them
//           from `cur`;  fit_out (nullable): receives the new objectives.
//   names: the objective list in shared memory (nullptr: read it from the kernel parameters; the
//          indexed constant-bank load is slow on a single warp's critical path)
template <bool STREAM = false>
__device__ __forceinline__ bool judge_move(const DevObjectives& ob, int obj, const Scalars& cur,
                                           const double* curfit, const Scalars& nw,
                                           unsigned long long E1, double* fit_out, bool want_snd,
                                           double& snd, const int* names = nullptr) {
    bool any_lt = false, obj_le = false, oneNeg = false;
    double sumSq = 0.0;
    for (int j = 0; j < ob.M; ++j) {
        const int name = names ? names[j] : ob.names[j];
        const double fn = objective(name, nw, E1);
        const double fc = curfit ? curfit[j] : objective(name, cur, E1);
        if (fit_out) {
            if (STREAM) __stcs(fit_out + j, fn);   // global output stream, written once: keep it out of the L1's way
            else fit_out[j] = fn;
        }
        if (fn < fc) any_lt = true;
        if (j == obj && fn <= fc) obj_le = true;
        if (want_snd) {
            const double pct = (fn - fc) / fc;
            if (pct < 0) oneNeg = true;
            sumSq += pct * pct;
        }
    }
    snd = want_snd ? (oneNeg ? -1.0 : sqrt(sumSq)) : 0.0;
    if (obj == -2) return true;
    return obj == -1 ? !any_lt : !obj_le;
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

// One individual's 16-bit state mirror (aln then alnInv, contiguous) into shared memory with ONE bulk
// copy of the TMA engine (cp.async.bulk, 1-D: no tensor map) completing on an mbarrier: no thread
// issues loads or stores for the staging, and the CTA's prologue runs beside it.
__device__ __forceinline__ void bulk_stage_begin(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b)
            : "memory");
    }
}
// every thread that reads the staged bytes waits for phase 0 of the barrier (after a __syncthreads
// that makes thread 0's mbarrier.init visible)
__device__ __forceinline__ void bulk_stage_wait(uint64_t* bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(b)
            : "memory");
    }
}

// Where one individual's packed state lives while a CTA works on it:
//   uint16_t : staged into shared memory (n2 < 32768; 4*n2 bytes), written back on release
//   uint32_t : used in place in HBM/L2 (any n2); same flag layout as the global arrays
template <typename T> struct State;
template <> struct State<uint16_t> {
    __device__ static __forceinline__ void acquire(const DevPop& pop, int p, int n2,
                                                   unsigned char* smem, uint16_t*& a, uint16_t*& i) {
        a = reinterpret_cast<uint16_t*>(smem);
        i = a + pop.n2p;
        stage_state<uint16_t>(pop, p, n2, a, i);
    }
    __device__ static __forceinline__ void release(const DevPop& pop, int p, const uint16_t* a,
                                                   const uint16_t* i) {
        unstage_state<uint16_t>(pop, p, a, i);
    }
};
template <> struct State<uint32_t> {
    __device__ static __forceinline__ void acquire(const DevPop& pop, int p, int, unsigned char*,
                                                   uint32_t*& a, uint32_t*& i) {
        a = pop.aln + (size_t)p * pop.n2p;
        i = pop.inv + (size_t)p * pop.n2p;
    }
    __device__ static __forceinline__ void release(const DevPop&, int, const uint32_t*, const uint32_t*) {}
};

}  // namespace ona
