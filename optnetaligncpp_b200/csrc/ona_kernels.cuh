// Kernels of the OptNetAlign scoring engine (sm_100a).  See ona_device.cuh for the
// per-candidate math and DESIGN.md for the roofline each kernel is bounded by.
#pragma once
#include "ona_device.cuh"

namespace ona {

// --------------------------------------------------------------------------------
// Upload: pack (aln, mask) -> packed aln, build packed inverse, verify permutation.
// Replaces the alnInv rebuild loops of Alignment.cpp:31-34, 383-385.
// --------------------------------------------------------------------------------
__global__ void k_pack_state(DevPop pop, int first, int count, int n2, const int* __restrict__ aln,
                             const unsigned char* __restrict__ mask) {
    const int p = blockIdx.y;
    if (p >= count) return;
    uint32_t* ga = pop.aln + (size_t)(first + p) * pop.n2p;
    uint32_t* gi = pop.inv + (size_t)(first + p) * pop.n2p;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pop.n2p; i += gridDim.x * blockDim.x) {
        if (i < n2) {
            const uint32_t v = (uint32_t)aln[(size_t)p * n2 + i];
            const uint32_t f = (mask && !mask[(size_t)p * n2 + i]) ? kGFlag : 0u;
            ga[i] = v | f;
            gi[v] = (uint32_t)i | f;
        } else {
            ga[i] = 0;
            gi[i] = 0;
        }
    }
}
__global__ void k_check_state(DevPop pop, int first, int count, int n2, int* __restrict__ bad) {
    const int p = blockIdx.y;
    if (p >= count) return;
    const uint32_t* ga = pop.aln + (size_t)(first + p) * pop.n2p;
    const uint32_t* gi = pop.inv + (size_t)(first + p) * pop.n2p;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n2; v += gridDim.x * blockDim.x) {
        const uint32_t i = gi[v] & kGMask;
        if (i >= (uint32_t)n2 || (ga[i] & kGMask) != (uint32_t)v) atomicAdd(bad, 1);
    }
}
__global__ void k_unpack_state(DevPop pop, int first, int count, int n2, int* __restrict__ aln,
                               int* __restrict__ inv, unsigned char* __restrict__ mask) {
    const int p = blockIdx.y;
    if (p >= count) return;
    const uint32_t* ga = pop.aln + (size_t)(first + p) * pop.n2p;
    const uint32_t* gi = pop.inv + (size_t)(first + p) * pop.n2p;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        if (aln) aln[(size_t)p * n2 + i] = (int)(ga[i] & kGMask);
        if (inv) inv[(size_t)p * n2 + i] = (int)(gi[i] & kGMask);
        if (mask) mask[(size_t)p * n2 + i] = (ga[i] & kGFlag) ? 0 : 1;
    }
}

// --------------------------------------------------------------------------------
// Block reductions
// --------------------------------------------------------------------------------
template <typename V>
__device__ __forceinline__ V warp_sum(V v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
// sums v over the block; result valid in thread 0.  scratch: >= 32 elements of V.
template <typename V>
__device__ __forceinline__ V block_sum(V v, V* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    V r = V(0);
    if (w == 0) {
        r = lane < nw ? scratch[lane] : V(0);
        r = warp_sum(r);
    }
    return r;
}

// --------------------------------------------------------------------------------
// F1-F4: population-batched from-scratch scoring, one CTA per individual.
//   conserved  = |{(u,v) in E1 : mask[u], mask[v], (aln[u],aln[v]) in E2}|   ics() numerator :629-658
//   induced    = |{(u,v) in E2 : both pre-images real and unmasked}|         fastICSDenominator :667-690
//   bits / goc = sum_{i<n1, mask[i]} S[i][aln[i]]                            sumBLAST/sumGOC :701-722
//   size       = #{i<n1 : mask[i]}                                           alnSize :724-732
//   refseed    = (int)((C/I)*I)                                              :444-446  (hazard H1)
// The individual's packed aln/inv are staged in shared memory once and gathered from there.
// --------------------------------------------------------------------------------
struct FullOut {
    int* cons_true;
    int* cons_ref;
    int* ind;
    double* bits;
    double* goc;
    double* size;
    double* fit;
};

template <typename T>
__global__ void __launch_bounds__(512)
k_full_fitness(DevGraph g, DevTable S, DevTable Gt, DevObjectives ob, DevPop pop, int first,
               int seed_mode, int sequential, double* __restrict__ seq_scratch, FullOut out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_aln = reinterpret_cast<T*>(smem_raw);
    T* s_inv = s_aln + pop.n2p;
    __shared__ double red_d[32];
    __shared__ int red_i[32];
    constexpr uint32_t FLAG = Pk<T>::FLAG, MASK = Pk<T>::MASK;
    const int pl = blockIdx.x, p = first + pl;
    stage_state<T>(pop, p, g.n2, s_aln, s_inv);
    __syncthreads();

    int c = 0, ind = 0, sz = 0;
    for (int e = threadIdx.x; e < g.ne1; e += blockDim.x) {
        const uint32_t ru = s_aln[__ldg(g.eu1 + e)], rv = s_aln[__ldg(g.ev1 + e)];
        if (!((ru | rv) & FLAG)) c += probe(g, (int)(ru & MASK), (int)(rv & MASK));
    }
    for (int e = threadIdx.x; e < g.ne2; e += blockDim.x) {
        const uint32_t iu = s_inv[__ldg(g.eu2 + e)], iv = s_inv[__ldg(g.ev2 + e)];
        ind += (iu < (uint32_t)g.n1 && iv < (uint32_t)g.n1) ? 1 : 0;
    }
    double sb = 0.0, sg = 0.0;
    double* seq_b = seq_scratch ? seq_scratch + (size_t)pl * 2 * g.n1 : nullptr;
    double* seq_g = seq_b ? seq_b + g.n1 : nullptr;
    for (int i = threadIdx.x; i < g.n1; i += blockDim.x) {
        const uint32_t r = s_aln[i];
        const bool m = !(r & FLAG);
        sz += m ? 1 : 0;
        double vb = 0.0, vg = 0.0;
        if (m) {
            if (S.kind) vb = table_lookup(S, i, (int)(r & MASK));
            if (Gt.kind) vg = table_lookup(Gt, i, (int)(r & MASK));
        }
        if (sequential) {
            seq_b[i] = vb;
            seq_g[i] = vg;
        } else {
            sb += vb;
            sg += vg;
        }
    }
    c = block_sum<int>(c, red_i);
    ind = block_sum<int>(ind, red_i);
    sz = block_sum<int>(sz, red_i);
    __shared__ double s_seq[2];
    if (sequential) {
        __syncthreads();
        // the reference's index-order fp64 sums (masked-off terms were skipped there; adding
        // their +0.0 here leaves every partial sum unchanged)
        if (threadIdx.x == 0 || threadIdx.x == 32) {
            const double* src = threadIdx.x == 0 ? seq_b : seq_g;
            double acc = 0.0;
            if ((threadIdx.x == 0 && S.kind) || (threadIdx.x == 32 && Gt.kind))
                for (int i = 0; i < g.n1; ++i) acc += src[i];
            s_seq[threadIdx.x == 0 ? 0 : 1] = acc;
        }
        __syncthreads();
        sb = s_seq[0];
        sg = s_seq[1];
    } else {
        sb = block_sum<double>(sb, red_d);
        sg = block_sum<double>(sg, red_d);
    }
    if (threadIdx.x == 0) {
        // currConservedCount = ics()*denom, truncated (Alignment.cpp:445); ics()==0 when denom==0
        const double denom = (double)ind;
        const double ics = denom == 0.0 ? 0.0 : __ddiv_rn((double)c, denom);
        const int cref = __double2int_rz(__dmul_rn(ics, denom));
        Scalars s;
        s.C = (seed_mode == 2) ? cref : c;
        s.I = ind;
        s.size = sz;
        s.bits = sb;
        s.goc = sg;
        if (seed_mode != 0) {
            pop.cons[p] = s.C;
            pop.ind[p] = s.I;
            pop.size[p] = s.size;
            pop.bits[p] = s.bits;
            pop.goc[p] = s.goc;
        }
        for (int j = 0; j < ob.M; ++j) {
            const double f = objective(ob.names[j], s, g.E1);
            if (seed_mode != 0) pop.fit[(size_t)p * ob.M + j] = f;
            if (out.fit) out.fit[(size_t)pl * ob.M + j] = f;
        }
        if (out.cons_true) out.cons_true[pl] = c;
        if (out.cons_ref) out.cons_ref[pl] = cref;
        if (out.ind) out.ind[pl] = ind;
        if (out.bits) out.bits[pl] = sb;
        if (out.goc) out.goc[pl] = sg;
        if (out.size) out.size[pl] = (double)sz;
    }
}

// recompute the cached objectives of slots from their counters (computeFitness, :559-602)
__global__ void k_refresh_fitness(DevGraph g, DevObjectives ob, DevPop pop, int first, int count) {
    const int pl = blockIdx.x * blockDim.x + threadIdx.x;
    if (pl >= count) return;
    const int p = first + pl;
    Scalars s;
    s.C = pop.cons[p]; s.I = pop.ind[p]; s.size = pop.size[p]; s.bits = pop.bits[p]; s.goc = pop.goc[p];
    for (int j = 0; j < ob.M; ++j) pop.fit[(size_t)p * ob.M + j] = objective(ob.names[j], s, g.E1);
}
// alnSize() of slots (after ona_pop_set_counters the size cache must match the masks)
__global__ void k_recount_size(DevGraph g, DevPop pop, int first, int count) {
    const int pl = blockIdx.x, p = first + pl;
    __shared__ int red_i[32];
    const uint32_t* ga = pop.aln + (size_t)p * pop.n2p;
    int sz = 0;
    for (int i = threadIdx.x; i < g.n1; i += blockDim.x) sz += (ga[i] & kGFlag) ? 0 : 1;
    sz = block_sum<int>(sz, red_i);
    if (threadIdx.x == 0) pop.size[p] = sz;
    (void)count;
}

// --------------------------------------------------------------------------------
// S1-S7: batched, state-preserving swap deltas.
// grid = (chunks, individuals); each CTA stages its individual's packed aln/inv in shared
// memory and scores a contiguous chunk of that individual's K candidates.
//   phase 1 (topology): groups of G lanes walk the <= 4 neighbour lists of one candidate
//                       (lanes stride the CSR rows -> coalesced column loads, gathers hit
//                       shared memory, adjacency tests are single-word probes of the
//                       L2-resident bit matrix), then shuffle-reduce the integer deltas;
//   phase 2 (per lane): one candidate per lane -- <= 8 score-table lookups, the objectives
//                       in the reference's fp64 operation order, accept bit, coalesced stores.
// --------------------------------------------------------------------------------
struct DeltaSummary {
    int* num_improved;    // [count] (atomic)
    double* block_best;   // [count][chunks] best swapNormalizedDelta in the chunk
    long long* block_arg; // [count][chunks] candidate index of the best (lowest index on ties)
};

template <typename T, int G>
__global__ void __launch_bounds__(256)
k_swap_deltas(DevGraph g, DevTable S, DevTable Gt, DevObjectives ob, DevPop pop, int first, int K,
              const int* __restrict__ xy, unsigned long long seed, long long island_offset,
              int obj, DevDeltaOut out, DeltaSummary sum, int* __restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_aln = reinterpret_cast<T*>(smem_raw);
    T* s_inv = s_aln + pop.n2p;
    __shared__ double s_bestv[8];
    __shared__ long long s_besti[8];
    const int pl = blockIdx.y, p = first + pl;
    stage_state<T>(pop, p, g.n2, s_aln, s_inv);

    Scalars cur;
    cur.C = pop.cons[p]; cur.I = pop.ind[p]; cur.size = pop.size[p];
    cur.bits = pop.bits[p]; cur.goc = pop.goc[p];
    double curfit[kMaxObj];
#pragma unroll
    for (int j = 0; j < kMaxObj; ++j) curfit[j] = j < ob.M ? objective(ob.names[j], cur, g.E1) : 0.0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int per = (K + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(K, lo + per);
    const unsigned long long ikey = rng_individual_key(seed, (unsigned long long)(island_offset + p));
    constexpr int NG = 32 / G;  // candidates evaluated concurrently per warp in phase 1
    const int grp = lane / G, gl = lane % G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));

    int n_imp = 0;
    double bestv = -2.0;
    long long besti = -1;

    for (int base = lo + warp * 32; base < hi; base += nwarps * 32) {
        const int c = base + lane;
        const bool valid = c < hi;
        int x = 0, y = 0;
        if (valid) {
            if (xy) {
                const int2 v = __ldg(reinterpret_cast<const int2*>(xy) + (size_t)pl * K + c);
                x = v.x; y = v.y;
                if ((unsigned)x >= (unsigned)g.n2 || (unsigned)y >= (unsigned)g.n2) {
                    atomicOr(err_flag, 1);  // reported by the host entry point as ONA_ERR_INVALID
                    x = 0; y = 0;
                }
            } else {
                rng_proposal(ikey, (unsigned long long)c, g.n1, g.n2, x, y);
            }
        }
        // ---- phase 1
        int my_dc = 0, my_di = 0;
        if (G == 1) {
            topo_delta<T, 1>(g, s_aln, s_inv, x, y, 0, my_dc, my_di);
        } else {
#pragma unroll 1
            for (int sub = 0; sub < G; ++sub) {
                const int ci = sub * NG + grp;
                const int cx = __shfl_sync(0xffffffffu, x, ci), cy = __shfl_sync(0xffffffffu, y, ci);
                int dc, di;
                topo_delta<T, G>(g, s_aln, s_inv, cx, cy, gl, dc, di);
                dc = group_sum<G>(dc, gmask);
                di = group_sum<G>(di, gmask);
                __syncwarp();
                const int rdc = __shfl_sync(0xffffffffu, dc, (lane % NG) * G);
                const int rdi = __shfl_sync(0xffffffffu, di, (lane % NG) * G);
                if (lane / NG == sub) { my_dc = rdc; my_di = rdi; }
            }
        }
        // ---- phase 2
        if (valid) {
            const ScoreDelta sd = score_delta<T>(g, S, Gt, s_aln, x, y);
            const Scalars nw = apply_delta(cur, my_dc, my_di, sd);
            double nfit[kMaxObj];
#pragma unroll
            for (int j = 0; j < kMaxObj; ++j) nfit[j] = j < ob.M ? objective(ob.names[j], nw, g.E1) : 0.0;
            const bool imp = accept_move(ob, obj, curfit, nfit);
            const size_t o = (size_t)pl * K + c;
            if (out.cons) out.cons[o] = nw.C;
            if (out.ind) out.ind[o] = nw.I;
            if (out.bits) out.bits[o] = nw.bits;
            if (out.goc) out.goc[o] = nw.goc;
            if (out.fit)
                for (int j = 0; j < ob.M; ++j) out.fit[o * ob.M + j] = nfit[j];
            if (out.improved) out.improved[o] = imp ? 1 : 0;
            n_imp += imp ? 1 : 0;
            if (out.snd || sum.block_best) {
                const double v = normalized_delta(ob, curfit, nfit);
                if (out.snd) out.snd[o] = v;
                if (v > bestv) { bestv = v; besti = c; }  // strict >, first wins (localSearch.cpp:141)
            }
        }
    }
    if (sum.num_improved) {
        n_imp = warp_sum<int>(n_imp);
        if (lane == 0 && n_imp) atomicAdd(sum.num_improved + pl, n_imp);
    }
    if (sum.block_best) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bestv, off);
            const long long oi = __shfl_xor_sync(0xffffffffu, besti, off);
            if (oi >= 0 && (besti < 0 || ov > bestv || (ov == bestv && oi < besti))) { bestv = ov; besti = oi; }
        }
        if (lane == 0) { s_bestv[warp] = bestv; s_besti[warp] = besti; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < nwarps; ++w) {
                const double ov = s_bestv[w];
                const long long oi = s_besti[w];
                if (oi >= 0 && (besti < 0 || ov > bestv || (ov == bestv && oi < besti))) { bestv = ov; besti = oi; }
            }
            sum.block_best[(size_t)pl * gridDim.x + blockIdx.x] = bestv;
            sum.block_arg[(size_t)pl * gridDim.x + blockIdx.x] = besti;
        }
    }
}

// --------------------------------------------------------------------------------
// S6: correctHillClimb with the reference's exact sequential semantics, one CTA per
// individual, state resident in shared memory for the whole climb.
//
// Speculative rounds: the next R = 32 proposals are all scored against the current state in
// parallel (32 groups of G lanes); warp 0 then resolves them IN ORDER.  Proposals are mostly
// rejected (~95 %), and a rejected proposal leaves aln/alnInv untouched, so every score up to
// and including the first accepted proposal is exactly what the sequential loop would have
// computed.  The first accepted proposal is committed, later ones in the round are discarded
// and re-proposed in the next round.  The fp64 accumulators after a rejected proposal are
// (acc1 - dX) - dY, not acc0 (H3); whenever that differs from acc0 in any bit the round's
// later decisions are re-taken with the perturbed accumulators, so accumulators and decisions
// stay bit-identical to the reference.
// --------------------------------------------------------------------------------
constexpr int kRound = 32;

template <typename T, int G>
__global__ void __launch_bounds__(kRound * G)
k_hill_climb(DevGraph g, DevTable S, DevTable Gt, DevObjectives ob, DevPop pop, int first,
             long long iters, const int* __restrict__ xy, unsigned long long seed,
             long long island_offset, const int* __restrict__ objs, int force_obj,
             unsigned char* __restrict__ accepted, long long* __restrict__ num_accepted,
             int* __restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_aln = reinterpret_cast<T*>(smem_raw);
    T* s_inv = s_aln + pop.n2p;
    __shared__ int s_x[kRound], s_y[kRound], s_dc[kRound], s_di[kRound];
    __shared__ long long s_pos;
    constexpr uint32_t FLAG = Pk<T>::FLAG, MASK = Pk<T>::MASK;

    const int pl = blockIdx.x, p = first + pl;
    stage_state<T>(pop, p, g.n2, s_aln, s_inv);
    if (threadIdx.x == 0) s_pos = 0;
    const int obj = force_obj != -3 ? force_obj : (objs ? objs[pl] : -1);
    const unsigned long long ikey = rng_individual_key(seed, (unsigned long long)(island_offset + p));
    const int grp = threadIdx.x / G, gl = threadIdx.x % G;
    const int lane = threadIdx.x & 31;
    const unsigned gmask = (G >= 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));

    // scalars live in warp 0's registers (replicated over its lanes)
    Scalars cur;
    cur.C = pop.cons[p]; cur.I = pop.ind[p]; cur.size = pop.size[p];
    cur.bits = pop.bits[p]; cur.goc = pop.goc[p];
    long long n_acc = 0;
    __syncthreads();

    for (;;) {
        const long long pos = s_pos;
        if (pos >= iters) break;
        const int nvalid = (int)min((long long)kRound, iters - pos);
        // ---- phase 1: group `grp` scores proposal pos+grp
        {
            int x = 0, y = 0;
            if (grp < nvalid) {
                if (xy) {
                    const int2 v = __ldg(reinterpret_cast<const int2*>(xy) + (size_t)pl * iters + pos + grp);
                    x = v.x; y = v.y;
                    if (x < 0 && force_obj == -2) { x = 0; y = 0; }  // ragged swap lists (ona_apply_swaps)
                    if ((unsigned)x >= (unsigned)g.n2 || (unsigned)y >= (unsigned)g.n2) {
                        atomicOr(err_flag, 1);
                        x = 0; y = 0;
                    }
                } else {
                    rng_proposal(ikey, (unsigned long long)(pos + grp), g.n1, g.n2, x, y);
                }
            }
            int dc, di;
            topo_delta<T, G>(g, s_aln, s_inv, x, y, gl, dc, di);
            if (G <= 32) {
                dc = group_sum<(G <= 32 ? G : 32)>(dc, gmask);
                di = group_sum<(G <= 32 ? G : 32)>(di, gmask);
            }
            if (gl == 0) { s_x[grp] = x; s_y[grp] = y; s_dc[grp] = dc; s_di[grp] = di; }
        }
        __syncthreads();
        // ---- phase 2: warp 0 resolves the round in order
        if (threadIdx.x < 32) {
            const int c = lane;
            const bool valid = c < nvalid;
            const int x = s_x[c], y = s_y[c], dc = s_dc[c], di = s_di[c];
            ScoreDelta sd = score_delta<T>(g, S, Gt, s_aln, valid ? x : 0, valid ? y : 0);
            int start = 0, consumed = nvalid, acc_lane = -1;
            for (;;) {
                double curfit[kMaxObj], nfit[kMaxObj];
                const Scalars nw = apply_delta(cur, dc, di, sd);
#pragma unroll
                for (int j = 0; j < kMaxObj; ++j) {
                    curfit[j] = j < ob.M ? objective(ob.names[j], cur, g.E1) : 0.0;
                    nfit[j] = j < ob.M ? objective(ob.names[j], nw, g.E1) : 0.0;
                }
                const bool acc = accept_move(ob, obj, curfit, nfit);
                double ub, ug;
                undo_accumulators(nw, sd, ub, ug);
                const bool pert = !acc && (__double_as_longlong(ub) != __double_as_longlong(cur.bits) ||
                                           __double_as_longlong(ug) != __double_as_longlong(cur.goc));
                const unsigned ev = __ballot_sync(0xffffffffu, valid && c >= start && (acc || pert));
                if (ev == 0) break;
                const int e = __ffs(ev) - 1;
                const bool e_acc = __shfl_sync(0xffffffffu, (int)acc, e) != 0;
                if (e_acc) {
                    // commit proposal e: doSwap's state change, Alignment.cpp:496-504,516-519
                    if (c == e) {
                        const uint32_t rx = s_aln[x], ry = s_aln[y];
                        s_aln[x] = (T)ry;
                        s_aln[y] = (T)rx;
                        s_inv[ry & MASK] = (T)((uint32_t)x | (ry & FLAG));
                        s_inv[rx & MASK] = (T)((uint32_t)y | (rx & FLAG));
                    }
                    cur.C = __shfl_sync(0xffffffffu, nw.C, e);
                    cur.I = __shfl_sync(0xffffffffu, nw.I, e);
                    cur.size = __shfl_sync(0xffffffffu, nw.size, e);
                    cur.bits = __shfl_sync(0xffffffffu, nw.bits, e);
                    cur.goc = __shfl_sync(0xffffffffu, nw.goc, e);
                    consumed = e + 1;
                    acc_lane = e;
                    n_acc++;
                    break;
                }
                // rejected, but the undo left different accumulator bits: carry them forward
                cur.bits = __shfl_sync(0xffffffffu, ub, e);
                cur.goc = __shfl_sync(0xffffffffu, ug, e);
                start = e + 1;
            }
            if (accepted && c < consumed) accepted[(size_t)pl * iters + pos + c] = (c == acc_lane) ? 1 : 0;
            if (c == 0) s_pos = pos + consumed;
        }
        __syncthreads();
    }
    // write the individual back
    unstage_state<T>(pop, p, s_aln, s_inv);
    if (threadIdx.x == 0) {
        pop.cons[p] = cur.C; pop.ind[p] = cur.I; pop.size[p] = cur.size;
        pop.bits[p] = cur.bits; pop.goc[p] = cur.goc;
        for (int j = 0; j < ob.M; ++j) pop.fit[(size_t)p * ob.M + j] = objective(ob.names[j], cur, g.E1);
        if (num_accepted) num_accepted[pl] = n_acc;
    }
}

// proposals of the device RNG, for host-side checking / replay
__global__ void k_rng_proposals(int n1, int n2, int count, int K, unsigned long long seed,
                                long long island_offset, int first, long long start,
                                int* __restrict__ xy) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)count * K) return;
    const int pl = (int)(idx / K);
    const long long c = idx % K;
    const unsigned long long ikey = rng_individual_key(seed, (unsigned long long)(island_offset + first + pl));
    int x, y;
    rng_proposal(ikey, (unsigned long long)(start + c), n1, n2, x, y);
    xy[2 * idx] = x;
    xy[2 * idx + 1] = y;
}

}  // namespace ona
