/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the NSGA-II ranking step that consumes the
 * hot path's objective vectors.  Same rules as ona_oracle.c: only
 * tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it, never the product.
 *
 * Restates /root/reference/nsga-ii.cpp:
 *   nonDominatedSort     :14-66     dominates          :121-134
 *   normalizeFitnesses   :68-84     setCrowdingDists   :89-118
 * in the order nsgaiimain.cpp:157-175 calls them.  setCrowdingDists sorts a front with
 * std::sort, whose result on tied keys is unspecified by the C++ standard; the reference's
 * Linux targets (Makefile:23,36) get GNU libstdc++'s introsort, so this file also restates
 * that published algorithm (bits/stl_algo.h, bits/stl_heap.h: median-of-3 introsort with a
 * 16-element threshold, depth limit 2*floor(log2 n) with a heap-sort fallback, final
 * insertion sort).
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks orc_nsga2 against the
 * unmodified nsga-ii.cpp (oracle/_ref, ref_nsga2) and orc_gnu_sort_by_key against this
 * toolchain's std::sort on tie-heavy and adversarial inputs; tests/test_oracle_golden.py
 * checks it against tests/golden/nsga2.npz and the known answers of tests.cpp:355-437.
 */
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include "ona_oracle.h"

/* nsga-ii.cpp:121-134 */
int orc_dominates(const double* f1, const double* f2, int M) {
    int one_bigger = 0;
    for (int i = 0; i < M; ++i) {
        if (f1[i] < f2[i]) return 0;
        if (f1[i] > f2[i]) one_bigger = 1;
    }
    return one_bigger;
}

/* ------------------------------------------------------------------ GNU std::sort
 * idx[0..n) sorted with comp(a,b) = keys[a] < keys[b]. */
typedef struct {
    const double* keys;
    int* a;
} Srt;
#define LESS(s, i, j) ((s)->keys[(s)->a[i]] < (s)->keys[(s)->a[j]])
#define LESSV(s, v, j) ((s)->keys[v] < (s)->keys[(s)->a[j]])
#define LESSIV(s, i, v) ((s)->keys[(s)->a[i]] < (s)->keys[v])

static void srt_swap(Srt* s, long i, long j) {
    int t = s->a[i];
    s->a[i] = s->a[j];
    s->a[j] = t;
}
/* __push_heap / __adjust_heap on the sub-array starting at `first` */
static void push_heap_(Srt* s, long first, long hole, long top, int value) {
    long parent = (hole - 1) / 2;
    while (hole > top && LESSIV(s, first + parent, value)) {
        s->a[first + hole] = s->a[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    s->a[first + hole] = value;
}
static void adjust_heap_(Srt* s, long first, long hole, long len, int value) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (LESS(s, first + child, first + (child - 1))) child--;
        s->a[first + hole] = s->a[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        s->a[first + hole] = s->a[first + (child - 1)];
        hole = child - 1;
    }
    push_heap_(s, first, hole, top, value);
}
/* partial_sort(first, last, last): make_heap + sort_heap */
static void heap_sort_(Srt* s, long first, long last) {
    const long len = last - first;
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            int value = s->a[first + parent];
            adjust_heap_(s, first, parent, len, value);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        int value = s->a[last];
        s->a[last] = s->a[first];
        adjust_heap_(s, first, 0, last - first, value);
    }
}
static void move_median_to_first_(Srt* s, long result, long a, long b, long c) {
    if (LESS(s, a, b)) {
        if (LESS(s, b, c)) srt_swap(s, result, b);
        else if (LESS(s, a, c)) srt_swap(s, result, c);
        else srt_swap(s, result, a);
    } else if (LESS(s, a, c)) srt_swap(s, result, a);
    else if (LESS(s, b, c)) srt_swap(s, result, c);
    else srt_swap(s, result, b);
}
static long unguarded_partition_(Srt* s, long first, long last, long pivot) {
    for (;;) {
        while (LESS(s, first, pivot)) ++first;
        --last;
        while (LESS(s, pivot, last)) --last;
        if (!(first < last)) return first;
        srt_swap(s, first, last);
        ++first;
    }
}
static void introsort_loop_(Srt* s, long first, long last, long depth) {
    while (last - first > 16) {
        if (depth == 0) {
            heap_sort_(s, first, last);
            return;
        }
        --depth;
        const long mid = first + (last - first) / 2;
        move_median_to_first_(s, first, first + 1, mid, last - 1);
        const long cut = unguarded_partition_(s, first + 1, last, first);
        introsort_loop_(s, cut, last, depth);
        last = cut;
    }
}
static void unguarded_linear_insert_(Srt* s, long last) {
    const int val = s->a[last];
    long next = last - 1;
    while (LESSV(s, val, next)) {
        s->a[last] = s->a[next];
        last = next;
        --next;
    }
    s->a[last] = val;
}
static void insertion_sort_(Srt* s, long first, long last) {
    if (first == last) return;
    for (long i = first + 1; i != last; ++i) {
        if (LESS(s, i, first)) {
            const int val = s->a[i];
            memmove(s->a + first + 1, s->a + first, (size_t)(i - first) * sizeof(int));
            s->a[first] = val;
        } else
            unguarded_linear_insert_(s, i);
    }
}
void orc_gnu_sort_by_key(int n, const double* keys, int* idx) {
    if (n <= 0) return;
    Srt s = {keys, idx};
    long lg = 0;
    for (long v = n; v > 1; v >>= 1) ++lg;
    introsort_loop_(&s, 0, n, 2 * lg);
    if (n > 16) {
        insertion_sort_(&s, 0, 16);
        for (long i = 16; i != n; ++i) unguarded_linear_insert_(&s, i);
    } else
        insertion_sort_(&s, 0, n);
}

/* the depth-limit fallback alone (std::partial_sort(first, last, last)), so that tests can pin it
 * without having to construct an input that exhausts introsort's depth limit */
void orc_gnu_heap_sort_by_key(int n, const double* keys, int* idx) {
    Srt s = {keys, idx};
    heap_sort_(&s, 0, n);
}

/* ------------------------------------------------------------------ orc_nsga2
 * Arguments as oracle/ref_shim.cpp:ref_nsga2. */
int orc_nsga2(int P, int M, const double* fitness, const unsigned char* has_norm, double* norm_io,
              int* dom_rank, int* front_order, int* front_offsets, int* n_fronts,
              int* order_after, double* crowd_dist) {
    if (P <= 0 || M <= 0) return -1;
    /* :17-36  domination counts and lists (index order) */
    int* count = (int*)calloc((size_t)P, sizeof(int));
    long* lp = (long*)calloc((size_t)P + 1, sizeof(long));
    for (int i = 0; i < P; ++i)
        for (int j = 0; j < P; ++j) {
            if (i == j) continue;
            const double *fi = fitness + (size_t)i * M, *fj = fitness + (size_t)j * M;
            if (orc_dominates(fi, fj, M)) lp[i + 1]++;
            else if (orc_dominates(fj, fi, M)) count[i]++;
        }
    for (int i = 0; i < P; ++i) lp[i + 1] += lp[i];
    int* lists = (int*)malloc(((size_t)lp[P] + 1) * sizeof(int));
    {
        long* fill = (long*)malloc((size_t)P * sizeof(long));
        for (int i = 0; i < P; ++i) fill[i] = lp[i];
        for (int i = 0; i < P; ++i)
            for (int j = 0; j < P; ++j)
                if (i != j && orc_dominates(fitness + (size_t)i * M, fitness + (size_t)j * M, M))
                    lists[fill[i]++] = j;
        free(fill);
    }
    int n = 0, nf = 0;
    front_offsets[0] = 0;
    for (int i = 0; i < P; ++i) {
        dom_rank[i] = -1;
        if (count[i] == 0) {
            dom_rank[i] = 0;
            front_order[n++] = i;
        }
    }
    nf = 1;
    front_offsets[1] = n;
    /* :38-54  peel: members in front order, their dominated lists in index order */
    for (int f = 0; f < nf && front_offsets[f + 1] > front_offsets[f]; ++f) {
        const int b = front_offsets[f], e = front_offsets[f + 1];
        for (int k = b; k < e; ++k) {
            const int p = front_order[k];
            for (long t = lp[p]; t < lp[p + 1]; ++t) {
                const int q = lists[t];
                if (--count[q] == 0) {
                    dom_rank[q] = f + 1;
                    front_order[n++] = q;
                }
            }
        }
        if (n > e) front_offsets[++nf] = n;
    }
    *n_fronts = nf;
    /* :68-84  normalise against the population maximum (starting from 0.0, strict >); only
     * individuals without stored values take the new ones (append-only vector, :79) */
    for (int m = 0; m < M; ++m) {
        double mx = 0.0;
        for (int i = 0; i < P; ++i)
            if (fitness[(size_t)i * M + m] > mx) mx = fitness[(size_t)i * M + m];
        for (int i = 0; i < P; ++i)
            if (!(has_norm && has_norm[i])) norm_io[(size_t)i * M + m] = fitness[(size_t)i * M + m] / mx;
    }
    /* :89-118 per front */
    memcpy(order_after, front_order, (size_t)P * sizeof(int));
    double* keys = (double*)malloc((size_t)P * sizeof(double));
    for (int f = 0; f < nf; ++f) {
        int* in = order_after + front_offsets[f];
        const int na = front_offsets[f + 1] - front_offsets[f];
        if (na == 0) continue;
        for (int k = 0; k < na; ++k) crowd_dist[in[k]] = 0.0;
        for (int m = 0; m < M; ++m) {
            for (int k = 0; k < na; ++k) keys[in[k]] = norm_io[(size_t)in[k] * M + m];
            orc_gnu_sort_by_key(na, keys, in);
            crowd_dist[in[0]] = DBL_MAX;
            crowd_dist[in[na - 1]] = DBL_MAX;
            const double denom = keys[in[na - 1]] - keys[in[0]];
            for (int k = 1; k < na - 1; ++k) {
                const double num = keys[in[k + 1]] - keys[in[k - 1]];
                crowd_dist[in[k]] += num / denom;
            }
        }
    }
    free(keys);
    free(lists);
    free(lp);
    free(count);
    return 0;
}
