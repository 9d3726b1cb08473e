"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/libona_oracle.so (ona_oracle.c).

The plain-C restatement of Alignment.cpp:491-826 and localSearch.cpp:18-120; see
ona_oracle.h for the reference line each function follows.
"""
import ctypes as C
import os

import numpy as np

from . import HERE, build

_LIB = None

I32P = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
F64P = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
U8P = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(HERE, "libona_oracle.so")
    src = os.path.join(HERE, "ona_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        build()
    L = C.CDLL(path)
    vp = C.c_void_p
    L.orc_problem_create.restype = vp
    L.orc_problem_create.argtypes = [C.c_int, C.c_int, C.c_long, I32P, I32P, I32P, I32P, vp, vp]
    L.orc_problem_free.argtypes = [vp]
    L.orc_aln_create.restype = vp
    L.orc_aln_create.argtypes = [vp, I32P, vp]
    L.orc_aln_clone.restype = vp
    L.orc_aln_clone.argtypes = [vp]
    L.orc_aln_free.argtypes = [vp]
    L.orc_aln_get_state.argtypes = [vp, vp, vp, vp]
    L.orc_aln_get_counters.argtypes = [vp] + [vp] * 4
    L.orc_aln_set_counters.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double]
    for f in ("orc_conserved_true", "orc_induced_true"):
        getattr(L, f).restype = C.c_long
        getattr(L, f).argtypes = [vp]
    for f in ("orc_ics", "orc_sum_blast", "orc_sum_goc", "orc_aln_size"):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [vp]
    L.orc_refseed_conserved.restype = C.c_int
    L.orc_refseed_conserved.argtypes = [C.c_long, C.c_long]
    L.orc_refseed_from_ics.restype = C.c_int
    L.orc_refseed_from_ics.argtypes = [C.c_double, C.c_long]
    L.orc_conserved_count.restype = C.c_int
    L.orc_conserved_count.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_induced_count.restype = C.c_int
    L.orc_induced_count.argtypes = [vp, C.c_int, C.c_int]
    L.orc_do_swap.argtypes = [vp, C.c_int, C.c_int]
    L.orc_compute_fitness.argtypes = [vp, I32P, C.c_int, vp]
    L.orc_hill_climb_replay.argtypes = [vp, I32P, C.c_long, I32P, C.c_int, C.c_int, vp]
    L.orc_swap_normalized_delta.restype = C.c_double
    L.orc_swap_normalized_delta.argtypes = [vp, I32P, C.c_int, C.c_int, C.c_int]
    L.orc_swap_delta_records.argtypes = [vp, I32P, C.c_int, C.c_int, I32P, C.c_long] + [vp] * 7
    L.orc_time_swap_deltas.restype = C.c_double
    L.orc_time_swap_deltas.argtypes = [vp, C.c_int, I32P, C.c_int, I32P, C.c_long, C.c_int, vp]
    L.orc_time_hill_climb.restype = C.c_double
    L.orc_time_hill_climb.argtypes = [vp, C.c_int, I32P, C.c_long, I32P, C.c_int, C.c_int, C.c_int]
    L.orc_time_full_fitness.restype = C.c_double
    L.orc_time_full_fitness.argtypes = [vp, C.c_int, I32P, C.c_int, C.c_long, C.c_int, vp]
    L.orc_apply_edits.argtypes = [vp, I32P, C.c_int]
    L.orc_mt_seed.argtypes = [C.c_uint32, vp]
    L.orc_mt_xy_stream.argtypes = [vp, C.c_int, C.c_int, C.c_long, I32P]
    L.orc_nsga2.restype = C.c_int
    L.orc_nsga2.argtypes = [C.c_int, C.c_int] + [vp] * 9
    L.orc_dominates.restype = C.c_int
    L.orc_dominates.argtypes = [vp, vp, C.c_int]
    L.orc_gnu_sort_by_key.argtypes = [C.c_int, vp, vp]
    L.orc_gnu_heap_sort_by_key.argtypes = [C.c_int, vp, vp]
    _LIB = L
    return L


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Problem:
    """Two networks (CSR, adjList order) + optional dense n1 x n2 score matrices."""

    def __init__(self, n1, n2, E1, rp1, col1, rp2, col2, S=None, G=None):
        self.n1, self.n2, self.E1 = int(n1), int(n2), int(E1)
        self.rp1, self.col1, self.rp2, self.col2 = _i32(rp1), _i32(col1), _i32(rp2), _i32(col2)
        self.S = None if S is None else np.ascontiguousarray(S, dtype=np.float64)
        self.G = None if G is None else np.ascontiguousarray(G, dtype=np.float64)
        self.h = lib().orc_problem_create(self.n1, self.n2, self.E1, self.rp1, self.col1, self.rp2,
                                          self.col2, _opt(self.S), _opt(self.G))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_problem_free(self.h)
            self.h = None


class Aln:
    def __init__(self, prob, aln=None, mask=None, _h=None):
        self.prob = prob
        if _h is not None:
            self.h = _h
            return
        aln = _i32(aln)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.h = lib().orc_aln_create(prob.h, aln, _opt(m))

    def clone(self):
        return Aln(self.prob, _h=lib().orc_aln_clone(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_aln_free(self.h)
            self.h = None

    def state(self):
        n = self.prob.n2
        a, inv, m = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.uint8)
        lib().orc_aln_get_state(self.h, _opt(a), _opt(inv), _opt(m))
        return a, inv, m

    def counters(self):
        c, i = C.c_int(), C.c_int()
        b, g = C.c_double(), C.c_double()
        lib().orc_aln_get_counters(self.h, C.addressof(c), C.addressof(i), C.addressof(b),
                                   C.addressof(g))
        return c.value, i.value, b.value, g.value

    def set_counters(self, cons, ind, bits, goc):
        lib().orc_aln_set_counters(self.h, cons, ind, bits, goc)

    def conserved_true(self):
        return lib().orc_conserved_true(self.h)

    def induced_true(self):
        return lib().orc_induced_true(self.h)

    def ics(self):
        return lib().orc_ics(self.h)

    def sum_blast(self):
        return lib().orc_sum_blast(self.h)

    def sum_goc(self):
        return lib().orc_sum_goc(self.h)

    def aln_size(self):
        return lib().orc_aln_size(self.h)

    def conserved_count(self, n1, n2, mask, ignore):
        return lib().orc_conserved_count(self.h, n1, n2, int(mask), ignore)

    def induced_count(self, n2, ignore):
        return lib().orc_induced_count(self.h, n2, ignore)

    def do_swap(self, x, y):
        lib().orc_do_swap(self.h, int(x), int(y))

    def apply_edits(self, ops):
        ops = _i32(ops).reshape(-1, 4)
        lib().orc_apply_edits(self.h, ops.reshape(-1), len(ops))

    def compute_fitness(self, names):
        names = _i32(names)
        out = np.empty(len(names), np.float64)
        lib().orc_compute_fitness(self.h, names, len(names), _opt(out))
        return out

    def hill_climb_replay(self, xy, names, obj=-1, want_accepted=False):
        xy, names = _i32(xy).reshape(-1), _i32(names)
        n = len(xy) // 2
        acc = np.empty(n, np.uint8) if want_accepted else None
        self.compute_fitness(names)
        lib().orc_hill_climb_replay(self.h, xy, n, names, len(names), obj, _opt(acc))
        return acc

    def swap_normalized_delta(self, names, x, y):
        names = _i32(names)
        return lib().orc_swap_normalized_delta(self.h, names, len(names), int(x), int(y))

    def swap_delta_records(self, xy, names, obj=-1, want_snd=False):
        xy, names = _i32(xy).reshape(-1), _i32(names)
        n, M = len(xy) // 2, len(names)
        r = dict(cons=np.empty(n, np.int32), ind=np.empty(n, np.int32), bits=np.empty(n, np.float64),
                 goc=np.empty(n, np.float64), fit=np.empty((n, M), np.float64),
                 improved=np.empty(n, np.uint8), snd=np.empty(n, np.float64) if want_snd else None)
        lib().orc_swap_delta_records(self.h, names, M, obj, xy, n, _opt(r["cons"]), _opt(r["ind"]),
                                     _opt(r["bits"]), _opt(r["goc"]), _opt(r["fit"]),
                                     _opt(r["improved"]), _opt(r["snd"]))
        return r


def _handles(alns):
    arr = (C.c_void_p * len(alns))(*[a.h for a in alns])
    return arr


def time_swap_deltas(alns, names, xy, nthreads):
    xy, names = _i32(xy).reshape(-1), _i32(names)
    sink = C.c_double()
    t = lib().orc_time_swap_deltas(_handles(alns), len(alns), names, len(names), xy, len(xy) // 2,
                                   nthreads, C.addressof(sink))
    return t, sink.value


def time_hill_climb(alns, names, xy, iters, obj, nthreads):
    xy, names = _i32(xy).reshape(-1), _i32(names)
    assert len(xy) == 2 * iters * len(alns)
    return lib().orc_time_hill_climb(_handles(alns), len(alns), xy, iters, names, len(names), obj,
                                     nthreads)


def time_full_fitness(alns, names, reps, nthreads):
    names = _i32(names)
    sink = C.c_double()
    t = lib().orc_time_full_fitness(_handles(alns), len(alns), names, len(names), reps, nthreads,
                                    C.addressof(sink))
    return t, sink.value


def nsga2(fitness, has_norm=None, norm=None):
    """orc_nsga2 (ona_oracle_nsga2.c); same result dictionary as oracle.ref.nsga2."""
    from .ref import _nsga2_call
    r = _nsga2_call(lib().orc_nsga2, fitness, has_norm, norm)
    if r.pop("seconds") != 0:
        raise ValueError("orc_nsga2: bad arguments")
    return r


def dominates(f1, f2):
    a, b = (np.ascontiguousarray(v, dtype=np.float64) for v in (f1, f2))
    return bool(lib().orc_dominates(a.ctypes.data, b.ctypes.data, len(a)))


def gnu_sort_by_key(keys, idx, heap=False):
    k, i = np.ascontiguousarray(keys, dtype=np.float64), np.array(idx, dtype=np.int32)
    (lib().orc_gnu_heap_sort_by_key if heap else lib().orc_gnu_sort_by_key)(len(i), k.ctypes.data,
                                                                            i.ctypes.data)
    return i


def mt_state(seed):
    """std::mt19937(seed) as its operator<< prints it: 624 state words and the position (624)."""
    st = np.empty(625, np.uint32)
    lib().orc_mt_seed(seed, st.ctypes.data_as(C.c_void_p))
    return st


def mt_xy_stream(state, n1, n2, iters):
    """the proposals correctHillClimb draws (localSearch.cpp:21-28) from an mt19937 in `state`;
    returns (xy [iters,2], state after the draws)."""
    st = np.ascontiguousarray(state, np.uint32).copy()
    xy = np.empty(2 * iters, np.int32)
    lib().orc_mt_xy_stream(st.ctypes.data_as(C.c_void_p), n1, n2, iters, xy)
    return xy.reshape(-1, 2), st
