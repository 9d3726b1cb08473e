"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libona_ref.so.

That library is the UNMODIFIED reference (Network.cpp, blastinfo.cpp, goc.cpp,
Alignment.cpp, localSearch.cpp, nsga-ii.cpp, Archive.cpp compiled where they lie under
/root/reference) plus the extern "C" glue in oracle/ref_shim.cpp.  It is built in the dev
container by oracle/Makefile and travels to the GPU box prebuilt; ``available()`` is False
where it did not.
"""
import ctypes as C
import os
import tempfile

import numpy as np

from . import HERE

_LIB = None
PATH = os.path.join(HERE, "_ref", "libona_ref.so")

I32P = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
F64P = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def available():
    if not os.path.exists(PATH):
        return False
    try:
        lib()
        return True
    except OSError:
        return False


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(PATH)
    vp = C.c_void_p
    L.ref_last_error.restype = C.c_char_p
    L.ref_net_load.restype = vp
    L.ref_net_load.argtypes = [C.c_char_p]
    L.ref_net_free.argtypes = [vp]
    L.ref_net_n.argtypes = [vp]
    L.ref_net_num_edges.restype = C.c_long
    L.ref_net_num_edges.argtypes = [vp]
    L.ref_net_adj_total.restype = C.c_long
    L.ref_net_adj_total.argtypes = [vp]
    L.ref_net_export_csr.argtypes = [vp, I32P, I32P]
    L.ref_net_adjmatrix_bit.argtypes = [vp, C.c_int, C.c_int]
    L.ref_net_node_name.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
    L.ref_blast_load.restype = vp
    L.ref_blast_load.argtypes = [vp, vp, C.c_char_p]
    L.ref_goc_load.restype = vp
    L.ref_goc_load.argtypes = [vp, vp, C.c_char_p, C.c_char_p]
    L.ref_dict_from_dense.restype = vp
    L.ref_dict_from_dense.argtypes = [C.c_int, C.c_int, F64P]
    L.ref_dict_free.argtypes = [vp]
    L.ref_dict_get.restype = C.c_double
    L.ref_dict_get.argtypes = [vp, C.c_int, C.c_int]
    L.ref_dict_row.argtypes = [vp, C.c_int, C.c_int, F64P]
    L.ref_aln_new.restype = vp
    L.ref_aln_new.argtypes = [vp, vp, vp, vp]
    L.ref_aln_clone.restype = vp
    L.ref_aln_clone.argtypes = [vp]
    L.ref_aln_free.argtypes = [vp]
    L.ref_aln_shuf.argtypes = [vp, C.c_uint32, C.c_int, C.c_int, C.c_int]
    L.ref_aln_set_state.argtypes = [vp, I32P, vp]
    L.ref_aln_size_n2.argtypes = [vp]
    L.ref_aln_get_state.argtypes = [vp, vp, vp, vp]
    L.ref_aln_get_counters.argtypes = [vp] * 5
    L.ref_aln_set_counters.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double]
    L.ref_aln_do_swap.argtypes = [vp, C.c_int, C.c_int]
    L.ref_aln_compute_fitness.argtypes = [vp, I32P, C.c_int, F64P]
    for f in ("ref_aln_ics", "ref_aln_fast_ics_denominator", "ref_aln_sum_blast", "ref_aln_sum_goc",
              "ref_aln_aln_size"):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [vp]
    L.ref_aln_conserved_count.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ref_aln_induced_count.argtypes = [vp, C.c_int, C.c_int]
    L.ref_aln_mutate.argtypes = [vp, C.c_uint32, C.c_float, C.c_int]
    L.ref_aln_on_bit.argtypes = [vp, C.c_int]
    L.ref_aln_greedy_match.argtypes = [vp, C.c_int]
    L.ref_steepest_ascent.argtypes = [vp, I32P, C.c_int, C.c_int]
    L.ref_proportional_search.argtypes = [vp, C.c_uint32, C.c_int, I32P, C.c_int, C.c_int, C.c_double]
    L.ref_aln_load.restype = vp
    L.ref_aln_load.argtypes = [vp, vp, C.c_char_p, vp, vp]
    L.ref_gen_mutate_ops.restype = C.c_int
    L.ref_gen_mutate_ops.argtypes = [C.c_uint32, C.c_int, C.c_float, C.c_int, vp, C.c_int]
    L.ref_gen_cross_draws.restype = C.c_int
    L.ref_gen_cross_draws.argtypes = [C.c_uint32, C.c_int, C.c_float, C.c_int, vp, vp, C.c_int]
    L.ref_aln_crossover.restype = vp
    L.ref_aln_crossover.argtypes = [C.c_uint32, C.c_float, vp, vp, C.c_int]
    L.ref_correct_hill_climb.argtypes = [vp, C.c_uint32, C.c_int, I32P, C.c_int, C.c_int]
    L.ref_gen_xy_stream.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_long, I32P]
    L.ref_hill_climb_replay.argtypes = [vp, I32P, C.c_long, I32P, C.c_int, C.c_int, vp]
    L.ref_swap_normalized_delta.restype = C.c_double
    L.ref_swap_normalized_delta.argtypes = [vp, I32P, C.c_int, C.c_int, C.c_int]
    L.ref_swap_delta_records.argtypes = [vp, I32P, C.c_int, C.c_int, I32P, C.c_long] + [vp] * 7
    L.ref_time_swap_deltas.restype = C.c_double
    L.ref_time_swap_deltas.argtypes = [vp, C.c_int, I32P, C.c_int, I32P, C.c_long, C.c_int, vp]
    L.ref_time_hill_climb.restype = C.c_double
    L.ref_time_hill_climb.argtypes = [vp, C.c_int, C.c_uint32, C.c_int, I32P, C.c_int, C.c_int,
                                      C.c_int]
    L.ref_time_full_fitness.restype = C.c_double
    L.ref_time_full_fitness.argtypes = [vp, C.c_int, I32P, C.c_int, C.c_long, C.c_int, C.c_int, vp]
    L.ref_hardware_threads.restype = C.c_int
    L.ref_nsga2.restype = C.c_double
    L.ref_nsga2.argtypes = [C.c_int, C.c_int] + [vp] * 9
    L.ref_dominates.restype = C.c_int
    L.ref_dominates.argtypes = [vp, vp, C.c_int]
    L.ref_std_sort_by_key.argtypes = [C.c_int, vp, vp]
    L.ref_std_partial_sort_by_key.argtypes = [C.c_int, vp, vp]
    _LIB = L
    return L


def _check(h, what):
    if not h:
        raise RuntimeError("%s: %s" % (what, lib().ref_last_error().decode()))
    return h


class Net:
    """reference ``Network`` (Network.h:98-108) loaded from an edge-list file."""

    def __init__(self, path):
        self.h = _check(lib().ref_net_load(path.encode()), "Network(%s)" % path)
        self.n = lib().ref_net_n(self.h)
        self.E = lib().ref_net_num_edges(self.h)

    def csr(self):
        rp = np.empty(self.n + 1, np.int32)
        col = np.empty(max(lib().ref_net_adj_total(self.h), 1), np.int32)
        lib().ref_net_export_csr(self.h, rp, col)
        return rp, col[: rp[-1]].copy()

    def adj_bit(self, u, v):
        return lib().ref_net_adjmatrix_bit(self.h, u, v)

    def name(self, u):
        buf = C.create_string_buffer(256)
        lib().ref_net_node_name(self.h, u, buf, 256)
        return buf.value.decode()

    @staticmethod
    def from_edges(edges, prefix):
        """Write ``prefix<u> prefix<v>`` lines to a temp file and load it.  Node ids are
        assigned by first appearance (Network.cpp:59-83), so callers pass edge lists that
        are already in first-appearance-canonical numbering (synth.canonicalize)."""
        with tempfile.NamedTemporaryFile("w", suffix=".net", delete=False) as f:
            for u, v in edges:
                f.write("%s%d\t%s%d\n" % (prefix, u, prefix, v))
            path = f.name
        try:
            return Net(path)
        finally:
            os.unlink(path)


class Dict:
    """reference BLASTDict / GOCDict (dense n1 x n2 doubles)."""

    def __init__(self, h, rows, cols):
        self.h, self.rows, self.cols = h, rows, cols

    @staticmethod
    def load_blast(net1, net2, path):
        return Dict(_check(lib().ref_blast_load(net1.h, net2.h, path.encode()), "loadBLASTInfo"),
                    net1.n, net2.n)

    @staticmethod
    def load_goc(net1, net2, f1, f2):
        return Dict(_check(lib().ref_goc_load(net1.h, net2.h, f1.encode(), f2.encode()), "loadGOC"),
                    net1.n, net2.n)

    @staticmethod
    def from_dense(M):
        M = np.ascontiguousarray(M, dtype=np.float64)
        return Dict(lib().ref_dict_from_dense(M.shape[0], M.shape[1], M), M.shape[0], M.shape[1])

    def dense(self):
        out = np.empty((self.rows, self.cols), np.float64)
        for i in range(self.rows):
            lib().ref_dict_row(self.h, i, self.cols, out[i])
        return out

    def get(self, i, j):
        return lib().ref_dict_get(self.h, i, j)


class Aln:
    """reference ``Alignment`` (Alignment.h:13-113)."""

    def __init__(self, net1=None, net2=None, bit=None, goc=None, _h=None):
        if _h is not None:
            self.h = _h
            return
        self.h = _check(lib().ref_aln_new(net1.h, net2.h, bit.h if bit else None,
                                          goc.h if goc else None), "Alignment()")

    @staticmethod
    def load(net1, net2, path, bit=None, goc=None):
        """the alignment-file constructor (Alignment.cpp:49-138)"""
        return Aln(_h=_check(lib().ref_aln_load(net1.h, net2.h, path.encode(), bit.h if bit else None,
                                               goc.h if goc else None), "Alignment(file)"))

    def clone(self):
        return Aln(_h=lib().ref_aln_clone(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_aln_free(self.h)
            self.h = None

    @property
    def n2(self):
        return lib().ref_aln_size_n2(self.h)

    def shuf(self, seed, uniformsize=False, smallstart=False, total=True):
        lib().ref_aln_shuf(self.h, seed, int(uniformsize), int(smallstart), int(total))

    def set_state(self, aln, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().ref_aln_set_state(self.h, _i32(aln), _opt(m))

    def state(self):
        n = self.n2
        a, inv, m = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.uint8)
        lib().ref_aln_get_state(self.h, _opt(a), _opt(inv), _opt(m))
        return a, inv, m

    def counters(self):
        c, i = C.c_int(), C.c_int()
        b, g = C.c_double(), C.c_double()
        lib().ref_aln_get_counters(self.h, C.addressof(c), C.addressof(i), C.addressof(b),
                                   C.addressof(g))
        return c.value, i.value, b.value, g.value

    def set_counters(self, cons, ind, bits, goc):
        lib().ref_aln_set_counters(self.h, cons, ind, bits, goc)

    def do_swap(self, x, y):
        lib().ref_aln_do_swap(self.h, int(x), int(y))

    def compute_fitness(self, names):
        names = _i32(names)
        out = np.empty(len(names), np.float64)
        lib().ref_aln_compute_fitness(self.h, names, len(names), out)
        return out

    def ics(self):
        return lib().ref_aln_ics(self.h)

    def fast_ics_denominator(self):
        return lib().ref_aln_fast_ics_denominator(self.h)

    def sum_blast(self):
        return lib().ref_aln_sum_blast(self.h)

    def sum_goc(self):
        return lib().ref_aln_sum_goc(self.h)

    def aln_size(self):
        return lib().ref_aln_aln_size(self.h)

    def conserved_count(self, n1, n2, mask, ignore):
        return lib().ref_aln_conserved_count(self.h, n1, n2, int(mask), ignore)

    def induced_count(self, n2, ignore):
        return lib().ref_aln_induced_count(self.h, n2, ignore)

    def greedy_match(self, bit=True):
        lib().ref_aln_greedy_match(self.h, int(bit))

    def on_bit(self, x):
        lib().ref_aln_on_bit(self.h, int(x))

    def mutate(self, seed, mutswappb, total=True):
        lib().ref_aln_mutate(self.h, seed, mutswappb, int(total))

    @staticmethod
    def crossover(seed, cxswappb, p1, p2, total=True):
        return Aln(_h=lib().ref_aln_crossover(seed, cxswappb, p1.h, p2.h, int(total)))

    def correct_hill_climb(self, seed, iters, names, obj=-1):
        names = _i32(names)
        self.compute_fitness(names)
        lib().ref_correct_hill_climb(self.h, seed, iters, names, len(names), obj)

    def proportional_search(self, seed, iters, names, obj, proportion):
        """proportionalSearch(mt19937(seed), ...) (localSearch.cpp:60-84)"""
        names = _i32(names)
        lib().ref_proportional_search(self.h, seed, iters, names, len(names), obj, proportion)

    def steepest_ascent(self, names, nthreads=2):
        """steepestAscentHillClimb until convergence (localSearch.cpp:122-203)"""
        names = _i32(names)
        lib().ref_steepest_ascent(self.h, names, len(names), nthreads)

    def hill_climb_replay(self, xy, names, obj=-1, want_accepted=False):
        xy, names = _i32(xy).reshape(-1), _i32(names)
        n = len(xy) // 2
        acc = np.empty(n, np.uint8) if want_accepted else None
        self.compute_fitness(names)
        lib().ref_hill_climb_replay(self.h, xy, n, names, len(names), obj, _opt(acc))
        return acc

    def swap_normalized_delta(self, names, x, y):
        names = _i32(names)
        return lib().ref_swap_normalized_delta(self.h, names, len(names), int(x), int(y))

    def swap_delta_records(self, xy, names, obj=-1, want_snd=False):
        xy, names = _i32(xy).reshape(-1), _i32(names)
        n, M = len(xy) // 2, len(names)
        r = dict(cons=np.empty(n, np.int32), ind=np.empty(n, np.int32), bits=np.empty(n, np.float64),
                 goc=np.empty(n, np.float64), fit=np.empty((n, M), np.float64),
                 improved=np.empty(n, np.uint8), snd=np.empty(n, np.float64) if want_snd else None)
        lib().ref_swap_delta_records(self.h, names, M, obj, xy, n, _opt(r["cons"]), _opt(r["ind"]),
                                     _opt(r["bits"]), _opt(r["goc"]), _opt(r["fit"]),
                                     _opt(r["improved"]), _opt(r["snd"]))
        return r


def gen_mutate_ops(seed, size, mutswappb, total):
    """the edit list Alignment::mutate(mt19937(seed), mutswappb, total) performs (Alignment.cpp:451-480)"""
    ops = np.zeros((2 * size, 4), np.int32)
    n = lib().ref_gen_mutate_ops(seed, size, mutswappb, int(total), ops.ctypes.data, len(ops))
    return ops[:n].copy()


def gen_cross_ops(seed, cxswappb, p1_state, p2_state, total):
    """(primary parent is p1?, edit list) of the UPMX constructor Alignment(mt19937(seed), cxswappb,
    p1, p2, total) (Alignment.cpp:152-256); p?_state = (aln, mask) arrays of the two parents."""
    size = len(p1_state[0])
    first, sel = C.c_int(0), np.zeros((size, 2), np.int32)
    n = lib().ref_gen_cross_draws(seed, size, cxswappb, int(total), C.addressof(first), sel.ctypes.data, size)
    par2 = p2_state if first.value else p1_state
    ops = np.zeros((n, 4), np.int32)
    ops[:, 0] = 3
    ops[:, 1] = sel[:n, 0]
    ops[:, 2] = np.asarray(par2[0])[sel[:n, 0]]
    ops[:, 3] = (np.asarray(par2[1])[sel[:n, 0]] != 0) | (sel[:n, 1] << 1) | (4 if total else 0)
    return bool(first.value), ops


def gen_xy_stream(seed, n1, n2, iters):
    """the proposals correctHillClimb draws from mt19937(seed) (localSearch.cpp:21-28)."""
    xy = np.empty(2 * iters, np.int32)
    lib().ref_gen_xy_stream(seed, n1, n2, iters, xy)
    return xy.reshape(-1, 2)


def mt_state(seed):
    """std::mt19937(seed) as operator<< prints it: 624 state words and the position."""
    st = np.empty(625, np.uint32)
    lib().ref_mt_state(C.c_uint32(seed), st.ctypes.data_as(C.c_void_p))
    return st


def gen_xy_stream_state(state, n1, n2, iters):
    """proposals of localSearch.cpp:21-28 from an explicit mt19937 state; returns (xy, state after)."""
    st = np.ascontiguousarray(state, np.uint32).copy()
    xy = np.empty(2 * iters, np.int32)
    lib().ref_gen_xy_stream_state(st.ctypes.data_as(C.c_void_p), C.c_int(n1), C.c_int(n2), C.c_long(iters),
                                  xy.ctypes.data_as(C.c_void_p))
    return xy.reshape(-1, 2), st


def _handles(alns):
    return (C.c_void_p * len(alns))(*[a.h for a in alns])


def time_swap_deltas(alns, names, xy, nthreads):
    xy, names = _i32(xy).reshape(-1), _i32(names)
    sink = C.c_double()
    t = lib().ref_time_swap_deltas(_handles(alns), len(alns), names, len(names), xy, len(xy) // 2,
                                   nthreads, C.addressof(sink))
    return t, sink.value


def time_hill_climb(alns, names, seed0, iters, obj, nthreads):
    names = _i32(names)
    return lib().ref_time_hill_climb(_handles(alns), len(alns), seed0, iters, names, len(names), obj,
                                     nthreads)


def time_full_fitness(alns, names, reps, nthreads, with_ics=True):
    names = _i32(names)
    sink = C.c_double()
    t = lib().ref_time_full_fitness(_handles(alns), len(alns), names, len(names), reps, nthreads,
                                    int(with_ics), C.addressof(sink))
    return t, sink.value


def hardware_threads():
    return lib().ref_hardware_threads()


def _nsga2_call(fn, fitness, has_norm, norm):
    """shared argument marshalling of ref_nsga2 / orc_nsga2 (oracle/port.py imports it)."""
    f = np.ascontiguousarray(fitness, dtype=np.float64)
    P, M = f.shape
    hn = None if has_norm is None else np.ascontiguousarray(has_norm, dtype=np.uint8)
    nio = np.zeros((P, M), np.float64) if norm is None else np.array(norm, dtype=np.float64, order="C")
    rank, order, after = (np.full(P, -1, np.int32) for _ in range(3))
    offs, nf, crowd = np.zeros(P + 1, np.int32), C.c_int(0), np.zeros(P, np.float64)
    t = fn(P, M, f.ctypes.data, hn.ctypes.data if hn is not None else None, nio.ctypes.data,
           rank.ctypes.data, order.ctypes.data, offs.ctypes.data, C.addressof(nf), after.ctypes.data,
           crowd.ctypes.data)
    return {"dom_rank": rank, "front_order": order, "front_offsets": offs[:nf.value + 1].copy(),
            "order_after": after, "crowd_dist": crowd, "norm": nio, "seconds": t}


def nsga2(fitness, has_norm=None, norm=None):
    """nonDominatedSort + normalizeFitnesses + setCrowdingDists on every front (nsga-ii.cpp:14-118)."""
    r = _nsga2_call(lib().ref_nsga2, fitness, has_norm, norm)
    if r["seconds"] < 0:
        raise RuntimeError("ref_nsga2: " + lib().ref_last_error().decode())
    return r


def dominates(f1, f2):
    a, b = (np.ascontiguousarray(v, dtype=np.float64) for v in (f1, f2))
    return bool(lib().ref_dominates(a.ctypes.data, b.ctypes.data, len(a)))


def std_sort_by_key(keys, idx, heap=False):
    """std::sort (or, heap=True, std::partial_sort over the whole range: introsort's fallback)."""
    k, i = np.ascontiguousarray(keys, dtype=np.float64), np.array(idx, dtype=np.int32)
    (lib().ref_std_partial_sort_by_key if heap else lib().ref_std_sort_by_key)(len(i), k.ctypes.data,
                                                                               i.ctypes.data)
    return i
