"""ctypes binding of the C ABI in include/ona_b200.h (libona_b200.so).

This is the Python-side host mirror used by the tests and bench.py; the C++ host mirror of
the reference's Alignment / localSearch interface lives in optnetaligncpp_b200/host/.
There is no CPU fallback: loading fails loudly when the CUDA library is missing, and
``Engine`` raises when no GPU is usable.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

# objective ids == the reference's enum fitnessName (Network.h:31-32)
ICS, EC, BITSCORE_SUM, EVALS_SUM, SIZE, GOC, S3, S3_DENOM, ICS_TIMES_EC, S3_VARIANT = range(10)
OBJECTIVE_NAMES = ["ICS", "EC", "BitscoreSum", "EvalsSum", "Size", "GOC", "S3", "S3Denom",
                   "ICS*EC", "S3Variant"]            # fitnessNameToStr, Network.cpp:12-37
SEED_NONE, SEED_TRUE, SEED_REFCOMPAT, SUM_SEQUENTIAL = 0, 1, 2, 16

EXPORTS = [
    "ona_device_count", "ona_engine_create", "ona_engine_fork", "ona_engine_destroy", "ona_last_error",
    "ona_set_stream", "ona_synchronize", "ona_host_alloc", "ona_host_free", "ona_kernel_launches",
    "ona_set_island_offset", "ona_set_objectives", "ona_pop_resize", "ona_pop_upload", "ona_pop_download",
    "ona_pop_set_counters", "ona_pop_get_counters", "ona_full_fitness", "ona_full_fitness_dev",
    "ona_swap_deltas", "ona_swap_deltas_rng", "ona_swap_deltas_dev", "ona_steepest_step", "ona_rng_proposals", "ona_hill_climb", "ona_hill_climb_lists", "ona_hill_climb_batched",
    "ona_hill_climb_mt", "ona_mt_proposals", "ona_probe_gather_rate", "ona_goc_matrix",
    "ona_apply_swaps", "ona_apply_edits", "ona_export_fitness_dev", "ona_nsga2_rank", "ona_nsga2_rank_dev",
]


EDIT_SWAP, EDIT_FLIP, EDIT_ONBIT, EDIT_CROSS = 0, 1, 2, 3


class OnaError(RuntimeError):
    pass


class _ScoreTable(C.Structure):
    _fields_ = [("dense", C.c_void_p), ("rows", C.c_void_p), ("cols", C.c_void_p),
                ("vals", C.c_void_p), ("nnz", C.c_longlong),
                ("ann_rowptr1", C.c_void_p), ("ann_terms1", C.c_void_p), ("ann_rowptr2", C.c_void_p),
                ("ann_terms2", C.c_void_p), ("ann_present1", C.c_void_p), ("ann_present2", C.c_void_p)]


class _Problem(C.Structure):
    _fields_ = [("n1", C.c_int), ("n2", C.c_int), ("num_edges1", C.c_longlong),
                ("rowptr1", C.c_void_p), ("col1", C.c_void_p), ("rowptr2", C.c_void_p),
                ("col2", C.c_void_p), ("bitscores", _ScoreTable), ("gocs", _ScoreTable),
                ("objectives", C.c_void_p), ("num_objectives", C.c_int)]


_LIB = None


def load_library(build_if_missing=True):
    """dlopen libona_b200.so (building it in-tree first when nvcc is around)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if build_if_missing:
        try:
            _build.build()
        except Exception:
            if not os.path.exists(_build.LIB):
                raise
    if not os.path.exists(_build.LIB):
        raise OnaError("libona_b200.so is missing: build it with `python -m optnetaligncpp_b200.build` "
                       "(there is no CPU fallback)")
    L = C.CDLL(_build.LIB)
    vp, ci, ll = C.c_void_p, C.c_int, C.c_longlong
    L.ona_last_error.restype = C.c_char_p
    L.ona_engine_create.argtypes = [ci, C.POINTER(_Problem), C.POINTER(vp)]
    L.ona_engine_fork.argtypes = [vp, C.POINTER(vp)]
    L.ona_hill_climb_lists.argtypes = [vp, ci, ci, ll, vp, vp, vp, vp]
    L.ona_engine_destroy.argtypes = [vp]
    L.ona_engine_destroy.restype = None
    L.ona_set_stream.argtypes = [vp, vp]
    L.ona_synchronize.argtypes = [vp]
    L.ona_host_alloc.restype = vp
    L.ona_host_alloc.argtypes = [ll]
    L.ona_host_free.argtypes = [vp]
    L.ona_host_free.restype = None
    L.ona_kernel_launches.restype = ll
    L.ona_kernel_launches.argtypes = [vp]
    L.ona_set_island_offset.argtypes = [vp, ll]
    L.ona_set_objectives.argtypes = [vp, vp, ci]
    L.ona_pop_resize.argtypes = [vp, ci]
    L.ona_pop_upload.argtypes = [vp, ci, ci, vp, vp]
    L.ona_pop_download.argtypes = [vp, ci, ci, vp, vp, vp]
    L.ona_pop_set_counters.argtypes = [vp, ci, ci, vp, vp, vp, vp]
    L.ona_pop_get_counters.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, vp]
    L.ona_full_fitness.argtypes = [vp, ci, ci, ci] + [vp] * 8
    L.ona_full_fitness_dev.argtypes = [vp, ci, ci, ci]
    L.ona_swap_deltas.argtypes = [vp, ci, ci, ci, vp, ci] + [vp] * 7
    L.ona_swap_deltas_rng.argtypes = [vp, ci, ci, ci, C.c_ulonglong, ci, vp, vp, vp]
    L.ona_swap_deltas_dev.argtypes = [vp, ci, ci, ci, vp, C.c_ulonglong, ci] + [vp] * 8
    L.ona_steepest_step.argtypes = [vp, ci, vp, vp, vp]
    L.ona_rng_proposals.argtypes = [vp, ci, ci, ci, C.c_ulonglong, ll, vp]
    L.ona_hill_climb.argtypes = [vp, ci, ci, ll, vp, C.c_ulonglong, vp, vp, vp]
    L.ona_hill_climb_batched.argtypes = [vp, ci, ci, ll, vp, C.c_ulonglong, vp, vp]
    L.ona_hill_climb_mt.argtypes = [vp, ci, ci, ll, vp, vp, vp, vp]
    L.ona_mt_proposals.argtypes = [vp, ci, ci, ci, ll, vp, vp]
    L.ona_probe_gather_rate.argtypes = [vp, ll, vp, vp]
    L.ona_goc_matrix.argtypes = [ci, ci, ci] + [vp] * 7
    L.ona_apply_swaps.argtypes = [vp, ci, ci, ci, vp]
    L.ona_export_fitness_dev.argtypes = [vp, ci, ci, vp]
    L.ona_apply_edits.argtypes = [vp, ci, ci, ci, vp]
    L.ona_nsga2_rank.argtypes = [vp, ci, ci] + [vp] * 9
    L.ona_nsga2_rank_dev.argtypes = [vp, ci, ci] + [vp] * 9
    _LIB = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Engine:
    """One scoring engine bound to one GPU: two networks, score tables, objective list and a
    population of alignment slots.  Mirrors the data an ``Alignment`` borrows
    (Alignment.h:73-98) and the batched forms of its hot-path methods."""

    def __init__(self, n1, n2, num_edges1, rp1, col1, rp2, col2, objectives, bits=None, goc=None,
                 device=0):
        """bits / goc: None, a dense [n1,n2] float64 array, or a (rows, cols, vals) COO triple."""
        self.L = load_library()
        self.n1, self.n2, self.M = int(n1), int(n2), len(objectives)
        self.objectives = _i32(objectives)
        keep = [_i32(rp1), _i32(col1), _i32(rp2), _i32(col2), self.objectives]
        pr = _Problem()
        pr.n1, pr.n2, pr.num_edges1 = self.n1, self.n2, int(num_edges1)
        pr.rowptr1, pr.col1, pr.rowptr2, pr.col2 = (_ptr(a) for a in keep[:4])
        pr.objectives, pr.num_objectives = _ptr(self.objectives), self.M
        for field, t in (("bitscores", bits), ("gocs", goc)):
            st = _ScoreTable()
            if t is None:
                pass
            elif isinstance(t, dict):          # term sets: {"ann1": [arrays], "ann2": [arrays]}
                arrs = []
                for ann in (t["ann1"], t["ann2"]):
                    rp = np.zeros(len(ann) + 1, np.int32)
                    np.cumsum([len(a) for a in ann], out=rp[1:])
                    tm = _i32(np.concatenate([np.asarray(a, np.int32) for a in ann] or [[]])) if rp[-1] else np.zeros(1, np.int32)
                    arrs += [rp, tm]
                keep += arrs
                st.ann_rowptr1, st.ann_terms1, st.ann_rowptr2, st.ann_terms2 = (_ptr(a) for a in arrs)
                for k, fld in (("present1", "ann_present1"), ("present2", "ann_present2")):
                    if t.get(k) is not None:
                        pr_ = np.ascontiguousarray(t[k], dtype=np.uint8)
                        keep.append(pr_)
                        setattr(st, fld, _ptr(pr_))
            elif isinstance(t, tuple):
                r, c, v = _i32(t[0]), _i32(t[1]), np.ascontiguousarray(t[2], dtype=np.float64)
                keep += [r, c, v]
                st.rows, st.cols, st.vals, st.nnz = _ptr(r), _ptr(c), _ptr(v), len(v)
            else:
                d = np.ascontiguousarray(t, dtype=np.float64)
                assert d.shape == (self.n1, self.n2)
                keep.append(d)
                st.dense = _ptr(d)
            setattr(pr, field, st)
        h = C.c_void_p()
        self.h = None
        self._check(self.L.ona_engine_create(int(device), C.byref(pr), C.byref(h)))
        self.h = h
        self.device = int(device)
        self.P = 0

    @classmethod
    def from_pair(cls, pair, objectives, use_bits=True, use_goc=False, goc_dense=None, device=0):
        """Engine for a synth.NetPair."""
        bits = None
        if use_bits and pair.bit_rows is not None:
            bits = (pair.bit_rows, pair.bit_cols, pair.bit_vals)
        goc = goc_dense if use_goc else None
        return cls(pair.n1, pair.n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, objectives,
                   bits=bits, goc=goc, device=device)

    def fork(self):
        """ona_engine_fork: a second queue (own slots, stream, lock) on the same device data."""
        f = object.__new__(Engine)
        f.L, f.n1, f.n2, f.M, f.objectives, f.P = self.L, self.n1, self.n2, self.M, self.objectives, 0
        f.device = self.device
        h = C.c_void_p()
        self._check(self.L.ona_engine_fork(self.h, C.byref(h)))
        f.h = h
        return f

    def _check(self, rc):
        if rc != 0:
            raise OnaError("ona error %d: %s" % (rc, self.L.ona_last_error().decode()))

    def close(self):
        if getattr(self, "h", None):
            self.L.ona_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing
    def set_stream(self, cuda_stream):
        self._check(self.L.ona_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self._check(self.L.ona_synchronize(self.h))

    def kernel_launches(self):
        return int(self.L.ona_kernel_launches(self.h))

    def set_island_offset(self, first_global):
        self._check(self.L.ona_set_island_offset(self.h, int(first_global)))

    def set_objectives(self, objectives):
        self.objectives = _i32(objectives)
        self.M = len(self.objectives)
        self._check(self.L.ona_set_objectives(self.h, _ptr(self.objectives), self.M))

    # ---- population
    def resize(self, P):
        """ona_pop_resize: slots below min(old, new) keep their state, new slots are empty."""
        self._check(self.L.ona_pop_resize(self.h, int(P)))
        self.P = int(P)

    def upload(self, aln, mask=None, first=0):
        """ona_pop_upload into slots first..first+len(aln); grows the population when the range
        ends beyond it (slots uploaded earlier keep their state)."""
        aln = _i32(aln).reshape(-1, self.n2)
        if first + len(aln) > self.P:
            self.resize(first + len(aln))
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8).reshape(-1, self.n2)
        self._check(self.L.ona_pop_upload(self.h, first, len(aln), _ptr(aln), _ptr(m)))

    def download(self, first=0, count=None):
        count = self.P - first if count is None else count
        a = np.empty((count, self.n2), np.int32)
        inv = np.empty((count, self.n2), np.int32)
        m = np.empty((count, self.n2), np.uint8)
        self._check(self.L.ona_pop_download(self.h, first, count, _ptr(a), _ptr(inv), _ptr(m)))
        return a, inv, m

    def set_counters(self, cons, ind, bits, goc, first=0):
        cons, ind = _i32(cons), _i32(ind)
        bits = np.ascontiguousarray(bits, dtype=np.float64)
        goc = np.ascontiguousarray(goc, dtype=np.float64)
        self._check(self.L.ona_pop_set_counters(self.h, first, len(cons), _ptr(cons), _ptr(ind),
                                                _ptr(bits), _ptr(goc)))

    def counters(self, first=0, count=None):
        count = self.P - first if count is None else count
        r = dict(cons=np.empty(count, np.int32), ind=np.empty(count, np.int32),
                 bits=np.empty(count, np.float64), goc=np.empty(count, np.float64),
                 size=np.empty(count, np.float64), fit=np.empty((count, self.M), np.float64))
        self._check(self.L.ona_pop_get_counters(self.h, first, count, _ptr(r["cons"]), _ptr(r["ind"]),
                                                _ptr(r["bits"]), _ptr(r["goc"]), _ptr(r["size"]),
                                                _ptr(r["fit"])))
        return r

    # ---- hot path
    def full_fitness(self, first=0, count=None, seed_mode=SEED_REFCOMPAT):
        count = self.P - first if count is None else count
        r = dict(cons_true=np.empty(count, np.int32), cons_ref=np.empty(count, np.int32),
                 ind=np.empty(count, np.int32), bits=np.empty(count, np.float64),
                 goc=np.empty(count, np.float64), size=np.empty(count, np.float64),
                 fit=np.empty((count, self.M), np.float64), ics=np.empty(count, np.float64))
        self._check(self.L.ona_full_fitness(self.h, first, count, seed_mode, _ptr(r["cons_true"]),
                                            _ptr(r["cons_ref"]), _ptr(r["ind"]), _ptr(r["bits"]),
                                            _ptr(r["goc"]), _ptr(r["size"]), _ptr(r["fit"]), _ptr(r["ics"])))
        return r

    def full_fitness_dev(self, first=0, count=None, seed_mode=SEED_REFCOMPAT):
        count = self.P - first if count is None else count
        self._check(self.L.ona_full_fitness_dev(self.h, first, count, seed_mode))

    def swap_deltas(self, xy, obj=-1, first=0, want=("cons", "ind", "bits", "goc", "fit", "improved",
                                                    "snd"), out=None):
        """xy: [count,K,2].  Returns per-candidate arrays (see include/ona_b200.h)."""
        xy = _i32(xy)
        count, K = xy.shape[0], xy.shape[1]
        if out is None:
            dt = dict(cons=np.int32, ind=np.int32, bits=np.float64, goc=np.float64, fit=np.float64,
                      improved=np.uint8, snd=np.float64)
            out = {k: np.empty((count, K, self.M) if k == "fit" else (count, K), dt[k]) for k in want}
        g = out.get
        self._check(self.L.ona_swap_deltas(self.h, first, count, K, _ptr(xy), int(obj), _ptr(g("cons")),
                                           _ptr(g("ind")), _ptr(g("bits")), _ptr(g("goc")),
                                           _ptr(g("fit")), _ptr(g("improved")), _ptr(g("snd"))))
        return out

    def swap_deltas_rng(self, K, seed, obj=-1, first=0, count=None, summary=True, best=False):
        count = self.P - first if count is None else count
        ni = np.empty(count, np.int32) if summary else None
        bs = np.empty(count, np.float64) if best else None
        bxy = np.empty((count, 2), np.int32) if best else None
        self._check(self.L.ona_swap_deltas_rng(self.h, first, count, int(K), int(seed), int(obj),
                                               _ptr(ni), _ptr(bs), _ptr(bxy)))
        return ni, bs, bxy

    def swap_deltas_dev(self, K, seed=0, obj=-1, first=0, count=None, xy_dev=None, cons=None,
                        ind=None, bits=None, goc=None, fit=None, improved=None, snd=None,
                        num_improved=None):
        """device-resident, asynchronous; every array argument is a raw device pointer (int)."""
        count = self.P - first if count is None else count
        vp = C.c_void_p
        self._check(self.L.ona_swap_deltas_dev(self.h, first, count, int(K), vp(xy_dev or 0), int(seed),
                                               int(obj), vp(cons or 0), vp(ind or 0), vp(bits or 0),
                                               vp(goc or 0), vp(fit or 0), vp(improved or 0),
                                               vp(snd or 0), vp(num_improved or 0)))

    def steepest_step(self, slot=0):
        """one exhaustive scan of steepestAscentHillClimb: (best swapNormalizedDelta, x, y)."""
        d, x, y = C.c_double(), C.c_int(), C.c_int()
        self._check(self.L.ona_steepest_step(self.h, int(slot), C.byref(d), C.byref(x), C.byref(y)))
        return d.value, x.value, y.value

    def rng_proposals(self, K, seed, first=0, count=None, start=0):
        count = self.P - first if count is None else count
        xy = np.empty((count, K, 2), np.int32)
        self._check(self.L.ona_rng_proposals(self.h, first, count, int(K), int(seed), int(start),
                                             _ptr(xy)))
        return xy

    def hill_climb(self, iters=None, xy=None, seed=0, objs=None, first=0, count=None,
                   want_accepted=False):
        """correctHillClimb on `count` slots; xy [count,iters,2] or None for device proposals."""
        if xy is not None:
            xy = _i32(xy)
            count, iters = xy.shape[0], xy.shape[1]
        count = self.P - first if count is None else count
        o = None if objs is None else _i32(np.broadcast_to(np.asarray(objs), (count,)))
        acc = np.empty((count, iters), np.uint8) if want_accepted else None
        nacc = np.empty(count, np.int64)
        self._check(self.L.ona_hill_climb(self.h, first, count, int(iters), _ptr(xy), int(seed),
                                          _ptr(o), _ptr(acc), _ptr(nacc)))
        return acc, nacc

    def hill_climb_lists(self, lists, objs=None, first=0):
        """ona_hill_climb_lists: one [iters,2] host array per individual."""
        arrs = [_i32(a) for a in lists]
        count, iters = len(arrs), arrs[0].shape[0]
        ptrs = (C.c_void_p * count)(*[a.ctypes.data for a in arrs])
        o = None if objs is None else _i32(np.broadcast_to(np.asarray(objs), (count,)))
        nacc = np.empty(count, np.int64)
        self._check(self.L.ona_hill_climb_lists(self.h, first, count, int(iters), ptrs, _ptr(o), None, _ptr(nacc)))
        return nacc

    def hill_climb_mt(self, iters, states, objs=None, first=0, want_accepted=False):
        """ona_hill_climb_mt: the reference's own proposal stream generated on the device from
        `states` ([count,625] uint32: mt19937 words + position); returns (accepted, num_accepted,
        states after the draws)."""
        st = np.ascontiguousarray(states, np.uint32).copy()
        count = st.shape[0]
        o = None if objs is None else _i32(np.broadcast_to(np.asarray(objs), (count,)))
        acc = np.empty((count, iters), np.uint8) if want_accepted else None
        nacc = np.empty(count, np.int64)
        self._check(self.L.ona_hill_climb_mt(self.h, first, count, int(iters), _ptr(st), _ptr(o), _ptr(acc),
                                             _ptr(nacc)))
        return acc, nacc, st

    def mt_proposals(self, n1, n2, iters, states):
        """ona_mt_proposals: (xy [count,iters,2], states after)."""
        st = np.ascontiguousarray(states, np.uint32).copy()
        count = st.shape[0]
        xy = np.empty((count, iters, 2), np.int32)
        self._check(self.L.ona_mt_proposals(self.h, int(n1), int(n2), count, int(iters), _ptr(st), _ptr(xy)))
        return xy, st

    def probe_gather_rate(self, table_bytes):
        """measurement aid: scattered 32-byte-sector reads per second from an L2-resident table."""
        r, ms = C.c_double(), C.c_double()
        self._check(self.L.ona_probe_gather_rate(self.h, int(table_bytes), C.byref(r), C.byref(ms)))
        return r.value, ms.value

    def hill_climb_batched(self, iters=None, xy=None, seed=0, objs=None, first=0, count=None):
        """throughput hill climb: rounds of 1024 proposals scored at once, a conflict-free accepted
        subset applied per round (see include/ona_b200.h).  Returns applied swaps per individual."""
        if xy is not None:
            xy = _i32(xy)
            count, iters = xy.shape[0], xy.shape[1]
        count = self.P - first if count is None else count
        o = None if objs is None else _i32(np.broadcast_to(np.asarray(objs), (count,)))
        napp = np.empty(count, np.int64)
        self._check(self.L.ona_hill_climb_batched(self.h, first, count, int(iters), _ptr(xy), int(seed),
                                                  _ptr(o), _ptr(napp)))
        return napp

    def apply_swaps(self, xy, first=0):
        xy = _i32(xy)
        self._check(self.L.ona_apply_swaps(self.h, first, xy.shape[0], xy.shape[1], _ptr(xy)))

    def apply_edits(self, ops, first=0):
        """ops [count, nops, 4] int32 = (kind, a, b, c): EDIT_SWAP / EDIT_FLIP / EDIT_ONBIT /
        EDIT_CROSS, see include/ona_b200.h (Alignment.cpp:197-252, :468-480, :491-533)."""
        ops = _i32(ops)
        self._check(self.L.ona_apply_edits(self.h, first, ops.shape[0], ops.shape[1], _ptr(ops)))

    def export_fitness_dev(self, dst_ptr, first=0, count=None):
        count = self.P - first if count is None else count
        self._check(self.L.ona_export_fitness_dev(self.h, first, count, C.c_void_p(dst_ptr)))

    def nsga2_rank(self, fitness, has_norm=None, norm=None):
        """nonDominatedSort + normalizeFitnesses + setCrowdingDists on every front
        (nsga-ii.cpp:14-118, in nsgaiimain.cpp:157-175's order) of a host [P, M] matrix."""
        f = np.ascontiguousarray(fitness, dtype=np.float64)
        P, M = f.shape
        hn = None if has_norm is None else np.ascontiguousarray(has_norm, dtype=np.uint8)
        nio = np.zeros((P, M), np.float64) if norm is None else np.array(norm, dtype=np.float64, order="C")
        rank, order, after = (np.full(P, -1, np.int32) for _ in range(3))
        offs, nf, crowd = np.zeros(P + 1, np.int32), C.c_int(0), np.zeros(P, np.float64)
        self._check(self.L.ona_nsga2_rank(self.h, P, M, _ptr(f), _ptr(hn) if hn is not None else None,
                                          _ptr(nio), _ptr(rank), _ptr(order), _ptr(offs),
                                          C.addressof(nf), _ptr(after), _ptr(crowd)))
        return {"dom_rank": rank, "front_order": order, "front_offsets": offs[:nf.value + 1].copy(),
                "order_after": after, "crowd_dist": crowd, "norm": nio}

    def nsga2_rank_dev(self, P, M, fitness_ptr, has_norm_ptr=None, norm_ptr=None, dom_rank_ptr=None,
                       front_order_ptr=None, front_offsets_ptr=None, order_after_ptr=None,
                       crowd_dist_ptr=None):
        """the same on device buffers (data_ptr()s); returns the number of fronts."""
        nf = C.c_int(0)
        self._check(self.L.ona_nsga2_rank_dev(self.h, P, M, *[C.c_void_p(p) if p else None for p in (
            fitness_ptr, has_norm_ptr, norm_ptr, dom_rank_ptr, front_order_ptr, front_offsets_ptr)],
            C.addressof(nf), *[C.c_void_p(p) if p else None for p in (order_after_ptr, crowd_dist_ptr)]))
        return nf.value


def goc_matrix(ann1, ann2, present1=None, present2=None, device=0):
    """ona_goc_matrix: loadGOC's dense Jaccard matrix (goc.cpp:16-52) from two lists of sorted term arrays,
    computed on the GPU.  present: which nodes have an entry in the reference's map (default: non-empty)."""
    L = load_library()

    def rows(ann, present):
        rp = np.zeros(len(ann) + 1, np.int32)
        rp[1:] = np.cumsum([len(a) for a in ann])
        t = np.concatenate([np.asarray(a, np.int32) for a in ann]) if rp[-1] else np.zeros(0, np.int32)
        p = np.array([len(a) > 0 for a in ann], np.uint8) if present is None else np.asarray(present, np.uint8)
        return rp, np.ascontiguousarray(t, np.int32), np.ascontiguousarray(p)
    rp1, t1, p1 = rows(ann1, present1)
    rp2, t2, p2 = rows(ann2, present2)
    out = np.empty((len(ann1), len(ann2)), np.float64)
    rc = L.ona_goc_matrix(int(device), len(ann1), len(ann2), _ptr(rp1), _ptr(t1), _ptr(p1), _ptr(rp2), _ptr(t2),
                          _ptr(p2), _ptr(out))
    if rc != 0:
        raise OnaError("ona error %d: %s" % (rc, L.ona_last_error().decode()))
    return out


def rng_proposals_host(seed, n1, n2, first_global, count, K, start=0):
    """numpy mirror of the device proposal RNG (csrc/ona_device.cuh: mix64, rng_proposal)."""
    u = np.uint64

    def mix64(z):
        z = (z + u(0x9E3779B97F4A7C15))
        z = (z ^ (z >> u(30))) * u(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> u(27))) * u(0x94D049BB133111EB)
        return z ^ (z >> u(31))

    with np.errstate(over="ignore"):
        gp = (np.arange(count, dtype=np.uint64) + u(first_global))
        key = mix64(mix64(np.full(count, seed, dtype=np.uint64)) + gp)
        i = np.arange(K, dtype=np.uint64) + u(start)
        r = mix64(key[:, None] + i[None, :])
        lo = r & u(0xFFFFFFFF)
        hi = r >> u(32)
        x = ((lo * u(n1)) >> u(32)).astype(np.int64)
        y = ((hi * u(n2 - 1)) >> u(32)).astype(np.int64)
    y = y + (y >= x)
    return np.stack([x, y], -1).astype(np.int32)
