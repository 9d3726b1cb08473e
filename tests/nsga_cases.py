"""Seeded objective matrices for the NSGA-II ranking tests (shared by the golden generator, the
oracle tests and the GPU parity tests).  Values mimic what computeFitness produces: ratios of
small integer counters (many exact ties, duplicates from cloned parents), continuous sums,
trade-off fronts, all-zero objectives (normalisation by max = 0, nsga-ii.cpp:71-81)."""
import numpy as np

KINDS = ["continuous", "coarse_grid", "clones", "one_active_objective", "tradeoff_negative",
         "counter_ratios", "single_front", "chain"]


def make_case(kind, P, M, seed):
    rng = np.random.default_rng(seed)
    if kind == "continuous":
        f = rng.random((P, M))
    elif kind == "coarse_grid":
        f = rng.integers(0, 6, (P, M)) / 7.0
    elif kind == "clones":
        f = rng.integers(0, 30, (P, M)) / 31.0
        src = rng.integers(0, P, max(1, P // 3))
        dst = rng.integers(0, P, max(1, P // 3))
        f[dst] = f[src]
    elif kind == "one_active_objective":
        f = np.zeros((P, M))
        f[:, 0] = rng.integers(0, 3, P)
    elif kind == "tradeoff_negative":  # last objective = -first: huge fronts, max stays 0 -> inf
        f = rng.random((P, M))
        f[:, -1] = -f[:, 0]
    elif kind == "counter_ratios":  # EC = C/E1, S3 = C/(E1 + I - C), bitscore sums
        C = rng.integers(0, 40, P).astype(np.float64)
        I = rng.integers(40, 80, P).astype(np.float64)
        cols = [C / 30000.0, C / (30000.0 + I - C), np.round(rng.random(P) * 50, 1)]
        f = np.stack([cols[m % 3] for m in range(M)], axis=1)
    elif kind == "single_front":  # points on a simplex: nobody dominates anybody
        f = rng.random((P, M))
        f /= f.sum(axis=1, keepdims=True)
        if M == 1:
            f[:] = 0.5
    elif kind == "chain":  # totally ordered: P fronts of one member
        f = np.repeat(rng.permutation(P)[:, None].astype(np.float64), M, axis=1)
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(f, dtype=np.float64)


def stale_norm(P, M, seed):
    """some individuals already carry normalised values from an earlier generation (:79)"""
    rng = np.random.default_rng(seed + 1000)
    return (rng.random(P) < 0.5).astype(np.uint8), rng.random((P, M))


def assert_same(got, want, tag=""):
    for key in ("dom_rank", "front_order", "front_offsets", "order_after"):
        assert np.array_equal(got[key], want[key]), (tag, key)
    assert np.array_equal(got["norm"], want["norm"], equal_nan=True), (tag, "norm")
    assert np.array_equal(got["crowd_dist"], want["crowd_dist"], equal_nan=True), (tag, "crowd_dist")


GOLDEN_CASES = [(kind, P, M, 100 + i, i % 3 == 2)
                for i, (kind, P, M) in enumerate([
                    ("continuous", 200, 2), ("continuous", 300, 3), ("coarse_grid", 150, 2),
                    ("coarse_grid", 240, 4), ("clones", 200, 2), ("clones", 260, 3),
                    ("one_active_objective", 90, 3), ("tradeoff_negative", 120, 2),
                    ("counter_ratios", 400, 2), ("counter_ratios", 400, 3), ("single_front", 130, 3),
                    ("chain", 40, 2), ("continuous", 1, 2), ("coarse_grid", 2, 1), ("clones", 17, 2)])]
