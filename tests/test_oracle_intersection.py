"""The reference's only known-answer vectors: the six import-time asserts of
tools/intersection.py:35-57, replayed against the oracle (and, in test_gpu_ops.py, the CUDA path)."""
import numpy as np

from oracle import countline as oc


def f(x):
    return np.array(x, dtype=float)


p1, q1 = f([0, 0]), f([1, 0])
p3, q3 = f([1, 2]), f([1, 1])
# (p, pr, q, qs, expected) -- tools/intersection.py:35-49
GOLDEN_SEGMENTS = [
    (p1, q1, f([1, -1]), f([0, 1]), True),
    (p1, q1, p3, q3, False),
    (p1, q1, f([1.01, 0]), f([2, 0]), False),
    (p3, q3, f([1, 2]), f([1, 3]), True),
]
# tools/intersection.py:51-57
GOLDEN_POLYLINES = [
    (p1, q1, f([[1, 2], [1, 1], [1, -1], [1, -2]]), True),
    (p1, q1, f([[1, 2], [1, 1], [3, 1], [3, -2]]), False),
]


def test_segment_asserts():
    for p, pr, q, qs, exp in GOLDEN_SEGMENTS:
        assert oc.intersection(p, pr, q, qs) == exp


def test_polyline_asserts():
    for p, q, pts, exp in GOLDEN_POLYLINES:
        assert oc.any_intersection(p, q, pts) == exp
