"""Shared comparison helpers for the parity tests."""
import math

import numpy as np

REL_TOL = 1e-6      # BASELINE.json north_star: floating-point statistics and p-values within 1e-6 relative


def close(a, b, rel=REL_TOL, abs_tol=0.0):
    """NaN == NaN, inf == inf, otherwise relative tolerance."""
    if a is None or b is None:
        return a is None and b is None
    a, b = float(a), float(b)
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= max(rel * max(abs(a), abs(b)), abs_tol)


def assert_close(a, b, what="", rel=REL_TOL, abs_tol=0.0):
    assert close(a, b, rel, abs_tol), "{}: {!r} vs {!r}".format(what, a, b)


def assert_close_list(a, b, what="", rel=REL_TOL, abs_tol=0.0):
    assert len(a) == len(b), "{}: length {} vs {}".format(what, len(a), len(b))
    for i, (x, y) in enumerate(zip(a, b)):
        assert_close(x, y, "{}[{}]".format(what, i), rel, abs_tol)
