"""Shared test helpers: assertion semantics of the reference's tests
(openmmapi/include/internal/AssertionUtilities.h:20-40 -- relative error with a floor of 1)."""
import numpy as np

TOL = 1e-4      # tests/TestSlicedNonbondedForce.h:27


def assert_equal_tol(expected, found, tol):
    scale = max(abs(expected), 1.0)
    assert abs(expected-found)/scale <= tol, f"expected {expected}, found {found} (tol {tol})"


def assert_equal_vec(expected, found, tol):
    expected, found = np.asarray(expected, float), np.asarray(found, float)
    norm = np.linalg.norm(expected)
    scale = norm if norm >= 1.0 else 1.0
    assert np.linalg.norm(expected-found)/scale <= tol, f"expected {expected}, found {found} (tol {tol})"


def force_rel_rms(forces, reference):
    return float(np.sqrt(((forces-reference)**2).sum()/(reference**2).sum()))
