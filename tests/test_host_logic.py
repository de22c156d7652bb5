"""Host-side logic of MembraneModel that needs no device: locator / value evaluation
(vectorised when provably equal to the per-row answer, odeSolver.py:140,177,183-187) and the
bounded mask cache."""
import numpy as np
import pytest

from knpemi_b200.odeSolver import MembraneModel, _MASK_CACHE_ENTRIES


def bare_model(n=500, strict=False, seed=0):
    m = object.__new__(MembraneModel)            # no handle: only the pure-Python helpers are used
    rng = np.random.default_rng(seed)
    m.dof_locations = rng.uniform(0, 1, (n, 3))
    m.nodes = n
    m.strict_locators = strict
    m._mask_cache = {}
    m._h = None
    return m


@pytest.mark.parametrize("strict", [False, True])
def test_locator_masks_equal_the_per_row_answer(strict):
    m = bare_model(strict=strict)
    X = m.dof_locations
    cases = [
        lambda x: x[0] < 0.3,                                   # vectorisable
        lambda x: bool(x[0] < 0.3 and x[1] > 0.5),              # raises on arrays -> per row
        lambda x: np.linalg.norm(x) < 0.9,                      # scalar for the whole matrix -> per row
        lambda x: True,                                         # constant
        lambda x: x[2] > 2.0,                                   # selects nothing
    ]
    for loc in cases:
        want = np.fromiter(map(loc, X), dtype=bool, count=len(X))
        got = m._mask(loc)
        if want.all():
            assert got is None
        else:
            assert np.array_equal(got, want)


def test_value_callables_equal_the_per_row_answer():
    m = bare_model()
    X = m.dof_locations
    rows = np.nonzero(X[:, 0] < 0.5)[0]

    class Const:                                               # dolfinx.fem.Constant stand-in
        def __float__(self):
            return 0.02

    cases = [lambda x: 1.5, lambda x: Const(), lambda x: x[0] * 2 + x[1], lambda x: float(np.sum(x)),
             lambda x: np.sum(x)]
    for f in cases:
        for sel in (None, rows):
            Xs = X if sel is None else X[sel]
            want = np.array([float(f(x)) for x in Xs])
            assert np.array_equal(m._values_of(f, sel), want)
    assert m._values_of(lambda x: 1.0, np.array([], dtype=int)).shape == (0,)


def test_mask_cache_is_bounded_and_keyed_by_identity():
    m = bare_model()
    loc = lambda x: x[0] < 0.3                                 # noqa: E731
    a = m._mask(loc)
    assert m._mask(loc) is a                                   # same callable -> cached array
    for k in range(3 * _MASK_CACHE_ENTRIES):
        m._mask(lambda x, k=k: x[0] < 0.01 * k)
    assert len(m._mask_cache) <= _MASK_CACHE_ENTRIES
    assert np.array_equal(m._mask(loc), a)                     # evicted entries are recomputed


def test_dof_sample_rows_cover_the_ends():
    m = bare_model(n=10_000)
    rows = m._sample_rows(10_000)
    assert rows[0] == 0 and rows[-1] == 9_999 and len(rows) <= 24
