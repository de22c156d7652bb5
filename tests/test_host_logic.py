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


def test_a_stateful_locator_is_not_served_from_the_cache():
    """The reference re-evaluates the locator at every call (odeSolver.py:100): a callable whose
    answer changes between calls must get a new mask, an unchanged one keeps its cached mask."""
    m = bare_model(n=4000)
    X = m.dof_locations
    edge = {"x": 0.3}
    loc = lambda x: x[0] < edge["x"]      # noqa: E731
    first = m._mask(loc)
    assert m._mask(loc) is first                                   # unchanged: cache hit
    edge["x"] = 0.7
    second = m._mask(loc)
    assert second is not first and np.array_equal(second, X[:, 0] < 0.7)
    edge["x"] = 2.0                                                # now selects everything
    assert m._mask(loc) is None
    edge["x"] = 0.1
    assert np.array_equal(m._mask(loc), X[:, 0] < 0.1)


def test_fixed_step_scheme_warns_about_ignored_tolerances():
    """rtol/atol only act on scheme="dp45"; with the default fixed-step scheme they are ignored
    and the constructor says so (the reference's LSODA call honours them, odeSolver.py:120)."""
    from ducks_for_tests import Space
    from knpemi_b200._cabi import KemError
    from knpemi_b200.models import hh_test
    import warnings
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        try:
            m = MembraneModel(hh_test, None, 1, Space(np.zeros((4, 3))), verbose=False, rtol=1e-6)
            m.close()
        except KemError:
            pass                                                   # no device on the CPU box
    assert any("rtol/atol are ignored" in str(x.message) for x in w)
    with pytest.raises(ValueError):
        MembraneModel(hh_test, None, 1, Space(np.zeros((4, 3))), verbose=False, unread_inputs="maybe")
    with pytest.raises(ValueError):
        MembraneModel(hh_test, None, 1, Space(np.zeros((4, 3))), verbose=False, exchange="lazy")


@pytest.mark.parametrize("taper", [0, 1])
@pytest.mark.parametrize("n", [0, 1, 1000, 65536, 300_000, 1_000_000, 10_000_000, 50_000_000, 123_456_789])
def test_chunk_plan_covers_the_range(n, taper):
    """The DOF chunks of the pipelined exchange: contiguous, complete, non-increasing; equal
    sixteenths by default, with a finely cut tail (a short pipeline drain) on request."""
    import ctypes as C
    from knpemi_b200 import _cabi
    off = (C.c_int64 * 128)()
    ln = (C.c_int64 * 128)()
    cnt = C.c_int(0)
    _cabi.check(_cabi.lib().kem_plan_chunks(n, 16, taper, off, ln, 128, C.byref(cnt)), "kem_plan_chunks")
    k = cnt.value
    if n == 0:
        assert k == 0
        return
    assert 1 <= k <= 96
    offs, lens = list(off[:k]), list(ln[:k])
    assert offs[0] == 0 and all(offs[i] + lens[i] == offs[i + 1] for i in range(k - 1))
    assert offs[-1] + lens[-1] == n and all(v > 0 for v in lens)
    assert all(lens[i] >= lens[i + 1] for i in range(k - 2))       # (the very last one is the remainder)
    if n >= 4_000_000:
        assert lens[0] >= n // 17 and max(lens) <= (n + 15) // 16 + 1024
        assert (lens[-1] <= 65536) if taper else (k == 16)


class _FakeLib:
    """Stands in for libknpemi_b200 in the host-array cache tests (no device here)."""

    def __init__(self, refuse=()):
        self.registered, self.refuse, self.log = {}, set(refuse), []

    def kem_host_register(self, ptr, nbytes):
        p = ptr.value
        if p in self.refuse:
            return -1
        self.registered[p] = self.registered.get(p, 0) + 1
        self.log.append(("reg", p))
        return 0

    def kem_host_unregister(self, ptr):
        p = ptr.value
        self.registered[p] -= 1
        if not self.registered[p]:
            del self.registered[p]
        self.log.append(("unreg", p))
        return 0

    def kem_last_error(self):
        return b"refused"


def test_host_array_cache_registers_on_the_second_sighting_of_a_living_owner(monkeypatch):
    """What the reference's callers hand over (utils.py:137-142, 190-191; run_2D.py:105-109):
    persistent getter targets are page-locked the second time they arrive, single-use trace
    Functions never are -- not even when a new array lands on a recycled address."""
    import gc
    from ducks_for_tests import Func
    from knpemi_b200 import _cabi, odeSolver
    fake = _FakeLib()
    monkeypatch.setattr(_cabi, "lib", lambda: fake)
    cache = odeSolver._HostArrayCache(max_pinned=2, min_bytes=1024)
    keep = Func(np.zeros(4096))
    assert cache.sight(keep, keep.x.array) is False            # first sighting: only noted
    assert cache.sight(keep, keep.x.array) is True             # second: registered
    assert cache.sight(keep, keep.x.array) is True and len(fake.log) == 1
    # too small, wrong dtype, non-contiguous: never
    assert cache.sight(keep, np.zeros(8)) is False
    assert cache.sight(keep, np.zeros(4096, dtype=np.float32)) is False
    assert cache.sight(keep, np.zeros(8192)[::2]) is False
    # a fresh owner per step at (possibly) the same address is a first sighting every time
    for _ in range(5):
        fresh = Func(np.zeros(4096))
        assert cache.sight(fresh, fresh.x.array) is False
        del fresh
        gc.collect()
    assert len(fake.registered) == 1
    # same memory offered by ANOTHER owner object is not the second sighting of the first
    a = np.zeros(4096)
    u1, u2 = Func(np.zeros(1)), Func(np.zeros(1))
    u1.x.array = a
    u2.x.array = a
    assert cache.sight(u1, a) is False and cache.sight(u2, a) is False and cache.sight(u2, a) is True
    # least-recently-used eviction unregisters
    third = Func(np.zeros(4096))
    cache.sight(third, third.x.array)
    assert cache.sight(third, third.x.array) is True
    assert len(cache.pinned) == 2 and ("unreg", keep.x.array.ctypes.data) in fake.log
    # a range the library refuses is remembered, not retried every step
    bad = Func(np.zeros(4096))
    fake.refuse.add(bad.x.array.ctypes.data)
    cache.sight(bad, bad.x.array)
    n_calls = len(fake.log)
    assert cache.sight(bad, bad.x.array) is False and cache.sight(bad, bad.x.array) is False
    assert len(fake.log) == n_calls
    cache.release_all()
    assert not fake.registered and not cache.pinned


def test_eliminated_ion_terms_follow_the_reference_expression():
    """update_pde_variables (utils.py:249-258): c_elim_sum = -(1/z_e) rho_z rho_tag, then
    += -(1/z_e) z_k c_k per solved ion.  The device kernels take the constant and the coefficients
    formed exactly like that."""
    from knpemi_b200.device_updates import eliminated_ion_terms
    ion_list = [{"name": "K", "z": 1.0}, {"name": "Cl", "z": -1.0}, {"name": "Na", "z": 1.0}]   # run_2D.py:252
    rho_z, rho_tag = -1.0, 41.3
    a0, coefs = eliminated_ion_terms(ion_list, rho_z, rho_tag)
    z_e = ion_list[-1]["z"]
    assert a0 == -(1.0 / z_e) * rho_z * rho_tag
    assert coefs == [-(1.0 / z_e) * 1.0, -(1.0 / z_e) * -1.0]
    rng = np.random.default_rng(0)
    c_K, c_Cl = rng.uniform(1, 150, 50), rng.uniform(1, 150, 50)
    want = a0
    for ion, c in zip(ion_list[:-1], (c_K, c_Cl)):            # the reference's loop, term by term
        want = want + -(1.0 / z_e) * ion["z"] * c
    assert np.array_equal(a0 + coefs[0] * c_K + coefs[1] * c_Cl, want)
    # electroneutrality: z_e c_elim + sum z_k c_k + rho_z rho_tag = 0
    assert np.allclose(z_e * want + c_K - c_Cl + rho_z * rho_tag, 0.0, atol=1e-12)
    with pytest.raises(Exception):
        from knpemi_b200.device_updates import _pack
        _pack([(1.0, 8)] * 9)


def test_address_is_the_data_pointer_for_every_kind_of_array():
    from knpemi_b200._cabi import address
    a = np.arange(32.0)
    ro = a.copy()
    ro.setflags(write=False)
    for arr in (a, a[3:], a[::2], ro, np.zeros(0), np.zeros((4, 3))[1]):
        assert address(arr) == arr.ctypes.data


class _RecordingLib:
    """Records what the setters, getters and the step hand to the library (no device here)."""

    def __init__(self):
        self.calls = []

    def kem_set_column(self, h, kind, col, ptr, n):
        self.calls.append(("set", kind, col, int(ptr), n))
        return 0

    def kem_get_column(self, h, kind, col, ptr, n):
        self.calls.append(("get", kind, col, int(ptr), n))
        return 0

    def kem_set_stimulus_mask(self, h, ptr, n):
        self.calls.append(("mask", None if ptr is None else int(ptr), n))
        return 0

    def kem_step(self, h, t, dt, n_sub, scheme, n_stim, cols, vals, flags):
        self.calls.append(("step", t, dt, n_sub, [cols[k] for k in range(n_stim)],
                           [vals[k] for k in range(n_stim)]))
        return 0

    def kem_last_error(self):
        return b""


def recording_model(n=300):
    import ctypes as C
    from collections import OrderedDict
    from knpemi_b200.models import hh_ideal
    m = bare_model(n)
    m.ode, m._lib, m._h = hh_ideal, _RecordingLib(), C.c_void_p(1)
    m.exchange, m.auto_register, m.verbose = "immediate", True, False
    m._pending, m._bound_out, m._prefetched = OrderedDict(), OrderedDict(), {}
    m._status_pending, m.states, m.n_sub, m._scheme_id, m.time = False, None, 25, 0, 0.0
    m._stim_mask_key, m._col_cache, m._stim_cache = "unset", {}, None
    return m


def test_the_call_sequence_hands_the_callers_own_arrays_and_current_stimulus_to_the_library():
    """The cached column lookups and stimulus arguments (what keeps a setter under 2 us) must
    not outlive a change of name, value or array."""
    from ducks_for_tests import Func
    m = recording_model()
    ode, lib = m.ode, m._lib
    u, v = Func(np.zeros(m.nodes)), Func(np.ones(m.nodes))
    for _ in range(2):
        m.set_parameter("K_e", u)
        m.set_parameter("K_e", v)
        m.get_parameter("I_ch_Na", u)
        m.set_membrane_potential(v)
    k_e, i_na, V = ode.parameter_indices("K_e"), ode.parameter_indices("I_ch_Na"), ode.state_indices("V")
    want = [("set", k_e, u.x.array.ctypes.data), ("set", k_e, v.x.array.ctypes.data),
            ("get", i_na, u.x.array.ctypes.data), ("set", V, v.x.array.ctypes.data)] * 2
    assert [(c[0], c[2], c[3]) for c in lib.calls] == want
    assert {c[1] for c in lib.calls if c[2] == V} != {c[1] for c in lib.calls if c[2] == k_e}   # state vs parameter
    with pytest.raises(ValueError):
        m.set_parameter("no_such_parameter", u)
    with pytest.raises(ValueError):
        m.get_state("no_such_state", u)

    lib.calls.clear()
    loc = lambda x: x[0] < 0.4                                  # noqa: E731
    m.step_lsoda(0.1, {"stim_amplitude": 10.0}, loc)
    m.step_lsoda(0.1, {"stim_amplitude": 10.0}, loc)            # same arguments: cached, mask not re-sent
    m.step_lsoda(0.1, {"stim_amplitude": 7.5}, loc)             # new value
    m.step_lsoda(0.1, {"g_K_bar": 7.5}, loc)                    # new name
    m.step_lsoda(0.1, {"stim_amplitude": 1.0, "g_K_bar": 2.0}, None)
    m.step_lsoda(0.1, None)
    steps = [c for c in lib.calls if c[0] == "step"]
    amp, gk = ode.parameter_indices("stim_amplitude"), ode.parameter_indices("g_K_bar")
    assert [(s[4], s[5]) for s in steps] == [([amp], [10.0]), ([amp], [10.0]), ([amp], [7.5]), ([gk], [7.5]),
                                             ([amp, gk], [1.0, 2.0]), ([], [])]
    assert [round(s[1], 12) for s in steps] == [0.0, 0.1, 0.2, 0.3, 0.4, 0.5]
    masks = [c for c in lib.calls if c[0] == "mask"]
    assert len(masks) == 2 and masks[0][1] is not None and masks[1][1] is None   # locator once, then "every row"
    with pytest.raises(ValueError):
        m.step_lsoda(0.1, {"no_such_parameter": 1.0})
