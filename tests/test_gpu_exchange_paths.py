"""The host<->device exchange around the step: residency policies for columns the right-hand
side never touches, literal outputs, validation order of kem_step_io, page-locked ranges, the
deferred (pipelined) drop-in sequence.  Every path must give the values the restated reference
class (oracle/membrane_oracle.py, after src/knpemi/odeSolver.py:130-166) gives."""
import gc

import numpy as np
import pytest

from ducks_for_tests import Func, Space
from test_gpu_api import close, make_pair
from workloads import SETUP, builtin, load_tables, synthetic_tables

pytestmark = pytest.mark.gpu

IONS_IN = ("K_e", "K_i", "Na_e", "Na_i", "Cl_e", "Cl_i")
IONS_OUT = ("I_ch_Na", "I_ch_K", "I_ch_Cl")


def _model(name, n, seed=3, **kw):
    from knpemi_b200.odeSolver import MembraneModel
    S, P, X, mask = synthetic_tables(name, n, seed=seed)
    m = MembraneModel(builtin(name), None, 1, Space(X), verbose=False, devices=kw.pop("devices", [0]), **kw)
    load_tables(m, S, P)
    return m, S, P, X


def _pinned(src):
    from knpemi_b200._cabi import pinned_empty
    a = pinned_empty(len(src))
    a[:] = src
    return a


@pytest.mark.parametrize("policy", ["shadow", "upload", "discard"])
def test_unread_input_policies_do_not_change_the_step(built, policy):
    """Cl_e / Cl_i are pushed every PDE step (utils.py:227-228) and never read by the HH
    right-hand side: wherever the policy puts them, states and currents are bitwise the same;
    only `discard` refuses to give them back."""
    from knpemi_b200._cabi import KemError
    name, n = "hh_ideal", 70_001
    ref, S, P, X = _model(name, n, unread_inputs="upload")
    m, _, _, _ = _model(name, n, unread_inputs=policy)
    rng = np.random.default_rng(11)
    loc = lambda x: x[0] < 20e-6      # noqa: E731
    for step in range(2):
        ins = {("parameter", k): _pinned(P[:, builtin(name).parameter_indices(k)] * (1 + 0.01 * rng.uniform(-1, 1, n)))
               for k in IONS_IN}
        ins[("state", "V")] = _pinned(S[:, 3] * (1 + 1e-3 * step))
        out = {mm: {("state", "V"): _pinned(np.zeros(n)), **{("parameter", k): _pinned(np.ones(n)) for k in IONS_OUT}}
               for mm in (ref, m)}
        for mm in (ref, m):
            mm.step_exchange(1e-4, ins, out[mm], {"stim_amplitude": 10.0}, loc)
        for key in out[ref]:
            assert np.array_equal(out[ref][key], out[m][key]), (policy, step, key)
    assert np.array_equal(np.asarray(ref.states), np.asarray(m.states))
    where = {"shadow": "host", "upload": "device", "discard": "discarded"}[policy]
    assert m.column_location("parameter", "Cl_e") == where
    got = Func(np.zeros(n))
    if policy == "discard":
        with pytest.raises(KemError, match="discarded"):
            m.get_parameter("Cl_e", got)
        # the plain setter follows the policy too, and any other policy brings the column back
        m.set_parameter("Cl_i", Func(np.full(n, 7.0)))
        assert m.column_location("parameter", "Cl_i") == "discarded"
        with pytest.raises(KemError, match="discarded"):
            m.step_exchange(1e-4, {}, {("parameter", "Cl_i"): _pinned(np.zeros(n))})
    else:
        m.get_parameter("Cl_e", got)
        assert np.array_equal(got.x.array, ins[("parameter", "Cl_e")])
    ref.close()
    m.close()


def test_literal_outputs_are_filled_on_the_host(built):
    """I_ch_Cl = 0.0 is a literal in the HH right-hand sides (mm_hh.py:225): after a step the
    getter and the exchange fill it without a device copy; a caller's write to the slot is
    honoured until the next step stores the literal again."""
    name, n = "hh_tissue", 50_021
    m, S, P, X = _model(name, n)
    col = builtin(name).parameter_indices("I_ch_Cl")
    m.set_parameter("I_ch_Cl", Func(np.full(n, 3.5)))            # before any step: an ordinary column
    got = Func(np.ones(n))
    m.get_parameter("I_ch_Cl", got)
    assert np.all(got.x.array == 3.5)
    m.step_lsoda(0.1, None)
    m.get_parameter("I_ch_Cl", got)
    assert np.all(got.x.array == 0.0) and not np.signbit(got.x.array).any()
    m.set_parameter("I_ch_Cl", Func(np.full(n, -2.0)))           # written again by the caller
    m.get_parameter("I_ch_Cl", got)
    assert np.all(got.x.array == -2.0)
    out = {("parameter", "I_ch_Cl"): _pinned(np.ones(n)), ("parameter", "I_ch_K"): _pinned(np.ones(n))}
    m.step_exchange(0.1, {("parameter", "I_ch_Cl"): _pinned(np.full(n, 9.0))}, out)
    assert np.all(out[("parameter", "I_ch_Cl")] == 0.0)          # the step stores after the input landed
    assert np.array_equal(np.asarray(m.parameters)[:, col], np.zeros(n))
    assert np.any(out[("parameter", "I_ch_K")] != 1.0)
    m.close()


def test_exchange_validates_everything_before_it_changes_anything(built):
    """A bad argument late in the call must leave the tables as they were (the first version
    flipped input columns to per-DOF storage over uninitialised memory before it noticed)."""
    from knpemi_b200._cabi import KEM_PARAM, KemError, check, kem_io_column
    import ctypes as C
    name, n = "hh_ideal", 20_011
    m, S, P, X = _model(name, n)
    ode = builtin(name)
    cm = ode.parameter_indices("Cm")
    assert m.column_location("parameter", "Cm") == "uniform"
    good = _pinned(np.full(n, 0.03))
    a_in = (kem_io_column * 1)()
    a_in[0].kind, a_in[0].col, a_in[0].host = KEM_PARAM, cm, good.ctypes.data
    a_out = (kem_io_column * 1)()
    a_out[0].kind, a_out[0].col, a_out[0].host = KEM_PARAM, 10_000, good.ctypes.data     # bad column
    flags = C.c_int(0)
    rc = m._lib.kem_step_io(m._h, 0.0, 1e-4, 25, 0, 0, None, None, 1, a_in, 1, a_out, C.byref(flags), None)
    assert rc < 0
    assert m.column_location("parameter", "Cm") == "uniform"     # untouched
    got = Func(np.zeros(n))
    m.get_parameter("Cm", got)
    assert np.all(got.x.array == SETUP[name]["uniform"]["Cm"])
    before = np.asarray(m.states).copy()
    with pytest.raises(KemError):
        m.step_exchange(1e-4, {("parameter", "Cm"): good}, {("state", "V"): np.zeros(n - 1)})   # too short
    assert np.array_equal(np.asarray(m.states), before)
    m.close()


def test_exchange_orders_inputs_before_the_stimulus(built):
    """Reference order: setters first (utils.py:227-233), then step_lsoda writes the stimulus
    (odeSolver.py:110-112).  An unmasked stimulus therefore overrides an input to the same
    column, a masked one overrides it on the masked rows -- also for a column the right-hand
    side never reads, whatever the residency policy."""
    name, n = "hh_ideal", 30_011
    for policy in ("shadow", "discard", "upload"):
        gpu, cpu, X, rng = make_pair(name, n, unread_inputs=policy)
        cfg = SETUP[name]
        for mm in (gpu, cpu):
            for k, v in {**cfg["uniform"], **cfg["varying"]}.items():
                mm.set_parameter_values({k: lambda x, v=v: v})
        amp = 0.5 + rng.uniform(size=n)
        cl = 100.0 + rng.uniform(size=n)
        loc = lambda x: x[0] < 20e-6      # noqa: E731
        # masked stimulus on the amplitude AND on a dead column, both also inputs of the call
        stim = {"stim_amplitude": 10.0, "Cl_e": -1.0}
        outs = {("parameter", "stim_amplitude"): _pinned(np.zeros(n)), ("parameter", "Cl_e"): _pinned(np.zeros(n)),
                ("state", "V"): _pinned(np.zeros(n))}
        gpu.step_exchange(cfg["dt"], {("parameter", "stim_amplitude"): _pinned(amp), ("parameter", "Cl_e"): _pinned(cl)},
                          outs, stim, loc)
        cpu.set_parameter("stim_amplitude", Func(amp))
        cpu.set_parameter("Cl_e", Func(cl))
        cpu.step_lsoda(cfg["dt"], stim, loc)
        for key, c in (("stim_amplitude", 8), ("Cl_e", 13)):
            assert np.array_equal(outs[("parameter", key)], cpu.parameters[:, c]), (policy, key)
        assert close(outs[("state", "V")], cpu.states[:, 3])
        # unmasked: every row gets the stimulus, the input is irrelevant
        gpu.step_exchange(cfg["dt"], {("parameter", "stim_amplitude"): _pinned(amp)}, outs, {"stim_amplitude": 2.0}, None)
        cpu.set_parameter("stim_amplitude", Func(amp))
        cpu.step_lsoda(cfg["dt"], {"stim_amplitude": 2.0}, None)
        assert np.all(outs[("parameter", "stim_amplitude")] == 2.0)
        assert gpu.column_location("parameter", "stim_amplitude") == "uniform"
        assert close(np.asarray(gpu.states), cpu.states)
        gpu.close()


def test_page_locked_ranges_are_checked_end_to_end(built):
    """cudaHostRegister pins whole pages.  A small array that starts inside a page a registered
    neighbour pinned is NOT page-locked beyond that page: it must take the staged path (the
    first-byte test of the first version sent it down the DMA path)."""
    from knpemi_b200 import _cabi
    from knpemi_b200._cabi import KemError, check, host_is_pinned
    import ctypes as C
    lib = _cabi.lib()
    page = 4096
    arena = np.zeros(64 * page // 8 + 1024)                     # one allocation, carved by hand
    base = (arena.ctypes.data + page - 1) // page * page
    off = (base - arena.ctypes.data) // 8
    big = arena[off: off + 8 * page // 8 + 100]                  # ends 800 bytes into its 9th page
    small = arena[off + len(big): off + len(big) + 4 * page // 8]   # starts in that page, runs on for 4 pages
    assert not host_is_pinned(big) and not host_is_pinned(small)
    check(lib.kem_host_register(C.c_void_p(big.ctypes.data), big.nbytes), "kem_host_register")
    assert host_is_pinned(big)
    assert host_is_pinned(big[10:200])                           # a view inside the range
    assert not host_is_pinned(small)                             # shares a pinned page, tail pageable
    # reference counting: a second registration needs a second release
    check(lib.kem_host_register(C.c_void_p(big.ctypes.data), big.nbytes), "kem_host_register")
    check(lib.kem_host_unregister(C.c_void_p(big.ctypes.data)), "kem_host_unregister")
    assert host_is_pinned(big)
    # a view INSIDE the registered range shares the owner's page-lock: registering it counts a
    # reference on the owner, unregistering it gives that reference back
    inner = big[16:400]
    check(lib.kem_host_register(C.c_void_p(inner.ctypes.data), inner.nbytes), "kem_host_register")
    check(lib.kem_host_unregister(C.c_void_p(big.ctypes.data)), "kem_host_unregister")
    assert host_is_pinned(big)                                   # still held through the view
    check(lib.kem_host_register(C.c_void_p(big.ctypes.data), big.nbytes), "kem_host_register")
    check(lib.kem_host_unregister(C.c_void_p(inner.ctypes.data)), "kem_host_unregister")
    assert host_is_pinned(big)
    # overlapping a registered range from another base is refused, not swallowed
    with pytest.raises(KemError, match="overlap"):
        check(lib.kem_host_register(C.c_void_p(big.ctypes.data + 8 * 50), big.nbytes), "kem_host_register")
    check(lib.kem_host_unregister(C.c_void_p(big.ctypes.data)), "kem_host_unregister")
    assert not host_is_pinned(big)
    # and the data path agrees: the same values through either path
    n = len(small)
    m, S, P, X = _model("hh_test", n)
    small[:] = np.linspace(-80.0, -60.0, n)
    check(lib.kem_host_register(C.c_void_p(big.ctypes.data), big.nbytes), "kem_host_register")
    u = Func(np.zeros(1))
    u.x.array = small
    m.set_membrane_potential(u)
    back = Func(np.zeros(n))
    m.get_membrane_potential(back)
    assert np.array_equal(back.x.array, small)
    check(lib.kem_host_unregister(C.c_void_p(big.ctypes.data)), "kem_host_unregister")
    m.close()


def test_two_models_share_one_registration(built):
    """register_host_array from two models on one array: closing the first must not unpin
    memory the second still uses (process-wide reference count)."""
    from knpemi_b200._cabi import host_is_pinned
    n = 40_003
    a, S, P, X = _model("hh_test", n, auto_register=False)
    b, _, _, _ = _model("hh_test", n, auto_register=False)
    u = Func(np.linspace(-80.0, -60.0, n))
    a.register_host_array(u)
    b.register_host_array(u)
    a.close()
    assert host_is_pinned(u.x.array)
    b.set_membrane_potential(u)
    back = Func(np.zeros(n))
    b.get_membrane_potential(back)
    assert np.array_equal(back.x.array, u.x.array)
    b.close()
    assert not host_is_pinned(u.x.array)


def test_arrays_that_come_back_are_page_locked_automatically(built):
    """The getter targets of solve_odes are the same Functions every step (run_2D.py:105-109):
    registered on their second sighting.  Arrays seen once (the fresh trace Functions of
    utils.py:190-191) never are -- not even when a new array lands on a recycled address."""
    from knpemi_b200 import odeSolver
    from knpemi_b200._cabi import host_is_pinned
    n = 60_007                                                  # 480 kB: above the cache's threshold
    m, S, P, X = _model("hh_tissue", n)
    cache = odeSolver._HOST_CACHE
    cache.release_all()
    phi, i_na = Func(S[:, 3].copy()), Func(np.zeros(n))
    for step in range(3):
        fresh = Func(P[:, 9] * (1 + 1e-3 * step))               # a new Function per step
        addr = fresh.x.array.ctypes.data
        m.set_parameter("K_e", fresh)
        assert not host_is_pinned(fresh.x.array)
        m.set_membrane_potential(phi)
        m.step_lsoda(0.1, None)
        m.get_membrane_potential(phi)
        m.get_parameter("I_ch_Na", i_na)
        del fresh
        gc.collect()
        assert host_is_pinned(phi.x.array)                       # setter + getter: seen twice in step 0
        assert host_is_pinned(i_na.x.array) == (step >= 1)       # one getter per step: second step
    assert len(cache.pinned) == 2
    # values are the same as without the cache
    ref, _, _, _ = _model("hh_tissue", n, auto_register=False)
    phi2 = Func(S[:, 3].copy())
    for step in range(3):
        ref.set_parameter("K_e", Func(P[:, 9] * (1 + 1e-3 * step)))
        ref.set_membrane_potential(phi2)
        ref.step_lsoda(0.1, None)
        ref.get_membrane_potential(phi2)
    assert np.array_equal(phi.x.array, phi2.x.array)
    cache.release_all()
    assert not host_is_pinned(phi.x.array)
    m.close()
    ref.close()


def test_deferred_exchange_is_bitwise_the_immediate_one(built):
    """exchange="deferred": setters record, step_lsoda enqueues the pipelined exchange, the
    first getter follows the kernel chunk by chunk.  Same bits as the default sequence; a failed
    integration is reported by that getter."""
    from knpemi_b200 import odeSolver
    name, n = "hh_ideal", 1_200_011                            # several tapered chunks
    odeSolver._HOST_CACHE.release_all()
    a, S, P, X = _model(name, n, exchange="immediate")
    b, _, _, _ = _model(name, n, exchange="deferred")
    ode = builtin(name)
    loc = lambda x: x[0] < 20e-6      # noqa: E731
    stim = {"stim_amplitude": 10.0}
    res, arrays = {}, {}
    for mm in (a, b):
        rng = np.random.default_rng(4)
        u_in = {k: Func(P[:, ode.parameter_indices(k)].copy()) for k in IONS_IN}
        phi = Func(S[:, 3].copy())
        u_out = {k: Func(np.ones(n)) for k in IONS_OUT}
        arrays[mm] = u_in
        for step in range(4):                                   # arrays are registered from step 1 on
            for k, u in u_in.items():
                u.x.array[:] = P[:, ode.parameter_indices(k)] * (1 + 0.01 * rng.uniform(-1, 1, n))
                mm.set_parameter(k, u)
            if step > 0:
                mm.set_membrane_potential(phi)
            mm.step_lsoda(1e-4, stim, loc)
            mm.get_membrane_potential(phi)
            for k, u in u_out.items():
                mm.get_parameter(k, u)
        res[mm] = (phi.x.array.copy(), {k: u.x.array.copy() for k, u in u_out.items()}, np.asarray(mm.states))
    assert np.array_equal(res[a][0], res[b][0])
    for k in IONS_OUT:
        assert np.array_equal(res[a][1][k], res[b][1][k]), k
    assert np.array_equal(res[a][2], res[b][2])
    # a recorded write is visible to everything else that reads the tables
    u = arrays[b]["K_e"]                                        # page-locked by now: only recorded
    u.x.array[:] = 3.25
    b.set_parameter("K_e", u)
    assert len(b._pending) == 1
    got = Func(np.zeros(n))
    b.get_parameter("K_e", got)
    assert np.all(got.x.array == 3.25) and not b._pending
    # failure surfaces at the getter that follows the enqueue-only step
    bad = Func(np.full(n, np.nan))
    b.set_membrane_potential(bad)
    b.set_membrane_potential(bad)                               # second sighting: page-locked, so recorded
    b.step_lsoda(1e-4, None)
    with pytest.raises(AssertionError):
        b.get_membrane_potential(got)
    a.close()
    b.close()
    odeSolver._HOST_CACHE.release_all()


def test_chunked_step_and_getter_overlap_change_nothing(built):
    from knpemi_b200._cabi import check, pinned_empty
    name, n = "hh_tissue", 900_017
    a, S, P, X = _model(name, n)
    b, _, _, _ = _model(name, n)
    check(b._lib.kem_set_step_chunks(b._h, 16), "kem_set_step_chunks")
    loc = lambda x: x[0] < 20e-6      # noqa: E731
    va, vb = pinned_empty(n), pinned_empty(n)
    for step in range(3):
        a.step_lsoda(0.1, {"stim_amplitude": 5.0}, loc)
        b.step_async(0.1, {"stim_amplitude": 5.0}, loc)
        ua, ub = Func(np.zeros(1)), Func(np.zeros(1))
        ua.x.array, ub.x.array = va, vb
        a.get_membrane_potential(ua)
        b.get_membrane_potential(ub)                            # follows the chunks on the copy stream
        assert np.array_equal(va, vb)
    assert np.array_equal(np.asarray(a.parameters), np.asarray(b.parameters))
    a.close()
    b.close()


def test_a_locator_that_changes_its_mind_is_evaluated_again(built):
    """The reference evaluates the locator at every call (odeSolver.py:100); the mask cache may
    not serve a stale mask when the callable's answer changes."""
    name, n = "hh_test", 5_003
    gpu, cpu, X, rng = make_pair(name, n)
    edge = {"x": 20e-6}
    loc = lambda x: x[0] < edge["x"]      # noqa: E731
    for k, new_edge in enumerate((20e-6, 20e-6, 45e-6, 5e-6)):
        edge["x"] = new_edge
        for mm in (gpu, cpu):
            mm.step_lsoda(0.1, {"stim_amplitude": 0.25 * (k + 1)}, loc)
        assert np.array_equal(np.asarray(gpu.parameters)[:, 7], cpu.parameters[:, 7]), k
    assert close(np.asarray(gpu.states), cpu.states)
    gpu.close()


def test_every_device_of_the_box_in_one_handle(built):
    """kem_create(n_dev > 1) over DISTINCT devices: per-device streams, device hopping in the
    exchange, bitwise the one-device result.  Runs on one device (trivially) when the box has one."""
    from knpemi_b200 import _cabi
    devs = list(range(_cabi.device_count()))
    name, n = "hh_ideal", 400_009
    one, S, P, X = _model(name, n, devices=[0])
    many, _, _, _ = _model(name, n, devices=devs)
    assert [d for d, _, _ in many.shard_ranges()] == devs
    loc = lambda x: x[0] < 20e-6      # noqa: E731
    ode = builtin(name)
    rng = np.random.default_rng(8)
    for step in range(3):
        ins = {("parameter", k): _pinned(P[:, ode.parameter_indices(k)] * (1 + 0.01 * rng.uniform(-1, 1, n)))
               for k in IONS_IN}
        ins[("state", "V")] = _pinned(S[:, 3])
        outs = {mm: {("state", "V"): _pinned(np.zeros(n)), ("parameter", "I_ch_Na"): _pinned(np.zeros(n))}
                for mm in (one, many)}
        for mm in (one, many):
            mm.step_exchange(1e-4, ins, outs[mm], {"stim_amplitude": 10.0}, loc)
            mm.step_lsoda(1e-4, {"stim_amplitude": 10.0}, loc)
        for key in outs[one]:
            assert np.array_equal(outs[one][key], outs[many][key])
    assert np.array_equal(np.asarray(one.states), np.asarray(many.states))
    assert np.array_equal(np.asarray(one.parameters), np.asarray(many.parameters))
    one.close()
    many.close()


def test_pipeline_tuning_does_not_change_results(built):
    """kem_set_io_tuning (chunks per device, copy streams) and the taper only reshape the
    pipeline: every setting gives the same bits."""
    import ctypes as C
    from knpemi_b200._cabi import check
    name, n = "hh_ideal", 1_500_003
    ode = builtin(name)
    ref = None
    for chunks, streams in ((0, 0), (1, 1), (4, 2), (32, 1), (48, 2)):
        m, S, P, X = _model(name, n)
        check(m._lib.kem_set_io_tuning(m._h, chunks, streams), "kem_set_io_tuning")
        rng = np.random.default_rng(2)
        ins = {("parameter", k): _pinned(P[:, ode.parameter_indices(k)] * (1 + 0.01 * rng.uniform(-1, 1, n)))
               for k in IONS_IN}
        ins[("state", "V")] = _pinned(S[:, 3])
        outs = {("state", "V"): _pinned(np.zeros(n)), **{("parameter", k): _pinned(np.ones(n)) for k in IONS_OUT}}
        for _ in range(2):
            m.step_exchange(1e-4, ins, outs, {"stim_amplitude": 10.0}, lambda x: x[0] < 20e-6)
        got = {k: np.array(v) for k, v in outs.items()}
        got["states"] = np.asarray(m.states)
        m.close()
        if ref is None:
            ref = got
        else:
            for k in ref:
                assert np.array_equal(ref[k], got[k]), (chunks, streams, k)
    cnt = C.c_int(0)
    from knpemi_b200 import _cabi
    assert _cabi.lib().kem_set_io_tuning(None, 0, 0) < 0            # null handle is an argument error
    _cabi.check(_cabi.lib().kem_plan_chunks(n, 48, 0, None, None, 0, C.byref(cnt)), "kem_plan_chunks")
    assert cnt.value == -(-n // 131072)        # chunks are never smaller than 2 x 64k DOFs
