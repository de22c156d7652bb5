"""MembraneModel drop-in behaviour on the GPU against the CPU restatement of the
reference class (oracle/membrane_oracle.py follows src/knpemi/odeSolver.py:6-189)."""
import os
import textwrap

import numpy as np
import pytest

from ducks_for_tests import FloatLike, Func, Space
from workloads import SETUP, builtin, load_tables, synthetic_tables

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def make_pair(name, n, seed=5, **kw):
    from knpemi_b200.odeSolver import MembraneModel
    from oracle.membrane_oracle import OracleMembraneModel
    ode = builtin(name)
    rng = np.random.default_rng(seed)
    X = rng.uniform(0.0, 62e-6, (n, 3))
    gpu = MembraneModel(ode, None, 1, Space(X), verbose=False, devices=kw.pop("devices", [0]), **kw)
    cpu = OracleMembraneModel(ode, None, 1, Space(X), oracle_name=name)
    return gpu, cpu, X, rng


def close(a, b, rtol=RTOL):
    scale = np.maximum(np.abs(b), 1e-3 * np.max(np.abs(b), axis=0, keepdims=True) + 1e-300)
    return float(np.max(np.abs(a - b) / scale)) < rtol


def test_solve_odes_call_sequence(built):
    """setup_membrane_model (utils.py:121-141) + update_ode_variables (utils.py:217-233)
    + step + getters (run_2D.py:98-109), repeated like the time loop does."""
    name, n = "hh_ideal", 3001
    gpu, cpu, X, rng = make_pair(name, n)
    cfg = SETUP[name]
    ions = [("K", 1.0), ("Cl", -1.0), ("Na", 1.0)]
    for m in (gpu, cpu):
        m.set_parameter_values({'Cm': lambda x: FloatLike(cfg["uniform"]["Cm"])})
        m.set_parameter_values({'psi': lambda x: FloatLike(cfg["uniform"]["psi"])})
        for ion, z in ions:
            m.set_parameter_values({f'z_{ion}': lambda x, z=z: z})
    I_ch = {m: {ion: Func(np.zeros(n)) for ion, _ in ions} for m in (gpu, cpu)}
    for m in (gpu, cpu):
        for ion, _ in ions:
            assert m.get_parameter("I_ch_" + ion, I_ch[m][ion]) is I_ch[m][ion]
    phi = {gpu: Func(np.zeros(n)), cpu: Func(np.zeros(n))}
    stim = {'stim_amplitude': cfg["stim"]}
    locator = lambda x: (x[0] < 20e-6)      # noqa: E731
    for k in range(4):
        traces = {f"{ion}_{side}": cfg["varying"][f"{ion}_{side}"] * (1 + 0.01 * rng.uniform(-1, 1, n))
                  for ion, _ in ions for side in ("e", "i")}
        for m in (gpu, cpu):
            for key, arr in traces.items():
                ret = m.set_parameter(key, Func(arr))
                assert ret is m.states
            if k > 0:
                m.set_membrane_potential(phi[m])
            ret = m.step_lsoda(dt=cfg["dt"], stimulus=stim, stimulus_locator=locator)
            assert ret is m.states
            m.get_membrane_potential(phi[m])
            for ion, _ in ions:
                m.get_parameter("I_ch_" + ion, I_ch[m][ion])
        assert close(phi[gpu].x.array, phi[cpu].x.array)
        for ion, _ in ions:
            assert close(I_ch[gpu][ion].x.array, I_ch[cpu][ion].x.array)
    assert gpu.time == cpu.time
    assert close(np.asarray(gpu.states), cpu.states)
    assert close(np.asarray(gpu.parameters), cpu.parameters)
    gpu.close()


def test_setters_getters_with_locators_are_exact(built):
    """Pure data movement: bit-identical tables after any mix of (masked) sets/gets."""
    gpu, cpu, X, rng = make_pair("hh_tissue", 777)
    n = 777
    loc_a = lambda x: x[0] < 30e-6                  # vectorisable           # noqa: E731
    loc_b = lambda x: bool(x[1] > 10e-6 and x[2] < 50e-6)   # only works per row   # noqa: E731
    loc_none = lambda x: False                      # selects nothing         # noqa: E731
    u1, u2 = rng.normal(size=n), rng.normal(size=n + 5)     # u may be longer than N
    for m in (gpu, cpu):
        m.set_state('V', Func(u1))
        m.set_state('m', Func(u2), locator=loc_a)
        m.set_parameter('K_e', Func(u2), locator=loc_b)
        m.set_parameter('Na_i', Func(u1), locator=loc_none)
        m.set_parameter_values({'Cm': lambda x: 1.0 + x[0], 'psi': lambda x: 0.04}, locator=loc_a)
        m.set_state_values({'h': lambda x: 0.5}, locator=loc_b)
        m.set_state_values({'n': lambda x: x[1] * 1e3})
    assert np.array_equal(np.asarray(gpu.states), cpu.states)
    assert np.array_equal(np.asarray(gpu.parameters), cpu.parameters)
    for which, loc in (('V', None), ('m', loc_a), ('h', loc_b), ('n', loc_none)):
        a, b = Func(np.full(n + 3, 9.0)), Func(np.full(n + 3, 9.0))
        gpu.get_state(which, a, locator=loc)
        cpu.get_state(which, b, locator=loc)
        assert np.array_equal(a.x.array, b.x.array)
    a, b = Func(np.zeros(n)), Func(np.zeros(n))
    gpu.get_parameter('Cm', a, locator=loc_b)
    cpu.get_parameter('Cm', b, locator=loc_b)
    assert np.array_equal(a.x.array, b.x.array)
    gpu.close()


def test_strict_locators_give_the_same_masks(built):
    gpu, cpu, X, rng = make_pair("glial_bench", 400, strict_locators=True)
    tricky = lambda x: np.linalg.norm(x) < 60e-6     # scalar for a whole matrix   # noqa: E731
    from knpemi_b200.odeSolver import MembraneModel
    lax = MembraneModel(builtin("glial_bench"), None, 1, Space(X), verbose=False, devices=[0])
    for m in (gpu, lax, cpu):
        m.set_state_values({'V': lambda x: -80.0}, locator=tricky)
    assert np.array_equal(np.asarray(gpu.states), cpu.states)
    assert np.array_equal(np.asarray(lax.states), cpu.states)
    gpu.close()
    lax.close()


def test_table_views(built):
    gpu, cpu, X, rng = make_pair("calibration", 11)
    assert gpu.states.shape == (11, 14) and gpu.parameters.shape == (11, 13)
    assert len(gpu.states) == 11
    idx = builtin("calibration").state_indices('K_e')
    col = 1 * gpu.states[:, idx]                       # run_calibration.py:70
    assert np.array_equal(col, cpu.states[:, idx])
    assert gpu.states[2, idx] == cpu.states[2, idx]
    gpu.states[:, idx] = np.arange(11.0)
    gpu.states[3:5, 0:2] = 0.25
    cpu.states[:, idx] = np.arange(11.0)
    cpu.states[3:5, 0:2] = 0.25
    assert np.array_equal(np.asarray(gpu.states), cpu.states)
    assert np.array_equal(gpu.states[1:4], cpu.states[1:4])
    with pytest.raises(IndexError):
        gpu.states[:, 14]
    with pytest.raises(ValueError):
        gpu.V_index                                    # calibration has V_n / V_g, no V
    gpu.close()


def test_unknown_names_raise_value_error(built):
    gpu, cpu, X, rng = make_pair("hh_test", 8)
    with pytest.raises(ValueError):
        gpu.set_parameter('nope', Func(np.zeros(8)))
    with pytest.raises(ValueError):
        gpu.step_lsoda(0.1, {'nope': 1.0})
    with pytest.raises(ValueError):
        gpu.set_state_values({'nope': lambda x: 0.0})
    assert gpu.V_index == 3
    with pytest.raises(AssertionError):
        from knpemi_b200.odeSolver import MembraneModel
        MembraneModel(builtin("hh_test"), None, "1", Space(X), verbose=False)   # tag must be int (:13)
    gpu.close()


def test_stimulus_is_sticky_and_time_accumulates(built):
    """odeSolver.py:108-112: the stimulus value stays in the table; :123 time += dt."""
    gpu, cpu, X, rng = make_pair("hh_tissue", 500)
    cfg = SETUP["hh_tissue"]
    for m in (gpu, cpu):
        for k, v in {**cfg["uniform"], **cfg["varying"]}.items():
            m.set_parameter_values({k: lambda x, v=v: v})
    loc = lambda x: x[0] < 20e-6     # noqa: E731
    c = builtin("hh_tissue").parameter_indices("stim_amplitude")
    for m in (gpu, cpu):
        m.step_lsoda(0.1, {'stim_amplitude': 5.0}, loc)
        m.step_lsoda(0.1, None)                        # stimulus=None -> {} (:94)
        m.step_lsoda(0.1, {}, loc)
    assert np.array_equal(gpu.parameters[:, c], cpu.parameters[:, c])
    assert set(np.unique(gpu.parameters[:, c])) == {0.0, 5.0}
    assert gpu.time == cpu.time == pytest.approx(0.3)
    assert close(np.asarray(gpu.states), cpu.states)
    # unmasked stimulus afterwards overwrites every row
    for m in (gpu, cpu):
        m.step_lsoda(0.1, {'stim_amplitude': 1.5})
    assert np.all(gpu.parameters[:, c] == 1.5)
    assert close(np.asarray(gpu.states), cpu.states)
    gpu.close()


def test_stimulus_discontinuities_in_time(built):
    """i_Stim jumps at mod(t, 30) = 0 and at t = 125 (mm_hh.py:182): stage times are
    formed identically on both sides, so steps across the jumps still agree."""
    gpu, cpu, X, rng = make_pair("hh_tissue", 300)
    cfg = SETUP["hh_tissue"]
    for m in (gpu, cpu):
        for k, v in {**cfg["uniform"], **cfg["varying"]}.items():
            m.set_parameter_values({k: lambda x, v=v: v})
        m.time = 119.7
    for _ in range(60):                                # crosses t = 120 and t = 125
        for m in (gpu, cpu):
            m.step_lsoda(0.1, {'stim_amplitude': 4.0})
    assert gpu.time == cpu.time
    assert close(np.asarray(gpu.states), cpu.states)
    assert close(np.asarray(gpu.parameters), cpu.parameters)
    gpu.close()


def test_nonfinite_state_raises_assertion_error(built):
    """Counterpart of `assert success` (odeSolver.py:121)."""
    gpu, cpu, X, rng = make_pair("hh_test", 64)
    gpu.states[5, 3] = np.nan
    with pytest.raises(AssertionError):
        gpu.step_lsoda(0.1, None)
    gpu.close()


@pytest.mark.parametrize("n", [0, 1, 31, 127, 129, 1025])
def test_ragged_sizes(built, n):
    gpu, cpu, X, rng = make_pair("hh_test", n)
    V = -70 + 5 * rng.uniform(-1, 1, n)
    for m in (gpu, cpu):
        m.set_state('V', Func(V))
        m.step_lsoda(0.1, {'stim_amplitude': 0.3}, lambda x: x[0] < 30e-6)
    assert np.asarray(gpu.states).shape == (n, 4)
    if n:
        assert close(np.asarray(gpu.states), cpu.states)
    gpu.close()


@pytest.mark.parametrize("name", ["hh_ideal", "calibration"])
def test_dof_ranges_over_several_shards_are_bitwise_identical(built, name):
    """contiguous DOF ranges per device, no collective, result bitwise equal
    to the single-device run.  Two/three shards on device 0 exercise the same code path
    as two/three GPUs."""
    from knpemi_b200.odeSolver import MembraneModel
    n = 10007
    S, P, X, mask = synthetic_tables(name, n, seed=9)
    outs = []
    for devs in ([0], [0, 0], [0, 0, 0]):
        m = MembraneModel(builtin(name), None, 1, Space(X), verbose=False, devices=devs)
        load_tables(m, S, P)
        for _ in range(3):
            m.step_lsoda(SETUP[name]["dt"], {'stim_amplitude': SETUP[name]["stim"]}, lambda x: x[0] < 20e-6)
        outs.append((np.asarray(m.states), np.asarray(m.parameters)))
        m.close()
    for s, p in outs[1:]:
        assert np.array_equal(s, outs[0][0]) and np.array_equal(p, outs[0][1])


def test_permutation_invariance_at_full_size(built):
    """Size-independent property at 10^6 DOFs: DOFs are independent, so permuting the rows
    permutes the results bitwise."""
    from knpemi_b200.odeSolver import MembraneModel
    name, n = "hh_ideal", 1_000_000
    S, P, X, mask = synthetic_tables(name, n, seed=1)
    perm = np.random.default_rng(0).permutation(n)
    res = []
    for order in (np.arange(n), perm):
        m = MembraneModel(builtin(name), None, 1, Space(X[order]), verbose=False, devices=[0])
        load_tables(m, S[order], P[order])
        for _ in range(2):
            m.step_lsoda(1e-4, {'stim_amplitude': 10.0}, lambda x: x[0] < 20e-6)
        res.append((np.asarray(m.states), np.asarray(m.parameters)))
        m.close()
    assert np.array_equal(res[0][0][perm], res[1][0])
    assert np.array_equal(res[0][1][perm], res[1][1])


@pytest.mark.parametrize("pinned", [False, True])
def test_step_exchange_equals_separate_calls(built, pinned):
    """kem_step_io (one pipelined call) == set_parameter x6 + set V + step + get V + get I_ch x3."""
    from knpemi_b200._cabi import PinnedArray
    from knpemi_b200.odeSolver import MembraneModel
    name, n = "hh_ideal", 300_007
    S, P, X, mask = synthetic_tables(name, n, seed=2)
    ode = builtin(name)
    a = MembraneModel(ode, None, 1, Space(X), verbose=False, devices=[0, 0])
    b = MembraneModel(ode, None, 1, Space(X), verbose=False, devices=[0])
    for m in (a, b):
        load_tables(m, S, P)
    keep = []

    def buf(src=None):
        if pinned:
            pa = PinnedArray(n)
            keep.append(pa)
            arr = pa.array
        else:
            arr = np.empty(n)
        arr[:] = 0.0 if src is None else src
        return arr

    rng = np.random.default_rng(4)
    loc = lambda x: x[0] < 20e-6      # noqa: E731
    for step in range(3):
        ins = {("parameter", k): buf(SETUP[name]["varying"][k] * (1 + 0.01 * rng.uniform(-1, 1, n)))
               for k in ("K_e", "K_i", "Na_e", "Na_i", "Cl_e", "Cl_i")}
        ins[("state", "V")] = buf(S[:, 3] * (1 + 0.001 * step))
        outs = {("state", "V"): buf(), ("parameter", "I_ch_Na"): buf(), ("parameter", "I_ch_K"): buf(),
                ("parameter", "I_ch_Cl"): buf()}
        times = a.step_exchange(1e-4, ins, outs, {'stim_amplitude': 10.0}, loc)
        assert times["ms_total"] > 0.0
        for (what, key), arr in ins.items():
            (b.set_state if what == "state" else b.set_parameter)(key, Func(arr))
        b.step_lsoda(1e-4, {'stim_amplitude': 10.0}, loc)
        for (what, key), arr in outs.items():
            u = Func(np.zeros(n))
            (b.get_state if what == "state" else b.get_parameter)(key, u)
            assert np.array_equal(u.x.array, arr), (step, key)
    assert np.array_equal(np.asarray(a.states), np.asarray(b.states))
    a.close()
    b.close()


def test_user_model_is_generated_and_compiled_at_run_time(built, tmp_path):
    """A model module that is not one of the six builtins goes through the generator and
    nvcc on first use (plugin protocol b2).  FitzHugh-Nagumo-like toy with an output slot."""
    import importlib.util
    src = textwrap.dedent('''
        import math
        import numpy as np
        def init_state_values(**values):
            return np.array([0.1, 0.0], dtype=np.float64)
        def init_parameter_values(**values):
            return np.array([0.7, 0.8, 12.5, 0.0, 0.0], dtype=np.float64)
        def state_indices(*names):
            d = {"V": 0, "w": 1}
            r = [d[n] for n in names]
            return r if len(r) > 1 else r[0]
        def parameter_indices(*names):
            d = {"a": 0, "b": 1, "tau": 2, "stim_amplitude": 3, "I_ch": 4}
            for n in names:
                if n not in d:
                    raise ValueError("Unknown param: '{0}'".format(n))
            r = [d[n] for n in names]
            return r if len(r) > 1 else r[0]
        def rhs_numba(t, states, values, parameters):
            a = parameters[0]
            b = parameters[1]
            tau = parameters[2]
            I = parameters[3] * np.exp(-np.mod(t, 5.0) / 2.0)
            cur = states[0] - states[0]**3 / 3 - states[1]
            parameters[4] = cur
            values[0] = cur + I
            values[1] = (states[0] + a - b * states[1]) / tau
    ''')
    path = tmp_path / "mm_toy_fhn.py"
    path.write_text(src)
    spec = importlib.util.spec_from_file_location("mm_toy_fhn", path)
    toy = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(toy)
    from knpemi_b200.odeSolver import MembraneModel
    n = 1000
    rng = np.random.default_rng(0)
    m = MembraneModel(toy, None, 7, Space(rng.uniform(0, 1, (n, 3))), verbose=False, devices=[0])
    assert m.output_columns == [4] and m.tag == 7 and m.prefix == "mm_toy_fhn"
    V0 = rng.uniform(-1, 1, n)
    m.set_state('V', Func(V0))
    for _ in range(5):
        m.step_lsoda(0.1, {'stim_amplitude': 0.5}, lambda x: x[0] < 0.5)
    got = np.asarray(m.states)
    # independent check: the same RK4 in numpy
    X = m.dof_locations
    y = np.stack([V0, np.zeros(n)], axis=1)
    stim = np.where(X[:, 0] < 0.5, 0.5, 0.0)

    def f(t, y):
        import math
        I = stim * math.exp(-math.fmod(t, 5.0) / 2.0)
        cur = y[:, 0] - y[:, 0] * (y[:, 0] * y[:, 0]) / 3 - y[:, 1]
        return np.stack([cur + I, (y[:, 0] + 0.7 - 0.8 * y[:, 1]) / 12.5], axis=1)

    t = 0.0
    for _ in range(5):
        h = 0.1 / 25
        for j in range(25):
            ta, tb, tc = t + j * h, t + (j + 0.5) * h, t + (j + 1.0) * h
            k1 = f(ta, y); k2 = f(tb, y + 0.5 * h * k1); k3 = f(tb, y + 0.5 * h * k2); k4 = f(tc, y + h * k3)
            y = y + h / 6.0 * (((k1 + 2 * k2) + 2 * k3) + k4)
        t = t + 0.1
    assert np.max(np.abs(got - y)) < 1e-12
    cur = y[:, 0] - y[:, 0] ** 3 / 3 - y[:, 1]
    assert np.max(np.abs(m.parameters[:, 4] - cur)) < 1e-12
    m.close()


def test_large_sub_step_counts_and_their_limit(built):
    """The host-evaluated time table lives in shared memory: (2 n_sub + 2) * NT doubles.  Above
    48 KB the launcher opts in to a larger carve-out; beyond 200 KB it refuses loudly."""
    from knpemi_b200._cabi import KemError
    from knpemi_b200.odeSolver import MembraneModel
    from oracle import cpu_oracle
    name = "hh_tissue"
    S, P, X, mask = synthetic_tables(name, 257, seed=6)
    m = MembraneModel(builtin(name), None, 1, Space(X), verbose=False, devices=[0], n_sub=4000)   # 125 KB
    load_tables(m, S, P)
    m.step_lsoda(0.1, None)
    cpu_oracle.step(name, S, P, 0.0, 0.1, 4000)
    assert close(np.asarray(m.states), S)
    with pytest.raises(KemError):
        m.step(0.1, None, n_sub=20000)                     # 625 KB of time table
    with pytest.raises(KemError):
        m.step(0.1, None, n_sub=0)
    m.close()


def test_step_exchange_accepts_dlpack_host_tensors(built):
    """north_star: 'a thin ctypes/C-ABI layer that passes DLPack/NumPy pointers' -- a CPU DLPack
    producer (here a torch tensor) is viewed without a copy."""
    torch = pytest.importorskip("torch")
    from knpemi_b200.odeSolver import MembraneModel
    n = 5000
    S, P, X, mask = synthetic_tables("hh_test", n, seed=2)
    a = MembraneModel(builtin("hh_test"), None, 1, Space(X), verbose=False, devices=[0])
    b = MembraneModel(builtin("hh_test"), None, 1, Space(X), verbose=False, devices=[0])
    for m in (a, b):
        load_tables(m, S, P)
    v_in = torch.from_numpy(S[:, 3].copy() * 1.01)
    v_out = torch.zeros(n, dtype=torch.float64)
    a.step_exchange(0.1, {("state", "V"): v_in}, {("state", "V"): v_out})
    b.set_membrane_potential(Func(v_in.numpy()))
    b.step_lsoda(0.1, None)
    assert np.array_equal(v_out.numpy(), b.states[:, 3])
    a.close()
    b.close()


def test_model_without_parameters(built, tmp_path):
    """Degenerate plugin: no parameter table at all (np = 0), one state."""
    import importlib.util
    src = textwrap.dedent('''
        import math
        import numpy as np
        def init_state_values(**values):
            return np.array([1.0], dtype=np.float64)
        def init_parameter_values(**values):
            return np.array([], dtype=np.float64)
        def state_indices(*names):
            return 0
        def parameter_indices(*names):
            raise ValueError("Unknown param: '{0}'".format(names[0]))
        def rhs_numba(t, states, values, parameters):
            values[0] = -2.0 * states[0]
    ''')
    path = tmp_path / "mm_decay.py"
    path.write_text(src)
    spec = importlib.util.spec_from_file_location("mm_decay", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from knpemi_b200.odeSolver import MembraneModel
    m = MembraneModel(mod, None, 1, Space(np.zeros((300, 3))), verbose=False, devices=[0])
    assert m.parameters.shape == (300, 0) and m.output_columns == []
    for _ in range(10):
        m.step_lsoda(0.05, None)
    assert np.allclose(np.asarray(m.states)[:, 0], np.exp(-1.0), rtol=1e-9)
    with pytest.raises(ValueError):
        m.step_lsoda(0.05, {'stim_amplitude': 1.0})
    m.close()


def test_parameters_the_rhs_never_reads_stay_on_the_host(built):
    """HH never reads Cl_e / Cl_i (mm_hh.py assigns them to unused locals): full-column writes to
    such slots are kept in a host shadow, every API path still sees the values."""
    gpu, cpu, X, rng = make_pair("hh_tissue", 4001, devices=[0, 0])
    n = 4001
    cl_e, cl_i, k_e = rng.normal(size=n), rng.normal(size=n), 3.0 + rng.normal(size=n) * 0.01
    for m in (gpu, cpu):
        m.set_parameter('Cl_e', Func(cl_e))
        m.set_parameter('K_e', Func(k_e))
    assert gpu.column_location('parameter', 'Cl_e') == "host"
    assert gpu.column_location('parameter', 'K_e') == "device"
    assert gpu.column_location('parameter', 'Cm') == "uniform"
    # reads, masked writes (forces the upload), value setters
    loc = lambda x: x[0] < 30e-6      # noqa: E731
    for m in (gpu, cpu):
        m.set_parameter('Cl_e', Func(cl_i), locator=loc)
        m.set_parameter_values({'Cl_i': lambda x: x[1] * 1e3}, locator=loc)
        m.set_parameter('Cl_i', Func(cl_e))                         # full write again: back to the shadow
    assert gpu.column_location('parameter', 'Cl_e') == "device"
    assert gpu.column_location('parameter', 'Cl_i') == "host"
    for name in ('Cl_e', 'Cl_i', 'K_e'):
        a, b = Func(np.zeros(n)), Func(np.zeros(n))
        gpu.get_parameter(name, a)
        cpu.get_parameter(name, b)
        assert np.array_equal(a.x.array, b.x.array), name
    # through the pipelined exchange, as input and as output
    from knpemi_b200._cabi import pinned_empty
    buf_in, buf_out, v_out = pinned_empty(n), pinned_empty(n), pinned_empty(n)
    buf_in[:] = rng.normal(size=n)
    cfg = SETUP["hh_tissue"]
    for m in (gpu, cpu):
        for k, v in {**cfg["uniform"], **{kk: vv for kk, vv in cfg["varying"].items() if kk != 'K_e'}}.items():
            if not k.startswith("Cl"):
                m.set_parameter_values({k: lambda x, v=v: v})
    gpu.step_exchange(0.1, {("parameter", "Cl_e"): buf_in}, {("parameter", "Cl_e"): buf_out, ("state", "V"): v_out})
    cpu.set_parameter('Cl_e', Func(np.array(buf_in)))
    cpu.step_lsoda(0.1, None)
    assert np.array_equal(buf_out, buf_in) and gpu.column_location('parameter', 'Cl_e') == "host"
    assert close(np.asarray(gpu.states), cpu.states)
    assert np.array_equal(np.asarray(gpu.parameters)[:, 13:15], cpu.parameters[:, 13:15])
    gpu.close()


def test_exchange_uploads_unread_pinned_inputs_when_the_host_is_shared(built, monkeypatch):
    """With more than two GPUs fed from one host, kem_step_io sends pinned inputs to unread slots
    over the link instead of through a host copy (csrc/kem_runtime.cu:shadow_pinned_inputs);
    the values every getter sees are the same either way, pageable inputs still use the shadow."""
    from knpemi_b200._cabi import pinned_empty
    n = 3000
    seen = {}
    for policy, local_world in (("shadow", "1"), ("upload", "4")):
        monkeypatch.setenv("LOCAL_WORLD_SIZE", local_world)
        gpu, cpu, X, rng = make_pair("hh_ideal", n, devices=[0])
        cfg = SETUP["hh_ideal"]
        for k, v in {**cfg["uniform"], **cfg["varying"]}.items():
            gpu.set_parameter_values({k: lambda x, v=v: v})
        cl_e, cl_i, v_out, cl_back = pinned_empty(n), rng.normal(size=n), pinned_empty(n), pinned_empty(n)
        cl_e[:] = rng.normal(size=n)
        gpu.step_exchange(cfg["dt"], {("parameter", "Cl_e"): cl_e, ("parameter", "Cl_i"): cl_i},
                          {("state", "V"): v_out})
        assert gpu.column_location('parameter', 'Cl_e') == ("host" if policy == "shadow" else "device")
        assert gpu.column_location('parameter', 'Cl_i') == "host"          # pageable: always the shadow
        gpu.step_exchange(cfg["dt"], {("parameter", "Cl_e"): cl_e}, {("state", "V"): v_out, ("parameter", "Cl_e"): cl_back})
        assert np.array_equal(cl_back, cl_e)
        a, b = Func(np.zeros(n)), Func(np.zeros(n))
        gpu.get_parameter('Cl_e', a)
        gpu.get_parameter('Cl_i', b)
        assert np.array_equal(a.x.array, cl_e) and np.array_equal(b.x.array, cl_i)
        seen[policy] = np.array(v_out)
        gpu.close()
    assert np.array_equal(seen["shadow"], seen["upload"])
