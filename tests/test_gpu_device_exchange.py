"""PDE<->ODE exchange with device-resident bulk vectors.
Oracle: NumPy fancy indexing (what the CG-1 trace + setter of the reference amount to)."""
import numpy as np
import pytest

from ducks_for_tests import Func, Space
from workloads import SETUP, builtin, load_tables, synthetic_tables

pytestmark = pytest.mark.gpu


def test_gather_scatter_and_potential_jump(built):
    from knpemi_b200._cabi import DeviceArray, KemError
    from knpemi_b200.odeSolver import MembraneModel
    name, n, n_bulk = "hh_ideal", 50_003, 400_000
    rng = np.random.default_rng(0)
    S, P, X, mask = synthetic_tables(name, n, seed=3)
    ode = builtin(name)
    a = MembraneModel(ode, None, 1, Space(X), verbose=False, devices=[0])     # device-resident exchange
    b = MembraneModel(ode, None, 1, Space(X), verbose=False, devices=[0])     # host setters/getters
    for m in (a, b):
        load_tables(m, S, P)
    map_e = rng.choice(n_bulk, n, replace=False)
    map_i = rng.choice(n_bulk, n, replace=False)
    a.register_trace_map(0, map_e)
    a.register_trace_map(1, map_i)
    K_e_bulk = 3.3 * (1 + 0.02 * rng.uniform(-1, 1, n_bulk))
    phi_e = 1e-3 * rng.normal(size=n_bulk)
    phi_i = -0.07 + 1e-3 * rng.normal(size=n_bulk)
    d_K, d_pe, d_pi = DeviceArray(0, K_e_bulk), DeviceArray(0, phi_e), DeviceArray(0, phi_i)
    a.gather_from_device('parameter', 'K_e', d_K.ptr, 0)
    a.set_membrane_potential_from_device(d_pi.ptr, 1, d_pe.ptr, 0)
    b.set_parameter('K_e', Func(K_e_bulk[map_e]))
    b.set_membrane_potential(Func(phi_i[map_i] - phi_e[map_e]))
    for m in (a, b):
        m.step_lsoda(1e-4, {'stim_amplitude': 10.0}, lambda x: x[0] < 20e-6)
    assert np.array_equal(np.asarray(a.states), np.asarray(b.states))
    assert np.array_equal(np.asarray(a.parameters), np.asarray(b.parameters))
    # ODE -> PDE: scatter I_ch_Na into a bulk-sized device vector
    d_out = DeviceArray(0, np.zeros(n_bulk))
    a.scatter_to_device('parameter', 'I_ch_Na', d_out.ptr, 1)
    want = np.zeros(n_bulk)
    want[map_i] = b.parameters[:, ode.parameter_indices('I_ch_Na')]
    assert np.array_equal(d_out.to_host(), want)
    # a uniform column scatters its value
    a.scatter_to_device('parameter', 'Cm', d_out.ptr, 1)
    assert np.all(d_out.to_host()[map_i] == SETUP[name]["uniform"]["Cm"])
    with pytest.raises(KemError, match="map not registered"):
        a.gather_from_device('parameter', 'K_e', d_K.ptr, 5)
    with pytest.raises(KemError):
        a.register_trace_map(2, -np.ones(n, dtype=np.int64))
    for d in (d_K, d_pe, d_pi, d_out):
        d.free()
    a.close()
    b.close()


def test_cuda_array_interface_columns(built):
    """Device-to-device column copies from / to CUDA arrays of another library (torch here)."""
    torch = pytest.importorskip("torch")
    from knpemi_b200._cabi import KemError
    from knpemi_b200.odeSolver import MembraneModel
    name, n = "hh_ideal", 20_001
    S, P, X, mask = synthetic_tables(name, n, seed=4)
    ode = builtin(name)
    a = MembraneModel(ode, None, 1, Space(X), verbose=False, devices=[0])
    b = MembraneModel(ode, None, 1, Space(X), verbose=False, devices=[0])
    for m in (a, b):
        load_tables(m, S, P)
    assert a.shard_ranges() == [(0, 0, n)]
    k_e = torch.tensor(3.3 * (1 + 0.01 * np.random.default_rng(0).uniform(-1, 1, n)), device="cuda:0")
    a.set_from_cuda_array('parameter', 'K_e', k_e)
    b.set_parameter('K_e', Func(k_e.cpu().numpy()))
    for m in (a, b):
        m.step_lsoda(1e-4, {'stim_amplitude': 10.0}, lambda x: x[0] < 20e-6)
    out = torch.zeros(n, dtype=torch.float64, device="cuda:0")
    a.get_to_cuda_array('state', 'V', out)
    assert np.array_equal(out.cpu().numpy(), b.states[:, 3])
    with pytest.raises(KemError):
        a.set_from_cuda_array('parameter', 'K_e', torch.zeros(n, dtype=torch.float32, device="cuda:0"))
    with pytest.raises(KemError):
        a.set_from_cuda_array('parameter', 'K_e', np.zeros(n))
    a.close()
    b.close()


def test_eliminated_ion_concentration_on_device(built):
    """f3: c_elim = -(1/z_e) (rho_z rho_tag + sum_k z_k c_k) (utils.py:247-267) over a bulk vector,
    and its membrane trace straight into the parameter column update_ode_variables would set
    (utils.py:219-228).  Checked against the same expression in NumPy, summed in the reference's
    order; the kernel may contract a*b+c into an FMA, hence one ulp."""
    from knpemi_b200._cabi import DeviceArray, KemError
    from knpemi_b200.device_updates import affine_combine, eliminated_ion_terms
    from knpemi_b200.odeSolver import MembraneModel
    rng = np.random.default_rng(1)
    # ion_list = [K, Cl, Na]: Na is eliminated (run_2D.py:252-253); rho from run_2D.py physical parameters
    ion_list = [{"name": "K", "z": 1.0}, {"name": "Cl", "z": -1.0}, {"name": "Na", "z": 1.0}]
    rho_z, rho_tag = -1.0, 77.0
    a0, coefs = eliminated_ion_terms(ion_list, rho_z, rho_tag)
    assert a0 == -(1.0 / 1.0) * rho_z * rho_tag and coefs == [-(1.0 / 1.0) * 1.0, -(1.0 / 1.0) * -1.0]
    for n_bulk in (0, 1, 1001, 400_000):
        c_K = 120.0 * (1 + 0.05 * rng.uniform(-1, 1, n_bulk))
        c_Cl = 140.0 * (1 + 0.05 * rng.uniform(-1, 1, n_bulk))
        want = a0 + coefs[0] * c_K + coefs[1] * c_Cl           # c_elim_sum of utils.py:249-258
        d_K, d_Cl, d_out = DeviceArray(0, c_K), DeviceArray(0, c_Cl), DeviceArray(0, np.zeros(max(n_bulk, 1)))
        affine_combine(0, n_bulk, d_out.ptr, a0, [(coefs[0], d_K.ptr), (coefs[1], d_Cl.ptr)])
        got = d_out.to_host()[:n_bulk]
        assert np.allclose(got, want, rtol=2.3e-16, atol=0), n_bulk
        if n_bulk > 1:                                          # unaligned views take the scalar path
            affine_combine(0, n_bulk - 1, d_out.ptr + 8, a0, [(coefs[0], d_K.ptr + 8), (coefs[1], d_Cl.ptr + 8)])
            assert np.allclose(d_out.to_host()[1:n_bulk], want[1:], rtol=2.3e-16, atol=0)
        for d in (d_K, d_Cl, d_out):
            d.free()
    # the membrane trace of the same combination, into Na_i of an HH model
    name, n, n_bulk = "hh_ideal", 30_011, 200_000
    S, P, X, mask = synthetic_tables(name, n, seed=6)
    ode = builtin(name)
    a = MembraneModel(ode, None, 1, Space(X), verbose=False, devices=[0])
    b = MembraneModel(ode, None, 1, Space(X), verbose=False, devices=[0])
    for m in (a, b):
        load_tables(m, S, P)
    trace = rng.choice(n_bulk, n, replace=False)
    a.register_trace_map(0, trace)
    c_K = 120.0 * (1 + 0.05 * rng.uniform(-1, 1, n_bulk))
    c_Cl = 140.0 * (1 + 0.05 * rng.uniform(-1, 1, n_bulk))
    d_K, d_Cl = DeviceArray(0, c_K), DeviceArray(0, c_Cl)
    a0_i = a0 + 6.0
    a.set_from_device_affine('parameter', 'Na_i', a0_i, [(coefs[0], d_K.ptr), (coefs[1], d_Cl.ptr)], 0)
    na_i = Func(np.zeros(n))
    a.get_parameter('Na_i', na_i)
    want = (a0_i + coefs[0] * c_K + coefs[1] * c_Cl)[trace]
    assert np.allclose(na_i.x.array, want, rtol=2.3e-16, atol=0)
    b.set_parameter('Na_i', na_i)                               # the host path with the same values
    for m in (a, b):
        m.step_lsoda(1e-4, {'stim_amplitude': 10.0}, lambda x: x[0] < 20e-6)
    assert np.array_equal(np.asarray(a.states), np.asarray(b.states))
    with pytest.raises(KemError):
        a.set_from_device_affine('parameter', 'Na_i', 0.0, [(1.0, d_K.ptr)] * 9, 0)
    for d in (d_K, d_Cl):
        d.free()
    a.close()
    b.close()
