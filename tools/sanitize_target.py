"""Small run of every kernel family for `compute-sanitizer --tool memcheck` (one tool per gpurun call)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT]
import numpy as np
from knpemi_b200._cabi import DeviceArray, PinnedArray
from knpemi_b200.ducks import ArrayFunction, PointSpace
from knpemi_b200.odeSolver import MembraneModel
from workloads import SETUP, builtin, load_tables, synthetic_tables

for name, scheme in (("hh_ideal", "rk4"), ("calibration", "rk4"), ("glial_tissue", "rk4"), ("hh_tissue", "dp45")):
    n = 1237
    S, P, X, mask = synthetic_tables(name, n, seed=1)
    m = MembraneModel(builtin(name), None, 1, PointSpace(X), devices=[0, 0], verbose=False, scheme=scheme)
    load_tables(m, S, P)
    loc = lambda x: x[0] < 20e-6
    for _ in range(2):
        m.step_lsoda(SETUP[name]["dt"], {"stim_amplitude": SETUP[name]["stim"]}, loc)
    m.set_state_values({builtin(name).STATES[0][0]: lambda x: 0.5}, locator=loc)
    u = ArrayFunction(n)
    m.get_state(builtin(name).STATES[0][0], u, locator=loc)
    if name == "hh_ideal":
        ins = {("parameter", "K_e"): PinnedArray(n).array, ("state", "V"): PinnedArray(n).array}
        ins[("parameter", "K_e")][:] = 3.3
        ins[("state", "V")][:] = -0.07
        outs = {("parameter", "I_ch_Na"): PinnedArray(n).array}
        m.step_exchange(1e-4, ins, outs, {"stim_amplitude": 10.0}, loc)
        m.register_trace_map(0, np.arange(n)[::-1].copy())
        d = DeviceArray(0, np.linspace(3, 4, n))
        m2 = MembraneModel(builtin(name), None, 1, PointSpace(X), devices=[0], verbose=False)
        m2.register_trace_map(0, np.arange(n)[::-1].copy())
        m2.gather_from_device("parameter", "K_e", d.ptr, 0)
        m2.scatter_to_device("parameter", "K_e", d.ptr, 0)
        # round 2: eliminated-ion kernels, unread-input policies, literal outputs, chunked step + getter
        from knpemi_b200.device_updates import affine_combine
        d2, d3 = DeviceArray(0, np.linspace(1, 2, n)), DeviceArray(0, np.zeros(n))
        affine_combine(0, n, d3.ptr, 0.5, [(1.0, d.ptr), (-1.0, d2.ptr)])
        affine_combine(0, n - 1, d3.ptr + 8, 0.5, [(1.0, d.ptr + 8)])
        m2.set_from_device_affine("parameter", "Na_i", 12.0, [(1e-3, d.ptr), (1e-3, d2.ptr)], 0)
        m2.close()
        for policy in ("shadow", "upload", "discard"):
            m3 = MembraneModel(builtin(name), None, 1, PointSpace(X), devices=[0], verbose=False, unread_inputs=policy)
            load_tables(m3, S, P)
            ins3 = {("parameter", k): PinnedArray(n).array for k in ("K_e", "Cl_e")}
            ins3[("parameter", "K_e")][:] = 3.3
            ins3[("parameter", "Cl_e")][:] = 100.0
            outs3 = {("parameter", k): PinnedArray(n).array for k in ("I_ch_Na", "I_ch_Cl")}
            m3.step_exchange(1e-4, ins3, outs3, {"stim_amplitude": 10.0}, loc)
            from knpemi_b200._cabi import check
            check(m3._lib.kem_set_step_chunks(m3._h, 16), "kem_set_step_chunks")
            m3.step_async(1e-4, {"stim_amplitude": 10.0}, loc)
            back = ArrayFunction(n)
            back.x.array = PinnedArray(n).array
            m3.get_membrane_potential(back)
            m3.close()
    print(name, scheme, float(np.asarray(m.states).sum()))
    m.close()
print("SANITIZE_TARGET_OK")
