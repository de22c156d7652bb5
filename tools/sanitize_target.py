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
        m2.close()
    print(name, scheme, float(np.asarray(m.states).sum()))
    m.close()
print("SANITIZE_TARGET_OK")
