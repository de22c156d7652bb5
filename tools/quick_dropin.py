"""Times the unmodified reference call sequence (7 setters, step_lsoda, 4 getters) with
pageable NumPy arrays, i.e. what solve_odes (run_2D.py:80-111) does per PDE step; with
`--register` the same calls again after MembraneModel.register_host_array on every array."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT]
import numpy as np
from knpemi_b200.ducks import ArrayFunction, PointSpace
from knpemi_b200.odeSolver import MembraneModel
from workloads import SETUP, builtin, load_tables, synthetic_tables

name = "hh_ideal"; n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
S, P, X, mask = synthetic_tables(name, n)
ode = builtin(name)
m = MembraneModel(ode, None, 1, PointSpace(X), devices=[0], verbose=False)
load_tables(m, S, P)
ins = {k: ArrayFunction(P[:, ode.parameter_indices(k)].copy()) for k in ("K_e", "K_i", "Na_e", "Na_i", "Cl_e", "Cl_i")}
phi = ArrayFunction(S[:, 3].copy())
I = {k: ArrayFunction(n) for k in ("Na", "K", "Cl")}
loc = lambda x: x[0] < 20e-6
register = "--register" in sys.argv
for it in range(8 if register else 5):
    if register and it == 4:
        t0 = time.perf_counter()
        for u in [*ins.values(), phi, *I.values()]:
            m.register_host_array(u)
        print(f"registered 11 arrays of {8 * n / 1e6:.0f} MB in {1e3 * (time.perf_counter() - t0):.1f} ms")
    t0 = time.perf_counter()
    for k, u in ins.items():
        m.set_parameter(k, u)
    m.set_membrane_potential(phi)
    t1 = time.perf_counter()
    m.step_lsoda(1e-4, {"stim_amplitude": 10.0}, loc)
    t2 = time.perf_counter()
    m.get_membrane_potential(phi)
    for k, u in I.items():
        m.get_parameter("I_ch_" + k, u)
    t3 = time.perf_counter()
    print(f"step {it}: set {1e3*(t1-t0):.1f} ms, step {1e3*(t2-t1):.1f} ms, get {1e3*(t3-t2):.1f} ms, "
          f"total {1e3*(t3-t0):.1f} ms -> {n/(t3-t0):.3e} DOF-steps/s "
          f"(H2D {7*8*n/(t1-t0)/1e9:.1f} GB/s, D2H {4*8*n/(t3-t2)/1e9:.1f} GB/s)")
