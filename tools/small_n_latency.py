"""Per-call latency of the drop-in API on a real-mesh-sized membrane (496 DOFs: the 2D idealized
mesh at res 3)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT]
import numpy as np
from knpemi_b200.ducks import ArrayFunction, PointSpace
from knpemi_b200.odeSolver import MembraneModel
from workloads import SETUP, builtin, load_tables, synthetic_tables

for n in (124, 496, 2952, 100_000):
    name = "hh_ideal"
    S, P, X, mask = synthetic_tables(name, n)
    ode = builtin(name)
    m = MembraneModel(ode, None, 1, PointSpace(X), devices=[0], verbose=False)
    load_tables(m, S, P)
    ins = {k: ArrayFunction(P[:, ode.parameter_indices(k)].copy()) for k in ("K_e", "K_i", "Na_e", "Na_i", "Cl_e", "Cl_i")}
    phi = ArrayFunction(S[:, 3].copy()); I = {k: ArrayFunction(n) for k in ("Na", "K", "Cl")}
    loc = lambda x: x[0] < 20e-6
    def pde_step():
        for k, u in ins.items(): m.set_parameter(k, u)
        m.set_membrane_potential(phi)
        m.step_lsoda(1e-4, {"stim_amplitude": 10.0}, loc)
        m.get_membrane_potential(phi)
        for k, u in I.items(): m.get_parameter("I_ch_" + k, u)
    for _ in range(20): pde_step()
    t0 = time.perf_counter()
    for _ in range(200): pde_step()
    full = (time.perf_counter() - t0) / 200
    t0 = time.perf_counter()
    for _ in range(200): m.step_lsoda(1e-4, {"stim_amplitude": 10.0}, loc)
    step = (time.perf_counter() - t0) / 200
    print(f"N={n:7d}: set x7 + step + get x4 = {1e6*full:7.0f} us per PDE step; step_lsoda alone {1e6*step:6.0f} us")
    m.close()
