"""Kernel-only throughput of the fused step against the membrane size, 10^4 .. 10^8 DOFs
(north star: "synthetic membrane DOF arrays (10^4 to 10^8 DOFs)").  Per-DOF inputs are written
column by column so the 10^8 case needs no 17 GB host table."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT]
import numpy as np
from knpemi_b200 import _cabi
from knpemi_b200.odeSolver import MembraneModel
from workloads import SETUP, builtin


class Space:
    def __init__(self, n):
        self.x = np.broadcast_to(np.zeros((1, 3)), (n, 3))

    def tabulate_dof_coordinates(self):
        return self.x


name = sys.argv[1] if len(sys.argv) > 1 else "hh_ideal"
scheme = sys.argv[2] if len(sys.argv) > 2 else "rk4"
ode, cfg = builtin(name), SETUP[name]
peak, _ = _cabi.fp64_peak(0)
print(f"model {name}, scheme {scheme}, FP64 peak {peak:.2f} TFLOP/s")
for n in (10_000, 100_000, 1_000_000, 10_000_000, 100_000_000):
    rng = np.random.default_rng(n)
    m = MembraneModel(ode, None, 1, Space(n), devices=[0], verbose=False, scheme=scheme)
    for k, v in cfg["uniform"].items():
        m.set_parameter_values({k: lambda x, v=v: v})
    for k, v in cfg["varying"].items():                       # per-DOF concentrations
        m._set_column(1, ode.parameter_indices(k), v * (1 + 0.02 * rng.uniform(-1, 1, n)))
    iv = ode.state_indices("V") if name != "calibration" else ode.state_indices("V_n")
    m._set_column(0, iv, ode.init_state_values()[iv] * (1 + 0.05 * rng.uniform(-1, 1, n)))
    mask = (rng.uniform(size=n) < 0.32).astype(np.uint8)
    _cabi.check(m._lib.kem_set_stimulus_mask(m._h, mask.ctypes.data, n))
    m._stim_mask_key = id(mask); m._mask_cache[id(None)] = (None, None)
    stim = {"stim_amplitude": cfg["stim"]}
    steps = 30 if n <= 10_000_000 else 5
    for _ in range(3):
        m.step_async(cfg["dt"], stim, None)
    m.synchronize()
    m.timer_begin()
    for _ in range(steps):
        m.step_async(cfg["dt"], stim, None)
    ms = m.timer_end() / steps
    print(f"  N = {n:>11,d}: {ms:9.4f} ms per PDE step -> {n / ms * 1e3:.3e} DOF-steps/s")
    m.close()
