"""Stand-alone consumer of the membrane stage: calibrate the initial conditions of the
neuron + glia + ECS compartment model by running it to steady state.

Same experiment as the reference's examples/calibrate_initial_conditions/run_calibration.py
(11 membrane DOFs on an interval mesh, 10 000 steps of dt = 0.1 ms, no stimulus, state columns
read through `membrane.states[:, idx]` every step), with the dolfinx interval mesh replaced by
11 points -- the stage only needs DOF coordinates.  Imports through the reference's own path
`knpemi.odeSolver` (compat shim).

    python examples/run_calibration_b200.py [--steps 10000] [--scheme rk4|dp45]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "knp-emi-fenics-x_b200", "compat"), os.path.join(ROOT, "knp-emi-fenics-x_b200")]

import numpy as np  # noqa: E402
from knpemi.odeSolver import MembraneModel  # noqa: E402
from knpemi_b200.ducks import PointSpace  # noqa: E402
from knpemi_b200.models import calibration as ode  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--scheme", default="rk4", choices=["rk4", "dp45"])
    args = ap.parse_args()

    M = 10
    Q = PointSpace(np.stack([np.linspace(0, 1, M + 1), np.zeros(M + 1), np.zeros(M + 1)], axis=1))
    membrane = MembraneModel(ode, None, 1, Q, scheme=args.scheme, verbose=False)
    names = [n for n, _ in ode.STATES]
    index = {n: ode.state_indices(n) for n in names}
    history = {n: [] for n in names}

    t0 = time.perf_counter()
    for _ in range(args.steps):
        membrane.step_lsoda(dt=0.1, stimulus={'stim_amplitude': 0})
        for n in ("V_n", "V_g"):                       # the reference records all 14 every step
            history[n].append(1 * membrane.states[:, index[n]])
    wall = time.perf_counter() - t0
    final = np.asarray(membrane.states)

    print("-------------------------------------------------------------")
    pretty = {"V_n": "phi_M_n_init", "V_g": "phi_M_g_init"}
    for n in names:
        print(f"{pretty.get(n, n + '_init')} =", final[2, index[n]])
    print("-------------------------------------------------------------")
    drift = np.abs(np.array(history["V_g"][-1]) - np.array(history["V_g"][-2])).max()
    print(f"{args.steps} PDE steps of 11 DOFs in {wall:.2f} s ({1e6 * wall / args.steps:.0f} us per step, "
          f"scheme {args.scheme}); last-step change of V_g: {drift:.2e} mV")


if __name__ == "__main__":
    main()
