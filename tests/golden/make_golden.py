"""Generate the committed golden fixtures from the REFERENCE's own compiled code.

Run in the build container only (needs /root/reference, numba):

    python tests/golden/make_golden.py

For each of the six reference model modules it
  1. imports the module verbatim from /root/reference (through the
     tests/shims/numbalsoda shim that supplies ``lsoda_sig``),
  2. checks oracle/knpemi_oracle.c's restated RHS bit-for-bit against the module's
     ``rhs_numba`` cfunc on 20 000 random points (aborts on any mismatch),
  3. stores 64 of those points (inputs and the cfunc's outputs) as
     ``rhs_<model>.npz``  -- pointwise known-answer vectors,
  4. drives the cfunc through scheme O1 (oracle/knpemi_oracle.c:kemo_step_fn takes
     a foreign function pointer) for a few PDE steps on 48 DOFs and stores inputs
     and end tables as ``traj_<model>.npz``.

The fixtures are what travels to the GPU box; /root/reference does not.
"""
from __future__ import annotations

import ctypes
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "tests", "shims"), os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT]

from knpemi_b200.models import REFERENCE_FILE  # noqa: E402
from oracle import cpu_oracle  # noqa: E402

REFERENCE_ROOT = os.environ.get("KNPEMI_REFERENCE", "/root/reference")
P = ctypes.POINTER(ctypes.c_double)

# per-model physical set-up: values the reference run scripts put into the U/P slots
SETUP = {
    # examples/idealized_geometries/run_2D.py:174-195,237-249,263
    "hh_ideal": dict(dt=1.0e-4, stim=10.0, fill=dict(
        Cm=0.02, psi=96485.0 / (8.314 * 300.0), z_Na=1.0, z_K=1.0, z_Cl=-1.0,
        Na_i=12.838513108648856, Na_e=100.71925900027354, K_i=124.15397583491901,
        K_e=3.3236967382705265, Cl_e=100.71925900027354 + 3.3236967382705265,
        Cl_i=12.838513108648856 + 124.15397583491901)),
    # examples/local_astrocyte_depolarization/run_stim_duration.py:216-242
    "hh_tissue": dict(dt=0.1, stim=5.0, fill=dict(
        Cm=1.0, psi=96500e3 / (8.315e3 * 307e3), z_Na=1.0, z_K=1.0, z_Cl=-1.0,
        Na_i=12.838513108648856, Na_e=100.71925900027354, K_i=124.15397583491901,
        K_e=3.3236967382705265, Cl_e=104.04295573854407, Cl_i=136.99248894356787)),
    "glial_tissue": dict(dt=0.1, stim=0.0, fill=dict(
        Cm=1.0, psi=96500e3 / (8.315e3 * 307e3), z_Na=1.0, z_K=1.0, z_Cl=-1.0,
        Na_i=15.775818906083778, Na_e=144.60625137617149, K_i=99.3100014897692,
        K_e=3.092970607490389, Cl_e=133.62525154406637, Cl_i=5.203660274163705)),
    "glial_bench": dict(dt=0.1, stim=0.0, fill=dict(
        Cm=1.0, psi=96500e3 / (8.315e3 * 307e3), z_Na=1.0, z_K=1.0, z_Cl=-1.0,
        Na_i=15.775818906083778, Na_e=144.60625137617149, K_i=99.3100014897692,
        K_e=3.092970607490389, Cl_e=133.62525154406637, Cl_i=5.203660274163705)),
    "calibration": dict(dt=0.1, stim=2.0, fill=dict()),
    "hh_test": dict(dt=0.1, stim=0.5, fill=dict()),
}


def load_reference(name):
    path = os.path.join(REFERENCE_ROOT, REFERENCE_FILE[name])
    spec = importlib.util.spec_from_file_location("reference_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def base_tables(ref, name, n, rng):
    """Seeded, physically plausible tables [n,ns], [n,np] for model `name`."""
    y0, p0 = ref.init_state_values(), ref.init_parameter_values()
    for k, v in SETUP[name]["fill"].items():
        p0[ref.parameter_indices(k)] = v
    ns = len(y0)
    states = np.tile(y0, (n, 1)) * (1.0 + 0.05 * rng.uniform(-1, 1, (n, ns)))
    if ns >= 4:   # gates stay in (0, 1)
        states[:, :3] = np.clip(np.tile(y0[:3], (n, 1)) + 0.02 * rng.uniform(-1, 1, (n, 3)), 1e-6, 1 - 1e-6)
    params = np.tile(p0, (n, 1)) * (1.0 + 0.02 * rng.uniform(-1, 1, (n, len(p0))))
    try:
        c = ref.parameter_indices("stim_amplitude")
        params[:, c] = np.where(rng.uniform(size=n) < 0.4, SETUP[name]["stim"], 0.0)
    except ValueError:
        pass
    return np.ascontiguousarray(states), np.ascontiguousarray(params)


def main():
    if not os.path.isdir(REFERENCE_ROOT):
        sys.exit(f"{REFERENCE_ROOT} not found: golden vectors can only be generated where the "
                 "reference is mounted")
    cpu_oracle.build(force=True)
    for name in REFERENCE_FILE:
        ref = load_reference(name)
        rng = np.random.default_rng(sum(map(ord, name)))
        dt = SETUP[name]["dt"]
        # ---- 1. pointwise pinning, 20 000 points
        n_pts = 20000
        Y, Pm = base_tables(ref, name, n_pts, rng)
        Y *= (1.0 + 0.25 * rng.uniform(-1, 1, Y.shape))          # wider spread than a trajectory sees
        if Y.shape[1] >= 4:
            Y[:, :3] = rng.uniform(0.0, 1.0, (n_pts, 3))
        T = rng.uniform(0.0, 400.0 * dt, n_pts)
        DY = np.zeros_like(Y)
        PA = Pm.copy()
        f = ref.rhs_numba.ctypes
        bad = 0
        for k in range(n_pts):
            y = Y[k].copy()
            f(T[k], y.ctypes.data_as(P), DY[k].ctypes.data_as(P), PA[k].ctypes.data_as(P))
            dy_o, p_o = cpu_oracle.rhs(name, T[k], Y[k], Pm[k])
            if not (np.array_equal(dy_o, DY[k], equal_nan=True) and np.array_equal(p_o, PA[k], equal_nan=True)):
                bad += 1
        if bad:
            sys.exit(f"{name}: oracle RHS differs from the reference cfunc on {bad}/{n_pts} points")
        keep = slice(0, 64)
        np.savez_compressed(os.path.join(HERE, f"rhs_{name}.npz"), t=T[keep], y=Y[keep], p=Pm[keep],
                            dy=DY[keep], p_after=PA[keep])
        # ---- 2. trajectories: the reference cfunc through scheme O1
        n, n_steps, n_sub = 48, 6, 25
        S0, P0 = base_tables(ref, name, n, rng)
        S, Pt = S0.copy(), P0.copy()
        t = 0.0
        for _ in range(n_steps):
            nbad = cpu_oracle.step_fn(ref.rhs_numba.address, S, Pt, t, dt, n_sub, 1)
            assert nbad == 0, name
            t = t + dt
        # the oracle's own restated RHS must reproduce it bit-for-bit
        S2, P2 = S0.copy(), P0.copy()
        t = 0.0
        for _ in range(n_steps):
            cpu_oracle.step(name, S2, P2, t, dt, n_sub, 1)
            t = t + dt
        if not (np.array_equal(S, S2) and np.array_equal(Pt, P2)):
            sys.exit(f"{name}: oracle trajectory differs from the reference-cfunc trajectory")
        np.savez_compressed(os.path.join(HERE, f"traj_{name}.npz"), states0=S0, params0=P0,
                            states=S, params=Pt, dt=dt, n_steps=n_steps, n_sub=n_sub)
        print(f"{name}: {n_pts} RHS points bit-exact; trajectory {n} DOFs x {n_steps} steps stored")


if __name__ == "__main__":
    main()
