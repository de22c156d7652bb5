"""Run the REFERENCE's own membrane glue, verbatim, and record what it does.

Build container only (needs /root/reference and numba):

    python tests/golden/make_glue_transcript.py

What is executed from /root/reference, unmodified:
  * ``src/knpemi/odeSolver.py``  -- the reference ``MembraneModel`` class itself,
  * ``src/knpemi/utils.py``      -- ``setup_membrane_model`` (:105-148) and
                                    ``update_ode_variables`` (:210-235),
  * ``examples/idealized_geometries/mm_hh.py`` -- the model module with its numba cfunc,
driven by a restatement of ``solve_odes`` (examples/idealized_geometries/run_2D.py:80-111:
update -> step_lsoda -> get_membrane_potential -> get_parameter("I_ch_"+ion)) with the
parameters of run_2D.py:174-195,237-266.

What is substituted (tests/shims/): ``dolfinx`` / ``scifem`` / ``mpi4py`` / ``ufl`` by stubs of
the few names the membrane side touches; ``interpolate_to_membrane`` (utils.py:150-207, the
scifem trace interpolation -- PDE side, out of scope) by seeded synthetic traces; and
``numbalsoda.lsoda`` (absent, un-pinned) by scheme O1 over the cfunc address the class hands it.
So every line of the reference class -- table construction, the three work-horses, the
stimulus loop, time accumulation -- runs as written; only the per-row integrator is ours, by
necessity.

Two fixtures are written next to this script:
  glue_hh_ideal.json  the CALL TRANSCRIPT: every MembraneModel method the reference's glue
                      invoked, in order, with scalar arguments inline and array arguments by key;
  glue_hh_ideal.npz   the arrays: inputs the glue passed in (by key) and the values it read
                      back after each getter.
tests/test_reference_glue.py regenerates and compares them here; the ``-m gpu`` replay
(tests/test_gpu_reference_glue.py) drives the CUDA backend through the transcript on the GPU
box, where /root/reference does not exist.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "tests", "shims"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE_ROOT = os.environ.get("KNPEMI_REFERENCE", "/root/reference")
N_DOF, N_STEPS, DT, SEED = 496, 4, 1.0e-4, 20240611
LOCATOR_SOURCE = "x[0] < 20e-6"                                    # run_2D.py:264


def load_reference():
    """(odeSolver module, utils module, mm_hh module) loaded from the reference tree by path;
    ``knpemi`` is a synthetic package so that ``src/knpemi/__init__.py`` (which pulls in the
    PDE side) is not executed."""
    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    pkg = types.ModuleType("knpemi")
    pkg.__path__ = []
    sys.modules["knpemi"] = pkg
    ode_solver = load("knpemi.odeSolver", "src/knpemi/odeSolver.py")
    pkg.odeSolver = ode_solver
    utils = load("knpemi.utils", "src/knpemi/utils.py")
    pkg.utils = utils
    mm_hh = load("mm_hh", "examples/idealized_geometries/mm_hh.py")
    return ode_solver, utils, mm_hh


class Space:
    def __init__(self, X):
        self.X = X

    def tabulate_dof_coordinates(self):
        return self.X


class Token:
    """Stands for a PDE-side concentration function: only its name travels."""

    def __init__(self, name):
        self.name = name


class Recorder:
    """Wraps a MembraneModel; logs every public call the glue makes."""

    def __init__(self, inner, calls, arrays):
        self.__dict__.update(_inner=inner, _calls=calls, _arrays=arrays)

    def __getattr__(self, name):
        attr = getattr(self._inner, name)
        if not callable(attr) or name.startswith("_"):
            return attr

        def logged(*args, **kw):
            entry = {"method": name}
            if name in ("set_parameter_values", "set_state_values"):
                # value callables are not serialisable: record what they return on the first DOF
                # (the glue only passes constants, utils.py:124-129)
                x0 = self._inner.dof_locations[0]
                entry["values"] = {k: float(f(x0)) for k, f in args[0].items()}
            elif name in ("set_parameter", "set_state", "set_membrane_potential"):
                u = args[-1] if name == "set_membrane_potential" else args[1]
                key = f"in{len(self._arrays):03d}_{name}" + ("" if name == "set_membrane_potential" else "_" + args[0])
                self._arrays[key] = np.array(u.x.array)
                entry.update(which=None if name == "set_membrane_potential" else args[0], array=key)
            elif name == "step_lsoda":
                entry.update(dt=float(kw.get("dt", args[0] if args else None)), stimulus=dict(kw.get("stimulus") or {}),
                             locator=LOCATOR_SOURCE if kw.get("stimulus_locator") is not None else None)
            out = attr(*args, **kw)
            if name in ("get_parameter", "get_state", "get_membrane_potential"):
                u = args[-1] if name == "get_membrane_potential" else args[1]
                key = f"out{len(self._arrays):03d}_{name}" + ("" if name == "get_membrane_potential" else "_" + args[0])
                self._arrays[key] = np.array(u.x.array)
                entry.update(which=None if name == "get_membrane_potential" else args[0], expect=key)
            self._calls.append(entry)
            return out
        return logged


def run_glue(model_class, ode, n=N_DOF, n_steps=N_STEPS, seed=SEED, record=True):
    """The reference's setup_membrane_model + solve_odes loop around `model_class`.
    Returns (calls, arrays, final states, final parameters)."""
    import dolfinx
    _, utils, _ = load_reference() if "knpemi.utils" not in sys.modules else (None, sys.modules["knpemi.utils"], None)
    rng = np.random.default_rng(seed)
    X = rng.uniform(0.0, 62e-6, (n, 3))
    Q = Space(X)
    calls, arrays = [], {}

    # run_2D.py:174-195, 237-266
    C_M, psi = 0.02, 96485.0 / (8.314 * 300.0)
    init = dict(Na_i=12.838513108648856, Na_e=100.71925900027354, K_i=124.15397583491901, K_e=3.3236967382705265)
    init["Cl_e"] = init["Na_e"] + init["K_e"]
    init["Cl_i"] = init["Na_i"] + init["K_i"]
    ion_list = [{"name": "K", "z": 1.0}, {"name": "Cl", "z": -1.0}, {"name": "Na", "z": 1.0}]
    ion_list[-1]["c_0"], ion_list[-1]["c_1"] = Token("Na_e"), Token("Na_i")
    c_prev = {0: [Token("K_e"), Token("Cl_e")], 1: [Token("K_i"), Token("Cl_i")]}
    stim_params = {"stimulus": {"stim_amplitude": 10}, "stimulus_locator": eval("lambda x: " + LOCATOR_SOURCE)}
    physical = {"C_M": dolfinx.fem.Constant(None, C_M), "psi": psi}

    # the class the glue instantiates (utils.py:122) is looked up in the utils module's namespace
    def factory(ode_, ct, tag, Q_):
        inner = model_class(ode_, ct, tag, Q_)
        return Recorder(inner, calls, arrays) if record else inner
    saved_class, saved_interp = utils.MembraneModel, utils.interpolate_to_membrane
    utils.MembraneModel = factory

    step_no = {"k": 0}

    def synthetic_traces(ue, ui, Q_, mesh, ct, subdomain_list, tag):
        """(qe, qi) as interpolate_to_membrane returns them (utils.py:190-207): fresh Functions."""
        out = []
        for tok in (ue, ui):
            f = dolfinx.fem.Function(Q_, name=tok.name)
            trng = np.random.default_rng([seed, step_no["k"], sum(map(ord, tok.name))])
            f.x.array[:] = init[tok.name] * (1.0 + 0.01 * trng.uniform(-1, 1, n))
            out.append(f)
        return tuple(out)
    utils.interpolate_to_membrane = synthetic_traces
    try:
        mem_models = utils.setup_membrane_model(stim_params, physical, {1: ode}, None, Q, ion_list)
        phi_M_prev = dolfinx.fem.Function(Q, name="phi_M")
        subdomain_list = {0: {}, 1: {"mem_models": mem_models}}
        # solve_odes, run_2D.py:80-111
        for k in range(n_steps):
            step_no["k"] = k
            for tag, subdomain in subdomain_list.items():
                if tag > 0:
                    for mem_model in subdomain["mem_models"]:
                        ode_model = mem_model["ode"]
                        utils.update_ode_variables(ode_model, c_prev, phi_M_prev, ion_list, subdomain_list,
                                                   None, None, tag, k)
                        ode_model.step_lsoda(dt=DT, stimulus=stim_params["stimulus"],
                                             stimulus_locator=stim_params["stimulus_locator"])
                        ode_model.get_membrane_potential(phi_M_prev)
                        for ion, I_ch_k in mem_model["I_ch_k"].items():
                            ode_model.get_parameter("I_ch_" + ion, I_ch_k)
            # the PDE solve would change phi_M here (update_pde_variables, utils.py:288-291)
            phi_M_prev.x.array[:] += 1.0e-4 * np.sin(1.0e5 * X[:, 1] + k)
    finally:
        utils.MembraneModel, utils.interpolate_to_membrane = saved_class, saved_interp
    m = mem_models[0]["ode"]
    return calls, arrays, np.array(m.states), np.array(m.parameters), X


def main():
    ode_solver, utils, mm_hh = load_reference()
    calls, arrays, S, P, X = run_glue(ode_solver.MembraneModel, mm_hh)
    arrays["dof_coordinates"] = X
    arrays["final_states"], arrays["final_parameters"] = S, P
    meta = {"generated_by": "tests/golden/make_glue_transcript.py",
            "reference_files": ["src/knpemi/odeSolver.py", "src/knpemi/utils.py",
                                "examples/idealized_geometries/mm_hh.py", "examples/idealized_geometries/run_2D.py:80-111"],
            "model": "hh_ideal", "n_dof": N_DOF, "n_steps": N_STEPS, "dt": DT, "integrator": "scheme O1 (RK4 x 25) in place "
            "of numbalsoda.lsoda, through the reference cfunc", "calls": calls}
    with open(os.path.join(HERE, "glue_hh_ideal.json"), "w") as f:
        json.dump(meta, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "glue_hh_ideal.npz"), **arrays)
    kinds = {}
    for c in calls:
        kinds[c["method"]] = kinds.get(c["method"], 0) + 1
    print(f"{len(calls)} calls recorded: {kinds}; {len(arrays)} arrays")


if __name__ == "__main__":
    main()
