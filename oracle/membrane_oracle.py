"""CPU restatement of the reference ``MembraneModel`` (TEST INFRASTRUCTURE).

Follows src/knpemi/odeSolver.py:6-189 method by method -- AoS ``float64[N,ns]``
/ ``float64[N,np]`` tables (:41-42), column scatter/gather through
``u.x.array`` (:130-166), per-row value setters (:168-188), sticky stimulus
written into the parameter table (:108-112), accumulated ``time`` (:123) --
with two substitutions, both stated in SURVEY.md 8(c):

* ``dolfinx.common.Timer`` -> ``time.perf_counter`` (dolfinx is not installed);
* ``numbalsoda.lsoda`` (absent, un-pinned) -> scheme O1 (RK4 x n_sub + current
  epilogue) in ``oracle/knpemi_oracle.c`` for :meth:`step_lsoda`, and
  ``scipy.integrate`` LSODA at the reference tolerances (rtol 1e-8, atol 1e-10,
  cold start per row, :116-120) for :meth:`step_lsoda_scipy` (scheme O2).

The product never imports this module.
"""
from __future__ import annotations

import time

import numpy as np

from . import cpu_oracle


class OracleMembraneModel:
    def __init__(self, ode, ft, tag, Q, oracle_name=None, n_sub=25, verbose=False):
        assert isinstance(tag, int)                                   # :13
        self.dof_locations = Q.tabulate_dof_coordinates()            # :32
        self.indices = np.arange(len(self.dof_locations))            # :34
        nodes = len(self.indices)
        self.nodes = nodes                                            # :38
        self.states = np.array([ode.init_state_values() for _ in range(nodes)])          # :41
        self.parameters = np.array([ode.init_parameter_values() for _ in range(nodes)])  # :42
        if nodes == 0:     # np.array([]) is 1-D; keep the [N, ncols] shape for an empty membrane
            self.states = np.zeros((0, len(ode.init_state_values())))
            self.parameters = np.zeros((0, len(ode.init_parameter_values())))
        self.tag = tag
        self.ode = ode
        self.prefix = ode.__name__
        self.time = 0
        self.n_sub = n_sub
        self.verbose = verbose
        self.oracle_name = oracle_name or ode.__name__.rsplit(".", 1)[-1]

    # --- :52-67
    def set_state(self, which, u, locator=None):
        return self._set_ODE('state', which, u, locator)

    def set_parameter(self, which, u, locator=None):
        return self._set_ODE('parameter', which, u, locator)

    def get_state(self, which, u, locator=None):
        return self._get_PDE('state', which, u, locator)

    def get_parameter(self, which, u, locator=None):
        return self._get_PDE('parameter', which, u, locator)

    # --- :70-76
    def set_state_values(self, value_dict, locator=None):
        return self._set_ODE_values('state', value_dict, locator)

    def set_parameter_values(self, value_dict, locator=None):
        return self._set_ODE_values('parameter', value_dict, locator)

    # --- :79-89
    def set_membrane_potential(self, u, locator=None):
        return self.set_state('V', u, locator=locator)

    def get_membrane_potential(self, u, locator=None):
        return self.get_state('V', u, locator=locator)

    @property
    def V_index(self):
        return self.ode.state_indices('V')

    # --- :92-127 with the integrator replaced by scheme O1
    def step_lsoda(self, dt, stimulus, stimulus_locator=None):
        if stimulus is None:
            stimulus = {}
        if stimulus_locator is None:
            stimulus_locator = lambda x: True                          # noqa: E731
        mask = np.fromiter(map(stimulus_locator, self.dof_locations), dtype=bool)  # :100
        t_begin = time.perf_counter()
        for key, value in stimulus.items():                            # :110-112 (sticky)
            self.parameters[mask, self.ode.parameter_indices(key)] = value
        bad = cpu_oracle.step(self.oracle_name, self.states, self.parameters,
                              float(self.time), float(dt), self.n_sub)
        assert bad == 0                                                # :121
        self.time = self.time + dt                                     # :106,123
        if self.verbose:
            print(f'\t{self.prefix} Stepped {self.nodes} ODES in {time.perf_counter() - t_begin}s')
        return self.states

    # --- scheme O2: the reference's own integrator family, per-row, cold-started
    def step_lsoda_scipy(self, dt, stimulus, stimulus_locator=None, rtol=1.0e-8, atol=1.0e-10):
        from scipy.integrate import solve_ivp
        if stimulus is None:
            stimulus = {}
        if stimulus_locator is None:
            stimulus_locator = lambda x: True                          # noqa: E731
        mask = np.fromiter(map(stimulus_locator, self.dof_locations), dtype=bool)
        name = self.oracle_name
        t0 = float(self.time)
        for row, is_stimulated in enumerate(mask):                     # :107
            p = self.parameters[row]
            if is_stimulated:
                for key, value in stimulus.items():
                    p[self.ode.parameter_indices(key)] = value

            def f(t, y, p=p):
                dy, p_after = cpu_oracle.rhs(name, t, y, p)
                p[:] = p_after                                         # RHS side effect on I_ch slots
                return dy

            sol = solve_ivp(f, (t0, t0 + dt), self.states[row].copy(), method='LSODA',
                            rtol=rtol, atol=atol, t_eval=[t0 + dt])
            assert sol.success                                         # :121
            self.states[row, :] = sol.y[:, -1]
        self.time = t0 + dt
        return self.states

    # --- work horses :130-188
    def _lidx(self, locator):
        lidx = np.arange(self.nodes)
        if locator is not None:
            lidx = lidx[np.fromiter(map(locator, self.dof_locations), dtype=bool)]
        return lidx

    def _set_ODE(self, what, which, u, locator=None):
        get_index, destination = {'state': (self.ode.state_indices, self.states),
                                  'parameter': (self.ode.parameter_indices, self.parameters)}[what]
        the_index = get_index(which)
        lidx = self._lidx(locator)
        source = u.x.array[:]
        if len(lidx) > 0:
            destination[lidx, the_index] = source[self.indices[lidx]]
        return self.states

    def _get_PDE(self, what, which, u, locator=None):
        get_index, source = {'state': (self.ode.state_indices, self.states),
                             'parameter': (self.ode.parameter_indices, self.parameters)}[what]
        the_index = get_index(which)
        lidx = self._lidx(locator)
        destination = u.x.array[:]
        if len(lidx) > 0:
            destination[self.indices[lidx]] = source[lidx, the_index]
        u.x.array[:] = destination
        return u

    def _set_ODE_values(self, what, value_dict, locator=None):
        destination, get_col = {'state': (self.states, self.ode.state_indices),
                                'parameter': (self.parameters, self.ode.parameter_indices)}[what]
        lidx = self._lidx(locator)
        if len(lidx) == 0:
            return destination
        coords = self.dof_locations[lidx]
        for param in value_dict:
            col = get_col(param)
            get_value = value_dict[param]
            for row, x in zip(lidx, coords):
                destination[row, col] = get_value(x)
        return destination
