"""Hodgkin-Huxley membrane with fixed Nernst potentials (mV, ms).

Builtin restatement of the reference's orphan model module tests/mm_test_ode.py
(states :12-15, parameters :38-70, right-hand side :126-169): 17 parameters,
no ``psi`` / ``z_*`` slots, all defaults directly usable (``Cm = 1``).
"""
import math

import numpy as np

from ._protocol import rhs_cfunc, table_functions

STATES = (
    ("m", 0.016648440745822956),
    ("h", 0.8542015627820805),
    ("n", 0.1882020248041632),
    ("V", -74.38609374462003),
)

PARAMETERS = (
    ("g_Na_bar", 120.0), ("g_K_bar", 36.0),
    ("g_leak_Na", 0.1), ("g_leak_K", 0.4),
    ("E_Na", 53.23236322443255), ("E_K", -93.46115007798299),
    ("Cm", 1.0), ("stim_amplitude", 0.0),
    ("I_ch_Na", 0.0), ("I_ch_K", 0.0), ("I_ch_Cl", 0.0),
    ("K_e", 3.32), ("Na_i", 12.83),
    ("m_K", 2.0), ("m_Na", 7.7), ("I_max", 50.0),
    ("E_Cl", 70.97802159265801),
)

(init_state_values, init_parameter_values,
 state_indices, parameter_indices) = table_functions(STATES, PARAMETERS)


@rhs_cfunc
def rhs_numba(t, states, values, parameters):
    g_Na_bar = parameters[0]
    g_K_bar = parameters[1]
    g_leak_Na = parameters[2]
    g_leak_K = parameters[3]
    E_Na = parameters[4]
    E_K = parameters[5]
    Cm = parameters[6]
    stim_amplitude = parameters[7]
    K_e = parameters[11]
    Na_i = parameters[12]
    m_K = parameters[13]
    m_Na = parameters[14]
    I_max = parameters[15]

    m = states[0]
    h = states[1]
    n = states[2]
    u = states[3] + 65.0

    alpha_m = 0.1*(25. - u)/(math.exp((25. - u)/10.) - 1)
    beta_m = 4.*math.exp(-u/18.)
    values[0] = (1 - m)*alpha_m - m*beta_m

    alpha_h = 0.07*math.exp(-u/20.)
    beta_h = 1./(math.exp((30. - u)/10.) + 1)
    values[1] = (1 - h)*alpha_h - h*beta_h

    alpha_n = 0.01*(10. - u)/(math.exp((10. - u)/10.) - 1.)
    beta_n = 0.125*math.exp(-u/80.)
    values[2] = (1 - n)*alpha_n - n*beta_n

    i_Stim = stim_amplitude*np.exp(-np.mod(t, 0.03)/0.002)*(t < 125)

    i_pump = I_max/((1 + m_K/K_e)**2*(1 + m_Na/Na_i)**3)

    i_Na = (g_leak_Na + g_Na_bar*h*math.pow(m, 3) + i_Stim)*(states[3] - E_Na) + 3*i_pump
    i_K = (g_leak_K + g_K_bar*math.pow(n, 4))*(states[3] - E_K) - 2*i_pump

    parameters[8] = i_Na
    parameters[9] = i_K
    parameters[10] = 0.0

    values[3] = (-i_K - i_Na)/Cm
