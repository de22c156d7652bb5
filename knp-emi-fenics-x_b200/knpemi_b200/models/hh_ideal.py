"""Hodgkin-Huxley neuron membrane, SI units (V, s, S/m^2), with Na/K pump.

Builtin restatement of the model the reference ships as
examples/idealized_geometries/mm_hh.py (states :12-21, parameters :39-78,
right-hand side :139-227).  Same slot numbering, same defaults, same
floating-point operation order in the right-hand side.
"""
import math

import numpy as np

from ._protocol import rhs_cfunc, table_functions

STATES = (
    ("m", 0.016648440745822956),
    ("h", 0.8542015627820805),
    ("n", 0.1882020248041632),
    ("V", -0.07438609374462003),
)

PARAMETERS = (
    ("g_Na_bar", 1200.0), ("g_K_bar", 360.0),
    ("g_leak_Na", 1.0), ("g_leak_K", 4.0),
    ("m_K", 2.0), ("m_Na", 7.7), ("I_max", 0.449),
    ("Cm", 0.0), ("stim_amplitude", 0.0),
    ("K_e", 0.0), ("K_i", 0.0), ("Na_e", 0.0), ("Na_i", 0.0),
    ("Cl_e", 0.0), ("Cl_i", 0.0),
    ("I_ch_Na", 0.0), ("I_ch_K", 0.0), ("I_ch_Cl", 0.0),
    ("z_Na", 0.0), ("z_K", 0.0), ("z_Cl", 0.0), ("psi", 0.0),
)

(init_state_values, init_parameter_values,
 state_indices, parameter_indices) = table_functions(STATES, PARAMETERS)


@rhs_cfunc
def rhs_numba(t, states, values, parameters):
    g_Na_bar = parameters[0]
    g_K_bar = parameters[1]
    g_leak_Na = parameters[2]
    g_leak_K = parameters[3]
    m_K = parameters[4]
    m_Na = parameters[5]
    I_max = parameters[6]
    Cm = parameters[7]
    stim_amplitude = parameters[8]
    K_e = parameters[9]
    K_i = parameters[10]
    Na_e = parameters[11]
    Na_i = parameters[12]
    z_K = parameters[19]
    psi = parameters[21]

    m = states[0]
    h = states[1]
    n = states[2]
    V = states[3]

    # Nernst potentials (the sodium one is scaled by z_K, as in the reference :169)
    E_Na = 1/psi/z_K * math.log(Na_e/Na_i)
    E_K = 1/psi/z_K * math.log(K_e/K_i)

    # shifted potential in mV used by the 1952 rate functions
    u = 1.0e3*(V + 65.0e-3)

    alpha_m = 0.1e3*(25. - u)/(math.exp((25. - u)/10.) - 1)
    beta_m = 4.e3*math.exp(-u/18.)
    values[0] = (1 - m)*alpha_m - m*beta_m

    alpha_h = 0.07e3*math.exp(-u/20.)
    beta_h = 1.e3/(math.exp((30. - u)/10.) + 1)
    values[1] = (1 - h)*alpha_h - h*beta_h

    alpha_n = 0.01e3*(10. - u)/(math.exp((10. - u)/10.) - 1.)
    beta_n = 0.125e3*math.exp(-u/80.)
    values[2] = (1 - n)*alpha_n - n*beta_n

    # synaptic conductance: decaying pulse train, switched off after 125 ms
    i_Stim = stim_amplitude*np.exp(-np.mod(t, 0.03)/0.002)*(t < 125e-3)

    i_pump = I_max/((1 + m_K/K_e)**2*(1 + m_Na/Na_i)**3)

    i_Na = (g_leak_Na + g_Na_bar*h*math.pow(m, 3) + i_Stim)*(V - E_Na) + 3*i_pump
    i_K = (g_leak_K + g_K_bar*math.pow(n, 4))*(V - E_K) - 2*i_pump

    parameters[15] = i_Na
    parameters[16] = i_K
    parameters[17] = 0.0

    values[3] = (-i_K - i_Na)/Cm
