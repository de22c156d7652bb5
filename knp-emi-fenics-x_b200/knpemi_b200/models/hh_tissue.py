"""Hodgkin-Huxley neuron membrane, tissue units (mV, ms, mS/cm^2).

Builtin restatement of examples/local_astrocyte_depolarization/mm_hh.py
(states :11-20, parameters :40-67, right-hand side :130-201): same 22-slot
parameter layout as ``hh_ideal``, rate functions in the
``(V+40)/(1-exp(-(V+40)/10))`` form.
"""
import math

import numpy as np

from ._protocol import rhs_cfunc, table_functions

STATES = (
    ("m", 0.015211986965658385),
    ("h", 0.8667432624969533),
    ("n", 0.17994146133363148),
    ("V", -75.09159534786934),
)

PARAMETERS = (
    ("g_Na_bar", 120.0), ("g_K_bar", 36.0),
    ("g_leak_Na", 0.1), ("g_leak_K", 0.4),
    ("m_K", 1.5), ("m_Na", 10.0), ("I_max", 58.0),
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

    E_Na = 1/psi/z_K * math.log(Na_e/Na_i)
    E_K = 1/psi/z_K * math.log(K_e/K_i)

    alpha_m = 0.1*(V + 40.0)/(1.0 - math.exp(-(V + 40.0)/10.0))
    beta_m = 4.0*math.exp(-(V + 65.0)/18.0)
    alpha_h = 0.07*math.exp(-(V + 65.0)/20.0)
    beta_h = 1.0/(1.0 + math.exp(-(V + 35.0)/10.0))
    alpha_n = 0.01*(V + 55.0)/(1.0 - math.exp(-(V + 55.0)/10.0))
    beta_n = 0.125*math.exp(-(V + 65)/80.0)

    values[0] = (1 - m)*alpha_m - m*beta_m
    values[1] = (1 - h)*alpha_h - h*beta_h
    values[2] = (1 - n)*alpha_n - n*beta_n

    i_Stim = stim_amplitude*np.exp(-np.mod(t, 30.0)/2.0)*(t < 125)

    i_pump = I_max/((1 + m_K/K_e)**2*(1 + m_Na/Na_i)**3)

    i_Na = (g_leak_Na + g_Na_bar*h*math.pow(m, 3) + i_Stim)*(V - E_Na) + 3*i_pump
    i_K = (g_leak_K + g_K_bar*math.pow(n, 4))*(V - E_K) - 2*i_pump

    parameters[15] = i_Na
    parameters[16] = i_K
    parameters[17] = 0.0

    values[3] = (-i_K - i_Na)/Cm
