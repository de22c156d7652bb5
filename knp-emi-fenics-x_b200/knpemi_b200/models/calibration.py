"""Neuron + glia + shared ECS compartment model used to calibrate initial
conditions (mV, ms, mM): 14 states, concentrations are ODE states.

Builtin restatement of examples/calibrate_initial_conditions/mm_calibration.py
(states :10-39, parameters :59-89, right-hand side :151-298).  There is no
state called ``V`` (``V_n`` / ``V_g`` instead) and no output parameter slot.
"""
import math

import numpy as np

from ._protocol import rhs_cfunc, table_functions

STATES = (
    ("m", 0.01365600905697864),
    ("h", 0.8804834256821714),
    ("n", 0.17041625484928405),
    ("V_n", -75.93151471235473),
    ("V_g", -85.85765274084892),
    ("K_e", 3.092970607490389),
    ("K_n", 124.13988964240784),
    ("K_g", 99.3100014897692),
    ("Na_e", 144.60625137617149),
    ("Na_n", 12.850454639128186),
    ("Na_g", 15.775818906083778),
    ("Cl_e", 133.62525154406637),
    ("Cl_n", 5.0),
    ("Cl_g", 5.203660274163705),
)

PARAMETERS = (
    ("g_Na_bar", 120.0), ("g_K_bar", 36.0),
    ("g_leak_Na_n", 0.1), ("g_leak_K_n", 0.4),
    ("g_leak_Na_g", 0.1), ("g_leak_K_g", 1.696),
    ("Cm", 1.0), ("stim_amplitude", 0.0),
    ("m_K", 1.5), ("m_Na", 10.0),
    ("I_max_n", 58.0), ("I_max_g", 10.75975),
    ("g_leak_Cl_g", 0.05),
)

(init_state_values, init_parameter_values,
 state_indices, parameter_indices) = table_functions(STATES, PARAMETERS)


@rhs_cfunc
def rhs_numba(t, states, values, parameters):
    temperature = 307e3
    R = 8.315e3
    F = 96500e3

    ICS_vol = 3.42e-11/2.0
    ECS_vol = 7.08e-11
    surface = 2.29e-6

    K_e_init = 3.092970607490389
    K_g_init = 99.3100014897692

    m = states[0]
    h = states[1]
    n = states[2]
    V_n = states[3]
    V_g = states[4]
    K_e = states[5]
    K_n = states[6]
    K_g = states[7]
    Na_e = states[8]
    Na_n = states[9]
    Na_g = states[10]
    Cl_e = states[11]
    Cl_n = states[12]
    Cl_g = states[13]

    g_Na_bar = parameters[0]
    g_K_bar = parameters[1]
    g_leak_Na_n = parameters[2]
    g_leak_K_n = parameters[3]
    g_leak_Na_g = parameters[4]
    g_leak_K_g = parameters[5]
    Cm = parameters[6]
    stim_amplitude = parameters[7]
    m_K = parameters[8]
    m_Na = parameters[9]
    I_max_n = parameters[10]
    I_max_g = parameters[11]
    g_leak_Cl_g = parameters[12]

    E_Na_n = R*temperature/F*np.log(Na_e/Na_n)
    E_K_n = R*temperature/F*np.log(K_e/K_n)
    E_Na_g = R*temperature/F*np.log(Na_e/Na_g)
    E_K_g = R*temperature/F*np.log(K_e/K_g)
    E_Cl_g = -R*temperature/F*np.log(Cl_e/Cl_g)
    E_K_init = R*temperature/F*np.log(K_e_init/K_g_init)

    alpha_m = 0.1*(V_n + 40.0)/(1.0 - math.exp(-(V_n + 40.0)/10.0))
    beta_m = 4.0*math.exp(-(V_n + 65.0)/18.0)
    alpha_h = 0.07*math.exp(-(V_n + 65.0)/20.0)
    beta_h = 1.0/(1.0 + math.exp(-(V_n + 35.0)/10.0))
    alpha_n = 0.01*(V_n + 55.0)/(1.0 - math.exp(-(V_n + 55.0)/10.0))
    beta_n = 0.125*math.exp(-(V_n + 65)/80.0)

    values[0] = (1 - m)*alpha_m - m*beta_m
    values[1] = (1 - h)*alpha_h - h*beta_h
    values[2] = (1 - n)*alpha_n - n*beta_n

    i_Stim = stim_amplitude*np.exp(-np.mod(t, 20.0)/2.0)

    i_pump_n = I_max_n/((1 + m_K/K_e)**2*(1 + m_Na/Na_n)**3)
    i_pump_g = I_max_g*(K_e/(K_e + m_K))*(Na_g**(1.5)/(Na_g**(1.5) + m_Na**(1.5)))

    # inward-rectifying K conductance of the glial membrane
    dphi = V_g - E_K_g
    A = 1 + np.exp(18.4/42.4)
    B = 1 + np.exp(-(0.1186e3 + E_K_init)/0.0441e3)
    C = 1 + np.exp((dphi + 0.0185e3)/0.0425e3)
    D = 1 + np.exp(-(0.1186e3 + V_g)/0.0441e3)
    g_Kir = np.sqrt(K_e/K_e_init)*(A*B)/(C*D)
    I_Kir = g_leak_K_g*g_Kir*(V_g - E_K_g)

    i_Na_n = (g_leak_Na_n + g_Na_bar*h*math.pow(m, 3) + i_Stim)*(V_n - E_Na_n) + 3*i_pump_n
    i_K_n = (g_leak_K_n + g_K_bar*math.pow(n, 4))*(V_n - E_K_n) - 2*i_pump_n
    i_Cl_n = 0.0

    i_Na_g = g_leak_Na_g*(V_g - E_Na_g) + 3*i_pump_g
    i_K_g = I_Kir - 2*i_pump_g
    i_Cl_g = g_leak_Cl_g*(V_g - E_Cl_g)

    values[3] = (-i_K_n - i_Na_n - i_Cl_n)/Cm
    values[4] = (-i_K_g - i_Na_g - i_Cl_g)/Cm

    # concentrations: current * surface / (F * volume)
    values[5] = i_K_n*surface/(F*ECS_vol) + i_K_g*surface/(F*ECS_vol)
    values[6] = -i_K_n*surface/(F*ICS_vol)
    values[7] = -i_K_g*surface/(F*ICS_vol)
    values[8] = i_Na_n*surface/(F*ECS_vol) + i_Na_g*surface/(F*ECS_vol)
    values[9] = -i_Na_n*surface/(F*ICS_vol)
    values[10] = -i_Na_g*surface/(F*ICS_vol)
    values[11] = -i_Cl_n*surface/(F*ECS_vol) - i_Cl_g*surface/(F*ECS_vol)
    values[12] = i_Cl_n*surface/(F*ICS_vol)
    values[13] = i_Cl_g*surface/(F*ICS_vol)
