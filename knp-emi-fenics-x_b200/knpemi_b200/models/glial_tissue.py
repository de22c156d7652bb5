"""Passive astrocyte membrane with Kir4.1 and Na/K pump (mV, ms, mS/cm^2).

Builtin restatement of examples/local_astrocyte_depolarization/mm_glial.py
(state :11-16, parameters :36-70, right-hand side :133-205).  One state (V);
the Kir reference potential is built from the ``K_e_init`` / ``K_i_init``
parameter slots and R, T, F literals (:168-170,177).
"""
import math

import numpy as np

from ._protocol import rhs_cfunc, table_functions

STATES = (("V", -85.84503411546689),)

PARAMETERS = (
    ("g_leak_Cl", 0.05), ("g_leak_Na", 0.1), ("g_leak_K", 1.696),
    ("Cm", 0.0), ("stim_amplitude", 0.0),
    ("I_ch_Na", 0.0), ("I_ch_K", 0.0), ("I_ch_Cl", 0.0),
    ("m_K", 1.5), ("m_Na", 10.0), ("I_max", 10.75975),
    ("K_e_init", 3.092970607490389), ("K_i_init", 99.3100014897692),
    ("K_e", 0.0), ("K_i", 0.0), ("Na_e", 0.0), ("Na_i", 0.0),
    ("Cl_e", 0.0), ("Cl_i", 0.0),
    ("z_Na", 0.0), ("z_K", 0.0), ("z_Cl", 0.0), ("psi", 0.0),
)

(init_state_values, init_parameter_values,
 state_indices, parameter_indices) = table_functions(STATES, PARAMETERS)


@rhs_cfunc
def rhs_numba(t, states, values, parameters):
    g_leak_Cl = parameters[0]
    g_leak_Na = parameters[1]
    g_leak_K = parameters[2]
    Cm = parameters[3]
    m_K = parameters[8]
    m_Na = parameters[9]
    I_max = parameters[10]
    K_e_init = parameters[11]
    K_i_init = parameters[12]
    K_e = parameters[13]
    K_i = parameters[14]
    Na_e = parameters[15]
    Na_i = parameters[16]
    Cl_e = parameters[17]
    Cl_i = parameters[18]
    z_K = parameters[20]
    z_Cl = parameters[21]
    psi = parameters[22]

    V = states[0]

    E_Na = 1/psi/z_K * math.log(Na_e/Na_i)
    E_K = 1/psi/z_K * math.log(K_e/K_i)
    E_Cl = 1/psi/z_Cl * math.log(Cl_e/Cl_i)

    temperature = 307e3
    R = 8.315e3
    F = 96500e3

    i_pump = I_max*(K_e/(K_e + m_K))*(Na_i**(1.5)/(Na_i**(1.5) + m_Na**(1.5)))

    # inward-rectifying K conductance
    E_K_init = R*temperature/F*np.log(K_e_init/K_i_init)
    dphi = V - E_K
    A = 1 + np.exp(18.5/42.4)
    B = 1 + np.exp(-(118.6 + E_K_init)/44.1)
    C = 1 + np.exp((dphi + 18.5)/42.4)
    D = 1 + np.exp(-(118.6 + V)/44.1)
    g_Kir = np.sqrt(K_e/K_e_init)*(A*B)/(C*D)

    i_Kir = g_leak_K*g_Kir*(V - E_K)
    i_Na = g_leak_Na*(V - E_Na) + 3*i_pump
    i_K = i_Kir - 2*i_pump
    i_Cl = g_leak_Cl*(V - E_Cl)

    parameters[5] = i_Na
    parameters[6] = i_K
    parameters[7] = i_Cl

    values[0] = (-i_K - i_Na - i_Cl)/Cm
